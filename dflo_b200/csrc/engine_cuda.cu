// The product: CUDA backend of Engine<> for sm_100a and the extern "C" entry points of
// include/dflo_b200.h.  One CUDA stream per ctx; whole time steps are captured once into a CUDA
// graph (one per starting solution buffer) and replayed, so a step costs one graph launch and no
// host round trip; the halo exchange of a sharded ctx is one NCCL group of send/recv pairs per
// RK stage on the same stream.  There is no CPU fallback: without a CUDA device create() returns
// DFLO_E_NO_DEVICE.
#include "abi_impl.h"
#include "row_kernel.cuh"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h> // types only: the library is bound at run time, see NcclApi

#include <cstdio>
#include <cstdlib>
#include <map>

namespace
{
   // NCCL is bound with dlopen when the first sharded ctx is created, not at link time: a host
   // process that already carries an NCCL (torch.distributed ships its own libnccl.so.2, newer
   // than the system one) must keep exactly one copy, and a single-GPU host needs none at all.
   struct NcclApi
   {
      void *handle = nullptr;
      decltype (&ncclGetUniqueId) GetUniqueId = nullptr;
      decltype (&ncclCommInitRank) CommInitRank = nullptr;
      decltype (&ncclCommDestroy) CommDestroy = nullptr;
      decltype (&ncclGetErrorString) GetErrorString = nullptr;
      decltype (&ncclGroupStart) GroupStart = nullptr;
      decltype (&ncclGroupEnd) GroupEnd = nullptr;
      decltype (&ncclSend) Send = nullptr;
      decltype (&ncclRecv) Recv = nullptr;
      decltype (&ncclAllReduce) AllReduce = nullptr;
      std::string error;

      bool load ()
      {
         if (handle) return true;
         const char *env = std::getenv ("DFLO_B200_NCCL_LIB");
         if (env) handle = dlopen (env, RTLD_NOW | RTLD_GLOBAL);
         if (!handle) handle = dlopen ("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD); // the copy the host process already has
         if (!handle) handle = dlopen ("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
         if (!handle) handle = dlopen ("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
         if (!handle)
         {
            error = std::string ("cannot load libnccl.so.2: ") + dlerror ();
            return false;
         }
#define DFLO_NCCL_SYM(name)                                                          \
   name = reinterpret_cast<decltype (name)> (dlsym (handle, "nccl" #name));           \
   if (!name)                                                                        \
   {                                                                                 \
      error = "libnccl lacks nccl" #name;                                            \
      handle = nullptr;                                                              \
      return false;                                                                  \
   }
         DFLO_NCCL_SYM (GetUniqueId)
         DFLO_NCCL_SYM (CommInitRank)
         DFLO_NCCL_SYM (CommDestroy)
         DFLO_NCCL_SYM (GetErrorString)
         DFLO_NCCL_SYM (GroupStart)
         DFLO_NCCL_SYM (GroupEnd)
         DFLO_NCCL_SYM (Send)
         DFLO_NCCL_SYM (Recv)
         DFLO_NCCL_SYM (AllReduce)
#undef DFLO_NCCL_SYM
         return true;
      }
   };
   NcclApi &nccl ()
   {
      static NcclApi api;
      return api;
   }

   template <class K>
   __global__ void __launch_bounds__ (K::THREADS, K::MIN_BLOCKS) phase_kernel (const typename K::Args a)
   {
      extern __shared__ __align__ (16) double dflo_smem[];
#pragma unroll
      for (int p = 0; p < K::NPHASE; ++p)
      {
         K::phase (p, a, dflo_smem, threadIdx.x, blockIdx.x);
         if (p + 1 < K::NPHASE) __syncthreads ();
      }
   }

   template <class K>
   __global__ void __launch_bounds__ (256) thread_kernel (const typename K::Args a, int n)
   {
      const int j = blockIdx.x * 256 + threadIdx.x;
      if (j < n) K::thread (a, j);
   }

   struct CudaBackend
   {
      cudaStream_t stream = nullptr;
      cudaEvent_t ev0 = nullptr, ev1 = nullptr;
      int device = 0, rank = 0, world = 1;
      int64_t launches = 0;
      cudaError_t first_error = cudaSuccess;
      ncclResult_t first_nccl_error = ncclSuccess;
      ncclComm_t comm = nullptr;
      bool use_graphs = true;
      bool persistent = true;
      bool row_kernel = true;  // register-blocked stage kernel for Qk (row_kernel.cuh); DFLO_B200_STAGE=tile selects the phase kernel
      int n_sm = 148;
      bool capturing = false;
      struct Graph
      {
         cudaGraphExec_t exec;
         int cur_after;
         int64_t launches_per_replay;
      };
      std::map<int, Graph> graphs;
      int64_t launches_at_capture = 0;

      void note (cudaError_t e)
      {
         if (e != cudaSuccess && first_error == cudaSuccess) first_error = e;
      }
      void note (ncclResult_t e)
      {
         if (e != ncclSuccess && first_nccl_error == ncclSuccess) first_nccl_error = e;
      }

      int open (int dev, int r, int w, const void *nccl_id, std::string &err)
      {
         int count = 0;
         if (cudaGetDeviceCount (&count) != cudaSuccess || count == 0)
         {
            (void) cudaGetLastError ();
            err = "no CUDA device visible; dflo_b200 has no CPU fallback";
            return DFLO_E_NO_DEVICE;
         }
         if (dev < 0 || dev >= count)
         {
            err = "device index out of range";
            return DFLO_E_INVALID;
         }
         device = dev;
         rank = r;
         world = w;
         note (cudaSetDevice (device));
         note (cudaStreamCreateWithFlags (&stream, cudaStreamNonBlocking));
         note (cudaEventCreate (&ev0));
         note (cudaEventCreate (&ev1));
         note (cudaDeviceGetAttribute (&n_sm, cudaDevAttrMultiProcessorCount, device));
         const char *ps = std::getenv ("DFLO_B200_PERSISTENT");
         persistent = ps ? (std::atoi (ps) != 0) : true;
         const char *sk = std::getenv ("DFLO_B200_STAGE");
         row_kernel = !(sk && std::string (sk) == "tile");
         const char *g = std::getenv ("DFLO_B200_GRAPHS");
         use_graphs = g ? (std::atoi (g) != 0) : true;
         if (world > 1)
         {
            if (!nccl_id)
            {
               err = "sharded context needs an NCCL unique id";
               return DFLO_E_INVALID;
            }
            if (!nccl ().load ())
            {
               err = nccl ().error;
               return DFLO_E_NCCL;
            }
            ncclUniqueId id;
            std::memcpy (&id, nccl_id, sizeof (id));
            const ncclResult_t rc = nccl ().CommInitRank (&comm, world, id, rank);
            if (rc != ncclSuccess)
            {
               err = std::string ("ncclCommInitRank: ") + nccl ().GetErrorString (rc);
               return DFLO_E_NCCL;
            }
         }
         return check (err);
      }

      void close ()
      {
         if (comm) nccl ().CommDestroy (comm);
         comm = nullptr;
         if (ev0) cudaEventDestroy (ev0);
         if (ev1) cudaEventDestroy (ev1);
         if (stream) cudaStreamDestroy (stream);
         stream = nullptr;
         ev0 = ev1 = nullptr;
      }

      template <class T> T *alloc (size_t n)
      {
         void *p = nullptr;
         note (cudaMalloc (&p, (n ? n : 1) * sizeof (T)));
         return static_cast<T *> (p);
      }
      void free (void *p)
      {
         if (p) note (cudaFree (p));
      }
      void h2d (void *d, const void *h, size_t b)
      {
         if (b) note (cudaMemcpyAsync (d, h, b, cudaMemcpyHostToDevice, stream));
      }
      void d2h (void *h, const void *d, size_t b)
      {
         if (!b) return;
         note (cudaMemcpyAsync (h, d, b, cudaMemcpyDeviceToHost, stream));
         note (cudaStreamSynchronize (stream));
      }
      void zero (void *d, size_t b) { note (cudaMemsetAsync (d, 0, b, stream)); }
      void sync () { note (cudaStreamSynchronize (stream)); }
      void *stream_handle () const { return (void *) stream; }

      int check (std::string &err)
      {
         note (cudaPeekAtLastError ());
         if (first_error != cudaSuccess)
         {
            err = std::string ("CUDA: ") + cudaGetErrorString (first_error);
            return DFLO_E_CUDA;
         }
         if (first_nccl_error != ncclSuccess)
         {
            err = std::string ("NCCL: ") + nccl ().GetErrorString (first_nccl_error);
            return DFLO_E_NCCL;
         }
         return DFLO_OK;
      }

      template <class K> void launch (int grid, const typename K::Args &a)
      {
         if (grid <= 0) return;
         constexpr size_t smem = K::SMEM_DOUBLES * sizeof (double);
         if (smem > 48 * 1024) // opt in to large dynamic shared memory once per kernel
         {
            static const cudaError_t rc = cudaFuncSetAttribute (phase_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
            note (rc);
         }
         ++launches;
         phase_kernel<K><<<grid, K::THREADS, smem, stream>>> (a);
         note (cudaPeekAtLastError ());
      }
      // The stage kernel.  Default: the pipelined persistent form (one producer warp streaming tiles
      // through two shared-memory stages); DFLO_B200_PERSISTENT=0 selects the one-tile-per-block form.
      int stage_prefetch_tiles () const
      {
         static const char *e = std::getenv ("DFLO_B200_PF_TILES");
         return e ? std::atoi (e) : n_sm * DFLO_ROW_MIN_BLOCKS;
      }
      int debug_flags () const
      {
         static const char *e = std::getenv ("DFLO_B200_DBG");
         return e ? std::atoi (e) : 0;
      }
      bool use_row_kernel (int basis, int n1) const { return row_kernel && basis == dflo::BASIS_QK && n1 >= 2; }
      // 1-D tables of the row kernel as constant-bank operands
      void prepare_tables (const dflo::FeTables &t)
      {
         if (t.basis != dflo::BASIS_QK) return;
         dflo::RowConst rc;
         std::memset (&rc, 0, sizeof (rc));
         for (int i = 0; i < t.n1; ++i)
         {
            for (int j = 0; j < t.n1; ++j) rc.dw[i][j] = t.dw[i][j];
            rc.e0[i] = t.e[0][i];
            rc.e1[i] = t.e[1][i];
            rc.gw[i] = t.gw[i];
         }
         note (cudaMemcpyToSymbolAsync (dflo::c_row, &rc, sizeof (rc), (size_t) t.n1 * sizeof (rc), cudaMemcpyHostToDevice, stream));
         note (cudaStreamSynchronize (stream));
      }
      template <int N1, int FLUX> void launch_row (int n_tiles, const dflo::StageArgs &a)
      {
         typedef dflo::RowShape<N1, FLUX> S;
         constexpr size_t smem = S::SMEM_DOUBLES * sizeof (double);
         static const cudaError_t rc = cudaFuncSetAttribute (dflo::row_stage_kernel<N1, FLUX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
         note (rc);
         ++launches;
         dflo::row_stage_kernel<N1, FLUX><<<n_tiles, S::THREADS, smem, stream>>> (a);
         note (cudaPeekAtLastError ());
      }
      template <class K> void launch_stage (int n_tiles, const typename K::Args &a)
      {
         if (n_tiles <= 0) return;
         if constexpr (K::BASIS_ID == dflo::BASIS_QK && K::N1_ID >= 2)
         {
            if (row_kernel)
            {
               launch_row<K::N1_ID, K::FLUX_ID> (n_tiles, a);
               return;
            }
         }
         if (!persistent)
         {
            launch<K> (n_tiles, a);
            return;
         }
         constexpr size_t smem = K::PERSIST_SMEM_DOUBLES * sizeof (double);
         static int blocks_per_sm = -1;
         if (blocks_per_sm < 0)
         {
            note (cudaFuncSetAttribute (dflo::stage_persistent_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
            int n = 0;
            note (cudaOccupancyMaxActiveBlocksPerMultiprocessor (&n, dflo::stage_persistent_kernel<K>, K::THREADS + 32, smem));
            blocks_per_sm = n > 0 ? n : 1;
         }
         const int grid = std::min (n_tiles, n_sm * blocks_per_sm);
         ++launches;
         dflo::stage_persistent_kernel<K><<<grid, K::THREADS + 32, smem, stream>>> (a, n_tiles);
         note (cudaPeekAtLastError ());
      }
      template <class K> void launch1d (int n, const typename K::Args &a)
      {
         if (n <= 0) return;
         ++launches;
         thread_kernel<K><<<(n + 255) / 256, 256, 0, stream>>> (a, n);
         note (cudaPeekAtLastError ());
      }

      // ---- CUDA graphs: one instantiated graph of a whole time step per starting buffer ----
      void drop_graphs ()
      {
         for (auto &kv : graphs) cudaGraphExecDestroy (kv.second.exec);
         graphs.clear ();
      }
      bool graph_launch (int key, int *cur_after)
      {
         auto it = graphs.find (key);
         if (it == graphs.end ()) return false;
         note (cudaGraphLaunch (it->second.exec, stream));
         launches += it->second.launches_per_replay;
         *cur_after = it->second.cur_after;
         return true;
      }
      bool capture_begin ()
      {
         if (!use_graphs || world > 1) return false; // sharded steps run eagerly (NCCL on the stream)
         if (cudaStreamBeginCapture (stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess)
         {
            (void) cudaGetLastError ();
            return false;
         }
         capturing = true;
         launches_at_capture = launches;
         return true;
      }
      void capture_end_and_launch (int key, int cur_after)
      {
         cudaGraph_t g = nullptr;
         capturing = false;
         note (cudaStreamEndCapture (stream, &g));
         if (!g) return;
         Graph entry;
         entry.cur_after = cur_after;
         entry.launches_per_replay = launches - launches_at_capture;
         const cudaError_t rc = cudaGraphInstantiate (&entry.exec, g, 0);
         cudaGraphDestroy (g);
         note (rc);
         if (rc != cudaSuccess) return;
         graphs[key] = entry;
         note (cudaGraphLaunch (entry.exec, stream)); // capture recorded the step, now run it
      }

      void timer_start () { note (cudaEventRecord (ev0, stream)); }
      void timer_stop () { note (cudaEventRecord (ev1, stream)); }
      float timer_ms ()
      {
         float ms = 0.0f;
         note (cudaEventSynchronize (ev1));
         note (cudaEventElapsedTime (&ms, ev0, ev1));
         return ms;
      }

      // ---- NCCL: halo as one group of send/recv pairs, dt/residual as tiny all-reduces ----
      void halo_begin ()
      {
         if (comm) note (nccl ().GroupStart ());
      }
      void halo_send (int peer, const double *buf, size_t count)
      {
         if (comm) note (nccl ().Send (buf, count, ncclDouble, peer, comm, stream));
      }
      void halo_recv (int peer, double *buf, size_t count)
      {
         if (comm) note (nccl ().Recv (buf, count, ncclDouble, peer, comm, stream));
      }
      void halo_end ()
      {
         if (comm) note (nccl ().GroupEnd ());
      }
      void halo_wait () {}
      void allreduce_min_dt (double *p)
      {
         if (comm) note (nccl ().AllReduce (p, p, 1, ncclDouble, ncclMin, comm, stream));
      }
      void allreduce_sum (double *p, int n)
      {
         if (comm) note (nccl ().AllReduce (p, p, n, ncclDouble, ncclSum, comm, stream));
      }
   };
}

DFLO_DEFINE_ABI (dflo_b200_, CudaBackend, dflo_ctx)

extern "C" int dflo_b200_nccl_unique_id (void *out128)
{
   if (!out128) return DFLO_E_INVALID;
   ncclUniqueId id;
   if (!nccl ().load () || nccl ().GetUniqueId (&id) != ncclSuccess) return DFLO_E_NCCL;
   static_assert (sizeof (id) == 128, "ncclUniqueId is 128 bytes");
   std::memcpy (out128, &id, sizeof (id));
   return DFLO_OK;
}
