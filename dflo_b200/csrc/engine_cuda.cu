// The product: CUDA backend of Engine<> for sm_100a and the extern "C" entry points of
// include/dflo_b200.h.  One CUDA stream per ctx; whole time steps are captured once into a CUDA
// graph (one per starting solution buffer) and replayed, so a step costs one graph launch and no
// host round trip; the stage kernels of a step are programmatic dependent launches.  The halo exchange
// of a sharded ctx goes over peer memory (p2p_halo.cuh: stores into the peers' ghost ranges over NVLink,
// fused into the row stage kernel where no limiter follows); one NCCL group of send/recv pairs per RK
// stage is the fallback when the peers' buffers cannot be mapped.  There is no CPU fallback: without a
// CUDA device create() returns DFLO_E_NO_DEVICE.
#include "abi_impl.h"
#include "p2p_halo.cuh"
#include "row_kernel.cuh"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h> // types only: the library is bound at run time, see NcclApi

#include <cstdio>
#include <cstdlib>
#include <map>
#include <typeinfo>
#include <vector>

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
      decltype (&ncclAllGather) AllGather = nullptr;
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
         DFLO_NCCL_SYM (AllGather)
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
      // Programmatic dependent launch (no-ops in an ordinary launch): wait for the predecessor on the stream, THEN let the
      // successor be scheduled -- a kernel that is allowed to read its inputs before its own wait (the first row stage
      // kernel of a step) may rely on everything before its predecessor being complete.
      dflo::pdl_wait ();
      dflo::pdl_launch_dependents ();
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
      dflo::pdl_wait (); // see phase_kernel
      dflo::pdl_launch_dependents ();
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
      // peer-memory halo (p2p_halo.cuh)
      bool p2p = false;
      dflo::P2PArgs p2p_args;
      double *p2p_U[3] = {nullptr, nullptr, nullptr}, *p2p_A[3] = {nullptr, nullptr, nullptr};
      double *p2p_peerU[dflo::P2P_MAX_WORLD][3], *p2p_peerA[dflo::P2P_MAX_WORLD][3];
      std::vector<void *> p2p_opened;
      unsigned long long *p2p_flags = nullptr, *p2p_epochs = nullptr;
      unsigned int *p2p_counter = nullptr;
      int p2p_grid = 1;
      bool p2p_fused = false;
      dflo::P2PFused *p2p_fused_dev = nullptr;
      unsigned int *p2p_send_counters = nullptr;
      bool p2p_fused_ok () const { return p2p && p2p_fused; }
      const dflo::P2PFused *p2p_fused_args (int buf) const { return p2p_fused_dev + buf; }
      // stand-alone exchange whose wait is left to the ghost-reading tiles of the next row stage kernel: an experiment
      // (DFLO_B200_P2P_DEFER=1), off by default -- bit for bit and hang-free at 2 and 8 GPUs but no faster (cfg4 on 8 GPUs:
      // 307 k against 313 k MDoF/s): what the strong-scaling runs lose is not the wait
      static bool p2p_defer_requested ()
      {
         static const char *e = std::getenv ("DFLO_B200_P2P_DEFER");
         return e && std::atoi (e) != 0;
      }
      bool p2p_deferred_ok () const { return p2p && p2p_defer_requested (); }
      const dflo::P2PFused *p2p_wait_args () const { return p2p_fused_dev + 3; } // no senders: the row kernel only waits

      // developer timeline (DFLO_B200_KTRACE=1): every launch bracketed by events on the ctx stream, steps run eagerly;
      // per-kernel averages are printed when the ctx closes.  Off by default: nothing is recorded.
      bool ktrace = false;
      struct KSpan
      {
         const char *name;
         cudaEvent_t a, b;
      };
      std::vector<KSpan> kspans;
      std::map<std::string, std::pair<double, long>> ktotals;
      void k_begin (const char *name)
      {
         if (!ktrace) return;
         KSpan sp;
         sp.name = name;
         cudaEventCreate (&sp.a);
         cudaEventCreate (&sp.b);
         cudaEventRecord (sp.a, stream);
         kspans.push_back (sp);
      }
      void k_end ()
      {
         if (ktrace) cudaEventRecord (kspans.back ().b, stream);
      }
      void k_collect ()
      {
         if (!ktrace) return;
         cudaStreamSynchronize (stream);
         for (auto &sp : kspans)
         {
            float ms = 0.f;
            cudaEventElapsedTime (&ms, sp.a, sp.b);
            auto &t = ktotals[sp.name];
            t.first += ms;
            t.second += 1;
            cudaEventDestroy (sp.a);
            cudaEventDestroy (sp.b);
         }
         kspans.clear ();
      }
      void k_report ()
      {
         if (!ktrace) return;
         k_collect ();
         for (auto &kv : ktotals)
            std::fprintf (stderr, "[ktrace rank %d] %-28s n %6ld  avg %8.2f us\n", rank, kv.first.c_str (), kv.second.second, 1e3 * kv.second.first / kv.second.second);
      }

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
         const char *kt = std::getenv ("DFLO_B200_KTRACE");
         ktrace = kt && std::atoi (kt) != 0;
         if (ktrace) use_graphs = false;
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
         k_report ();
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
      void sync ()
      {
         note (cudaStreamSynchronize (stream));
         if (ktrace && !kspans.empty () && !capturing) k_collect ();
      }
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

      // launch configuration with the programmatic-serialization attribute when every kernel of a step is chained that way
      // (DFLO_B200_PDL >= 3): the next kernel's blocks are scheduled while the last blocks of this one drain
      template <class... Params, class... Args>
      void launch_cfg (void (*kernel) (Params...), dim3 grid, dim3 block, size_t smem, Args... args)
      {
         cudaLaunchConfig_t cfg = {};
         cfg.gridDim = grid;
         cfg.blockDim = block;
         cfg.dynamicSmemBytes = smem;
         cfg.stream = stream;
         cudaLaunchAttribute at[1];
         at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
         at[0].val.programmaticStreamSerializationAllowed = 1;
         cfg.attrs = at;
         cfg.numAttrs = pdl_level () >= 3 ? 1 : 0;
         note (cudaLaunchKernelEx (&cfg, kernel, args...));
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
         k_begin (typeid (K).name ());
         launch_cfg (phase_kernel<K>, dim3 (grid), dim3 (K::THREADS), smem, a);
         k_end ();
         note (cudaPeekAtLastError ());
      }
      // Stage-kernel forms: Qk runs the register-blocked row kernel (launch_row; DFLO_B200_STAGE=tile selects the generic
      // tile kernel instead), Pk degree 1-2 the thread-per-cell kernel (cell_stage.cuh), mapping = q1 the mapped kernel.
      // What is left for the generic tile kernel (P3, degree 0) runs in its pipelined persistent form (one producer
      // warp streaming tiles through two shared-memory stages) unless DFLO_B200_PERSISTENT=0.
      int n_sms () const { return n_sm; }
      int stage_prefetch_tiles () const
      {
         static const char *e = std::getenv ("DFLO_B200_PF_TILES");
         return e ? std::atoi (e) : n_sm * DFLO_ROW_MIN_BLOCKS;
      }
      bool limiter_block_form () const
      {
         static const char *e = std::getenv ("DFLO_B200_LIMITER");
         return e && std::string (e) == "block";
      }
      int debug_flags () const
      {
         static const char *e = std::getenv ("DFLO_B200_DBG");
         return e ? std::atoi (e) : 0;
      }
      // programmatic dependent launch (DFLO_B200_PDL: 0 off, 1 first stage kernel beside the time-step kernels, 2 also
      // stage after stage, 3 every kernel of the ctx chained that way); off while tracing (the event pairs serialise anyway)
      int pdl_level () const
      {
         static const char *e = std::getenv ("DFLO_B200_PDL");
         return ktrace ? 0 : e ? std::atoi (e) : 2; // 3 measured no faster (cfg3 6 % slower): profiles/r02_pdl_levels.md
      }
      bool keep_graphs_sharded () const
      {
         static const char *e = std::getenv ("DFLO_B200_KEEP_GRAPHS");
         return p2p && !(e && std::atoi (e) == 0); // with NCCL on the stream the steps are not captured to begin with
      }
      bool use_row_kernel (int basis, int n1) const { return row_kernel && basis == dflo::BASIS_QK && n1 >= 2; }
      // 1-D tables of the row kernel as constant-bank operands
      void prepare_tables (const dflo::FeTables &t, const std::vector<double> &flat)
      {
         if (t.basis == dflo::BASIS_PK && t.n1 >= 2 && t.n1 <= 3) // thread-per-cell Pk stage kernel: the flat stage table
         {
            note (cudaMemcpyToSymbolAsync (dflo::c_pk_tab, flat.data (), flat.size () * sizeof (double),
                                           (size_t) t.n1 * dflo::PK_TAB_MAX * sizeof (double), cudaMemcpyHostToDevice, stream));
            note (cudaStreamSynchronize (stream));
         }
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
         k_begin (a.mode == dflo::MODE_RHS ? "row_stage(rhs)" : a.ark == 0.0 ? "row_stage(rk0)" : "row_stage(rk>0)");
         if (a.pdl)
         {
            // programmatic dependent launch: blocks may be scheduled while the predecessor's last blocks still run
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3 (n_tiles);
            cfg.blockDim = dim3 (S::THREADS);
            cfg.dynamicSmemBytes = smem;
            cfg.stream = stream;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            note (cudaLaunchKernelEx (&cfg, dflo::row_stage_kernel<N1, FLUX>, a));
         }
         else
            dflo::row_stage_kernel<N1, FLUX><<<n_tiles, S::THREADS, smem, stream>>> (a);
         k_end ();
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
      // thread-per-cell Pk stage kernel (cell_stage.cuh); DFLO_B200_PK=tile selects the tile kernel
      bool use_pk_cell_kernel () const
      {
         static const char *e = std::getenv ("DFLO_B200_PK");
         return !(e && std::string (e) == "tile");
      }
      void note_cell_stage () {}
      template <class K> void launch1d (int n, const typename K::Args &a)
      {
         if (n <= 0) return;
         ++launches;
         k_begin (typeid (K).name ());
         launch_cfg (thread_kernel<K>, dim3 ((n + 255) / 256), dim3 (256), 0, a, n);
         k_end ();
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
         static const char *gs = std::getenv ("DFLO_B200_GRAPHS_SHARDED");
         // sharded steps are captured when the halo goes over peer memory; with NCCL on the stream they run eagerly unless asked
         if (!use_graphs || (world > 1 && !p2p && !(gs && std::atoi (gs)))) return false;
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
      // returns true if the step was also finalised (dt clipped and published, accumulator reset)
      bool allreduce_min_dt (double *p, bool finalize = false, double time_step = 0.0)
      {
         if (p2p)
         {
            dflo::P2PArgs a = p2p_args;
            a.dt_val = p;
            ++launches;
            k_begin ("dt_min_kernel");
            if (pdl_level () >= 1)
            {
               cudaLaunchConfig_t cfg = {};
               cfg.gridDim = dim3 (1);
               cfg.blockDim = dim3 (32);
               cfg.stream = stream;
               cudaLaunchAttribute at[1];
               at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
               at[0].val.programmaticStreamSerializationAllowed = 1;
               cfg.attrs = at;
               cfg.numAttrs = 1;
               note (cudaLaunchKernelEx (&cfg, dflo::dt_min_kernel, a, finalize ? 1 : 0, time_step));
            }
            else
               dflo::dt_min_kernel<<<1, 32, 0, stream>>> (a, finalize ? 1 : 0, time_step);
            k_end ();
            note (cudaPeekAtLastError ());
            return finalize;
         }
         else if (comm)
            note (nccl ().AllReduce (p, p, 1, ncclDouble, ncclMin, comm, stream));
         return false;
      }

      // ---- peer-memory halo: map the peers' buffers once (CUDA IPC handles over the NCCL communicator) ----
      template <class Peers>
      void p2p_setup (double **U, double **AVG, const Peers &peers, const std::vector<int *> *d_send_cells, int D, const int *d_send_entries,
                      int n_send_tiles)
      {
         const char *env = std::getenv ("DFLO_B200_P2P");
         if (!comm || world < 2 || world > dflo::P2P_MAX_WORLD || (env && std::atoi (env) == 0)) return;
         const int W = world, NH = 7; // handles per rank: U0 U1 U2 A0 A1 A2 flags
         p2p_flags = alloc<unsigned long long> (4 * W);
         p2p_epochs = alloc<unsigned long long> (4);
         p2p_counter = alloc<unsigned int> (1);
         zero (p2p_flags, 4 * W * sizeof (unsigned long long));
         {
            // the dt slots (words 2W .. 4W) are armed with -1: a non-negative value is a message
            std::vector<double> arm (2 * W, -1.0);
            h2d (p2p_flags + 2 * W, arm.data (), arm.size () * sizeof (double));
            sync ();
         }
         zero (p2p_counter, sizeof (unsigned int));
         const unsigned long long ones[4] = {1, 1, 1, 0};
         h2d (p2p_epochs, ones, sizeof (ones));
         std::vector<cudaIpcMemHandle_t> hs ((size_t) W * NH);
         int ok = 1;
         void *mine[NH] = {U[0], U[1], U[2], AVG[0], AVG[1], AVG[2], p2p_flags};
         for (int i = 0; i < NH; ++i)
            if (cudaIpcGetMemHandle (&hs[(size_t) rank * NH + i], mine[i]) != cudaSuccess) ok = 0;
         (void) cudaGetLastError ();
         char *d_h = alloc<char> (hs.size () * sizeof (cudaIpcMemHandle_t));
         const size_t per = NH * sizeof (cudaIpcMemHandle_t);
         h2d (d_h + rank * per, &hs[(size_t) rank * NH], per);
         note (nccl ().AllGather (d_h + rank * per, d_h, per, ncclChar, comm, stream));
         d2h (hs.data (), d_h, hs.size () * sizeof (cudaIpcMemHandle_t));
         free (d_h);
         std::memset (&p2p_args, 0, sizeof (p2p_args));
         auto open_handle = [&] (const cudaIpcMemHandle_t &h) -> void * {
            void *ptr = nullptr;
            if (cudaIpcOpenMemHandle (&ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess)
            {
               (void) cudaGetLastError ();
               ok = 0;
               return nullptr;
            }
            p2p_opened.push_back (ptr);
            return ptr;
         };
         for (int r = 0; r < W && ok; ++r)
            p2p_args.all_flags[r] = r == rank ? p2p_flags : static_cast<unsigned long long *> (open_handle (hs[(size_t) r * NH + 6]));
         int np = 0, ns = 0;
         size_t items = 0;
         for (size_t pi = 0; pi < peers.size () && ok; ++pi)
         {
            const int r = peers[pi].rank;
            for (int i = 0; i < 3 && ok; ++i)
            {
               p2p_peerU[np][i] = static_cast<double *> (open_handle (hs[(size_t) r * NH + i]));
               p2p_peerA[np][i] = static_cast<double *> (open_handle (hs[(size_t) r * NH + 3 + i]));
            }
            p2p_args.peer_flags[np] = p2p_args.all_flags[r];
            p2p_args.peer_rank[np] = r;
            for (int k = 0; k < 2; ++k)
            {
               const int n = (int) peers[pi].send_cells[k].size ();
               if (!n) continue;
               dflo::P2PSeg sg;
               sg.cells = d_send_cells[k][pi];
               sg.n = n;
               sg.peer = np;
               sg.dst_cell0 = peers[pi].dst_start[k];
               p2p_args.seg[ns++] = sg;
               items += (size_t) n * D / 2;
            }
            ++np;
         }
         // every rank must take the same path
         int *d_ok = alloc<int> (1);
         h2d (d_ok, &ok, sizeof (int));
         note (nccl ().AllReduce (d_ok, d_ok, 1, ncclInt, ncclMin, comm, stream));
         d2h (&ok, d_ok, sizeof (int));
         free (d_ok);
         if (!ok)
         {
            p2p_teardown ();
            return;
         }
         for (int i = 0; i < 3; ++i)
         {
            p2p_U[i] = U[i];
            p2p_A[i] = AVG[i];
         }
         p2p_args.my_flags = p2p_flags;
         p2p_args.epochs = p2p_epochs;
         p2p_args.counter = p2p_counter;
         p2p_args.nseg = ns;
         p2p_args.npeers = np;
         p2p_args.me = rank;
         p2p_args.world = W;
         p2p_args.D = D;
         if (std::getenv ("DFLO_B200_P2P_TRACE"))
         {
            p2p_args.trace = alloc<unsigned long long> (4 * 4096);
            zero (p2p_args.trace, 4 * 4096 * sizeof (unsigned long long));
         }
         p2p_grid = (int) std::min<size_t> (n_sm, std::max<size_t> (1, items / 512));
         // fused form: one descriptor per output buffer
         p2p_send_counters = alloc<unsigned int> (2);
         zero (p2p_send_counters, 2 * sizeof (unsigned int));
         p2p_fused_dev = alloc<dflo::P2PFused> (4);
         for (int b = 0; b < 4; ++b)
         {
            dflo::P2PFused f;
            std::memset (&f, 0, sizeof (f));
            for (int q = 0; q < np; ++q)
            {
               f.dstU[q] = p2p_peerU[q][b < 3 ? b : 0];
               f.dstA[q] = p2p_peerA[q][b < 3 ? b : 0];
               f.peer_flags[q] = p2p_args.peer_flags[q];
               f.peer_rank[q] = p2p_args.peer_rank[q];
            }
            f.my_flags = p2p_flags;
            f.epochs = p2p_epochs;
            f.send_counter = p2p_send_counters;
            f.block_counter = p2p_send_counters + 1;
            f.send_entries = d_send_entries;
            f.n_send_tiles = b < 3 ? n_send_tiles : 0; // [3]: wait-only
            f.npeers = np;
            f.me = rank;
            f.world = W;
            h2d (p2p_fused_dev + b, &f, sizeof (f));
         }
         sync ();
         const char *fe = std::getenv ("DFLO_B200_P2P_FUSED");
         p2p_fused = n_send_tiles > 0 && !(fe && std::atoi (fe) == 0);
         p2p = true;
      }
      void p2p_teardown ()
      {
         if (p2p && p2p_args.trace) // developer timeline: per exchange, ns relative to the previous exchange's end
         {
            std::vector<unsigned long long> t (4 * 4096);
            sync ();
            d2h (t.data (), p2p_args.trace, t.size () * sizeof (unsigned long long));
            unsigned long long e_last = 0;
            d2h (&e_last, p2p_epochs, sizeof (e_last));
            const int hi = (int) std::min<unsigned long long> (e_last, 4096);
            for (int e = std::max (2, hi - 12); e < hi; ++e)
               std::fprintf (stderr, "[p2p rank %d] exch %d: since prev end %7.2f us | push start->copied %6.2f | fence+flag %6.2f | wait peers %6.2f\n", rank, e,
                             (t[4 * e] - t[4 * (e - 1) + 3]) * 1e-3, (t[4 * e + 1] - t[4 * e]) * 1e-3, (t[4 * e + 2] - t[4 * e + 1]) * 1e-3,
                             (t[4 * e + 3] - t[4 * e + 2]) * 1e-3);
            free (p2p_args.trace);
            p2p_args.trace = nullptr;
         }
         if (comm && (p2p || !p2p_opened.empty ()))
         {
            // nobody unmaps or frees while a peer may still be storing into it
            int *d = alloc<int> (1);
            zero (d, sizeof (int));
            note (nccl ().AllReduce (d, d, 1, ncclInt, ncclSum, comm, stream));
            sync ();
            free (d);
         }
         for (void *q : p2p_opened) cudaIpcCloseMemHandle (q);
         p2p_opened.clear ();
         free (p2p_flags);
         free (p2p_epochs);
         free (p2p_counter);
         free (p2p_fused_dev);
         free (p2p_send_counters);
         p2p_fused_dev = nullptr;
         p2p_send_counters = nullptr;
         p2p_fused = false;
         p2p_flags = p2p_epochs = nullptr;
         p2p_counter = nullptr;
         p2p = false;
      }
      // one stage's exchange of buffer `buf`: push to the peers, then wait for theirs
      bool p2p_exchange (int buf, bool wait = true)
      {
         if (!p2p) return false;
         dflo::P2PArgs a = p2p_args;
         a.wait = wait ? 1 : 0;
         a.srcU = p2p_U[buf];
         a.srcA = p2p_A[buf];
         for (int p = 0; p < a.npeers; ++p)
         {
            a.dstU[p] = p2p_peerU[p][buf];
            a.dstA[p] = p2p_peerA[p][buf];
         }
         ++launches;
         static const int dbg = std::getenv ("DFLO_B200_P2P_DBG") ? std::atoi (std::getenv ("DFLO_B200_P2P_DBG")) : 0; // timing experiments only
         if (dbg & 4) return true;
         if (dbg & 1) a.nseg = 0;
         k_begin ("halo_push_kernel");
         launch_cfg (dflo::halo_push_kernel, dim3 (p2p_grid), dim3 (256), 0, a);
         k_end ();
         note (cudaPeekAtLastError ());
         return true;
      }
      // after exchanges whose wait was deferred: nothing of the peers' last exchange is still in flight when this returns
      void p2p_drain ()
      {
         if (!p2p) return;
         ++launches;
         dflo::halo_wait_kernel<<<1, 32, 0, stream>>> (p2p_args);
         note (cudaPeekAtLastError ());
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
