// Halo exchange over peer memory (NVLink 5 / NVSwitch P2P) for a cell-sharded context, one
// process per GPU.  Replaces the per-stage NCCL send/recv group (and the per-step dt all-reduce)
// of the src_mpi ghost exchange (reference src_mpi/claw.cc:331-340, 579) by plain stores into the
// peers' ghost ranges:
//
//   * every rank maps its peers' solution / cell-average buffers and a small flag block through
//     CUDA IPC once, at context creation (the handles travel over the NCCL communicator);
//   * after the stage (and limiter) kernels of a stage, ONE kernel on the same stream
//       - gathers the owned cells each peer needs and stores them straight into that peer's
//         ghost range (contiguous per peer: no unpack), solution and cell averages,
//       - fences system-wide and raises the "data" flag on every peer.
//     No "ready to receive" handshake is needed: the stores go into the peer's copy of the buffer
//     this stage WROTE, whose ghost range the peer does not read before its next stage (it has
//     not started that stage: it waits for this very flag), and the stage kernel never stores the
//     solution of redundantly updated ghost cells (only their means, which are bit-identical to
//     the ones sent);
//       - and, in its last block, waits for the peers' data flags, so the next kernel on the
//         stream may read the ghosts;
// Epochs are device-side counters, so the whole time step -- exchanges included -- is captured in
// a CUDA graph and replayed without touching the host.
#pragma once

#include <cuda_runtime.h>

namespace dflo
{
   constexpr int P2P_MAX_WORLD = 8;
   constexpr int P2P_MAX_SEG = 2 * (P2P_MAX_WORLD - 1);

   struct P2PSeg
   {
      const int *cells; // owned local cells to send, in the peer's ghost order
      int n;
      int peer;         // index into the peer tables below
      int dst_cell0;    // first cell of the peer's ghost range these land in
   };

   // flag block of a rank, in 8-byte words indexed by SOURCE rank: (unused)[W] | data[W] | dt slots, even epochs[W] | odd epochs[W]
   struct P2PArgs
   {
      const double *srcU, *srcA;
      double *dstU[P2P_MAX_WORLD], *dstA[P2P_MAX_WORLD];       // per peer index: the peer's buffers of this exchange
      unsigned long long *peer_flags[P2P_MAX_WORLD];           // per peer index
      int peer_rank[P2P_MAX_WORLD];
      P2PSeg seg[P2P_MAX_SEG];
      unsigned long long *all_flags[P2P_MAX_WORLD];            // per rank (dt exchange); [me] = my_flags
      unsigned long long *my_flags;
      unsigned long long *epochs;                              // local: [0] exchanges published, [2] dt reductions
      unsigned int *counter;
      double *dt_val;
      unsigned long long *trace; // optional timeline (DFLO_B200_P2P_TRACE): 4 globaltimer stamps per exchange
      int nseg, npeers, me, world, D;
      int wait;                  // 1: the kernel does not end before the peers' data of this exchange is here; 0: the readers wait
   };
   // What the fused (stage-kernel) form of the exchange needs, resident in device memory, one
   // instance per output buffer: tiles whose halo holds ghost cells wait for the peers' previous
   // exchange before they stage it; tiles that own cells a peer needs store them into the peer's
   // ghost range right after their own write-back, and the last such tile raises the data flags.
   struct P2PFused
   {
      double *dstU[P2P_MAX_WORLD], *dstA[P2P_MAX_WORLD];
      unsigned long long *peer_flags[P2P_MAX_WORLD];
      int peer_rank[P2P_MAX_WORLD];
      unsigned long long *my_flags, *epochs;
      unsigned int *send_counter, *block_counter;
      const int *send_entries; // 3 ints per entry: local cell, peer index, destination cell on the peer
      int n_send_tiles, npeers, me, world;
   };

   __device__ __forceinline__ unsigned long long global_ns ()
   {
      unsigned long long t;
      asm volatile ("mov.u64 %0, %%globaltimer;" : "=l"(t));
      return t;
   }

   __device__ __forceinline__ void st_release_sys (unsigned long long *p, unsigned long long v)
   {
      asm volatile ("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
   }
   __device__ __forceinline__ unsigned long long ld_acquire_sys (const unsigned long long *p)
   {
      unsigned long long v;
      asm volatile ("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
      return v;
   }

   __global__ void __launch_bounds__ (256) halo_push_kernel (const P2PArgs a)
   {
      // programmatic dependent launch: wait for the kernel that wrote what is sent, then let the successor be scheduled
      asm volatile ("griddepcontrol.wait;" ::: "memory");
      asm volatile ("griddepcontrol.launch_dependents;" ::: "memory");
      __shared__ unsigned long long s_epoch;
      if (threadIdx.x == 0)
      {
         s_epoch = *reinterpret_cast<volatile unsigned long long *> (a.epochs);
         if (a.trace && blockIdx.x == 0 && s_epoch < 4096) a.trace[4 * s_epoch] = global_ns ();
      }
      __syncthreads ();
      const int stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
      for (int s = 0; s < a.nseg; ++s)
      {
         const P2PSeg sg = a.seg[s];
         double *dU = a.dstU[sg.peer] + (size_t) sg.dst_cell0 * a.D;
         double *dA = a.dstA[sg.peer] + (size_t) sg.dst_cell0 * 4;
         const int D2 = a.D / 2; // D = 4 n_s is even: move 16 bytes per thread
         for (int i = t0; i < sg.n * D2; i += stride)
         {
            const int r = i / D2, c = i - r * D2;
            reinterpret_cast<double2 *> (dU)[i] = reinterpret_cast<const double2 *> (a.srcU + (size_t) sg.cells[r] * a.D)[c];
         }
         for (int i = t0; i < sg.n * 2; i += stride)
            reinterpret_cast<double2 *> (dA)[i] = reinterpret_cast<const double2 *> (a.srcA + (size_t) sg.cells[i >> 1] * 4)[i & 1];
      }
      __threadfence (); // device scope is enough here: the last block's system-scope release below is cumulative
      __syncthreads ();
      if (threadIdx.x == 0)
      {
         const unsigned int done = atomicAdd (a.counter, 1u);
         if (done == gridDim.x - 1)
         {
            *a.counter = 0;
            if (a.trace && s_epoch < 4096) a.trace[4 * s_epoch + 1] = global_ns ();
            __threadfence_system ();
            for (int p = 0; p < a.npeers; ++p) st_release_sys (a.peer_flags[p] + a.world + a.me, s_epoch); // data has landed
            a.epochs[0] = s_epoch + 1;
            if (a.trace && s_epoch < 4096) a.trace[4 * s_epoch + 2] = global_ns ();
            // ... and the kernel does not end before the peers' data is here -- unless the readers of the ghost cells wait
            // themselves (row stage kernel: its ghost-reading tiles, last in the tile order, check the flags)
            if (a.wait)
               for (int p = 0; p < a.npeers; ++p)
                  while (ld_acquire_sys (a.my_flags + a.world + a.peer_rank[p]) < s_epoch) {}
            if (a.trace && s_epoch < 4096) a.trace[4 * s_epoch + 3] = global_ns ();
         }
      }
   }

   // the wait alone: everything the peers published up to the last exchange of this rank has landed
   __global__ void halo_wait_kernel (const P2PArgs a)
   {
      if (threadIdx.x != 0) return;
      const unsigned long long e = *reinterpret_cast<volatile unsigned long long *> (a.epochs) - 1;
      for (int p = 0; p < a.npeers; ++p)
         while (ld_acquire_sys (a.my_flags + a.world + a.peer_rank[p]) < e) {}
   }

   // global minimum of *dt_val over the ranks (compute_time_step, reference src_mpi/claw.cc:579),
   // optionally followed by the finalisation of the step (DtFinalizeKernel's work: a launch less).
   // The value itself is the message: every rank stores it into slot [parity][me] of each other
   // rank's flag block (one 8-byte store: no data/flag pair, no system fence), spins until its own
   // slots [parity][r] turn non-negative, and re-arms them with -1 for the exchange after next.
   __global__ void dt_min_kernel (const P2PArgs a, int finalize, double time_step)
   {
      // launched as a programmatic dependent of the per-cell time-step kernel and itself the predecessor of the first
      // stage kernel: that one may start at once (it reads dt in its last phase, behind its own wait for this kernel);
      // this one waits here for the local minimum to be complete
      asm volatile ("griddepcontrol.launch_dependents;" ::: "memory");
      asm volatile ("griddepcontrol.wait;" ::: "memory");
      const int r = threadIdx.x;
      const unsigned long long e = *reinterpret_cast<volatile unsigned long long *> (a.epochs + 2);
      const int par = (int) (e & 1ull);
      const double v = *a.dt_val;
      __shared__ double got[P2P_MAX_WORLD];
      if (r < a.world) got[r] = v;
      if (r < a.world && r != a.me)
      {
         reinterpret_cast<volatile double *> (a.all_flags[r] + (2 + par) * a.world)[a.me] = v;
         volatile double *mine = reinterpret_cast<volatile double *> (a.my_flags + (2 + par) * a.world) + r;
         double w;
         while ((w = *mine) < 0.0) {}
         *mine = -1.0;
         got[r] = w;
      }
      __syncthreads ();
      if (r == 0)
      {
         double m = v;
         for (int q = 0; q < a.world; ++q) m = (got[q] < m) ? got[q] : m;
         double *time = a.dt_val - 2; // dt_val = time + 2 (the accumulator)
         if (finalize) // claw.cc:468-476, as DtFinalizeKernel
         {
            double dt = m;
            if (time_step != -2.0) // -2: time step type = local, the minimum as it is
            {
               if (dt > 0 && time_step > 0) dt = (time_step < dt) ? time_step : dt;
               if (time[0] + dt > time[3]) dt = time[3] - time[0];
            }
            time[1] = dt;
            time[2] = 1.0e20;
         }
         else
            *a.dt_val = m;
         a.epochs[2] = e + 1;
      }
   }
}
