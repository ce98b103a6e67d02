// Kernels of the explicit RK stage, written as "phase" functions: phase(p, args, smem, tid, bid)
// is the work of one thread between two block barriers.  launch.cuh turns a kernel class into a
// __global__ function (phases unrolled, __syncthreads between them); tests/emu runs the very
// same phase code thread by thread on the CPU so index arithmetic is checked against the oracle
// before GPU time is spent.  No state lives in registers across phases: everything a later
// phase needs is in shared memory.
//
// Layout: DoFs are kept in the reference's order, u[cell*D + comp*NS + node], so
// a block of CPB consecutive cells is one contiguous run of CPB*D doubles that is moved
// HBM <-> shared memory with fully coalesced flat copies.  A group of G = (k+1)^2 threads works
// on one cell: one thread per Gauss point (volume flux), per face point (Riemann flux, 4(k+1)
// per cell, looped) and per DoF (residual + RK combine).
//
// Cell-centric faces: every cell evaluates the numerical flux of all four of its faces, so the
// residual, M^-1, the RK combine and the cell average are produced in one pass with no atomics
// and rhs never touches HBM.  For an interior face both adjacent cells evaluate the SAME
// function call the reference makes once (plus side = the cell MeshWorker::loop visits the face
// from, reference src/assemble_explicit.cc:440-451), so both obtain bit-identical fluxes and
// the scheme stays conservative to round-off exactly like the reference's single evaluation.
#pragma once

#include "euler.cuh"
#include "tables.h"

#include <stddef.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define DFLO_DEV __device__ __forceinline__
#else
#define DFLO_DEV inline
#endif

namespace dflo
{
#if !defined(__CUDACC__)
   struct double2 // the CPU emulation's stand-in for the CUDA vector type
   {
      double x, y;
   };
#endif
   enum { FACE_OWNER = 1, FACE_PERIODIC = 2, FACE_FLIP = 4, JOB_SHARED = 8 };
   enum { MODE_STAGE = 0, MODE_RHS = 1 };
   enum { ERR_NEGATIVE_STATE = 1, ERR_POSLIM_ROOT = 2 };

   constexpr int cells_per_block (int G) { return G >= 128 ? 1 : 128 / G; }
   constexpr int n_scalar (int basis, int n1) { return basis == BASIS_QK ? n1 * n1 : n1 * (n1 + 1) / 2; }
   constexpr int n_gll (int n1) { return ((n1 - 1) + 3) % 2 == 0 ? ((n1 - 1) + 3) / 2 : ((n1 - 1) + 4) / 2; }

   // Tile shape of the stage kernel per degree: tx x ty cells, (k+1)^2 threads per cell, 256 threads
   // per block (the host renumbers cells tile-major with the same numbers, partition.h).
   constexpr int tile_nx (int n1) { return n1 == 1 ? 16 : n1 == 2 ? 8 : n1 == 3 ? 7 : n1 == 4 ? 4 : 5; }
   constexpr int tile_ny (int n1) { return n1 == 1 ? 16 : n1 == 2 ? 8 : n1 == 3 ? 4 : n1 == 4 ? 4 : 2; }

   // flat table layouts (built by pack_stage_tables / pack_limiter_tables in tables_pack.h)
   constexpr int stage_table_size (int basis, int n1)
   {
      return basis == BASIS_QK ? n1 * n1 + 3 * n1 : 3 * n1 * n1 * n_scalar (basis, n1) + 4 * n1 * n_scalar (basis, n1) + n1;
   }
   constexpr int limiter_table_size (int basis, int n1)
   {
      return basis == BASIS_QK ? 3 * n1 + n_gll (n1) * n1 : 2 * n_gll (n1) * n1 * n_scalar (basis, n1);
   }

   // One unique face of a tile (built by build_tile_jobs, partition.h)
   struct alignas (16) FaceJob
   {
      int a;       // slot of the visiting cell in the tile * 4 + its local face number
      int nb;      // neighbour: local cell id, or -1 - local boundary face
      int slot_b;  // where the neighbour's DoFs sit in shared memory (tile slot, or tile_cells + halo
                   // slot), or -1: read them from global memory
      int flags;   // FACE_* | JOB_SHARED
   };

   // One tile of the stage kernel: cells [c0, c0+ncb), its staged halo cells halo_cells[h0 .. h0+nh)
   // and its unique faces jobs[j0 .. j0+nj)
   struct alignas (32) TileDesc
   {
      int c0, ncb, h0, nh, j0, nj, pad0, pad1;
   };

   struct P2PFused; // p2p_halo.cuh (CUDA backend only)

   struct StageArgs
   {
      const double *u;        // current_solution   [n_local][D]
      const double *u_old;    // old_solution
      double *out;            // MODE_STAGE: updated solution (a different buffer than u); MODE_RHS: right_hand_side
      const double *avg;      // cell_average of u   [n_local][4]
      double *avg_out;        // cell_average of the updated solution
      const TileDesc *tiles;  // [n_tiles]
      const int *halo_cells;
      const FaceJob *jobs;
      const double *geom;     // [n_local][4] x0, y0, hx, hy
      const double *bc_g;     // [n_bfaces][n_q_face][4]
      const int *bkind;       // [n_bfaces]
      const double *tab;      // flat stage tables
      const double *time;     // device scalars: [0] elapsed time, [1] dt
      const double *dt_cell;  // optional per-cell dt (local time stepping), else nullptr
      const int *rowdesc;     // tile descriptors of the register-blocked Qk kernel (row_desc.h), or nullptr
      int n_cells_u;          // cells held by u / u_old (bounds the L2 prefetch hints)
      int pf_tiles;           // row kernel: prefetch distance in tiles (resident blocks of the device)
      int n_tiles_owned;      // tiles >= this one are redundantly updated ghost cells: only their means are stored
      const P2PFused *fx;     // row kernel: halo exchange over peer memory fused into the stage kernel, or nullptr
      int pdl;                // row kernel launched as a programmatic dependent of its predecessor on the stream: 0 no;
                              // 1 the predecessor only wrote the time scalars (everything else may be read at once);
                              // 2 the predecessor wrote u / cell averages (nothing of this step's data may be read before the wait)
      int dbg;                // developer timing experiments only (DFLO_B200_DBG): 1 no Riemann solves, 2 no volume fluxes, 4 no edge jobs
      int mode;
      int compat_mpi;
      double ark;
      double gravity;
      const double *ext_force; // [n_local][n_q][2] external force at the cell quadrature points (src_mpi/assemble_explicit.cc:56-58),
                               // or nullptr: the hard-wired (0,-1) of src/ (equation.h:829-850)
   };

#if defined(__CUDACC__)
   // ---- sm_100a asynchronous bulk copies (TMA unit, 1-D): global -> shared completing on an
   //      mbarrier, shared -> global as a bulk group ----
   // Programmatic dependent launch (sm_90+): a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start
   // while its predecessor on the stream still runs; it must not read what the predecessor writes before pdl_wait (),
   // which returns when the predecessor has completed and flushed.  Both are no-ops in an ordinary launch.
   __device__ __forceinline__ void pdl_launch_dependents () { asm volatile ("griddepcontrol.launch_dependents;" ::: "memory"); }
   __device__ __forceinline__ void pdl_wait () { asm volatile ("griddepcontrol.wait;" ::: "memory"); }
   __device__ __forceinline__ unsigned smem_addr (const void *p) { return (unsigned) __cvta_generic_to_shared (p); }
   __device__ __forceinline__ void mbar_init (void *bar, unsigned count)
   {
      asm volatile ("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr (bar)), "r"(count) : "memory");
      asm volatile ("fence.mbarrier_init.release.cluster;" ::: "memory");
   }
   __device__ __forceinline__ void mbar_expect_tx (void *bar, unsigned bytes)
   {
      asm volatile ("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr (bar)), "r"(bytes) : "memory");
   }
   __device__ __forceinline__ void mbar_wait (void *bar, unsigned parity)
   {
      asm volatile ("{\n"
                    ".reg .pred p;\n"
                    "WAIT_%=:\n"
                    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                    "@p bra DONE_%=;\n"
                    "bra WAIT_%=;\n"
                    "DONE_%=:\n"
                    "}" ::"r"(smem_addr (bar)), "r"(parity) : "memory");
   }
   // same, for a warp that has nothing else to do: let the hardware park it (suspend-time hint)
   __device__ __forceinline__ void mbar_wait_parked (void *bar, unsigned parity)
   {
      asm volatile ("{\n"
                    ".reg .pred p;\n"
                    "WAIT_%=:\n"
                    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
                    "@p bra DONE_%=;\n"
                    "bra WAIT_%=;\n"
                    "DONE_%=:\n"
                    "}" ::"r"(smem_addr (bar)), "r"(parity), "r"(20000u) : "memory");
   }
   __device__ __forceinline__ void bulk_g2s (void *dst, const void *src, unsigned bytes, void *bar)
   {
      asm volatile ("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr (dst)),
                    "l"(src), "r"(bytes), "r"(smem_addr (bar))
                    : "memory");
   }
   __device__ __forceinline__ void bulk_s2g (void *dst, const void *src, unsigned bytes)
   {
      asm volatile ("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_addr (src)), "r"(bytes) : "memory");
      asm volatile ("cp.async.bulk.commit_group;" ::: "memory");
   }
   __device__ __forceinline__ void bulk_s2g_wait () { asm volatile ("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
   __device__ __forceinline__ void fence_async_smem () { asm volatile ("fence.proxy.async.shared::cta;" ::: "memory"); }
   __device__ __forceinline__ void mbar_arrive (void *bar)
   {
      asm volatile ("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr (bar)) : "memory");
   }
   template <int ID, int COUNT>
   __device__ __forceinline__ void named_barrier () { asm volatile ("bar.sync %0, %1;" ::"n"(ID), "n"(COUNT) : "memory"); }
#endif

   template <int BASIS, int N1, int FLUX>
   struct StageKernel
   {
      typedef StageArgs Args;
      static constexpr int NQ = N1 * N1;
      static constexpr int G = NQ;
      static constexpr int NS = n_scalar (BASIS, N1);
      static constexpr int D = 4 * NS;
      static constexpr int TC = tile_nx (N1) * tile_ny (N1);        // cells per tile
      static constexpr int NH = 2 * (tile_nx (N1) + tile_ny (N1));  // staged halo cells per tile
      static constexpr int THREADS = (TC * G + 31) / 32 * 32;
#ifndef DFLO_STAGE_MIN_BLOCKS
#define DFLO_STAGE_MIN_BLOCKS 4
#endif
      static constexpr int MIN_BLOCKS = DFLO_STAGE_MIN_BLOCKS; // resident blocks per SM the register budget is held to
      static constexpr int NPHASE = 6;
      static constexpr int FLUX_ID = FLUX;
      static constexpr int BASIS_ID = BASIS;
      static constexpr int N1_ID = N1;
      static constexpr int TAB = stage_table_size (BASIS, N1);
      // shared memory carve-up (in doubles); every bulk-copy destination is 16-byte aligned
      static constexpr int O_TAB = 2;                                // [0,2): the mbarrier
      static constexpr int O_U = O_TAB + (TAB + 1) / 2 * 2;
      static constexpr int O_F = O_U + (TC + NH) * D;
      static constexpr int O_H = O_F + TC * 8 * NQ;
      static constexpr int O_W = O_H + TC * 16 * N1;
      static constexpr int O_UOLD = O_W + (BASIS == BASIS_PK ? TC * 4 * NQ : 0);   // old_solution of the tile
      static constexpr int O_GEOM = O_UOLD + TC * D;                             // x0 y0 hx hy per tile cell
      static constexpr int O_AVG = O_GEOM + TC * 4;                              // cell averages, tile + halo (LxF only)
      static constexpr int O_JOBS = O_AVG + (flux_uses_averages (FLUX) ? (TC + NH) * 4 : 0);
      static constexpr int O_DT = O_JOBS + TC * 4 * 2;                           // FaceJob = 2 doubles
      static constexpr int SMEM_DOUBLES = O_DT + 2;
      // pipelined (persistent) form: barriers | tables | work arrays | old_solution | dt | 2 input stages
      static constexpr int PERSIST_BLOCKS = 3;                       // resident blocks per SM aimed at
      static constexpr int P_TAB = 8;                                // [0,8): six mbarriers
      static constexpr int P_F = P_TAB + (TAB + 1) / 2 * 2;
      static constexpr int P_H = P_F + TC * 8 * NQ;
      static constexpr int P_W = P_H + TC * 16 * N1;
      static constexpr int P_UOLD = P_W + (BASIS == BASIS_PK ? TC * 4 * NQ : 0);
      static constexpr int P_DT = P_UOLD + TC * D;
      static constexpr int P_STAGE = P_DT + 2;
      static constexpr int S_U = 0;                                  // offsets inside one input stage
      static constexpr int S_GEOM = S_U + (TC + NH) * D;
      static constexpr int S_AVG = S_GEOM + TC * 4;
      static constexpr int S_JOBS = S_AVG + (flux_uses_averages (FLUX) ? (TC + NH) * 4 : 0);
      static constexpr int S_TD = S_JOBS + TC * 4 * 2;
      static constexpr int STAGE_DOUBLES = S_TD + 4;
      static constexpr int PERSIST_SMEM_DOUBLES = P_STAGE + 2 * STAGE_DOUBLES;

      // table accessors ------------------------------------------------------------------------
      // Qk: dw[N1*N1] e0[N1] e1[N1] gw[N1]
      // Pk: phi[NQ*NS] dphix[NQ*NS] dphiy[NQ*NS] phiface[4*N1*NS] gw[N1]
      static DFLO_DEV const double *t_dw (const double *tb) { return tb; }
      static DFLO_DEV const double *t_e (const double *tb, int side) { return tb + N1 * N1 + side * N1; }
      static DFLO_DEV const double *t_gw (const double *tb)
      {
         return BASIS == BASIS_QK ? tb + N1 * N1 + 2 * N1 : tb + 3 * NQ * NS + 4 * N1 * NS;
      }
      static DFLO_DEV const double *t_phi (const double *tb) { return tb; }
      static DFLO_DEV const double *t_dphix (const double *tb) { return tb + NQ * NS; }
      static DFLO_DEV const double *t_dphiy (const double *tb) { return tb + 2 * NQ * NS; }
      static DFLO_DEV const double *t_phiface (const double *tb) { return tb + 3 * NQ * NS; }

      // Trace of one cell (ucell = its D DoFs, shared or global memory) at point q of its face f.
      // The same fma chain is used for a cell's own trace and for its neighbour's, so whichever
      // block evaluates a face feeds bit-identical states into the Riemann solver.
      static DFLO_DEV void trace (const double *tb, const double *ucell, int f, int q, double W[4])
      {
         if (BASIS == BASIS_QK)
         {
            const double *e = t_e (tb, f & 1);
            const int base = (f < 2) ? N1 * q : q;
            const int stride = (f < 2) ? 1 : N1;
#pragma unroll
            for (int c = 0; c < 4; ++c)
            {
               double s = 0.0;
#pragma unroll
               for (int a = 0; a < N1; ++a) s = fma (e[a], ucell[c * NS + base + a * stride], s);
               W[c] = s;
            }
         }
         else
         {
            const double *pf = t_phiface (tb) + (f * N1 + q) * NS;
#pragma unroll
            for (int c = 0; c < 4; ++c)
            {
               double s = 0.0;
#pragma unroll
               for (int m = 0; m < NS; ++m) s = fma (pf[m], ucell[c * NS + m], s);
               W[c] = s;
            }
         }
      }

      // where the pieces of one tile live in shared memory
      struct Views
      {
         double *tb, *su, *sF, *sH, *sW, *sUold, *sGeom, *sAvg, *sDt;
         FaceJob *sJobs;
      };

      // one-tile-per-block form (generic phase_kernel, and the CPU emulation of tests/emu)
      static DFLO_DEV void phase (int p, const Args &A, double *sm, int tid, int bid)
      {
         const Views v = {sm + O_TAB, sm + O_U, sm + O_F, sm + O_H, sm + O_W, sm + O_UOLD, sm + O_GEOM, sm + O_AVG, sm + O_DT,
                          reinterpret_cast<FaceJob *> (sm + O_JOBS)};
         work (p, A, A.tiles[bid], v, sm, tid, false);
      }

      // p = 0: stage the tile (one-tile-per-block form only); 1: volume fluxes + own traces; 2: face
      // fluxes; 3: residual, M^-1, RK combine; 4: write back + row sums; 5: cell averages.  `persistent`: called from the pipelined
      // kernel below (inputs already staged by its producer warp, plain stores on the way out).
      static DFLO_DEV void work (int p, const Args &A, const TileDesc &td, const Views &v, double *sm, int tid, bool persistent)
      {
         const int c0 = td.c0, ncb = td.ncb;
         double *tb = v.tb, *su = v.su, *sF = v.sF, *sH = v.sH, *sW = v.sW, *sUold = v.sUold, *sGeom = v.sGeom, *sAvg = v.sAvg;
         FaceJob *sJobs = v.sJobs;
         const int slot = tid / G, lq = tid % G;
         const bool active = slot < ncb;
         const int cell = c0 + slot;
         const bool need_old = A.mode == MODE_STAGE && A.ark != 0.0;

         if (p == 0)
         {
            // Stage EVERYTHING this tile reads from global memory in shared memory, as bulk async
            // copies completing on one mbarrier: the tile's DoFs (one contiguous run), its halo
            // cells, old_solution of the tile, geometry, the unique-face list and (LxF) the cell
            // averages.  Later phases touch global memory only to write.
            const int h0 = td.h0, nh = td.nh, nj = td.nj;
#if defined(__CUDA_ARCH__)
            const unsigned cell_bytes = (unsigned) (D * sizeof (double));
            if (tid == 0)
            {
               unsigned bytes = (unsigned) (ncb + nh) * cell_bytes + (unsigned) ncb * 32u + (unsigned) nj * 16u;
               if (need_old) bytes += (unsigned) ncb * cell_bytes;
               if (flux_uses_averages (FLUX)) bytes += (unsigned) (ncb + nh) * 32u;
               mbar_init (sm, 1);
               mbar_expect_tx (sm, bytes);
            }
            __syncthreads ();
            if (tid == 0)
            {
               bulk_g2s (su, A.u + (size_t) c0 * D, (unsigned) ncb * cell_bytes, sm);
               if (need_old) bulk_g2s (sUold, A.u_old + (size_t) c0 * D, (unsigned) ncb * cell_bytes, sm);
               bulk_g2s (sGeom, A.geom + (size_t) c0 * 4, (unsigned) ncb * 32u, sm);
               if (nj) bulk_g2s (sJobs, A.jobs + td.j0, (unsigned) nj * 16u, sm);
               if (flux_uses_averages (FLUX)) bulk_g2s (sAvg, A.avg + (size_t) c0 * 4, (unsigned) ncb * 32u, sm);
            }
            else if (tid <= nh)
            {
               const int hc = A.halo_cells[h0 + tid - 1];
               bulk_g2s (su + (TC + tid - 1) * D, A.u + (size_t) hc * D, cell_bytes, sm);
               if (flux_uses_averages (FLUX)) bulk_g2s (sAvg + (TC + tid - 1) * 4, A.avg + (size_t) hc * 4, 32u, sm);
            }
            if (tid == THREADS - 1) v.sDt[0] = A.time[1];
            for (int i = tid; i < TAB; i += THREADS) tb[i] = A.tab[i];
            mbar_wait (sm, 0);
#else
            for (int i = tid; i < ncb * D; i += THREADS) su[i] = A.u[(size_t) c0 * D + i];
            for (int i = tid; i < nh * D; i += THREADS) su[TC * D + i] = A.u[(size_t) A.halo_cells[h0 + i / D] * D + i % D];
            if (need_old)
               for (int i = tid; i < ncb * D; i += THREADS) sUold[i] = A.u_old[(size_t) c0 * D + i];
            for (int i = tid; i < ncb * 4; i += THREADS) sGeom[i] = A.geom[(size_t) c0 * 4 + i];
            for (int i = tid; i < nj; i += THREADS) sJobs[i] = A.jobs[td.j0 + i];
            if (flux_uses_averages (FLUX))
            {
               for (int i = tid; i < ncb * 4; i += THREADS) sAvg[i] = A.avg[(size_t) c0 * 4 + i];
               for (int i = tid; i < nh * 4; i += THREADS) sAvg[TC * 4 + i] = A.avg[(size_t) A.halo_cells[h0 + i / 4] * 4 + i % 4];
            }
            if (tid == THREADS - 1) v.sDt[0] = A.time[1];
            for (int i = tid; i < TAB; i += THREADS) tb[i] = A.tab[i];
#endif
         }
         else if (p == 1)
         {
            // ---- volume: Cartesian fluxes at Gauss point lq (assemble_explicit.cc:57-79) ----
            if (active)
            {
               const double *uc = su + slot * D;
               double W[4], Fx[4], Fy[4];
               if (BASIS == BASIS_QK)
               {
                  // collocation: the interpolation loop 65-75 is the identity
#pragma unroll
                  for (int c = 0; c < 4; ++c) W[c] = uc[c * NS + lq];
               }
               else
               {
                  const double *ph = t_phi (tb) + lq * NS;
#pragma unroll
                  for (int c = 0; c < 4; ++c)
                  {
                     double s = 0.0;
#pragma unroll
                     for (int m = 0; m < NS; ++m) s = fma (ph[m], uc[c * NS + m], s);
                     W[c] = s;
                     sW[(slot * 4 + c) * NQ + lq] = s;
                  }
               }
               flux_matrix (W, Fx, Fy);
#pragma unroll
               for (int c = 0; c < 4; ++c)
               {
                  sF[(slot * 8 + c) * NQ + lq] = Fx[c];
                  sF[(slot * 8 + 4 + c) * NQ + lq] = Fy[c];
               }
            }
            // ---- traces of the cell on its own four faces, one (face, point) per thread: each
            //      half-warp reads one cell, which is free of shared-memory bank conflicts.  They
            //      are parked in the slots the fluxes will overwrite (sH). ----
            if (active)
            {
               const double *uc = su + slot * D;
               for (int i = lq; i < 4 * N1; i += G)
               {
                  double W[4];
                  trace (tb, uc, i / N1, i % N1, W);
#pragma unroll
                  for (int c = 0; c < 4; ++c) sH[(slot * 4 * N1 + i) * 4 + c] = W[c];
               }
            }
         }
         else if (p == 2)
         {
            // ---- faces: one numerical flux per UNIQUE face point of the tile, along the visiting
            //      cell's outward normal (assemble_explicit.cc:176-206, 303-341; periodic: src_mpi
            //      186-260); a face shared by two cells of the tile serves both.  A job owns the
            //      sH slots of its face point(s): it reads the traces parked there, then overwrites
            //      them with the flux. ----
            const int njq = td.nj * N1;
            for (int idx = tid; idx < njq; idx += THREADS)
            {
               const FaceJob job = sJobs[idx / N1];
               const int q = idx % N1;
               const int sa = job.a >> 2, f = job.a & 3;
               const int nb = job.nb, fl = job.flags;
               const double nx = (f == 0) ? -1.0 : (f == 1) ? 1.0 : 0.0;
               const double ny = (f == 2) ? -1.0 : (f == 3) ? 1.0 : 0.0;
               double Wo[4], Wn[4], Ao[4], An[4], H[4];
               double *ho = sH + ((sa * 4 + f) * N1 + q) * 4;
#pragma unroll
               for (int c = 0; c < 4; ++c) Wo[c] = ho[c];
               if (flux_uses_averages (FLUX)) // the only flux that reads the cell averages (equation.h:357-359)
               {
#pragma unroll
                  for (int c = 0; c < 4; ++c) Ao[c] = sAvg[sa * 4 + c];
               }
               bool plus = true; // the visiting cell is the "plus" side of the flux call
               if (nb >= 0)
               {
                  const int qn = (fl & FACE_FLIP) ? N1 - 1 - q : q;
                  if (fl & JOB_SHARED)
                  {
                     const double *hn = sH + ((job.slot_b * 4 + (f ^ 1)) * N1 + qn) * 4;
#pragma unroll
                     for (int c = 0; c < 4; ++c) Wn[c] = hn[c];
                  }
                  else if (job.slot_b >= 0)
                     trace (tb, su + job.slot_b * D, f ^ 1, qn, Wn);
                  else
                     trace (tb, A.u + (size_t) nb * D, f ^ 1, qn, Wn);
                  if (flux_uses_averages (FLUX))
                  {
#pragma unroll
                     for (int c = 0; c < 4; ++c) An[c] = job.slot_b >= 0 ? sAvg[job.slot_b * 4 + c] : A.avg[(size_t) nb * 4 + c];
                  }
                  plus = (fl & (FACE_OWNER | FACE_PERIODIC)) != 0;
               }
               else
               {
                  const int bf = -1 - nb;
                  const int kind = A.bkind[bf];
                  double g[4];
#pragma unroll
                  for (int c = 0; c < 4; ++c) g[c] = A.bc_g[((size_t) bf * N1 + q) * 4 + c];
                  compute_wminus (kind, nx, ny, Wo, g, Wn);
                  if (flux_uses_averages (FLUX))
                  {
                     if (A.compat_mpi) // src_mpi/assemble_explicit.cc:296-321
                        compute_wminus (kind, nx, ny, Ao, g, An);
                     else // src/assemble_explicit.cc:203-204: own average on both sides
                     {
#pragma unroll
                        for (int c = 0; c < 4; ++c) An[c] = Ao[c];
                     }
                  }
               }
               double L[4], R[4], AL[4], AR[4];
#pragma unroll
               for (int c = 0; c < 4; ++c)
               {
                  L[c] = plus ? Wo[c] : Wn[c];
                  R[c] = plus ? Wn[c] : Wo[c];
                  AL[c] = plus ? Ao[c] : An[c];
                  AR[c] = plus ? An[c] : Ao[c];
               }
               const double sg = plus ? 1.0 : -1.0;
               numerical_flux<FLUX> (sg * nx, sg * ny, L, R, AL, AR, H);
#pragma unroll
               for (int c = 0; c < 4; ++c) ho[c] = sg * H[c];
               if (fl & JOB_SHARED) // the neighbour's outward normal is -n
               {
                  double *hn = sH + ((job.slot_b * 4 + (f ^ 1)) * N1 + q) * 4;
#pragma unroll
                  for (int c = 0; c < 4; ++c) hn[c] = -(sg * H[c]);
               }
            }
         }
         else if (p == 3)
         {
            if (!active || lq >= NS) return;
            const double hx = sGeom[slot * 4 + 2], hy = sGeom[slot * 4 + 3];
            const double *gw = t_gw (tb);
            const double dt = A.dt_cell ? A.dt_cell[cell] : v.sDt[0];
            double *uc = su + slot * D;
            const double *H = sH + slot * 16 * N1;
            double r[4];
            double invm;
            if (BASIS == BASIS_QK)
            {
               // rhs_i = sum_q F.grad(phi_i) JxW - sum_faces sum_q H phi_i JxW  (85-115, 209-244,
               // 344-382) with phi_i the Lagrange function of Gauss node (a,b)
               const int a = lq % N1, b = lq / N1;
               const double *dw = t_dw (tb);
               const double *e0 = t_e (tb, 0), *e1 = t_e (tb, 1);
               const double wa = gw[a], wb = gw[b];
#pragma unroll
               for (int c = 0; c < 4; ++c)
               {
                  double sx = 0.0, sy = 0.0;
#pragma unroll
                  for (int ap = 0; ap < N1; ++ap) sx = fma (sF[(slot * 8 + c) * NQ + ap + N1 * b], dw[ap * N1 + a], sx);
#pragma unroll
                  for (int bp = 0; bp < N1; ++bp) sy = fma (sF[(slot * 8 + 4 + c) * NQ + a + N1 * bp], dw[bp * N1 + b], sy);
                  const double fx = H[((1 * N1) + b) * 4 + c] * e1[a] + H[((0 * N1) + b) * 4 + c] * e0[a];
                  const double fy = H[((3 * N1) + a) * 4 + c] * e1[b] + H[((2 * N1) + a) * 4 + c] * e0[b];
                  r[c] = hy * wb * (sx - fx) + hx * wa * (sy - fy);
               }
               if (A.gravity != 0.0)
               {
                  double W[4], Gv[4];
#pragma unroll
                  for (int c = 0; c < 4; ++c) W[c] = uc[c * NS + lq];
                  if (A.ext_force)
                     forcing_ext (W, A.ext_force[((size_t) cell * NQ + lq) * 2], A.ext_force[((size_t) cell * NQ + lq) * 2 + 1], Gv);
                  else
                     forcing (W, Gv);
#pragma unroll
                  for (int c = 0; c < 4; ++c) r[c] += A.gravity * Gv[c] * (wa * wb * hx * hy);
               }
               invm = 1.0 / (wa * wb * hx * hy); // claw.cc:228-258 on a Cartesian cell
            }
            else
            {
               const int m = lq;
               const double *dpx = t_dphix (tb), *dpy = t_dphiy (tb), *pf = t_phiface (tb), *ph = t_phi (tb);
#pragma unroll
               for (int c = 0; c < 4; ++c)
               {
                  double s = 0.0;
                  for (int q = 0; q < NQ; ++q)
                  {
                     const double w2 = gw[q % N1] * gw[q / N1];
                     s += w2 * (sF[(slot * 8 + c) * NQ + q] * dpx[q * NS + m] * hy + sF[(slot * 8 + 4 + c) * NQ + q] * dpy[q * NS + m] * hx);
                  }
                  for (int f = 0; f < 4; ++f)
                  {
                     const double len = (f < 2) ? hy : hx;
                     for (int q = 0; q < N1; ++q) s -= gw[q] * len * H[((f * N1) + q) * 4 + c] * pf[(f * N1 + q) * NS + m];
                  }
                  r[c] = s;
               }
               if (A.gravity != 0.0)
               {
                  for (int q = 0; q < NQ; ++q)
                  {
                     double W[4], Gv[4];
#pragma unroll
                     for (int c = 0; c < 4; ++c) W[c] = sW[(slot * 4 + c) * NQ + q];
                     if (A.ext_force)
                        forcing_ext (W, A.ext_force[((size_t) cell * NQ + q) * 2], A.ext_force[((size_t) cell * NQ + q) * 2 + 1], Gv);
                     else
                        forcing (W, Gv);
                     const double w = gw[q % N1] * gw[q / N1] * hx * hy * ph[q * NS + m];
#pragma unroll
                     for (int c = 0; c < 4; ++c) r[c] += A.gravity * Gv[c] * w;
                  }
               }
               invm = 1.0 / (hx * hy); // orthonormal modes
            }
            if (A.mode == MODE_RHS)
            {
#pragma unroll
               for (int c = 0; c < 4; ++c) uc[c * NS + lq] = r[c];
            }
            else
            {
               // solve() + RK combine: claw.cc:694-713, 757-760
#pragma unroll
               for (int c = 0; c < 4; ++c)
               {
                  const double un = uc[c * NS + lq] + dt * r[c] * invm;
                  if (A.ark != 0.0)
                     uc[c * NS + lq] = (1.0 - A.ark) * un + A.ark * sUold[slot * D + c * NS + lq];
                  else
                     uc[c * NS + lq] = un;
               }
            }
#if defined(__CUDA_ARCH__)
            if (!persistent) fence_async_smem (); // make this thread's shared-memory writes visible to the bulk-copy engine
#endif
         }
         else if (p == 4)
         {
            double *dst = A.out + (size_t) c0 * D;
            // td.pad0: a tile of redundantly updated ghost cells -- only its means are kept, the
            // solution itself arrives with the halo exchange (possibly before this tile runs)
            const bool keep = td.pad0 == 0;
#if defined(__CUDA_ARCH__)
            if (!keep) {}
            else if (persistent) // fire-and-forget 16-byte stores: the stage buffer is free as soon as they are issued
            {
               const double2 *s2 = reinterpret_cast<const double2 *> (su);
               double2 *d2 = reinterpret_cast<double2 *> (dst);
               for (int i = tid; i < ncb * D / 2; i += THREADS) d2[i] = s2[i];
            }
            else if (tid == 0) // the tile goes back as one bulk copy shared -> global
               bulk_s2g (dst, su, (unsigned) (ncb * D * sizeof (double)));
#else
            if (keep)
               for (int i = tid; i < ncb * D; i += THREADS) dst[i] = su[i];
#endif
            // compute_cell_average of the updated solution, claw.cc:562-597, in two steps: weighted
            // sums along x, one (cell, component, row) per thread (scratch = the flux array sF)
            if (A.mode == MODE_STAGE && BASIS == BASIS_QK)
            {
               const double *gw = t_gw (tb);
               for (int j = tid; j < ncb * 4 * N1; j += THREADS)
               {
                  const int b = j % N1, sc = j / N1; // sc = slot * 4 + component
                  double r = 0.0;
#pragma unroll
                  for (int a = 0; a < N1; ++a) r = fma (gw[a], su[sc * NS + a + N1 * b], r);
                  sF[j] = gw[b] * r;
               }
            }
         }
         else // p == 5
         {
            if (A.mode == MODE_STAGE)
            {
               for (int j = tid; j < ncb * 4; j += THREADS)
               {
                  double v;
                  if (BASIS == BASIS_QK)
                  {
                     v = 0.0;
#pragma unroll
                     for (int b = 0; b < N1; ++b) v += sF[j * N1 + b];
                  }
                  else
                     v = su[(j / 4) * D + (j % 4) * NS];
                  A.avg_out[(size_t) c0 * 4 + j] = v;
               }
            }
#if defined(__CUDA_ARCH__)
            if (!persistent && tid == 0) bulk_s2g_wait (); // shared memory must outlive the copy
#endif
         }
      }
   };

#if defined(__CUDACC__)
   // Pipelined form of the stage kernel: one persistent block per SM slot walks over the tiles
   // t = blockIdx.x, blockIdx.x + gridDim.x, ...  K::THREADS consumer threads run phases 1-3 of
   // StageKernel::work on the tile in input stage (it & 1) while ONE producer warp, two tiles
   // ahead, streams the next tiles' inputs into the other stage with bulk async copies (TMA):
   //   full[s]   producer -> consumers   stage s holds tile it (expect_tx byte count)
   //   empty[s]  consumers -> producer   all consumer threads are done with stage s
   //   full_old / empty_old              the same hand-shake for old_solution of the tile
   // so no consumer ever waits on a global-memory load; consumers synchronise among themselves
   // with a named barrier that the producer warp does not take part in.
   template <class K>
   __global__ void __launch_bounds__ (K::THREADS + 32, K::PERSIST_BLOCKS) stage_persistent_kernel (const StageArgs A, int n_tiles)
   {
      extern __shared__ __align__ (16) double sm[];
      const int tid = threadIdx.x;
      unsigned long long *bars = reinterpret_cast<unsigned long long *> (sm); // full0 full1 empty0 empty1 full_old empty_old
      const bool need_old = A.mode == MODE_STAGE && A.ark != 0.0;
      constexpr unsigned cell_bytes = (unsigned) (K::D * sizeof (double));
      if (tid == 0)
      {
         mbar_init (&bars[0], 1);
         mbar_init (&bars[1], 1);
         mbar_init (&bars[2], K::THREADS);
         mbar_init (&bars[3], K::THREADS);
         mbar_init (&bars[4], 1);
         mbar_init (&bars[5], 1);
         sm[K::P_DT] = A.time[1];
      }
      for (int i = tid; i < K::TAB; i += K::THREADS + 32) sm[K::P_TAB + i] = A.tab[i];
      __syncthreads ();

      if (tid >= K::THREADS)
      {
         // ---------------- producer warp ----------------
         const int lane = tid - K::THREADS;
         int it = 0;
         for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it)
         {
            const int s = it & 1;
            double *st = sm + K::P_STAGE + s * K::STAGE_DOUBLES;
            if (it >= 2) mbar_wait_parked (&bars[2 + s], ((it >> 1) - 1) & 1);
            const TileDesc td = A.tiles[t];
            if (lane == 0)
            {
               unsigned bytes = (unsigned) (td.ncb + td.nh) * cell_bytes + (unsigned) td.ncb * 32u + (unsigned) td.nj * 16u + 32u;
               if (flux_uses_averages (K::FLUX_ID)) bytes += (unsigned) (td.ncb + td.nh) * 32u;
               mbar_expect_tx (&bars[s], bytes);
               bulk_g2s (st + K::S_U, A.u + (size_t) td.c0 * K::D, (unsigned) td.ncb * cell_bytes, &bars[s]);
               bulk_g2s (st + K::S_GEOM, A.geom + (size_t) td.c0 * 4, (unsigned) td.ncb * 32u, &bars[s]);
               if (td.nj) bulk_g2s (st + K::S_JOBS, A.jobs + td.j0, (unsigned) td.nj * 16u, &bars[s]);
               bulk_g2s (st + K::S_TD, A.tiles + t, 32u, &bars[s]);
               if (flux_uses_averages (K::FLUX_ID)) bulk_g2s (st + K::S_AVG, A.avg + (size_t) td.c0 * 4, (unsigned) td.ncb * 32u, &bars[s]);
            }
            __syncwarp ();
            for (int h = lane; h < td.nh; h += 32)
            {
               const int hc = A.halo_cells[td.h0 + h];
               bulk_g2s (st + K::S_U + (K::TC + h) * K::D, A.u + (size_t) hc * K::D, cell_bytes, &bars[s]);
               if (flux_uses_averages (K::FLUX_ID)) bulk_g2s (st + K::S_AVG + (K::TC + h) * 4, A.avg + (size_t) hc * 4, 32u, &bars[s]);
            }
            if (need_old)
            {
               if (it >= 1) mbar_wait_parked (&bars[5], (it - 1) & 1);
               if (lane == 0)
               {
                  mbar_expect_tx (&bars[4], (unsigned) td.ncb * cell_bytes);
                  bulk_g2s (sm + K::P_UOLD, A.u_old + (size_t) td.c0 * K::D, (unsigned) td.ncb * cell_bytes, &bars[4]);
               }
            }
         }
         return;
      }

      // ---------------- consumers ----------------
      int it = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it)
      {
         const int s = it & 1;
         double *st = sm + K::P_STAGE + s * K::STAGE_DOUBLES;
         const typename K::Views v = {sm + K::P_TAB, st + K::S_U, sm + K::P_F, sm + K::P_H, sm + K::P_W, sm + K::P_UOLD, st + K::S_GEOM,
                                      st + K::S_AVG, sm + K::P_DT, reinterpret_cast<FaceJob *> (st + K::S_JOBS)};
         mbar_wait (&bars[s], (it >> 1) & 1);
         const TileDesc td = *reinterpret_cast<const TileDesc *> (st + K::S_TD);
         K::work (1, A, td, v, nullptr, tid, true);
         named_barrier<1, K::THREADS> ();
         K::work (2, A, td, v, nullptr, tid, true);
         named_barrier<1, K::THREADS> ();
         if (need_old) mbar_wait (&bars[4], it & 1);
         K::work (3, A, td, v, nullptr, tid, true);
         named_barrier<1, K::THREADS> ();
         if (need_old && tid == 0) mbar_arrive (&bars[5]);
         K::work (4, A, td, v, nullptr, tid, true);
         named_barrier<1, K::THREADS> ();
         K::work (5, A, td, v, nullptr, tid, true);
         // the next tile's phase 1 writes sF, which phase 5 is still reading in slower warps
         named_barrier<1, K::THREADS> ();
         mbar_arrive (&bars[2 + s]);
      }
   }
#endif

   //---------------------------------------------------------------------------------------------
   // Cell averages of a solution vector (claw.cc:562-597), used after set_solution
   //---------------------------------------------------------------------------------------------
   struct AvgArgs
   {
      const double *u;
      double *avg;
      const double *gw; // [N1] device
      const double *gx; // [N1] device (mapping = q1)
      const double *verts; // [n_cells][8] for mapping = q1, else nullptr
      int n_cells, basis, n1, ns;
   };

   DFLO_DEV void cell_average_thread (const AvgArgs &A, int j)
   {
      if (j >= A.n_cells * 4) return;
      const int cell = j / 4, c = j % 4;
      const double *uc = A.u + (size_t) cell * 4 * A.ns + c * A.ns;
      double v;
      if (A.basis == BASIS_QK && A.verts)
      {
         // mapping = q1: sum_q u_q JxW_q / measure (claw.cc:562-597 under MappingQ1)
         const double *q = A.verts + (size_t) cell * 8;
         v = 0.0;
         for (int b = 0; b < A.n1; ++b)
            for (int a = 0; a < A.n1; ++a)
            {
               const double xi = A.gx[a], eta = A.gx[b];
               const double xxi = (q[2] - q[0]) * (1.0 - eta) + (q[6] - q[4]) * eta, xeta = (q[4] - q[0]) * (1.0 - xi) + (q[6] - q[2]) * xi;
               const double yxi = (q[3] - q[1]) * (1.0 - eta) + (q[7] - q[5]) * eta, yeta = (q[5] - q[1]) * (1.0 - xi) + (q[7] - q[3]) * xi;
               v += uc[a + A.n1 * b] * (A.gw[a] * A.gw[b] * (xxi * yeta - xeta * yxi));
            }
         v /= 0.5 * ((q[6] - q[0]) * (q[5] - q[3]) - (q[4] - q[2]) * (q[7] - q[1]));
      }
      else if (A.basis == BASIS_QK)
      {
         v = 0.0;
         for (int b = 0; b < A.n1; ++b)
            for (int a = 0; a < A.n1; ++a) v += A.gw[a] * A.gw[b] * uc[a + A.n1 * b];
      }
      else
         v = uc[0];
      A.avg[j] = v;
   }

   //---------------------------------------------------------------------------------------------
   // TVB limiter (limiter.cc:224-516) fused with the positivity limiter (positivity.cc:16-208),
   // in place on the freshly updated solution.  Reads only the cell's own DoFs and the cell
   // averages of the cell and its four face neighbours, so cells are independent exactly as in
   // the reference's serial loops.
   //---------------------------------------------------------------------------------------------
   struct LimiterArgs
   {
      double *u;              // in place
      const double *avg;
      const int *nbr;
      const unsigned char *fflags;
      const double *geom;
      const double *tab;      // flat limiter tables
      const double *shock;    // shock_indicator per cell (TVB only where > 1, limiter.cc:263, 406), or nullptr: every cell
      int *flags_out;         // [n_local] bit0 TVB rewrote, bit1 theta1<1, bit2 theta2<1
      unsigned int *err;      // device error word
      int n_compute;
      int tvb, char_lim, pos_lim, cam; // tvb: 0 none, 1 TVB, 2 minmax (Qk, LimiterCellKernel only)
      double M, beta;
   };

   template <int BASIS, int N1>
   struct LimiterKernel
   {
      typedef LimiterArgs Args;
      static constexpr int NQ = N1 * N1;
      static constexpr int G = NQ;
      static constexpr int NS = n_scalar (BASIS, N1);
      static constexpr int D = 4 * NS;
      static constexpr int CPB = cells_per_block (G);
      static constexpr int THREADS = CPB * G;
      static constexpr int MIN_BLOCKS = 1;
      static constexpr int NPHASE = 11;
      static constexpr int NGLL = n_gll (N1);
      static constexpr int NPOS = NGLL * N1;
      static constexpr int TAB = limiter_table_size (BASIS, N1);
      static constexpr int O_U = TAB;
      static constexpr int O_D = O_U + CPB * D;        // slopes Dx[4], Dy[4] per cell
      static constexpr int O_N = O_D + CPB * 8;        // limited slopes
      static constexpr int O_P = O_N + CPB * 8;        // point values / theta candidates [2*NPOS]
      static constexpr int O_T = O_P + CPB * 2 * NPOS; // theta1, theta2
      static constexpr int O_FLAG = O_T + CPB * 2;     // per-cell flags (stored as doubles) + block flag
      static constexpr int SMEM_DOUBLES = O_FLAG + CPB + 1;

      static int grid (int n_compute) { return (n_compute + CPB - 1) / CPB; }

      // Qk tables: gw[N1] gx[N1] gdiff[N1] gl_interp[NGLL*N1];  Pk: phipos[2][NPOS][NS]
      static DFLO_DEV const double *t_gw (const double *tb) { return tb; }
      static DFLO_DEV const double *t_gx (const double *tb) { return tb + N1; }
      static DFLO_DEV const double *t_gdiff (const double *tb) { return tb + 2 * N1; }
      static DFLO_DEV const double *t_gli (const double *tb) { return tb + 3 * N1; }
      static DFLO_DEV const double *t_phipos (const double *tb) { return tb; }

      // value of component c at positivity point pt of set (0: GLL x Gauss, 1: Gauss x GLL)
      static DFLO_DEV double point_value (const double *tb, const double *uc, int set, int pt, int c)
      {
         double s = 0.0;
         if (BASIS == BASIS_QK)
         {
            const double *gli = t_gli (tb);
            if (set == 0)
            {
               const int j = pt % NGLL, b = pt / NGLL;
#pragma unroll
               for (int a = 0; a < N1; ++a) s = fma (gli[j * N1 + a], uc[c * NS + a + N1 * b], s);
            }
            else
            {
               const int a = pt % N1, j = pt / N1;
#pragma unroll
               for (int b = 0; b < N1; ++b) s = fma (gli[j * N1 + b], uc[c * NS + a + N1 * b], s);
            }
         }
         else
         {
            const double *pp = t_phipos (tb) + (set * NPOS + pt) * NS;
#pragma unroll
            for (int m = 0; m < NS; ++m) s = fma (pp[m], uc[c * NS + m], s);
         }
         return s;
      }

      static DFLO_DEV void phase (int p, const Args &A, double *sm, int tid, int bid)
      {
         const int c0 = bid * CPB;
         const int ncb = (A.n_compute - c0 < CPB) ? A.n_compute - c0 : CPB;
         double *tb = sm;
         double *su = sm + O_U;
         double *sD = sm + O_D;
         double *sN = sm + O_N;
         double *sP = sm + O_P;
         double *sT = sm + O_T;
         double *sFlag = sm + O_FLAG;
         const int slot = tid / G, lq = tid % G;
         const bool active = slot < ncb;
         const int cell = c0 + slot;
         const double eps = 1.0e-13;
         if (N1 == 1) return; // degree 0: both limiters return at once (limiter.cc:227, positivity.cc:19)

         if (p == 0)
         {
            for (int i = tid; i < TAB; i += THREADS) tb[i] = A.tab[i];
            const double *src = A.u + (size_t) c0 * D;
            for (int i = tid; i < ncb * D; i += THREADS) su[i] = src[i];
            for (int i = tid; i < CPB + 1; i += THREADS) sFlag[i] = 0.0;
         }
         else if (p == 1)
         {
            // mean slopes of the cell, limiter.cc:268-281 (Qk) / 412-420 (Pk)
            if (!active || !A.tvb) return;
            const double *uc = su + slot * D;
            for (int r = lq; r < 8; r += G)
            {
               const int c = r % 4, dir = r / 4;
               double s = 0.0;
               if (BASIS == BASIS_QK)
               {
                  const double *gw = t_gw (tb), *gd = t_gdiff (tb);
                  for (int b = 0; b < N1; ++b)
                     for (int a = 0; a < N1; ++a)
                        s += (dir == 0 ? gd[a] * gw[b] : gw[a] * gd[b]) * uc[c * NS + a + N1 * b];
               }
               else
                  s = uc[c * NS + (dir == 0 ? 1 : N1)] * 1.7320508075688772; // sqrt(3); base index 1 / k+1
               sD[slot * 8 + r] = s;
            }
         }
         else if (p == 2)
         {
            // one thread per cell, packed into the first lanes: differences of means,
            // characteristic projection, minmod (limiter.cc:283-345 / 425-486)
            if (!A.tvb || tid >= ncb) return;
            const int s = tid, cl = c0 + tid;
            if (A.shock && !(A.shock[cl] > 1.0)) return; // limiter.cc:263, 406
            const double hx = A.geom[(size_t) cl * 4 + 2], hy = A.geom[(size_t) cl * 4 + 3];
            const double dx = sqrt (hx * hx + hy * hy) / 1.4142135623730951; // diameter / sqrt(dim)
            const double Mdx2 = A.M * dx * dx;
            const double beta = (BASIS == BASIS_QK) ? A.beta : 0.5 * A.beta;
            double av[4], Dx[4], Dy[4], dbx[4], dfx[4], dby[4], dfy[4];
#pragma unroll
            for (int c = 0; c < 4; ++c)
            {
               av[c] = A.avg[(size_t) cl * 4 + c];
               if (BASIS == BASIS_QK)
               {
                  // Dx = dx * mean(d/dx), mean gradient on the unit cell divided by hx
                  Dx[c] = dx * (sD[s * 8 + c] / hx);
                  Dy[c] = dx * (sD[s * 8 + 4 + c] / hy);
               }
               else
               {
                  Dx[c] = sD[s * 8 + c];
                  Dy[c] = sD[s * 8 + 4 + c];
               }
               dbx[c] = dfx[c] = Dx[c];
               dby[c] = dfy[c] = Dy[c];
            }
            const double ang_mom = Dx[1] - Dy[0];
            // lcell/rcell/bcell/tcell (claw.cc:357-379); periodic partners count as neighbours
            // (src_mpi/claw.cc:417-465)
            for (int f = 0; f < 4; ++f)
            {
               const int nb = A.nbr[(size_t) cl * 4 + f];
               if (nb < 0) continue;
#pragma unroll
               for (int c = 0; c < 4; ++c)
               {
                  const double an = A.avg[(size_t) nb * 4 + c];
                  if (f == 0) dbx[c] = av[c] - an;
                  if (f == 1) dfx[c] = an - av[c];
                  if (f == 2) dby[c] = av[c] - an;
                  if (f == 3) dfy[c] = an - av[c];
               }
            }
            EigenMatrices em;
            if (A.char_lim)
            {
               compute_eigen_matrix (av, em);
               transform_to_char (em.Lx, dbx);
               transform_to_char (em.Lx, dfx);
               transform_to_char (em.Ly, dby);
               transform_to_char (em.Ly, dfy);
               transform_to_char (em.Lx, Dx);
               transform_to_char (em.Ly, Dy);
            }
            double Dxn[4], Dyn[4], change_x = 0.0, change_y = 0.0;
#pragma unroll
            for (int c = 0; c < 4; ++c)
            {
               Dxn[c] = minmod (Dx[c], beta * dbx[c], beta * dfx[c], Mdx2);
               Dyn[c] = minmod (Dy[c], beta * dby[c], beta * dfy[c], Mdx2);
               change_x += fabs (Dxn[c] - Dx[c]);
               change_y += fabs (Dyn[c] - Dy[c]);
            }
            change_x /= 4;
            change_y /= 4;
            if (change_x + change_y > 1.0e-10)
            {
               if (BASIS == BASIS_QK)
               {
#pragma unroll
                  for (int c = 0; c < 4; ++c)
                  {
                     Dxn[c] /= dx;
                     Dyn[c] /= dx;
                  }
               }
               if (A.char_lim)
               {
                  transform_to_con (em.Rx, Dxn);
                  transform_to_con (em.Ry, Dyn);
               }
               if (BASIS == BASIS_PK && A.cam) // limiter.cc:496-500
               {
                  Dyn[0] = 0.5 * (Dyn[0] - (ang_mom - Dxn[1]));
                  Dxn[1] = ang_mom + Dyn[0];
               }
#pragma unroll
               for (int c = 0; c < 4; ++c)
               {
                  sN[s * 8 + c] = Dxn[c];
                  sN[s * 8 + 4 + c] = Dyn[c];
               }
               sFlag[s] = 1.0;
               sFlag[CPB] = 1.0; // benign race: every writer stores the same value
            }
         }
         else if (p == 3)
         {
            // rewrite the cell as mean + limited linear part (limiter.cc:356-366 / 501-511)
            if (!active || !A.tvb || lq >= NS || sFlag[slot] == 0.0) return;
            double *uc = su + slot * D;
            if (BASIS == BASIS_QK)
            {
               const double *gx = t_gx (tb);
               const int a = lq % N1, b = lq / N1;
               const double x0 = A.geom[(size_t) cell * 4 + 0], y0 = A.geom[(size_t) cell * 4 + 1];
               const double hx = A.geom[(size_t) cell * 4 + 2], hy = A.geom[(size_t) cell * 4 + 3];
               const double dr0 = (x0 + gx[a] * hx) - (x0 + 0.5 * hx);
               const double dr1 = (y0 + gx[b] * hy) - (y0 + 0.5 * hy);
#pragma unroll
               for (int c = 0; c < 4; ++c)
                  uc[c * NS + lq] = A.avg[(size_t) cell * 4 + c] + dr0 * sN[slot * 8 + c] + dr1 * sN[slot * 8 + 4 + c];
            }
            else
            {
               const int m = lq;
               if (m == 0) return;
#pragma unroll
               for (int c = 0; c < 4; ++c)
               {
                  double v = 0.0;
                  if (m == 1) v = sN[slot * 8 + c] / 1.7320508075688772;
                  if (m == N1) v = sN[slot * 8 + 4 + c] / 1.7320508075688772;
                  uc[c * NS + m] = v;
               }
            }
         }
         else if (p == 4)
         {
            // positivity, density at the GLL x Gauss point sets (positivity.cc:68-78)
            if (!active || !A.pos_lim) return;
            const double *uc = su + slot * D;
            if (lq == 0)
            {
               double av[4];
#pragma unroll
               for (int c = 0; c < 4; ++c) av[c] = A.avg[(size_t) cell * 4 + c];
               if (std_min (av[RHO], pressure (av)) < eps) // positivity.cc:26-39
               {
#if defined(__CUDA_ARCH__)
                  atomicOr (A.err, (unsigned int) ERR_NEGATIVE_STATE);
#else
                  *A.err |= ERR_NEGATIVE_STATE;
#endif
               }
            }
            for (int idx = lq; idx < 2 * NPOS; idx += G)
               sP[slot * 2 * NPOS + idx] = point_value (tb, uc, idx / NPOS, idx % NPOS, RHO);
         }
         else if (p == 5)
         {
            if (!A.pos_lim || tid >= ncb) return;
            const int s = tid, cl = c0 + tid;
            double rho_min = 1.0e20;
            for (int i = 0; i < 2 * NPOS; ++i) rho_min = std_min (rho_min, sP[s * 2 * NPOS + i]);
            const double density_average = A.avg[(size_t) cl * 4 + RHO];
            const double rat = fabs (density_average - eps) / (fabs (density_average - rho_min) + 1.0e-13);
            const double theta1 = std_min (rat, 1.0);
            sT[s * 2] = theta1;
            if (theta1 < 1.0)
            {
               sFlag[s] += 2.0;
               sFlag[CPB] = 1.0;
            }
         }
         else if (p == 6)
         {
            // scale density about its mean (positivity.cc:85-110)
            if (!active || !A.pos_lim || lq >= NS) return;
            const double theta1 = sT[slot * 2];
            if (!(theta1 < 1.0)) return;
            double *uc = su + slot * D;
            if (BASIS == BASIS_QK)
               uc[RHO * NS + lq] = theta1 * uc[RHO * NS + lq] + (1.0 - theta1) * A.avg[(size_t) cell * 4 + RHO];
            else if (lq > 0)
               uc[RHO * NS + lq] *= theta1;
         }
         else if (p == 7)
         {
            // pressure at every point; where negative, the admissible fraction t towards the
            // mean (positivity.cc:134-179)
            if (!active || !A.pos_lim) return;
            const double *uc = su + slot * D;
            double av[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) av[c] = A.avg[(size_t) cell * 4 + c];
            for (int idx = lq; idx < 2 * NPOS; idx += G)
            {
               const int set = idx / NPOS, pt = idx % NPOS;
               const double mx = point_value (tb, uc, set, pt, 0);
               const double my = point_value (tb, uc, set, pt, 1);
               const double rho = point_value (tb, uc, set, pt, RHO);
               const double E = point_value (tb, uc, set, pt, ENE);
               const double pre = GM1 * (E - 0.5 * (mx * mx + my * my) / rho);
               double t = 1.0;
               if (pre < eps)
               {
                  const double drho = rho - av[RHO];
                  const double dm0 = mx - av[0], dm1 = my - av[1];
                  const double dE = E - av[ENE];
                  const double a1 = 2.0 * drho * dE - (dm0 * dm0 + dm1 * dm1);
                  double b1 = 2.0 * drho * (av[ENE] - eps / GM1) + 2.0 * av[RHO] * dE - 2.0 * (av[0] * dm0 + av[1] * dm1);
                  double c1 = 2.0 * av[RHO] * av[ENE] - (av[0] * av[0] + av[1] * av[1]) - 2.0 * eps * av[RHO] / GM1;
                  b1 /= a1;
                  c1 /= a1;
                  const double Dd = sqrt (fabs (b1 * b1 - 4.0 * c1));
                  const double t1 = 0.5 * (-b1 - Dd), t2 = 0.5 * (-b1 + Dd);
                  if (t1 > -1.0e-12 && t1 < 1.0 + 1.0e-12)
                     t = t1;
                  else if (t2 > -1.0e-12 && t2 < 1.0 + 1.0e-12)
                     t = t2;
                  else
                  {
                     t = 0.0;
#if defined(__CUDA_ARCH__)
                     atomicOr (A.err, (unsigned int) ERR_POSLIM_ROOT);
#else
                     *A.err |= ERR_POSLIM_ROOT;
#endif
                  }
                  t = std_min (1.0, t);
                  t = std_max (0.0, t);
                  if (fabs (1.0 - t) < 1.0e-14) t = 0.0;
               }
               sP[slot * 2 * NPOS + idx] = t;
            }
         }
         else if (p == 8)
         {
            if (!A.pos_lim || tid >= ncb) return;
            const int s = tid;
            double theta2 = 1.0;
            for (int i = 0; i < 2 * NPOS; ++i) theta2 = std_min (theta2, sP[s * 2 * NPOS + i]);
            sT[s * 2 + 1] = theta2;
            if (theta2 < 1.0)
            {
               sFlag[s] += 4.0;
               sFlag[CPB] = 1.0;
            }
         }
         else if (p == 9)
         {
            // scale all components about the mean (positivity.cc:182-206), then write back the
            // block only if some cell in it changed
            if (active && A.pos_lim && lq < NS)
            {
               const double theta2 = sT[slot * 2 + 1];
               if (theta2 < 1.0)
               {
                  double *uc = su + slot * D;
#pragma unroll
                  for (int c = 0; c < 4; ++c)
                  {
                     if (BASIS == BASIS_QK)
                        uc[c * NS + lq] = theta2 * uc[c * NS + lq] + (1.0 - theta2) * A.avg[(size_t) cell * 4 + c];
                     else if (lq > 0)
                        uc[c * NS + lq] *= theta2;
                  }
               }
            }
            if (active && lq == 0 && A.flags_out) A.flags_out[cell] = (int) sFlag[slot];
         }
         else // p == 10: write the block back only if some cell in it changed
         {
            if (sFlag[CPB] == 0.0) return;
            double *dst = A.u + (size_t) c0 * D;
            for (int i = tid; i < ncb * D; i += THREADS) dst[i] = su[i];
         }
      }
   };

   //---------------------------------------------------------------------------------------------
   // The same limiters as LimiterKernel, one THREAD per cell (thread_kernel form).  The block form
   // above spends most of its time in block barriers around steps that only one thread per cell
   // can do (the characteristic minmod decision, the theta reductions); here a thread walks
   // through all steps of its cell on its own, reading the cell's DoFs straight from global memory
   // (consecutive loads of a thread hit the L1 sectors its earlier loads brought in) and writing
   // only if the cell changes -- which on most cells of most stages it does not.  Every
   // arithmetic expression is the one of LimiterKernel, in the same order, so both forms give
   // bit-identical results.
   //---------------------------------------------------------------------------------------------
   template <int BASIS, int N1, int MINMAX = 0> // MINMAX = 1: the minmax limiter instead of TVB, an instantiation of its own
   struct LimiterCellKernel
   {
      typedef LimiterArgs Args;
      typedef LimiterKernel<BASIS, N1> LK;
      static constexpr int NS = LK::NS, D = LK::D, NGLL = LK::NGLL, NPOS = LK::NPOS;
      // block form: CPB cells are staged with coalesced loads into shared memory (odd row stride:
      // the per-thread walks over a cell are then free of bank conflicts), one thread per cell
      // works on its row, changed cells are written back.  A block is ONE warp (DFLO_LIM_THREADS = 32): the kernel is a
      // load phase followed by a long arithmetic phase, and many small blocks interleave the two better than a few large
      // ones -- measured on the Q3 limiter of cfg5: 128 threads x 3 blocks 113 us, 64 x 6 110 us, 32 x 12 103 us.
#ifndef DFLO_LIM_THREADS
#define DFLO_LIM_THREADS 32
#endif
      static constexpr int THREADS = DFLO_LIM_THREADS;
      static constexpr int CPB = THREADS;
      // register budget as resident WARPS per SM: up to Q2 / P3 sixteen (128 registers) -- except the P2 TVB + positivity
      // chain (cfg3), which spills 528 bytes at 128 registers -- and twelve (168 registers) up to Q3; beyond, what fits
#ifndef DFLO_LIM_P2_WARPS
#define DFLO_LIM_P2_WARPS 12
#endif
#ifndef DFLO_LIM_Q3_WARPS
#define DFLO_LIM_Q3_WARPS 12
#endif
      static constexpr int WARPS = (BASIS == BASIS_PK && N1 == 3 && MINMAX == 0) ? DFLO_LIM_P2_WARPS : D <= 40 ? 16 : D <= 64 ? DFLO_LIM_Q3_WARPS : 4;
      static constexpr int MIN_BLOCKS = WARPS * 32 / THREADS;
      static constexpr int NPHASE = 2;
      static constexpr int ROW = D + 1;
      static constexpr int SMEM_DOUBLES = CPB * ROW;
      static int grid (int n_compute) { return (n_compute + CPB - 1) / CPB; }

      static DFLO_DEV void phase (int p, const Args &A, double *sm, int tid, int bid)
      {
         if (N1 == 1) return; // degree 0: both limiters return at once
         const int c0 = bid * CPB;
         const int ncb = (A.n_compute - c0 < CPB) ? A.n_compute - c0 : CPB;
         if (p == 0)
         {
            const double *src = A.u + (size_t) c0 * D;
#if defined(__CUDA_ARCH__)
            // what the cell walk reads behind dependent addresses: towards L1 now
            if (tid < ncb)
            {
               asm volatile ("prefetch.global.L1 [%0];" ::"l"(A.geom + (size_t) (c0 + tid) * 4));
               asm volatile ("prefetch.global.L1 [%0];" ::"l"(A.avg + (size_t) (c0 + tid) * 4));
            }
#endif
#if defined(__CUDA_ARCH__)
            // up to Q2 / P3: asynchronous 8-byte copies straight into the padded rows (the odd row stride rules out wider
            // ones), the whole block in flight at once, no staging registers.  They allocate in L1; with the 66 KB blocks
            // of Q3 too little of it is left beside the shared memory (measured: 205 us instead of 110 on cfg5), so the
            // larger cells go through registers
            if (D <= 40)
            {
               const unsigned smb = (unsigned) __cvta_generic_to_shared (sm);
#pragma unroll 8
               for (int i = tid; i < ncb * D; i += THREADS)
                  asm volatile ("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smb + 8u * (unsigned) ((i / D) * ROW + (i % D))), "l"(src + i) : "memory");
            }
            else
#endif
            if (ncb == CPB) // full block: 16-byte loads, all of a round in flight
            {
               constexpr int NV = D / 2, UB = NV < 16 ? NV : 16; // D is even: NV double2 per cell, CPB * NV in the block
               const double2 *s2 = reinterpret_cast<const double2 *> (src);
#pragma unroll
               for (int k0 = 0; k0 < NV; k0 += UB)
               {
                  double2 v[UB];
#pragma unroll
                  for (int k = 0; k < UB; ++k)
                     if (k0 + k < NV) v[k] = s2[tid + (k0 + k) * THREADS];
#pragma unroll
                  for (int k = 0; k < UB; ++k)
                     if (k0 + k < NV)
                     {
                        const int i = 2 * (tid + (k0 + k) * THREADS);
                        double *d = sm + (i / D) * ROW + (i % D);
                        d[0] = v[k].x;
                        d[1] = v[k].y;
                     }
               }
            }
            else
               for (int i = tid; i < ncb * D; i += THREADS) sm[(i / D) * ROW + (i % D)] = src[i];
#if defined(__CUDA_ARCH__)
            if (tid < ncb && A.tvb) // the neighbours' means
            {
               const int4 nb4 = *reinterpret_cast<const int4 *> (A.nbr + (size_t) (c0 + tid) * 4);
               if (nb4.x >= 0) asm volatile ("prefetch.global.L1 [%0];" ::"l"(A.avg + (size_t) nb4.x * 4));
               if (nb4.y >= 0) asm volatile ("prefetch.global.L1 [%0];" ::"l"(A.avg + (size_t) nb4.y * 4));
               if (nb4.z >= 0) asm volatile ("prefetch.global.L1 [%0];" ::"l"(A.avg + (size_t) nb4.z * 4));
               if (nb4.w >= 0) asm volatile ("prefetch.global.L1 [%0];" ::"l"(A.avg + (size_t) nb4.w * 4));
            }
            asm volatile ("cp.async.wait_all;" ::: "memory");
#endif
         }
         else if (tid < ncb)
         {
            double *uc = sm + tid * ROW;
            if (cell_work (A, c0 + tid, uc))
            {
               double *dst = A.u + (size_t) (c0 + tid) * D;
               for (int i = 0; i < D; ++i) dst[i] = uc[i];
            }
         }
      }

      // apply_limiter_minmax_Qk of the MPI tree (src_mpi/limiter.cc:400-553) on one cell: slopes
      // scaled so that the linear reconstruction at the four face centres stays inside the range
      // of the neighbour means.  As in the reference, without characteristic limiting the range
      // starts from 0 (zero-initialised avg_min / avg_max, :438), and only true interior faces
      // bring neighbours (!at_boundary, :454: periodic partners do not).  The mean gradient is
      // taken with the Gauss(k+1) rule of the TVB limiter instead of the reference's QGauss(nq)
      // (:407-413) -- both are exact for the integrand.  Returns 1 if the cell was rewritten.
      struct MinmaxIn
      {
         const double *avg, *geom, *tab;
         const int *nbr;
         const unsigned char *fflags;
         double M;
         int char_lim;
      };
      static DFLO_DEV int minmax_cell (const MinmaxIn A, int cell, double *uc)
      {
         const double *tb = A.tab;
         const double hx = A.geom[(size_t) cell * 4 + 2], hy = A.geom[(size_t) cell * 4 + 3];
         const double dx = sqrt (hx * hx + hy * hy) / 1.4142135623730951;
         const double Mdx2 = A.M * dx * dx;
         double av[4], avc[4], amin[4] = {0.0, 0.0, 0.0, 0.0}, amax[4] = {0.0, 0.0, 0.0, 0.0}, Dx[4], Dy[4];
#pragma unroll
         for (int c = 0; c < 4; ++c) avc[c] = av[c] = A.avg[(size_t) cell * 4 + c];
         EigenStream es;
         if (A.char_lim)
         {
            compute_eigen_stream (av, es);
            transform_to_char (es.L, avc);
#pragma unroll
            for (int c = 0; c < 4; ++c) amin[c] = amax[c] = avc[c];
         }
         for (int f = 0; f < 4; ++f)
         {
            const int nb = A.nbr[(size_t) cell * 4 + f];
            if (nb < 0 || (A.fflags[(size_t) cell * 4 + f] & FACE_PERIODIC)) continue;
            double an[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) an[c] = A.avg[(size_t) nb * 4 + c];
            if (A.char_lim) transform_to_char (es.L, an);
#pragma unroll
            for (int c = 0; c < 4; ++c)
            {
               amin[c] = std_min (amin[c], an[c]);
               amax[c] = std_max (amax[c], an[c]);
            }
         }
         const double *gw = LK::t_gw (tb), *gd = LK::t_gdiff (tb);
#pragma unroll
         for (int c = 0; c < 4; ++c)
         {
            double sx = 0.0, sy = 0.0;
            for (int b = 0; b < N1; ++b)
               for (int a = 0; a < N1; ++a)
               {
                  const double u = uc[c * NS + a + N1 * b];
                  sx += (gd[a] * gw[b]) * u;
                  sy += (gw[a] * gd[b]) * u;
               }
            Dx[c] = sx / hx;
            Dy[c] = sy / hy;
         }
         if (A.char_lim)
         {
            transform_to_char (es.L, Dx);
            transform_to_char (es.L, Dy);
         }
         double theta[4] = {1.0, 1.0, 1.0, 1.0}, change = 0.0;
#pragma unroll
         for (int c = 0; c < 4; ++c)
         {
            const double dumin = amin[c] - avc[c], dumax = amax[c] - avc[c];
            if (dumax - dumin > Mdx2)
               for (int f = 0; f < 4; ++f)
               {
                  // face centre - cell centre (src_mpi/limiter.cc:500)
                  const double dr0 = (f == 0) ? -0.5 * hx : (f == 1) ? 0.5 * hx : 0.0;
                  const double dr1 = (f == 2) ? -0.5 * hy : (f == 3) ? 0.5 * hy : 0.0;
                  const double du = dr0 * Dx[c] + dr1 * Dy[c];
                  if (du > 0.0)
                     theta[c] = std_min (theta[c], dumax / du);
                  else if (du < 0.0)
                     theta[c] = std_min (theta[c], dumin / du);
               }
            change += theta[c];
         }
         change /= 4;
         if (!(change < 0.99)) return 0;
#pragma unroll
         for (int c = 0; c < 4; ++c)
         {
            Dx[c] *= theta[c];
            Dy[c] *= theta[c];
         }
         if (A.char_lim)
         {
            transform_to_con (es.R, Dx);
            transform_to_con (es.R, Dy);
         }
         const double *gx = LK::t_gx (tb);
         const double x0 = A.geom[(size_t) cell * 4 + 0], y0 = A.geom[(size_t) cell * 4 + 1];
         for (int b = 0; b < N1; ++b)
            for (int a = 0; a < N1; ++a)
            {
               const double dr0 = (x0 + gx[a] * hx) - (x0 + 0.5 * hx);
               const double dr1 = (y0 + gx[b] * hy) - (y0 + 0.5 * hy);
#pragma unroll
               for (int c = 0; c < 4; ++c) uc[c * NS + a + N1 * b] = av[c] + dr0 * Dx[c] + dr1 * Dy[c];
            }
         return 1;
      }

      // range of the nodal values for the positivity fast path.  (Tracking only max |value| for the momentum components
      // would save a quarter of these -- a std_min / std_max pair costs as much as five fp64 multiply-adds -- but the
      // bound |m| <= (wp - wm) max|m| it leads to is too loose for supersonic flow, where the kinetic energy is most of
      // E: on cfg5 every cell then fell through to the exact evaluation, 205 us instead of 110.)
      static DFLO_DEV void track (int c, double u, double (&lo)[4], double (&hi)[4])
      {
         lo[c] = std_min (lo[c], u);
         hi[c] = std_max (hi[c], u);
      }

      // all limiter steps of one cell on its DoFs uc (in place); returns true if they changed
      static DFLO_DEV bool cell_work (const Args &A, int cell, double *uc)
      {
         const double *tb = A.tab;
         const double eps = 1.0e-13;
         double av[4];
#pragma unroll
         for (int c = 0; c < 4; ++c) av[c] = A.avg[(size_t) cell * 4 + c];
         int flag = 0;
         // Qk: smallest / largest nodal value per component, kept up to date for the positivity
         // fast path below
         double lo[4] = {1.0e300, 1.0e300, 1.0e300, 1.0e300}, hi[4] = {-1.0e300, -1.0e300, -1.0e300, -1.0e300};
         bool have_bounds = false;

         if (MINMAX && BASIS == BASIS_QK && A.tvb == 2 && (!A.shock || A.shock[cell] > 1.0)) // src_mpi/limiter.cc:437
         {
            const typename LimiterCellKernel::MinmaxIn in = {A.avg, A.geom, A.tab, A.nbr, A.fflags, A.M, A.char_lim};
            flag |= minmax_cell (in, cell, uc);
         }
         if (!MINMAX && A.tvb == 1 && (!A.shock || A.shock[cell] > 1.0)) // limiter.cc:263, 406
         {
            const double hx = A.geom[(size_t) cell * 4 + 2], hy = A.geom[(size_t) cell * 4 + 3];
            const double dx = sqrt (hx * hx + hy * hy) / 1.4142135623730951; // diameter / sqrt(dim)
            const double Mdx2 = A.M * dx * dx;
            const double beta = (BASIS == BASIS_QK) ? A.beta : 0.5 * A.beta;
            double Dx[4], Dy[4], dbx[4], dfx[4], dby[4], dfy[4];
            // mean slopes of the cell, limiter.cc:268-281 (Qk) / 412-420 (Pk)
#pragma unroll
            for (int c = 0; c < 4; ++c)
            {
               double sx = 0.0, sy = 0.0;
               if (BASIS == BASIS_QK)
               {
                  const double *gw = LK::t_gw (tb), *gd = LK::t_gdiff (tb);
                  for (int b = 0; b < N1; ++b)
                     for (int a = 0; a < N1; ++a)
                     {
                        const double u = uc[c * NS + a + N1 * b];
                        sx += (gd[a] * gw[b]) * u;
                        sy += (gw[a] * gd[b]) * u;
                        if (A.pos_lim) track (c, u, lo, hi);
                     }
                  Dx[c] = dx * (sx / hx);
                  Dy[c] = dx * (sy / hy);
               }
               else
               {
                  Dx[c] = uc[c * NS + 1] * 1.7320508075688772; // sqrt(3); base index 1 / k+1
                  Dy[c] = uc[c * NS + N1] * 1.7320508075688772;
               }
               dbx[c] = dfx[c] = Dx[c];
               dby[c] = dfy[c] = Dy[c];
            }
            have_bounds = BASIS == BASIS_QK;
            const double ang_mom = Dx[1] - Dy[0];
            // lcell/rcell/bcell/tcell (claw.cc:357-379); periodic partners count as neighbours
            for (int f = 0; f < 4; ++f)
            {
               const int nb = A.nbr[(size_t) cell * 4 + f];
               if (nb < 0) continue;
#pragma unroll
               for (int c = 0; c < 4; ++c)
               {
                  const double an = A.avg[(size_t) nb * 4 + c];
                  if (f == 0) dbx[c] = av[c] - an;
                  if (f == 1) dfx[c] = an - av[c];
                  if (f == 2) dby[c] = av[c] - an;
                  if (f == 3) dfy[c] = an - av[c];
               }
            }
            if (A.char_lim)
            {
               EigenLeft el; // the right eigenvectors are needed only if the cell is actually limited
               compute_eigen_left (av, el);
               transform_to_char (el.Lx, Dx);
               transform_to_char (el.Ly, Dy);
               // TVB with M > 0: where every characteristic slope is below M dx^2 minmod returns it
               // unchanged (limiter.cc:20-21) whatever the neighbours are: nothing more to do
               bool small = true;
#pragma unroll
               for (int c = 0; c < 4; ++c) small = small && fabs (Dx[c]) < Mdx2 && fabs (Dy[c]) < Mdx2;
               if (!small)
               {
                  transform_to_char (el.Lx, dbx);
                  transform_to_char (el.Lx, dfx);
                  transform_to_char (el.Ly, dby);
                  transform_to_char (el.Ly, dfy);
               }
            }
            double Dxn[4], Dyn[4], change_x = 0.0, change_y = 0.0;
#pragma unroll
            for (int c = 0; c < 4; ++c)
            {
               Dxn[c] = minmod (Dx[c], beta * dbx[c], beta * dfx[c], Mdx2);
               Dyn[c] = minmod (Dy[c], beta * dby[c], beta * dfy[c], Mdx2);
               change_x += fabs (Dxn[c] - Dx[c]);
               change_y += fabs (Dyn[c] - Dy[c]);
            }
            change_x /= 4;
            change_y /= 4;
            if (change_x + change_y > 1.0e-10)
            {
               if (BASIS == BASIS_QK)
               {
#pragma unroll
                  for (int c = 0; c < 4; ++c)
                  {
                     Dxn[c] /= dx;
                     Dyn[c] /= dx;
                  }
               }
               if (A.char_lim)
               {
                  EigenRight er;
                  compute_eigen_right (av, er);
                  transform_to_con (er.Rx, Dxn);
                  transform_to_con (er.Ry, Dyn);
               }
               if (BASIS == BASIS_PK && A.cam) // limiter.cc:496-500
               {
                  Dyn[0] = 0.5 * (Dyn[0] - (ang_mom - Dxn[1]));
                  Dxn[1] = ang_mom + Dyn[0];
               }
               flag |= 1;
               // rewrite the cell as mean + limited linear part (limiter.cc:356-366 / 501-511)
               if (BASIS == BASIS_QK)
               {
                  const double *gx = LK::t_gx (tb);
                  const double x0 = A.geom[(size_t) cell * 4 + 0], y0 = A.geom[(size_t) cell * 4 + 1];
                  for (int b = 0; b < N1; ++b)
                     for (int a = 0; a < N1; ++a)
                     {
                        const double dr0 = (x0 + gx[a] * hx) - (x0 + 0.5 * hx);
                        const double dr1 = (y0 + gx[b] * hy) - (y0 + 0.5 * hy);
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                        {
                           const double v = av[c] + dr0 * Dxn[c] + dr1 * Dyn[c];
                           uc[c * NS + a + N1 * b] = v;
                           if (a + b == 0)
                           {
                              lo[c] = 1.0e300;
                              hi[c] = -1.0e300;
                           }
                           track (c, v, lo, hi);
                        }
                     }
               }
               else
               {
                  for (int m = 1; m < NS; ++m)
#pragma unroll
                     for (int c = 0; c < 4; ++c)
                     {
                        double v = 0.0;
                        if (m == 1) v = Dxn[c] / 1.7320508075688772;
                        if (m == N1) v = Dyn[c] / 1.7320508075688772;
                        uc[c * NS + m] = v;
                     }
               }
            }
         }

         if (A.pos_lim)
         {
            if (std_min (av[RHO], pressure (av)) < eps) // positivity.cc:26-39
            {
#if defined(__CUDA_ARCH__)
               atomicOr (A.err, (unsigned int) ERR_NEGATIVE_STATE);
#else
               *A.err |= ERR_NEGATIVE_STATE;
#endif
            }
            // Fast path.  Rigorous bounds on every component at every positivity point from the
            // cell's own coefficients (Qk: nodal range times the positive / negative weight sums of
            // the 1-D interpolation to the GLL points; Pk: mean +- sum |mode| max|basis|).  If even
            // the bounds keep density and pressure away from the thresholds by a margin far above
            // round-off, the exact evaluation below would return theta1 = theta2 = 1: nothing to do.
            {
               double vlo[4], vhi[4];
               if (BASIS == BASIS_QK)
               {
                  if (!have_bounds)
                     for (int m = 0; m < NS; ++m)
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                        {
                           track (c, uc[c * NS + m], lo, hi);
                        }
                  const double *gli = LK::t_gli (tb);
                  double wp = 1.0, wm = 0.0; // the Gauss direction of a point set is the identity
                  for (int j = 0; j < NGLL; ++j)
                  {
                     double p = 0.0, q = 0.0;
                     for (int a = 0; a < N1; ++a)
                     {
                        const double w = gli[j * N1 + a];
                        p += std_max (w, 0.0);
                        q += std_min (w, 0.0);
                     }
                     wp = std_max (wp, p);
                     wm = std_min (wm, q);
                  }
#pragma unroll
                  for (int c = 0; c < 4; ++c)
                  {
                     vlo[c] = wp * lo[c] + wm * hi[c];
                     vhi[c] = wp * hi[c] + wm * lo[c];
                  }
               }
               else
               {
                  const double *cmax = tb + LK::TAB; // max |phi_m| over the positivity points (pack_limiter_tables)
                  double spread[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
                  for (int m = 1; m < NS; ++m)
                  {
                     const double cm = cmax[m];
#pragma unroll
                     for (int c = 0; c < 4; ++c) spread[c] += cm * fabs (uc[c * NS + m]);
                  }
#pragma unroll
                  for (int c = 0; c < 4; ++c)
                  {
                     vlo[c] = uc[c * NS] - spread[c];
                     vhi[c] = uc[c * NS] + spread[c];
                  }
               }
               const double m0 = std_max (fabs (vlo[0]), fabs (vhi[0])), m1 = std_max (fabs (vlo[1]), fabs (vhi[1]));
               const double scale = 1.0 + std_max (fabs (vhi[ENE]), std_max (fabs (vhi[RHO]), std_max (m0, m1)));
               const double margin = 1.0e-9 * scale;
               if (vlo[RHO] > margin && GM1 * (vlo[ENE] - 0.5 * (m0 * m0 + m1 * m1) / vlo[RHO]) > margin)
               {
                  if (A.flags_out) A.flags_out[cell] = flag;
                  return flag != 0;
               }
            }
            // density at the GLL x Gauss point sets (positivity.cc:68-78)
            double rho_min = 1.0e20;
            for (int i = 0; i < 2 * NPOS; ++i) rho_min = std_min (rho_min, LK::point_value (tb, uc, i / NPOS, i % NPOS, RHO));
            const double rat = fabs (av[RHO] - eps) / (fabs (av[RHO] - rho_min) + 1.0e-13);
            const double theta1 = std_min (rat, 1.0);
            if (theta1 < 1.0)
            {
               flag += 2;
               // scale density about its mean (positivity.cc:85-110)
               for (int m = 0; m < NS; ++m)
               {
                  if (BASIS == BASIS_QK)
                     uc[RHO * NS + m] = theta1 * uc[RHO * NS + m] + (1.0 - theta1) * av[RHO];
                  else if (m > 0)
                     uc[RHO * NS + m] *= theta1;
               }
            }
            // pressure at every point; where negative, the admissible fraction t towards the mean
            // (positivity.cc:134-179)
            double theta2 = 1.0;
            for (int idx = 0; idx < 2 * NPOS; ++idx)
            {
               const int set = idx / NPOS, pt = idx % NPOS;
               const double mx = LK::point_value (tb, uc, set, pt, 0);
               const double my = LK::point_value (tb, uc, set, pt, 1);
               const double rho = LK::point_value (tb, uc, set, pt, RHO);
               const double E = LK::point_value (tb, uc, set, pt, ENE);
               const double pre = GM1 * (E - 0.5 * (mx * mx + my * my) / rho);
               double t = 1.0;
               if (pre < eps)
               {
                  const double drho = rho - av[RHO];
                  const double dm0 = mx - av[0], dm1 = my - av[1];
                  const double dE = E - av[ENE];
                  const double a1 = 2.0 * drho * dE - (dm0 * dm0 + dm1 * dm1);
                  double b1 = 2.0 * drho * (av[ENE] - eps / GM1) + 2.0 * av[RHO] * dE - 2.0 * (av[0] * dm0 + av[1] * dm1);
                  double c1 = 2.0 * av[RHO] * av[ENE] - (av[0] * av[0] + av[1] * av[1]) - 2.0 * eps * av[RHO] / GM1;
                  b1 /= a1;
                  c1 /= a1;
                  const double Dd = sqrt (fabs (b1 * b1 - 4.0 * c1));
                  const double t1 = 0.5 * (-b1 - Dd), t2 = 0.5 * (-b1 + Dd);
                  if (t1 > -1.0e-12 && t1 < 1.0 + 1.0e-12)
                     t = t1;
                  else if (t2 > -1.0e-12 && t2 < 1.0 + 1.0e-12)
                     t = t2;
                  else
                  {
                     t = 0.0;
#if defined(__CUDA_ARCH__)
                     atomicOr (A.err, (unsigned int) ERR_POSLIM_ROOT);
#else
                     *A.err |= ERR_POSLIM_ROOT;
#endif
                  }
                  t = std_min (1.0, t);
                  t = std_max (0.0, t);
                  if (fabs (1.0 - t) < 1.0e-14) t = 0.0;
               }
               theta2 = std_min (theta2, t);
            }
            if (theta2 < 1.0)
            {
               flag += 4;
               // scale all components about the mean (positivity.cc:182-206)
               for (int m = 0; m < NS; ++m)
#pragma unroll
                  for (int c = 0; c < 4; ++c)
                  {
                     if (BASIS == BASIS_QK)
                        uc[c * NS + m] = theta2 * uc[c * NS + m] + (1.0 - theta2) * av[c];
                     else if (m > 0)
                        uc[c * NS + m] *= theta2;
                  }
            }
         }
         if (A.flags_out) A.flags_out[cell] = flag;
         return flag != 0;
      }
   };

   //---------------------------------------------------------------------------------------------
   // KXRCF shock indicator, compute_shock_indicator_kxrcf (indicator.cc:50-198), one thread per
   // cell: the jump of density (or energy) over the inflow part of the cell's interior faces,
   // normalised by diameter^((k+1)/2) x inflow measure x cell mean.  Reads the freshly updated
   // solution of the cell and of its face neighbours BEFORE any cell is limited, which is why it
   // is its own pass (the reference, too, fills shock_indicator for all cells first, claw.cc:763).
   //---------------------------------------------------------------------------------------------
   struct IndicatorArgs
   {
      const double *u, *avg;
      const int *nbr;
      const unsigned char *fflags;
      const double *geom;
      const double *tab;   // flat stage tables (end-point values / face tables, Gauss weights)
      double *shock;
      int n_cells, component;
   };

   template <int BASIS, int N1>
   struct IndicatorKernel
   {
      typedef IndicatorArgs Args;
      typedef StageKernel<BASIS, N1, FLUX_LXF> SK;
      static constexpr int NS = SK::NS, D = SK::D;

      // component c of cell ucell at point q of its face f (the fma chain of StageKernel::trace)
      static DFLO_DEV double trace1 (const double *tb, const double *ucell, int f, int q, int c)
      {
         double s = 0.0;
         if (BASIS == BASIS_QK)
         {
            const double *e = SK::t_e (tb, f & 1);
            const int base = (f < 2) ? N1 * q : q, stride = (f < 2) ? 1 : N1;
            for (int a = 0; a < N1; ++a) s = fma (e[a], ucell[c * NS + base + a * stride], s);
         }
         else
         {
            const double *pf = SK::t_phiface (tb) + (f * N1 + q) * NS;
            for (int m = 0; m < NS; ++m) s = fma (pf[m], ucell[c * NS + m], s);
         }
         return s;
      }

      static DFLO_DEV void thread (const Args &A, int cell)
      {
         if (cell >= A.n_cells) return;
         const double *tb = A.tab;
         const double *gw = SK::t_gw (tb);
         const double hx = A.geom[(size_t) cell * 4 + 2], hy = A.geom[(size_t) cell * 4 + 3];
         const double *av = A.avg + (size_t) cell * 4;
         const double vel[2] = {av[0] / av[RHO], av[1] / av[RHO]};                       // :108-110
         double ind = 0.0, inflow_measure = 0.0;
         for (int f = 0; f < 4; ++f)
         {
            const int nb = A.nbr[(size_t) cell * 4 + f];
            const int fl = A.fflags[(size_t) cell * 4 + f];
            if (nb < 0 || (fl & FACE_PERIODIC)) continue;                                // at_boundary(f): nothing, :181-186
            const double nx = (f == 0) ? -1.0 : (f == 1) ? 1.0 : 0.0, ny = (f == 2) ? -1.0 : (f == 3) ? 1.0 : 0.0;
            const int inflow_status = (vel[0] * nx + vel[1] * ny < 0);                   // :125
            const double len = (f < 2) ? hy : hx;
            for (int q = 0; q < N1; ++q)
            {
               const double own = trace1 (tb, A.u + (size_t) cell * D, f, q, A.component);
               const double nbv = trace1 (tb, A.u + (size_t) nb * D, f ^ 1, q, A.component);
               const double jxw = gw[q] * len;
               ind += inflow_status * (own - nbv) * jxw;
               inflow_measure += inflow_status * jxw;
            }
         }
         const double diameter = sqrt (hx * hx + hy * hy);
         const double denominator = pow (diameter, 0.5 * N1) * inflow_measure * av[A.component]; // :189-194
         A.shock[cell] = fabs (ind) / denominator;
      }
   };

   //---------------------------------------------------------------------------------------------
   // compute_time_step_cartesian, claw.cc:484-511: per-cell dt from the cell averages
   //---------------------------------------------------------------------------------------------
   DFLO_DEV double cell_time_step (const double *avg, const double *geom, double cfl, int degree)
   {
      const double hx = geom[2], hy = geom[3];
      const double h = sqrt (hx * hx + hy * hy) / 1.4142135623730951;
      const double sonic = sound_speed (avg);
      const double density = avg[RHO];
      double max_eigenvalue = 0.0;
      max_eigenvalue += (sonic + fabs (avg[0] / density)) / h;
      max_eigenvalue += (sonic + fabs (avg[1] / density)) / h;
      return cfl / max_eigenvalue / (2.0 * degree + 1.0);
   }
}
