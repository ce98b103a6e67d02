// Register-blocked stage kernel for the Qk basis on sm_100a (CUDA only).
//
// Same arithmetic as StageKernel<BASIS_QK,...> in kernels.cuh -- assemble_system
// (reference src/assemble_explicit.cc:30-452), M^-1 and the RK combine (src/claw.cc:694-713,
// 757-760) and compute_cell_average (src/claw.cc:562-597) in one launch -- but organised around
// registers instead of shared memory:
//
//   * a tile cell is worked on by N1 = k+1 threads.  Thread l first holds Gauss ROW l of the cell
//     (all N1 nodes, 4 components) and does everything that couples along x: F_x at its nodes, the
//     contraction with the 1-D derivative matrix, the left/right traces, the right-face Riemann
//     problem, the x lifting.  Then it holds Gauss COLUMN l of a cell and does the same along y.
//     The dense i x q loops of the reference become fully unrolled FMA chains on registers whose
//     table operands (D.w, l_a(0), l_a(1), w_a) are constant-bank immediates.
//   * every face is solved once per tile, along +e_x / +e_y, by the thread that owns the low-side
//     cell's row/column (row_desc.h); tile-edge low faces and the traces of halo cells are the job
//     of one extra warp.  Neighbouring threads exchange one 32-byte trace and one 32-byte flux
//     per face point through shared memory; nothing else is shared except the y part of the
//     residual, which is transposed back to row order through one padded buffer.
//   * the tile and its halo cells arrive by per-cell bulk async copies (TMA) into a padded layout
//     (cell stride D+2 doubles) that makes both the row reads (128-bit) and the column reads
//     (64-bit, even/odd cell interleave for N1 = 4) free of bank conflicts; old_solution is
//     prefetched into L2 by one bulk prefetch and read straight into registers at the end.
//
// Per Q3 cell this is ~8 k SASS thread-instructions (4.0 k fp64) against 11.5 k (4.1 k) in the phase
// kernel, and a third of its shared-memory wavefronts.
#pragma once

#include "kernels.cuh"
#include "p2p_halo.cuh"
#include "row_desc.h"

namespace dflo
{
   struct RowConst
   {
      double dw[MAX_N1][MAX_N1]; // dw[ap][a] = l_a'(x_ap) w_ap
      double e0[MAX_N1], e1[MAX_N1], gw[MAX_N1];
   };
   __constant__ RowConst c_row[MAX_N1 + 1]; // indexed by N1

   template <int N1, int FLUX>
   struct RowShape
   {
      static constexpr int NS = N1 * N1, D = 4 * NS;
      static constexpr int TC = row_tc (N1), NH = row_nh (N1);
      static constexpr int CS = D + 2;                      // padded cell stride in shared memory (doubles)
      #ifndef DFLO_ROW_EXTRA
#define DFLO_ROW_EXTRA 32
#endif
      static constexpr int MAIN = TC * N1, EXTRA = DFLO_ROW_EXTRA, THREADS = MAIN + EXTRA;
      #ifndef DFLO_ROW_MIN_BLOCKS
#define DFLO_ROW_MIN_BLOCKS 4
#endif
#ifndef DFLO_ROW_MIN_BLOCKS_N3
#define DFLO_ROW_MIN_BLOCKS_N3 6
#endif
      static constexpr int MIN_BLOCKS = N1 == 3 ? DFLO_ROW_MIN_BLOCKS_N3 : N1 <= 4 ? DFLO_ROW_MIN_BLOCKS : 2;
      static constexpr int DESC_INTS = rowd_ints (TC, NH);
      static constexpr int OFF_HALO = rowd_off_halo (), OFF_NBHI = rowd_off_nbhi (NH), OFF_LJOB = rowd_off_ljob (TC, NH),
                           OFF_GJOB = rowd_off_gjob (TC, NH);
      // shared memory carve-up in doubles; every bulk-copy destination is 16-byte aligned.
      //   [O_U, O_X)   the tile's cells (padded), live to the end
      //   [O_X, ...)   halo cells | ghost traces | low-face traces -> fluxes: dead once the fluxes
      //                have been read into registers; the y part of the residual (sR, row order,
      //                padded like su) is then written over this region, and the partial cell
      //                averages over its tail
      static constexpr int O_U = 2;                               // [0,2): mbarrier
      static constexpr int O_X = O_U + TC * CS;
      static constexpr int O_G = O_X + NH * CS;                   // ghost traces [NH][N1][4]
      static constexpr int O_T = O_G + NH * N1 * 4;               // 2 halves x [2 dirs][TC][N1] double2
      static constexpr int T_HALF = 2 * TC * N1 * 2;
      static constexpr int X_END_A = O_T + 2 * T_HALF;
      static constexpr int O_R = O_X;
      static constexpr int O_PART = O_R + TC * CS;                // [TC][N1][4]
      static constexpr int X_END_B = O_PART + TC * N1 * 4;
      static constexpr int O_GEOM = X_END_A > X_END_B ? X_END_A : X_END_B;   // x0 y0 hx hy of the tile cells
      static constexpr int O_AVG = O_GEOM + TC * 4;               // cell averages tile + halo (LxF only)
      static constexpr int O_DESC = O_AVG + (flux_uses_averages (FLUX) ? (TC + NH) * 4 : 0);
      static constexpr int SMEM_DOUBLES = O_DESC + DESC_INTS / 2;
      static_assert (MAIN % 32 == 0, "whole warps");
      static_assert (THREADS >= TC + NH + 2, "one staging copy per thread");
      static_assert (CS % 2 == 0 && DESC_INTS % 4 == 0, "16-byte aligned bulk copies");
   };

   // which (cell slot, line) a main thread works on in its column role, and the inverse (the
   // thread position of (slot, line)), used to index the bottom-face trace/flux array so that
   // those accesses are linear in the thread index.  N1 = 4: the two half-warps of a warp take
   // the even and the odd cells of the warp's 8 cells -- with the padded stride CS the 16 threads
   // of a 64-bit shared-memory access then fall on 32 distinct banks.
   template <int N1>
   __device__ __forceinline__ void col_role (int tid, int &slot, int &line)
   {
      if (N1 == 4)
      {
         const int lane = tid & 31, j = lane & 15;
         slot = (tid >> 5) * 8 + 2 * (j >> 2) + (lane >> 4);
         line = j & 3;
      }
      else
      {
         slot = tid / N1;
         line = tid % N1;
      }
   }
   template <int N1>
   __device__ __forceinline__ int col_pos (int slot, int line)
   {
      if (N1 == 4)
      {
         const int cw = slot & 7;
         return (slot >> 3) * 32 + (cw & 1) * 16 + (cw >> 1) * 4 + line;
      }
      return slot * N1 + line;
   }

   // N1 consecutive doubles (16-byte aligned when N1 is even)
   template <int N1>
   __device__ __forceinline__ void load_line (const double *p, double v[N1])
   {
      if (N1 % 2 == 0)
      {
#pragma unroll
         for (int a = 0; a < N1; a += 2)
         {
            const double2 t = *reinterpret_cast<const double2 *> (p + a);
            v[a] = t.x;
            v[a + 1] = t.y;
         }
      }
      else
      {
#pragma unroll
         for (int a = 0; a < N1; ++a) v[a] = p[a];
      }
   }
   template <int N1>
   __device__ __forceinline__ void store_line (double *p, const double v[N1])
   {
      if (N1 % 2 == 0)
      {
#pragma unroll
         for (int a = 0; a < N1; a += 2) *reinterpret_cast<double2 *> (p + a) = make_double2 (v[a], v[a + 1]);
      }
      else
      {
#pragma unroll
         for (int a = 0; a < N1; ++a) p[a] = v[a];
      }
   }

   // a 4-component point value stored as two double2 halves T_HALF apart
   template <int T_HALF>
   __device__ __forceinline__ void load_pt (const double *sT, int p, double W[4])
   {
      const double2 a = *reinterpret_cast<const double2 *> (sT + 2 * p);
      const double2 b = *reinterpret_cast<const double2 *> (sT + T_HALF + 2 * p);
      W[0] = a.x;
      W[1] = a.y;
      W[2] = b.x;
      W[3] = b.y;
   }
   template <int T_HALF>
   __device__ __forceinline__ void store_pt (double *sT, int p, const double W[4])
   {
      *reinterpret_cast<double2 *> (sT + 2 * p) = make_double2 (W[0], W[1]);
      *reinterpret_cast<double2 *> (sT + T_HALF + 2 * p) = make_double2 (W[2], W[3]);
   }
   __device__ __forceinline__ void load4 (const double *p, double W[4])
   {
      const double2 a = *reinterpret_cast<const double2 *> (p);
      const double2 b = *reinterpret_cast<const double2 *> (p + 2);
      W[0] = a.x;
      W[1] = a.y;
      W[2] = b.x;
      W[3] = b.y;
   }
   __device__ __forceinline__ void store4 (double *p, const double W[4])
   {
      *reinterpret_cast<double2 *> (p) = make_double2 (W[0], W[1]);
      *reinterpret_cast<double2 *> (p + 2) = make_double2 (W[2], W[3]);
   }

   // Trace of a staged cell (padded layout) at point q of its face f -- the fma chain of
   // StageKernel::trace, so that a halo cell's trace is bit-identical to the one its own tile forms.
   template <int N1>
   __device__ __forceinline__ void row_trace (const double *ucell, int f, int q, double W[4])
   {
      constexpr int NS = N1 * N1;
      const RowConst &T = c_row[N1];
      const int base = (f < 2) ? N1 * q : q;
      const int stride = (f < 2) ? 1 : N1;
#pragma unroll
      for (int c = 0; c < 4; ++c)
      {
         double s = 0.0;
#pragma unroll
         for (int a = 0; a < N1; ++a) s = fma ((f & 1) ? T.e1[a] : T.e0[a], ucell[c * NS + base + a * stride], s);
         W[c] = s;
      }
   }

   // face flux with the direction known only at run time (extra warps): exchange the momentum
   // components around the x form
   template <int FLUX>
   __device__ __forceinline__ void face_flux_rt (int dir, bool plus_low, const double Wlo[4], const double Whi[4], const double Alo[4],
                                                 const double Ahi[4], double H[4])
   {
      double L[4], R[4], AL[4], AR[4], G[4];
      L[0] = dir ? Wlo[1] : Wlo[0];
      L[1] = dir ? Wlo[0] : Wlo[1];
      R[0] = dir ? Whi[1] : Whi[0];
      R[1] = dir ? Whi[0] : Whi[1];
      L[2] = Wlo[2];
      L[3] = Wlo[3];
      R[2] = Whi[2];
      R[3] = Whi[3];
      if (flux_uses_averages (FLUX))
      {
         AL[0] = dir ? Alo[1] : Alo[0];
         AL[1] = dir ? Alo[0] : Alo[1];
         AR[0] = dir ? Ahi[1] : Ahi[0];
         AR[1] = dir ? Ahi[0] : Ahi[1];
         AL[2] = Alo[2];
         AL[3] = Alo[3];
         AR[2] = Ahi[2];
         AR[3] = Ahi[3];
      }
      face_flux_axis<FLUX, 0> (plus_low, L, R, AL, AR, G);
      H[0] = dir ? G[1] : G[0];
      H[1] = dir ? G[0] : G[1];
      H[2] = G[2];
      H[3] = G[3];
   }

   template <int N1, int FLUX>
   __global__ void __launch_bounds__ (RowShape<N1, FLUX>::THREADS, RowShape<N1, FLUX>::MIN_BLOCKS) row_stage_kernel (const StageArgs A)
   {
      typedef RowShape<N1, FLUX> S;
      constexpr int NS = S::NS, D = S::D, TC = S::TC, CS = S::CS, TH = S::T_HALF;
      extern __shared__ __align__ (16) double sm[];
      double *su = sm + S::O_U, *sR = sm + S::O_R, *sT = sm + S::O_T, *sG = sm + S::O_G, *sGeom = sm + S::O_GEOM, *sAvg = sm + S::O_AVG;
      int *sdesc = reinterpret_cast<int *> (sm + S::O_DESC);
      const RowConst &T = c_row[N1];
      const int tid = threadIdx.x;
      const int *gdesc = A.rowdesc + (size_t) blockIdx.x * S::DESC_INTS;
      pdl_launch_dependents (); // the next kernel on the stream may be scheduled into the tail of this one
      const int c0 = gdesc[0], ncb = gdesc[1], nh = gdesc[2];
      // halo thread i = tid - TC: its cell id is requested together with the header, not after it (one trip less
      // in front of the copies); entries past nh are inside the descriptor and unused
      int halo_cell = 0;
      if (tid >= TC && tid < TC + S::NH) halo_cell = gdesc[S::OFF_HALO + tid - TC];
      const bool need_old = A.mode == MODE_STAGE && A.ark != 0.0;
      constexpr unsigned cell_bytes = (unsigned) (D * sizeof (double));
      // launched into the tail of the previous stage kernel: its output is this kernel's input (the descriptor words
      // above are static data)
      if (A.pdl == 2) pdl_wait ();

      // ---- stage the tile: per-cell bulk copies into the padded layout, descriptor, geometry ----
      if (tid == 0)
      {
         if (A.fx && gdesc[5])
         {
            // fused halo exchange: this tile reads ghost cells -- the peers' stores of the previous
            // exchange must have landed (boundary tiles run first, so this rarely spins)
            const P2PFused &F = *A.fx;
            // epochs[0] = number of the NEXT exchange to publish; the last published one is needed here
            const unsigned long long e = *reinterpret_cast<volatile unsigned long long *> (F.epochs) - 1;
            for (int p = 0; p < F.npeers; ++p)
               while (ld_acquire_sys (F.my_flags + F.world + F.peer_rank[p]) < e) {}
         }
         unsigned bytes = (unsigned) (ncb + nh) * cell_bytes + (unsigned) ncb * 32u + (unsigned) (S::DESC_INTS * sizeof (int));
         if (flux_uses_averages (FLUX)) bytes += (unsigned) (ncb + nh) * 32u;
         mbar_init (sm, 1);
         mbar_expect_tx (sm, bytes);
      }
      __syncthreads ();
      if (tid < ncb || (tid >= TC && tid - TC < nh))
      {
         const int cell = tid < TC ? c0 + tid : halo_cell;
         const int slot = tid;
         bulk_g2s (su + slot * CS, A.u + (size_t) cell * D, cell_bytes, sm);
         if (flux_uses_averages (FLUX)) bulk_g2s (sAvg + slot * 4, A.avg + (size_t) cell * 4, 32u, sm);
      }
      else if (tid == S::THREADS - 1)
      {
         bulk_g2s (sdesc, gdesc, (unsigned) (S::DESC_INTS * sizeof (int)), sm);
         bulk_g2s (sGeom, A.geom + (size_t) c0 * 4, (unsigned) ncb * 32u, sm);
         if (need_old)
            asm volatile ("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(A.u_old + (size_t) c0 * D), "r"((unsigned) ncb * cell_bytes) : "memory");
      }
      else if (tid == S::THREADS - 2)
      {
         // Pull the inputs of the tile that will run in this block's slot a wave later from HBM
         // into L2 now (a hint: on a uniform tiling that tile starts pf_tiles * TC cells further on)
         const int bt = (int) blockIdx.x + A.pf_tiles;
         if (bt < (int) gridDim.x)
         {
            const size_t pc0 = (size_t) c0 + (size_t) A.pf_tiles * TC;
            if (pc0 + TC <= (size_t) A.n_cells_u)
            {
               asm volatile ("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(A.u + pc0 * D), "r"((unsigned) TC * cell_bytes) : "memory");
               if (need_old) asm volatile ("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(A.u_old + pc0 * D), "r"((unsigned) TC * cell_bytes) : "memory");
            }
            asm volatile ("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(A.rowdesc + (size_t) bt * S::DESC_INTS), "r"((unsigned) (S::DESC_INTS * sizeof (int))) : "memory");
         }
      }
      mbar_wait (sm, 0);

      const bool main_thread = tid < S::MAIN;
      // row role: cell slot rs, Gauss row rb;  column role: cell slot cs, Gauss column ca
      const int rs = tid / N1, rb = tid % N1;
      int cs, ca;
      col_role<N1> (tid, cs, ca);
      const bool row_on = main_thread && rs < ncb, col_on = main_thread && cs < ncb;
      const int pT_row = rs * N1 + rb;                              // low-face slots of the own cell
      const int pT_col = TC * N1 + col_pos<N1> (cs, ca);

      double WR[4], WT[4]; // own traces on the right / top face, later the fluxes there (along +e)

      // ================= P1: traces =================
      if (row_on)
      {
         const double *uc = su + rs * CS + rb * N1;
         double WL[4];
#pragma unroll
         for (int c = 0; c < 4; ++c)
         {
            double u[N1];
            load_line<N1> (uc + c * NS, u);
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int a = 0; a < N1; ++a)
            {
               s0 = fma (T.e0[a], u[a], s0);
               s1 = fma (T.e1[a], u[a], s1);
            }
            WL[c] = s0;
            WR[c] = s1;
         }
         store_pt<TH> (sT, pT_row, WL);
      }
      if (col_on)
      {
         const double *uc = su + cs * CS + ca;
         double WB[4];
#pragma unroll
         for (int c = 0; c < 4; ++c)
         {
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int b = 0; b < N1; ++b)
            {
               const double u = uc[c * NS + b * N1];
               s0 = fma (T.e0[b], u, s0);
               s1 = fma (T.e1[b], u, s1);
            }
            WB[c] = s0;
            WT[c] = s1;
         }
         store_pt<TH> (sT, pT_col, WB);
      }
      if (!main_thread)
      {
         // G jobs: low-face traces of the cells beyond the tile's high edges
         const int nG = (A.dbg & 4) ? 0 : sdesc[4];
         for (int j = tid - S::MAIN; j < nG * N1; j += S::EXTRA)
         {
            const int g = j / N1, q = j % N1;
            const int code = sdesc[S::OFF_GJOB + g];
            const int slot = code & 0xffff, dir = (code >> 16) & 1;
            const int qq = (code & ROWD_FLIP) ? N1 - 1 - q : q;
            double W[4];
            row_trace<N1> (su + slot * CS, 2 * dir, qq, W);
            store4 (sG + j * 4, W);
         }
      }
      __syncthreads ();

      // ================= P2: Riemann problems =================
      if (row_on)
      {
         const int code = sdesc[S::OFF_NBHI + 2 * rs];
         double Wn[4], Ao[4], An[4], H[4];
         if (flux_uses_averages (FLUX)) load4 (sAvg + rs * 4, Ao);
         if (code >= 0)
         {
            const int idx = code & 0xffff;
            if (idx < TC)
               load_pt<TH> (sT, idx * N1 + rb, Wn);
            else
               load4 (sG + ((idx - TC) * N1 + rb) * 4, Wn);
            if (flux_uses_averages (FLUX)) load4 (sAvg + (idx < TC ? idx : (sdesc[S::OFF_GJOB + idx - TC] & 0xffff)) * 4, An);
            if (A.dbg & 1)
            {
#pragma unroll
               for (int c = 0; c < 4; ++c) H[c] = WR[c] + Wn[c];
            }
            else
               face_flux_axis<FLUX, 0> ((code & ROWD_PLUS) != 0, WR, Wn, Ao, An, H);
            if (idx < TC) store_pt<TH> (sT, idx * N1 + rb, H);
         }
         else
         {
            // physical boundary on the right: W- from the boundary condition (assemble_explicit.cc:176-206)
            const int bf = -1 - code;
            const int kind = A.bkind[bf];
            double g[4];
            load4 (A.bc_g + ((size_t) bf * N1 + rb) * 4, g);
            compute_wminus (kind, 1.0, 0.0, WR, g, Wn);
            if (flux_uses_averages (FLUX))
            {
               if (A.compat_mpi)
                  compute_wminus (kind, 1.0, 0.0, Ao, g, An);
               else
               {
#pragma unroll
                  for (int c = 0; c < 4; ++c) An[c] = Ao[c];
               }
            }
            face_flux_axis<FLUX, 0> (true, WR, Wn, Ao, An, H);
         }
#pragma unroll
         for (int c = 0; c < 4; ++c) WR[c] = H[c];
      }
      if (col_on)
      {
         const int code = sdesc[S::OFF_NBHI + 2 * cs + 1];
         double Wn[4], Ao[4], An[4], H[4];
         if (flux_uses_averages (FLUX)) load4 (sAvg + cs * 4, Ao);
         if (code >= 0)
         {
            const int idx = code & 0xffff;
            if (idx < TC)
               load_pt<TH> (sT, TC * N1 + col_pos<N1> (idx, ca), Wn);
            else
               load4 (sG + ((idx - TC) * N1 + ca) * 4, Wn);
            if (flux_uses_averages (FLUX)) load4 (sAvg + (idx < TC ? idx : (sdesc[S::OFF_GJOB + idx - TC] & 0xffff)) * 4, An);
            if (A.dbg & 1)
            {
#pragma unroll
               for (int c = 0; c < 4; ++c) H[c] = WT[c] + Wn[c];
            }
            else
               face_flux_axis<FLUX, 1> ((code & ROWD_PLUS) != 0, WT, Wn, Ao, An, H);
            if (idx < TC) store_pt<TH> (sT, TC * N1 + col_pos<N1> (idx, ca), H);
         }
         else
         {
            const int bf = -1 - code;
            const int kind = A.bkind[bf];
            double g[4];
            load4 (A.bc_g + ((size_t) bf * N1 + ca) * 4, g);
            compute_wminus (kind, 0.0, 1.0, WT, g, Wn);
            if (flux_uses_averages (FLUX))
            {
               if (A.compat_mpi)
                  compute_wminus (kind, 0.0, 1.0, Ao, g, An);
               else
               {
#pragma unroll
                  for (int c = 0; c < 4; ++c) An[c] = Ao[c];
               }
            }
            face_flux_axis<FLUX, 1> (true, WT, Wn, Ao, An, H);
         }
#pragma unroll
         for (int c = 0; c < 4; ++c) WT[c] = H[c];
      }
      if (!main_thread)
      {
         // L jobs: low faces of tile cells whose neighbour is a halo cell, a periodic partner or a boundary
         const int nL = (A.dbg & 4) ? 0 : sdesc[3];
         for (int j = tid - S::MAIN; j < nL * N1; j += S::EXTRA)
         {
            const int job = j / N1, q = j % N1;
            const int w0 = sdesc[S::OFF_LJOB + 2 * job], nb = sdesc[S::OFF_LJOB + 2 * job + 1];
            const int s = (w0 & 0xffff) >> 1, dir = w0 & 1;
            const bool plus_own = (w0 & ROWD_PLUS) != 0;
            const int p = dir ? TC * N1 + col_pos<N1> (s, q) : s * N1 + q;
            double Wo[4], Wn[4], Ao[4], An[4], H[4];
            load_pt<TH> (sT, p, Wo);
            if (flux_uses_averages (FLUX)) load4 (sAvg + s * 4, Ao);
            if (nb >= 0)
            {
               const int qn = (w0 & ROWD_FLIP) ? N1 - 1 - q : q;
               row_trace<N1> (su + nb * CS, 2 * dir + 1, qn, Wn);
               if (flux_uses_averages (FLUX)) load4 (sAvg + nb * 4, An);
            }
            else
            {
               const int bf = -1 - nb;
               const int kind = A.bkind[bf];
               double g[4];
               load4 (A.bc_g + ((size_t) bf * N1 + q) * 4, g);
               const double nx = dir ? 0.0 : -1.0, ny = dir ? -1.0 : 0.0;
               compute_wminus (kind, nx, ny, Wo, g, Wn);
               if (flux_uses_averages (FLUX))
               {
                  if (A.compat_mpi)
                     compute_wminus (kind, nx, ny, Ao, g, An);
                  else
                  {
#pragma unroll
                     for (int c = 0; c < 4; ++c) An[c] = Ao[c];
                  }
               }
            }
            face_flux_rt<FLUX> (dir, !plus_own, Wn, Wo, An, Ao, H);
            store_pt<TH> (sT, p, H);
         }
      }
      __syncthreads ();

      // ================= P3: volume terms and lifting =================
      // the low-face fluxes move to registers; after the barrier the exchange arrays are dead and
      // sR is written over them
      double HL[4], HB[4];
      if (row_on) load_pt<TH> (sT, pT_row, HL);
      if (col_on) load_pt<TH> (sT, pT_col, HB);
      __syncthreads ();
      double rrow[4][N1], uold[4][N1];
      if (col_on)
      {
         // y part: F_y at the column's nodes, contraction with D.w along y, top/bottom lifting
         const double hx = sGeom[cs * 4 + 2];
         const double *uc = su + cs * CS + ca;
         double Fy[4][N1];
#pragma unroll
         for (int b = 0; b < N1; ++b)
         {
            // F_y(W) = F_x with the momentum components exchanged
            const double W[4] = {uc[1 * NS + b * N1], uc[0 * NS + b * N1], uc[2 * NS + b * N1], uc[3 * NS + b * N1]};
            double F[4];
            if (A.dbg & 2)
            {
#pragma unroll
               for (int c = 0; c < 4; ++c) F[c] = W[c];
            }
            else
               flux_x (W, F);
            Fy[0][b] = F[1];
            Fy[1][b] = F[0];
            Fy[2][b] = F[2];
            Fy[3][b] = F[3];
         }
         const double hw = hx * T.gw[ca];
         double *rc = sR + cs * CS + ca;
#pragma unroll
         for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int b = 0; b < N1; ++b)
            {
               double sy = 0.0;
#pragma unroll
               for (int bp = 0; bp < N1; ++bp) sy = fma (Fy[c][bp], T.dw[bp][b], sy);
               const double fy = WT[c] * T.e1[b] - HB[c] * T.e0[b];
               rc[c * NS + b * N1] = hw * (sy - fy);
            }
      }
      if (row_on)
      {
         const double hx = sGeom[rs * 4 + 2], hy = sGeom[rs * 4 + 3];
         const double *uc = su + rs * CS + rb * N1;
         double u[4][N1], Fx[4][N1];
#pragma unroll
         for (int c = 0; c < 4; ++c) load_line<N1> (uc + c * NS, u[c]);
#pragma unroll
         for (int a = 0; a < N1; ++a)
         {
            const double W[4] = {u[0][a], u[1][a], u[2][a], u[3][a]};
            double F[4];
            if (A.dbg & 2)
            {
#pragma unroll
               for (int c = 0; c < 4; ++c) F[c] = W[c];
            }
            else
               flux_x (W, F);
#pragma unroll
            for (int c = 0; c < 4; ++c) Fx[c][a] = F[c];
         }
         const double hw = hy * T.gw[rb];
#pragma unroll
         for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int a = 0; a < N1; ++a)
            {
               double sx = 0.0;
#pragma unroll
               for (int ap = 0; ap < N1; ++ap) sx = fma (Fx[c][ap], T.dw[ap][a], sx);
               const double fx = WR[c] * T.e1[a] - HL[c] * T.e0[a];
               rrow[c][a] = hw * (sx - fx);
            }
         if (A.gravity != 0.0) // assemble_explicit.cc:78, 108-111
         {
#pragma unroll
            for (int a = 0; a < N1; ++a)
            {
               const double W[4] = {u[0][a], u[1][a], u[2][a], u[3][a]};
               double Gv[4];
               if (A.ext_force)
               {
                  const double *f = A.ext_force + ((size_t) (c0 + rs) * NS + rb * N1 + a) * 2;
                  forcing_ext (W, f[0], f[1], Gv);
               }
               else
                  forcing (W, Gv);
               const double w = A.gravity * (T.gw[a] * T.gw[rb] * hx * hy);
#pragma unroll
               for (int c = 0; c < 4; ++c) rrow[c][a] += Gv[c] * w;
            }
         }
      }
      // old_solution of the own row: requested before the barrier so the loads fly while the block
      // drains P3 (the lines were pulled into L2 by the bulk prefetch at the start)
      if (row_on && need_old)
      {
#pragma unroll
         for (int c = 0; c < 4; ++c) load_line<N1> (A.u_old + ((size_t) (c0 + rs) * D + rb * N1) + c * NS, uold[c]);
      }
      __syncthreads ();

      // ================= P4: M^-1, Euler step, RK combine, write-back, cell average =================
      double *sPart = sm + S::O_PART; // [TC][N1][4] partial cell averages
      const unsigned row_mask = __ballot_sync (0xffffffffu, row_on);
      if (row_on)
      {
         const int cell = c0 + rs;
         const double hx = sGeom[rs * 4 + 2], hy = sGeom[rs * 4 + 3];
         // first stage of a step launched beside the time-step kernels (pdl == 1): dt is the only thing they write
         if (A.pdl == 1) pdl_wait ();
         const double dt = A.dt_cell ? A.dt_cell[cell] : A.time[1];
         const size_t goff = (size_t) cell * D + rb * N1;
         const double wbh = T.gw[rb] * hx * hy;
         // a tile of redundantly updated ghost cells keeps only its means: the solution itself
         // arrives with the halo exchange (possibly before this tile runs)
         const bool owned_tile = (int) blockIdx.x < A.n_tiles_owned;
         double part[4];
#pragma unroll
         for (int c = 0; c < 4; ++c)
         {
            double ry[N1], v[N1];
            load_line<N1> (sR + rs * CS + c * NS + rb * N1, ry);
            if (A.mode == MODE_RHS)
            {
#pragma unroll
               for (int a = 0; a < N1; ++a) v[a] = rrow[c][a] + ry[a];
            }
            else
            {
               double u[N1];
               load_line<N1> (su + rs * CS + c * NS + rb * N1, u);
#pragma unroll
               for (int a = 0; a < N1; ++a)
               {
                  const double invm = fast_rcp (T.gw[a] * wbh); // claw.cc:228-258 on a Cartesian cell
                  const double un = u[a] + dt * (rrow[c][a] + ry[a]) * invm;
                  v[a] = need_old ? (1.0 - A.ark) * un + A.ark * uold[c][a] : un;
               }
            }
            if (owned_tile) store_line<N1> (A.out + goff + c * NS, v);
            double s = 0.0;
#pragma unroll
            for (int a = 0; a < N1; ++a) s = fma (T.gw[a], v[a], s);
            part[c] = T.gw[rb] * s;
         }
         if (A.mode == MODE_STAGE)
         {
            if (N1 == 2 || N1 == 4)
            {
               // compute_cell_average (claw.cc:562-597): the N1 row sums of a cell sit in adjacent
               // lanes; the first lane adds them in row order
               const int lane0 = (tid & 31) - rb;
               double v[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
               for (int b = 0; b < N1; ++b)
#pragma unroll
                  for (int c = 0; c < 4; ++c) v[c] += __shfl_sync (row_mask, part[c], lane0 + b);
               if (rb == 0) store4 (A.avg_out + (size_t) cell * 4, v);
            }
            else
               store4 (sPart + (rs * N1 + rb) * 4, part);
         }
      }
      if (A.mode == MODE_STAGE && N1 != 2 && N1 != 4)
      {
         __syncthreads ();
         for (int j = tid; j < ncb * 4; j += S::THREADS) // compute_cell_average of the updated solution, claw.cc:562-597
         {
            const int s = j >> 2, c = j & 3;
            double v = 0.0;
#pragma unroll
            for (int b = 0; b < N1; ++b) v += sPart[(s * N1 + b) * 4 + c];
            A.avg_out[(size_t) c0 * 4 + j] = v;
         }
      }

      // ================= fused halo exchange over peer memory (p2p_halo.cuh) =================
      // Only the tiles that own cells a peer needs do anything here (the descriptor copy in shared memory says so):
      // the other blocks retire without touching global memory again.
      if (A.fx && A.mode == MODE_STAGE && sdesc[7] > 0 && A.fx->n_send_tiles > 0) // n_send_tiles == 0: wait-only descriptor (the exchange is a kernel of its own)
      {
         const P2PFused &F = *A.fx;
         const int n_send = sdesc[7];
         __syncthreads (); // the tile's write-back is complete and visible to the block
         const int *ent = F.send_entries + 3 * (size_t) sdesc[6];
         constexpr int D2 = D / 2;
         for (int i = tid; i < n_send * D2; i += S::THREADS)
         {
            const int e = i / D2, c = i - e * D2;
            const int cell = ent[3 * e], pi = ent[3 * e + 1], dc = ent[3 * e + 2];
            reinterpret_cast<double2 *> (F.dstU[pi] + (size_t) dc * D)[c] = __ldcg (reinterpret_cast<const double2 *> (A.out + (size_t) cell * D) + c);
         }
         for (int i = tid; i < n_send * 2; i += S::THREADS)
         {
            const int e = i >> 1;
            const int cell = ent[3 * e], pi = ent[3 * e + 1], dc = ent[3 * e + 2];
            reinterpret_cast<double2 *> (F.dstA[pi] + (size_t) dc * 4)[i & 1] = __ldcg (reinterpret_cast<const double2 *> (A.avg_out + (size_t) cell * 4) + (i & 1));
         }
         __threadfence ();
         __syncthreads ();
         if (tid == 0)
         {
            const unsigned int done = atomicAdd (F.send_counter, 1u);
            if (done == (unsigned int) F.n_send_tiles - 1)
            {
               // last sending tile of this stage: everything the peers need is on its way
               *F.send_counter = 0;
               __threadfence_system ();
               const unsigned long long e = *reinterpret_cast<volatile unsigned long long *> (F.epochs);
               for (int p = 0; p < F.npeers; ++p) st_release_sys (F.peer_flags[p] + F.world + F.me, e);
               // The epoch advances here, not when the launch retires (that took an atomic per block).  The tiles that
               // wait for the peers are the boundary tiles, first in the tile order, and have read the epoch long before
               // the last sender finishes; a waiting tile that starts later reads the advanced value and waits for the
               // flags of THIS exchange instead -- raised unconditionally by every peer, and implying the earlier ones.
               *reinterpret_cast<volatile unsigned long long *> (F.epochs) = e + 1;
            }
         }
      }
   }
}
