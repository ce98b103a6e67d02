// Stage kernel for the Pk (orthonormal Legendre) basis, one THREAD per cell.
//
// Same fused stage as StageKernel / row_stage_kernel -- assemble_system (cell, face and boundary
// workers, src/assemble_explicit.cc:30-452), M^-1 (claw.cc:228-258), forward-Euler update, SSP-RK
// combine (claw.cc:694-713, 757-760), cell average (claw.cc:562-597) -- but organised like the
// limiter cell kernel: a thread keeps the D = 4 (k+1)(k+2)/2 modal coefficients of its cell and the
// residual in registers and walks through volume points, faces and update on its own; the basis
// tables are constant-bank operands of fully unrolled FMA chains.  A block of 128 cells is staged
// with coalesced loads into shared memory (odd row stride: the per-thread walks are free of bank
// conflicts), neighbours inside the block are read from there, and the updated cells go back
// through a second shared buffer with coalesced stores that apply the RK combine on the way.  Every face is evaluated from both sides with the same arguments in the same order (the
// reference's "plus" side first, MeshWorker owner rule of assemble_explicit.cc:440 / both sides for
// periodic pairs, src_mpi/assemble_explicit.cc:186-260), so the two cells subtract bit-identical
// fluxes: conservative to round-off without atomics, and a sharded run equals the single-GPU run
// bit for bit.  Neighbour coefficients come through L1/L2 (x neighbours are the adjacent threads'
// cells).  The phase-structured tile kernel (kernels.cuh) stays selectable: DFLO_B200_PK=tile.
#pragma once

#include "kernels.cuh"

namespace dflo
{
   struct CellStageArgs
   {
      const double *u;        // current_solution [n_local][D]
      const double *u_old;    // old_solution
      double *out;            // MODE_STAGE: updated solution (a different buffer than u); MODE_RHS: right_hand_side
      const double *avg;      // cell_average of u [n_local][4] (LxF, kep)
      double *avg_out;        // cell_average of the updated solution
      const int *nbr;         // [n_local][4] local neighbour, or -1 - boundary face
      const unsigned char *fflags; // [n_local][4] FACE_*
      const double *geom;     // [n_local][4] x0, y0, hx, hy
      const double *bc_g;     // [n_bfaces][n_q_face][4]
      const int *bkind;       // [n_bfaces]
      const double *tab;      // flat stage tables (pack_stage_tables); the CUDA build reads its constant-memory copy
      const double *time;     // device scalars: [0] elapsed time, [1] dt
      const double *dt_cell;  // optional per-cell dt, else nullptr
      const double *ext_force; // [n_local][n_q][2] or nullptr (see StageArgs)
      int n_compute;          // cells updated: owned (+ ghost layer 1 when a limiter follows)
      int n_keep;             // cells >= n_keep are redundantly updated ghost cells: only their means are stored
      int mode, compat_mpi;
      double ark, gravity;
   };

#if !defined(__CUDACC__)
   struct double2 // the CPU emulation's stand-in for the CUDA vector type
   {
      double x, y;
   };
#endif
   constexpr int PK_TAB_MAX = stage_table_size (BASIS_PK, 4);
#if defined(__CUDACC__)
   __constant__ double c_pk_tab[5][PK_TAB_MAX]; // indexed by N1 = k+1
#endif

   template <int N1, int FLUX>
   struct PkCellStageKernel
   {
      typedef CellStageArgs Args;
      static constexpr int NQ = N1 * N1;
      static constexpr int NS = n_scalar (BASIS_PK, N1);
      static constexpr int D = 4 * NS;
#ifndef DFLO_PK_THREADS
#define DFLO_PK_THREADS 128
#endif
#ifndef DFLO_PK_UNROLL_STAGE
#define DFLO_PK_UNROLL_STAGE 1
#endif
#ifndef DFLO_PK_UB
#define DFLO_PK_UB 12
#endif
#ifndef DFLO_PK_MIN_BLOCKS
#define DFLO_PK_MIN_BLOCKS 1
#endif
      static constexpr int THREADS = DFLO_PK_THREADS;
      static constexpr int CPB = THREADS;      // cells per block
      static constexpr int MIN_BLOCKS = DFLO_PK_MIN_BLOCKS;
      static constexpr int NPHASE = 3;
      static constexpr int ROW = D + 1;        // odd row stride
      static constexpr int SMEM_DOUBLES = 2 * CPB * ROW;
      static int grid (int n_compute) { return (n_compute + CPB - 1) / CPB; }
      // Pk: phi[NQ*NS] dphix[NQ*NS] dphiy[NQ*NS] phiface[4*N1*NS] gw[N1]
      static constexpr int O_PHI = 0, O_DPX = NQ * NS, O_DPY = 2 * NQ * NS, O_PF = 3 * NQ * NS, O_GW = 3 * NQ * NS + 4 * N1 * NS;

      static DFLO_DEV double T (const double *tb, int i)
      {
#if defined(__CUDA_ARCH__)
         (void) tb;
         return c_pk_tab[N1][i];
#else
         return tb[i];
#endif
      }

      // trace at point q of face f: the fma chain of StageKernel::trace, from registers or from memory
      static DFLO_DEV void trace_own (const double *tb, const double (&u)[4][NS], int f, int q, double W[4])
      {
#pragma unroll
         for (int c = 0; c < 4; ++c)
         {
            double s = 0.0;
#pragma unroll
            for (int m = 0; m < NS; ++m) s = fma (T (tb, O_PF + (f * N1 + q) * NS + m), u[c][m], s);
            W[c] = s;
         }
      }

      // face F of the cell (0 left, 1 right, 2 bottom, 3 top): interior / periodic face or boundary face
      template <int F>
      static DFLO_DEV void face_term (const Args &A, const double *tb, int cell, int c0, int ncb, const double *sm, const double (&u)[4][NS],
                                      double (&r)[4][NS], const double (&Ao)[4], double hx, double hy)
      {
            const int nb = A.nbr[(size_t) cell * 4 + F];
            const int fl = A.fflags[(size_t) cell * 4 + F];
            const double nx = (F == 0) ? -1.0 : (F == 1) ? 1.0 : 0.0;
            const double ny = (F == 2) ? -1.0 : (F == 3) ? 1.0 : 0.0;
            const double len = (F < 2) ? hy : hx;
            double un[4][NS], An[4] = {0.0, 0.0, 0.0, 0.0};
            if (nb >= 0)
            {
               // a neighbour staged by this block is read from its row
               const double *pn = ((unsigned) (nb - c0) < (unsigned) ncb) ? sm + (nb - c0) * ROW : A.u + (size_t) nb * D;
#pragma unroll
               for (int c = 0; c < 4; ++c)
#pragma unroll
                  for (int m = 0; m < NS; ++m) un[c][m] = pn[c * NS + m];
               if (flux_uses_averages (FLUX))
               {
#pragma unroll
                  for (int c = 0; c < 4; ++c) An[c] = A.avg[(size_t) nb * 4 + c];
               }
            }
#pragma unroll
            for (int q = 0; q < N1; ++q)
            {
               double Wo[4], Wn[4], H[4];
               trace_own (tb, u, F, q, Wo);
               bool plus = true; // this cell is the "plus" side of the flux call
               if (nb >= 0)
               {
                  if (fl & FACE_FLIP)
                     trace_own (tb, un, F ^ 1, N1 - 1 - q, Wn);
                  else
                     trace_own (tb, un, F ^ 1, q, Wn);
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
               // the axis-specialised Riemann problem of the row kernel (euler.cuh face_flux_axis): states on the
               // low / high coordinate side of the face, the reference's plus side marked; H = flux along +e_DIR,
               // the same bits from whichever of the two cells evaluates it
               constexpr bool low = (F & 1) != 0; // this cell sits on the low side of its faces 1 and 3
               face_flux_axis<FLUX, F / 2> (low ? plus : !plus, low ? Wo : Wn, low ? Wn : Wo, low ? Ao : An, low ? An : Ao, H);
               const double wl = (low ? 1.0 : -1.0) * (T (tb, O_GW + q) * len);
#pragma unroll
               for (int c = 0; c < 4; ++c)
               {
                  const double h = wl * H[c];
#pragma unroll
                  for (int m = 0; m < NS; ++m) r[c][m] = fma (-h, T (tb, O_PF + (F * N1 + q) * NS + m), r[c][m]);
               }
            }
               }

      // p = 0: stage the block's cells; 1: one thread per cell, result into the second buffer; 2: RK combine,
      // coalesced write-back, cell averages
      static DFLO_DEV void phase (int p, const Args &A, double *sm, int tid, int bid)
      {
         const int c0 = bid * CPB;
         const int ncb = (A.n_compute - c0 < CPB) ? A.n_compute - c0 : CPB;
         double *smB = sm + CPB * ROW;
         if (p == 0)
         {
            const double *src = A.u + (size_t) c0 * D;
            if (DFLO_PK_UNROLL_STAGE && ncb == CPB) // full block: 16-byte loads, UB of them in flight per thread
            {
               constexpr int UB = DFLO_PK_UB, NV = D / 2; // D is even: NV double2 per cell, CPB * NV in the block
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
         }
         else if (p == 1)
         {
            if (tid < ncb) cell_work (A, c0, ncb, tid, sm, smB + tid * ROW);
         }
         else
         {
            const bool combine = A.mode == MODE_STAGE && A.ark != 0.0;
            const double *uo = A.u_old + (size_t) c0 * D;
            double *dst = A.out + (size_t) c0 * D;
            if (DFLO_PK_UNROLL_STAGE && ncb == CPB && combine) // full block: 16-byte accesses, UB old_solution loads in flight
            {
               constexpr int UB = DFLO_PK_UB, NV = D / 2;
               const double2 *o2 = reinterpret_cast<const double2 *> (uo);
               double2 *d2 = reinterpret_cast<double2 *> (dst);
#pragma unroll
               for (int k0 = 0; k0 < NV; k0 += UB)
               {
                  double2 w[UB];
#pragma unroll
                  for (int k = 0; k < UB; ++k)
                     if (k0 + k < NV) w[k] = o2[tid + (k0 + k) * THREADS];
#pragma unroll
                  for (int k = 0; k < UB; ++k)
                     if (k0 + k < NV)
                     {
                        const int i = 2 * (tid + (k0 + k) * THREADS);
                        const int lc = i / D, kk = i % D; // kk even: both entries belong to the same cell
                        double2 v;
                        v.x = (1.0 - A.ark) * smB[lc * ROW + kk] + A.ark * w[k].x; // claw.cc:757-760
                        v.y = (1.0 - A.ark) * smB[lc * ROW + kk + 1] + A.ark * w[k].y;
                        if (c0 + lc < A.n_keep) d2[tid + (k0 + k) * THREADS] = v;
                        if (kk % NS == 0) A.avg_out[(size_t) (c0 + lc) * 4 + kk / NS] = v.x;
                        if ((kk + 1) % NS == 0) A.avg_out[(size_t) (c0 + lc) * 4 + (kk + 1) / NS] = v.y;
                     }
               }
            }
            else
               for (int i = tid; i < ncb * D; i += THREADS)
               {
                  const int lc = i / D, k = i % D;
                  double v = smB[lc * ROW + k];
                  if (combine) v = (1.0 - A.ark) * v + A.ark * uo[i]; // claw.cc:757-760
                  // cells >= n_keep: redundantly updated ghost cells, only their means are kept
                  if (A.mode == MODE_RHS || c0 + lc < A.n_keep) dst[i] = v;
                  if (A.mode == MODE_STAGE && k % NS == 0) A.avg_out[(size_t) (c0 + lc) * 4 + k / NS] = v; // mode 0 = mean
               }
         }
      }

      // residual and Euler update of cell c0 + lc; res = its row of the second buffer
      static DFLO_DEV void cell_work (const Args &A, int c0, int ncb, int lc, const double *sm, double *res)
      {
         const int cell = c0 + lc;
         const double *tb = A.tab;
         const double hx = A.geom[(size_t) cell * 4 + 2], hy = A.geom[(size_t) cell * 4 + 3];
#if defined(__CUDA_ARCH__)
         // neighbours outside the block: pull their coefficients towards L1 now, they are read after the volume term
#pragma unroll
         for (int f = 0; f < 4; ++f)
         {
            const int nb = A.nbr[(size_t) cell * 4 + f];
            if (nb >= 0 && (unsigned) (nb - c0) >= (unsigned) ncb)
            {
               const char *pn = reinterpret_cast<const char *> (A.u + (size_t) nb * D);
#pragma unroll
               for (int b = 0; b < D * 8; b += 128) asm volatile ("prefetch.global.L1 [%0];" ::"l"(pn + b));
               if ((D * 8) % 128 != 0) asm volatile ("prefetch.global.L1 [%0];" ::"l"(pn + D * 8 - 8));
            }
         }
#endif
         double u[4][NS], r[4][NS];
         {
            const double *uc = sm + lc * ROW;
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
               for (int m = 0; m < NS; ++m)
               {
                  u[c][m] = uc[c * NS + m];
                  r[c][m] = 0.0;
               }
         }

         // ---- cell term, assemble_explicit.cc:30-120: rhs_i += sum_q F(W_q).grad(phi_i) JxW (+ forcing) ----
#pragma unroll
         for (int q = 0; q < NQ; ++q)
         {
            double W[4], Fx[4], Fy[4];
#pragma unroll
            for (int c = 0; c < 4; ++c)
            {
               double s = 0.0;
#pragma unroll
               for (int m = 0; m < NS; ++m) s = fma (T (tb, O_PHI + q * NS + m), u[c][m], s);
               W[c] = s;
            }
            flux_matrix (W, Fx, Fy);
            const double w2 = T (tb, O_GW + q % N1) * T (tb, O_GW + q / N1);
            const double wx = w2 * hy, wy = w2 * hx;
#pragma unroll
            for (int c = 0; c < 4; ++c)
            {
               const double ax = wx * Fx[c], ay = wy * Fy[c];
#pragma unroll
               for (int m = 0; m < NS; ++m) r[c][m] = fma (ax, T (tb, O_DPX + q * NS + m), fma (ay, T (tb, O_DPY + q * NS + m), r[c][m]));
            }
            if (A.gravity != 0.0) // assemble_explicit.cc:78, 108-111
            {
               double Gv[4];
               if (A.ext_force)
                  forcing_ext (W, A.ext_force[((size_t) cell * NQ + q) * 2], A.ext_force[((size_t) cell * NQ + q) * 2 + 1], Gv);
               else
                  forcing (W, Gv);
               const double wg = A.gravity * (w2 * hx * hy);
#pragma unroll
               for (int c = 0; c < 4; ++c)
               {
                  const double g = wg * Gv[c];
#pragma unroll
                  for (int m = 0; m < NS; ++m) r[c][m] = fma (g, T (tb, O_PHI + q * NS + m), r[c][m]);
               }
            }
         }

         // ---- face and boundary terms, assemble_explicit.cc:127-427 ----
         double Ao[4] = {0.0, 0.0, 0.0, 0.0};
         if (flux_uses_averages (FLUX))
         {
#pragma unroll
            for (int c = 0; c < 4; ++c) Ao[c] = A.avg[(size_t) cell * 4 + c];
         }
         face_term<0> (A, tb, cell, c0, ncb, sm, u, r, Ao, hx, hy);
         face_term<1> (A, tb, cell, c0, ncb, sm, u, r, Ao, hx, hy);
         face_term<2> (A, tb, cell, c0, ncb, sm, u, r, Ao, hx, hy);
         face_term<3> (A, tb, cell, c0, ncb, sm, u, r, Ao, hx, hy);

         // ---- M^-1 (orthonormal modes: 1/|K|, claw.cc:228-258) and Euler step (claw.cc:694-713); MODE_RHS: the residual ----
         if (A.mode == MODE_RHS)
         {
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
               for (int m = 0; m < NS; ++m) res[c * NS + m] = r[c][m];
            return;
         }
         const double invm = 1.0 / (hx * hy);
         const double dt = A.dt_cell ? A.dt_cell[cell] : A.time[1];
#pragma unroll
         for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int m = 0; m < NS; ++m) res[c * NS + m] = u[c][m] + dt * r[c][m] * invm;
      }
   };
}
