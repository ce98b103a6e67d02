// Stage kernel for the Pk (orthonormal Legendre) basis: one THREAD per cell, every face solved ONCE per block.
//
// Same fused stage as StageKernel / row_stage_kernel -- assemble_system (cell, face and boundary
// workers, src/assemble_explicit.cc:30-452), M^-1 (claw.cc:228-258), forward-Euler update, SSP-RK
// combine (claw.cc:694-713, 757-760), cell average (claw.cc:562-597).  A block of 128 consecutive
// cells is staged with coalesced loads into shared memory (odd row stride: the per-thread walks are
// free of bank conflicts) and goes through five barrier-separated phases:
//
//   0  stage the block's cells;
//   1  thread = cell: its LOW faces (0: left, 2: bottom).  The Pk trace on a face is a polynomial of
//      degree k in the tangential variable, so the two cells are first reduced to k+1 tangential
//      coefficients per component (the normal direction summed out with L_i(0) / L_i(1)), the k+1
//      Riemann problems are posed along +e_x / +e_y exactly like the row kernel's (face_flux_axis: the
//      reference's "plus" side marked, MeshWorker owner rule of assemble_explicit.cc:440), and the
//      weighted fluxes are projected back on the tangential Legendre modes: k+1 MOMENTS per component.
//      They go into the cell's own slot of that face and, when the neighbour is staged by the same
//      block, into the neighbour's slot of its high face;
//   2  the high faces (1: right, 3: top) nobody solved -- neighbour outside the block, boundary faces,
//      periodic pairs (both sides integrate their own problem, src_mpi/assemble_explicit.cc:186-260) --
//      were pushed on a job list in phase 1 and are solved now, one thread per job, by the same code;
//   3  thread = cell: volume term (structural zeros of the derivative tables skipped), lifting of the
//      four moment sets, M^-1 = 1/|K|, Euler step, into the cell's own row in place;
//   4  RK combine with old_solution and coalesced write-back, cell means.
//
// Whoever solves a face gets bit-identical arguments in the same order, and each cell adds volume
// term and faces 0..3 in a fixed order: conservative to round-off without atomics, and a sharded run
// equals the single-GPU run bit for bit.  The sum-factorised traces and lifts re-associate the
// reference's dense i x q loops (round-off level differences; the tests state the tolerance).
#pragma once

#include "kernels.cuh"

// the compute phases as functions of their own: ptxas allocates registers phase by phase (inlined into one body they
// cost 248 registers, or spills under the cap)
#ifndef DFLO_PK_NOINLINE
#define DFLO_PK_NOINLINE 1
#endif
#if defined(__CUDACC__) && DFLO_PK_NOINLINE
#define DFLO_PHASE_FN __device__ __noinline__
#else
#define DFLO_PHASE_FN DFLO_DEV
#endif

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
      int pf_blocks;          // L2 prefetch distance in blocks (0: none): the cells of block bid + pf_blocks are requested at block start
      double ark, gravity;
   };

   // The cell kernel lets the block solve an interior face once, from the cell that has it as a LOW face (0, 2), and hands the
   // moments to the neighbour's face F ^ 1: every interior, non-periodic, unflipped face must be seen as F / F ^ 1 by its two
   // cells, with the same flags.  True for the conforming Cartesian meshes the reference admits for Pk (parameters.cc:536-550).
   // (Only pairs of updated cells can share a block.)
   inline bool pk_cell_mesh_ok (const int *nbr, const unsigned char *fflags, int n_compute)
   {
      for (int c = 0; c < n_compute; ++c)
         for (int f = 0; f < 4; ++f)
         {
            const int nb = nbr[4 * (size_t) c + f];
            if (nb < 0 || nb >= n_compute) continue;
            const int skip = FACE_PERIODIC | FACE_FLIP;
            if (fflags[4 * (size_t) c + f] & skip) continue;
            if (nbr[4 * (size_t) nb + (f ^ 1)] != c || (fflags[4 * (size_t) nb + (f ^ 1)] & skip)) return false;
         }
      return true;
   }

   constexpr int PK_TAB_MAX = stage_table_size (BASIS_PK, 4);
#if defined(__CUDACC__)
   __constant__ double c_pk_tab[5][PK_TAB_MAX]; // indexed by N1 = k+1
#endif


   template <int N1, int FLUX>
   struct PkCellStageKernel
   {
      typedef CellStageArgs Args;
      static constexpr int K = N1 - 1;
      static constexpr int NQ = N1 * N1;
      static constexpr int NS = n_scalar (BASIS_PK, N1);
      static constexpr int D = 4 * NS;
#ifndef DFLO_PK_THREADS
#define DFLO_PK_THREADS 128
#endif
#ifndef DFLO_PK_UB
#define DFLO_PK_UB 12
#endif
#ifndef DFLO_PK_QUNROLL
#define DFLO_PK_QUNROLL 0 // the face loop over the k+1 Riemann problems stays rolled: measured faster (115 vs 117 us, cfg3) at 2/3 of the code
#endif
#ifndef DFLO_PK_MIN_BLOCKS
#define DFLO_PK_MIN_BLOCKS 3
#endif
      static constexpr int THREADS = DFLO_PK_THREADS;
      static constexpr int CPB = THREADS;      // cells per block
      static constexpr int MIN_BLOCKS = DFLO_PK_MIN_BLOCKS;
      static constexpr int NPHASE = 5;
      static constexpr int ROW = D + 1;        // odd row stride
      static constexpr int NM = 4 * N1;        // moments per face: component x tangential mode
      // shared memory: rows [CPB][ROW] | moments [4 faces][NM][CPB] (cell fastest) | job list [2 CPB] + 2 counters (ints)
      static constexpr int O_MOM = CPB * ROW, O_JOB = O_MOM + 4 * NM * CPB;
      static constexpr int SMEM_DOUBLES = O_JOB + CPB + 1;
      static int grid (int n_compute) { return (n_compute + CPB - 1) / CPB; }
      // Pk: phi[NQ*NS] dphix[NQ*NS] dphiy[NQ*NS] phiface[4*N1*NS] gw[N1]
      static constexpr int O_PHI = 0, O_DPX = NQ * NS, O_DPY = 2 * NQ * NS, O_PF = 3 * NQ * NS, O_GW = 3 * NQ * NS + 4 * N1 * NS;

      // mode (i, j) = L_i(x) L_j(y), i + j <= k, in the order of tables.cc / claw.cc:104-114 (j outer, i inner)
      static DFLO_DEV constexpr int mode (int i, int j) { return j * (K + 1) - j * (j - 1) / 2 + i; }

      static DFLO_DEV double T (const double *tb, int i)
      {
#if defined(__CUDA_ARCH__)
         (void) tb;
         return c_pk_tab[N1][i];
#else
         return tb[i];
#endif
      }
      // 1-D tables read out of the 2-D ones (L_0 = 1 exactly): L_t at Gauss point q; L_i at the end of face F
      static DFLO_DEV double LG (const double *tb, int t, int q) { return T (tb, O_PHI + q * NS + mode (t, 0)); }
      static DFLO_DEV double LE (const double *tb, int F, int i) { return T (tb, O_PF + (F * N1) * NS + (F < 2 ? mode (i, 0) : mode (0, i))); }

      static DFLO_DEV int bump (int *p)
      {
#if defined(__CUDA_ARCH__)
         return atomicAdd (p, 1);
#else
         return (*p)++;
#endif
      }

      // tangential coefficients of the trace on face F of the cell whose coefficients are at p: a[c][t]
      template <int F>
      static DFLO_DEV void face_coefficients (const double *tb, const double *p, double (&a)[4][N1])
      {
#pragma unroll
         for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int t = 0; t < N1; ++t)
            {
               // normal index n = 0 .. k - t; L_0 = 1
               double s = p[c * NS + (F < 2 ? mode (0, t) : mode (t, 0))];
#pragma unroll
               for (int n = 1; n < N1; ++n)
                  if (n + t <= K) s = fma (LE (tb, F, n), p[c * NS + (F < 2 ? mode (n, t) : mode (t, n))], s);
               a[c][t] = s;
            }
      }

      // Face F of `cell` seen from that cell: moments b[c][t] = sum_q w_q |face| H_c(q) L_t(g_q) of the numerical flux H
      // along +e_x (F < 2) / +e_y; interior, periodic and boundary faces (assemble_explicit.cc:127-427)
      template <int F>
      static DFLO_DEV void face_moments (const Args &A, const double *tb, int cell, int c0, int ncb, const double *sm, const double *own,
                                         double (&b)[4][N1])
      {
         const int nb = A.nbr[(size_t) cell * 4 + F];
         const int fl = A.fflags[(size_t) cell * 4 + F];
         const double nx = (F == 0) ? -1.0 : (F == 1) ? 1.0 : 0.0;
         const double ny = (F == 2) ? -1.0 : (F == 3) ? 1.0 : 0.0;
         const double len = A.geom[(size_t) cell * 4 + (F < 2 ? 3 : 2)];
         double ao[4][N1], an[4][N1];
         double Ao[4] = {0.0, 0.0, 0.0, 0.0}, An[4] = {0.0, 0.0, 0.0, 0.0};
         face_coefficients<F> (tb, own, ao);
         if (flux_uses_averages (FLUX))
         {
#pragma unroll
            for (int c = 0; c < 4; ++c) Ao[c] = A.avg[(size_t) cell * 4 + c];
         }
         bool plus = true; // this cell is the "plus" side of the flux call
         if (nb >= 0)
         {
            // a neighbour staged by this block is read from its row (two calls: shared / global loads, no generic ones)
            if ((unsigned) (nb - c0) < (unsigned) ncb)
               face_coefficients<(F ^ 1)> (tb, sm + (nb - c0) * ROW, an);
            else
               face_coefficients<(F ^ 1)> (tb, A.u + (size_t) nb * D, an);
            if (fl & FACE_FLIP) // the neighbour's face runs backwards: t -> (-1)^t in the symmetric Gauss points
            {
#pragma unroll
               for (int c = 0; c < 4; ++c)
#pragma unroll
                  for (int t = 1; t < N1; t += 2) an[c][t] = -an[c][t];
            }
            if (flux_uses_averages (FLUX))
            {
#pragma unroll
               for (int c = 0; c < 4; ++c) An[c] = A.avg[(size_t) nb * 4 + c];
            }
            plus = (fl & (FACE_OWNER | FACE_PERIODIC)) != 0;
         }
#pragma unroll
         for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int t = 0; t < N1; ++t) b[c][t] = 0.0;
#if DFLO_PK_QUNROLL
#pragma unroll
#else
#pragma unroll 1
#endif
         for (int q = 0; q < N1; ++q)
         {
            double Wo[4], Wn[4], H[4];
#pragma unroll
            for (int c = 0; c < 4; ++c)
            {
               double s = ao[c][0];
#pragma unroll
               for (int t = 1; t < N1; ++t) s = fma (LG (tb, t, q), ao[c][t], s);
               Wo[c] = s;
            }
            if (nb >= 0)
            {
#pragma unroll
               for (int c = 0; c < 4; ++c)
               {
                  double s = an[c][0];
#pragma unroll
                  for (int t = 1; t < N1; ++t) s = fma (LG (tb, t, q), an[c][t], s);
                  Wn[c] = s;
               }
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
            const double wl = T (tb, O_GW + q) * len;
#pragma unroll
            for (int c = 0; c < 4; ++c)
            {
               const double h = wl * H[c];
               b[c][0] += h;
#pragma unroll
               for (int t = 1; t < N1; ++t) b[c][t] = fma (h, LG (tb, t, q), b[c][t]);
            }
         }
      }

      // does the block solve face F (1 or 3) of this cell from the neighbour's side in phase 1?
      static DFLO_DEV bool covered (const Args &A, int cell, int F, int c0, int ncb)
      {
         const int nb = A.nbr[(size_t) cell * 4 + F];
         if (nb < 0 || (unsigned) (nb - c0) >= (unsigned) ncb) return false;
         // the neighbour sees this cell across its face F ^ 1 with the same flags (checked on the host: pk_cell_mesh_ok)
         return !(A.fflags[(size_t) cell * 4 + F] & (FACE_PERIODIC | FACE_FLIP));
      }

      static DFLO_DEV void store_moments (double *mom, int F, int lc, const double (&b)[4][N1])
      {
#pragma unroll
         for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int t = 0; t < N1; ++t) mom[((F * 4 + c) * N1 + t) * CPB + lc] = b[c][t];
      }

      template <int F>
      static DFLO_PHASE_FN void job_face (const Args &A, int c0, int ncb, int lc, const double *sm, double *mom)
      {
         double b[4][N1];
         face_moments<F> (A, A.tab, c0 + lc, c0, ncb, sm, sm + lc * ROW, b);
         store_moments (mom, F, lc, b);
      }

      template <int F>
      static DFLO_PHASE_FN void low_face (const Args &A, int c0, int ncb, int lc, const double *sm, double *mom)
      {
         double b[4][N1];
         face_moments<F> (A, A.tab, c0 + lc, c0, ncb, sm, sm + lc * ROW, b);
         store_moments (mom, F, lc, b);
         const int nb = A.nbr[(size_t) (c0 + lc) * 4 + F];
         // the neighbour's copy; it uses it only if `covered` says so, otherwise a phase-2 job overwrites it
         if (nb >= 0 && (unsigned) (nb - c0) < (unsigned) ncb && !(A.fflags[(size_t) (c0 + lc) * 4 + F] & (FACE_PERIODIC | FACE_FLIP)))
            store_moments (mom, F ^ 1, nb - c0, b);
      }

      static DFLO_DEV void phase (int p, const Args &A, double *sm, int tid, int bid)
      {
         const int c0 = bid * CPB;
         const int ncb = (A.n_compute - c0 < CPB) ? A.n_compute - c0 : CPB;
         double *mom = sm + O_MOM;
         int *jobs = reinterpret_cast<int *> (sm + O_JOB), *cnt = jobs + 2 * CPB;
         if (p == 0)
         {
            if (tid == 0) cnt[0] = cnt[1] = 0;
            const double *src = A.u + (size_t) c0 * D;
#if defined(__CUDA_ARCH__)
            // what the later phases read behind dependent addresses, requested now: old_solution of the block into L2 (one
            // bulk prefetch), the cell's face tables and the time-step scalar towards L1
            if (tid == 0 && A.mode == MODE_STAGE && A.ark != 0.0)
               asm volatile ("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(A.u_old + (size_t) c0 * D), "r"((unsigned) (ncb * D * sizeof (double))) : "memory");
            if (tid == 32 && A.pf_blocks > 0) // a block that runs on this SM soon: its cells into L2
            {
               const int pc0 = (bid + A.pf_blocks) * CPB;
               if (pc0 + CPB <= A.n_compute)
                  asm volatile ("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(A.u + (size_t) pc0 * D), "r"((unsigned) (CPB * D * sizeof (double))) : "memory");
            }
            if (tid < ncb)
            {
               asm volatile ("prefetch.global.L1 [%0];" ::"l"(A.geom + (size_t) (c0 + tid) * 4));
               if (tid % 8 == 0) asm volatile ("prefetch.global.L1 [%0];" ::"l"(A.nbr + (size_t) (c0 + tid) * 4));
               if (tid % 32 == 0) asm volatile ("prefetch.global.L1 [%0];" ::"l"(A.fflags + (size_t) (c0 + tid) * 4));
               if (tid == 0) asm volatile ("prefetch.global.L1 [%0];" ::"l"(A.time));
            }
#endif
            if (ncb == CPB) // full block: 16-byte loads, UB of them in flight per thread
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
#if defined(__CUDA_ARCH__)
            // neighbours outside the block: pull their coefficients towards L1, the face phases read them
            if (tid < ncb)
            {
               const int4 nb4 = *reinterpret_cast<const int4 *> (A.nbr + (size_t) (c0 + tid) * 4);
               const int nbs[4] = {nb4.x, nb4.y, nb4.z, nb4.w};
#pragma unroll
               for (int f = 0; f < 4; ++f)
                  if (nbs[f] >= 0 && (unsigned) (nbs[f] - c0) >= (unsigned) ncb)
                  {
                     const char *pn = reinterpret_cast<const char *> (A.u + (size_t) nbs[f] * D);
#pragma unroll
                     for (int b = 0; b < D * 8; b += 128) asm volatile ("prefetch.global.L1 [%0];" ::"l"(pn + b));
                     if ((D * 8) % 128 != 0) asm volatile ("prefetch.global.L1 [%0];" ::"l"(pn + D * 8 - 8));
                  }
            }
#endif
         }
         else if (p == 1)
         {
            if (tid >= ncb) return;
            // high faces nobody else solves: F = 1 jobs from the front of the list, F = 3 jobs from its back
            if (!covered (A, c0 + tid, 1, c0, ncb)) jobs[bump (cnt)] = tid;
            if (!covered (A, c0 + tid, 3, c0, ncb)) jobs[2 * CPB - 1 - bump (cnt + 1)] = tid;
            low_face<0> (A, c0, ncb, tid, sm, mom);
            low_face<2> (A, c0, ncb, tid, sm, mom);
         }
         else if (p == 2)
         {
            const int n1 = cnt[0], n1r = (n1 + 31) & ~31, n3 = cnt[1]; // warps do not mix the two kinds
            for (int k = tid; k < n1r + n3; k += THREADS)
            {
               if (k < n1)
                  job_face<1> (A, c0, ncb, jobs[k], sm, mom);
               else if (k >= n1r)
                  job_face<3> (A, c0, ncb, jobs[2 * CPB - 1 - (k - n1r)], sm, mom);
            }
         }
         else if (p == 3)
         {
            if (tid < ncb) cell_work (A, c0, tid, sm + tid * ROW, mom);
         }
         else
         {
            const bool combine = A.mode == MODE_STAGE && A.ark != 0.0;
            const double *uo = A.u_old + (size_t) c0 * D;
            double *dst = A.out + (size_t) c0 * D;
            if (ncb == CPB && A.mode == MODE_STAGE) // full block: 16-byte accesses, UB old_solution loads in flight
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
                     if (combine && k0 + k < NV) w[k] = o2[tid + (k0 + k) * THREADS];
#pragma unroll
                  for (int k = 0; k < UB; ++k)
                     if (k0 + k < NV)
                     {
                        const int i = 2 * (tid + (k0 + k) * THREADS);
                        const int lc = i / D, kk = i % D; // kk even: both entries belong to the same cell
                        double2 v;
                        v.x = sm[lc * ROW + kk];
                        v.y = sm[lc * ROW + kk + 1];
                        if (combine) // claw.cc:757-760
                        {
                           v.x = (1.0 - A.ark) * v.x + A.ark * w[k].x;
                           v.y = (1.0 - A.ark) * v.y + A.ark * w[k].y;
                        }
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
                  double v = sm[lc * ROW + k];
                  if (combine) v = (1.0 - A.ark) * v + A.ark * uo[i]; // claw.cc:757-760
                  // cells >= n_keep: redundantly updated ghost cells, only their means are kept
                  if (A.mode == MODE_RHS || c0 + lc < A.n_keep) dst[i] = v;
                  if (A.mode == MODE_STAGE && k % NS == 0) A.avg_out[(size_t) (c0 + lc) * 4 + k / NS] = v; // mode 0 = mean
               }
         }
      }

      // r -= (outward flux through face F) tested with the basis: lifting of the moments
      template <int F>
      static DFLO_DEV void lift (const double *tb, const double *mom, int lc, double (&r)[4][NS])
      {
#pragma unroll
         for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int t = 0; t < N1; ++t)
            {
               const double bt = mom[((F * 4 + c) * N1 + t) * CPB + lc];
               const double sb = (F & 1) ? -bt : bt; // outward normal = +e on the high faces
#pragma unroll
               for (int n = 0; n < N1; ++n)
                  if (n + t <= K)
                  {
                     const int m = F < 2 ? mode (n, t) : mode (t, n);
                     if (n == 0)
                        r[c][m] += sb;
                     else
                        r[c][m] = fma (sb, LE (tb, F, n), r[c][m]);
                  }
            }
      }

      // volume term, lifting and Euler update of cell c0 + lc; row = its coefficients, replaced by the result
      static DFLO_PHASE_FN void cell_work (const Args &A, int c0, int lc, double *row, const double *mom)
      {
         const int cell = c0 + lc;
         const double *tb = A.tab;
         const double hx = A.geom[(size_t) cell * 4 + 2], hy = A.geom[(size_t) cell * 4 + 3];
         const double dt = A.mode == MODE_RHS ? 0.0 : A.dt_cell ? A.dt_cell[cell] : A.time[1];
         double u[4][NS], r[4][NS];
#pragma unroll
         for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int m = 0; m < NS; ++m)
            {
               u[c][m] = row[c * NS + m];
               r[c][m] = 0.0;
            }

         // ---- cell term, assemble_explicit.cc:30-120: rhs_i += sum_q F(W_q).grad(phi_i) JxW (+ forcing) ----
#pragma unroll
         for (int q = 0; q < NQ; ++q)
         {
            double W[4], Fx[4], Fy[4];
#pragma unroll
            for (int c = 0; c < 4; ++c)
            {
               double s = u[c][0]; // phi_0 = 1
#pragma unroll
               for (int m = 1; m < NS; ++m) s = fma (T (tb, O_PHI + q * NS + m), u[c][m], s);
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
               for (int j = 0; j < N1; ++j)
#pragma unroll
                  for (int i = 0; i < N1; ++i)
                     if (i + j <= K) // d/dx of L_0(x) L_j(y) and d/dy of L_i(x) L_0(y) vanish
                     {
                        const int m = mode (i, j);
                        if (i > 0) r[c][m] = fma (ax, T (tb, O_DPX + q * NS + m), r[c][m]);
                        if (j > 0) r[c][m] = fma (ay, T (tb, O_DPY + q * NS + m), r[c][m]);
                     }
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
         lift<0> (tb, mom, lc, r);
         lift<1> (tb, mom, lc, r);
         lift<2> (tb, mom, lc, r);
         lift<3> (tb, mom, lc, r);

         // ---- M^-1 (orthonormal modes: 1/|K|, claw.cc:228-258) and Euler step (claw.cc:694-713); MODE_RHS: the residual ----
         if (A.mode == MODE_RHS)
         {
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
               for (int m = 0; m < NS; ++m) row[c * NS + m] = r[c][m];
            return;
         }
         const double invm = 1.0 / (hx * hy);
#pragma unroll
         for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int m = 0; m < NS; ++m) row[c * NS + m] = u[c][m] + dt * r[c][m] * invm;
      }
   };
}
