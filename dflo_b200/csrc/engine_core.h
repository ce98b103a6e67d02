// Orchestration of the explicit RK stage, independent of where the kernels execute.
// Engine<Backend> owns the device buffers and issues the kernels of kernels.cuh through the
// backend: CudaBackend (engine_cuda.cu, the product: CUDA streams, graphs, NCCL) or the CPU
// emulation backend of tests/emu (test infrastructure only; it is never part of the shipped
// library).  Mirrors ConservationLaw<dim>::iterate_explicit and compute_time_step
// (reference src/claw.cc:444-511, 725-772).
#pragma once

#include "../../include/dflo_b200.h"
#include "expr.h"
#include "kernels.cuh"
#include "cell_stage.cuh"
#include "mapped_stage.cuh"
#include "partition.h"
#include "tables.h"
#include "tables_pack.h"

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace dflo
{
   //---------------------------------------------------------------------------------------------
   // small thread-per-item kernels (thread(args, global_thread_index))
   //---------------------------------------------------------------------------------------------
   struct CellAverageKernel
   {
      typedef AvgArgs Args;
      static DFLO_DEV void thread (const Args &A, int j) { cell_average_thread (A, j); }
   };

   // Gather between the reference DoF layout (global cell order, optional deal.II dof map) and the
   // engine's local cell order.
   struct LayoutArgs
   {
      double *local;             // [n_cells_local][D]
      double *ext;               // staging copy of the caller's vector
      const int *l2g;            // local -> global cell
      const uint32_t *dof_map;   // optional
      int64_t ext_offset;        // ext[0] corresponds to reference position ext_offset (no dof_map)
      int n_cells, D, to_local;
   };
   struct LayoutKernel
   {
      typedef LayoutArgs Args;
      static DFLO_DEV void thread (const Args &A, int j)
      {
         if (j >= A.n_cells * A.D) return;
         const int l = j / A.D, i = j % A.D;
         const int64_t ref = (int64_t) A.l2g[l] * A.D + i;
         const int64_t e = A.dof_map ? (int64_t) A.dof_map[ref] : ref - A.ext_offset;
         if (A.to_local)
            A.local[j] = A.ext[e];
         else
            A.ext[e] = A.local[j];
      }
   };

   // rows of 4 (cell averages) or 1 (flags) between local and global cell order
   struct RowGatherArgs
   {
      const double *src;
      double *dst;
      const int *idx; // dst row r <- src row idx[r]
      int n_rows, width;
   };
   struct RowGatherKernel
   {
      typedef RowGatherArgs Args;
      static DFLO_DEV void thread (const Args &A, int j)
      {
         if (j >= A.n_rows * A.width) return;
         const int r = j / A.width, c = j % A.width;
         A.dst[j] = A.src[(size_t) A.idx[r] * A.width + c];
      }
   };

   // boundary values g(x,t) at the face quadrature points from the compiled expressions
   struct BcEvalArgs
   {
      double *bc_g;              // [n_bfaces][nqf][4]
      const int *bf_cell, *bf_face, *bf_id;
      const double *geom;
      const double *gx;          // Gauss nodes [nqf]
      const ExprInstr *code;     // all programs back to back
      const int *prog_start;     // [10*4][2] begin, end of each program; empty program => keep the stored value
      const double *time;        // [0] t, [1] dt
      int n_bfaces, nqf;
      int use_t_plus_dt;
      const double *verts;       // mapping = q1: cell vertices [n_local][8], else nullptr
   };
   struct BcEvalKernel
   {
      typedef BcEvalArgs Args;
      // one thread per (boundary face, face point, component): the four expressions of a point are interpreted side by side
      // (a time-dependent deck such as the double Mach reflection re-evaluates its boundary twice per step; one thread
      // walking through all four programs took 30 us per pass on cfg4)
      static DFLO_DEV void thread (const Args &A, int jc)
      {
         const int j = jc >> 2, c = jc & 3;
         if (j >= A.n_bfaces * A.nqf) return;
         const int bf = j / A.nqf, q = j % A.nqf;
         const int cell = A.bf_cell[bf], f = A.bf_face[bf], id = A.bf_id[bf];
         const double *g = A.geom + (size_t) cell * 4;
         const double s = A.gx[q];
         double x = g[0] + (f == 0 ? 0.0 : f == 1 ? 1.0 : s) * g[2];
         double y = g[1] + (f == 2 ? 0.0 : f == 3 ? 1.0 : s) * g[3];
         if (A.verts) // the bilinear map of the face point
         {
            const double *v = A.verts + (size_t) cell * 8;
            const double xi = f == 0 ? 0.0 : f == 1 ? 1.0 : s, eta = f == 2 ? 0.0 : f == 3 ? 1.0 : s;
            const double n0 = (1.0 - xi) * (1.0 - eta), n1 = xi * (1.0 - eta), n2 = (1.0 - xi) * eta, n3 = xi * eta;
            x = n0 * v[0] + n1 * v[2] + n2 * v[4] + n3 * v[6];
            y = n0 * v[1] + n1 * v[3] + n2 * v[5] + n3 * v[7];
         }
         const double t = A.time[0] + (A.use_t_plus_dt ? A.time[1] : 0.0);
         const int p0 = A.prog_start[2 * (id * 4 + c)], p1 = A.prog_start[2 * (id * 4 + c) + 1];
         if (p1 > p0) A.bc_g[(size_t) j * 4 + c] = expr_eval (A.code + p0, p1 - p0, x, y, t);
      }
   };

   // external force f_d(x,y) of the MPI tree at the cell quadrature points, from the two compiled
   // expressions, with t = 0 (src_mpi/assemble_explicit.cc:56-58: vector_value_list on a
   // FunctionParser whose time is never set)
   struct ExtForceEvalArgs
   {
      double *ext_force;         // [n_cells][n1*n1][2]
      const double *geom;
      const double *gx;          // Gauss nodes [n1]
      const ExprInstr *code;     // program of f_0 then f_1
      int len0, len1;
      int n_cells, n1;
   };
   struct ExtForceEvalKernel
   {
      typedef ExtForceEvalArgs Args;
      static DFLO_DEV void thread (const Args &A, int j)
      {
         const int nq = A.n1 * A.n1;
         if (j >= A.n_cells * nq) return;
         const int cell = j / nq, q = j % nq;
         const double *g = A.geom + (size_t) cell * 4;
         const double x = g[0] + A.gx[q % A.n1] * g[2];
         const double y = g[1] + A.gx[q / A.n1] * g[3];
         A.ext_force[(size_t) j * 2] = expr_eval (A.code, A.len0, x, y, 0.0);
         A.ext_force[(size_t) j * 2 + 1] = expr_eval (A.code + A.len0, A.len1, x, y, 0.0);
      }
   };

   // time scalars on the device: [0] t, [1] dt, [2] dt accumulator (min over cells), [3] final time
   struct DtArgs
   {
      const double *avg;
      const double *geom;
      double *time;
      int n_cells, degree;
      double cfl;
      // single-GPU: the block that retires last also finalises the step (DtFinalizeKernel's work),
      // saving a launch; sharded contexts reduce over the ranks in between and keep it separate
      unsigned int *done;
      int finalize, nblocks;
      double time_step;
      // mapping = q1: compute_time_step_q (claw.cc:518-557) from the solution itself
      const double *u, *verts, *dtq;
      int n1;
      // time step type = local: the per-cell values are kept, the minimum is neither capped nor clipped
      double *dt_cell;
   };
   struct DtKernel // phase kernel: per-cell dt, block minimum, one atomic per block
   {
      typedef DtArgs Args;
      static constexpr int THREADS = 256;
      static constexpr int MIN_BLOCKS = 1;
      static constexpr int NPHASE = 4;
      static constexpr int SMEM_DOUBLES = 2 * THREADS;
      // mapping = cartesian: a thread per cell.  mapping = q1 (compute_time_step_q): a thread per (cell, point) of the 4 x 4
      // equispaced points, 16 cells per block.
      static constexpr int QCELLS = THREADS / 16;
      static int grid (int n, bool mapped = false) { return mapped ? (n + QCELLS - 1) / QCELLS : (n + THREADS - 1) / THREADS; }
      static DFLO_DEV void phase (int p, const Args &A, double *sm, int tid, int bid)
      {
         const bool mapped = A.verts != nullptr;
         const int cell = mapped ? bid * QCELLS + tid / 16 : bid * THREADS + tid;
         if (p == 0)
         {
            double d = 1.0e20;
            if (mapped)
            {
               // |v| + c (equation.h:98-116) at point (jx, jy) = tid % 16 of QIterated(QTrapez,3); dtq[j][a] = l_a(j/3)
               d = 0.0;
               if (cell < A.n_cells)
               {
                  const int n1 = A.n1, nq = n1 * n1, jx = tid & 3, jy = (tid >> 2) & 3;
                  const double *uc = A.u + (size_t) cell * 4 * nq;
                  double W[4];
                  for (int c = 0; c < 4; ++c)
                  {
                     double s = 0.0;
                     for (int b = 0; b < n1; ++b)
                     {
                        double sb = 0.0;
                        for (int a = 0; a < n1; ++a) sb += A.dtq[jx * n1 + a] * uc[c * nq + a + n1 * b];
                        s += A.dtq[jy * n1 + b] * sb;
                     }
                     W[c] = s;
                  }
                  d = sqrt (W[0] * W[0] + W[1] * W[1]) / W[RHO] + sound_speed (W);
               }
            }
            else if (cell < A.n_cells)
            {
               d = cell_time_step (A.avg + (size_t) cell * 4, A.geom + (size_t) cell * 4, A.cfl, A.degree);
               // a NaN or non-positive cell value (negative density / pressure in the mean) must not poison the
               // block minimum, nor take part in the bit-pattern atomicMin below: the cell is skipped, as
               // `std::min (global_dt, dt(c))` skips a NaN in the reference (claw.cc:508)
               if (A.dt_cell) A.dt_cell[cell] = d; // dt(c), claw.cc:506 -- used as it is by solve(), :709
               if (!(d > 0.0)) d = 1.0e20;
            }
            sm[tid] = d;
         }
         else if (p == 1)
         {
            // mapping = q1: largest eigenvalue of the cell's 16 points -> dt(c) = cfl h / lambda / (2k+1), h = diameter / sqrt(2)
            if (mapped)
            {
               double d = 1.0e20;
               const int c = bid * QCELLS + tid;
               if (tid < QCELLS && c < A.n_cells)
               {
                  double lam = 0.0;
                  for (int i = 0; i < 16; ++i) lam = std_max (lam, sm[tid * 16 + i]); // std::max skips a NaN like the reference's loop
                  d = A.cfl * (q1_diameter (A.verts + (size_t) c * 8) / 1.4142135623730951) / lam / (2.0 * A.degree + 1.0);
                  if (A.dt_cell) A.dt_cell[c] = d;
                  if (!(d > 0.0)) d = 1.0e20;
               }
               sm[THREADS + tid] = d;
            }
         }
         else if (p == 2)
         {
            const double *v = sm + (mapped ? THREADS : 0);
            if (tid < 16)
            {
               double m = v[tid];
               for (int i = tid + 16; i < THREADS; i += 16) m = std_min (m, v[i]);
               sm[tid] = m; // mapped: sm[0..15] (the first cell's point values) are dead by now
            }
         }
         else if (tid == 0)
         {
            double m = sm[0];
            for (int i = 1; i < 16; ++i) m = std_min (m, sm[i]);
#if defined(__CUDA_ARCH__)
            // positive doubles order like their bit patterns
            atomicMin ((unsigned long long *) (A.time + 2), (unsigned long long) __double_as_longlong (m));
            bool last = false;
            if (A.finalize)
            {
               __threadfence ();
               last = atomicAdd (A.done, 1u) == (unsigned int) A.nblocks - 1;
               if (last)
               {
                  *A.done = 0;
                  __threadfence ();
               }
            }
            const double acc = last ? *reinterpret_cast<volatile double *> (A.time + 2) : 0.0;
#else
            if (m < A.time[2]) A.time[2] = m;
            const bool last = A.finalize && ++*A.done == (unsigned int) A.nblocks;
            if (last) *A.done = 0;
            const double acc = A.time[2];
#endif
            if (last) // claw.cc:468-476, as DtFinalizeKernel
            {
               double dt = acc;
               if (!A.dt_cell)
               {
               if (dt > 0 && A.time_step > 0) dt = std_min (dt, A.time_step);
               if (A.time[0] + dt > A.time[3]) dt = A.time[3] - A.time[0];
               }
               A.time[1] = dt;
               A.time[2] = 1.0e20;
            }
         }
      }
   };
   struct DtFinalizeArgs
   {
      double *time;
      double time_step;
      int local; // time step type = local: the minimum as it is (claw.cc:469: the capping block is for "global" only)
   };
   struct DtFinalizeKernel // claw.cc:468-476
   {
      typedef DtFinalizeArgs Args;
      static DFLO_DEV void thread (const Args &A, int j)
      {
         if (j != 0) return;
         double dt = A.time[2];
         if (!A.local)
         {
         if (dt > 0 && A.time_step > 0) dt = std_min (dt, A.time_step);
         if (A.time[0] + dt > A.time[3]) dt = A.time[3] - A.time[0];
         }
         A.time[1] = dt;
         A.time[2] = 1.0e20;
      }
   };
   struct FixedDtKernel // claw.cc:457-461: "time step" of the input file when cfl <= 0 (global time stepping)
   {
      typedef DtFinalizeArgs Args;
      static DFLO_DEV void thread (const Args &A, int j)
      {
         if (j != 0) return;
         // the reference leaves global_dt untouched in this branch; the sensible reading (and what the
         // host front end did in round 1) is dt = time_step, clipped so that t + dt <= final time
         double dt = A.time_step;
         if (A.time[0] + dt > A.time[3]) dt = A.time[3] - A.time[0];
         A.time[1] = dt;
         A.time[2] = 1.0e20;
      }
   };
   struct AdvanceTimeKernel // claw.cc:1072
   {
      typedef DtFinalizeArgs Args;
      static DFLO_DEV void thread (const Args &A, int j)
      {
         if (j == 0) A.time[0] += A.time[1];
      }
   };

   struct SumSqArgs
   {
      const double *v;
      double *out;
      int64_t n;
      int nblocks;
   };
   struct SumSqKernel // right_hand_side.l2_norm()^2, claw.cc:749
   {
      typedef SumSqArgs Args;
      static constexpr int THREADS = 256;
      static constexpr int MIN_BLOCKS = 1;
      static constexpr int NPHASE = 3;
      static constexpr int SMEM_DOUBLES = THREADS;
      static int grid (int64_t n) { int64_t g = (n + THREADS * 8 - 1) / (THREADS * 8); return (int) (g < 1 ? 1 : g > 1184 ? 1184 : g); }
      static DFLO_DEV void phase (int p, const Args &A, double *sm, int tid, int bid)
      {
         if (p == 0)
         {
            double s = 0.0;
            const int64_t stride = (int64_t) A.nblocks * THREADS;
            for (int64_t i = (int64_t) bid * THREADS + tid; i < A.n; i += stride) s += A.v[i] * A.v[i];
            sm[tid] = s;
         }
         else if (p == 1)
         {
            if (tid < 16)
            {
               double s = sm[tid];
               for (int i = tid + 16; i < THREADS; i += 16) s += sm[i];
               sm[tid] = s;
            }
         }
         else if (tid == 0)
         {
            double s = 0.0;
            for (int i = 0; i < 16; ++i) s += sm[i];
#if defined(__CUDA_ARCH__)
            atomicAdd (A.out, s);
#else
            *A.out += s;
#endif
         }
      }
   };

   struct PackArgs
   {
      const double *src;
      double *dst;
      const int *cells;
      int n_cells, width;
   };
   struct PackKernel // halo pack: dst[r][:] = src[cells[r]][:]
   {
      typedef PackArgs Args;
      static DFLO_DEV void thread (const Args &A, int j)
      {
         if (j >= A.n_cells * A.width) return;
         const int r = j / A.width, c = j % A.width;
         A.dst[j] = A.src[(size_t) A.cells[r] * A.width + c];
      }
   };

   //---------------------------------------------------------------------------------------------
   // compile-time dispatch over (basis, degree, flux)
   //---------------------------------------------------------------------------------------------
   template <class BK, int BASIS, int N1>
   struct StageDispatch
   {
      static void run (BK &bk, int flux, int n_tiles, const StageArgs &a)
      {
         switch (flux)
         {
            case FLUX_LXF: bk.template launch_stage<StageKernel<BASIS, N1, FLUX_LXF>> (n_tiles, a); break;
            case FLUX_SW: bk.template launch_stage<StageKernel<BASIS, N1, FLUX_SW>> (n_tiles, a); break;
            case FLUX_KFVS: bk.template launch_stage<StageKernel<BASIS, N1, FLUX_KFVS>> (n_tiles, a); break;
            case FLUX_ROE: bk.template launch_stage<StageKernel<BASIS, N1, FLUX_ROE>> (n_tiles, a); break;
            case FLUX_KEP: bk.template launch_stage<StageKernel<BASIS, N1, FLUX_KEP>> (n_tiles, a); break;
            default: bk.template launch_stage<StageKernel<BASIS, N1, FLUX_HLLC>> (n_tiles, a); break;
         }
      }
   };

   template <class BK>
   void launch_stage (BK &bk, int basis, int n1, int flux, int n_tiles, const StageArgs &a)
   {
      if (n1 == 1) // degree 0: Q0 == P0
         StageDispatch<BK, BASIS_QK, 1>::run (bk, flux, n_tiles, a);
      else if (basis == BASIS_QK)
      {
         if (n1 == 2) StageDispatch<BK, BASIS_QK, 2>::run (bk, flux, n_tiles, a);
         else if (n1 == 3) StageDispatch<BK, BASIS_QK, 3>::run (bk, flux, n_tiles, a);
         else if (n1 == 4) StageDispatch<BK, BASIS_QK, 4>::run (bk, flux, n_tiles, a);
         else StageDispatch<BK, BASIS_QK, 5>::run (bk, flux, n_tiles, a);
      }
      else
      {
         if (n1 == 2) StageDispatch<BK, BASIS_PK, 2>::run (bk, flux, n_tiles, a);
         else if (n1 == 3) StageDispatch<BK, BASIS_PK, 3>::run (bk, flux, n_tiles, a);
         else StageDispatch<BK, BASIS_PK, 4>::run (bk, flux, n_tiles, a);
      }
   }

   // thread-per-cell stage kernel of the Pk basis (cell_stage.cuh)
   template <class BK, int N1>
   void launch_cell_stage_n (BK &bk, int flux, const CellStageArgs &a)
   {
      switch (flux)
      {
         case FLUX_LXF: bk.template launch<PkCellStageKernel<N1, FLUX_LXF>> (PkCellStageKernel<N1, FLUX_LXF>::grid (a.n_compute), a); break;
         case FLUX_SW: bk.template launch<PkCellStageKernel<N1, FLUX_SW>> (PkCellStageKernel<N1, FLUX_LXF>::grid (a.n_compute), a); break;
         case FLUX_KFVS: bk.template launch<PkCellStageKernel<N1, FLUX_KFVS>> (PkCellStageKernel<N1, FLUX_LXF>::grid (a.n_compute), a); break;
         case FLUX_ROE: bk.template launch<PkCellStageKernel<N1, FLUX_ROE>> (PkCellStageKernel<N1, FLUX_LXF>::grid (a.n_compute), a); break;
         case FLUX_KEP: bk.template launch<PkCellStageKernel<N1, FLUX_KEP>> (PkCellStageKernel<N1, FLUX_LXF>::grid (a.n_compute), a); break;
         default: bk.template launch<PkCellStageKernel<N1, FLUX_HLLC>> (PkCellStageKernel<N1, FLUX_LXF>::grid (a.n_compute), a); break;
      }
   }
   template <class BK>
   void launch_cell_stage (BK &bk, int n1, int flux, const CellStageArgs &a)
   {
      // P1 and P2; P3 (40 coefficients + 40 residuals per thread) does not fit the register file and stays on the tile kernel
      if (n1 == 2) launch_cell_stage_n<BK, 2> (bk, flux, a);
      else launch_cell_stage_n<BK, 3> (bk, flux, a);
   }

   // mapped (mapping = q1) stage kernel, Qk (mapped_stage.cuh)
   template <class BK, int N1>
   void launch_mapped_stage_n (BK &bk, int flux, const MappedStageArgs &a)
   {
      switch (flux)
      {
         case FLUX_LXF: bk.template launch<MappedStageKernel<N1, FLUX_LXF>> (MappedStageKernel<N1, FLUX_LXF>::grid (a.n_compute), a); break;
         case FLUX_SW: bk.template launch<MappedStageKernel<N1, FLUX_SW>> (MappedStageKernel<N1, FLUX_LXF>::grid (a.n_compute), a); break;
         case FLUX_KFVS: bk.template launch<MappedStageKernel<N1, FLUX_KFVS>> (MappedStageKernel<N1, FLUX_LXF>::grid (a.n_compute), a); break;
         case FLUX_ROE: bk.template launch<MappedStageKernel<N1, FLUX_ROE>> (MappedStageKernel<N1, FLUX_LXF>::grid (a.n_compute), a); break;
         case FLUX_KEP: bk.template launch<MappedStageKernel<N1, FLUX_KEP>> (MappedStageKernel<N1, FLUX_LXF>::grid (a.n_compute), a); break;
         default: bk.template launch<MappedStageKernel<N1, FLUX_HLLC>> (MappedStageKernel<N1, FLUX_LXF>::grid (a.n_compute), a); break;
      }
   }
   template <class BK>
   void launch_mapped_stage (BK &bk, int n1, int flux, const MappedStageArgs &a)
   {
      if (n1 == 1) launch_mapped_stage_n<BK, 1> (bk, flux, a);
      else if (n1 == 2) launch_mapped_stage_n<BK, 2> (bk, flux, a);
      else if (n1 == 3) launch_mapped_stage_n<BK, 3> (bk, flux, a);
      else if (n1 == 4) launch_mapped_stage_n<BK, 4> (bk, flux, a);
      else launch_mapped_stage_n<BK, 5> (bk, flux, a);
   }

   template <class BK>
   void launch_limiter (BK &bk, int basis, int n1, const LimiterArgs &a)
   {
      if (n1 == 1) return;
      // one thread per cell (LimiterCellKernel); DFLO_B200_LIMITER=block selects the block-per-cells form
#define DFLO_LIM(B, N)                                                                          \
   do                                                                                           \
   {                                                                                            \
      if (bk.limiter_block_form ())                                                             \
         bk.template launch<LimiterKernel<B, N>> (LimiterKernel<B, N>::grid (a.n_compute), a);  \
      else                                                                                      \
         bk.template launch<LimiterCellKernel<B, N>> (LimiterCellKernel<B, N>::grid (a.n_compute), a); \
   } while (0)
      if (basis == BASIS_QK && a.tvb == 2) // minmax limiter (src_mpi/limiter.cc:400-553): instantiations of their own
      {
         if (n1 == 2) bk.template launch<LimiterCellKernel<BASIS_QK, 2, 1>> (LimiterCellKernel<BASIS_QK, 2, 1>::grid (a.n_compute), a);
         else if (n1 == 3) bk.template launch<LimiterCellKernel<BASIS_QK, 3, 1>> (LimiterCellKernel<BASIS_QK, 3, 1>::grid (a.n_compute), a);
         else if (n1 == 4) bk.template launch<LimiterCellKernel<BASIS_QK, 4, 1>> (LimiterCellKernel<BASIS_QK, 4, 1>::grid (a.n_compute), a);
         else bk.template launch<LimiterCellKernel<BASIS_QK, 5, 1>> (LimiterCellKernel<BASIS_QK, 5, 1>::grid (a.n_compute), a);
      }
      else if (basis == BASIS_QK)
      {
         if (n1 == 2) DFLO_LIM (BASIS_QK, 2);
         else if (n1 == 3) DFLO_LIM (BASIS_QK, 3);
         else if (n1 == 4) DFLO_LIM (BASIS_QK, 4);
         else DFLO_LIM (BASIS_QK, 5);
      }
      else
      {
         if (n1 == 2) DFLO_LIM (BASIS_PK, 2);
         else if (n1 == 3) DFLO_LIM (BASIS_PK, 3);
         else DFLO_LIM (BASIS_PK, 4);
      }
#undef DFLO_LIM
   }

   template <class BK>
   void launch_indicator (BK &bk, int basis, int n1, const IndicatorArgs &a)
   {
      if (n1 == 1) return;
#define DFLO_IND(B, N) bk.template launch1d<IndicatorKernel<B, N>> (a.n_cells, a)
      if (basis == BASIS_QK)
      {
         if (n1 == 2) DFLO_IND (BASIS_QK, 2);
         else if (n1 == 3) DFLO_IND (BASIS_QK, 3);
         else if (n1 == 4) DFLO_IND (BASIS_QK, 4);
         else DFLO_IND (BASIS_QK, 5);
      }
      else
      {
         if (n1 == 2) DFLO_IND (BASIS_PK, 2);
         else if (n1 == 3) DFLO_IND (BASIS_PK, 3);
         else DFLO_IND (BASIS_PK, 4);
      }
#undef DFLO_IND
   }

   //---------------------------------------------------------------------------------------------
   // The engine
   //---------------------------------------------------------------------------------------------
   template <class BK>
   class Engine
   {
   public:
      BK bk;
      dflo_params prm;
      FeTables tab;
      LocalMesh lm;
      std::string error;
      int n_rk;
      double ark[3];

      // device state
      double *U[3] = {nullptr, nullptr, nullptr};
      double *AVG[3] = {nullptr, nullptr, nullptr};
      int cur = 0, old = 0;
      double *rhs = nullptr;
      double *d_time = nullptr;       // t, dt, dt accumulator, final time
      double *d_scratch = nullptr;    // [1] reductions
      int *d_nbr = nullptr, *d_halo_cells = nullptr, *d_rowdesc = nullptr, *d_send_entries = nullptr;
      FaceJob *d_jobs = nullptr;
      TileDesc *d_tiles = nullptr;
      unsigned char *d_fflags = nullptr;
      bool pk_cell_ok = false; // the mesh admits the thread-per-cell Pk stage kernel (cell_stage.cuh: pk_cell_mesh_ok)
      double *d_ext_force = nullptr; // [n_local][n_q][2], allocated by set_external_force
      int *d_hang_of = nullptr, *d_hang = nullptr; // faces with hanging nodes (LocalMesh::hang_of, hang)
      double *d_dt_cell = nullptr; // [n_local] dt(cell) of time step type = local
      double *d_verts = nullptr, *d_dtq = nullptr; // mapping = q1: cell vertices [n_local][8], l_a(j/3) [4][n1]
      unsigned char *d_nbr_face = nullptr;
      double *d_geom = nullptr, *d_bc_g = nullptr, *d_stage_tab = nullptr, *d_lim_tab = nullptr, *d_gw = nullptr, *d_gx = nullptr;
      int *d_bkind = nullptr, *d_bf_cell = nullptr, *d_bf_face = nullptr, *d_bf_id = nullptr, *d_l2g = nullptr, *d_flags = nullptr;
      unsigned int *d_err = nullptr;
      double *d_shock = nullptr;       // shock_indicator per local cell
      ExprInstr *d_code = nullptr;
      int *d_prog_start = nullptr;
      int *d_prog_start_t = nullptr;  // same table with the time-independent programs emptied: the per-stage refresh
      double *d_ext = nullptr;        // staging for set/get_solution
      size_t ext_capacity = 0;
      uint32_t *d_dofmap = nullptr;
      size_t dofmap_capacity = 0;
      // halo
      std::vector<int *> d_send_cells[2];
      std::vector<double *> d_send_u[2], d_send_avg[2];
      // boundary expressions
      std::vector<ExprInstr> programs[DFLO_MAX_BOUNDARIES][4];
      bool have_programs = false, programs_time_dependent = false;
      bool program_uses_t[DFLO_MAX_BOUNDARIES][4] = {};
      int n_global_bfaces = 0;

      // halo exchange fused into the stage kernel: row kernel, peer memory mapped, nothing between
      // the stage kernel and the exchange (no limiter)
      bool mapped () const { return prm.mapping == DFLO_MAPPING_Q1; }
      // meshes with hanging nodes run through the mapped stage kernel whatever the mapping (it is the one that knows sub-faces)
      bool hanging = false;
      bool general_kernel () const { return mapped () || hanging; }
      // the register-blocked Qk kernel serves mapping = cartesian on the device
      bool row_kernel () { return !general_kernel () && bk.use_row_kernel (tab.basis, tab.n1); }
      bool fused_halo () { return !lm.peers.empty () && !tvb () && !pos () && row_kernel () && bk.p2p_fused_ok (); }
      // the exchange follows the limiter as a kernel of its own, and every reader of ghost cells inside a step is a row
      // stage kernel: the exchange does not wait for the peers, the ghost-reading tiles of the next stage kernel do
      bool deferred_halo () { return !lm.peers.empty () && (tvb () || pos ()) && !kxrcf () && row_kernel () && bk.p2p_deferred_ok (); }
      int D () const { return tab.D; }
      // a slope limiter that reads the neighbours' new means runs after the stage kernel: TVB, or the
      // minmax limiter of the MPI tree (src_mpi/limiter.cc:36-70)
      bool tvb () const { return (prm.limiter_type == DFLO_LIMITER_TVB || prm.limiter_type == DFLO_LIMITER_MINMAX) && tab.k > 0; }
      bool pos () const { return prm.pos_lim && tab.k > 0; }
      bool kxrcf () const { return tvb () && prm.shock_indicator != DFLO_INDICATOR_LIMITER; }

      int init (const dflo_flat_mesh &mesh, const dflo_params &p, int rank, int world)
      {
         prm = p;
         if (p.basis != DFLO_BASIS_QK && p.basis != DFLO_BASIS_PK) return fail (DFLO_E_INVALID, "unknown basis");
         if (!build_tables (p.basis, p.degree, tab)) return fail (DFLO_E_UNSUPPORTED, "degree out of range (Qk 0..4, Pk 0..3)");
         if (p.flux_type < 0 || p.flux_type > 5) return fail (DFLO_E_INVALID, "unknown flux");
         if (p.shock_indicator < 0 || p.shock_indicator > 2) return fail (DFLO_E_INVALID, "unknown shock indicator");
         if (p.limiter_type < DFLO_LIMITER_NONE || p.limiter_type > DFLO_LIMITER_MINMAX) return fail (DFLO_E_INVALID, "unknown limiter type");
         if (p.limiter_type == DFLO_LIMITER_MINMAX && p.basis != DFLO_BASIS_QK) // src_mpi/parameters.cc:610-611
            return fail (DFLO_E_UNSUPPORTED, "minmax limiter is implemented only for Qk");
         if (mesh.n_cells <= 0) return fail (DFLO_E_INVALID, "empty mesh");
         if (p.mapping != DFLO_MAPPING_CARTESIAN && p.mapping != DFLO_MAPPING_Q1) return fail (DFLO_E_UNSUPPORTED, "mapping: cartesian or q1");
         hanging = mesh.n_hanging_faces > 0;
         if (hanging)
         {
            // the TVB / minmax limiters' neighbour lists are same-level (claw.cc:336-380 asserts level or level - 1 but the
            // limiters are not restated for it); the Qk basis only (as under mapping = q1).  The positivity limiter is local
            // to a cell and runs as it is
            if (p.basis != DFLO_BASIS_QK || p.limiter_type != DFLO_LIMITER_NONE)
               return fail (DFLO_E_UNSUPPORTED, "faces with hanging nodes: Qk basis without the TVB / minmax limiter only");
            if (!mesh.cell_vertices || !mesh.neighbor_face || !mesh.hanging) return fail (DFLO_E_INVALID, "hanging nodes need cell_vertices, neighbor_face and the hanging table");
         }
         if (p.mapping == DFLO_MAPPING_Q1)
         {
            // src/parameters.cc:545-549: TVB and Pk need Cartesian grids.  The positivity limiter runs on mapped cells as it is:
            // positivity.cc:46-47 evaluates the solution at GLL x Gauss points of the UNIT cell (FEValues with update_values
            // only) and scales about cell_average, which compute_cell_average took with the mapped JxW (claw.cc:562-597)
            if (p.basis != DFLO_BASIS_QK) return fail (DFLO_E_UNSUPPORTED, "mapping = q1: Pk basis can only be used with Cartesian grids");
            if (p.limiter_type != DFLO_LIMITER_NONE) return fail (DFLO_E_UNSUPPORTED, "mapping = q1: TVB limiter works on cartesian grids only");
            if (!mesh.cell_vertices || !mesh.neighbor_face) return fail (DFLO_E_INVALID, "mapping = q1 needs cell_vertices and neighbor_face in the flat mesh");
         }
         else
         {
            // mapping = cartesian: every neighbour sits on the opposite face and runs along it in the same direction,
            // and (when the vertices are given) every cell is the rectangle cell_origin / cell_size describe
            for (int c = 0; c < mesh.n_cells; ++c)
               for (int f = 0; f < 4; ++f)
               {
                  if (mesh.neighbor[4 * (size_t) c + f] < 0) continue;
                  const int fl = mesh.face_flags[4 * (size_t) c + f];
                  if ((mesh.neighbor_face && mesh.neighbor_face[4 * (size_t) c + f] != (f ^ 1)) || ((fl & DFLO_FACE_FLIP) && !(fl & DFLO_FACE_PERIODIC)))
                     return fail (DFLO_E_UNSUPPORTED, "neighbouring cells are not equally oriented: use mapping = q1");
               }
            if (mesh.cell_vertices)
               for (int c = 0; c < mesh.n_cells; ++c)
               {
                  const double *q = mesh.cell_vertices + 8 * (size_t) c, *o = mesh.cell_origin + 2 * (size_t) c, *h = mesh.cell_size + 2 * (size_t) c;
                  const double tol = 1e-12 * (fabs (h[0]) + fabs (h[1]));
                  const bool rect = fabs (q[0] - o[0]) <= tol && fabs (q[1] - o[1]) <= tol && fabs (q[2] - (o[0] + h[0])) <= tol && fabs (q[3] - o[1]) <= tol
                                    && fabs (q[4] - o[0]) <= tol && fabs (q[5] - (o[1] + h[1])) <= tol && fabs (q[6] - (o[0] + h[0])) <= tol
                                    && fabs (q[7] - (o[1] + h[1])) <= tol;
                  if (!rect) return fail (DFLO_E_UNSUPPORTED, "cell is not an axis-aligned rectangle: use mapping = q1");
               }
         }
         // claw.cc:457-461: cfl <= 0 selects the fixed time step of the input file, which must then be given
         if (!(p.cfl > 0.0) && !(p.time_step > 0.0)) return fail (DFLO_E_INVALID, "cfl <= 0 needs a positive time step");
         for (int b = 0; b < mesh.n_boundary_faces; ++b)
            if (mesh.bface_id[b] < 0 || mesh.bface_id[b] >= DFLO_MAX_BOUNDARIES) return fail (DFLO_E_INVALID, "boundary id out of range");
         // claw.cc:141-159
         if (tab.k == 0) { n_rk = 1; ark[0] = 0.0; }
         else if (tab.k == 1) { n_rk = 2; ark[0] = 0.0; ark[1] = 0.5; }
         else { n_rk = 3; ark[0] = 0.0; ark[1] = 3.0 / 4.0; ark[2] = 1.0 / 3.0; }
         const int layers = tvb () ? 2 : 1;
         std::string e;
         // stage-kernel flavour decides the tile shape: register-blocked row kernel (Qk, CUDA) or the
         // generic phase kernel
         const bool row = row_kernel ();
         const int tx = row ? row_tx (tab.n1) : tile_nx (tab.n1), ty = row ? row_ty (tab.n1) : tile_ny (tab.n1);
         // tiles on the partition cut come first (their cells travel while the interior is worked on); last only in the
         // deferred-wait experiment, where they are the ones that wait for the peers
         if (!build_local_mesh (mesh, rank, world, layers, tx, ty, lm, e, row, !((tvb () || pos ()) && BK::p2p_defer_requested ()))) return fail (DFLO_E_INVALID, e);
         // the 1-D kernels (layout, pack, cell averages) index DoFs with 32-bit ints
         if ((int64_t) lm.n_local * D () > (int64_t) 0x7fffffff)
            return fail (DFLO_E_UNSUPPORTED, "more than 2^31-1 DoFs on one rank: shard the mesh over more GPUs");
         bk.prepare_tables (tab, pack_stage_tables (tab));
         if (row) d_rowdesc = upload (lm.rowdesc);
         n_global_bfaces = mesh.n_boundary_faces;

         const size_t nd = (size_t) lm.n_local * D ();
         for (int i = 0; i < 3; ++i)
         {
            U[i] = bk.template alloc<double> (nd);
            AVG[i] = bk.template alloc<double> ((size_t) lm.n_local * 4);
            bk.zero (U[i], nd * sizeof (double));
            bk.zero (AVG[i], (size_t) lm.n_local * 4 * sizeof (double));
         }
         if (prm.local_time_step)
         {
            if (prm.cfl <= 0.0) return fail (DFLO_E_INVALID, "time step type = local needs a cfl");
            d_dt_cell = bk.template alloc<double> (lm.n_local);
            bk.zero (d_dt_cell, (size_t) lm.n_local * sizeof (double));
         }
         d_time = bk.template alloc<double> (4);
         const double t0[4] = {0.0, 0.0, 1.0e20, 1.0e20};
         bk.h2d (d_time, t0, sizeof (t0));
         d_scratch = bk.template alloc<double> (4); // [0] reductions, [2] block counter of the dt kernel
         bk.zero (d_scratch, 4 * sizeof (double));
         d_nbr = upload (lm.nbr);
         d_halo_cells = upload (pad1 (lm.halo_cells));
         d_jobs = upload_aligned_jobs ();
         {
            std::vector<TileDesc> td (std::max (1, lm.n_tiles));
            for (int t = 0; t < lm.n_tiles; ++t)
            {
               td[t].c0 = lm.tile_start[t];
               td[t].ncb = lm.tile_start[t + 1] - lm.tile_start[t];
               td[t].h0 = lm.halo_start[t];
               td[t].nh = lm.halo_start[t + 1] - lm.halo_start[t];
               td[t].j0 = lm.job_start[t];
               td[t].nj = lm.job_start[t + 1] - lm.job_start[t];
               td[t].pad0 = t >= lm.n_tiles_owned; // ghost tile: means only
               td[t].pad1 = 0;
            }
            d_tiles = upload (td);
         }
         d_fflags = upload (lm.fflags);
         pk_cell_ok = tab.basis == BASIS_PK && pk_cell_mesh_ok (lm.nbr.data (), lm.fflags.data (), lm.n_compute);
         d_geom = upload (lm.geom);
         d_l2g = upload (lm.l2g);
         if (hanging)
         {
            d_hang_of = upload (lm.hang_of);
            d_hang = upload (lm.hang);
         }
         if (general_kernel ())
         {
            d_verts = upload (lm.verts);
            d_nbr_face = upload (lm.nbr_face);
            // l_a(j/3), j = 0..3: the solution at the points of QIterated(QTrapez,3) (compute_time_step_q, claw.cc:522)
            std::vector<double> dtq (4 * (size_t) tab.n1);
            for (int j = 0; j < 4; ++j)
               for (int a = 0; a < tab.n1; ++a)
               {
                  double l = 1.0;
                  for (int m = 0; m < tab.n1; ++m)
                     if (m != a) l *= (j / 3.0 - tab.gx[m]) / (tab.gx[a] - tab.gx[m]);
                  dtq[(size_t) j * tab.n1 + a] = l;
               }
            d_dtq = upload (dtq);
         }
         std::vector<int> kinds (std::max<size_t> (1, lm.bf_id.size ()), 0);
         for (size_t b = 0; b < lm.bf_id.size (); ++b) kinds[b] = prm.bc_kind[lm.bf_id[b]];
         d_bkind = upload (kinds);
         d_bf_cell = upload (pad1 (lm.bf_cell));
         d_bf_face = upload (pad1 (lm.bf_face));
         d_bf_id = upload (pad1 (lm.bf_id));
         const size_t ng = std::max<size_t> (1, lm.bf_id.size ()) * tab.n1 * 4;
         d_bc_g = bk.template alloc<double> (ng);
         bk.zero (d_bc_g, ng * sizeof (double));
         d_stage_tab = upload (pack_stage_tables (tab));
         d_lim_tab = upload (pack_limiter_tables (tab));
         d_gw = upload (std::vector<double> (tab.gw, tab.gw + tab.n1));
         d_gx = upload (std::vector<double> (tab.gx, tab.gx + tab.n1));
         d_shock = bk.template alloc<double> (lm.n_local);
         {
            std::vector<double> all ((size_t) lm.n_local, 1.0e20); // indicator.cc:18-22
            bk.h2d (d_shock, all.data (), all.size () * sizeof (double));
            bk.sync ();
         }
         d_flags = bk.template alloc<int> (lm.n_local);
         bk.zero (d_flags, lm.n_local * sizeof (int));
         d_err = bk.template alloc<unsigned int> (1);
         bk.zero (d_err, sizeof (unsigned int));
         d_prog_start = bk.template alloc<int> (DFLO_MAX_BOUNDARIES * 4 * 2);
         d_prog_start_t = bk.template alloc<int> (DFLO_MAX_BOUNDARIES * 4 * 2);
         for (auto &pr : lm.peers)
            for (int k = 0; k < 2; ++k)
            {
               d_send_cells[k].push_back (upload (pad1 (pr.send_cells[k])));
               d_send_u[k].push_back (bk.template alloc<double> (std::max<size_t> (1, pr.send_cells[k].size ()) * D ()));
               d_send_avg[k].push_back (bk.template alloc<double> (std::max<size_t> (1, pr.send_cells[k].size ()) * 4));
            }
         // direct peer-memory halo (NVLink P2P) when the backend can map the peers' buffers
         if (!lm.peers.empty ())
         {
            d_send_entries = upload (pad1 (lm.send_entries));
            bk.p2p_setup (U, AVG, lm.peers, d_send_cells, D (), d_send_entries, lm.n_send_tiles);
         }
         return bk.check (error);
      }

      void release ()
      {
         bk.sync ();
         bk.p2p_teardown ();
         bk.drop_graphs ();
         for (int i = 0; i < 3; ++i)
         {
            bk.free (U[i]);
            bk.free (AVG[i]);
         }
         void *ptrs[] = {d_hang_of, d_hang, d_dt_cell, d_verts, d_dtq, d_nbr_face, d_ext_force, rhs, d_time, d_scratch, d_nbr, d_fflags, d_geom, d_bc_g, d_stage_tab, d_lim_tab, d_gw, d_gx, d_bkind, d_bf_cell,
                         d_bf_face, d_bf_id, d_l2g, d_flags, d_err, d_shock, d_code, d_prog_start, d_prog_start_t, d_ext, d_dofmap, d_halo_cells, d_jobs, d_tiles, d_rowdesc, d_send_entries};
         for (void *p : ptrs) bk.free (p);
         for (int k = 0; k < 2; ++k)
         {
            for (auto p : d_send_cells[k]) bk.free (p);
            for (auto p : d_send_u[k]) bk.free (p);
            for (auto p : d_send_avg[k]) bk.free (p);
         }
      }

      //------------------------------------------------------------------------------------------
      int set_solution (const double *u, const uint32_t *dof_map, size_t n)
      {
         if (n != (size_t) lm.n_global * D ()) return fail (DFLO_E_INVALID, "set_solution: wrong vector length");
         bk.halo_wait ();
         int rc = stage_external (u, dof_map, n, true);
         if (rc) return rc;
         layout (U[cur], dof_map != nullptr, true);
         compute_cell_average (cur, lm.n_owned);
         exchange_halo (cur);
         old = cur;
         // a step graph is a function of the starting buffer alone, on sharded contexts too: the exchange epochs are
         // device-side counters (p2p_halo.cuh), advanced by the exchange above on every rank alike, so the captured
         // steps stay valid (DFLO_B200_KEEP_GRAPHS=0 restores the r01 behaviour of capturing again)
         if (!lm.peers.empty () && !bk.keep_graphs_sharded ()) bk.drop_graphs ();
         return bk.check (error);
      }

      int get_solution (double *u, const uint32_t *dof_map, size_t n) { return get_vector (U[cur], u, dof_map, n); }

      int get_rhs (double *r, const uint32_t *dof_map, size_t n)
      {
         if (!rhs) return fail (DFLO_E_INVALID, "get_rhs before assemble_rhs");
         return get_vector (rhs, r, dof_map, n);
      }

      int get_cell_average (double *avg)
      {
         bk.halo_wait ();
         bk.sync ();
         std::vector<double> loc ((size_t) lm.n_owned * 4);
         bk.d2h (loc.data (), AVG[cur], loc.size () * sizeof (double));
         for (int l = 0; l < lm.n_owned; ++l)
            for (int c = 0; c < 4; ++c) avg[(size_t) lm.l2g[l] * 4 + c] = loc[(size_t) l * 4 + c];
         return bk.check (error);
      }

      int get_shock_indicator (double *ind)
      {
         bk.sync ();
         std::vector<double> loc (lm.n_owned);
         bk.d2h (loc.data (), d_shock, loc.size () * sizeof (double));
         for (int l = 0; l < lm.n_owned; ++l) ind[lm.l2g[l]] = loc[l];
         return bk.check (error);
      }

      int get_limited_flags (int32_t *flags)
      {
         bk.sync ();
         std::vector<int> loc (lm.n_owned);
         bk.d2h (loc.data (), d_flags, loc.size () * sizeof (int));
         for (int l = 0; l < lm.n_owned; ++l) flags[lm.l2g[l]] = loc[l];
         return bk.check (error);
      }

      int commit_step ()
      {
         old = cur;
         return DFLO_OK;
      }

      //------------------------------------------------------------------------------------------
      int set_boundary_values (const double *g)
      {
         const int nqf = tab.n1;
         std::vector<double> loc (std::max<size_t> (1, lm.bf_global.size ()) * nqf * 4, 0.0);
         for (size_t b = 0; b < lm.bf_global.size (); ++b)
            std::memcpy (&loc[b * nqf * 4], g + (size_t) lm.bf_global[b] * nqf * 4, sizeof (double) * nqf * 4);
         bk.h2d (d_bc_g, loc.data (), loc.size () * sizeof (double));
         return bk.check (error);
      }

      int set_boundary_expression (int id, int comp, const char *text)
      {
         if (id < 0 || id >= DFLO_MAX_BOUNDARIES || comp < 0 || comp > 3) return fail (DFLO_E_INVALID, "boundary id/component out of range");
         ExprCompiler cc;
         std::string e;
         std::vector<ExprInstr> code;
         bool uses_t = false;
         if (!cc.compile (text, code, e, &uses_t)) return fail (DFLO_E_EXPR, e);
         programs[id][comp] = code;
         program_uses_t[id][comp] = uses_t;
         // flatten all programs; the kernel reads program k as code[start[2k] .. start[2k+1])
         std::vector<ExprInstr> all;
         std::vector<int> start (DFLO_MAX_BOUNDARIES * 4 * 2, 0), start_t (DFLO_MAX_BOUNDARIES * 4 * 2, 0);
         for (int b = 0; b < DFLO_MAX_BOUNDARIES; ++b)
            for (int c = 0; c < 4; ++c)
            {
               const int k = b * 4 + c;
               start[2 * k] = start_t[2 * k] = all.size ();
               all.insert (all.end (), programs[b][c].begin (), programs[b][c].end ());
               start[2 * k + 1] = all.size ();
               start_t[2 * k + 1] = program_uses_t[b][c] ? (int) all.size () : start_t[2 * k]; // empty: keep the stored value
            }
         bk.sync ();
         bk.free (d_code);
         d_code = bk.template alloc<ExprInstr> (std::max<size_t> (1, all.size ()));
         bk.h2d (d_code, all.data (), all.size () * sizeof (ExprInstr));
         bk.h2d (d_prog_start, start.data (), start.size () * sizeof (int));
         bk.h2d (d_prog_start_t, start_t.data (), start_t.size () * sizeof (int));
         bk.sync ();
         have_programs = true;
         programs_time_dependent = programs_time_dependent || uses_t;
         bk.drop_graphs ();
         eval_boundary (false); // values at the current time
         return bk.check (error);
      }

      // "f_0 value" / "f_1 value" of the MPI tree (src_mpi/parameters.cc:355-360, 488-497): the external
      // force replaces the hard-wired (0,-1) of src/ in the forcing term gravity * G
      // (src_mpi/equation.h:1189-1202, src_mpi/assemble_explicit.cc:84, 108-111)
      int set_external_force (const char *fx, const char *fy)
      {
         ExprCompiler cc;
         std::string e;
         std::vector<ExprInstr> c0, c1;
         if (!cc.compile (fx, c0, e, nullptr)) return fail (DFLO_E_EXPR, e);
         if (!cc.compile (fy, c1, e, nullptr)) return fail (DFLO_E_EXPR, e);
         std::vector<ExprInstr> all (c0);
         all.insert (all.end (), c1.begin (), c1.end ());
         bk.sync ();
         const size_t nq = (size_t) tab.n1 * tab.n1;
         if (!d_ext_force) d_ext_force = bk.template alloc<double> ((size_t) lm.n_local * nq * 2);
         ExprInstr *d_fcode = bk.template alloc<ExprInstr> (std::max<size_t> (1, all.size ()));
         bk.h2d (d_fcode, all.data (), all.size () * sizeof (ExprInstr));
         ExtForceEvalArgs a;
         a.ext_force = d_ext_force;
         a.geom = d_geom;
         a.gx = d_gx;
         a.code = d_fcode;
         a.len0 = c0.size ();
         a.len1 = c1.size ();
         a.n_cells = lm.n_local;
         a.n1 = tab.n1;
         bk.template launch1d<ExtForceEvalKernel> (a.n_cells * (int) nq, a);
         bk.sync ();
         bk.free (d_fcode);
         bk.drop_graphs ();
         return bk.check (error);
      }

      //------------------------------------------------------------------------------------------
      // assemble_system, assemble_explicit.cc:433-452
      int assemble_rhs (double t_bc)
      {
         bk.halo_wait ();
         if (!rhs)
         {
            rhs = bk.template alloc<double> ((size_t) lm.n_local * D ());
            bk.zero (rhs, (size_t) lm.n_local * D () * sizeof (double));
         }
         set_time (0, t_bc);
         eval_boundary (false);
         StageArgs a = stage_args (0, MODE_RHS);
         a.out = rhs;
         if (fused_halo ()) a.fx = bk.p2p_fused_args (cur); // waits for the last fused exchange, publishes nothing
         run_stage (a, true);
         return bk.check (error);
      }

      // one pass of the rk loop body, claw.cc:747-766, with host-supplied dt and BC time
      int rk_stage (int rk, double t_bc, double dt, double *res_norm)
      {
         if (rk < 0 || rk >= n_rk) return fail (DFLO_E_INVALID, "rk out of range");
         bk.halo_wait ();
         if (res_norm)
         {
            int rc = assemble_rhs (t_bc);
            if (rc) return rc;
            const double z = 0.0;
            bk.h2d (d_scratch, &z, sizeof (double));
            SumSqArgs s;
            s.v = rhs;
            s.out = d_scratch;
            s.n = (int64_t) lm.n_owned * D ();
            s.nblocks = SumSqKernel::grid (s.n);
            bk.template launch<SumSqKernel> (s.nblocks, s);
            bk.allreduce_sum (d_scratch, 1);
            double ss = 0.0;
            bk.sync ();
            bk.d2h (&ss, d_scratch, sizeof (double));
            *res_norm = sqrt (ss);
         }
         const double td[2] = {t_bc, dt};
         bk.h2d (d_time, td, sizeof (td));
         eval_boundary (false);
         enqueue_stage (rk);
         if (deferred_halo ()) bk.p2p_drain (); // stage-by-stage use: the exchange is complete when the call returns
         return bk.check (error);
      }

      // claw.cc:997-1003: limit the initial condition (TVB only)
      int limit_initial_condition ()
      {
         bk.halo_wait ();
         if (tvb ())
         {
            LimiterArgs a = limiter_args (cur);
            a.pos_lim = 0;
            enqueue_indicator (cur);
            launch_limiter (bk, tab.basis, tab.n1, a);
            exchange_halo (cur);
         }
         old = cur;
         return bk.check (error);
      }

      // compute_time_step, claw.cc:444-511 (global time step)
      int compute_dt (double elapsed, double final_time, double *dt)
      {
         bk.halo_wait ();
         const double t0[4] = {elapsed, 0.0, 1.0e20, final_time};
         bk.h2d (d_time, t0, sizeof (t0));
         enqueue_dt ();
         bk.sync ();
         double t[2];
         bk.d2h (t, d_time, sizeof (t));
         *dt = t[1];
         return bk.check (error);
      }

      // n whole steps on the device, claw.cc:1026-1110
      int advance (int n_steps, double final_time, double *elapsed, double *last_dt)
      {
         bk.halo_wait ();
         const double t0[4] = {*elapsed, 0.0, 1.0e20, final_time};
         bk.h2d (d_time, t0, sizeof (t0));
         bk.timer_start ();
         for (int s = 0; s < n_steps; ++s)
         {
            // at the start of a step old == cur, so the buffer rotation of the whole step is
            // a function of cur alone: one CUDA graph per starting buffer
            int cur_after = 0;
            if (bk.graph_launch (cur, &cur_after))
            {
               cur = cur_after;
               old = cur;
            }
            else
            {
               const int key = cur;
               const bool capturing = bk.capture_begin ();
               enqueue_step ();
               if (capturing) bk.capture_end_and_launch (key, cur);
            }
         }
         bk.timer_stop ();
         if (deferred_halo ()) bk.p2p_drain (); // the last exchange of the last step: nothing in flight when the call returns
         bk.sync ();
         double t[2];
         bk.d2h (t, d_time, sizeof (t));
         *elapsed = t[0];
         if (last_dt) *last_dt = t[1];
         int rc = bk.check (error);
         if (rc) return rc;
         return poll_error ();
      }

      // Instrumentation: average device time of the stage kernel of RK stage rk alone, timed
      // with events on the ctx stream, L2 flushed (a scratch buffer larger than L2 is rewritten)
      // before every repetition.  The solution state is not modified.
      int time_stage_kernel (int rk, int reps, size_t flush_bytes, float *avg_ms)
      {
         if (rk < 0 || rk >= n_rk || reps < 1) return fail (DFLO_E_INVALID, "time_stage_kernel: bad arguments");
         bk.halo_wait ();
         void *scratch = flush_bytes ? bk.template alloc<char> (flush_bytes) : nullptr;
         const int out = free_buffer ();
         StageArgs a = stage_args (rk, MODE_STAGE);
         a.out = U[out];
         a.avg_out = AVG[out];
         double total = 0.0;
         for (int r = 0; r < reps; ++r)
         {
            if (scratch) bk.zero (scratch, flush_bytes);
            bk.timer_start ();
            run_stage (a, false);
            bk.timer_stop ();
            total += bk.timer_ms ();
         }
         bk.sync ();
         bk.free (scratch);
         *avg_ms = (float) (total / reps);
         return bk.check (error);
      }

      int poll_error ()
      {
         bk.sync ();
         unsigned int e = 0;
         bk.d2h (&e, d_err, sizeof (e));
         if (e & ERR_NEGATIVE_STATE) return fail (DFLO_E_NEGATIVE_STATE, "Fatal: Negative states");
         if (e & ERR_POSLIM_ROOT) return fail (DFLO_E_POSLIM_ROOT, "Problem in positivity limiter");
         return bk.check (error);
      }

   private:
      int fail (int code, const std::string &msg)
      {
         error = msg;
         return code;
      }

      template <class T>
      static std::vector<T> pad1 (const std::vector<T> &v)
      {
         std::vector<T> r = v;
         if (r.empty ()) r.push_back (T ());
         return r;
      }

      FaceJob *upload_aligned_jobs ()
      {
         static_assert (sizeof (FaceJob) == 4 * sizeof (int), "FaceJob is 4 ints");
         const size_t n = std::max<size_t> (1, lm.jobs.size () / 4);
         FaceJob *d = bk.template alloc<FaceJob> (n);
         if (!lm.jobs.empty ()) bk.h2d (d, lm.jobs.data (), lm.jobs.size () * sizeof (int));
         return d;
      }

      template <class T>
      T *upload (const std::vector<T> &v)
      {
         T *d = bk.template alloc<T> (std::max<size_t> (1, v.size ()));
         if (!v.empty ()) bk.h2d (d, v.data (), v.size () * sizeof (T));
         return d;
      }

      void set_time (int slot, double v) { bk.h2d (d_time + slot, &v, sizeof (double)); }

      int free_buffer () const
      {
         for (int i = 0; i < 3; ++i)
            if (i != cur && i != old) return i;
         return 0;
      }

      // the stage kernel in the form that fits the basis: thread-per-cell for P1 / P2 (cell_stage.cuh; measured on
      // B200, 512 x 512 cells HLLC: P1 62 us against 127 us for the tile kernel, P2 136 / 216, P3 1101 / 487), row kernel
      // for Qk on the device, else the tile kernel.  owned_only: right-hand side of the owned cells only.
      void run_stage (const StageArgs &a, bool owned_only)
      {
         if (general_kernel ())
         {
            MappedStageArgs m;
            m.u = a.u;
            m.u_old = a.u_old;
            m.out = a.out;
            m.avg = a.avg;
            m.avg_out = a.avg_out;
            m.nbr = d_nbr;
            m.nbr_face = d_nbr_face;
            m.fflags = d_fflags;
            m.verts = d_verts;
            m.bc_g = a.bc_g;
            m.bkind = a.bkind;
            m.tab = a.tab;
            m.time = a.time;
            m.ext_force = a.ext_force;
            m.dt_cell = a.dt_cell;
            m.hang_of = d_hang_of;
            m.hang = d_hang;
            m.n_compute = owned_only ? lm.n_owned : lm.n_compute;
            m.n_keep = lm.n_owned;
            m.mode = a.mode;
            m.compat_mpi = a.compat_mpi;
            m.ark = a.ark;
            m.gravity = a.gravity;
            launch_mapped_stage (bk, tab.n1, prm.flux_type, m);
            return;
         }
         if (tab.basis == BASIS_PK && tab.n1 >= 2 && tab.n1 <= 3 && bk.use_pk_cell_kernel () && pk_cell_ok)
         {
            CellStageArgs c;
            c.u = a.u;
            c.u_old = a.u_old;
            c.out = a.out;
            c.avg = a.avg;
            c.avg_out = a.avg_out;
            c.nbr = d_nbr;
            c.fflags = d_fflags;
            c.geom = a.geom;
            c.bc_g = a.bc_g;
            c.bkind = a.bkind;
            c.tab = a.tab;
            c.time = a.time;
            c.dt_cell = a.dt_cell;
            c.ext_force = a.ext_force;
            c.n_compute = owned_only ? lm.n_owned : lm.n_compute;
            c.n_keep = lm.n_owned;
            c.mode = a.mode;
            c.compat_mpi = a.compat_mpi;
            c.ark = a.ark;
            c.gravity = a.gravity;
            c.pf_blocks = bk.stage_prefetch_tiles () > 0 ? bk.n_sms () : 0;
            bk.note_cell_stage ();
            launch_cell_stage (bk, tab.n1, prm.flux_type, c);
            return;
         }
         launch_stage (bk, tab.basis, tab.n1, prm.flux_type, owned_only ? lm.n_tiles_owned : lm.n_tiles, a);
      }

      StageArgs stage_args (int rk, int mode)
      {
         StageArgs a;
         a.u = U[cur];
         a.u_old = U[old];
         a.out = nullptr;
         a.avg = AVG[cur];
         a.avg_out = nullptr;
         a.tiles = d_tiles;
         a.halo_cells = d_halo_cells;
         a.jobs = d_jobs;
         a.geom = d_geom;
         a.bc_g = d_bc_g;
         a.bkind = d_bkind;
         a.tab = d_stage_tab;
         a.time = d_time;
         a.dt_cell = d_dt_cell;
         a.rowdesc = d_rowdesc;
         a.n_cells_u = lm.n_local;
         a.pf_tiles = bk.stage_prefetch_tiles ();
         a.dbg = bk.debug_flags ();
         a.pdl = 0;
         a.n_tiles_owned = lm.n_tiles_owned;
         a.fx = deferred_halo () ? bk.p2p_wait_args () : nullptr;
         a.mode = mode;
         a.compat_mpi = prm.compat == DFLO_COMPAT_MPI;
         a.ark = ark[rk];
         a.gravity = prm.gravity;
         a.ext_force = d_ext_force;
         return a;
      }

      LimiterArgs limiter_args (int buf)
      {
         LimiterArgs a;
         a.u = U[buf];
         a.avg = AVG[buf];
         a.nbr = d_nbr;
         a.fflags = d_fflags;
         a.geom = d_geom;
         a.tab = d_lim_tab;
         a.shock = kxrcf () ? d_shock : nullptr;
         a.flags_out = d_flags;
         a.err = d_err;
         a.n_compute = lm.n_owned;
         a.tvb = tvb () ? (prm.limiter_type == DFLO_LIMITER_MINMAX ? 2 : 1) : 0;
         a.char_lim = prm.char_lim;
         a.pos_lim = pos ();
         a.cam = prm.conserve_angular_momentum;
         a.M = prm.M;
         a.beta = prm.beta;
         return a;
      }

      // compute_shock_indicator (claw.cc:763, 1000) of buffer buf; "limiter" type: all cells, no pass
      void enqueue_indicator (int buf)
      {
         if (!kxrcf ()) return;
         IndicatorArgs a;
         a.u = U[buf];
         a.avg = AVG[buf];
         a.nbr = d_nbr;
         a.fflags = d_fflags;
         a.geom = d_geom;
         a.tab = d_stage_tab;
         a.shock = d_shock;
         a.n_cells = lm.n_owned;
         a.component = prm.shock_indicator == DFLO_INDICATOR_DENSITY ? RHO : ENE;
         launch_indicator (bk, tab.basis, tab.n1, a);
      }

      void compute_cell_average (int buf, int n_cells)
      {
         AvgArgs a;
         a.u = U[buf];
         a.avg = AVG[buf];
         a.gw = d_gw;
         a.gx = d_gx;
         a.verts = mapped () ? d_verts : nullptr;
         a.n_cells = n_cells;
         a.basis = tab.basis;
         a.n1 = tab.n1;
         a.ns = tab.ns;
         bk.template launch1d<CellAverageKernel> (n_cells * 4, a);
      }

      // time_dependent_only: the per-stage refresh inside a step -- programs without t keep their values
      void eval_boundary (bool t_plus_dt, bool time_dependent_only = false)
      {
         if (!have_programs || lm.bf_id.empty ()) return;
         BcEvalArgs a;
         a.bc_g = d_bc_g;
         a.bf_cell = d_bf_cell;
         a.bf_face = d_bf_face;
         a.bf_id = d_bf_id;
         a.geom = d_geom;
         a.gx = d_gx;
         a.code = d_code;
         a.prog_start = time_dependent_only ? d_prog_start_t : d_prog_start;
         a.time = d_time;
         a.n_bfaces = lm.bf_id.size ();
         a.nqf = tab.n1;
         a.use_t_plus_dt = t_plus_dt;
         a.verts = mapped () ? d_verts : nullptr;
         bk.template launch1d<BcEvalKernel> (4 * a.n_bfaces * a.nqf, a);
      }

      void enqueue_dt ()
      {
         if (prm.cfl <= 0.0) // claw.cc:457-461: the time step of the input file, no cell loop
         {
            DtFinalizeArgs f;
            f.time = d_time;
            f.time_step = prm.time_step;
            f.local = 0;
            bk.template launch1d<FixedDtKernel> (1, f);
            return;
         }
         DtArgs a;
         a.avg = AVG[cur];
         a.geom = d_geom;
         a.time = d_time;
         a.n_cells = d_dt_cell ? lm.n_compute : lm.n_owned; // redundantly updated ghost cells need their dt(cell) too
         a.degree = tab.k;
         a.cfl = prm.cfl;
         a.dt_cell = d_dt_cell;
         a.done = reinterpret_cast<unsigned int *> (d_scratch + 2);
         a.nblocks = DtKernel::grid (a.n_cells, mapped ());
         a.finalize = lm.peers.empty ();
         a.time_step = prm.time_step;
         a.u = U[cur];
         a.verts = mapped () ? d_verts : nullptr;
         a.dtq = d_dtq;
         a.n1 = tab.n1;
         bk.template launch<DtKernel> (a.nblocks, a);
         if (a.finalize) return;
         if (bk.allreduce_min_dt (d_time + 2, true, prm.local_time_step ? -2.0 : prm.time_step)) return; // reduced over the ranks and finalised in one launch
         DtFinalizeArgs f;
         f.time = d_time;
         f.time_step = prm.time_step;
         f.local = prm.local_time_step;
         bk.template launch1d<DtFinalizeKernel> (1, f);
      }

      // stage kernel + limiters + halo for rk, reading d_time for dt
      void enqueue_stage (int rk, int pdl = 0)
      {
         const int out = free_buffer ();
         StageArgs a = stage_args (rk, MODE_STAGE);
         a.out = U[out];
         a.avg_out = AVG[out];
         a.pdl = pdl;
         const bool fused = fused_halo ();
         if (fused) a.fx = bk.p2p_fused_args (out);
         run_stage (a, false);
         if (tvb () || pos ())
         {
            LimiterArgs l = limiter_args (out);
            // KXRCF on a sharded context: the indicator of an owned cell reads its neighbours'
            // post-update DoFs; for ghost neighbours they arrive with an extra exchange of the
            // pre-limiter state (the exchange after the limiter then overwrites them)
            if (kxrcf () && !lm.peers.empty ()) exchange_halo (out);
            enqueue_indicator (out);
            launch_limiter (bk, tab.basis, tab.n1, l);
         }
         cur = out;
         if (!fused) exchange_halo (cur, !deferred_halo ());
      }

      void enqueue_step ()
      {
         enqueue_dt ();
         for (int rk = 0; rk < n_rk; ++rk)
         {
            // bc time: t for rk 0, t+dt afterwards (src/claw.cc:736-745); always t in src_mpi
            // refreshed only when the BC time changes: t for rk 0, then t+dt once (src) or never (src_mpi)
            const bool plus_dt = rk > 0 && prm.compat == DFLO_COMPAT_SRC;
            const bool bc_refresh = programs_time_dependent && (rk == 0 || (rk == 1 && plus_dt));
            if (bc_refresh) eval_boundary (plus_dt, true);
            // programmatic dependent launch (row kernel): the first stage starts beside the time-step kernels -- they
            // only write the time scalars, which the stage reads in its last phase; a later stage is scheduled into the
            // tail of the stage before it when nothing sits between the two launches (no limiter, no stand-alone exchange)
            int pdl = 0;
            if (!bc_refresh && row_kernel () && prm.cfl > 0.0)
            {
               const bool back_to_back = !(tvb () || pos ()) && (lm.peers.empty () || fused_halo ());
               if (rk == 0 && bk.pdl_level () >= 1) pdl = 1;
               else if (rk > 0 && back_to_back && bk.pdl_level () >= 2) pdl = 2;
            }
            enqueue_stage (rk, pdl);
         }
         DtFinalizeArgs f;
         f.time = d_time;
         f.time_step = prm.time_step;
         f.local = 0;
         bk.template launch1d<AdvanceTimeKernel> (1, f);
         old = cur;
      }

      void exchange_halo (int buf, bool wait = true)
      {
         if (lm.peers.empty ()) return;
         if (bk.p2p_exchange (buf, wait)) return;
         for (size_t p = 0; p < lm.peers.size (); ++p)
            for (int k = 0; k < 2; ++k)
            {
               const int n = lm.peers[p].send_cells[k].size ();
               if (!n) continue;
               PackArgs a;
               a.src = U[buf];
               a.dst = d_send_u[k][p];
               a.cells = d_send_cells[k][p];
               a.n_cells = n;
               a.width = D ();
               bk.template launch1d<PackKernel> (n * D (), a);
               a.src = AVG[buf];
               a.dst = d_send_avg[k][p];
               a.width = 4;
               bk.template launch1d<PackKernel> (n * 4, a);
            }
         bk.halo_begin ();
         for (size_t p = 0; p < lm.peers.size (); ++p)
            for (int k = 0; k < 2; ++k)
            {
               const HaloPeer &pr = lm.peers[p];
               const int ns = pr.send_cells[k].size ();
               if (ns)
               {
                  bk.halo_send (pr.rank, d_send_u[k][p], (size_t) ns * D ());
                  bk.halo_send (pr.rank, d_send_avg[k][p], (size_t) ns * 4);
               }
               if (pr.recv_count[k])
               {
                  bk.halo_recv (pr.rank, U[buf] + (size_t) pr.recv_start[k] * D (), (size_t) pr.recv_count[k] * D ());
                  bk.halo_recv (pr.rank, AVG[buf] + (size_t) pr.recv_start[k] * 4, (size_t) pr.recv_count[k] * 4);
               }
            }
         bk.halo_end ();
      }

      // caller vector -> device staging
      int stage_external (const double *u, const uint32_t *dof_map, size_t n, bool to_device)
      {
         const size_t need = dof_map ? n : (size_t) lm.n_owned * D ();
         if (ext_capacity < need)
         {
            bk.sync ();
            bk.free (d_ext);
            d_ext = bk.template alloc<double> (need);
            ext_capacity = need;
         }
         if (dof_map)
         {
            // a bad map would read or write device memory out of bounds in LayoutKernel
            for (size_t i = 0; i < n; ++i)
               if ((size_t) dof_map[i] >= n) return fail (DFLO_E_INVALID, "dof_map entry out of range");
            if (dofmap_capacity < n)
            {
               bk.sync ();
               bk.free (d_dofmap);
               d_dofmap = bk.template alloc<uint32_t> (n);
               dofmap_capacity = n;
            }
            bk.h2d (d_dofmap, dof_map, n * sizeof (uint32_t));
         }
         if (to_device)
         {
            if (dof_map)
               bk.h2d (d_ext, u, n * sizeof (double));
            else
               bk.h2d (d_ext, u + (size_t) lm.begin * D (), need * sizeof (double));
         }
         return DFLO_OK;
      }

      void layout (double *local, bool with_map, bool to_local)
      {
         LayoutArgs a;
         a.local = local;
         a.ext = d_ext;
         a.l2g = d_l2g;
         a.dof_map = with_map ? d_dofmap : nullptr;
         a.ext_offset = (int64_t) lm.begin * D ();
         a.n_cells = lm.n_owned;
         a.D = D ();
         a.to_local = to_local;
         bk.template launch1d<LayoutKernel> (lm.n_owned * D (), a);
      }

      int get_vector (const double *local, double *u, const uint32_t *dof_map, size_t n)
      {
         if (n != (size_t) lm.n_global * D ()) return fail (DFLO_E_INVALID, "get: wrong vector length");
         bk.halo_wait ();
         int rc = stage_external (nullptr, dof_map, n, false);
         if (rc) return rc;
         if (dof_map)
         {
            // other ranks' entries must survive: start from the caller's content
            bk.h2d (d_ext, u, n * sizeof (double));
         }
         layout (const_cast<double *> (local), dof_map != nullptr, false);
         bk.sync ();
         if (dof_map)
            bk.d2h (u, d_ext, n * sizeof (double));
         else
            bk.d2h (u + (size_t) lm.begin * D (), d_ext, (size_t) lm.n_owned * D () * sizeof (double));
         return bk.check (error);
      }
   };
}
