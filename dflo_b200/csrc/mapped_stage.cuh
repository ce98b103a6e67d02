// Stage kernel for `mapping = q1` (MappingQ1: straight-sided quadrilaterals), Qk basis -- SURVEY.md 8(f) row 2.
//
// assemble_system (reference src/assemble_explicit.cc:30-452) with the geometry FEValues / FEFaceValues deliver under
// MappingQ1 (src/claw.cc:165-190), M^-1 (src/claw.cc:228-258: diagonal, 1 / (w_a w_b det J) at the Gauss nodes -- "not exact
// for general cells", as the reference notes), the RK combine (694-713, 757-760) and compute_cell_average (562-597) in
// one launch.  On the collocated Gauss-node basis the mapped scheme keeps the sum-factorised form of the Cartesian one:
//
//   volume   r_(a,b) += w_b sum_a' F1~(a',b) D[a'][a] w_a'  +  w_a sum_b' F2~(a,b') D[b'][b] w_b'
//            with the contravariant fluxes  F1~ = y_eta F_x - x_eta F_y,  F2~ = -y_xi F_x + x_xi F_y  (= det J . J^-1 F)
//   faces    r_(a,b) -= H_0(b) l_a(0) + H_1(b) l_a(1) + H_2(a) l_b(0) + H_3(a) l_b(1),   H_f(q) = w_q |edge_f| H(n_f, W+, W-)
//
// A face is straight, so its outward normal and its JxW = w_q |edge| are constants of the face; the neighbour may see
// the face as any of its own four faces (neighbor_face) and run along it in the opposite direction (DFLO_FACE_FLIP).
// Which cell is the reference's "plus" side (MeshWorker owner, assemble_explicit.cc:440; both for periodic pairs) is
// honoured: the non-owner evaluates the owner's Riemann problem (owner's normal = minus its own, bit for bit, because
// both subtract the same two vertices) and negates, so the two cells subtract the same bits -- conservative to
// round-off without atomics, and sharded = single GPU bit for bit.
//
// One thread per (cell, Gauss node); CPB cells per block; neighbour traces straight from global memory (L2).  This
// is a phase kernel: the same code runs thread by thread on the CPU emulation backend of tests/emu.  It is the
// general-geometry path, not the fast one: the register-blocked row kernel serves mapping = cartesian.
#pragma once

#include "kernels.cuh"

namespace dflo
{
   struct MappedStageArgs
   {
      const double *u, *u_old;
      double *out;
      const double *avg;
      double *avg_out;
      const int *nbr;               // [n_local][4]
      const unsigned char *nbr_face; // [n_local][4]
      const unsigned char *fflags;  // [n_local][4]
      const double *verts;          // [n_local][8]
      const double *bc_g;           // [n_bfaces][N1][4]
      const int *bkind;
      const double *tab;            // flat Qk stage tables: dw[N1][N1] e0[N1] e1[N1] gw[N1]
      const double *time;
      const double *ext_force;      // [n_local][NQ][2] or nullptr
      const double *dt_cell;        // time step type = local: dt(cell), else nullptr
      const int *hang_of, *hang;    // faces with hanging nodes: [n_local][4] index or -1; per face {fine cell, its face, backwards?} x 2; or nullptr
      int n_compute, n_keep;
      int mode, compat_mpi;
      double ark, gravity;
   };

   // geometry of a bilinear cell: Jacobian entries at (xi, eta)
   DFLO_DEV void q1_jacobian (const double *v, double xi, double eta, double &xxi, double &xeta, double &yxi, double &yeta)
   {
      xxi = (v[2] - v[0]) * (1.0 - eta) + (v[6] - v[4]) * eta;
      xeta = (v[4] - v[0]) * (1.0 - xi) + (v[6] - v[2]) * xi;
      yxi = (v[3] - v[1]) * (1.0 - eta) + (v[7] - v[5]) * eta;
      yeta = (v[5] - v[1]) * (1.0 - xi) + (v[7] - v[3]) * xi;
   }
   // outward unit normal and length of face f: the straight edge between the face's two vertices (deal.II order:
   // face 0 = v0 v2, 1 = v1 v3, 2 = v0 v1, 3 = v2 v3); (t_y, -t_x) points out of faces 1 and 2
   DFLO_DEV void q1_face (const double *v, int f, double &nx, double &ny, double &len)
   {
      const int a = f == 0 ? 0 : f == 1 ? 1 : f == 2 ? 0 : 2, b = f == 0 ? 2 : f == 1 ? 3 : f == 2 ? 1 : 3;
      const double tx = v[2 * b] - v[2 * a], ty = v[2 * b + 1] - v[2 * a + 1];
      len = sqrt (tx * tx + ty * ty);
      const double s = (f == 1 || f == 2) ? 1.0 : -1.0;
      nx = s * ty / len;
      ny = -s * tx / len;
   }
   DFLO_DEV double q1_measure (const double *v)
   {
      return 0.5 * ((v[6] - v[0]) * (v[5] - v[3]) - (v[4] - v[2]) * (v[7] - v[1]));
   }
   DFLO_DEV double q1_diameter (const double *v)
   {
      const double d1 = sqrt ((v[6] - v[0]) * (v[6] - v[0]) + (v[7] - v[1]) * (v[7] - v[1]));
      const double d2 = sqrt ((v[4] - v[2]) * (v[4] - v[2]) + (v[5] - v[3]) * (v[5] - v[3]));
      return d1 > d2 ? d1 : d2;
   }

   template <int N1, int FLUX>
   struct MappedStageKernel
   {
      typedef MappedStageArgs Args;
      static constexpr int NQ = N1 * N1, D = 4 * NQ;
#ifndef DFLO_MAPPED_BLOCK
#define DFLO_MAPPED_BLOCK 128 // threads per block; one-warp blocks (32) measured 0.3675 vs 0.374 ms per step on the q1 bench entry: within noise, not taken
#endif
      static constexpr int CPB = NQ >= DFLO_MAPPED_BLOCK ? 1 : DFLO_MAPPED_BLOCK / NQ; // cells per block
      static constexpr int THREADS = (CPB * NQ + 31) / 32 * 32;
#ifndef DFLO_MAPPED_MIN_BLOCKS
#define DFLO_MAPPED_MIN_BLOCKS 6
#endif
      // 85 registers, 24 warps per SM.  Left alone the allocation had grown to 196 registers with the sub-face branches of the
      // hanging-node faces (2 blocks, 8 warps per SM), although most launches never take them.  Measured on the 256 x 256 Q3
      // skewed mesh, ms per step: 2 blocks 0.715, 4 blocks 0.427, 6 blocks 0.374 (a few bytes of spills), 8 blocks 0.398
      static constexpr int MIN_BLOCKS = DFLO_MAPPED_MIN_BLOCKS * 128 / DFLO_MAPPED_BLOCK;
      static constexpr int NPHASE = 4;
      static constexpr int NTAB = N1 * N1 + 3 * N1;
      static constexpr int O_TAB = 0;
      static constexpr int O_U = (NTAB + 1) / 2 * 2;
      static constexpr int O_F = O_U + CPB * D;                      // [CPB][NQ][8] contravariant fluxes
      static constexpr int O_H = O_F + CPB * NQ * 8;                 // [CPB][4 faces][2 halves][N1][4] weighted face fluxes (second half: hanging nodes)
      static constexpr int O_P = O_H + CPB * 32 * N1;                // [CPB][NQ][4] the nodes' parts of the cell means
      static constexpr int SMEM_DOUBLES = O_P + CPB * NQ * 4;
      static int grid (int n_cells) { return (n_cells + CPB - 1) / CPB; }

      // trace of a cell (D DoFs at `uc`) at point p of its face f -- one fma chain for own and neighbour traces
      static DFLO_DEV void trace (const double *tb, const double *uc, int f, int p, double W[4])
      {
         const double *e = tb + N1 * N1 + (f & 1) * N1;
         const int base = (f < 2) ? N1 * p : p, stride = (f < 2) ? 1 : N1;
#pragma unroll
         for (int c = 0; c < 4; ++c)
         {
            double s = 0.0;
#pragma unroll
            for (int a = 0; a < N1; ++a) s = fma (e[a], uc[c * NQ + base + a * stride], s);
            W[c] = s;
         }
      }

      // trace of a COARSE cell at point qc of child `child` of its face f (FESubfaceValues): the trace along the normal of every
      // tangential line, interpolated to the sub-face point with S[child][qc][.] -- one chain for both sides of the face
      static DFLO_DEV void subtrace (const double *tb, const double *S, const double *uc, int f, int child, int qc, double W[4])
      {
         const double *e = tb + N1 * N1 + (f & 1) * N1;
         const double *Sq = S + (child * N1 + qc) * N1;
         const int stride = (f < 2) ? 1 : N1;
#pragma unroll
         for (int c = 0; c < 4; ++c)
         {
            double s = 0.0;
#pragma unroll
            for (int t = 0; t < N1; ++t)
            {
               const int base = (f < 2) ? N1 * t : t;
               double in = 0.0;
#pragma unroll
               for (int a = 0; a < N1; ++a) in = fma (e[a], uc[c * NQ + base + a * stride], in);
               s = fma (Sq[t], in, s);
            }
            W[c] = s;
         }
      }

      static DFLO_DEV void phase (int ph, const Args &A, double *sm, int tid, int bid)
      {
         double *tb = sm + O_TAB, *su = sm + O_U, *sF = sm + O_F, *sH = sm + O_H;
         const int slot = tid / NQ, q = tid % NQ;
         const int cell = bid * CPB + slot;
         const bool active = slot < CPB && cell < A.n_compute;
         const int a = q % N1, b = q / N1;
         const double *gw = tb + N1 * N1 + 2 * N1;
         const double *gx = A.tab + NTAB;      // Gauss nodes follow the stage tables (tables_pack.h) ...
         const double *S = A.tab + NTAB + N1;  // ... and the sub-face interpolation table S[child][q][a]

         if (ph == 0)
         {
            for (int i = tid; i < NTAB; i += THREADS) tb[i] = A.tab[i];
            if (active)
            {
               const double *uc = A.u + (size_t) cell * D;
#pragma unroll
               for (int c = 0; c < 4; ++c) su[slot * D + c * NQ + q] = uc[c * NQ + q];
            }
            return;
         }
         if (!active) return;
         const double *v = A.verts + (size_t) cell * 8;
         double *uc = su + slot * D;

         if (ph == 1)
         {
            // volume fluxes at the own node
            {
               double xxi, xeta, yxi, yeta;
               q1_jacobian (v, gx[a], gx[b], xxi, xeta, yxi, yeta);
               const double W[4] = {uc[q], uc[NQ + q], uc[2 * NQ + q], uc[3 * NQ + q]};
               const double Wy[4] = {W[1], W[0], W[2], W[3]};
               double Fx[4], G[4];
               flux_x (W, Fx);
               flux_x (Wy, G); // F_y(W) = F_x with the momentum components exchanged
               const double Fy[4] = {G[1], G[0], G[2], G[3]};
               double *f = sF + (slot * NQ + q) * 8;
#pragma unroll
               for (int c = 0; c < 4; ++c)
               {
                  f[c] = yeta * Fx[c] - xeta * Fy[c];
                  f[4 + c] = -yxi * Fx[c] + xxi * Fy[c];
               }
            }
            // face fluxes: (face, half, point) items of the cell spread over its NQ threads; the second half of a face only
            // exists where the face has a hanging node
            for (int j = q; j < 8 * N1; j += NQ)
            {
               const int f = j / (2 * N1), half = (j / N1) & 1, p = j % N1;
               const int fl = A.fflags[(size_t) cell * 4 + f];
               if (half && !(fl & DFLO_FACE_HANGING)) continue;
               // The branches only gather the two states, the normal and the weight; ONE Riemann solve follows (five inlined
               // copies of it cost the kernel 84 registers).  swap: the other cell is the one that integrates the face -- its
               // problem, posed with its normal, is evaluated with the same bits and this cell takes minus the flux.
               double Wp[4], Wm[4], Ap[4], Am[4], H[4];
               double nx, ny, len;
               bool swap = false;
               int pw = p; // Gauss weight of the point
               if (flux_uses_averages (FLUX))
               {
#pragma unroll
                  for (int c = 0; c < 4; ++c) Ap[c] = Am[c] = A.avg[(size_t) cell * 4 + c];
               }
               if (fl & DFLO_FACE_HANGING)
               {
                  // coarse side of a face with a hanging node: the fine cell behind this half is the one that integrates
                  // (MeshWorker); its Riemann problem -- its normal, its face points, its JxW
                  const int *hg = A.hang + 6 * (size_t) A.hang_of[(size_t) cell * 4 + f] + 3 * half;
                  const int fine = hg[0], nf = hg[1];
                  const int pf = hg[2] ? N1 - 1 - p : p; // the fine side's number of this point
                  q1_face (A.verts + (size_t) fine * 8, nf, nx, ny, len);
                  trace (tb, A.u + (size_t) fine * D, nf, pf, Wm);
                  subtrace (tb, S, uc, f, half, p, Wp);
                  if (flux_uses_averages (FLUX))
                  {
#pragma unroll
                     for (int c = 0; c < 4; ++c) Am[c] = A.avg[(size_t) fine * 4 + c];
                  }
                  swap = true;
                  pw = pf;
               }
               else
               {
                  q1_face (v, f, nx, ny, len);
                  trace (tb, uc, f, p, Wp);
                  const int nb = A.nbr[(size_t) cell * 4 + f];
                  if (nb >= 0)
                  {
                     const int nf = A.nbr_face[(size_t) cell * 4 + f];
                     const int pn = (fl & DFLO_FACE_FLIP) ? N1 - 1 - p : p;
                     if (fl & DFLO_FACE_COARSER) // fine side: the neighbour's trace on this half of its face, at this cell's points
                        subtrace (tb, S, A.u + (size_t) nb * D, nf, (fl & DFLO_FACE_CHILD1) ? 1 : 0, pn, Wm);
                     else
                     {
                        trace (tb, A.u + (size_t) nb * D, nf, pn, Wm);
                        if (!(fl & (DFLO_FACE_OWNER | DFLO_FACE_PERIODIC)))
                        {
                           // the owner's Riemann problem, its normal = -(nx, ny)
                           swap = true;
                           nx = -nx;
                           ny = -ny;
                        }
                     }
                     if (flux_uses_averages (FLUX))
                     {
#pragma unroll
                        for (int c = 0; c < 4; ++c) Am[c] = A.avg[(size_t) nb * 4 + c];
                     }
                  }
                  else
                  {
                     // boundary: W- from the boundary condition (assemble_explicit.cc:176-206)
                     const int bf = -1 - nb;
                     const int kind = A.bkind[bf];
                     double g[4];
#pragma unroll
                     for (int c = 0; c < 4; ++c) g[c] = A.bc_g[((size_t) bf * N1 + p) * 4 + c];
                     compute_wminus (kind, nx, ny, Wp, g, Wm);
                     if (flux_uses_averages (FLUX) && A.compat_mpi) compute_wminus (kind, nx, ny, Ap, g, Am); // else: own average on both sides
                  }
               }
               {
                  double P[4], M[4], AP[4], AM[4];
#pragma unroll
                  for (int c = 0; c < 4; ++c)
                  {
                     P[c] = swap ? Wm[c] : Wp[c];
                     M[c] = swap ? Wp[c] : Wm[c];
                     AP[c] = swap ? Am[c] : Ap[c];
                     AM[c] = swap ? Ap[c] : Am[c];
                  }
                  numerical_flux<FLUX> (nx, ny, P, M, AP, AM, H);
               }
               const double w = (swap ? -1.0 : 1.0) * (gw[pw] * len); // fe_v.JxW(q) on a straight face
               double *h = sH + (((slot * 4 + f) * 2 + half) * N1 + p) * 4;
#pragma unroll
               for (int c = 0; c < 4; ++c) h[c] = w * H[c];
            }
            return;
         }

         if (ph == 2)
         {
            // residual at the own node, M^-1, RK combine, write-back, mean parts
            const double *dw = tb, *e0 = tb + N1 * N1, *e1 = e0 + N1;
            double xxi, xeta, yxi, yeta;
            q1_jacobian (v, gx[a], gx[b], xxi, xeta, yxi, yeta);
            const double det = xxi * yeta - xeta * yxi;
            const double wq = gw[a] * gw[b] * det; // JxW at the node
            const double *F = sF + slot * NQ * 8;
            // lifting: the flux of face f at this node's tangential position t -- on a face with a hanging node the two halves'
            // fluxes tested with the coarse basis at the sub-face points, sum_half sum_q H(half, q) S[half][q][t]
            double L[4][4]; // [face][component]
#pragma unroll
            for (int f = 0; f < 4; ++f)
            {
               const int t = f < 2 ? b : a;
               const double *Hf = sH + ((slot * 4 + f) * 2) * N1 * 4;
               if (A.hang_of && (A.fflags[(size_t) cell * 4 + f] & DFLO_FACE_HANGING))
               {
#pragma unroll
                  for (int c = 0; c < 4; ++c) L[f][c] = 0.0;
                  for (int half = 0; half < 2; ++half)
                     for (int qq = 0; qq < N1; ++qq)
                     {
                        const double sv = S[(half * N1 + qq) * N1 + t];
#pragma unroll
                        for (int c = 0; c < 4; ++c) L[f][c] = fma (sv, Hf[(half * N1 + qq) * 4 + c], L[f][c]);
                     }
               }
               else
               {
#pragma unroll
                  for (int c = 0; c < 4; ++c) L[f][c] = Hf[t * 4 + c];
               }
            }
            double r[4];
#pragma unroll
            for (int c = 0; c < 4; ++c)
            {
               double sx = 0.0, sy = 0.0;
#pragma unroll
               for (int ap = 0; ap < N1; ++ap) sx = fma (F[(ap + N1 * b) * 8 + c], dw[ap * N1 + a], sx);
#pragma unroll
               for (int bp = 0; bp < N1; ++bp) sy = fma (F[(a + N1 * bp) * 8 + 4 + c], dw[bp * N1 + b], sy);
               r[c] = gw[b] * sx + gw[a] * sy - (L[0][c] * e0[a] + L[1][c] * e1[a] + L[2][c] * e0[b] + L[3][c] * e1[b]);
            }
            const double W[4] = {uc[q], uc[NQ + q], uc[2 * NQ + q], uc[3 * NQ + q]};
            if (A.gravity != 0.0) // assemble_explicit.cc:78, 108-111
            {
               double Gv[4];
               if (A.ext_force)
                  forcing_ext (W, A.ext_force[((size_t) cell * NQ + q) * 2], A.ext_force[((size_t) cell * NQ + q) * 2 + 1], Gv);
               else
                  forcing (W, Gv);
#pragma unroll
               for (int c = 0; c < 4; ++c) r[c] += A.gravity * Gv[c] * wq;
            }
            double vnew[4];
            if (A.mode == MODE_RHS)
            {
#pragma unroll
               for (int c = 0; c < 4; ++c) vnew[c] = r[c];
            }
            else
            {
               const double dt = A.dt_cell ? A.dt_cell[cell] : A.time[1];
               const double invm = 1.0 / wq; // claw.cc:228-258 with a collocated basis: M_ii = JxW_i
               const bool need_old = A.ark != 0.0;
#pragma unroll
               for (int c = 0; c < 4; ++c)
               {
                  const double un = W[c] + dt * r[c] * invm;
                  vnew[c] = need_old ? (1.0 - A.ark) * un + A.ark * A.u_old[(size_t) cell * D + c * NQ + q] : un;
               }
            }
            if (cell < A.n_keep)
            {
#pragma unroll
               for (int c = 0; c < 4; ++c) A.out[(size_t) cell * D + c * NQ + q] = vnew[c];
            }
            // compute_cell_average (claw.cc:562-597): sum_q u_q JxW_q / measure -- the node's part
            double *part = sm + O_P + (slot * NQ + q) * 4;
#pragma unroll
            for (int c = 0; c < 4; ++c) part[c] = vnew[c] * wq;
            return;
         }

         // ph == 3: cell averages of the updated solution
         if (A.mode == MODE_STAGE)
            for (int c = q; c < 4; c += NQ)
            {
               double s = 0.0;
               for (int k = 0; k < NQ; ++k) s += sm[O_P + (slot * NQ + k) * 4 + c];
               A.avg_out[(size_t) cell * 4 + c] = s / q1_measure (v);
            }
      }
   };
}
