// Finite-element tables of the hot path on the unit square, built once on the host and copied
// to the device.  They hold what dflo obtains per cell and per stage from deal.II FEValues
// (reference src/claw.cc:416-437, src/assemble_explicit.cc:39-41): on Cartesian cells the
// mapped quantities are the unit-cell tables times hx, hy.
#pragma once

namespace dflo
{
   constexpr int MAX_N1 = 5;            // Qk up to degree 4
   constexpr int MAX_NS = 10;           // Pk up to degree 3
   constexpr int MAX_NQ = MAX_N1 * MAX_N1;
   constexpr int MAX_NPOS = MAX_N1 * MAX_N1;

   enum Basis { BASIS_QK = 0, BASIS_PK = 1 };

   struct FeTables
   {
      int basis, k, n1, ns, nq, D, ngll, npos;

      // 1-D Gauss rule on [0,1] with k+1 points (QGauss<1>(k+1))
      double gx[MAX_N1], gw[MAX_N1];
      // Qk (Lagrange on the Gauss nodes, FE_DGQArbitraryNodes): derivative matrix
      // dmat[ap][a] = l_a'(gx[ap]), its quadrature-weighted form dw[ap][a] = dmat[ap][a]*gw[ap],
      // end-point values e[0][a] = l_a(0), e[1][a] = l_a(1), and gdiff[a] = sum_ap dw[ap][a]
      double dmat[MAX_N1][MAX_N1], dw[MAX_N1][MAX_N1];
      double e[2][MAX_N1];
      double gdiff[MAX_N1];
      // Gauss-Lobatto nodes (QGaussLobatto<1>(N), positivity.cc:43-45) and, for Qk, the 1-D
      // interpolation l_a(gll[j])
      double gll[MAX_N1];
      double gl_interp[MAX_N1][MAX_N1];
      // Pk (orthonormal Legendre, FE_DGP): dense tables at the (k+1)^2 Gauss points (x fastest),
      // at the k+1 points of each face, and at the two positivity point sets
      int px[MAX_NS], py[MAX_NS];
      double phi[MAX_NQ][MAX_NS], dphix[MAX_NQ][MAX_NS], dphiy[MAX_NQ][MAX_NS];
      double phiface[4][MAX_N1][MAX_NS];
      double phipos[2][MAX_NPOS][MAX_NS];
   };

   // returns false if (basis, degree) is outside the supported range
   bool build_tables (int basis, int degree, FeTables &t);
   // phi[ns], dphix[ns], dphiy[ns]: every scalar basis function and its unit-cell gradient at (x,y) in [0,1]^2
   void eval_basis (const FeTables &t, double x, double y, double *phi, double *dphix, double *dphiy);
}
