// Host-side construction of the FE tables (see tables.h).  Conventions follow what the
// reference relies on in deal.II: Gauss / Gauss-Lobatto rules mapped
// to [0,1], tensor-product points with x fastest, FE_DGP index order "y degree outer, x degree
// inner" (reference src/claw.cc:104-114) with L_i(x) = sqrt(2i+1) P_i(2x-1)
// (src/limiter.cc:395,417 relies on the sqrt(3)).
#include "tables.h"

#include <cmath>
#include <cstring>

namespace dflo
{
   namespace
   {
      // Legendre P_n and P_n' at t in (-1,1) via Bonnet's recursion
      void leg (int n, double t, double &p, double &dp)
      {
         double a = 1.0, b = t;
         if (n == 0)
         {
            p = 1.0;
            dp = 0.0;
            return;
         }
         for (int j = 1; j < n; ++j)
         {
            const double c = ((2 * j + 1) * t * b - j * a) / (j + 1);
            a = b;
            b = c;
         }
         p = b;
         dp = n * (a - t * b) / (1.0 - t * t);
      }

      void gauss01 (int n, double *x, double *w)
      {
         const double pi = 3.14159265358979323846;
         for (int i = 0; i < (n + 1) / 2; ++i)
         {
            double t = std::cos (pi * (i + 0.75) / (n + 0.5)); // descending roots
            double p, dp;
            for (int it = 0; it < 50; ++it)
            {
               leg (n, t, p, dp);
               const double d = p / dp;
               t -= d;
               if (std::fabs (d) < 1e-17) break;
            }
            leg (n, t, p, dp);
            const double wt = 2.0 / ((1.0 - t * t) * dp * dp);
            x[n - 1 - i] = 0.5 * (1.0 + t);
            x[i] = 0.5 * (1.0 - t);
            w[n - 1 - i] = w[i] = 0.5 * wt;
         }
         if (n % 2 == 1) x[n / 2] = 0.5;
      }

      void lobatto01 (int n, double *x)
      {
         const double pi = 3.14159265358979323846;
         const int m = n - 1;
         x[0] = 0.0;
         x[n - 1] = 1.0;
         for (int i = 1; i < n - 1; ++i)
         {
            double t = -std::cos (pi * i / m);
            for (int it = 0; it < 50; ++it)
            {
               double p, dp;
               leg (m, t, p, dp);
               const double d2p = (2.0 * t * dp - m * (m + 1.0) * p) / (1.0 - t * t);
               const double d = dp / d2p;
               t -= d;
               if (std::fabs (d) < 1e-17) break;
            }
            x[i] = 0.5 * (1.0 + t);
         }
         for (int i = 0; i < n / 2; ++i)
         {
            const double s = 0.5 * (x[i] + 1.0 - x[n - 1 - i]);
            x[i] = s;
            x[n - 1 - i] = 1.0 - s;
         }
         if (n % 2 == 1) x[n / 2] = 0.5;
      }

      double lagrange (const double *xs, int n, int a, double x)
      {
         double l = 1.0;
         for (int j = 0; j < n; ++j)
            if (j != a) l *= (x - xs[j]) / (xs[a] - xs[j]);
         return l;
      }

      double lagrange_deriv (const double *xs, int n, int a, double x)
      {
         double s = 0.0;
         for (int m = 0; m < n; ++m)
         {
            if (m == a) continue;
            double t = 1.0 / (xs[a] - xs[m]);
            for (int j = 0; j < n; ++j)
               if (j != a && j != m) t *= (x - xs[j]) / (xs[a] - xs[j]);
            s += t;
         }
         return s;
      }

      // orthonormal Legendre on [0,1] and derivative; end points handled in closed form
      void leg01 (int i, double x, double &L, double &dL)
      {
         const double t = 2.0 * x - 1.0;
         double p, dp;
         if (std::fabs (1.0 - std::fabs (t)) < 1e-14)
         {
            const double s = t > 0 ? 1.0 : -1.0;
            p = (i % 2) ? s : 1.0;
            dp = ((i % 2) ? 1.0 : s) * 0.5 * i * (i + 1.0);
         }
         else
            leg (i, t, p, dp);
         const double nrm = std::sqrt (2.0 * i + 1.0);
         L = nrm * p;
         dL = 2.0 * nrm * dp;
      }
   }

   bool build_tables (int basis, int degree, FeTables &t)
   {
      std::memset (&t, 0, sizeof (t));
      if (degree < 0) return false;
      if (basis == BASIS_QK && degree + 1 > MAX_N1) return false;
      if (basis == BASIS_PK && (degree + 1) * (degree + 2) / 2 > MAX_NS) return false;
      if (basis != BASIS_QK && basis != BASIS_PK) return false;
      t.basis = basis;
      t.k = degree;
      t.n1 = degree + 1;
      const int n1 = t.n1;
      t.nq = n1 * n1;
      gauss01 (n1, t.gx, t.gw);
      for (int ap = 0; ap < n1; ++ap)
         for (int a = 0; a < n1; ++a)
         {
            t.dmat[ap][a] = lagrange_deriv (t.gx, n1, a, t.gx[ap]);
            t.dw[ap][a] = t.dmat[ap][a] * t.gw[ap];
         }
      for (int a = 0; a < n1; ++a)
      {
         t.e[0][a] = lagrange (t.gx, n1, a, 0.0);
         t.e[1][a] = lagrange (t.gx, n1, a, 1.0);
         double s = 0.0;
         for (int ap = 0; ap < n1; ++ap) s += t.dw[ap][a];
         t.gdiff[a] = s;
      }
      // positivity.cc:43: N = (k+3)/2 rounded up
      t.ngll = (degree + 3) % 2 == 0 ? (degree + 3) / 2 : (degree + 4) / 2;
      if (t.ngll > MAX_N1) return false;
      lobatto01 (t.ngll, t.gll);
      for (int j = 0; j < t.ngll; ++j)
         for (int a = 0; a < n1; ++a) t.gl_interp[j][a] = lagrange (t.gx, n1, a, t.gll[j]);
      t.npos = t.ngll * n1;

      if (basis == BASIS_QK)
      {
         t.ns = n1 * n1;
      }
      else
      {
         int m = 0;
         for (int j = 0; j <= degree; ++j)
            for (int i = 0; i <= degree - j; ++i)
            {
               t.px[m] = i;
               t.py[m] = j;
               ++m;
            }
         t.ns = m;
         auto eval = [&] (int mm, double x, double y, double &v, double &dx, double &dy) {
            double lx, dlx, ly, dly;
            leg01 (t.px[mm], x, lx, dlx);
            leg01 (t.py[mm], y, ly, dly);
            v = lx * ly;
            dx = dlx * ly;
            dy = lx * dly;
         };
         for (int b = 0; b < n1; ++b)
            for (int a = 0; a < n1; ++a)
               for (int mm = 0; mm < t.ns; ++mm)
                  eval (mm, t.gx[a], t.gx[b], t.phi[a + n1 * b][mm], t.dphix[a + n1 * b][mm], t.dphiy[a + n1 * b][mm]);
         for (int f = 0; f < 4; ++f)
            for (int q = 0; q < n1; ++q)
            {
               const double x = f == 0 ? 0.0 : f == 1 ? 1.0 : t.gx[q];
               const double y = f == 2 ? 0.0 : f == 3 ? 1.0 : t.gx[q];
               double dx, dy;
               for (int mm = 0; mm < t.ns; ++mm) eval (mm, x, y, t.phiface[f][q][mm], dx, dy);
            }
         // positivity point sets: X = GLL(N) in x (fastest) times Gauss in y; Y = transpose
         for (int b = 0; b < n1; ++b)
            for (int a = 0; a < t.ngll; ++a)
            {
               double dx, dy;
               for (int mm = 0; mm < t.ns; ++mm) eval (mm, t.gll[a], t.gx[b], t.phipos[0][a + t.ngll * b][mm], dx, dy);
            }
         for (int b = 0; b < t.ngll; ++b)
            for (int a = 0; a < n1; ++a)
            {
               double dx, dy;
               for (int mm = 0; mm < t.ns; ++mm) eval (mm, t.gx[a], t.gll[b], t.phipos[1][a + n1 * b][mm], dx, dy);
            }
      }
      t.D = 4 * t.ns;
      return true;
   }

   // value and unit-cell gradient of every scalar basis function at a point of the unit square (output path:
   // what DataOut::build_patches evaluates on its sub-cell vertices)
   void eval_basis (const FeTables &t, double x, double y, double *phi, double *dphix, double *dphiy)
   {
      if (t.basis == BASIS_QK)
      {
         for (int b = 0; b < t.n1; ++b)
            for (int a = 0; a < t.n1; ++a)
            {
               const double la = lagrange (t.gx, t.n1, a, x), lb = lagrange (t.gx, t.n1, b, y);
               phi[a + t.n1 * b] = la * lb;
               dphix[a + t.n1 * b] = lagrange_deriv (t.gx, t.n1, a, x) * lb;
               dphiy[a + t.n1 * b] = la * lagrange_deriv (t.gx, t.n1, b, y);
            }
      }
      else
      {
         for (int m = 0; m < t.ns; ++m)
         {
            double lx, dlx, ly, dly;
            leg01 (t.px[m], x, lx, dlx);
            leg01 (t.py[m], y, ly, dly);
            phi[m] = lx * ly;
            dphix[m] = dlx * ly;
            dphiy[m] = lx * dly;
         }
      }
   }
}
