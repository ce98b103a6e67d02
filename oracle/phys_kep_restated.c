/* TEST INFRASTRUCTURE ONLY (oracle/). Not part of the product; the product never links this.
 *
 * Plain-C restatement of the kinetic-energy-preserving, entropy-stable flux that only the
 * reference's MPI tree has (/root/reference/src_mpi/equation.h: logavg 27-45, kep_diff_matrix
 * 749-837, kep_flux 842-921; dispatched from src_mpi/claw.h:363-370).  Pinned bit for bit against
 * that header's own object code (oracle/_ref/libphys_reference_mpi.so) by tests/golden/flux_kat.npz.
 */
#include <math.h>

#define GAMMA 1.4
#define RHO 2
#define ENE 3

/* src_mpi/equation.h:27-45 */
static double logavg (double a, double b)
{
   double xi = b / a;
   double f = (xi - 1.0) / (xi + 1.0);
   double u = f * f;
   double F;
   if (u < 1.0e-2)
   {
      double u2 = u * u;
      double u3 = u2 * u;
      F = 1.0 + u / 3.0 + u2 / 5.0 + u3 / 7.0;
   }
   else
      F = log (xi) / 2.0 / f;
   return 0.5 * (a + b) / F;
}

/* src_mpi/equation.h:749-837 */
static void kep_diff_matrix (const double normal[2], const double W_l[4], const double W_r[4], double Dm[4][4])
{
   static const double BETA = 1.0 / 6.0;
   double rhol = W_l[RHO];
   double rhor = W_r[RHO];
   double rho = logavg (rhol, rhor);
   double v_l[2], v_r[2], vel[2];
   double v2_l = 0, v2_r = 0;
   double vnl = 0, vnr = 0;
   double vel_normal = 0, v2 = 0;
   for (int d = 0; d < 2; ++d)
   {
      v_l[d] = W_l[d] / W_l[RHO];
      v_r[d] = W_r[d] / W_r[RHO];
      v2_l += v_l[d] * v_l[d];
      v2_r += v_r[d] * v_r[d];
      vnl += v_l[d] * normal[d];
      vnr += v_r[d] * normal[d];
      vel[d] = 0.5 * (v_l[d] + v_r[d]);
      vel_normal += vel[d] * normal[d];
      v2 += vel[d] * vel[d];
   }
   double p_l = (GAMMA - 1) * (W_l[ENE] - 0.5 * W_l[RHO] * v2_l);
   double p_r = (GAMMA - 1) * (W_r[ENE] - 0.5 * W_r[RHO] * v2_r);
   double betal = 0.5 * rhol / p_l;
   double betar = 0.5 * rhor / p_r;
   double beta = logavg (betal, betar);
   double a = sqrt (0.5 * GAMMA / beta);
   double p = 0.5 * (rhol + rhor) / (betal + betar);
   double H = a * a / (GAMMA - 1.0) + 0.5 * v2;
   double v1 = vel[0] * normal[1] - vel[1] * normal[0];
   double R[4][4] = {
      {1, 1, 0, 1},
      {vel[0] - a * normal[0], vel[0], normal[1], vel[0] + a * normal[0]},
      {vel[1] - a * normal[1], vel[1], -normal[0], vel[1] + a * normal[1]},
      {H - a * vel_normal, 0.5 * v2, v1, H + a * vel_normal}};
   double al = sqrt (GAMMA * p_l / rhol);
   double ar = sqrt (GAMMA * p_r / rhor);
   double LambdaL[4] = {vnl - al, vnl, vnl, vnl + al};
   double LambdaR[4] = {vnr - ar, vnr, vnr, vnr + ar};
   double l2, l3;
   l2 = l3 = fabs (vel_normal);
   double Lambda[4] = {fabs (vel_normal - a) + BETA * fabs (LambdaL[0] - LambdaR[0]), l2, l3,
                       fabs (vel_normal + a) + BETA * fabs (LambdaL[3] - LambdaR[3])};
   double S[4] = {0.5 * rho / GAMMA, (GAMMA - 1.0) * rho / GAMMA, p, 0.5 * rho / GAMMA};
   double D[4] = {Lambda[0] * S[0], Lambda[1] * S[1], Lambda[2] * S[2], Lambda[3] * S[3]};
   for (int i = 0; i < 4; ++i)
   {
      for (int j = 0; j < i; ++j) Dm[i][j] = Dm[j][i];
      for (int j = i; j < 4; ++j)
      {
         Dm[i][j] = 0;
         for (int k = 0; k < 4; ++k) Dm[i][j] += R[i][k] * D[k] * R[j][k];
      }
   }
}

/* src_mpi/equation.h:842-921 */
void phys_kep_flux_restated (const double normal[2], const double W_l[4], const double W_r[4], const double Aplus[4],
                             const double Aminus[4], double normal_flux[4])
{
   double rhol = W_l[RHO];
   double rhor = W_r[RHO];
   double rho = logavg (rhol, rhor);
   double v_l[2], v_r[2], vel[2];
   double v2_l = 0, v2_r = 0;
   double vel_normal = 0;
   for (int d = 0; d < 2; ++d)
   {
      v_l[d] = W_l[d] / W_l[RHO];
      v_r[d] = W_r[d] / W_r[RHO];
      v2_l += v_l[d] * v_l[d];
      v2_r += v_r[d] * v_r[d];
      vel[d] = 0.5 * (v_l[d] + v_r[d]);
      vel_normal += vel[d] * normal[d];
   }
   double vel2 = 0.5 * (v2_l + v2_r);
   double p_l = (GAMMA - 1) * (W_l[ENE] - 0.5 * W_l[RHO] * v2_l);
   double p_r = (GAMMA - 1) * (W_r[ENE] - 0.5 * W_r[RHO] * v2_r);
   double betal = 0.5 * rhol / p_l;
   double betar = 0.5 * rhor / p_r;
   double beta = logavg (betal, betar);
   double p = 0.5 * (rhol + rhor) / (betal + betar);
   /* central flux */
   normal_flux[RHO] = rho * vel_normal;
   for (int d = 0; d < 2; ++d) normal_flux[d] = normal[d] * p + vel[d] * normal_flux[RHO];
   normal_flux[ENE] = 0.5 * (1.0 / ((GAMMA - 1.0) * beta) - vel2) * normal_flux[RHO] + normal_flux[0] * vel[0] + normal_flux[1] * vel[1];
   double Dm[4][4];
   kep_diff_matrix (normal, Aplus, Aminus, Dm);
   /* jump in entropy: s = log(p) - gamma*log(rho) */
   double ds = log (p_r / p_l) - GAMMA * log (rhor / rhol);
   double dV[4] = {-ds / (GAMMA - 1.0) - (betar * v2_r - betal * v2_l), 2.0 * (betar * v_r[0] - betal * v_l[0]),
                   2.0 * (betar * v_r[1] - betal * v_l[1]), -2.0 * (betar - betal)};
   double Diff[4] = {0.0, 0.0, 0.0, 0.0};
   for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) Diff[i] += Dm[i][j] * dV[j];
   normal_flux[RHO] -= 0.5 * Diff[0];
   normal_flux[0] -= 0.5 * Diff[1];
   normal_flux[1] -= 0.5 * Diff[2];
   normal_flux[ENE] -= 0.5 * Diff[3];
}

/* src_mpi/equation.h:299-335: left / right eigenvector matrices along the streamline direction
 * (kx,ky) = (cos, sin)(atan2(v,u)), used by the minmax limiter (src_mpi/limiter.cc:450).  Row-major 4x4,
 * variable order of the transforms (rho, m_x, m_y, E) as in the x / y matrices. */
void phys_eigen_stream_restated (const double W[4], double R[16], double L[16])
{
   double g1 = GAMMA - 1.0;
   double rho = W[RHO];
   double E = W[ENE];
   double u = W[0] / rho;
   double v = W[1] / rho;
   double q2 = u * u + v * v;
   double p = g1 * (E - 0.5 * rho * q2);
   double c2 = GAMMA * p / rho;
   double c = sqrt (c2);
   double beta = 0.5 / c2;
   double phi2 = 0.5 * g1 * q2;
   double h = c2 / g1 + 0.5 * q2;
   double theta = atan2 (v, u);
   double kx = cos (theta);
   double ky = sin (theta);
   double uk = u * kx + v * ky;

   R[0] = 1;         R[1] = 0;                 R[2] = 1;           R[3] = 1;
   R[4] = u;         R[5] = ky;                R[6] = u + kx * c;  R[7] = u - kx * c;
   R[8] = v;         R[9] = -kx;               R[10] = v + ky * c; R[11] = v - ky * c;
   R[12] = 0.5 * q2; R[13] = ky * u - kx * v;  R[14] = h + c * uk; R[15] = h - c * uk;

   L[0] = 1 - phi2 / c2;          L[1] = g1 * u / c2;              L[2] = g1 * v / c2;               L[3] = -g1 / c2;
   L[4] = -(ky * u - kx * v);     L[5] = ky;                       L[6] = -kx;                       L[7] = 0;
   L[8] = beta * (phi2 - c * uk); L[9] = beta * (kx * c - g1 * u); L[10] = beta * (ky * c - g1 * v); L[11] = beta * g1;
   L[12] = beta * (phi2 + c * uk); L[13] = -beta * (kx * c + g1 * u); L[14] = -beta * (ky * c + g1 * v); L[15] = beta * g1;
}

/* src_mpi/equation.h:1189-1202 */
void phys_ext_forcing_restated (const double W[4], const double f[2], double G[4])
{
   int d;
   G[RHO] = 0.0;
   G[ENE] = 0.0;
   for (d = 0; d < 2; ++d)
   {
      G[d] = W[RHO] * f[d];
      G[ENE] += W[d] * f[d];
   }
}
