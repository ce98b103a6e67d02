/* TEST INFRASTRUCTURE ONLY (oracle/). Not part of the product; the product never links this.
 *
 * Plain-C restatement of the point-wise Euler arithmetic of dflo's EulerEquations<2>
 * (/root/reference/src/equation.h), keeping the reference's order of floating-point
 * operations so that, compiled without FMA contraction, it reproduces the reference's object
 * code bit for bit (tests/test_oracle_physics.py checks that against oracle/_ref and against
 * tests/golden/physics_kat.npz).  Each function cites the lines it follows.
 * Components: 0 = rho*u, 1 = rho*v, 2 = rho (density_component), 3 = E (energy_component).
 */
#include "phys.h"
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <math.h>
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#define GAMMA 1.4 /* equation.cc:33 */
#define RHO 2
#define ENE 3

const char *phys_impl_name (void) { return "restated:phys_restated.c"; }

/* equation.h:67-79 */
static double kinetic_energy (const double W[4])
{
   double ke = 0;
   ke += W[0] * W[0];
   ke += W[1] * W[1];
   ke *= 0.5 / W[RHO];
   return ke;
}

/* equation.h:84-92 */
double phys_pressure (const double W[4]) { return (GAMMA - 1.0) * (W[ENE] - kinetic_energy (W)); }

/* equation.h:142-152 */
double phys_sound_speed (const double W[4]) { return sqrt (GAMMA * phys_pressure (W) / W[RHO]); }

/* equation.h:97-114 (|v| + c) */
double phys_max_eigenvalue (const double W[4])
{
   const double p = phys_pressure (W);
   double vel = 0;
   vel += W[0] * W[0];
   vel += W[1] * W[1];
   vel = sqrt (vel) / W[RHO];
   return vel + sqrt (GAMMA * p / W[RHO]);
}

/* equation.h:119-137 (|v.n| + c), used by the LxF dissipation */
static double max_eigenvalue_normal (const double W[4], const double n[2])
{
   const double p = phys_pressure (W);
   const double sonic = sqrt (GAMMA * p / W[RHO]);
   double vel = 0;
   vel += W[0] * n[0];
   vel += W[1] * n[1];
   vel /= W[RHO];
   return fabs (vel) + sonic;
}

/* equation.h:158-193 */
void phys_flux_matrix (const double W[4], double F[8])
{
   const double p = phys_pressure (W);
   for (int d = 0; d < 2; ++d)
   {
      for (int e = 0; e < 2; ++e) F[2 * d + e] = W[d] * W[e] / W[RHO];
      F[2 * d + d] += p;
   }
   for (int d = 0; d < 2; ++d) F[2 * RHO + d] = W[d];
   for (int d = 0; d < 2; ++d) F[2 * ENE + d] = W[d] / W[RHO] * (W[ENE] + p);
}

/* equation.h:829-850: gravity -1 along y */
void phys_forcing (const double W[4], double G[4])
{
   const double g = -1.0;
   G[0] = 0;
   G[1] = g * W[RHO];
   G[RHO] = 0;
   G[ENE] = g * W[1];
}

/* equation.h:324-377 */
static void lxf (const double n[2], const double Wp[4], const double Wm[4], const double Ap[4],
                 const double Am[4], double H[4])
{
   double vnp = 0, vnm = 0;
   for (int d = 0; d < 2; ++d)
   {
      vnp += Wp[d] * n[d];
      vnm += Wm[d] * n[d];
   }
   vnp /= Wp[RHO];
   vnm /= Wm[RHO];
   const double pp = phys_pressure (Wp), pm = phys_pressure (Wm);
   const double lp = max_eigenvalue_normal (Ap, n), lm = max_eigenvalue_normal (Am, n);
   const double lambda = lp < lm ? lm : lp; /* std::max(lp, lm) */
   for (int d = 0; d < 2; ++d)
      H[d] = 0.5 * (pp * n[d] + Wp[d] * vnp + pm * n[d] + Wm[d] * vnm);
   H[RHO] = 0.5 * (Wp[RHO] * vnp + Wm[RHO] * vnm);
   H[ENE] = 0.5 * ((Wp[ENE] + pp) * vnp + (Wm[ENE] + pm) * vnm);
   for (int c = 0; c < 4; ++c) H[c] += 0.5 * lambda * (Wp[c] - Wm[c]);
}

static double dmax (double a, double b) { return a < b ? b : a; } /* std::max(a,b) */
static double dmin (double a, double b) { return b < a ? b : a; } /* std::min(a,b) */

/* equation.h:382-464 */
static void steger_warming (const double n[2], const double Wp[4], const double Wm[4], double H[4])
{
   double pf[4], mf[4];
   double vnp = 0, vnm = 0, q2p = 0, q2m = 0;
   for (int d = 0; d < 2; ++d)
   {
      vnp += Wp[d] * n[d];
      vnm += Wm[d] * n[d];
      q2p += Wp[d] * Wp[d];
      q2m += Wm[d] * Wm[d];
   }
   vnp /= Wp[RHO];
   vnm /= Wm[RHO];
   q2p /= Wp[RHO] * Wp[RHO];
   q2m /= Wm[RHO] * Wm[RHO];
   const double pp = phys_pressure (Wp), pm = phys_pressure (Wm);
   const double cp = sqrt (GAMMA * pp / Wp[RHO]);
   const double cm = sqrt (GAMMA * pm / Wm[RHO]);

   const double l1p = dmax (vnp, 0.0), l2p = dmax (vnp + cp, 0.0), l3p = dmax (vnp - cp, 0.0);
   const double ap = 2.0 * (GAMMA - 1.0) * l1p + l2p + l3p;
   const double fp = 0.5 * Wp[RHO] / GAMMA;
   for (int d = 0; d < 2; ++d) pf[d] = ap * Wp[d] / Wp[RHO] + cp * (l2p - l3p) * n[d];
   pf[RHO] = ap;
   pf[ENE] = 0.5 * ap * q2p + cp * vnp * (l2p - l3p) + cp * cp * (l2p + l3p) / (GAMMA - 1.0);

   const double l1m = dmin (vnm, 0.0), l2m = dmin (vnm + cm, 0.0), l3m = dmin (vnm - cm, 0.0);
   const double am = 2.0 * (GAMMA - 1.0) * l1m + l2m + l3m;
   const double fm = 0.5 * Wm[RHO] / GAMMA;
   for (int d = 0; d < 2; ++d) mf[d] = am * Wm[d] / Wm[RHO] + cm * (l2m - l3m) * n[d];
   mf[RHO] = am;
   mf[ENE] = 0.5 * am * q2m + cm * vnm * (l2m - l3m) + cm * cm * (l2m + l3m) / (GAMMA - 1.0);

   for (int c = 0; c < 4; ++c) H[c] = fp * pf[c] + fm * mf[c];
}

/* Roe-averaged quantities shared by equation.h:481-515 (roe) and 575-616 (hllc) */
struct side
{
   double v[2], v2, vn, p;
};

static void side_state (const double W[4], const double n[2], struct side *s)
{
   s->v2 = 0;
   s->vn = 0;
   for (int d = 0; d < 2; ++d)
   {
      s->v[d] = W[d] / W[RHO];
      s->v2 += s->v[d] * s->v[d];
      s->vn += s->v[d] * n[d];
   }
   s->p = (GAMMA - 1) * (W[ENE] - 0.5 * W[RHO] * s->v2);
}

/* equation.h:469-556 */
static void roe (const double n[2], const double Wl[4], const double Wr[4], double H[4])
{
   const double rl = sqrt (Wl[RHO]), rr = sqrt (Wr[RHO]);
   const double fl = rl / (rl + rr), fr = 1.0 - fl;
   struct side L, R;
   side_state (Wl, n, &L);
   side_state (Wr, n, &R);
   double vel[2], dv[2], veln = 0, v2 = 0, vdotdv = 0;
   for (int d = 0; d < 2; ++d)
   {
      vel[d] = L.v[d] * fl + R.v[d] * fr;
      veln += vel[d] * n[d];
      v2 += vel[d] * vel[d];
      dv[d] = R.v[d] - L.v[d];
      vdotdv += vel[d] * dv[d];
   }
   const double hl = GAMMA * L.p / Wl[RHO] / (GAMMA - 1) + 0.5 * L.v2;
   const double hr = GAMMA * R.p / Wr[RHO] / (GAMMA - 1) + 0.5 * R.v2;
   const double dens = rl * rr;
   const double h = hl * fl + hr * fr;
   const double c = sqrt ((GAMMA - 1.0) * (h - 0.5 * v2));
   const double drho = Wr[RHO] - Wl[RHO];
   const double dp = R.p - L.p;
   const double dvn = R.vn - L.vn;

   const double a1 = (dp - dens * c * dvn) / (2.0 * c * c);
   const double a2 = drho - dp / (c * c);
   const double a3 = (dp + dens * c * dvn) / (2.0 * c * c);

   double l1 = fabs (veln - c), l2 = fabs (veln), l3 = fabs (veln + c);
   const double delta = 0.1 * c; /* Harten fix on the acoustic waves only, 528-531 */
   if (l1 < delta) l1 = 0.5 * (l1 * l1 / delta + delta);
   if (l3 < delta) l3 = 0.5 * (l3 * l3 / delta + delta);

   double D[4];
   D[RHO] = l1 * a1 + l2 * a2 + l3 * a3;
   D[ENE] = l1 * a1 * (h - c * veln) + l2 * a2 * 0.5 * v2 + l2 * dens * (vdotdv - veln * dvn)
            + l3 * a3 * (h + c * veln);
   H[RHO] = 0.5 * (Wl[RHO] * L.vn + Wr[RHO] * R.vn - D[RHO]);
   H[ENE] = 0.5 * (Wl[RHO] * hl * L.vn + Wr[RHO] * hr * R.vn - D[ENE]);
   const double pavg = 0.5 * (L.p + R.p);
   for (int d = 0; d < 2; ++d)
   {
      D[d] = (vel[d] - n[d] * c) * l1 * a1 + vel[d] * l2 * a2 + (dv[d] - n[d] * dvn) * l2 * dens
             + (vel[d] + n[d] * c) * l3 * a3;
      H[d] = n[d] * pavg + 0.5 * (Wl[d] * L.vn + Wr[d] * R.vn) - 0.5 * D[d];
   }
}

/* equation.h:563-681 */
static void hllc (const double n[2], const double Wl[4], const double Wr[4], double H[4])
{
   const double rl = sqrt (Wl[RHO]), rr = sqrt (Wr[RHO]);
   const double fl = rl / (rl + rr), fr = 1.0 - fl;
   struct side L, R;
   side_state (Wl, n, &L);
   side_state (Wr, n, &R);
   double vel[2], veln = 0, v2 = 0;
   for (int d = 0; d < 2; ++d)
   {
      vel[d] = L.v[d] * fl + R.v[d] * fr;
      veln += vel[d] * n[d];
      v2 += vel[d] * vel[d];
   }
   const double hl = (Wl[ENE] + L.p) / Wl[RHO], hr = (Wr[ENE] + R.p) / Wr[RHO];
   const double cl = sqrt (GAMMA * L.p / Wl[RHO]), cr = sqrt (GAMMA * R.p / Wr[RHO]);
   const double el = Wl[ENE] / Wl[RHO], er = Wr[ENE] / Wr[RHO];
   const double h = hl * fl + hr * fr;
   const double c = sqrt ((GAMMA - 1.0) * (h - 0.5 * v2));
   const double sl = dmin (veln - c, L.vn - cl);
   const double sr = dmax (veln + c, R.vn + cr);
   const double sm = (L.p - R.p - Wl[RHO] * L.vn * (sl - L.vn) + Wr[RHO] * R.vn * (sr - R.vn))
                     / (Wr[RHO] * (sr - R.vn) - Wl[RHO] * (sl - L.vn));
   const double ps = Wr[RHO] * (R.vn - sr) * (R.vn - sm) + R.p;

   if (sm >= 0.0)
   {
      if (sl > 0.0)
      {
         H[RHO] = Wl[RHO] * L.vn;
         for (int d = 0; d < 2; ++d) H[d] = Wl[RHO] * L.v[d] * L.vn + L.p * n[d];
         H[ENE] = el * Wl[RHO] * L.vn + L.p * L.vn;
      }
      else
      {
         const double inv = 1.0 / (sl - sm);
         const double smu = sl - L.vn;
         const double rhos = Wl[RHO] * smu * inv;
         double ms[2];
         for (int d = 0; d < 2; ++d) ms[d] = (Wl[RHO] * L.v[d] * smu + (ps - L.p) * n[d]) * inv;
         const double es = (smu * el * Wl[RHO] - L.p * L.vn + ps * sm) * inv;
         H[RHO] = rhos * sm;
         for (int d = 0; d < 2; ++d) H[d] = ms[d] * sm + ps * n[d];
         H[ENE] = (es + ps) * sm;
      }
   }
   else
   {
      if (sr >= 0.0)
      {
         const double inv = 1.0 / (sr - sm);
         const double smu = sr - R.vn;
         const double rhos = Wr[RHO] * smu * inv;
         double ms[2];
         for (int d = 0; d < 2; ++d) ms[d] = (Wr[RHO] * R.v[d] * smu + (ps - R.p) * n[d]) * inv;
         const double es = (smu * er * Wr[RHO] - R.p * R.vn + ps * sm) * inv;
         H[RHO] = rhos * sm;
         for (int d = 0; d < 2; ++d) H[d] = ms[d] * sm + ps * n[d];
         H[ENE] = (es + ps) * sm;
      }
      else
      {
         H[RHO] = Wr[RHO] * R.vn;
         for (int d = 0; d < 2; ++d) H[d] = Wr[RHO] * R.v[d] * R.vn + R.p * n[d];
         H[ENE] = er * Wr[RHO] * R.vn + R.p * R.vn;
      }
   }
}

/* equation.h:686-709: Abramowitz & Stegun 7.1.26, NOT libm erf */
static double as_erf (double xarg)
{
   const double a1 = 0.254829592, a2 = -0.284496736, a3 = 1.421413741, a4 = -1.453152027,
                a5 = 1.061405429, p = 0.3275911;
   int sign = 1;
   if (xarg < 0) sign = -1;
   const double x = fabs (xarg);
   const double t = 1.0 / (1.0 + p * x);
   const double y = 1.0 - (((((a5 * t + a4) * t) + a3) * t + a2) * t + a1) * t * exp (-x * x);
   return sign * y;
}

/* equation.h:714-751 */
static void kinetic_split (int sign, const double n[2], const double W[4], double H[4])
{
   double vn = 0;
   for (int d = 0; d < 2; ++d) vn += W[d] * n[d];
   vn /= W[RHO];
   const double p = phys_pressure (W);
   const double beta = 0.5 * W[RHO] / p;
   const double s = vn * sqrt (beta);
   const double A = 0.5 * (1.0 + sign * as_erf (s));
   const double B = 0.5 * sign * exp (-s * s) / sqrt (M_PI * beta);
   const double uf = vn * A + B;
   for (int d = 0; d < 2; ++d) H[d] = p * n[d] * A + W[d] * uf;
   H[RHO] = W[RHO] * uf;
   H[ENE] = (W[ENE] + p) * vn * A + (W[ENE] + 0.5 * p) * B;
}

/* equation.h:756-782 */
static void kfvs (const double n[2], const double Wp[4], const double Wm[4], double H[4])
{
   double pf[4], mf[4];
   kinetic_split (+1, n, Wp, pf);
   kinetic_split (-1, n, Wm, mf);
   for (int c = 0; c < 4; ++c) H[c] = pf[c] + mf[c];
}

/* claw.h:271-325 */
void phys_numerical_flux (int flux_type, const double n[2], const double Wp[4], const double Wm[4],
                          const double Ap[4], const double Am[4], double out[4])
{
   switch (flux_type)
   {
      case PHYS_FLUX_LXF: lxf (n, Wp, Wm, Ap, Am, out); break;
      case PHYS_FLUX_SW: steger_warming (n, Wp, Wm, out); break;
      case PHYS_FLUX_KFVS: kfvs (n, Wp, Wm, out); break;
      case PHYS_FLUX_ROE: roe (n, Wp, Wm, out); break;
      case PHYS_FLUX_HLLC: hllc (n, Wp, Wm, out); break;
      case PHYS_FLUX_KEP: phys_kep_flux_restated (n, Wp, Wm, Ap, Am, out); break;
      default: out[0] = out[1] = out[2] = out[3] = NAN;
   }
}

/* equation.h:939-1033 */
void phys_wminus (int kind, const double n[2], const double Wp[4], const double g[4], double Wm[4])
{
   switch (kind)
   {
      case PHYS_BC_INFLOW:
      case PHYS_BC_FARFIELD:
         for (int c = 0; c < 4; ++c) Wm[c] = g[c];
         break;
      case PHYS_BC_OUTFLOW:
         for (int c = 0; c < 4; ++c) Wm[c] = Wp[c];
         break;
      case PHYS_BC_PRESSURE:
      {
         const double rho = Wp[RHO];
         double ke = 0;
         for (int d = 0; d < 2; ++d) ke += Wp[d] * Wp[d];
         ke *= 0.5 / rho;
         for (int d = 0; d < 2; ++d) Wm[d] = Wp[d];
         Wm[RHO] = rho;
         Wm[ENE] = g[ENE] / (GAMMA - 1.0) + ke;
         break;
      }
      case PHYS_BC_SLIP:
      {
         double vn = 0;
         for (int d = 0; d < 2; ++d) vn += Wp[d] * n[d];
         for (int d = 0; d < 2; ++d) Wm[d] = Wp[d] - 2.0 * vn * n[d];
         Wm[RHO] = Wp[RHO];
         Wm[ENE] = Wp[ENE];
         break;
      }
      default:
         for (int c = 0; c < 4; ++c) Wm[c] = NAN;
   }
}

/* equation.h:225-265 */
void phys_eigen (const double W[4], double Rx[16], double Lx[16], double Ry[16], double Ly[16])
{
   const double g1 = GAMMA - 1.0;
   const double rho = W[RHO], E = W[ENE];
   const double u = W[0] / rho, v = W[1] / rho;
   const double q2 = u * u + v * v;
   const double p = g1 * (E - 0.5 * rho * q2);
   const double c2 = GAMMA * p / rho;
   const double c = sqrt (c2);
   const double beta = 0.5 / c2;
   const double phi2 = 0.5 * g1 * q2;
   const double h = c2 / g1 + 0.5 * q2;

   const double rx[16] = {1, 0, 1, 1, u, 0, u + c, u - c, v, -1, v, v, 0.5 * q2, -v, h + c * u, h - c * u};
   const double ry[16] = {1, 0, 1, 1, u, 1, u, u, v, 0, v + c, v - c, 0.5 * q2, u, h + c * v, h - c * v};
   const double lx[16] = {1 - phi2 / c2, g1 * u / c2, g1 * v / c2, -g1 / c2,
                          v, 0, -1, 0,
                          beta * (phi2 - c * u), beta * (c - g1 * u), -beta * g1 * v, beta * g1,
                          beta * (phi2 + c * u), -beta * (c + g1 * u), -beta * g1 * v, beta * g1};
   const double ly[16] = {1 - phi2 / c2, g1 * u / c2, g1 * v / c2, -g1 / c2,
                          -u, 1, 0, 0,
                          beta * (phi2 - c * v), -beta * g1 * u, beta * (c - g1 * v), beta * g1,
                          beta * (phi2 + c * v), -beta * g1 * u, -beta * (c + g1 * v), beta * g1};
   for (int i = 0; i < 16; ++i)
   {
      Rx[i] = rx[i];
      Ry[i] = ry[i];
      Lx[i] = lx[i];
      Ly[i] = ly[i];
   }
}

/* equation.h:270-285: reorder to (rho, m_x, m_y, E), multiply by L, result stays in that order */
void phys_to_char (const double L[16], double W[4])
{
   const double V[4] = {W[RHO], W[0], W[1], W[ENE]};
   for (int i = 0; i < 4; ++i)
   {
      W[i] = 0;
      for (int j = 0; j < 4; ++j) W[i] += L[4 * i + j] * V[j];
   }
}

/* equation.h:290-306 */
void phys_to_con (const double R[16], double W[4])
{
   double V[4];
   for (int i = 0; i < 4; ++i)
   {
      V[i] = 0;
      for (int j = 0; j < 4; ++j) V[i] += R[4 * i + j] * W[j];
   }
   W[RHO] = V[0];
   W[ENE] = V[3];
   W[0] = V[1];
   W[1] = V[2];
}
