// Point-wise compressible-Euler arithmetic of the hot path, written for the GPU.
//
// Same formulas as dflo's EulerEquations<2> (reference src/equation.h, lines cited per
// function) but arranged for fp64 throughput on sm_100a: one reciprocal per density instead of
// repeated divisions, shared sub-expressions (exp(-s^2) of the KFVS flux is evaluated once),
// explicit fma where the operation order matters for reproducibility between the two cells
// that evaluate the same face.  Results agree with the reference arithmetic to a few ulp, not
// bit for bit; tests/ state the tolerance.
//
// Everything is DFLO_HD (host+device) so the identical code is exercised on the CPU by
// tests/emu before it ever runs on a GPU.  Components: 0 = rho*u, 1 = rho*v, 2 = rho, 3 = E
// (equation.h:26-28); gamma = 1.4 (equation.cc:33).
#pragma once

#include <math.h>

#if defined(__CUDACC__)
#define DFLO_HD __host__ __device__ __forceinline__
#else
#define DFLO_HD inline
#endif

namespace dflo
{
   constexpr double GAMMA = 1.4;
   constexpr double GM1 = GAMMA - 1.0;
   constexpr int RHO = 2;
   constexpr int ENE = 3;

   enum FluxType { FLUX_LXF = 0, FLUX_SW = 1, FLUX_KFVS = 2, FLUX_ROE = 3, FLUX_HLLC = 4, FLUX_KEP = 5 };
   // fluxes that read the two cell averages besides the two traces (lxf: equation.h:357-359; kep:
   // its dissipation matrix is built from them, src_mpi/equation.h:900)
   DFLO_HD constexpr bool flux_uses_averages (int flux) { return flux == FLUX_LXF || flux == FLUX_KEP; }
   enum BCKind { BC_INFLOW = 0, BC_OUTFLOW = 1, BC_SLIP = 2, BC_PRESSURE = 3, BC_FARFIELD = 4, BC_PERIODIC = 5 };

   // std::max / std::min with the C++ library's exact semantics ((a<b)?b:a and (b<a)?b:a).  They
   // differ from fmax/fmin when an argument is NaN, and the reference's branches on wave speeds
   // computed from non-physical traces (negative density at an unlimited shock) depend on that.
   // Reciprocal, reciprocal square root and square root for the flux kernels.  On the device: the
   // 20-bit hardware seed (MUFU.RCP64H / MUFU.RSQ64H) refined by two Newton steps in fma
   // arithmetic -- within 1-2 ulp of the correctly rounded result for normal arguments, with none
   // of the range checks and slow-path subroutine calls of the IEEE-rounded `1.0/x` / `sqrt(x)`
   // sequences (the flux arguments are densities, sound speeds and their sums; a negative or NaN
   // argument still yields the negative / NaN result the branches of the callers expect).  On the
   // host (tests/emu) the exact operations are used.
   DFLO_HD double fast_rcp (double x)
   {
#if defined(__CUDA_ARCH__)
      double r;
      asm ("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
      double e = fma (-x, r, 1.0);
      r = fma (r, e, r);
      e = fma (-x, r, 1.0);
      return fma (r, e, r);
#else
      return 1.0 / x;
#endif
   }
   DFLO_HD double fast_rsqrt (double x)
   {
#if defined(__CUDA_ARCH__)
      double y;
      asm ("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
      const double hx = 0.5 * x;
      double e = fma (-hx * y, y, 0.5);  // (1 - x y^2)/2
      y = fma (y, e, y);
      e = fma (-hx * y, y, 0.5);
      return fma (y, e, y);
#else
      return 1.0 / sqrt (x);
#endif
   }
   DFLO_HD double fast_sqrt (double x)
   {
#if defined(__CUDA_ARCH__)
      if (x == 0.0) return 0.0;
      const double y = fast_rsqrt (x);
      const double s = x * y;
      return fma (fma (-s, s, x), 0.5 * y, s); // one correction step on the square root itself
#else
      return sqrt (x);
#endif
   }

   DFLO_HD double std_max (double a, double b) { return (a < b) ? b : a; }
   DFLO_HD double std_min (double a, double b) { return (b < a) ? b : a; }

   // equation.h:84-92
   DFLO_HD double pressure (const double W[4])
   {
      return GM1 * (W[ENE] - 0.5 * (W[0] * W[0] + W[1] * W[1]) / W[RHO]);
   }

   // equation.h:142-152
   DFLO_HD double sound_speed (const double W[4]) { return sqrt (GAMMA * pressure (W) / W[RHO]); }

   // equation.h:158-193: Cartesian flux components Fx[c], Fy[c]
   DFLO_HD void flux_matrix (const double W[4], double Fx[4], double Fy[4])
   {
      const double r = fast_rcp (W[RHO]);
      const double u = W[0] * r, v = W[1] * r;
      const double p = GM1 * (W[ENE] - 0.5 * (W[0] * u + W[1] * v));
      Fx[0] = W[0] * u + p;
      Fx[1] = W[1] * u;
      Fx[RHO] = W[0];
      Fx[ENE] = u * (W[ENE] + p);
      Fy[0] = W[0] * v;
      Fy[1] = W[1] * v + p;
      Fy[RHO] = W[1];
      Fy[ENE] = v * (W[ENE] + p);
   }

   // equation.h:829-850: G = (0, -rho, 0, -rho v), unit gravity along -y
   DFLO_HD void forcing (const double W[4], double G[4])
   {
      G[0] = 0.0;
      G[1] = -W[RHO];
      G[RHO] = 0.0;
      G[ENE] = -W[1];
   }

   // src_mpi/equation.h:1189-1202: forcing vector of an external force f = (f_x, f_y); the hard-wired
   // forcing of src/ above is the case f = (0,-1)
   DFLO_HD void forcing_ext (const double W[4], double fx, double fy, double G[4])
   {
      G[0] = W[RHO] * fx;
      G[1] = W[RHO] * fy;
      G[RHO] = 0.0;
      G[ENE] = W[0] * fx + W[1] * fy;
   }

   // |v.n| + c of a cell average, equation.h:119-137
   DFLO_HD double max_eigenvalue_normal (const double A[4], double nx, double ny)
   {
      const double r = fast_rcp (A[RHO]);
      const double p = GM1 * (A[ENE] - 0.5 * (A[0] * A[0] + A[1] * A[1]) * r);
      return fabs ((A[0] * nx + A[1] * ny) * r) + fast_sqrt (GAMMA * p * r);
   }

   // equation.h:324-377
   DFLO_HD void lxf_flux (double nx, double ny, const double Wp[4], const double Wm[4], const double Ap[4],
                          const double Am[4], double H[4])
   {
      const double rp = fast_rcp (Wp[RHO]), rm = fast_rcp (Wm[RHO]);
      const double vnp = (Wp[0] * nx + Wp[1] * ny) * rp;
      const double vnm = (Wm[0] * nx + Wm[1] * ny) * rm;
      const double pp = GM1 * (Wp[ENE] - 0.5 * (Wp[0] * Wp[0] + Wp[1] * Wp[1]) * rp);
      const double pm = GM1 * (Wm[ENE] - 0.5 * (Wm[0] * Wm[0] + Wm[1] * Wm[1]) * rm);
      const double lambda = std_max (max_eigenvalue_normal (Ap, nx, ny), max_eigenvalue_normal (Am, nx, ny));
      const double psum = pp + pm;
      H[0] = 0.5 * (psum * nx + Wp[0] * vnp + Wm[0] * vnm + lambda * (Wp[0] - Wm[0]));
      H[1] = 0.5 * (psum * ny + Wp[1] * vnp + Wm[1] * vnm + lambda * (Wp[1] - Wm[1]));
      H[RHO] = 0.5 * (Wp[RHO] * vnp + Wm[RHO] * vnm + lambda * (Wp[RHO] - Wm[RHO]));
      H[ENE] = 0.5 * ((Wp[ENE] + pp) * vnp + (Wm[ENE] + pm) * vnm + lambda * (Wp[ENE] - Wm[ENE]));
   }

   // equation.h:382-464
   DFLO_HD void steger_warming_flux (double nx, double ny, const double Wp[4], const double Wm[4], double H[4])
   {
      const double rp = fast_rcp (Wp[RHO]), rm = fast_rcp (Wm[RHO]);
      const double up = Wp[0] * rp, vp = Wp[1] * rp, um = Wm[0] * rm, vm = Wm[1] * rm;
      const double vnp = up * nx + vp * ny, vnm = um * nx + vm * ny;
      const double q2p = up * up + vp * vp, q2m = um * um + vm * vm;
      const double pp = GM1 * (Wp[ENE] - 0.5 * Wp[RHO] * q2p);
      const double pm = GM1 * (Wm[ENE] - 0.5 * Wm[RHO] * q2m);
      const double cp = fast_sqrt (GAMMA * pp * rp), cm = fast_sqrt (GAMMA * pm * rm);

      const double l1p = std_max (vnp, 0.0), l2p = std_max (vnp + cp, 0.0), l3p = std_max (vnp - cp, 0.0);
      const double ap = 2.0 * GM1 * l1p + l2p + l3p;
      const double fp = 0.5 * Wp[RHO] / GAMMA;
      const double l1m = std_min (vnm, 0.0), l2m = std_min (vnm + cm, 0.0), l3m = std_min (vnm - cm, 0.0);
      const double am = 2.0 * GM1 * l1m + l2m + l3m;
      const double fm = 0.5 * Wm[RHO] / GAMMA;

      const double dlp = cp * (l2p - l3p), dlm = cm * (l2m - l3m);
      H[0] = fp * (ap * up + dlp * nx) + fm * (am * um + dlm * nx);
      H[1] = fp * (ap * vp + dlp * ny) + fm * (am * vm + dlm * ny);
      H[RHO] = fp * ap + fm * am;
      H[ENE] = fp * (0.5 * ap * q2p + vnp * dlp + cp * cp * (l2p + l3p) / GM1)
               + fm * (0.5 * am * q2m + vnm * dlm + cm * cm * (l2m + l3m) / GM1);
   }

   // equation.h:469-556
   DFLO_HD void roe_flux (double nx, double ny, const double Wl[4], const double Wr[4], double H[4])
   {
      const double sl = fast_sqrt (Wl[RHO]), sr = fast_sqrt (Wr[RHO]);
      const double fl = sl * fast_rcp (sl + sr), fr = 1.0 - fl;
      const double rl = fast_rcp (Wl[RHO]), rr = fast_rcp (Wr[RHO]);
      const double ul = Wl[0] * rl, vl = Wl[1] * rl, ur = Wr[0] * rr, vr = Wr[1] * rr;
      const double v2l = ul * ul + vl * vl, v2r = ur * ur + vr * vr;
      const double vnl = ul * nx + vl * ny, vnr = ur * nx + vr * ny;
      const double u = ul * fl + ur * fr, v = vl * fl + vr * fr;
      const double vn = u * nx + v * ny, v2 = u * u + v * v;
      const double du = ur - ul, dv = vr - vl;
      const double vdotdv = u * du + v * dv;
      const double pl = GM1 * (Wl[ENE] - 0.5 * Wl[RHO] * v2l);
      const double pr = GM1 * (Wr[ENE] - 0.5 * Wr[RHO] * v2r);
      const double hl = (GAMMA / GM1) * pl * rl + 0.5 * v2l;
      const double hr = (GAMMA / GM1) * pr * rr + 0.5 * v2r;
      const double dens = sl * sr;
      const double h = hl * fl + hr * fr;
      const double c2 = GM1 * (h - 0.5 * v2);
      const double c = fast_sqrt (c2);
      const double ic2 = fast_rcp (c2);
      const double drho = Wr[RHO] - Wl[RHO], dp = pr - pl, dvn = vnr - vnl;

      const double a1 = (dp - dens * c * dvn) * (0.5 * ic2);
      const double a2 = drho - dp * ic2;
      const double a3 = (dp + dens * c * dvn) * (0.5 * ic2);

      double l1 = fabs (vn - c), l3 = fabs (vn + c);
      const double l2 = fabs (vn);
      const double delta = 0.1 * c; // Harten fix on the acoustic waves only (528-531)
      const double idelta = fast_rcp (delta);
      if (l1 < delta) l1 = 0.5 * (l1 * l1 * idelta + delta);
      if (l3 < delta) l3 = 0.5 * (l3 * l3 * idelta + delta);

      const double w1 = l1 * a1, w2 = l2 * a2, w3 = l3 * a3, w4 = l2 * dens;
      const double Drho = w1 + w2 + w3;
      const double Dene = w1 * (h - c * vn) + w2 * 0.5 * v2 + w4 * (vdotdv - vn * dvn) + w3 * (h + c * vn);
      const double D0 = (u - nx * c) * w1 + u * w2 + (du - nx * dvn) * w4 + (u + nx * c) * w3;
      const double D1 = (v - ny * c) * w1 + v * w2 + (dv - ny * dvn) * w4 + (v + ny * c) * w3;
      const double pavg = 0.5 * (pl + pr);
      H[RHO] = 0.5 * (Wl[RHO] * vnl + Wr[RHO] * vnr - Drho);
      H[ENE] = 0.5 * (Wl[RHO] * hl * vnl + Wr[RHO] * hr * vnr - Dene);
      H[0] = nx * pavg + 0.5 * (Wl[0] * vnl + Wr[0] * vnr) - 0.5 * D0;
      H[1] = ny * pavg + 0.5 * (Wl[1] * vnl + Wr[1] * vnr) - 0.5 * D1;
   }

   // equation.h:563-681 (HLLC after SU2 v2.0.2)
   DFLO_HD void hllc_flux (double nx, double ny, const double Wl[4], const double Wr[4], double H[4])
   {
      const double sql = fast_sqrt (Wl[RHO]), sqr = fast_sqrt (Wr[RHO]);
      const double fl = sql * fast_rcp (sql + sqr), fr = 1.0 - fl;
      const double rl = fast_rcp (Wl[RHO]), rr = fast_rcp (Wr[RHO]);
      const double ul = Wl[0] * rl, vl = Wl[1] * rl, ur = Wr[0] * rr, vr = Wr[1] * rr;
      const double v2l = ul * ul + vl * vl, v2r = ur * ur + vr * vr;
      const double vnl = ul * nx + vl * ny, vnr = ur * nx + vr * ny;
      const double u = ul * fl + ur * fr, v = vl * fl + vr * fr;
      const double vn = u * nx + v * ny, v2 = u * u + v * v;
      const double pl = GM1 * (Wl[ENE] - 0.5 * Wl[RHO] * v2l);
      const double pr = GM1 * (Wr[ENE] - 0.5 * Wr[RHO] * v2r);
      const double hl = (Wl[ENE] + pl) * rl, hr = (Wr[ENE] + pr) * rr;
      const double cl = fast_sqrt (GAMMA * pl * rl), cr = fast_sqrt (GAMMA * pr * rr);
      const double h = hl * fl + hr * fr;
      const double c = fast_sqrt (GM1 * (h - 0.5 * v2));
      const double s_l = std_min (vn - c, vnl - cl);
      const double s_r = std_max (vn + c, vnr + cr);
      const double ml = Wl[RHO] * (s_l - vnl), mr = Wr[RHO] * (s_r - vnr);
      const double s_m = (pl - pr - ml * vnl + mr * vnr) * fast_rcp (mr - ml);
      const double ps = Wr[RHO] * (vnr - s_r) * (vnr - s_m) + pr;

      if (s_m >= 0.0)
      {
         if (s_l > 0.0)
         {
            H[RHO] = Wl[RHO] * vnl;
            H[0] = Wl[0] * vnl + pl * nx;
            H[1] = Wl[1] * vnl + pl * ny;
            H[ENE] = (Wl[ENE] + pl) * vnl;
         }
         else
         {
            const double inv = fast_rcp (s_l - s_m);
            const double smu = s_l - vnl;
            const double dps = ps - pl;
            H[RHO] = Wl[RHO] * smu * inv * s_m;
            H[0] = (Wl[0] * smu + dps * nx) * inv * s_m + ps * nx;
            H[1] = (Wl[1] * smu + dps * ny) * inv * s_m + ps * ny;
            H[ENE] = ((smu * Wl[ENE] - pl * vnl + ps * s_m) * inv + ps) * s_m;
         }
      }
      else
      {
         if (s_r >= 0.0)
         {
            const double inv = fast_rcp (s_r - s_m);
            const double smu = s_r - vnr;
            const double dps = ps - pr;
            H[RHO] = Wr[RHO] * smu * inv * s_m;
            H[0] = (Wr[0] * smu + dps * nx) * inv * s_m + ps * nx;
            H[1] = (Wr[1] * smu + dps * ny) * inv * s_m + ps * ny;
            H[ENE] = ((smu * Wr[ENE] - pr * vnr + ps * s_m) * inv + ps) * s_m;
         }
         else
         {
            H[RHO] = Wr[RHO] * vnr;
            H[0] = Wr[0] * vnr + pr * nx;
            H[1] = Wr[1] * vnr + pr * ny;
            H[ENE] = (Wr[ENE] + pr) * vnr;
         }
      }
   }

   // equation.h:714-751 with ERF of 686-709 (Abramowitz-Stegun 7.1.26, NOT libm erf)
   DFLO_HD void kinetic_split_flux (double sign, double nx, double ny, const double W[4], double H[4])
   {
      const double r = fast_rcp (W[RHO]);
      const double vn = (W[0] * nx + W[1] * ny) * r;
      const double p = GM1 * (W[ENE] - 0.5 * (W[0] * W[0] + W[1] * W[1]) * r);
      const double beta = 0.5 * W[RHO] * fast_rcp (p);
      const double sb = fast_sqrt (beta);
      const double s = vn * sb;
      const double ex = exp (-s * s);
      // ERF(s)
      const double x = fabs (s);
      const double t = fast_rcp (1.0 + 0.3275911 * x);
      const double y = 1.0
                       - (((((1.061405429 * t + -1.453152027) * t) + 1.421413741) * t + -0.284496736) * t + 0.254829592)
                            * t * ex;
      const double erf_s = (s < 0) ? -y : y;
      const double A = 0.5 * (1.0 + sign * erf_s);
      const double B = 0.5 * sign * ex * fast_rcp (1.7724538509055160273 * sb); // sqrt(pi*beta)
      const double uf = vn * A + B;
      H[0] = p * nx * A + W[0] * uf;
      H[1] = p * ny * A + W[1] * uf;
      H[RHO] = W[RHO] * uf;
      H[ENE] = (W[ENE] + p) * vn * A + (W[ENE] + 0.5 * p) * B;
   }

   // equation.h:756-782
   DFLO_HD void kfvs_flux (double nx, double ny, const double Wp[4], const double Wm[4], double H[4])
   {
      double pf[4], mf[4];
      kinetic_split_flux (+1.0, nx, ny, Wp, pf);
      kinetic_split_flux (-1.0, nx, ny, Wm, mf);
      for (int c = 0; c < 4; ++c) H[c] = pf[c] + mf[c];
   }

   // Kinetic-energy preserving, entropy-stable flux of the MPI tree (src_mpi/equation.h: logavg 27-45,
   // kep_diff_matrix 749-837 built from the two CELL AVERAGES, kep_flux 842-921).  Written once for a
   // general normal; the axis forms below call it with n = (1, 0).
   DFLO_HD double logavg (double a, double b)
   {
      const double xi = b / a;
      const double f = (xi - 1.0) / (xi + 1.0);
      const double u = f * f;
      double F;
      if (u < 1.0e-2)
      {
         const double u2 = u * u, u3 = u2 * u;
         F = 1.0 + u / 3.0 + u2 / 5.0 + u3 / 7.0;
      }
      else
         F = log (xi) / 2.0 / f;
      return 0.5 * (a + b) / F;
   }
   DFLO_HD void kep_flux (double nx, double ny, const double Wl[4], const double Wr[4], const double Al[4], const double Ar[4], double H[4])
   {
      // central part, from the two traces
      double p, beta, rho, vel[2], vel2, betal, betar, vl[2], vr[2], v2l, v2r, pl, pr;
      {
         const double rl = 1.0 / Wl[RHO], rr = 1.0 / Wr[RHO];
         vl[0] = Wl[0] * rl; vl[1] = Wl[1] * rl; vr[0] = Wr[0] * rr; vr[1] = Wr[1] * rr;
         v2l = vl[0] * vl[0] + vl[1] * vl[1];
         v2r = vr[0] * vr[0] + vr[1] * vr[1];
         vel[0] = 0.5 * (vl[0] + vr[0]);
         vel[1] = 0.5 * (vl[1] + vr[1]);
         vel2 = 0.5 * (v2l + v2r);
         pl = GM1 * (Wl[ENE] - 0.5 * Wl[RHO] * v2l);
         pr = GM1 * (Wr[ENE] - 0.5 * Wr[RHO] * v2r);
         betal = 0.5 * Wl[RHO] / pl;
         betar = 0.5 * Wr[RHO] / pr;
         rho = logavg (Wl[RHO], Wr[RHO]);
         beta = logavg (betal, betar);
         p = 0.5 * (Wl[RHO] + Wr[RHO]) / (betal + betar);
      }
      const double vn = vel[0] * nx + vel[1] * ny;
      const double frho = rho * vn;
      const double f0 = nx * p + vel[0] * frho, f1 = ny * p + vel[1] * frho;
      const double fE = 0.5 * (1.0 / (GM1 * beta) - vel2) * frho + f0 * vel[0] + f1 * vel[1];
      // dissipation matrix R |Lambda| S R^T from the two averages
      double Dm[4][4];
      {
         const double rl = 1.0 / Al[RHO], rr = 1.0 / Ar[RHO];
         const double al0 = Al[0] * rl, al1 = Al[1] * rl, ar0 = Ar[0] * rr, ar1 = Ar[1] * rr;
         const double a2l = al0 * al0 + al1 * al1, a2r = ar0 * ar0 + ar1 * ar1;
         const double vnl = al0 * nx + al1 * ny, vnr = ar0 * nx + ar1 * ny;
         const double w0 = 0.5 * (al0 + ar0), w1 = 0.5 * (al1 + ar1);
         const double wn = w0 * nx + w1 * ny, w2 = w0 * w0 + w1 * w1;
         const double qpl = GM1 * (Al[ENE] - 0.5 * Al[RHO] * a2l), qpr = GM1 * (Ar[ENE] - 0.5 * Ar[RHO] * a2r);
         const double bl = 0.5 * Al[RHO] / qpl, br = 0.5 * Ar[RHO] / qpr;
         const double rhoA = logavg (Al[RHO], Ar[RHO]);
         const double betaA = logavg (bl, br);
         const double a = sqrt (0.5 * GAMMA / betaA);
         const double pA = 0.5 * (Al[RHO] + Ar[RHO]) / (bl + br);
         const double Hh = a * a / GM1 + 0.5 * w2;
         const double v1 = w0 * ny - w1 * nx;
         const double R[4][4] = {{1, 1, 0, 1},
                                 {w0 - a * nx, w0, ny, w0 + a * nx},
                                 {w1 - a * ny, w1, -nx, w1 + a * ny},
                                 {Hh - a * wn, 0.5 * w2, v1, Hh + a * wn}};
         const double cl = sqrt (GAMMA * qpl / Al[RHO]), cr = sqrt (GAMMA * qpr / Ar[RHO]);
         const double L0 = fabs (wn - a) + (1.0 / 6.0) * fabs ((vnl - cl) - (vnr - cr));
         const double L3 = fabs (wn + a) + (1.0 / 6.0) * fabs ((vnl + cl) - (vnr + cr));
         const double L12 = fabs (wn);
         const double Dd[4] = {L0 * (0.5 * rhoA / GAMMA), L12 * (GM1 * rhoA / GAMMA), L12 * pA, L3 * (0.5 * rhoA / GAMMA)};
         for (int i = 0; i < 4; ++i)
            for (int j = i; j < 4; ++j)
            {
               double s = 0;
               for (int k = 0; k < 4; ++k) s += R[i][k] * Dd[k] * R[j][k];
               Dm[i][j] = Dm[j][i] = s;
            }
      }
      // jump in the entropy variables of the traces
      const double ds = log (pr / pl) - GAMMA * log (Wr[RHO] / Wl[RHO]);
      const double dV[4] = {-ds / GM1 - (betar * v2r - betal * v2l), 2.0 * (betar * vr[0] - betal * vl[0]),
                            2.0 * (betar * vr[1] - betal * vl[1]), -2.0 * (betar - betal)};
      double Diff[4] = {0.0, 0.0, 0.0, 0.0};
      for (int i = 0; i < 4; ++i)
         for (int j = 0; j < 4; ++j) Diff[i] += Dm[i][j] * dV[j];
      H[RHO] = frho - 0.5 * Diff[0];
      H[0] = f0 - 0.5 * Diff[1];
      H[1] = f1 - 0.5 * Diff[2];
      H[ENE] = fE - 0.5 * Diff[3];
   }

   // claw.h:271-325 with the switch resolved at compile time
   template <int FLUX>
   DFLO_HD void numerical_flux (double nx, double ny, const double Wp[4], const double Wm[4], const double Ap[4],
                                const double Am[4], double H[4])
   {
      if (FLUX == FLUX_LXF)
         lxf_flux (nx, ny, Wp, Wm, Ap, Am, H);
      else if (FLUX == FLUX_SW)
         steger_warming_flux (nx, ny, Wp, Wm, H);
      else if (FLUX == FLUX_KFVS)
         kfvs_flux (nx, ny, Wp, Wm, H);
      else if (FLUX == FLUX_ROE)
         roe_flux (nx, ny, Wp, Wm, H);
      else if (FLUX == FLUX_KEP)
         kep_flux (nx, ny, Wp, Wm, Ap, Am, H);
      else
         hllc_flux (nx, ny, Wp, Wm, H);
   }

   //---------------------------------------------------------------------------------------------
   // Axis-aligned forms.  On the Cartesian cells the engine accepts every face normal is +-e_x or
   // +-e_y, and every flux of claw.h:271-325 is antisymmetric, H(-n, Wb, Wa) = -H(n, Wa, Wb), so a
   // face is always solved along +e_x (or +e_y, by exchanging the two momentum components) with
   // the LOW-side cell's trace as the first state.  These are the general formulas above with
   // n = (1,0) substituted (x*1 = x, x*0 = 0 exactly) and the vanishing terms dropped.
   //---------------------------------------------------------------------------------------------
   // s = sqrt(x) and r = 1/x from ONE reciprocal-square-root seed
   DFLO_HD void sqrt_and_rcp (double x, double &s, double &r)
   {
#if defined(__CUDA_ARCH__)
      const double y = fast_rsqrt (x);
      const double s0 = x * y;
      s = fma (fma (-s0, s0, x), 0.5 * y, s0);
      r = y * y;
#else
      s = sqrt (x);
      r = 1.0 / x;
#endif
   }

   // equation.h:158-193, x column only: F_x(W)
   DFLO_HD void flux_x (const double W[4], double Fx[4])
   {
      const double r = fast_rcp (W[RHO]);
      const double u = W[0] * r, v = W[1] * r;
      const double p = GM1 * (W[ENE] - 0.5 * (W[0] * u + W[1] * v));
      Fx[0] = W[0] * u + p;
      Fx[1] = W[1] * u;
      Fx[RHO] = W[0];
      Fx[ENE] = u * (W[ENE] + p);
   }

   DFLO_HD double max_eigenvalue_x (const double A[4])
   {
      const double r = fast_rcp (A[RHO]);
      const double p = GM1 * (A[ENE] - 0.5 * (A[0] * A[0] + A[1] * A[1]) * r);
      return fabs (A[0] * r) + fast_sqrt (GAMMA * p * r);
   }

   DFLO_HD void lxf_flux_x (const double Wp[4], const double Wm[4], const double Ap[4], const double Am[4], double H[4])
   {
      const double rp = fast_rcp (Wp[RHO]), rm = fast_rcp (Wm[RHO]);
      const double vnp = Wp[0] * rp, vnm = Wm[0] * rm;
      const double pp = GM1 * (Wp[ENE] - 0.5 * (Wp[0] * Wp[0] + Wp[1] * Wp[1]) * rp);
      const double pm = GM1 * (Wm[ENE] - 0.5 * (Wm[0] * Wm[0] + Wm[1] * Wm[1]) * rm);
      const double lambda = std_max (max_eigenvalue_x (Ap), max_eigenvalue_x (Am));
      H[0] = 0.5 * ((pp + pm) + Wp[0] * vnp + Wm[0] * vnm + lambda * (Wp[0] - Wm[0]));
      H[1] = 0.5 * (Wp[1] * vnp + Wm[1] * vnm + lambda * (Wp[1] - Wm[1]));
      H[RHO] = 0.5 * (Wp[RHO] * vnp + Wm[RHO] * vnm + lambda * (Wp[RHO] - Wm[RHO]));
      H[ENE] = 0.5 * ((Wp[ENE] + pp) * vnp + (Wm[ENE] + pm) * vnm + lambda * (Wp[ENE] - Wm[ENE]));
   }

   DFLO_HD void steger_warming_flux_x (const double Wp[4], const double Wm[4], double H[4])
   {
      const double rp = fast_rcp (Wp[RHO]), rm = fast_rcp (Wm[RHO]);
      const double up = Wp[0] * rp, vp = Wp[1] * rp, um = Wm[0] * rm, vm = Wm[1] * rm;
      const double q2p = up * up + vp * vp, q2m = um * um + vm * vm;
      const double pp = GM1 * (Wp[ENE] - 0.5 * Wp[RHO] * q2p);
      const double pm = GM1 * (Wm[ENE] - 0.5 * Wm[RHO] * q2m);
      const double cp = fast_sqrt (GAMMA * pp * rp), cm = fast_sqrt (GAMMA * pm * rm);
      const double l1p = std_max (up, 0.0), l2p = std_max (up + cp, 0.0), l3p = std_max (up - cp, 0.0);
      const double ap = 2.0 * GM1 * l1p + l2p + l3p;
      const double fp = 0.5 * Wp[RHO] / GAMMA;
      const double l1m = std_min (um, 0.0), l2m = std_min (um + cm, 0.0), l3m = std_min (um - cm, 0.0);
      const double am = 2.0 * GM1 * l1m + l2m + l3m;
      const double fm = 0.5 * Wm[RHO] / GAMMA;
      const double dlp = cp * (l2p - l3p), dlm = cm * (l2m - l3m);
      H[0] = fp * (ap * up + dlp) + fm * (am * um + dlm);
      H[1] = fp * (ap * vp) + fm * (am * vm);
      H[RHO] = fp * ap + fm * am;
      H[ENE] = fp * (0.5 * ap * q2p + up * dlp + cp * cp * (l2p + l3p) / GM1)
               + fm * (0.5 * am * q2m + um * dlm + cm * cm * (l2m + l3m) / GM1);
   }

   DFLO_HD void roe_flux_x (const double Wl[4], const double Wr[4], double H[4])
   {
      double sl, rl, sr, rr;
      sqrt_and_rcp (Wl[RHO], sl, rl);
      sqrt_and_rcp (Wr[RHO], sr, rr);
      const double fl = sl * fast_rcp (sl + sr), fr = 1.0 - fl;
      const double ul = Wl[0] * rl, vl = Wl[1] * rl, ur = Wr[0] * rr, vr = Wr[1] * rr;
      const double v2l = ul * ul + vl * vl, v2r = ur * ur + vr * vr;
      const double u = ul * fl + ur * fr, v = vl * fl + vr * fr;
      const double v2 = u * u + v * v;
      const double du = ur - ul, dv = vr - vl;
      const double pl = GM1 * (Wl[ENE] - 0.5 * Wl[RHO] * v2l);
      const double pr = GM1 * (Wr[ENE] - 0.5 * Wr[RHO] * v2r);
      const double hl = (GAMMA / GM1) * pl * rl + 0.5 * v2l;
      const double hr = (GAMMA / GM1) * pr * rr + 0.5 * v2r;
      const double dens = sl * sr;
      const double h = hl * fl + hr * fr;
      const double c2 = GM1 * (h - 0.5 * v2);
      double c, ic2;
      sqrt_and_rcp (c2, c, ic2);
      const double drho = Wr[RHO] - Wl[RHO], dp = pr - pl;
      const double t = dens * c * du;
      const double a1 = (dp - t) * (0.5 * ic2);
      const double a2 = drho - dp * ic2;
      const double a3 = (dp + t) * (0.5 * ic2);
      double l1 = fabs (u - c), l3 = fabs (u + c);
      const double l2 = fabs (u);
      const double delta = 0.1 * c;            // Harten fix on the acoustic waves only (528-531)
      const double idelta = 10.0 * (c * ic2);  // 1 / delta
      if (l1 < delta) l1 = 0.5 * (l1 * l1 * idelta + delta);
      if (l3 < delta) l3 = 0.5 * (l3 * l3 * idelta + delta);
      const double w1 = l1 * a1, w2 = l2 * a2, w3 = l3 * a3, w4 = l2 * dens;
      const double Drho = w1 + w2 + w3;
      const double cu = c * u;
      const double Dene = w1 * (h - cu) + w2 * 0.5 * v2 + w4 * (v * dv) + w3 * (h + cu);
      const double D0 = (u - c) * w1 + u * w2 + (u + c) * w3;
      const double D1 = v * Drho + dv * w4;
      H[RHO] = 0.5 * (Wl[0] + Wr[0] - Drho);
      H[ENE] = 0.5 * ((Wl[ENE] + pl) * ul + (Wr[ENE] + pr) * ur - Dene);
      H[0] = 0.5 * (pl + pr) + 0.5 * (Wl[0] * ul + Wr[0] * ur) - 0.5 * D0;
      H[1] = 0.5 * (Wl[1] * ul + Wr[1] * ur) - 0.5 * D1;
   }

   DFLO_HD void hllc_flux_x (const double Wl[4], const double Wr[4], double H[4])
   {
      double sql, rl, sqr, rr;
      sqrt_and_rcp (Wl[RHO], sql, rl);
      sqrt_and_rcp (Wr[RHO], sqr, rr);
      const double fl = sql * fast_rcp (sql + sqr), fr = 1.0 - fl;
      const double ul = Wl[0] * rl, vl = Wl[1] * rl, ur = Wr[0] * rr, vr = Wr[1] * rr;
      const double v2l = ul * ul + vl * vl, v2r = ur * ur + vr * vr;
      const double u = ul * fl + ur * fr, v = vl * fl + vr * fr;
      const double v2 = u * u + v * v;
      const double pl = GM1 * (Wl[ENE] - 0.5 * Wl[RHO] * v2l);
      const double pr = GM1 * (Wr[ENE] - 0.5 * Wr[RHO] * v2r);
      const double hl = (Wl[ENE] + pl) * rl, hr = (Wr[ENE] + pr) * rr;
      const double cl = fast_sqrt (GAMMA * pl * rl), cr = fast_sqrt (GAMMA * pr * rr);
      const double h = hl * fl + hr * fr;
      const double c = fast_sqrt (GM1 * (h - 0.5 * v2));
      const double s_l = std_min (u - c, ul - cl);
      const double s_r = std_max (u + c, ur + cr);
      const double ml = Wl[RHO] * (s_l - ul), mr = Wr[RHO] * (s_r - ur);
      const double s_m = (pl - pr - ml * ul + mr * ur) * fast_rcp (mr - ml);
      const double ps = Wr[RHO] * (ur - s_r) * (ur - s_m) + pr;
      if (s_m >= 0.0)
      {
         if (s_l > 0.0)
         {
            H[RHO] = Wl[RHO] * ul;
            H[0] = Wl[0] * ul + pl;
            H[1] = Wl[1] * ul;
            H[ENE] = (Wl[ENE] + pl) * ul;
         }
         else
         {
            const double inv = fast_rcp (s_l - s_m);
            const double smu = s_l - ul;
            H[RHO] = Wl[RHO] * smu * inv * s_m;
            H[0] = (Wl[0] * smu + (ps - pl)) * inv * s_m + ps;
            H[1] = (Wl[1] * smu) * inv * s_m;
            H[ENE] = ((smu * Wl[ENE] - pl * ul + ps * s_m) * inv + ps) * s_m;
         }
      }
      else
      {
         if (s_r >= 0.0)
         {
            const double inv = fast_rcp (s_r - s_m);
            const double smu = s_r - ur;
            H[RHO] = Wr[RHO] * smu * inv * s_m;
            H[0] = (Wr[0] * smu + (ps - pr)) * inv * s_m + ps;
            H[1] = (Wr[1] * smu) * inv * s_m;
            H[ENE] = ((smu * Wr[ENE] - pr * ur + ps * s_m) * inv + ps) * s_m;
         }
         else
         {
            H[RHO] = Wr[RHO] * ur;
            H[0] = Wr[0] * ur + pr;
            H[1] = Wr[1] * ur;
            H[ENE] = (Wr[ENE] + pr) * ur;
         }
      }
   }

   DFLO_HD void kinetic_split_flux_x (double sign, const double W[4], double H[4])
   {
      const double r = fast_rcp (W[RHO]);
      const double vn = W[0] * r;
      const double p = GM1 * (W[ENE] - 0.5 * (W[0] * W[0] + W[1] * W[1]) * r);
      const double beta = 0.5 * W[RHO] * fast_rcp (p);
      const double sb = fast_sqrt (beta);
      const double s = vn * sb;
      const double ex = exp (-s * s);
      const double x = fabs (s);
      const double t = fast_rcp (1.0 + 0.3275911 * x);
      const double y = 1.0
                       - (((((1.061405429 * t + -1.453152027) * t) + 1.421413741) * t + -0.284496736) * t + 0.254829592)
                            * t * ex;
      const double erf_s = (s < 0) ? -y : y;
      const double A = 0.5 * (1.0 + sign * erf_s);
      const double B = 0.5 * sign * ex * fast_rcp (1.7724538509055160273 * sb);
      const double uf = vn * A + B;
      H[0] = p * A + W[0] * uf;
      H[1] = W[1] * uf;
      H[RHO] = W[RHO] * uf;
      H[ENE] = (W[ENE] + p) * vn * A + (W[ENE] + 0.5 * p) * B;
   }

   DFLO_HD void kfvs_flux_x (const double Wp[4], const double Wm[4], double H[4])
   {
      double pf[4], mf[4];
      kinetic_split_flux_x (+1.0, Wp, pf);
      kinetic_split_flux_x (-1.0, Wm, mf);
      for (int c = 0; c < 4; ++c) H[c] = pf[c] + mf[c];
   }

   // Numerical flux along +e_DIR (DIR = 0: x, 1: y) between the low-side state Wl and the
   // high-side state Wr; Al, Ar = the two cell averages (read by LxF only).  The y form is the x
   // form with the momentum components exchanged on the way in and out.
   template <int FLUX, int DIR>
   DFLO_HD void numerical_flux_axis (const double Wl[4], const double Wr[4], const double Al[4], const double Ar[4], double H[4])
   {
      const double L[4] = {Wl[DIR], Wl[1 - DIR], Wl[RHO], Wl[ENE]};
      const double R[4] = {Wr[DIR], Wr[1 - DIR], Wr[RHO], Wr[ENE]};
      double G[4];
      if (flux_uses_averages (FLUX))
      {
         const double AL[4] = {Al[DIR], Al[1 - DIR], Al[RHO], Al[ENE]};
         const double AR[4] = {Ar[DIR], Ar[1 - DIR], Ar[RHO], Ar[ENE]};
         if (FLUX == FLUX_LXF)
            lxf_flux_x (L, R, AL, AR, G);
         else
            kep_flux (1.0, 0.0, L, R, AL, AR, G);
      }
      else if (FLUX == FLUX_SW)
         steger_warming_flux_x (L, R, G);
      else if (FLUX == FLUX_KFVS)
         kfvs_flux_x (L, R, G);
      else if (FLUX == FLUX_ROE)
         roe_flux_x (L, R, G);
      else
         hllc_flux_x (L, R, G);
      H[DIR] = G[0];
      H[1 - DIR] = G[1];
      H[RHO] = G[RHO];
      H[ENE] = G[ENE];
   }

   // The flux through a face with normal direction e_DIR exactly as the reference evaluates it:
   // numerical_normal_flux (n, W+, W-) with W+ the trace of the cell that integrates the face
   // ("plus") and n its outward normal.  plus_low: that cell is the one on the low-coordinate side
   // (n = +e_DIR); otherwise n = -e_DIR, which is the +e_DIR problem mirrored in DIR (normal
   // momentum negated on the way in and out -- every intermediate is identical or negated, so
   // the result is bit-for-bit the mirrored one).  Returns the flux along +e_DIR.
   template <int FLUX, int DIR>
   DFLO_HD void face_flux_axis (bool plus_low, const double Wlo[4], const double Whi[4], const double Alo[4], const double Ahi[4],
                                double H[4])
   {
      const double sg = plus_low ? 1.0 : -1.0;
      double P[4], M[4], AP[4], AM[4], G[4];
#pragma unroll
      for (int c = 0; c < 4; ++c)
      {
         P[c] = plus_low ? Wlo[c] : Whi[c];
         M[c] = plus_low ? Whi[c] : Wlo[c];
      }
      P[DIR] *= sg;
      M[DIR] *= sg;
      if (flux_uses_averages (FLUX))
      {
#pragma unroll
         for (int c = 0; c < 4; ++c)
         {
            AP[c] = plus_low ? Alo[c] : Ahi[c];
            AM[c] = plus_low ? Ahi[c] : Alo[c];
         }
         AP[DIR] *= sg;
         AM[DIR] *= sg;
      }
      numerical_flux_axis<FLUX, DIR> (P, M, AP, AM, G);
      // reference flux along n = sg*e_DIR is (sg*G[DIR], G[others]); along +e_DIR: times sg
#pragma unroll
      for (int c = 0; c < 4; ++c) H[c] = sg * G[c];
      H[DIR] = G[DIR];
   }

   // equation.h:939-1033
   DFLO_HD void compute_wminus (int kind, double nx, double ny, const double Wp[4], const double g[4], double Wm[4])
   {
      if (kind == BC_INFLOW || kind == BC_FARFIELD)
      {
         for (int c = 0; c < 4; ++c) Wm[c] = g[c];
      }
      else if (kind == BC_SLIP)
      {
         const double vdotn = Wp[0] * nx + Wp[1] * ny;
         Wm[0] = Wp[0] - 2.0 * vdotn * nx;
         Wm[1] = Wp[1] - 2.0 * vdotn * ny;
         Wm[RHO] = Wp[RHO];
         Wm[ENE] = Wp[ENE];
      }
      else if (kind == BC_PRESSURE)
      {
         const double ke = 0.5 * (Wp[0] * Wp[0] + Wp[1] * Wp[1]) / Wp[RHO];
         Wm[0] = Wp[0];
         Wm[1] = Wp[1];
         Wm[RHO] = Wp[RHO];
         Wm[ENE] = g[ENE] / GM1 + ke;
      }
      else // outflow
      {
         for (int c = 0; c < 4; ++c) Wm[c] = Wp[c];
      }
   }

   // Left/right eigenvector matrices at a cell average, equation.h:225-265.  Stored in the
   // reordered variable order (rho, m_x, m_y, E) that transform_to_char/con use (270-306).
   struct EigenMatrices
   {
      double Rx[4][4], Lx[4][4], Ry[4][4], Ly[4][4];
   };

   DFLO_HD void compute_eigen_matrix (const double W[4], EigenMatrices &m)
   {
      const double g1 = GM1;
      const double rho = W[RHO], E = W[ENE], ir = fast_rcp (rho); // reciprocals: see compute_eigen_left
      const double u = W[0] * ir, v = W[1] * ir;
      const double q2 = u * u + v * v;
      const double p = g1 * (E - 0.5 * rho * q2);
      const double c2 = GAMMA * p * ir;
      const double c = fast_sqrt (c2);
      const double ic2 = fast_rcp (c2);
      const double beta = 0.5 * ic2;
      const double phi2 = 0.5 * g1 * q2;
      const double h = c2 * (1.0 / g1) + 0.5 * q2;

      m.Rx[0][0] = 1;        m.Rx[0][1] = 0;   m.Rx[0][2] = 1;         m.Rx[0][3] = 1;
      m.Rx[1][0] = u;        m.Rx[1][1] = 0;   m.Rx[1][2] = u + c;     m.Rx[1][3] = u - c;
      m.Rx[2][0] = v;        m.Rx[2][1] = -1;  m.Rx[2][2] = v;         m.Rx[2][3] = v;
      m.Rx[3][0] = 0.5 * q2; m.Rx[3][1] = -v;  m.Rx[3][2] = h + c * u; m.Rx[3][3] = h - c * u;

      m.Ry[0][0] = 1;        m.Ry[0][1] = 0;   m.Ry[0][2] = 1;         m.Ry[0][3] = 1;
      m.Ry[1][0] = u;        m.Ry[1][1] = 1;   m.Ry[1][2] = u;         m.Ry[1][3] = u;
      m.Ry[2][0] = v;        m.Ry[2][1] = 0;   m.Ry[2][2] = v + c;     m.Ry[2][3] = v - c;
      m.Ry[3][0] = 0.5 * q2; m.Ry[3][1] = u;   m.Ry[3][2] = h + c * v; m.Ry[3][3] = h - c * v;

      m.Lx[0][0] = 1 - phi2 * ic2;        m.Lx[0][1] = g1 * u * ic2;         m.Lx[0][2] = g1 * v * ic2;   m.Lx[0][3] = -g1 * ic2;
      m.Lx[1][0] = v;                     m.Lx[1][1] = 0;                    m.Lx[1][2] = -1;             m.Lx[1][3] = 0;
      m.Lx[2][0] = beta * (phi2 - c * u); m.Lx[2][1] = beta * (c - g1 * u);  m.Lx[2][2] = -beta * g1 * v; m.Lx[2][3] = beta * g1;
      m.Lx[3][0] = beta * (phi2 + c * u); m.Lx[3][1] = -beta * (c + g1 * u); m.Lx[3][2] = -beta * g1 * v; m.Lx[3][3] = beta * g1;

      m.Ly[0][0] = 1 - phi2 * ic2;        m.Ly[0][1] = g1 * u * ic2;   m.Ly[0][2] = g1 * v * ic2;         m.Ly[0][3] = -g1 * ic2;
      m.Ly[1][0] = -u;                    m.Ly[1][1] = 1;              m.Ly[1][2] = 0;                    m.Ly[1][3] = 0;
      m.Ly[2][0] = beta * (phi2 - c * v); m.Ly[2][1] = -beta * g1 * u; m.Ly[2][2] = beta * (c - g1 * v);  m.Ly[2][3] = beta * g1;
      m.Ly[3][0] = beta * (phi2 + c * v); m.Ly[3][1] = -beta * g1 * u; m.Ly[3][2] = -beta * (c + g1 * v); m.Ly[3][3] = beta * g1;
   }

   // The same matrices in two halves (identical expressions, hence identical values): the limiter
   // needs the left eigenvectors of every cell but the right ones only where it actually limits.
   struct EigenLeft
   {
      double Lx[4][4], Ly[4][4];
   };
   struct EigenRight
   {
      double Rx[4][4], Ry[4][4];
   };
   DFLO_HD void compute_eigen_left (const double W[4], EigenLeft &m)
   {
      const double g1 = GM1;
      // one reciprocal of the density and one of c^2 instead of the reference's eleven divisions (the device has no
      // fp64 divider: each IEEE quotient is a ~20-instruction sequence)
      const double rho = W[RHO], E = W[ENE], ir = fast_rcp (rho);
      const double u = W[0] * ir, v = W[1] * ir;
      const double q2 = u * u + v * v;
      const double p = g1 * (E - 0.5 * rho * q2);
      const double c2 = GAMMA * p * ir;
      const double c = fast_sqrt (c2);
      const double ic2 = fast_rcp (c2);
      const double beta = 0.5 * ic2;
      const double phi2 = 0.5 * g1 * q2;

      m.Lx[0][0] = 1 - phi2 * ic2;        m.Lx[0][1] = g1 * u * ic2;         m.Lx[0][2] = g1 * v * ic2;   m.Lx[0][3] = -g1 * ic2;
      m.Lx[1][0] = v;                     m.Lx[1][1] = 0;                    m.Lx[1][2] = -1;             m.Lx[1][3] = 0;
      m.Lx[2][0] = beta * (phi2 - c * u); m.Lx[2][1] = beta * (c - g1 * u);  m.Lx[2][2] = -beta * g1 * v; m.Lx[2][3] = beta * g1;
      m.Lx[3][0] = beta * (phi2 + c * u); m.Lx[3][1] = -beta * (c + g1 * u); m.Lx[3][2] = -beta * g1 * v; m.Lx[3][3] = beta * g1;

      m.Ly[0][0] = 1 - phi2 * ic2;        m.Ly[0][1] = g1 * u * ic2;   m.Ly[0][2] = g1 * v * ic2;         m.Ly[0][3] = -g1 * ic2;
      m.Ly[1][0] = -u;                    m.Ly[1][1] = 1;              m.Ly[1][2] = 0;                    m.Ly[1][3] = 0;
      m.Ly[2][0] = beta * (phi2 - c * v); m.Ly[2][1] = -beta * g1 * u; m.Ly[2][2] = beta * (c - g1 * v);  m.Ly[2][3] = beta * g1;
      m.Ly[3][0] = beta * (phi2 + c * v); m.Ly[3][1] = -beta * g1 * u; m.Ly[3][2] = -beta * (c + g1 * v); m.Ly[3][3] = beta * g1;
   }
   DFLO_HD void compute_eigen_right (const double W[4], EigenRight &m)
   {
      const double g1 = GM1;
      const double rho = W[RHO], E = W[ENE], ir = fast_rcp (rho);
      const double u = W[0] * ir, v = W[1] * ir;
      const double q2 = u * u + v * v;
      const double p = g1 * (E - 0.5 * rho * q2);
      const double c2 = GAMMA * p * ir;
      const double c = fast_sqrt (c2);
      const double h = c2 * (1.0 / g1) + 0.5 * q2;

      m.Rx[0][0] = 1;        m.Rx[0][1] = 0;   m.Rx[0][2] = 1;         m.Rx[0][3] = 1;
      m.Rx[1][0] = u;        m.Rx[1][1] = 0;   m.Rx[1][2] = u + c;     m.Rx[1][3] = u - c;
      m.Rx[2][0] = v;        m.Rx[2][1] = -1;  m.Rx[2][2] = v;         m.Rx[2][3] = v;
      m.Rx[3][0] = 0.5 * q2; m.Rx[3][1] = -v;  m.Rx[3][2] = h + c * u; m.Rx[3][3] = h - c * u;

      m.Ry[0][0] = 1;        m.Ry[0][1] = 0;   m.Ry[0][2] = 1;         m.Ry[0][3] = 1;
      m.Ry[1][0] = u;        m.Ry[1][1] = 1;   m.Ry[1][2] = u;         m.Ry[1][3] = u;
      m.Ry[2][0] = v;        m.Ry[2][1] = 0;   m.Ry[2][2] = v + c;     m.Ry[2][3] = v - c;
      m.Ry[3][0] = 0.5 * q2; m.Ry[3][1] = u;   m.Ry[3][2] = h + c * v; m.Ry[3][3] = h - c * v;
   }

   // src_mpi/equation.h:299-335: eigenvector matrices along the streamline direction (kx,ky), the
   // projection of the minmax limiter (src_mpi/limiter.cc:450).  The reference obtains (kx,ky) as
   // (cos,sin)(atan2(v,u)); here it is the normalised velocity, (1,0) for a gas at rest -- the same
   // direction to round-off without the fp64 trigonometric slow paths.
   struct EigenStream
   {
      double R[4][4], L[4][4];
   };
   DFLO_HD void compute_eigen_stream (const double W[4], EigenStream &m)
   {
      const double g1 = GM1;
      const double rho = W[RHO], E = W[ENE], ir = fast_rcp (rho); // reciprocals: see compute_eigen_left
      const double u = W[0] * ir, v = W[1] * ir;
      const double q2 = u * u + v * v;
      const double p = g1 * (E - 0.5 * rho * q2);
      const double c2 = GAMMA * p * ir;
      const double c = fast_sqrt (c2);
      const double ic2 = fast_rcp (c2);
      const double beta = 0.5 * ic2;
      const double phi2 = 0.5 * g1 * q2;
      const double h = c2 * (1.0 / g1) + 0.5 * q2;
      const double q = sqrt (q2);
      const double kx = (q > 0.0) ? u / q : 1.0, ky = (q > 0.0) ? v / q : 0.0;
      const double uk = u * kx + v * ky;

      m.R[0][0] = 1;        m.R[0][1] = 0;               m.R[0][2] = 1;          m.R[0][3] = 1;
      m.R[1][0] = u;        m.R[1][1] = ky;              m.R[1][2] = u + kx * c; m.R[1][3] = u - kx * c;
      m.R[2][0] = v;        m.R[2][1] = -kx;             m.R[2][2] = v + ky * c; m.R[2][3] = v - ky * c;
      m.R[3][0] = 0.5 * q2; m.R[3][1] = ky * u - kx * v; m.R[3][2] = h + c * uk; m.R[3][3] = h - c * uk;

      m.L[0][0] = 1 - phi2 / c2;          m.L[0][1] = g1 * u / c2;               m.L[0][2] = g1 * v / c2;               m.L[0][3] = -g1 / c2;
      m.L[1][0] = -(ky * u - kx * v);     m.L[1][1] = ky;                        m.L[1][2] = -kx;                       m.L[1][3] = 0;
      m.L[2][0] = beta * (phi2 - c * uk); m.L[2][1] = beta * (kx * c - g1 * u);  m.L[2][2] = beta * (ky * c - g1 * v);  m.L[2][3] = beta * g1;
      m.L[3][0] = beta * (phi2 + c * uk); m.L[3][1] = -beta * (kx * c + g1 * u); m.L[3][2] = -beta * (ky * c + g1 * v); m.L[3][3] = beta * g1;
   }

   // equation.h:270-285: W (conserved order) -> characteristic (result in matrix row order)
   DFLO_HD void transform_to_char (const double L[4][4], double W[4])
   {
      const double V[4] = {W[RHO], W[0], W[1], W[ENE]};
      for (int i = 0; i < 4; ++i)
      {
         double s = 0.0;
         for (int j = 0; j < 4; ++j) s += L[i][j] * V[j];
         W[i] = s;
      }
   }

   // equation.h:290-306
   DFLO_HD void transform_to_con (const double R[4][4], double W[4])
   {
      double V[4];
      for (int i = 0; i < 4; ++i)
      {
         double s = 0.0;
         for (int j = 0; j < 4; ++j) s += R[i][j] * W[j];
         V[i] = s;
      }
      W[RHO] = V[0];
      W[ENE] = V[3];
      W[0] = V[1];
      W[1] = V[2];
   }

   // TVB minmod, limiter.cc:15-30
   DFLO_HD double minmod (double a, double b, double c, double Mdx2)
   {
      const double aa = fabs (a);
      if (aa < Mdx2) return a;
      if (a * b > 0 && b * c > 0)
      {
         const double s = (a > 0) ? 1.0 : -1.0;
         return s * std_min (aa, std_min (fabs (b), fabs (c)));
      }
      return 0.0;
   }
}
