// TEST INFRASTRUCTURE ONLY (oracle/). Built only where /root/reference exists, into oracle/_ref/.
//
// phys.h implemented by calling the reference's OWN object code: this file #includes
// /root/reference/src/equation.h unmodified (through -I/root/reference/src and the deal.II
// stub under oracle/dealii_stub) and forwards to EulerEquations<2>. No reference source is
// copied into this repository; the only definition supplied here is the out-of-class constant
// gas_gamma, which the reference defines in src/equation.cc:33 (that file also holds the
// deal.II Postprocessor and cannot be compiled without deal.II).
#include "equation.h" // the reference's header, found via -I/root/reference/src
#include "phys.h"

template <> const double EulerEquations<2>::gas_gamma = 1.4; // src/equation.cc:33

namespace
{
   // What dealii::Table<2,double>::operator[] hands to the flux templates: a row accessor
   // whose const operator[] yields a writable reference (see the reference's own remark,
   // equation.h:896-938).
   struct Row
   {
      typedef double value_type;
      double *p;
      explicit Row (const double *q) : p (const_cast<double *> (q)) {}
      double &operator[] (unsigned int i) const { return p[i]; }
      double *begin () const { return p; }
   };

   dealii::Tensor<1, 2> tensor (const double n[2])
   {
      dealii::Tensor<1, 2> t;
      t[0] = n[0];
      t[1] = n[1];
      return t;
   }

   dealii::Vector<double> vec4 (const double a[4])
   {
      dealii::Vector<double> v (4);
      for (int c = 0; c < 4; ++c) v[c] = a[c];
      return v;
   }

   typedef EulerEquations<2> EE;
}

extern "C" {

const char *phys_impl_name (void) { return "reference:src/equation.h"; }

void phys_numerical_flux (int flux_type, const double n[2], const double Wp[4], const double Wm[4],
                          const double Ap[4], const double Am[4], double out[4])
{
   const dealii::Tensor<1, 2> normal = tensor (n);
   const Row wp (Wp), wm (Wm);
   double (&f)[4] = *reinterpret_cast<double (*)[4]> (out);
   switch (flux_type) // same dispatch as claw.h:283-324
   {
      case PHYS_FLUX_LXF: EE::lxf_flux (normal, wp, wm, vec4 (Ap), vec4 (Am), f); break;
      case PHYS_FLUX_SW: EE::steger_warming_flux (normal, wp, wm, f); break;
      case PHYS_FLUX_KFVS: EE::kfvs_flux (normal, wp, wm, f); break;
      case PHYS_FLUX_ROE: EE::roe_flux (normal, wp, wm, f); break;
      case PHYS_FLUX_HLLC: EE::hllc_flux (normal, wp, wm, f); break;
      case PHYS_FLUX_KEP: phys_kep_flux_restated (n, Wp, Wm, Ap, Am, out); break; // src_mpi only, see phys_kep_restated.c
      default: assert (false);
   }
}

void phys_flux_matrix (const double W[4], double F[8])
{
   double flux[4][2];
   EE::compute_flux_matrix (Row (W), flux);
   for (int c = 0; c < 4; ++c)
      for (int d = 0; d < 2; ++d) F[2 * c + d] = flux[c][d];
}

void phys_forcing (const double W[4], double G[4])
{
   double (&g)[4] = *reinterpret_cast<double (*)[4]> (G);
   EE::compute_forcing_vector (Row (W), g);
}

void phys_wminus (int kind, const double n[2], const double Wp[4], const double g[4], double Wm[4])
{
   const EE::BoundaryKind k = static_cast<EE::BoundaryKind> (kind);
   EE::compute_Wminus (k, tensor (n), Row (Wp), vec4 (g), Row (Wm));
}

void phys_eigen (const double W[4], double Rx[16], double Lx[16], double Ry[16], double Ly[16])
{
   double rx[4][4], lx[4][4], ry[4][4], ly[4][4];
   EE::compute_eigen_matrix (vec4 (W), rx, lx, ry, ly);
   for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j)
      {
         Rx[4 * i + j] = rx[i][j];
         Lx[4 * i + j] = lx[i][j];
         Ry[4 * i + j] = ry[i][j];
         Ly[4 * i + j] = ly[i][j];
      }
}

void phys_to_char (const double L[16], double W[4])
{
   double l[4][4];
   for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) l[i][j] = L[4 * i + j];
   dealii::Vector<double> w = vec4 (W);
   EE::transform_to_char (l, w);
   for (int c = 0; c < 4; ++c) W[c] = w[c];
}

void phys_to_con (const double R[16], double W[4])
{
   double r[4][4];
   for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) r[i][j] = R[4 * i + j];
   dealii::Vector<double> w = vec4 (W);
   EE::transform_to_con (r, w);
   for (int c = 0; c < 4; ++c) W[c] = w[c];
}

double phys_pressure (const double W[4]) { return EE::compute_pressure<double> (Row (W)); }
double phys_sound_speed (const double W[4]) { return EE::sound_speed (vec4 (W)); }
double phys_max_eigenvalue (const double W[4]) { return EE::max_eigenvalue (vec4 (W)); }
}
