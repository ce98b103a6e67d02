// TEST INFRASTRUCTURE ONLY (oracle/). Built only where /root/reference exists, into oracle/_ref/.
//
// The reference's MPI-tree physics header, /root/reference/src_mpi/equation.h, included
// UNMODIFIED (through -I/root/reference/src_mpi and the deal.II stub) so that its kep_flux is the
// reference's own object code.  It is a library of its own with hidden visibility: the class is
// called EulerEquations<dim> in both trees, and the two must never meet in one link.
#include "equation.h" // src_mpi

template <> const double EulerEquations<2>::gas_gamma = 1.4; // src_mpi/equation.cc

namespace
{
   struct Row
   {
      typedef double value_type;
      double *p;
      explicit Row (const double *q) : p (const_cast<double *> (q)) {}
      double &operator[] (unsigned int i) const { return p[i]; }
   };
   dealii::Vector<double> vec4 (const double a[4])
   {
      dealii::Vector<double> v (4);
      for (int c = 0; c < 4; ++c) v[c] = a[c];
      return v;
   }
}

extern "C" __attribute__ ((visibility ("default"))) const char *phys_mpi_impl_name (void) { return "reference:src_mpi/equation.h"; }

// src_mpi/claw.h:363-370
extern "C" __attribute__ ((visibility ("default"))) void phys_mpi_kep_flux (const double n[2], const double Wl[4], const double Wr[4],
                                                                          const double Al[4], const double Ar[4], double out[4])
{
   dealii::Tensor<1, 2> normal;
   normal[0] = n[0];
   normal[1] = n[1];
   double (&f)[4] = *reinterpret_cast<double (*)[4]> (out);
   EulerEquations<2>::kep_flux (normal, Row (Wl), Row (Wr), vec4 (Al), vec4 (Ar), f);
}

// src_mpi/equation.h:299-335, the streamline-direction eigenvector matrices the minmax limiter
// projects with (src_mpi/limiter.cc:450); row-major 4x4
extern "C" __attribute__ ((visibility ("default"))) void phys_mpi_eigen_stream (const double W[4], double R[16], double L[16])
{
   double (&r)[4][4] = *reinterpret_cast<double (*)[4][4]> (R);
   double (&l)[4][4] = *reinterpret_cast<double (*)[4][4]> (L);
   EulerEquations<2>::compute_eigen_matrix (vec4 (W), r, l);
}

// src_mpi/equation.h:1189-1202, forcing vector of an external force f (src_mpi/assemble_explicit.cc:56-58, 84)
extern "C" __attribute__ ((visibility ("default"))) void phys_mpi_ext_forcing (const double W[4], const double f[2], double G[4])
{
   dealii::Vector<double> ef (2);
   ef[0] = f[0];
   ef[1] = f[1];
   double (&g)[4] = *reinterpret_cast<double (*)[4]> (G);
   EulerEquations<2>::compute_forcing_vector (Row (W), ef, g);
}
