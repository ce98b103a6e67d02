// TEST INFRASTRUCTURE ONLY (oracle/): minimal stand-ins for the deal.II names that
// /root/reference/src/equation.h mentions, so that the reference's own flux / boundary /
// eigenvector arithmetic (equation.h:67-1033) compiles UNMODIFIED with plain g++.
// Nothing here restates reference code; it only supplies the container types the
// reference templates are instantiated with. deal.II itself is not installed in this image.
#ifndef DFLO_ORACLE_DEALII_STUB_CORE_H
#define DFLO_ORACLE_DEALII_STUB_CORE_H

#include <cassert>
#include <cmath>
#include <cstddef>
#include <string>
#include <vector>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#define Assert(cond, exc) assert(cond)

namespace dealii
{
   struct ExcNotImplemented {};

   template <int rank, int dim>
   struct Tensor
   {
      double v[dim > 0 ? dim : 1];
      Tensor () { for (int d = 0; d < dim; ++d) v[d] = 0.0; }
      double &operator[] (unsigned int i) { return v[i]; }
      const double &operator[] (unsigned int i) const { return v[i]; }
   };

   template <int dim>
   struct Point : public Tensor<1, dim> {};

   // dealii::Vector<double>: operator[], operator(), begin(), value_type, "= scalar".
   template <typename T>
   struct Vector
   {
      typedef T value_type;
      std::vector<T> data;
      Vector () {}
      explicit Vector (unsigned int n) : data (n, T (0)) {}
      unsigned int size () const { return data.size (); }
      T &operator[] (unsigned int i) { return data[i]; }
      const T &operator[] (unsigned int i) const { return data[i]; }
      T &operator() (unsigned int i) { return data[i]; }
      const T &operator() (unsigned int i) const { return data[i]; }
      T *begin () { return data.data (); }
      const T *begin () const { return data.data (); }
      Vector &operator= (const T s) { for (auto &x : data) x = s; return *this; }
   };

   namespace DataComponentInterpretation
   {
      enum DataComponentInterpretation { component_is_scalar, component_is_part_of_vector };
   }

   enum UpdateFlags { update_default = 0, update_values = 1, update_gradients = 2 };

   template <int dim, int spacedim = dim> class DoFHandler;
   template <int dim, int spacedim = dim> class Mapping;
   template <int dim> class QMidpoint;
   template <int dim, int spacedim = dim> class FEValues;
   template <int dim> struct DataPostprocessor { virtual ~DataPostprocessor () {} };
}

#endif
