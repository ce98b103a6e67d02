/* TEST INFRASTRUCTURE ONLY: deal.II stub for compiling the reference's src_mpi/equation.h
 * unmodified (declarations the header names but the point-wise physics never touches). */
#pragma once
#include "../../dealii_stub_core.h"
namespace dealii
{
   namespace LinearAlgebra
   {
      namespace distributed
      {
         template <typename T> class Vector;
      }
   }
   namespace DataPostprocessorInputs
   {
      template <int dim> struct Vector;
   }
}
