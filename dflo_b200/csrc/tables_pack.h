// Flat table images in the exact order the kernels index them (kernels.cuh: StageKernel /
// LimiterKernel table accessors).  Built once on the host, copied to the device, and copied by
// every block into shared memory at kernel start.
#pragma once

#include "tables.h"

#include <algorithm>
#include <cmath>
#include <vector>

namespace dflo
{
   // Qk: dw[N1*N1] e0[N1] e1[N1] gw[N1]
   // Pk: phi[NQ*NS] dphix[NQ*NS] dphiy[NQ*NS] phiface[4*N1*NS] gw[N1]
   inline std::vector<double> pack_stage_tables (const FeTables &t)
   {
      std::vector<double> o;
      const int n1 = t.n1;
      if (t.basis == BASIS_QK)
      {
         for (int ap = 0; ap < n1; ++ap)
            for (int a = 0; a < n1; ++a) o.push_back (t.dw[ap][a]);
         for (int s = 0; s < 2; ++s)
            for (int a = 0; a < n1; ++a) o.push_back (t.e[s][a]);
      }
      else
      {
         for (int q = 0; q < t.nq; ++q)
            for (int m = 0; m < t.ns; ++m) o.push_back (t.phi[q][m]);
         for (int q = 0; q < t.nq; ++q)
            for (int m = 0; m < t.ns; ++m) o.push_back (t.dphix[q][m]);
         for (int q = 0; q < t.nq; ++q)
            for (int m = 0; m < t.ns; ++m) o.push_back (t.dphiy[q][m]);
         for (int f = 0; f < 4; ++f)
            for (int q = 0; q < n1; ++q)
               for (int m = 0; m < t.ns; ++m) o.push_back (t.phiface[f][q][m]);
      }
      for (int a = 0; a < n1; ++a) o.push_back (t.gw[a]);
      if (t.basis == BASIS_QK) // read by the mapped stage kernel only; the others take exactly stage_table_size doubles
      {
         for (int a = 0; a < n1; ++a) o.push_back (t.gx[a]);
         // sub-face interpolation (hanging nodes): S[child][q][a] = l_a ((x_q + child) / 2), the 1-D Lagrange basis on the Gauss
         // nodes at the Gauss points of one half of a face
         for (int child = 0; child < 2; ++child)
            for (int q = 0; q < n1; ++q)
               for (int a = 0; a < n1; ++a)
               {
                  const double x = 0.5 * (t.gx[q] + child);
                  double l = 1.0;
                  for (int m = 0; m < n1; ++m)
                     if (m != a) l *= (x - t.gx[m]) / (t.gx[a] - t.gx[m]);
                  o.push_back (l);
               }
      }
      return o;
   }

   // Qk: gw[N1] gx[N1] gdiff[N1] gl_interp[NGLL*N1];  Pk: phipos[2][NPOS][NS] | cmax[NS]
   inline std::vector<double> pack_limiter_tables (const FeTables &t)
   {
      std::vector<double> o;
      const int n1 = t.n1;
      if (t.basis == BASIS_QK)
      {
         for (int a = 0; a < n1; ++a) o.push_back (t.gw[a]);
         for (int a = 0; a < n1; ++a) o.push_back (t.gx[a]);
         for (int a = 0; a < n1; ++a) o.push_back (t.gdiff[a]);
         for (int j = 0; j < t.ngll; ++j)
            for (int a = 0; a < n1; ++a) o.push_back (t.gl_interp[j][a]);
      }
      else
      {
         for (int s = 0; s < 2; ++s)
            for (int p = 0; p < t.npos; ++p)
               for (int m = 0; m < t.ns; ++m) o.push_back (t.phipos[s][p][m]);
         // behind the limiter_table_size doubles the block form copies: max |basis function| over the positivity points, per
         // mode (the rigorous bound of the cell limiter's fast path)
         for (int m = 0; m < t.ns; ++m)
         {
            double cm = 0.0;
            for (int s = 0; s < 2; ++s)
               for (int p = 0; p < t.npos; ++p) cm = std::max (cm, std::fabs (t.phipos[s][p][m]));
            o.push_back (cm);
         }
      }
      if (o.empty ()) o.push_back (0.0);
      return o;
   }
}
