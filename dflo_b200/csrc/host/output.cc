// VTU writers, see output.h.  ASCII data arrays (%.10g); file structure follows what deal.II's DataOut writes
// for a discontinuous field: no vertex is shared between cells.
#include "output.h"

#include <cmath>
#include <cstdio>
#include <vector>

namespace dflo
{
   namespace
   {
      const double gas_gamma = 1.4; // src/equation.cc:33

      void vtu_head (FILE *fp, int n_points, int n_cells, double time, unsigned int cycle, bool with_flags)
      {
         std::fprintf (fp, "<?xml version=\"1.0\"?>\n");
         if (with_flags) std::fprintf (fp, "<!-- time %.16g cycle %u -->\n", time, cycle); // DataOutBase::VtkFlags (time, cycle)
         std::fprintf (fp, "<VTKFile type=\"UnstructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\">\n<UnstructuredGrid>\n"
                           "<Piece NumberOfPoints=\"%d\" NumberOfCells=\"%d\">\n<Points>\n"
                           "<DataArray type=\"Float64\" NumberOfComponents=\"3\" format=\"ascii\">\n", n_points, n_cells);
      }

      // nsub x nsub quads on the (nsub+1)^2 lexicographic points of each of the nc cells written
      void vtu_cells (FILE *fp, int nc, int nsub)
      {
         const int np1 = nsub + 1, npc = np1 * np1, n_out = nc * nsub * nsub;
         std::fprintf (fp, "</DataArray>\n</Points>\n<Cells>\n<DataArray type=\"Int32\" Name=\"connectivity\" format=\"ascii\">\n");
         for (int cell = 0; cell < nc; ++cell)
            for (int j = 0; j < nsub; ++j)
               for (int i = 0; i < nsub; ++i)
               {
                  const int p0 = cell * npc + i + np1 * j;
                  std::fprintf (fp, "%d %d %d %d\n", p0, p0 + 1, p0 + 1 + np1, p0 + np1);
               }
         std::fprintf (fp, "</DataArray>\n<DataArray type=\"Int32\" Name=\"offsets\" format=\"ascii\">\n");
         for (int i = 1; i <= n_out; ++i) std::fprintf (fp, "%d\n", 4 * i);
         std::fprintf (fp, "</DataArray>\n<DataArray type=\"UInt8\" Name=\"types\" format=\"ascii\">\n");
         for (int i = 0; i < n_out; ++i) std::fprintf (fp, "9\n");
         std::fprintf (fp, "</DataArray>\n</Cells>\n");
      }

      void vtu_points (FILE *fp, const FlatMesh &flat, int nsub, int c0, int c1)
      {
         for (int cell = c0; cell < c1; ++cell)
            for (int j = 0; j <= nsub; ++j)
               for (int i = 0; i <= nsub; ++i)
               {
                  double x = flat.origin[2 * cell] + (double) i / nsub * flat.size[2 * cell];
                  double y = flat.origin[2 * cell + 1] + (double) j / nsub * flat.size[2 * cell + 1];
                  if (!flat.cartesian) flat.map (cell, (double) i / nsub, (double) j / nsub, x, y); // build_patches with MappingQ1
                  std::fprintf (fp, "%.10g %.10g 0\n", x, y);
               }
      }
   }

   namespace
   {
      // values of the 4 conserved variables [point][4] and of |grad rho|^2 [point] at the (nsub+1)^2 patch
      // vertices of cells [cell_begin, cell_end)
      void eval_patches (const FeTables &tab, const FlatMesh &flat, const double *u, bool schlieren_plot, int cell_begin, int cell_end,
                         std::vector<double> &val, std::vector<double> &schl)
      {
         const int nc = cell_end - cell_begin, ns = tab.ns, D = tab.D;
         const int nsub = tab.k > 0 ? tab.k : 1, np1 = nsub + 1, npc = np1 * np1;
         u += (size_t) cell_begin * D; // rows below are relative to the first cell written
         // basis values and unit-cell gradients at the equispaced patch vertices
         std::vector<double> phi ((size_t) npc * ns), dpx ((size_t) npc * ns), dpy ((size_t) npc * ns);
         for (int j = 0; j < np1; ++j)
            for (int i = 0; i < np1; ++i)
            {
               const size_t v = (size_t) (i + np1 * j) * ns;
               eval_basis (tab, (double) i / nsub, (double) j / nsub, &phi[v], &dpx[v], &dpy[v]);
            }
         val.assign ((size_t) nc * npc * 4, 0.0);
         schl.assign (schlieren_plot ? (size_t) nc * npc : 0, 0.0);
         for (int cell = 0; cell < nc; ++cell)
            for (int v = 0; v < npc; ++v)
            {
               const double *pv = &phi[(size_t) v * ns];
               for (int c = 0; c < 4; ++c)
               {
                  const double *uc = &u[(size_t) cell * D + c * ns];
                  double s = 0.0;
                  for (int m = 0; m < ns; ++m) s += pv[m] * uc[m];
                  val[((size_t) cell * npc + v) * 4 + c] = s;
               }
               if (schlieren_plot)
               {
                  // duh[density] . duh[density] in real coordinates (src/equation.cc:122-124)
                  const double *ur = &u[(size_t) cell * D + 2 * ns];
                  double gx = 0.0, gy = 0.0;
                  for (int m = 0; m < ns; ++m)
                  {
                     gx += dpx[(size_t) v * ns + m] * ur[m];
                     gy += dpy[(size_t) v * ns + m] * ur[m];
                  }
                  if (flat.cartesian)
                  {
                     gx /= flat.size[2 * (cell_begin + cell)];
                     gy /= flat.size[2 * (cell_begin + cell) + 1];
                  }
                  else // J^-T times the unit-cell gradient at the patch vertex
                  {
                     const int nsub1 = (int) std::lround (std::sqrt ((double) npc));
                     double J[4];
                     flat.jacobian (cell_begin + cell, (double) (v % nsub1) / (nsub1 - 1), (double) (v / nsub1) / (nsub1 - 1), J);
                     const double det = J[0] * J[3] - J[1] * J[2], gu = gx, gv = gy;
                     gx = (J[3] * gu - J[2] * gv) / det;
                     gy = (-J[1] * gu + J[0] * gv) / det;
                  }
                  schl[(size_t) cell * npc + v] = gx * gx + gy * gy;
               }
            }
      }

      // component_names (src/equation.h) then Postprocessor::get_names (src/equation.cc:130-145)
      const char *const output_names[8] = {"XMomentum", "YMomentum", "Density", "Energy", "XVelocity", "YVelocity", "Pressure", "schlieren_plot"};

      double output_value (int k, const double *w, const std::vector<double> &schl, size_t p)
      {
         if (k < 4) return w[k];
         if (k == 4) return w[0] / w[2];
         if (k == 5) return w[1] / w[2];
         if (k == 6) return (gas_gamma - 1.0) * (w[3] - 0.5 * (w[0] * w[0] + w[1] * w[1]) / w[2]);
         return schl[p];
      }
   }

   bool write_solution_vtu (const FeTables &tab, const FlatMesh &flat, const double *u, bool schlieren_plot, double time,
                            unsigned int cycle, const std::string &path, int cell_begin, int cell_end, int subdomain)
   {
      if (cell_end < 0) cell_end = flat.n_cells ();
      if (cell_begin < 0 || cell_begin > cell_end || cell_end > flat.n_cells ()) return false;
      FILE *fp = std::fopen (path.c_str (), "w");
      if (!fp) return false;
      const int nc = cell_end - cell_begin;
      const int nsub = tab.k > 0 ? tab.k : 1, npc = (nsub + 1) * (nsub + 1);
      std::vector<double> val, schl;
      eval_patches (tab, flat, u, schlieren_plot, cell_begin, cell_end, val, schl);
      vtu_head (fp, npc * nc, nc * nsub * nsub, time, cycle, true);
      vtu_points (fp, flat, nsub, cell_begin, cell_end);
      vtu_cells (fp, nc, nsub);
      std::fprintf (fp, "<PointData>\n");
      // DataOutBase::write_vtu writes the vector-valued ranges first -- the momentum (component_interpretation,
      // src/equation.h:48-59) and the velocity (Postprocessor, src/equation.cc:147-166), each as one 3-component array
      // named by its components joined with "__" -- and then the scalars in the order they were added
      for (int k0 : {0, 4})
      {
         std::fprintf (fp, "<DataArray type=\"Float64\" Name=\"%s__%s\" NumberOfComponents=\"3\" format=\"ascii\">\n", output_names[k0],
                       output_names[k0 + 1]);
         for (size_t p = 0; p < (size_t) nc * npc; ++p)
            std::fprintf (fp, "%.10g %.10g 0\n", output_value (k0, &val[p * 4], schl, p), output_value (k0 + 1, &val[p * 4], schl, p));
         std::fprintf (fp, "</DataArray>\n");
      }
      for (int k : {2, 3, 6, 7})
      {
         if (k == 7 && !schlieren_plot) continue;
         std::fprintf (fp, "<DataArray type=\"Float64\" Name=\"%s\" format=\"ascii\">\n", output_names[k]);
         for (size_t p = 0; p < (size_t) nc * npc; ++p) std::fprintf (fp, "%.10g\n", output_value (k, &val[p * 4], schl, p));
         std::fprintf (fp, "</DataArray>\n");
      }
      if (subdomain >= 0) // locally_owned_subdomain of every cell of the piece, src_mpi/output.cc:51-54
      {
         std::fprintf (fp, "<DataArray type=\"Float64\" Name=\"subdomain\" format=\"ascii\">\n");
         for (size_t p = 0; p < (size_t) nc * npc; ++p) std::fprintf (fp, "%d\n", subdomain);
         std::fprintf (fp, "</DataArray>\n");
      }
      std::fprintf (fp, "</PointData>\n</Piece>\n</UnstructuredGrid>\n</VTKFile>\n");
      return std::fclose (fp) == 0;
   }

   // "output: format = tecplot" (src/output.cc:51-52, 65-66): DataOut::write_tecplot -- ASCII FEBLOCK zone of
   // quadrilaterals: all x, all y, then every variable over the same patch vertices, then 1-based connectivity
   bool write_solution_tecplot (const FeTables &tab, const FlatMesh &flat, const double *u, bool schlieren_plot, double time,
                                const std::string &path)
   {
      FILE *fp = std::fopen (path.c_str (), "w");
      if (!fp) return false;
      const int nc = flat.n_cells ();
      const int nsub = tab.k > 0 ? tab.k : 1, np1 = nsub + 1, npc = np1 * np1;
      std::vector<double> val, schl;
      eval_patches (tab, flat, u, schlieren_plot, 0, nc, val, schl);
      const int n_arrays = schlieren_plot ? 8 : 7;
      std::fprintf (fp, "# This file was generated by dflo_b200 in the layout of the deal.II tecplot writer.\n#\n"
                        "# For a description of the Tecplot format see the Tecplot documentation.\n#\nVariables=\"x\", \"y\"");
      for (int k = 0; k < n_arrays; ++k) std::fprintf (fp, ", \"%s\"", output_names[k]);
      std::fprintf (fp, "\nzone t=\"time=%.10g\" f=feblock, n=%d, e=%d, et=quadrilateral\n", time, nc * npc, nc * nsub * nsub);
      for (int d = 0; d < 2; ++d)
      {
         for (int cell = 0; cell < nc; ++cell)
            for (int j = 0; j < np1; ++j)
               for (int i = 0; i < np1; ++i)
               {
                  double xy[2] = {flat.origin[2 * cell] + (double) i / nsub * flat.size[2 * cell], flat.origin[2 * cell + 1] + (double) j / nsub * flat.size[2 * cell + 1]};
                  if (!flat.cartesian) flat.map (cell, (double) i / nsub, (double) j / nsub, xy[0], xy[1]);
                  std::fprintf (fp, "%.10g\n", xy[d]);
               }
         std::fprintf (fp, "\n");
      }
      for (int k = 0; k < n_arrays; ++k)
      {
         for (size_t p = 0; p < (size_t) nc * npc; ++p) std::fprintf (fp, "%.10g\n", output_value (k, &val[p * 4], schl, p));
         std::fprintf (fp, "\n");
      }
      for (int cell = 0; cell < nc; ++cell)
         for (int j = 0; j < nsub; ++j)
            for (int i = 0; i < nsub; ++i)
            {
               const int p0 = cell * npc + i + np1 * j + 1;
               std::fprintf (fp, "%d %d %d %d\n", p0, p0 + 1, p0 + 1 + np1, p0 + np1);
            }
      return std::fclose (fp) == 0;
   }

   // shock.plt (src/output.cc:80-84)
   bool write_shock_tecplot (const FlatMesh &flat, const double *mu_shock, const double *shock_indicator, const std::string &path)
   {
      FILE *fp = std::fopen (path.c_str (), "w");
      if (!fp) return false;
      const int nc = flat.n_cells ();
      std::fprintf (fp, "# This file was generated by dflo_b200 in the layout of the deal.II tecplot writer.\n#\n"
                        "Variables=\"x\", \"y\", \"mu_shock\", \"shock_indicator\"\n"
                        "zone t=\"\" f=feblock, n=%d, e=%d, et=quadrilateral\n", 4 * nc, nc);
      for (int d = 0; d < 2; ++d)
      {
         for (int cell = 0; cell < nc; ++cell)
            for (int v = 0; v < 4; ++v) std::fprintf (fp, "%.10g\n", flat.vertices[8 * (size_t) cell + 2 * v + d]);
         std::fprintf (fp, "\n");
      }
      for (int i = 0; i < 4 * nc; ++i) std::fprintf (fp, "%.10g\n", mu_shock ? mu_shock[i / 4] : 0.0);
      std::fprintf (fp, "\n");
      for (int i = 0; i < 4 * nc; ++i) std::fprintf (fp, "%.10g\n", shock_indicator[i / 4]);
      std::fprintf (fp, "\n");
      for (int cell = 0; cell < nc; ++cell) std::fprintf (fp, "%d %d %d %d\n", 4 * cell + 1, 4 * cell + 2, 4 * cell + 4, 4 * cell + 3);
      return std::fclose (fp) == 0;
   }

   // cell loop of compute_angular_momentum, src/claw.cc:604-635: QGauss(k+1)^2, cross = x m_y - y m_x
   double angular_momentum (const FeTables &tab, const FlatMesh &flat, const double *u, int cell_begin, int cell_end)
   {
      if (cell_end < 0) cell_end = flat.n_cells ();
      const int n1 = tab.n1, nq = n1 * n1, ns = tab.ns, D = tab.D;
      double total = 0.0;
      for (int cell = cell_begin; cell < cell_end; ++cell)
      {
         const double x0 = flat.origin[2 * cell], y0 = flat.origin[2 * cell + 1], hx = flat.size[2 * cell], hy = flat.size[2 * cell + 1];
         const double *mx = &u[(size_t) cell * D], *my = mx + ns;
         for (int q = 0; q < nq; ++q)
         {
            double vx = 0.0, vy = 0.0;
            if (tab.basis == BASIS_QK) // collocated: the DoF is the value at the Gauss point
            {
               vx = mx[q];
               vy = my[q];
            }
            else
               for (int m = 0; m < ns; ++m)
               {
                  vx += tab.phi[q][m] * mx[m];
                  vy += tab.phi[q][m] * my[m];
               }
            double x = x0 + tab.gx[q % n1] * hx, y = y0 + tab.gx[q / n1] * hy, jxw = tab.gw[q % n1] * tab.gw[q / n1] * hx * hy;
            if (!flat.cartesian) // MappingQ1: mapped point, w det J
            {
               double J[4];
               flat.map (cell, tab.gx[q % n1], tab.gx[q / n1], x, y);
               flat.jacobian (cell, tab.gx[q % n1], tab.gx[q / n1], J);
               jxw = tab.gw[q % n1] * tab.gw[q / n1] * (J[0] * J[3] - J[1] * J[2]);
            }
            total += (x * vy - y * vx) * jxw;
         }
      }
      return total;
   }

   bool write_visit_record (const std::vector<std::vector<std::string>> &all_files, const std::string &path)
   {
      FILE *fp = std::fopen (path.c_str (), "w");
      if (!fp) return false;
      if (!all_files.empty ()) std::fprintf (fp, "!NBLOCKS %zu\n", all_files[0].size ());
      for (const auto &step : all_files)
         for (const auto &f : step) std::fprintf (fp, "%s\n", f.c_str ());
      return std::fclose (fp) == 0;
   }

   bool write_shock_vtu (const FlatMesh &flat, const double *mu_shock, const double *shock_indicator, const std::string &path)
   {
      FILE *fp = std::fopen (path.c_str (), "w");
      if (!fp) return false;
      const int nc = flat.n_cells ();
      vtu_head (fp, 4 * nc, nc, 0.0, 0, false);
      vtu_points (fp, flat, 1, 0, nc);
      vtu_cells (fp, nc, 1);
      // DataOut stores cell data as the same value on every vertex of the cell's patch and writes point data only
      std::fprintf (fp, "<PointData>\n<DataArray type=\"Float64\" Name=\"mu_shock\" format=\"ascii\">\n");
      for (int i = 0; i < 4 * nc; ++i) std::fprintf (fp, "%.10g\n", mu_shock ? mu_shock[i / 4] : 0.0);
      std::fprintf (fp, "</DataArray>\n<DataArray type=\"Float64\" Name=\"shock_indicator\" format=\"ascii\">\n");
      for (int i = 0; i < 4 * nc; ++i) std::fprintf (fp, "%.10g\n", shock_indicator[i / 4]);
      std::fprintf (fp, "</DataArray>\n</PointData>\n</Piece>\n</UnstructuredGrid>\n</VTKFile>\n");
      return std::fclose (fp) == 0;
   }
}
