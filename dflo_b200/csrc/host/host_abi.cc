// C ABI of the host front end: meshes and expressions (include/dflo_host.h).
#include "../../../include/dflo_host.h"
#include "../expr.h"
#include "host_error.h"
#include "mesh.h"
#include "mesh_handle.h"
#include "output.h"
#include "../tables.h"

#include <cstring>
#include <string>
#include <vector>


namespace dflo
{
   std::string &host_error ()
   {
      static thread_local std::string e;
      return e;
   }
}

extern "C" {

const char *dflo_host_last_error (void) { return dflo::host_error ().c_str (); }

dflo_mesh *dflo_mesh_create (const char *kind, const double *a, int n)
{
   if (!kind)
   {
      dflo::host_error () = "mesh kind is null";
      return nullptr;
   }
   const std::string k = kind;
   dflo_mesh *m = new dflo_mesh;
   if (k == "rectangle" && n == 10)
   {
      const int ids[4] = {(int) a[6], (int) a[7], (int) a[8], (int) a[9]};
      m->pm = dflo::make_rectangle ((int) a[0], (int) a[1], a[2], a[3], a[4], a[5], ids);
   }
   else if (k == "rectangle_skew" && n == 12)
   {
      const int ids[4] = {(int) a[6], (int) a[7], (int) a[8], (int) a[9]};
      m->pm = dflo::make_skewed_rectangle ((int) a[0], (int) a[1], a[2], a[3], a[4], a[5], ids, a[10], (int) a[11]);
   }
   else if (k == "rectangle_refined" && (n == 14 || n == 15))
   {
      const int ids[4] = {(int) a[6], (int) a[7], (int) a[8], (int) a[9]};
      m->pm = dflo::make_refined_rectangle ((int) a[0], (int) a[1], a[2], a[3], a[4], a[5], ids, (int) a[10], (int) a[11], (int) a[12], (int) a[13],
                                            n == 15 ? (int) a[14] : 0);
   }
   else if (k == "compression_corner" && n == 3)
      m->pm = dflo::make_compression_corner ((int) a[0], (int) a[1], (int) a[2]);
   else if (k == "isentropic_vortex" && n == 1)
      m->pm = dflo::make_isentropic_vortex_grid ((int) a[0]);
   else if (k == "sod_tube" && n == 2)
      m->pm = dflo::make_sod_tube ((int) a[0], (int) a[1]);
   else if (k == "double_mach" && n == 1)
      m->pm = dflo::make_double_mach_grid ((int) a[0]);
   else if (k == "forward_step" && n == 1)
      m->pm = dflo::make_forward_step_grid (a[0]);
   else
   {
      dflo::host_error () = "unknown mesh kind or wrong argument count: " + k;
      delete m;
      return nullptr;
   }
   return m;
}

dflo_mesh *dflo_mesh_read_gmsh (const char *path)
{
   dflo_mesh *m = new dflo_mesh;
   std::string e;
   if (!path || !dflo::read_gmsh2 (path, m->pm, e))
   {
      dflo::host_error () = e;
      delete m;
      return nullptr;
   }
   return m;
}

int dflo_mesh_write_gmsh (const dflo_mesh *m, const char *path)
{
   return (m && path && dflo::write_gmsh2 (path, m->pm)) ? DFLO_OK : DFLO_E_INVALID;
}

void dflo_mesh_destroy (dflo_mesh *m) { delete m; }
int dflo_mesh_n_vertices (const dflo_mesh *m) { return m->pm.n_vertices (); }
int dflo_mesh_n_cells (const dflo_mesh *m) { return m->pm.n_cells (); }
int dflo_mesh_n_blines (const dflo_mesh *m) { return m->pm.n_blines (); }
const double *dflo_mesh_vertices (const dflo_mesh *m) { return m->pm.vertices.data (); }
const int *dflo_mesh_cells (const dflo_mesh *m) { return m->pm.cells.data (); }
const int *dflo_mesh_blines (const dflo_mesh *m) { return m->pm.blines.data (); }
const int *dflo_mesh_bline_ids (const dflo_mesh *m) { return m->pm.bline_id.data (); }

int dflo_mesh_flatten (dflo_mesh *m, const int bc_kind[DFLO_MAX_BOUNDARIES], const int periodic_pair[DFLO_MAX_BOUNDARIES])
{
   if (!m || !bc_kind || !periodic_pair) return DFLO_E_INVALID;
   std::string e;
   if (!dflo::flatten (m->pm, bc_kind, periodic_pair, m->flat, e))
   {
      dflo::host_error () = e;
      return DFLO_E_INVALID;
   }
   m->view = m->flat.view ();
   m->flattened = true;
   return DFLO_OK;
}

const dflo_flat_mesh *dflo_mesh_flat (const dflo_mesh *m) { return (m && m->flattened) ? &m->view : nullptr; }

int dflo_expr_eval (const char *expr, int n, const double *x, const double *y, double t, double *out)
{
   dflo::ExprCompiler cc;
   std::vector<dflo::ExprInstr> code;
   std::string e;
   if (!expr || !cc.compile (expr, code, e))
   {
      dflo::host_error () = e;
      return DFLO_E_EXPR;
   }
   for (int i = 0; i < n; ++i) out[i] = dflo::expr_eval (code.data (), (int) code.size (), x[i], y[i], t);
   return DFLO_OK;
}

int dflo_host_write_solution_piece_vtu (const dflo_mesh *m, int basis, int degree, const double *u, size_t n, int schlieren_plot,
                                        double time, unsigned int cycle, int cell_begin, int cell_end, int subdomain, const char *path)
{
   dflo::FeTables tab;
   if (!m || !m->flattened || !u || !path || !dflo::build_tables (basis, degree, tab) || n != (size_t) m->flat.n_cells () * tab.D
       || cell_begin < 0 || cell_end > m->flat.n_cells () || (cell_end >= 0 && cell_begin > cell_end))
   {
      dflo::host_error () = "write_solution_vtu: mesh not flattened, unsupported element, wrong vector length or bad cell range";
      return DFLO_E_INVALID;
   }
   if (!dflo::write_solution_vtu (tab, m->flat, u, schlieren_plot != 0, time, cycle, path, cell_begin, cell_end, subdomain))
   {
      dflo::host_error () = std::string ("cannot write ") + path;
      return DFLO_E_INVALID;
   }
   return DFLO_OK;
}

int dflo_host_write_solution_vtu (const dflo_mesh *m, int basis, int degree, const double *u, size_t n, int schlieren_plot,
                                  double time, unsigned int cycle, const char *path)
{
   return dflo_host_write_solution_piece_vtu (m, basis, degree, u, n, schlieren_plot, time, cycle, 0, -1, -1, path);
}

int dflo_host_write_solution_tecplot (const dflo_mesh *m, int basis, int degree, const double *u, size_t n, int schlieren_plot,
                                      double time, const char *path)
{
   dflo::FeTables tab;
   if (!m || !m->flattened || !u || !path || !dflo::build_tables (basis, degree, tab) || n != (size_t) m->flat.n_cells () * tab.D)
   {
      dflo::host_error () = "write_solution_tecplot: mesh not flattened, unsupported element or wrong vector length";
      return DFLO_E_INVALID;
   }
   if (!dflo::write_solution_tecplot (tab, m->flat, u, schlieren_plot != 0, time, path))
   {
      dflo::host_error () = std::string ("cannot write ") + path;
      return DFLO_E_INVALID;
   }
   return DFLO_OK;
}

int dflo_host_angular_momentum (const dflo_mesh *m, int basis, int degree, const double *u, size_t n, double *value)
{
   dflo::FeTables tab;
   if (!m || !m->flattened || !u || !value || !dflo::build_tables (basis, degree, tab) || n != (size_t) m->flat.n_cells () * tab.D)
   {
      dflo::host_error () = "angular_momentum: mesh not flattened, unsupported element or wrong vector length";
      return DFLO_E_INVALID;
   }
   *value = dflo::angular_momentum (tab, m->flat, u);
   return DFLO_OK;
}

int dflo_host_write_shock_vtu (const dflo_mesh *m, const double *mu_shock, const double *shock_indicator, const char *path)
{
   if (!m || !m->flattened || !shock_indicator || !path || !dflo::write_shock_vtu (m->flat, mu_shock, shock_indicator, path))
   {
      dflo::host_error () = "write_shock_vtu: mesh not flattened or file not writable";
      return DFLO_E_INVALID;
   }
   return DFLO_OK;
}
}
