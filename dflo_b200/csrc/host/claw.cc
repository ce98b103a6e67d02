#include "claw.h"

#include "../../../include/dflo_host.h"
#include "../expr.h"
#include "host_error.h"
#include "mesh_handle.h"
#include "output.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <sstream>
#include <sys/stat.h>

namespace dflo
{
   namespace
   {
      std::string dirname_of (const std::string &p)
      {
         const size_t s = p.find_last_of ('/');
         return s == std::string::npos ? std::string (".") : p.substr (0, s);
      }
   }

   // src/claw.cc:70-160 (constructor + read_parameters) and the mesh input of run(), :956-967
   ConservationLaw::ConservationLaw (const std::string &input_filename, const std::string &mesh_override,
                                     const std::string &overrides, int compat_)
      : compat (compat_)
   {
      std::string e;
      if (!parameters.parse_file (input_filename, e) || (!overrides.empty () && !parameters.parse_text (overrides, e))
          || !parameters.finish (e) || !parameters.to_engine_params (compat, engine_params, e))
      {
         error = e;
         return;
      }
      if (!build_tables (engine_params.basis, engine_params.degree, tab))
      {
         error = "degree out of range (Qk 0..4, Pk 0..3)";
         return;
      }
      if (!mesh_override.empty ())
      {
         std::istringstream is (mesh_override);
         std::string kind;
         is >> kind;
         std::vector<double> a;
         double v;
         while (is >> v) a.push_back (v);
         dflo_mesh *m = dflo_mesh_create (kind.c_str (), a.data (), (int) a.size ());
         if (!m)
         {
            error = dflo_host_last_error ();
            return;
         }
         const int nv = dflo_mesh_n_vertices (m), nc = dflo_mesh_n_cells (m), nb = dflo_mesh_n_blines (m);
         pm.vertices.assign (dflo_mesh_vertices (m), dflo_mesh_vertices (m) + 2 * nv);
         pm.cells.assign (dflo_mesh_cells (m), dflo_mesh_cells (m) + 4 * nc);
         pm.blines.assign (dflo_mesh_blines (m), dflo_mesh_blines (m) + 2 * nb);
         pm.bline_id.assign (dflo_mesh_bline_ids (m), dflo_mesh_bline_ids (m) + nb);
         dflo_mesh_destroy (m);
      }
      else
      {
         if (parameters.mesh_type != "gmsh")
         {
            error = "only 'mesh type = gmsh' (format 2) can be read";
            return;
         }
         std::string path = parameters.mesh_filename;
         if (!path.empty () && path[0] != '/') path = dirname_of (input_filename) + "/" + path;
         if (!read_gmsh2 (path, pm, e))
         {
            error = e;
            return;
         }
      }
      if (!flatten (pm, engine_params.bc_kind, parameters.periodic_pair, flat, e))
      {
         error = e;
         return;
      }
      flat_view = flat.view ();
   }

   ConservationLaw::~ConservationLaw ()
   {
      if (ctx) dflo_b200_destroy (ctx);
   }

   // IC functions of src/ic.cc:12-97 (src_mpi/ic.cc:44-61 for the advected vortex) or the
   // FunctionParser expressions of subsection "initial condition"
   void ConservationLaw::initial_value (double x, double y, double w[4]) const
   {
      const double gamma = 1.4;
      const std::string &f = parameters.ic_function;
      if (f == "isenvort")
      {
         const double beta = 5.0, x0 = 0.0, y0 = 0.0;
         const double a1 = 0.5 * beta / M_PI, a2 = (gamma - 1.0) * a1 * a1 / 2.0;
         const double r2 = (x - x0) * (x - x0) + (y - y0) * (y - y0);
         const double rho = std::pow (1.0 - a2 * std::exp (1.0 - r2), 1.0 / (gamma - 1.0));
         double vex = -a1 * (y - y0) * std::exp (0.5 * (1.0 - r2));
         double vey = +a1 * (x - x0) * std::exp (0.5 * (1.0 - r2));
         double pre = std::pow (rho, gamma);
         if (compat == DFLO_COMPAT_MPI) // IsentropicVortex(0.5, 0.0, 5.0, 0.0, 0.0), src_mpi/ic.cc:144
         {
            vex += 0.5;
            pre /= gamma;
         }
         w[0] = rho * vex;
         w[1] = rho * vey;
         w[2] = rho;
         w[3] = pre / (gamma - 1.0) + 0.5 * rho * (vex * vex + vey * vey);
      }
      else if (f == "rt")
      {
         const double Lx = 0.5, Ly = 1.5, A = 0.01, P0 = 2.5;
         const double rho = y < 0.0 ? 1.0 : 2.0;
         const double vel = A * (1.0 + std::cos (2.0 * M_PI * x / Lx)) / 2.0 * (1.0 + std::cos (2.0 * M_PI * y / Ly)) / 2.0;
         const double pressure = P0 - parameters.gravity * rho * y;
         w[0] = 0.0;
         w[1] = rho * vel;
         w[2] = rho;
         w[3] = pressure / (gamma - 1.0) + 0.5 * rho * vel * vel;
      }
      else if (f == "vortsys")
      {
         const double beta = 5.0, Rc = 4.0;
         const double a1 = 0.5 * beta / M_PI, a2 = (gamma - 1.0) * a1 * a1 / 2.0;
         const double xs[3] = {0.0, Rc * std::cos (30.0 * M_PI / 180.0), -Rc * std::cos (30.0 * M_PI / 180.0)};
         const double ys[3] = {-Rc, Rc * std::sin (30.0 * M_PI / 180.0), Rc * std::sin (30.0 * M_PI / 180.0)};
         double rho = 0, vex = 0, vey = 0;
         for (int i = 0; i < 3; ++i)
         {
            const double r2 = (x - xs[i]) * (x - xs[i]) + (y - ys[i]) * (y - ys[i]);
            rho += std::pow (1.0 - a2 * std::exp (1.0 - r2), 1.0 / (gamma - 1.0));
            vex += -a1 * (y - ys[i]) * std::exp (0.5 * (1.0 - r2));
            vey += +a1 * (x - xs[i]) * std::exp (0.5 * (1.0 - r2));
         }
         rho -= 2.0;
         vex /= 3.0;
         vey /= 3.0;
         double pre = std::pow (rho, gamma);
         if (std::fabs (x) < 0.1 && std::fabs (y) < 0.1) pre = 50.0;
         w[0] = rho * vex;
         w[1] = rho * vey;
         w[2] = rho;
         w[3] = pre / (gamma - 1.0) + 0.5 * rho * (vex * vex + vey * vey);
      }
   }

   // set_initial_condition: Qk interpolates at the support (= Gauss) points, Pk projects with
   // QGauss(k+1) and the diagonal mass matrix (src/ic.cc:104-182)
   int ConservationLaw::set_initial_condition (std::vector<double> &u, std::string &err) const
   {
      const int nc = flat.n_cells (), n1 = tab.n1, ns = tab.ns, D = tab.D, nq = tab.nq;
      u.assign ((size_t) nc * D, 0.0);
      std::vector<ExprInstr> code[4];
      const bool use_expr = parameters.ic_function == "none";
      if (use_expr)
      {
         ExprCompiler cc;
         std::string e;
         // FunctionParser::initialize throws on a syntax error (src/parameters.cc:524-526): report it
         // instead of running on a component that silently evaluates to zero
         for (int c = 0; c < 4; ++c)
            if (!cc.compile (parameters.ic_expr[c], code[c], e))
            {
               err = "initial condition: w_" + std::to_string (c) + " value: " + e;
               return DFLO_E_EXPR;
            }
      }
      std::vector<double> f ((size_t) nq * 4);
      for (int cell = 0; cell < nc; ++cell)
      {
         const double x0 = flat.origin[2 * cell], y0 = flat.origin[2 * cell + 1];
         const double hx = flat.size[2 * cell], hy = flat.size[2 * cell + 1];
         for (int b = 0; b < n1; ++b)
            for (int a = 0; a < n1; ++a)
            {
               double x = x0 + tab.gx[a] * hx, y = y0 + tab.gx[b] * hy;
               if (!flat.cartesian) flat.map (cell, tab.gx[a], tab.gx[b], x, y); // mapping = q1: the mapped support point
               double w[4];
               if (use_expr)
                  for (int c = 0; c < 4; ++c) w[c] = expr_eval (code[c].data (), (int) code[c].size (), x, y, 0.0);
               else
                  initial_value (x, y, w);
               for (int c = 0; c < 4; ++c) f[(size_t) (a + n1 * b) * 4 + c] = w[c];
            }
         double *uc = &u[(size_t) cell * D];
         if (tab.basis == BASIS_QK)
         {
            for (int c = 0; c < 4; ++c)
               for (int q = 0; q < nq; ++q) uc[c * ns + q] = f[(size_t) q * 4 + c];
         }
         else
         {
            for (int c = 0; c < 4; ++c)
               for (int m = 0; m < ns; ++m)
               {
                  double r = 0.0;
                  for (int q = 0; q < nq; ++q) r += f[(size_t) q * 4 + c] * tab.phi[q][m] * (tab.gw[q % n1] * tab.gw[q / n1] * hx * hy);
                  uc[c * ns + m] = r * (1.0 / (hx * hy));
               }
         }
      }
      return DFLO_OK;
   }

   // setup_system (src/claw.cc:270-386) + the start of run() (:981-1003)
   int ConservationLaw::setup_system (int device, int rank_, int world_, const void *nccl_unique_id)
   {
      if (!ok ()) return DFLO_E_INVALID;
      if (ctx) dflo_b200_destroy (ctx);
      ctx = nullptr;
      rank = world_ > 1 ? rank_ : 0;
      world = world_ > 1 ? world_ : 1;
      int rc = world > 1 ? dflo_b200_create_sharded (&flat_view, &engine_params, device, rank, world, nccl_unique_id, &ctx)
                         : dflo_b200_create (&flat_view, &engine_params, device, &ctx);
      if (rc)
      {
         error = dflo_b200_last_error (nullptr);
         return rc;
      }
      // boundary expressions g(x,y,t) of every boundary that uses its values
      for (int b = 0; b < DFLO_MAX_BOUNDARIES; ++b)
      {
         const int kind = engine_params.bc_kind[b];
         if (kind == DFLO_BC_INFLOW || kind == DFLO_BC_FARFIELD || kind == DFLO_BC_PRESSURE)
            for (int c = 0; c < 4; ++c)
            {
               rc = dflo_b200_set_boundary_expression (ctx, b, c, parameters.boundary_expr[b][c].c_str ());
               if (rc)
               {
                  error = dflo_b200_last_error (ctx);
                  return rc;
               }
            }
      }
      // MPI tree: the forcing term is gravity * (rho f, m.f) with the external force of the deck
      // (src_mpi/assemble_explicit.cc:56-58, 84); src/ keeps its hard-wired f = (0,-1)
      if (compat == DFLO_COMPAT_MPI && parameters.gravity != 0.0)
      {
         rc = dflo_b200_set_external_force (ctx, parameters.external_force[0].c_str (), parameters.external_force[1].c_str ());
         if (rc)
         {
            error = dflo_b200_last_error (ctx);
            return rc;
         }
      }
      std::vector<double> u;
      rc = set_initial_condition (u, error);
      if (rc) return rc;
      rc = dflo_b200_set_solution (ctx, u.data (), nullptr, u.size ());   // also cell averages (:997)
      if (!rc) rc = dflo_b200_limit_initial_condition (ctx);              // :1000-1002
      if (rc) error = dflo_b200_last_error (ctx);
      elapsed_time = 0.0;
      time_iter = 0;
      next_output_time = parameters.output_time_step; // src/claw.cc:1016-1017
      next_output_iter = parameters.output_iter_step;
      return rc;
   }

   // compute_time_step, src/claw.cc:444-478
   int ConservationLaw::compute_time_step ()
   {
      if (parameters.cfl <= 0.0) // time step given in the input file
      {
         global_dt = parameters.time_step;
         if (elapsed_time + global_dt > parameters.final_time) global_dt = parameters.final_time - elapsed_time;
         return DFLO_OK;
      }
      return dflo_b200_compute_dt (ctx, elapsed_time, parameters.final_time, &global_dt);
   }

   // iterate_explicit, src/claw.cc:725-772
   int ConservationLaw::iterate_explicit (double &res_norm0, double &res_norm, bool want_norm)
   {
      const int n_rk = dflo_b200_n_rk (ctx);
      for (int rk = 0; rk < n_rk; ++rk)
      {
         // bc time (src/claw.cc:736-745; src_mpi always uses elapsed_time, src_mpi/claw.cc:769-773)
         const double bc_time = (rk == 0 || compat == DFLO_COMPAT_MPI) ? elapsed_time : elapsed_time + global_dt;
         const int rc = dflo_b200_rk_stage (ctx, rk, bc_time, global_dt, want_norm ? &res_norm : nullptr);
         if (rc) return rc;
         if (rk == 0) res_norm0 = res_norm;
         if (want_norm) std::printf ("   %-16.3e %04d        %-5.2e\n", res_norm, 0, 0.0);
      }
      return DFLO_OK;
   }

   // run, src/claw.cc:1026-1110
   int ConservationLaw::run (int max_steps, bool verbose)
   {
      if (!ctx) return DFLO_E_INVALID;
      int steps = 0;
      if (output_enabled && time_iter == 0 && output_file_number == 0) // initial solution, src/claw.cc:1010
      {
         const int rc = output_results (output_dir);
         if (rc) return rc;
      }
      while (elapsed_time < parameters.final_time && (max_steps < 0 || steps < max_steps))
      {
         int rc;
         if (verbose)
         {
            rc = compute_time_step ();
            if (rc) return rc;
            std::printf ("\nIt=%d, T=%g, dt=%g, cfl=%g\n   Number of active cells:       %d\n   Number of degrees of freedom: %d\n\n",
                         time_iter + 1, elapsed_time + global_dt, global_dt, parameters.cfl, flat.n_cells (), n_dofs ());
            double r0 = 1.0, r = 1.0;
            rc = iterate_explicit (r0, r, true);
            if (rc) return rc;
            elapsed_time += global_dt;
            rc = dflo_b200_commit_step (ctx); // old_solution = current_solution, :1110
            if (!rc) rc = dflo_b200_poll_error (ctx);
         }
         else if (parameters.cfl > 0.0)
         {
            // whole step on the device: compute_time_step + iterate_explicit + commit
            rc = dflo_b200_advance (ctx, 1, parameters.final_time, &elapsed_time, &global_dt);
         }
         else
         {
            rc = compute_time_step ();
            double r0 = 1.0, r = 1.0;
            if (!rc) rc = iterate_explicit (r0, r, false);
            elapsed_time += global_dt;
            if (!rc) rc = dflo_b200_commit_step (ctx);
         }
         if (rc)
         {
            error = dflo_b200_last_error (ctx);
            return rc;
         }
         ++time_iter;
         ++steps;
         if (parameters.ang_mom_step > 0 && time_iter % parameters.ang_mom_step == 0 && world == 1) // src/claw.cc:1075-1076
         {
            double am = 0.0;
            rc = compute_angular_momentum (am);
            if (rc) return rc;
            std::printf ("Total angular momentum: %18.8e %24.14e\n", elapsed_time, am);
         }
         // "Save solution for visualization", src/claw.cc:1093-1099
         if (output_enabled && (elapsed_time >= next_output_time || time_iter == next_output_iter
                                || std::fabs (elapsed_time - parameters.final_time) < 1.0e-13))
         {
            rc = output_results (output_dir);
            if (rc) return rc;
            next_output_time = elapsed_time + parameters.output_time_step;
            next_output_iter = time_iter + parameters.output_iter_step;
         }
      }
      return DFLO_OK;
   }

   int ConservationLaw::get_solution (std::vector<double> &u)
   {
      u.assign ((size_t) n_dofs (), 0.0);
      return dflo_b200_get_solution (ctx, u.data (), nullptr, u.size ());
   }

   // compute_angular_momentum, src/claw.cc:604-635 (this rank's cells)
   int ConservationLaw::compute_angular_momentum (double &value)
   {
      std::vector<double> u;
      const int rc = get_solution (u);
      if (rc) return rc;
      int64_t c0 = 0, c1 = flat.n_cells ();
      if (world > 1) dflo_b200_cell_range (ctx, &c0, &c1);
      value = angular_momentum (tab, flat, u.data (), (int) c0, (int) c1);
      return DFLO_OK;
   }

   // output_results: host copy of current_solution -> VTU (host/output.cc).
   // path = a file name: this rank's cells into that file.  path = "" / "dir/": the reference's own numbering --
   //   src (one process):   dir/solution-NNN.vtu ("Writing file ..." line included) + dir/shock.vtu with mu_shock and
   //                        shock_indicator (src/output.cc:33-79);
   //   src_mpi, or sharded: dir/output/solution-NNNN.RRR.vtu = the cells this rank owns plus the "subdomain" array,
   //                        and on rank 0 dir/master_file.visit listing the pieces of every output so far
   //                        (src_mpi/output.cc:34-86; the directory is created if missing).
   int ConservationLaw::output_results (const std::string &path_in)
   {
      std::vector<double> u;
      int rc = get_solution (u); // sharded: fills the owned range only
      if (rc) return rc;
      int64_t c0 = 0, c1 = flat.n_cells ();
      if (world > 1) dflo_b200_cell_range (ctx, &c0, &c1);
      std::string path = path_in, dir;
      const bool numbered = path.empty () || path.back () == '/';
      const bool pieces = numbered && (world > 1 || compat == DFLO_COMPAT_MPI);
      const bool tecplot = parameters.output_format == "tecplot"; // src only; src_mpi always writes vtu
      if (pieces)
      {
         dir = path;
         ::mkdir ((dir + "output").c_str (), 0777);
         std::vector<std::string> names;
         for (int r = 0; r < world; ++r)
         {
            char name[64];
            std::snprintf (name, sizeof (name), "output/solution-%04u.%03d.vtu", output_file_number, r);
            names.push_back (name);
         }
         path = dir + names[rank];
         all_files.push_back (names);
      }
      else if (numbered)
      {
         dir = path;
         char name[64];
         std::snprintf (name, sizeof (name), "solution-%03u.%s", output_file_number, tecplot ? "plt" : "vtu"); // output.cc:48-52
         path = dir + name;
         std::printf ("Writing file %s\n", path.c_str ());
      }
      const bool written = (tecplot && numbered && !pieces)
                              ? write_solution_tecplot (tab, flat, u.data (), parameters.schlieren_plot, elapsed_time, path)
                              : write_solution_vtu (tab, flat, u.data (), parameters.schlieren_plot, elapsed_time, output_file_number, path,
                                                    (int) c0, (int) c1, pieces ? rank : -1);
      if (!written)
      {
         error = "cannot write " + path;
         return DFLO_E_INVALID;
      }
      if (pieces && rank == 0 && !write_visit_record (all_files, dir + "master_file.visit"))
      {
         error = "cannot write " + dir + "master_file.visit";
         return DFLO_E_INVALID;
      }
      if (numbered) ++output_file_number;
      if (numbered && !pieces) return write_shock_file (dir + (tecplot ? "shock.plt" : "shock.vtu"));
      return DFLO_OK;
   }

   // "Write shock indicator", src/output.cc:70-79.  mu_shock is the shock-capturing viscosity of the implicit
   // path (zero on the explicit path); shock_indicator is the one of the last stage.
   int ConservationLaw::write_shock_file (const std::string &path)
   {
      std::vector<double> ind ((size_t) flat.n_cells (), 0.0);
      const int rc = dflo_b200_get_shock_indicator (ctx, ind.data ());
      if (rc) return rc;
      const bool tecplot = path.size () > 4 && path.compare (path.size () - 4, 4, ".plt") == 0;
      if (!(tecplot ? write_shock_tecplot (flat, nullptr, ind.data (), path) : write_shock_vtu (flat, nullptr, ind.data (), path)))
      {
         error = "cannot write " + path;
         return DFLO_E_INVALID;
      }
      return DFLO_OK;
   }
}

//-------------------------------------------------------------------------------------------------
// C ABI (include/dflo_host.h, dflo_claw_*)
//-------------------------------------------------------------------------------------------------
struct dflo_claw
{
   dflo::ConservationLaw *claw;
   dflo_mesh *mesh_view;
};

extern "C" {

dflo_claw *dflo_claw_create (const char *prm_path, const char *mesh_override, const char *overrides, int compat)
{
   if (!prm_path)
   {
      dflo::host_error () = "input.prm path is null";
      return nullptr;
   }
   dflo::ConservationLaw *c = new dflo::ConservationLaw (prm_path, mesh_override ? mesh_override : "", overrides ? overrides : "", compat);
   if (!c->ok ())
   {
      dflo::host_error () = c->error;
      delete c;
      return nullptr;
   }
   dflo_claw *h = new dflo_claw;
   h->claw = c;
   h->mesh_view = nullptr;
   return h;
}

void dflo_claw_destroy (dflo_claw *c)
{
   if (!c) return;
   delete c->mesh_view;
   delete c->claw;
   delete c;
}

const dflo_params *dflo_claw_params (const dflo_claw *c) { return &c->claw->engine_params; }
const int *dflo_claw_periodic_pairs (const dflo_claw *c) { return c->claw->parameters.periodic_pair; }
int dflo_claw_n_dofs (const dflo_claw *c) { return c->claw->n_dofs (); }
double dflo_claw_final_time (const dflo_claw *c) { return c->claw->parameters.final_time; }
// the driver's mesh as a dflo_mesh handle (owned by the claw object; do not destroy)
dflo_mesh *dflo_claw_mesh (dflo_claw *c)
{
   if (!c->mesh_view)
   {
      c->mesh_view = new dflo_mesh;
      c->mesh_view->pm = c->claw->pm;
      c->mesh_view->flat = c->claw->flat;
      c->mesh_view->view = c->mesh_view->flat.view ();
      c->mesh_view->flattened = true;
   }
   return c->mesh_view;
}
const char *dflo_claw_boundary_expression (const dflo_claw *c, int id, int comp)
{
   if (id < 0 || id >= DFLO_MAX_BOUNDARIES || comp < 0 || comp > 3) return nullptr;
   return c->claw->parameters.boundary_expr[id][comp].c_str ();
}
int dflo_claw_initial_condition (dflo_claw *c, double *u, size_t n)
{
   std::vector<double> v;
   const int rc = c->claw->set_initial_condition (v, c->claw->error);
   if (rc)
   {
      dflo::host_error () = c->claw->error;
      return rc;
   }
   if (v.size () != n) return DFLO_E_INVALID;
   std::memcpy (u, v.data (), n * sizeof (double));
   return DFLO_OK;
}
int dflo_claw_setup (dflo_claw *c, int device, int rank, int world, const void *id)
{
   const int rc = c->claw->setup_system (device, rank, world, id);
   if (rc) dflo::host_error () = c->claw->error;
   return rc;
}
dflo_ctx *dflo_claw_engine (dflo_claw *c) { return c->claw->ctx; }
int dflo_claw_run (dflo_claw *c, int max_steps, int verbose, double *elapsed, int *steps_done)
{
   const int it0 = c->claw->time_iter;
   const int rc = c->claw->run (max_steps, verbose != 0);
   if (rc) dflo::host_error () = c->claw->error;
   if (elapsed) *elapsed = c->claw->elapsed_time;
   if (steps_done) *steps_done = c->claw->time_iter - it0;
   return rc;
}
int dflo_claw_get_solution (dflo_claw *c, double *u, size_t n)
{
   return dflo_b200_get_solution (c->claw->ctx, u, nullptr, n);
}
int dflo_claw_write_vtu (dflo_claw *c, const char *path)
{
   const int rc = c->claw->output_results (path ? path : "");
   if (rc) dflo::host_error () = c->claw->error;
   return rc;
}
int dflo_claw_angular_momentum (dflo_claw *c, double *value)
{
   return value ? c->claw->compute_angular_momentum (*value) : DFLO_E_INVALID;
}
void dflo_claw_set_output (dflo_claw *c, const char *dir)
{
   c->claw->output_enabled = dir != nullptr;
   c->claw->output_dir = dir ? dir : "";
   if (!c->claw->output_dir.empty () && c->claw->output_dir.back () != '/') c->claw->output_dir += '/';
}
}
