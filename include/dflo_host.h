/* dflo_host.h -- C ABI of the host-side front end of the standalone driver (mesh, input.prm,
 * initial condition, the ConservationLaw time loop).  These are the parts of dflo that stay on
 * the host and that a deal.II build of dflo already has (GridIn, ParameterHandler,
 * VectorTools::interpolate; reference src/claw.cc:953-1003, src/parameters.cc, src/ic.cc); they
 * are exported so that tests and bench.py can drive the engine through the same code the
 * dflo_b200 executable uses.  Nothing here touches the GPU except dflo_claw_*, which calls the
 * engine only through include/dflo_b200.h.
 */
#ifndef DFLO_HOST_H
#define DFLO_HOST_H

#include "dflo_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

const char *dflo_host_last_error (void);

/* ---- meshes (gmsh .msh v2 as read by GridIn::read_msh, or the built-in generators that
 *      reproduce the reference's transfinite .geo files) ---- */
typedef struct dflo_mesh dflo_mesh;
/* kind / args:
 *   "rectangle"          nx ny x0 x1 y0 y1 id_left id_right id_bottom id_top
 *   "rectangle_skew"     nx ny x0 x1 y0 y1 id_left id_right id_bottom id_top amplitude rotate   (general quadrilaterals
 *                        for mapping = q1: interior vertices displaced smoothly; rotate != 0: mixed cell orientations)
 *   "rectangle_refined"  nx ny x0 x1 y0 y1 id_left id_right id_bottom id_top i0 i1 j0 j1 [rotate]   (cells i0 <= i < i1, j0 <= j < j1 split
 *                        into four: faces with hanging nodes around the patch)
 *   "compression_corner" nx1 nx2 ny (cells)            (examples/compression_corner/corner.geo: trapezoids, mapping = q1)
 *   "isentropic_vortex"  n_cells_per_side              (examples/isentropic_vortex/grid.geo)
 *   "sod_tube"           nx ny                         (examples/sod_shock_tube/tube.geo)
 *   "double_mach"        ny_cells                      (examples/double_mach_reflection/grid.geo)
 *   "forward_step"       cl                            (examples/forward_step/step.geo) */
dflo_mesh *dflo_mesh_create (const char *kind, const double *args, int n_args);
dflo_mesh *dflo_mesh_read_gmsh (const char *path);
int dflo_mesh_write_gmsh (const dflo_mesh *m, const char *path);
void dflo_mesh_destroy (dflo_mesh *m);
int dflo_mesh_n_vertices (const dflo_mesh *m);
int dflo_mesh_n_cells (const dflo_mesh *m);
int dflo_mesh_n_blines (const dflo_mesh *m);
const double *dflo_mesh_vertices (const dflo_mesh *m);  /* [nv][2] */
const int *dflo_mesh_cells (const dflo_mesh *m);        /* [nc][4] deal.II lexicographic */
const int *dflo_mesh_blines (const dflo_mesh *m);       /* [nb][2] */
const int *dflo_mesh_bline_ids (const dflo_mesh *m);    /* [nb] */
/* neighbour lists, MeshWorker face ownership, periodic partners, boundary-face list */
int dflo_mesh_flatten (dflo_mesh *m, const int bc_kind[DFLO_MAX_BOUNDARIES], const int periodic_pair[DFLO_MAX_BOUNDARIES]);
const dflo_flat_mesh *dflo_mesh_flat (const dflo_mesh *m);

/* ---- expressions (deal.II FunctionParser strings) ---- */
/* values of expr at n points, variables x,y,t; returns 0 or DFLO_E_EXPR */
int dflo_expr_eval (const char *expr, int n, const double *x, const double *y, double t, double *out);

/* ---- input.prm + the ConservationLaw driver ---- */
typedef struct dflo_claw dflo_claw;
/* Parse an input.prm (deal.II ParameterHandler syntax, the schema of src/parameters.cc) and build
 * the mesh: "mesh file" is read as gmsh v2 relative to the .prm unless mesh_override != NULL, in
 * which case mesh_override = "<kind> <args...>" selects a built-in generator.  degree < 0 keeps the
 * file's value; overrides = extra "set key = value" / subsection text applied after the file. */
dflo_claw *dflo_claw_create (const char *prm_path, const char *mesh_override, const char *overrides, int compat);
void dflo_claw_destroy (dflo_claw *c);
const dflo_params *dflo_claw_params (const dflo_claw *c);
const int *dflo_claw_periodic_pairs (const dflo_claw *c);      /* [10] partner id or -1 */
dflo_mesh *dflo_claw_mesh (dflo_claw *c);
int dflo_claw_n_dofs (const dflo_claw *c);
double dflo_claw_final_time (const dflo_claw *c);
/* boundary expression of (boundary id, component) as written in the file */
const char *dflo_claw_boundary_expression (const dflo_claw *c, int id, int comp);
/* set_initial_condition (src/ic.cc:104-182) on the host: u in the reference DoF layout */
int dflo_claw_initial_condition (dflo_claw *c, double *u, size_t n);
/* setup_system + set_initial_condition + initial limiting on `device` (src/claw.cc:981-1003);
 * sharded when world > 1 */
int dflo_claw_setup (dflo_claw *c, int device, int rank, int world, const void *nccl_unique_id);
dflo_ctx *dflo_claw_engine (dflo_claw *c);
/* the time loop of ConservationLaw::run (src/claw.cc:1026-1110): up to max_steps steps or until
 * final time; prints the reference's per-step lines when verbose */
int dflo_claw_run (dflo_claw *c, int max_steps, int verbose, double *elapsed_time, int *steps_done);
int dflo_claw_get_solution (dflo_claw *c, double *u, size_t n);
/* output_results (src/output.cc:33-87, src_mpi/output.cc:34-86).  path = a file name: the solution (this rank's cells)
 * only.  path = NULL, "" or "dir/": the reference's numbered files in that directory, counter advancing --
 *   compat src, one process: solution-NNN.vtu + shock.vtu;
 *   compat mpi, or a sharded setup: output/solution-NNNN.RRR.vtu (this rank's cells, extra array "subdomain") and, on
 *   rank 0, master_file.visit.
 * Every cell is written as degree x degree sub-quads like DataOut::build_patches (mapping, fe.degree); point data in
 * the order of DataOutBase::write_vtu: the vector ranges XMomentum__YMomentum and XVelocity__YVelocity (3 components),
 * then Density Energy Pressure [schlieren_plot] (src/equation.h:32-59, src/equation.cc:59-166). */
int dflo_claw_write_vtu (dflo_claw *c, const char *path);
/* dir != NULL: dflo_claw_run writes the initial solution and then follows "output: time step / iter step" and the
 * final time like src/claw.cc:1010-1017, 1093-1099 (dir "" = working directory, as the reference); NULL: off (default) */
void dflo_claw_set_output (dflo_claw *c, const char *dir);
/* compute_angular_momentum (src/claw.cc:604-635) of the cells this rank owns: int (x m_y - y m_x) with QGauss(k+1)^2.
 * dflo_claw_run prints the reference's "Total angular momentum:" line every `output: compute angular momentum` steps
 * (src/claw.cc:1075-1076) on unsharded setups. */
int dflo_claw_angular_momentum (dflo_claw *c, double *value);
int dflo_host_angular_momentum (const dflo_mesh *m, int basis, int degree, const double *u, size_t n, double *value);
/* the same writers on a host copy of the solution (reference DoF layout, n = n_cells * 4 * n_s), no engine involved;
 * mesh must be flattened.  shock file: mu_shock may be NULL (zeros); both arrays are written the way DataOut writes
 * cell vectors, as point data constant on the four vertices of each cell. */
int dflo_host_write_solution_vtu (const dflo_mesh *m, int basis, int degree, const double *u, size_t n, int schlieren_plot,
                                  double time, unsigned int cycle, const char *path);
/* "output: format = tecplot" (src/output.cc:51-52, 65-66): the same patches and variables as an ASCII FEBLOCK zone of
 * quadrilaterals; dflo_claw_write_vtu / the run loop write solution-NNN.plt + shock.plt instead of .vtu when the input
 * file asks for it (src tree only; src_mpi always writes vtu) */
int dflo_host_write_solution_tecplot (const dflo_mesh *m, int basis, int degree, const double *u, size_t n, int schlieren_plot,
                                      double time, const char *path);
/* the piece of cells [cell_begin, cell_end) (cell_end < 0: to the last cell); subdomain >= 0 adds the "subdomain" array */
int dflo_host_write_solution_piece_vtu (const dflo_mesh *m, int basis, int degree, const double *u, size_t n, int schlieren_plot,
                                        double time, unsigned int cycle, int cell_begin, int cell_end, int subdomain, const char *path);
int dflo_host_write_shock_vtu (const dflo_mesh *m, const double *mu_shock, const double *shock_indicator, const char *path);

#ifdef __cplusplus
}
#endif
#endif
