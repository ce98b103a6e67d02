/* dflo_b200.h -- C ABI of the B200-native explicit DG residual / RK-stage engine.
 *
 * This is the drop-in boundary for dflo's explicit hot path.  dflo (cpraveen/dflo) has no
 * plugin/FFI API of its own; the seam this library replaces is the body of the RK loop in
 * ConservationLaw<dim>::iterate_explicit (reference src/claw.cc:732-771), i.e.
 *     assemble_system(integrator)            src/assemble_explicit.cc:433-452
 *     solve() rk3 branch + RK combine        src/claw.cc:694-713, 757-760
 *     compute_cell_average()                 src/claw.cc:562-597
 *     compute_shock_indicator()              src/indicator.cc:15-31 (limiter), 50-198 (KXRCF density / energy)
 *     apply_limiter() TVB Qk/Pk              src/limiter.cc:224-516
 *     apply_positivity_limiter()             src/positivity.cc:16-208
 * plus compute_time_step() (src/claw.cc:444-511).  INTEGRATION.md shows the call sites a dflo
 * maintainer would add.  Every entry point is plain C: opaque handle, raw pointers and sizes,
 * int return code (0 = ok, negative = DFLO_E_*).  Caller owns every pointer it passes; the
 * library copies during the call and never retains host pointers.  A ctx is bound to one CUDA
 * device and is NOT thread-safe (one host thread per ctx), mirroring the reference where the
 * stage loop itself is serial (src/claw.cc:726-772).
 *
 * All floating point is IEEE fp64.  DoF layout at this boundary is the reference's: for the
 * serial code global = cell*D + comp*n_s + node (FESystem of a DG element); pass
 * dof_map to translate any other deal.II numbering.  Component order [rho*u, rho*v, rho, E]
 * (src/equation.h:26-28).
 */
#ifndef DFLO_B200_H
#define DFLO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DFLO_B200_ABI_VERSION 4
#define DFLO_MAX_BOUNDARIES 10 /* Parameters::AllParameters::max_n_boundaries, src/parameters.h:370 */

/* error codes */
enum
{
   DFLO_OK = 0,
   DFLO_E_INVALID = -1,        /* bad argument / inconsistent mesh */
   DFLO_E_UNSUPPORTED = -2,    /* hanging nodes, non-Cartesian mapping, degree out of range */
   DFLO_E_CUDA = -3,           /* CUDA runtime failure, see dflo_b200_last_error() */
   DFLO_E_NEGATIVE_STATE = -4, /* "Fatal: Negative states", src/positivity.cc:33-37 */
   DFLO_E_POSLIM_ROOT = -5,    /* "Problem in positivity limiter", src/positivity.cc:160-169 */
   DFLO_E_NCCL = -6,
   DFLO_E_EXPR = -7,           /* boundary expression did not parse */
   DFLO_E_NO_DEVICE = -8       /* no CUDA device: the engine has no CPU fallback */
};

/* Parameters::Flux::FluxType, src/parameters.h:229; kep (kinetic-energy preserving, entropy stable) exists in the
 * MPI tree only: src_mpi/parameters.cc:150-180, src_mpi/equation.h:842-921 */
enum { DFLO_FLUX_LXF = 0, DFLO_FLUX_SW = 1, DFLO_FLUX_KFVS = 2, DFLO_FLUX_ROE = 3, DFLO_FLUX_HLLC = 4, DFLO_FLUX_KEP = 5 };
/* EulerEquations::BoundaryKind, src/equation.h:862-869; periodic from src_mpi/equation.h */
enum { DFLO_BC_INFLOW = 0, DFLO_BC_OUTFLOW = 1, DFLO_BC_SLIP = 2, DFLO_BC_PRESSURE = 3, DFLO_BC_FARFIELD = 4, DFLO_BC_PERIODIC = 5 };
/* Parameters::AllParameters::BasisType, src/parameters.h:390 */
enum { DFLO_BASIS_QK = 0, DFLO_BASIS_PK = 1 };
/* Parameters::Limiter::LimiterType, src/parameters.h:243; minmax (Barth-Jespersen type, Qk only) exists in the
 * MPI tree only: src_mpi/parameters.h:235, src_mpi/limiter.cc:400-553 */
enum { DFLO_LIMITER_NONE = 0, DFLO_LIMITER_TVB = 1, DFLO_LIMITER_MINMAX = 2 };
/* Parameters::Limiter::ShockIndType, src/parameters.h: which cells the TVB limiter may touch.
 * limiter: every cell (src/indicator.cc:18-22); density / energy: the KXRCF indicator of that variable
 * (src/indicator.cc:50-198), cells with indicator > 1 (src/limiter.cc:263, 406) */
enum { DFLO_INDICATOR_LIMITER = 0, DFLO_INDICATOR_DENSITY = 1, DFLO_INDICATOR_ENERGY = 2 };
/* which tree's semantics where src/ and src_mpi/ differ (SURVEY.md 8a "semantic forks") */
enum { DFLO_COMPAT_SRC = 0, DFLO_COMPAT_MPI = 1 };

/* per-(cell,face) flag bits in dflo_flat_mesh::face_flags */
enum
{
   DFLO_FACE_OWNER = 1,    /* this cell integrates the face (MeshWorker::loop visits an interior face
                              once, from the cell that compares smaller; src/assemble_explicit.cc:440):
                              the numerical flux is evaluated with THIS cell as "plus" side */
   DFLO_FACE_PERIODIC = 2, /* neighbour reached through a periodic pair: both sides integrate with
                              their own normal (src_mpi/assemble_explicit.cc:186-260); it is a TVB
                              neighbour like any other (src_mpi/claw.cc:417-465) */
   DFLO_FACE_FLIP = 4,     /* periodic face_flip: neighbour's face points run backwards
                              (src_mpi/assemble_explicit.cc:247-250) */
   /* hanging nodes (one level of refinement across a face, as deal.II keeps it; MeshWorker integrates such a face from the
    * FINE side on sub-faces): */
   DFLO_FACE_COARSER = 8,  /* the neighbour is coarser: this face is one half of its face `neighbor_face` ... */
   DFLO_FACE_CHILD1 = 16,  /* ... the second half along the coarse cell's line (else the first) */
   DFLO_FACE_HANGING = 32  /* this face has a hanging node: two finer neighbours, listed in dflo_flat_mesh::hanging;
                              `neighbor` holds the first of them */
};

/* "mapping" of input.prm (src/claw.cc:165-190): cartesian = MappingCartesian (axis-aligned rectangles, the fast kernels);
 * q1 = MappingQ1 (straight-sided quadrilaterals: per-point Jacobians, general normals, compute_time_step_q) -- Qk basis,
 * no TVB / positivity limiter (src/parameters.cc:545-549 refuses TVB and Pk off Cartesian grids). */
enum { DFLO_MAPPING_CARTESIAN = 0, DFLO_MAPPING_Q1 = 1 };

/* Mesh topology and geometry flattened once from the host mesh (deal.II Triangulation +
 * DoFHandler in dflo; src/claw.cc:270-386). */
typedef struct
{
   int32_t n_cells;
   const double *cell_origin;   /* [n_cells][2] lower-left vertex */
   const double *cell_size;     /* [n_cells][2] hx, hy */
   const int32_t *neighbor;     /* [n_cells][4] face order left,right,bottom,top (deal.II faces 0..3):
                                   cell index >= 0, or -1 - boundary_face_index */
   const uint8_t *face_flags;   /* [n_cells][4] DFLO_FACE_* */
   int32_t n_boundary_faces;
   const int32_t *bface_cell;   /* [n_boundary_faces] */
   const int32_t *bface_face;   /* [n_boundary_faces] local face number 0..3 */
   const int32_t *bface_id;     /* [n_boundary_faces] boundary id 0..9 */
   /* mapping = q1 only (may be NULL for mapping = cartesian, where cell_origin / cell_size say it all and the
    * neighbour across face f sees the face as its face f ^ 1): */
   const double *cell_vertices;   /* [n_cells][4][2] vertices in deal.II's lexicographic order (cell->vertex(0..3)) */
   const uint8_t *neighbor_face;  /* [n_cells][4] the neighbour's local number of the shared face; DFLO_FACE_FLIP in
                                     face_flags when the two cells run along the face in opposite directions */
   /* faces with a hanging node (needs cell_vertices / neighbor_face; no TVB / minmax limiter): */
   int32_t n_hanging_faces;
   const int32_t *hanging;        /* [n_hanging_faces][6] coarse cell, its face, then the two fine cells in the order of the
                                     coarse line: fine cell 0, its face, fine cell 1, its face */
} dflo_flat_mesh;

/* The subset of Parameters::AllParameters (src/parameters.h:112-414) the hot path reads. */
typedef struct
{
   int32_t basis;                      /* DFLO_BASIS_* ("basis") */
   int32_t degree;                     /* "degree": Qk 0..4, Pk 0..3 */
   int32_t flux_type;                  /* DFLO_FLUX_* ("flux") */
   int32_t limiter_type;               /* DFLO_LIMITER_* ("type" in subsection limiter) */
   int32_t char_lim;                   /* "characteristic limiter" */
   int32_t pos_lim;                    /* "positivity limiter" */
   int32_t conserve_angular_momentum;  /* Pk only, src/limiter.cc:496-500 */
   int32_t compat;                     /* DFLO_COMPAT_* */
   double M;                           /* TVB constant */
   double beta;                        /* minmod beta */
   double gravity;                     /* "gravity" multiplier, src/assemble_explicit.cc:108 */
   double cfl;                         /* "cfl" */
   double time_step;                   /* "time step" (<=0: unused), src/claw.cc:471-472 */
   int32_t bc_kind[DFLO_MAX_BOUNDARIES];   /* DFLO_BC_* per boundary id */
   int32_t shock_indicator;            /* DFLO_INDICATOR_* ("shock indicator" in subsection limiter) */
   int32_t mapping;                    /* DFLO_MAPPING_* ("mapping"); 0 = cartesian */
   int32_t local_time_step;            /* "time step type = local" (src/claw.cc:444-478, 694-713): every cell advances with its own
                                          dt(cell); the clock moves by the smallest one, which is neither capped by "time step" nor
                                          clipped at the final time.  dflo_b200_rk_stage then uses the per-cell values of the last
                                          dflo_b200_compute_dt, its dt argument only sets the BC time */
   int32_t reserved1;
} dflo_params;

typedef struct dflo_ctx dflo_ctx;

/* ---- life cycle (replaces ConservationLaw::setup_system, src/claw.cc:270-386) ---- */
int dflo_b200_abi_version (void);
int dflo_b200_create (const dflo_flat_mesh *mesh, const dflo_params *prm, int device, dflo_ctx **out);
/* Cell-sharded context, one process per GPU.  Every rank passes the SAME global mesh; the library
 * partitions by cell id into `world` contiguous ranges and builds the halo lists.  nccl_unique_id:
 * 128 bytes from dflo_b200_nccl_unique_id() on rank 0, broadcast by the caller. */
int dflo_b200_create_sharded (const dflo_flat_mesh *mesh, const dflo_params *prm, int device, int rank,
                              int world, const void *nccl_unique_id, dflo_ctx **out);
int dflo_b200_nccl_unique_id (void *out128);
void dflo_b200_destroy (dflo_ctx *ctx);
const char *dflo_b200_strerror (int code);
const char *dflo_b200_last_error (const dflo_ctx *ctx);

/* ---- sizes ---- */
int dflo_b200_dofs_per_cell (const dflo_ctx *ctx);
int dflo_b200_n_q_face (const dflo_ctx *ctx);
int dflo_b200_n_rk (const dflo_ctx *ctx);            /* src/claw.cc:141-159 */
double dflo_b200_ark (const dflo_ctx *ctx, int rk);
int64_t dflo_b200_n_cells_owned (const dflo_ctx *ctx);
int64_t dflo_b200_cell_range (const dflo_ctx *ctx, int64_t *begin, int64_t *end); /* owned global range */

/* ---- state: current_solution / old_solution (src/claw.h:212-213) ----
 * u has n = n_cells_global*D entries in the reference layout; dof_map (may be NULL = identity)
 * gives for reference position cell*D+i the index into u (deal.II's get_dof_indices).  A sharded
 * ctx reads / writes only the entries of its owned cells.  set_solution also sets old_solution and
 * recomputes the cell averages (src/claw.cc:997). */
int dflo_b200_set_solution (dflo_ctx *ctx, const double *u, const uint32_t *dof_map, size_t n);
int dflo_b200_get_solution (dflo_ctx *ctx, double *u, const uint32_t *dof_map, size_t n);
int dflo_b200_get_cell_average (dflo_ctx *ctx, double *avg /* [n_cells_global][4] */);
int dflo_b200_commit_step (dflo_ctx *ctx);           /* old_solution = current_solution, claw.cc:1110 */

/* ---- boundary data g(x,t) (FunctionParser, src/assemble_explicit.cc:161-165) ----
 * Either hand the values at the face quadrature points for the coming stage(s) ... */
int dflo_b200_set_boundary_values (dflo_ctx *ctx, const double *g /* [n_boundary_faces][n_q_face][4] */);
/* ... or the muparser-style expression in x,y,t of one component (src/parameters.cc:470-511); it is
 * compiled to bytecode and evaluated on the device at the BC time of every stage. */
int dflo_b200_set_boundary_expression (dflo_ctx *ctx, int boundary_id, int component, const char *expr);
/* ---- external force of the MPI tree ("f_0 value" / "f_1 value", src_mpi/parameters.cc:355-360, 488-497) ----
 * Two muparser-style expressions in x,y, evaluated once at the cell quadrature points (with t = 0 like the
 * reference, whose FunctionParser time is never set: src_mpi/assemble_explicit.cc:56-58).  From then on the forcing
 * term gravity * G of the cell integral uses G = (rho f, m.f) (src_mpi/equation.h:1189-1202) instead of the
 * hard-wired f = (0,-1) of src/equation.h:829-850. */
int dflo_b200_set_external_force (dflo_ctx *ctx, const char *fx_expr, const char *fy_expr);

/* ---- the hot path ---- */
/* assemble_system: right_hand_side from current_solution (src/assemble_explicit.cc:433-452) */
int dflo_b200_assemble_rhs (dflo_ctx *ctx, double t_bc);
int dflo_b200_get_rhs (dflo_ctx *ctx, double *rhs, const uint32_t *dof_map, size_t n);
/* one pass of the rk loop body, src/claw.cc:747-766.  res_norm may be NULL (skips the reduction). */
int dflo_b200_rk_stage (dflo_ctx *ctx, int rk, double t_bc, double dt, double *res_norm);
/* compute_time_step, global time step (src/claw.cc:444-511) */
int dflo_b200_compute_dt (dflo_ctx *ctx, double elapsed_time, double final_time, double *dt);
/* initial limiting, src/claw.cc:997-1003: cell average, indicator, apply_limiter (no positivity) */
int dflo_b200_limit_initial_condition (dflo_ctx *ctx);
/* n whole time steps of the loop src/claw.cc:1026-1110 (compute_time_step, iterate_explicit with
 * bc_time = t for rk 0 and t+dt afterwards in COMPAT_SRC, old = current), entirely on the device
 * (CUDA graph replay, no host round trip per step).  *elapsed_time is read and advanced. */
int dflo_b200_advance (dflo_ctx *ctx, int n_steps, double final_time, double *elapsed_time, double *last_dt);
/* poll the device error word (DFLO_E_NEGATIVE_STATE / DFLO_E_POSLIM_ROOT), synchronises */
int dflo_b200_poll_error (dflo_ctx *ctx);
/* shock_indicator of the last stage (src/indicator.cc), [n_cells_global]; 1e20 everywhere for DFLO_INDICATOR_LIMITER */
int dflo_b200_get_shock_indicator (dflo_ctx *ctx, double *ind);
/* per-cell limiter activity of the last stage: bit0 TVB rewrote the cell, bit1 theta1<1, bit2 theta2<1 */
int dflo_b200_get_limited_flags (dflo_ctx *ctx, int32_t *flags /* [n_cells_global] */);

/* ---- instrumentation ---- */
/* number of kernels launched by this ctx since creation */
int64_t dflo_b200_launch_count (const dflo_ctx *ctx);
/* CUDA stream the ctx launches on (cudaStream_t as void*), for event timing by the caller */
void *dflo_b200_stream (const dflo_ctx *ctx);
int dflo_b200_synchronize (dflo_ctx *ctx);
/* average device time (ms) of the stage kernel of RK stage rk alone over `reps` launches, CUDA
 * events on the ctx stream, with flush_bytes of scratch rewritten before each launch to evict L2
 * (0 = no flush); the solution state is left untouched */
int dflo_b200_time_stage_kernel (dflo_ctx *ctx, int rk, int reps, size_t flush_bytes, float *avg_ms);
/* device-resident time of the last dflo_b200_advance call, measured with CUDA events on the ctx stream */
int dflo_b200_last_advance_ms (dflo_ctx *ctx, float *ms);

#ifdef __cplusplus
}
#endif
#endif
