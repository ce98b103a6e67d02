/* TEST INFRASTRUCTURE ONLY (oracle/). Not part of the product; the product never links this.
 *
 * C interface of the CPU oracle: a restatement of dflo's explicit DG path
 * (/root/reference/src/assemble_explicit.cc, claw.cc, limiter.cc, positivity.cc) on top of a
 * restatement of the deal.II pieces that path relies on.
 * PARITY PIN: the point-wise physics is pinned bit-for-bit against the reference's own
 * equation.h (oracle/_ref); the assembly loops have NO reference test/golden vector to pin
 * against (the reference ships no tests and cannot be built without deal.II), so at that level
 * parity is "unpinned" and rests on line-by-line correspondence + analytic invariants.
 */
#ifndef DFLO_ORACLE_H
#define DFLO_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

enum { ORACLE_BASIS_QK = 0, ORACLE_BASIS_PK = 1 };         /* parameters.h:390 */
enum { ORACLE_LIMITER_NONE = 0, ORACLE_LIMITER_TVB = 1, ORACLE_LIMITER_MINMAX = 2 };  /* parameters.h:243, src_mpi/parameters.h:235 */
enum { ORACLE_BC_PERIODIC = 5 };                           /* src_mpi/equation.h BoundaryKind::periodic */
enum { ORACLE_COMPAT_SRC = 0, ORACLE_COMPAT_MPI = 1 };
enum { ORACLE_MAPPING_CARTESIAN = 0, ORACLE_MAPPING_Q1 = 1 };

typedef struct
{
   int basis, degree;
   int flux_type;        /* phys.h PHYS_FLUX_* */
   int limiter_type;     /* ORACLE_LIMITER_* */
   int char_lim, pos_lim, conserve_angular_momentum;
   double M, beta;       /* TVB parameters, limiter.cc:258,266 */
   double gravity;       /* parameters.gravity, assemble_explicit.cc:108 */
   double cfl;
   int bc_kind[10];      /* per boundary id: PHYS_BC_* or ORACLE_BC_PERIODIC */
   int periodic_pair[10];/* partner boundary id when periodic, else -1 */
   int compat;           /* ORACLE_COMPAT_SRC: LxF boundary lambda from own average twice
                            (src/assemble_explicit.cc:203-204); _MPI: BC-reflected average
                            (src_mpi/assemble_explicit.cc:296-321) */
   int n_threads;        /* >1: cells integrated in parallel chunks, serial copier (WorkStream) */
   int shock_indicator;  /* 0 limiter (all cells, indicator.cc:18-22), 1 density, 2 energy (KXRCF, indicator.cc:50-198) */
   int mapping;          /* 0 cartesian (MappingCartesian), 1 q1 (MappingQ1: straight-sided quadrilaterals, claw.cc:165-190);
                            q1: Qk only, no TVB (parameters.cc:545-549), compute_time_step_q (claw.cc:518-557) */
   int local_time_step;  /* "time step type = local": solve() multiplies by dt(cell) (claw.cc:709); the clock moves by the
                            smallest dt(cell), neither capped by "time step" nor clipped at the final time (claw.cc:469-476) */
} oracle_params;

typedef struct oracle_ctx oracle_ctx;

/* Mesh in "gmsh-like" primitive form: vertices, quads in deal.II lexicographic vertex order
 * (SURVEY A1), boundary lines with ids (unlisted boundary faces get id 0). */
oracle_ctx *oracle_create (int n_vertices, const double *vertices /*[nv][2]*/, int n_cells,
                           const int *cells /*[nc][4]*/, int n_blines, const int *blines /*[nb][2]*/,
                           const int *bline_id /*[nb]*/, const oracle_params *prm);
void oracle_destroy (oracle_ctx *);
/* compute_shock_indicator of the current solution / cell averages; values per cell */
void oracle_compute_shock_indicator (oracle_ctx *);
void oracle_get_shock_indicator (const oracle_ctx *, double *ind /*[nc]*/);
const char *oracle_last_error (void);

int oracle_n_cells (const oracle_ctx *);
int oracle_dofs_per_cell (const oracle_ctx *);
int oracle_n_q_face (const oracle_ctx *);
int oracle_n_q_cell (const oracle_ctx *);
int oracle_n_bfaces (const oracle_ctx *);     /* non-periodic boundary faces, ordered by (cell, face) */
int oracle_n_rk (const oracle_ctx *);         /* claw.cc:141-159 */
double oracle_ark (const oracle_ctx *, int rk);

/* topology as the oracle derived it: nbr[c][f] = neighbour cell (>=0, incl. periodic) or
 * -1-bface_index; */
void oracle_get_neighbors (const oracle_ctx *, int *nbr /*[nc][4]*/);
void oracle_get_bfaces (const oracle_ctx *, int *cell, int *face, int *bid, double *xq /*[nbf][nqf][2]*/);
void oracle_get_cell_qpoints (const oracle_ctx *, double *xq /*[nc][nq][2]*/);
void oracle_get_tables (const oracle_ctx *, double *gauss_x, double *gauss_w /*[k+1]*/);

/* IC: Qk interpolate at support points (ic.cc:104-120) / Pk L2 projection (ic.cc:128-168);
 * f = IC function values at the QGauss(k+1)^2 points of every cell, [nc][nq][4]. */
void oracle_set_initial_condition (oracle_ctx *, const double *f);
void oracle_set_solution (oracle_ctx *, const double *u);   /* sets current AND old */
void oracle_get_solution (const oracle_ctx *, double *u);
void oracle_commit_step (oracle_ctx *);                     /* old_solution = current (claw.cc:1110) */
void oracle_set_external_force (oracle_ctx *, const double *f /*[nc][nq][2] or NULL*/); /* src_mpi/assemble_explicit.cc:56-58 */
void oracle_set_bc_values (oracle_ctx *, const double *g /*[nbf][nqf][4]*/);

void oracle_compute_cell_average (oracle_ctx *);            /* claw.cc:562-597 */
void oracle_get_cell_average (const oracle_ctx *, double *avg /*[nc][4]*/);
void oracle_assemble (oracle_ctx *);                        /* assemble_explicit.cc:433-452 */
void oracle_get_rhs (const oracle_ctx *, double *rhs);
double oracle_compute_dt (oracle_ctx *, double elapsed, double final_time, double time_step); /* claw.cc:444-511 */
void oracle_apply_limiter (oracle_ctx *);                   /* limiter.cc:35-65 */
int oracle_apply_positivity (oracle_ctx *);                 /* positivity.cc; 0 ok, -1 negative state, -2 root failure */
void oracle_get_limited_flags (const oracle_ctx *, int *flags /*[nc]: bit0 TVB rewrote, bit1 theta1<1, bit2 theta2<1*/);
/* claw.cc:747-766 for one rk; returns 0 or the positivity error code; *res_norm = l2 of rhs */
int oracle_rk_stage (oracle_ctx *, int rk, double dt, double *res_norm);
/* convenience for timing: n_steps full steps with the current (time-independent) BC values and
 * dt recomputed every step; returns elapsed physical time */
double oracle_run_steps (oracle_ctx *, int n_steps, int *err);

#ifdef __cplusplus
}
#endif
#endif
