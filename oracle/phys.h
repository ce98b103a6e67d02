/* TEST INFRASTRUCTURE ONLY (oracle/). Not part of the product; the product never links this.
 *
 * Physics interface of the CPU oracle: the point-wise Euler arithmetic that dflo keeps in
 * EulerEquations<2> (/root/reference/src/equation.h). Two interchangeable implementations:
 *   phys_restated.c    plain-C restatement (travels with the repo)
 *   phys_reference.cc  thin wrappers around the reference's own, unmodified equation.h
 *                      (compiled where it lies; output only under oracle/_ref/)
 * Enumerations keep the reference's numeric values:
 *   flux_type      Parameters::Flux::FluxType {lxf, sw, kfvs, roe, hllc}   src/parameters.h:229
 *   boundary kind  EulerEquations::BoundaryKind {inflow, outflow, no_penetration, pressure,
 *                  farfield}                                               src/equation.h:862-869
 * Component order [rho*u, rho*v, rho, E] (equation.h:26-28), gamma = 1.4 (equation.cc:33).
 */
#ifndef DFLO_ORACLE_PHYS_H
#define DFLO_ORACLE_PHYS_H

#ifdef __cplusplus
extern "C" {
#endif

enum { PHYS_FLUX_LXF = 0, PHYS_FLUX_SW = 1, PHYS_FLUX_KFVS = 2, PHYS_FLUX_ROE = 3, PHYS_FLUX_HLLC = 4,
       PHYS_FLUX_KEP = 5 /* src_mpi only: src_mpi/parameters.cc:150-180, src_mpi/equation.h:842-921 */ };
enum { PHYS_BC_INFLOW = 0, PHYS_BC_OUTFLOW = 1, PHYS_BC_SLIP = 2, PHYS_BC_PRESSURE = 3, PHYS_BC_FARFIELD = 4 };

const char *phys_impl_name (void);
/* restated kep_flux of the MPI tree (phys_kep_restated.c), used by both implementations: src/equation.h has none */
void phys_kep_flux_restated (const double n[2], const double Wl[4], const double Wr[4], const double Al[4],
                             const double Ar[4], double out[4]);

/* restated streamline eigenvector matrices (src_mpi/equation.h:299-335) and external forcing vector
 * (src_mpi/equation.h:1189-1202) of the MPI tree (phys_kep_restated.c), used by both implementations */
void phys_eigen_stream_restated (const double W[4], double R[16], double L[16]);
void phys_ext_forcing_restated (const double W[4], const double f[2], double G[4]);

/* claw.h:271-325 numerical_normal_flux -> equation.h lxf/sw/kfvs/roe/hllc */
void phys_numerical_flux (int flux_type, const double n[2], const double Wp[4], const double Wm[4],
                          const double Ap[4], const double Am[4], double out[4]);
/* equation.h:158-193; F[2*c+d] */
void phys_flux_matrix (const double W[4], double F[8]);
/* equation.h:829-850 */
void phys_forcing (const double W[4], double G[4]);
/* equation.h:939-1033 */
void phys_wminus (int kind, const double n[2], const double Wp[4], const double g[4], double Wm[4]);
/* equation.h:225-265; row-major 4x4 */
void phys_eigen (const double W[4], double Rx[16], double Lx[16], double Ry[16], double Ly[16]);
/* equation.h:270-285 / 290-306 (note the internal reordering rho,m,E) */
void phys_to_char (const double L[16], double W[4]);
void phys_to_con (const double R[16], double W[4]);
/* equation.h:84-92, 142-152, 97-114 */
double phys_pressure (const double W[4]);
double phys_sound_speed (const double W[4]);
double phys_max_eigenvalue (const double W[4]);

#ifdef __cplusplus
}
#endif
#endif
