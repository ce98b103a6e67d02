// Developer micro-benchmark of the thread-per-cell Pk stage kernel (dflo_b200/csrc/cell_stage.cuh):
// P2 / HLLC on a periodic nx x ny box of smooth data, timed alone with CUDA events.  Compiles in
// seconds (one instantiation), so launch-shape variants (-DPK_MIN_BLOCKS=.., -DPK_THREADS=..) can be
// compared without rebuilding the library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I dflo_b200/csrc scripts/micro/pk_cell_bench.cu \
//        dflo_b200/csrc/tables.cc -o scripts/micro/pk_cell_bench
#ifdef PK_V1
#include "cell_stage_v1.cuh"
#else
#include "cell_stage.cuh"
#endif
#include "tables.h"
#include "tables_pack.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace dflo;
typedef PkCellStageKernel<3, FLUX_HLLC> K;

__global__ void __launch_bounds__ (K::THREADS, K::MIN_BLOCKS) bench_kernel (const CellStageArgs a)
{
   extern __shared__ __align__ (16) double smem[];
#ifdef PK_NOUNROLL
#pragma unroll 1
#else
#pragma unroll
#endif
   for (int p = 0; p < K::NPHASE; ++p)
   {
#ifdef PK_PHASE_MASK
      if (!((PK_PHASE_MASK >> p) & 1)) continue;
#endif
      K::phase (p, a, smem, threadIdx.x, blockIdx.x);
      if (p + 1 < K::NPHASE) __syncthreads ();
   }
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf ("%s: %s\n", #x, cudaGetErrorString (e)); exit (1); } } while (0)

int main (int argc, char **argv)
{
   const int nx = argc > 1 ? atoi (argv[1]) : 1600, ny = argc > 2 ? atoi (argv[2]) : 160, reps = 20;
   const int TX = argc > 3 ? atoi (argv[3]) : nx, TY = argc > 4 ? atoi (argv[4]) : 1; // cell order: TX x TY tiles (default: row-major)
   const int nc = nx * ny, D = K::D;
   auto id = [&] (int i, int j) { return ((j / TY) * (nx / TX) + i / TX) * TX * TY + (j % TY) * TX + i % TX; };
   FeTables tab;
   if (!build_tables (BASIS_PK, 2, tab)) return 1;
   std::vector<double> flat = pack_stage_tables (tab);
   CK (cudaMemcpyToSymbol (c_pk_tab, flat.data (), flat.size () * sizeof (double), (size_t) 3 * PK_TAB_MAX * sizeof (double)));
   std::vector<double> u ((size_t) nc * D, 0.0), geom ((size_t) nc * 4), time = {0.0, 1e-5, 0.0, 1.0};
   std::vector<int> nbr ((size_t) nc * 4);
   std::vector<unsigned char> ff ((size_t) nc * 4);
   for (int j = 0; j < ny; ++j)
      for (int i = 0; i < nx; ++i)
      {
         const int c = id (i, j);
         const double x = (i + 0.5) / nx, y = (j + 0.5) / ny;
         const double rho = 1.0 + 0.2 * sin (6.28 * x) * cos (6.28 * y), vx = 0.3, vy = 0.1, p = 1.0 + 0.1 * cos (6.28 * x);
         u[(size_t) c * D + 0 * K::NS] = rho * vx;
         u[(size_t) c * D + 1 * K::NS] = rho * vy;
         u[(size_t) c * D + 2 * K::NS] = rho;
         u[(size_t) c * D + 3 * K::NS] = p / 0.4 + 0.5 * rho * (vx * vx + vy * vy);
         for (int m = 1; m < K::NS; ++m) u[(size_t) c * D + 2 * K::NS + m] = 0.01 / m;
         geom[4 * c + 0] = (double) i / nx;
         geom[4 * c + 1] = (double) j / ny;
         geom[4 * c + 2] = 1.0 / nx;
         geom[4 * c + 3] = 1.0 / ny;
         nbr[4 * c + 0] = id ((i + nx - 1) % nx, j);
         nbr[4 * c + 1] = id ((i + 1) % nx, j);
         nbr[4 * c + 2] = id (i, (j + ny - 1) % ny);
         nbr[4 * c + 3] = id (i, (j + 1) % ny);
         ff[4 * c + 0] = ff[4 * c + 2] = 0;
         ff[4 * c + 1] = ff[4 * c + 3] = FACE_OWNER;
      }
   double *d_u, *d_uo, *d_out, *d_avg, *d_avgo, *d_geom, *d_time, *d_tab, *d_flush;
   int *d_nbr;
   unsigned char *d_ff;
   const size_t nb = (size_t) nc * D * sizeof (double), flush = 256u << 20;
   CK (cudaMalloc (&d_u, nb)); CK (cudaMalloc (&d_uo, nb)); CK (cudaMalloc (&d_out, nb));
   CK (cudaMalloc (&d_avg, (size_t) nc * 32)); CK (cudaMalloc (&d_avgo, (size_t) nc * 32)); CK (cudaMalloc (&d_geom, (size_t) nc * 32));
   CK (cudaMalloc (&d_time, 32)); CK (cudaMalloc (&d_tab, flat.size () * 8)); CK (cudaMalloc (&d_nbr, (size_t) nc * 16)); CK (cudaMalloc (&d_ff, (size_t) nc * 4));
   CK (cudaMalloc (&d_flush, flush));
   CK (cudaMemcpy (d_u, u.data (), nb, cudaMemcpyHostToDevice)); CK (cudaMemcpy (d_uo, u.data (), nb, cudaMemcpyHostToDevice));
   CK (cudaMemset (d_avg, 0, (size_t) nc * 32));
   CK (cudaMemcpy (d_geom, geom.data (), (size_t) nc * 32, cudaMemcpyHostToDevice)); CK (cudaMemcpy (d_time, time.data (), 32, cudaMemcpyHostToDevice));
   CK (cudaMemcpy (d_tab, flat.data (), flat.size () * 8, cudaMemcpyHostToDevice));
   CK (cudaMemcpy (d_nbr, nbr.data (), (size_t) nc * 16, cudaMemcpyHostToDevice)); CK (cudaMemcpy (d_ff, ff.data (), (size_t) nc * 4, cudaMemcpyHostToDevice));
   CellStageArgs a;
   a.u = d_u; a.u_old = d_uo; a.out = d_out; a.avg = d_avg; a.avg_out = d_avgo; a.nbr = d_nbr; a.fflags = d_ff; a.geom = d_geom;
   a.bc_g = nullptr; a.bkind = nullptr; a.tab = d_tab; a.time = d_time; a.dt_cell = nullptr; a.ext_force = nullptr;
   a.n_compute = nc; a.n_keep = nc; a.mode = MODE_STAGE; a.compat_mpi = 0; a.ark = 0.75; a.gravity = 0.0; a.pf_blocks = argc > 5 ? atoi (argv[5]) : 148; if (argc > 6) a.ark = atof (argv[6]);
   const size_t smem = K::SMEM_DOUBLES * sizeof (double);
   CK (cudaFuncSetAttribute (bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
   int occ = 0;
   CK (cudaOccupancyMaxActiveBlocksPerMultiprocessor (&occ, bench_kernel, K::THREADS, smem));
   cudaEvent_t e0, e1;
   CK (cudaEventCreate (&e0)); CK (cudaEventCreate (&e1));
   double total = 0.0;
   for (int r = 0; r < reps + 3; ++r)
   {
      CK (cudaMemsetAsync (d_flush, r, flush));
      CK (cudaEventRecord (e0));
      bench_kernel<<<K::grid (nc), K::THREADS, smem>>> (a);
      CK (cudaEventRecord (e1));
      CK (cudaEventSynchronize (e1));
      float ms;
      CK (cudaEventElapsedTime (&ms, e0, e1));
      if (r >= 3) total += ms;
   }
   CK (cudaGetLastError ());
   std::vector<double> out ((size_t) nc * D);
   CK (cudaMemcpy (out.data (), d_out, nb, cudaMemcpyDeviceToHost));
   double cs = 0.0; // order-independent check value: cells visited in lattice order
   for (int j = 0; j < ny; ++j)
      for (int i = 0; i < nx; ++i)
         for (int k = 0; k < D; ++k) cs += out[(size_t) id (i, j) * D + k] * (1.0 + 0.001 * ((i * 7 + j * 13 + k) % 17));
   printf ("cells %d tiles %dx%d threads %d blocks/SM %d smem %zu: %.1f us per launch, checksum %.15e\n", nc, TX, TY, K::THREADS, occ, smem, 1e3 * total / reps, cs);
   return 0;
}
