// Micro-benchmark: what does it cost to bring one 8x4-cell Q3 tile (32 cells x 512 B) and its 24 halo cells into
// shared memory, in the access pattern of row_stage_kernel, at 4 blocks of 160 threads per SM?  Variants:
//   0  per-cell bulk copies (TMA unit), own + halo, padded rows (what the kernel does)      56 copies / tile
//   1  per-cell bulk copies, own cells only                                                 32 copies / tile
//   2  one dense bulk copy of the own cells (16 KB) + per-cell halo copies                  25 copies / tile
//   3  one dense bulk copy of the own cells only                                             1 copy  / tile
//   4  plain 16-byte coalesced loads of own + halo cells into the padded rows (no TMA)
//   5  like 4, own cells only
//   6  persistent: 592 blocks loop over the tiles, per-cell copies of tile i+1 (own + halo) in flight while tile i
//      is "worked on" (a dependent FMA chain of `work` iterations per thread)
//   7  one tile per block with the same `work` (the non-pipelined counterpart of 6)
// Every variant finishes with a checksum of the staged data so nothing is optimised away.  Also prints plain HBM copy
// bandwidth for reference.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o staging_bench staging_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

constexpr int D = 64, TC = 32, NH = 24, CS = D + 2, THREADS = 160;
constexpr int NX = 256, NY = 256, TX = 8, TY = 4;

__device__ __forceinline__ unsigned smem_addr (const void *p) { return (unsigned) __cvta_generic_to_shared (p); }
__device__ __forceinline__ void mbar_init (void *bar, unsigned count)
{
   asm volatile ("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr (bar)), "r"(count) : "memory");
   asm volatile ("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx (void *bar, unsigned bytes)
{
   asm volatile ("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr (bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait (void *bar, unsigned parity)
{
   asm volatile ("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(smem_addr (bar)),
                 "r"(parity)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s (void *dst, const void *src, unsigned bytes, void *bar)
{
   asm volatile ("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr (dst)), "l"(src), "r"(bytes),
                 "r"(smem_addr (bar))
                 : "memory");
}

struct Args
{
   const double *u;
   const int *halo; // [n_tiles][NH]
   double *out;     // [n_tiles]
   int n_tiles, variant, work;
};

__device__ __forceinline__ double consume (const double *su, int n_cells, int stride, int work)
{
   double s = 0.0;
   for (int i = threadIdx.x; i < n_cells * D / 2; i += THREADS)
   {
      const int cell = i / (D / 2), j = i % (D / 2);
      const double2 v = *reinterpret_cast<const double2 *> (su + cell * stride + 2 * j);
      s += v.x + v.y;
   }
   for (int i = 0; i < work; ++i) s = fma (s, 1.0000001, 1e-9);
   return s;
}

__device__ __forceinline__ void issue_tile (const Args &A, int tile, double *su, void *bar, bool with_halo, bool dense_own)
{
   const int tid = threadIdx.x;
   const int c0 = tile * TC;
   if (tid == 0)
   {
      const unsigned bytes = (unsigned) (TC + (with_halo ? NH : 0)) * D * 8u;
      mbar_expect_tx (bar, bytes);
   }
   __syncthreads ();
   if (dense_own)
   {
      if (tid == 0) bulk_g2s (su, A.u + (size_t) c0 * D, TC * D * 8u, bar);
      if (with_halo && tid >= TC && tid < TC + NH) bulk_g2s (su + TC * D + (tid - TC) * CS, A.u + (size_t) A.halo[tile * NH + tid - TC] * D, D * 8u, bar);
   }
   else if (tid < TC)
      bulk_g2s (su + tid * CS, A.u + (size_t) (c0 + tid) * D, D * 8u, bar);
   else if (with_halo && tid < TC + NH)
      bulk_g2s (su + tid * CS, A.u + (size_t) A.halo[tile * NH + tid - TC] * D, D * 8u, bar);
}

__global__ void __launch_bounds__ (THREADS, 4) staging_kernel (const Args A)
{
   extern __shared__ __align__ (16) double sm[];
   const int tid = threadIdx.x;
   const int v = A.variant;
   if (v <= 3 || v == 7)
   {
      double *su = sm + 2;
      const int tile = blockIdx.x;
      if (tid == 0) mbar_init (sm, 1);
      __syncthreads ();
      const bool with_halo = (v == 0 || v == 2 || v == 7), dense = (v == 2 || v == 3);
      issue_tile (A, tile, su, sm, with_halo, dense);
      mbar_wait (sm, 0);
      double s = consume (su, TC, dense ? D : CS, A.work);
      if (with_halo) s += consume (su + (dense ? TC * D : TC * CS), NH, CS, 0);
      if (s == 123.456) A.out[tile] = s;
      if (tid == 0) A.out[tile] = s;
   }
   else if (v == 4 || v == 5)
   {
      double *su = sm + 2;
      const int tile = blockIdx.x, c0 = tile * TC;
      const int ncell = v == 4 ? TC + NH : TC;
      for (int i = tid; i < ncell * D / 2; i += THREADS)
      {
         const int cell = i / (D / 2), j = i % (D / 2);
         const int g = cell < TC ? c0 + cell : A.halo[tile * NH + cell - TC];
         *reinterpret_cast<double2 *> (su + cell * CS + 2 * j) = *reinterpret_cast<const double2 *> (A.u + (size_t) g * D + 2 * j);
      }
      __syncthreads ();
      double s = consume (su, ncell, CS, A.work);
      if (tid == 0) A.out[tile] = s;
   }
   else if (v == 8 || v == 9)
   {
      if (tid == 0) A.out[blockIdx.x] = 1.0; // block launch cost only
   }
   else if (v == 10)
   {
      // persistent, single buffer: issue -> wait -> consume per tile (block launch amortised, nothing overlapped)
      double *su = sm + 2;
      if (tid == 0) mbar_init (sm, 1);
      __syncthreads ();
      int it = 0;
      for (int tile = blockIdx.x; tile < A.n_tiles; tile += gridDim.x, ++it)
      {
         issue_tile (A, tile, su, sm, true, false);
         mbar_wait (sm, it & 1);
         double s = consume (su, TC, CS, A.work) + consume (su + TC * CS, NH, CS, 0);
         if (tid == 0) A.out[tile] = s;
         __syncthreads ();
      }
   }
   else if (v == 11)
   {
      // persistent, own cells double buffered, halo single buffered after the own cells
      double *buf[2] = {sm + 4, sm + 4 + TC * CS};
      if (tid == 0)
      {
         mbar_init (sm, 1);
         mbar_init (sm + 1, 1);
      }
      __syncthreads ();
      int tile = blockIdx.x, it = 0;
      if (tile < A.n_tiles) issue_tile (A, tile, buf[0], sm, false, false);
      for (; tile < A.n_tiles; tile += gridDim.x, ++it)
      {
         const int b = it & 1, next = tile + gridDim.x;
         if (next < A.n_tiles) issue_tile (A, next, buf[b ^ 1], sm + (b ^ 1), false, false);
         mbar_wait (sm + b, (it >> 1) & 1);
         double s = consume (buf[b], TC, CS, A.work);
         if (tid == 0) A.out[tile] = s;
         __syncthreads ();
      }
   }
   else if (v == 6)
   {
      // persistent, double buffered: [bar0 bar1 | buf0 | buf1]
      double *buf[2] = {sm + 4, sm + 4 + (TC + NH) * CS};
      if (tid == 0)
      {
         mbar_init (sm, 1);
         mbar_init (sm + 1, 1);
      }
      __syncthreads ();
      int tile = blockIdx.x, it = 0;
      if (tile < A.n_tiles) issue_tile (A, tile, buf[0], sm, true, false);
      for (; tile < A.n_tiles; tile += gridDim.x, ++it)
      {
         const int b = it & 1, next = tile + gridDim.x;
         if (next < A.n_tiles) issue_tile (A, next, buf[b ^ 1], sm + (b ^ 1), true, false);
         mbar_wait (sm + b, (it >> 1) & 1);
         double s = consume (buf[b], TC, CS, A.work) + consume (buf[b] + TC * CS, NH, CS, 0);
         if (tid == 0) A.out[tile] = s;
         __syncthreads ();
      }
   }
}

__global__ void copy_kernel (const double2 *a, double2 *b, size_t n)
{
   for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) b[i] = a[i];
}

int main (int argc, char **argv)
{
   const int n_cells = NX * NY, n_tiles = n_cells / TC;
   std::vector<int> halo ((size_t) n_tiles * NH);
   // tile-major numbering: tile (ti, tj), cell (i, j) in tile -> ((tj * (NX/TX) + ti) * TC + j * TX + i
   auto cell_id = [&] (int x, int y) {
      x = (x + NX) % NX;
      y = (y + NY) % NY;
      return ((y / TY) * (NX / TX) + x / TX) * TC + (y % TY) * TX + x % TX;
   };
   for (int t = 0; t < n_tiles; ++t)
   {
      const int ti = t % (NX / TX), tj = t / (NX / TX), x0 = ti * TX, y0 = tj * TY;
      int k = 0;
      for (int j = 0; j < TY; ++j) halo[(size_t) t * NH + k++] = cell_id (x0 - 1, y0 + j);
      for (int j = 0; j < TY; ++j) halo[(size_t) t * NH + k++] = cell_id (x0 + TX, y0 + j);
      for (int i = 0; i < TX; ++i) halo[(size_t) t * NH + k++] = cell_id (x0 + i, y0 - 1);
      for (int i = 0; i < TX; ++i) halo[(size_t) t * NH + k++] = cell_id (x0 + i, y0 + TY);
   }
   double *u, *out, *flush;
   int *d_halo;
   const size_t nbytes = (size_t) n_cells * D * 8;
   cudaMalloc (&u, nbytes);
   cudaMalloc (&out, n_tiles * 8);
   cudaMalloc (&flush, 256u << 20);
   cudaMalloc (&d_halo, halo.size () * 4);
   cudaMemset (u, 0, nbytes);
   cudaMemcpy (d_halo, halo.data (), halo.size () * 4, cudaMemcpyHostToDevice);
   const size_t smem = (4 + 2 * (TC + NH) * CS) * 8;
   cudaFuncSetAttribute (staging_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
   cudaEvent_t e0, e1;
   cudaEventCreate (&e0);
   cudaEventCreate (&e1);
   const int works[] = {0, 500};
   for (int flushed = 0; flushed < 2; ++flushed)
      for (int v = 0; v <= 11; ++v)
         for (int w : works)
         {
                        if (w && v != 6 && v != 7 && v != 0 && v != 10 && v != 11) continue;
            Args A{u, d_halo, out, n_tiles, v, w};
            const size_t sm_bytes = v == 6 ? smem : (2 + (TC + NH) * CS) * 8 + 12800; // ~42.5 KB like the stage kernel: 4 blocks / SM
            const int grid = v == 6 ? 444 : (v == 10 || v == 11 || v == 9) ? 592 : n_tiles;
            float total = 0;
            const int reps = 20;
            for (int r = 0; r < reps + 2; ++r)
            {
               if (flushed) cudaMemsetAsync (flush, r, 256u << 20);
               cudaEventRecord (e0);
               staging_kernel<<<grid, THREADS, sm_bytes>>> (A);
               cudaEventRecord (e1);
               cudaEventSynchronize (e1);
               float ms;
               cudaEventElapsedTime (&ms, e0, e1);
               if (r >= 2) total += ms;
            }
            cudaError_t err = cudaGetLastError ();
            printf ("%s variant %d work %4d: %7.2f us  (%s)\n", flushed ? "L2 flushed" : "L2 warm   ", v, w, 1e3 * total / reps, cudaGetErrorString (err));
         }
   {
      float total = 0;
      for (int r = 0; r < 12; ++r)
      {
         cudaMemsetAsync (flush, r, 256u << 20);
         cudaEventRecord (e0);
         copy_kernel<<<148 * 8, 256>>> ((const double2 *) u, (double2 *) flush, nbytes / 16);
         cudaEventRecord (e1);
         cudaEventSynchronize (e1);
         float ms;
         cudaEventElapsedTime (&ms, e0, e1);
         if (r >= 2) total += ms;
      }
      printf ("plain copy of the 33.5 MB vector: %7.2f us (%.0f GB/s read+write)\n", 1e3 * total / 10, 2 * nbytes / (total / 10 * 1e-3) / 1e9);
   }
   return 0;
}
