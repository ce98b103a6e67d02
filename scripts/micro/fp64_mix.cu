// Developer micro-benchmark: what fp64 instruction streams other than pure DFMA chains achieve on a whole B200.
//  (a) DFMA / DMUL / DADD / std_min-max (DSETP + selects) streams at ILP 4, whole chip, 8 warps per SM;
//  (b) the HLLC flux of euler.cuh on registers (no memory traffic) at 4 .. 32 warps per SM: the arithmetic ceiling of
//      the Pk stage kernel's face phase as a function of occupancy.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I dflo_b200/csrc scripts/micro/fp64_mix.cu -o scripts/micro/fp64_mix
#include "euler.cuh"

#include <cstdio>
#include <cuda_runtime.h>
using namespace dflo;

template <int MODE>
__global__ void stream_kernel (double *out, int iters, double a, double b)
{
   double x0 = threadIdx.x, x1 = 1 + threadIdx.x, x2 = 2, x3 = 3;
   for (int it = 0; it < iters; ++it)
   {
#pragma unroll
      for (int r = 0; r < 16; ++r)
      {
         if (MODE == 0) { x0 = fma (x0, a, b); x1 = fma (x1, a, b); x2 = fma (x2, a, b); x3 = fma (x3, a, b); }
         if (MODE == 1) { x0 = x0 * a; x1 = x1 * a; x2 = x2 * a; x3 = x3 * a; }
         if (MODE == 2) { x0 = x0 + b; x1 = x1 + b; x2 = x2 + b; x3 = x3 + b; }
         if (MODE == 3) { x0 = std_min (x0, x1 + r); x1 = std_max (x1, x2); x2 = std_min (x2, x3); x3 = std_max (x3, x0); }
      }
   }
   out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3;
}

__global__ void hllc_kernel (double *out, int iters, double seed)
{
   double Wl[4] = {0.3 + 1e-3 * threadIdx.x, 0.1, 1.0 + seed, 2.5}, Wr[4] = {0.25, 0.12, 0.9 + seed, 2.4}, acc[4] = {0, 0, 0, 0};
   for (int it = 0; it < iters; ++it)
   {
      double H[4];
      hllc_flux_x (Wl, Wr, H);
#pragma unroll
      for (int c = 0; c < 4; ++c)
      {
         acc[c] += H[c];
         Wl[c] += 1e-9 * H[c]; // the next problem depends on this one: one chain per thread, like one face point after the other
      }
   }
   out[blockIdx.x * blockDim.x + threadIdx.x] = acc[0] + acc[1] + acc[2] + acc[3];
}

template <class F> float time_ms (F f)
{
   cudaEvent_t e0, e1;
   cudaEventCreate (&e0); cudaEventCreate (&e1);
   f ();
   cudaEventRecord (e0);
   f ();
   cudaEventRecord (e1);
   cudaEventSynchronize (e1);
   float ms;
   cudaEventElapsedTime (&ms, e0, e1);
   return ms;
}

int main ()
{
   double *out;
   cudaMalloc (&out, 8 * 148 * 1024 * 4);
   const int iters = 4000;
   const char *names[4] = {"DFMA", "DMUL", "DADD", "std_min/max"};
   for (int m = 0; m < 4; ++m)
   {
      float ms = 0;
      if (m == 0) ms = time_ms ([&] { stream_kernel<0><<<148, 256>>> (out, iters, 0.999, 0.001); });
      if (m == 1) ms = time_ms ([&] { stream_kernel<1><<<148, 256>>> (out, iters, 0.999, 0.001); });
      if (m == 2) ms = time_ms ([&] { stream_kernel<2><<<148, 256>>> (out, iters, 0.999, 0.001); });
      if (m == 3) ms = time_ms ([&] { stream_kernel<3><<<148, 256>>> (out, iters, 0.999, 0.001); });
      const double ops = 148.0 * 8 * iters * 64; // warp-level source operations
      printf ("%-12s 8 warps/SM: %.3f ms, %.3f source ops per cycle per SM (1.965 GHz)\n", names[m], ms, ops / (ms * 1e-3 * 1.965e9 * 148));
   }
   for (int warps : {4, 8, 12, 16, 24, 32})
   {
      const int it = 2000;
      float ms = time_ms ([&] { hllc_kernel<<<148, 32 * warps>>> (out, it, 0.01); });
      printf ("HLLC on registers, %2d warps/SM: %.3f ms, %.1f cycles per flux per warp, %.3f fluxes per cycle per SM\n", warps, ms,
              ms * 1e-3 * 1.965e9 / it, (double) warps * it / (ms * 1e-3 * 1.965e9));
   }
   return 0;
}
