// Does a non-fp64 instruction issue in the shadow of a DFMA (fp64 pipe busy 2 cycles per warp
// instruction on B200)?  Build: nvcc -arch=sm_100a -O3.  One warp per scheduler, ILP 4.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k (double *out, long long *cyc, int iters, double a, double b, int ia)
{
   double x0 = threadIdx.x, x1 = 1 + threadIdx.x, x2 = 2, x3 = 3;
   int i0 = threadIdx.x, i1 = 1, i2 = 2, i3 = 3;
   float f0 = threadIdx.x, f1 = 1, f2 = 2, f3 = 3;
   long long t0 = clock64 ();
   for (int it = 0; it < iters; ++it)
   {
#pragma unroll
      for (int r = 0; r < 16; ++r)
      {
         if (MODE != 1) { x0 = fma (x0, a, b); x1 = fma (x1, a, b); x2 = fma (x2, a, b); x3 = fma (x3, a, b); }
         if (MODE == 1 || MODE == 2) { i0 = i0 * ia + 1; i1 = i1 * ia + 2; i2 = i2 * ia + 3; i3 = i3 * ia + 4; }
         if (MODE == 3) { f0 = fmaf (f0, (float) a, 1.f); f1 = fmaf (f1, (float) a, 1.f); f2 = fmaf (f2, (float) a, 1.f); f3 = fmaf (f3, (float) a, 1.f); }
      }
   }
   long long t1 = clock64 ();
   out[threadIdx.x] = x0 + x1 + x2 + x3 + i0 + i1 + i2 + i3 + f0 + f1 + f2 + f3;
   if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <int MODE> void run (const char *what)
{
   double *out; long long *cyc;
   cudaMalloc (&out, 8 * 128); cudaMalloc (&cyc, 8);
   k<MODE><<<1, 128>>> (out, cyc, 2000, 0.999, 0.001, 3);
   cudaDeviceSynchronize ();
   long long c; cudaMemcpy (&c, cyc, 8, cudaMemcpyDeviceToHost);
   printf ("%-40s %.2f cycles per group of 4\n", what, (double) c / (2000.0 * 16));
}
int main ()
{
   run<0> ("4 DFMA");
   run<1> ("4 IMAD");
   run<2> ("4 DFMA + 4 IMAD (independent)");
   run<3> ("4 DFMA + 4 FFMA (independent)");
   return 0;
}
