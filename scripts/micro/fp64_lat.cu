// Micro-benchmarks of the fp64 pipe on B200 (sm_100a): dependent DFMA latency, per-SM throughput
// as a function of resident warps and ILP, MUFU.RSQ64H latency.  Build: nvcc -arch=sm_100a -O3.
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void dfma_chain (double *out, long long *cyc, int iters, double a, double b)
{
   double x[ILP];
#pragma unroll
   for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
   __syncthreads ();
   long long t0 = clock64 ();
   for (int it = 0; it < iters; ++it)
   {
#pragma unroll
      for (int r = 0; r < 16; ++r)
#pragma unroll
         for (int i = 0; i < ILP; ++i) x[i] = fma (x[i], a, b);
   }
   long long t1 = clock64 ();
   double s = 0;
#pragma unroll
   for (int i = 0; i < ILP; ++i) s += x[i];
   out[blockIdx.x * blockDim.x + threadIdx.x] = s;
   if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void rsq_chain (double *out, long long *cyc, int iters)
{
   double x = 1.5 + threadIdx.x * 1e-3;
   long long t0 = clock64 ();
   for (int it = 0; it < iters; ++it)
   {
#pragma unroll
      for (int r = 0; r < 16; ++r)
      {
         double y;
         asm volatile ("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
         x = y + 1.0;
      }
   }
   long long t1 = clock64 ();
   out[threadIdx.x] = x;
   if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <int ILP>
void run (int threads, int blocks, const char *what)
{
   double *out; long long *cyc;
   cudaMalloc (&out, sizeof (double) * threads * blocks);
   cudaMalloc (&cyc, sizeof (long long) * blocks);
   const int iters = 2000;
   dfma_chain<ILP><<<blocks, threads>>> (out, cyc, iters, 0.999, 0.001);
   cudaDeviceSynchronize ();
   cudaEvent_t e0, e1; cudaEventCreate (&e0); cudaEventCreate (&e1);
   cudaEventRecord (e0);
   dfma_chain<ILP><<<blocks, threads>>> (out, cyc, iters, 0.999, 0.001);
   cudaEventRecord (e1); cudaEventSynchronize (e1);
   float ms; cudaEventElapsedTime (&ms, e0, e1);
   long long c; cudaMemcpy (&c, cyc, sizeof (c), cudaMemcpyDeviceToHost);
   const double n = (double) iters * 16 * ILP;
   printf ("%-28s ILP %2d threads %4d blocks %4d: %.2f cycles per DFMA per warp-chain step, %.2f cyc/instr/warp, device %.1f GFMA/s (%.2f TFLOP/s)\n", what, ILP, threads, blocks,
           (double) c / (iters * 16), (double) c / n, n * threads * blocks / ms / 1e6, 2 * n * threads * blocks / ms / 1e9);
   cudaFree (out); cudaFree (cyc);
}
int main ()
{
   run<1> (32, 1, "latency (1 warp)");
   run<2> (32, 1, "1 warp");
   run<4> (32, 1, "1 warp");
   run<8> (32, 1, "1 warp");
   run<1> (128, 1, "1 warp/SMSP");
   run<2> (128, 1, "1 warp/SMSP");
   run<4> (128, 1, "1 warp/SMSP");
   run<1> (256, 1, "2 warps/SMSP");
   run<1> (512, 1, "4 warps/SMSP");
   run<1> (1024, 1, "8 warps/SMSP");
   run<4> (512, 148 * 2, "full device 8 w/SMSP");
   run<8> (1024, 148 * 2, "full device 16 w/SMSP");
   double *out; long long *cyc; cudaMalloc (&out, 8 * 32); cudaMalloc (&cyc, 8);
   rsq_chain<<<1, 32>>> (out, cyc, 1000); cudaDeviceSynchronize ();
   long long c; cudaMemcpy (&c, cyc, 8, cudaMemcpyDeviceToHost);
   printf ("rsqrt.approx.f64 + DADD dependent pair: %.2f cycles\n", (double) c / 16000);
   return 0;
}
