// Developer probe: compiles a few representative stage kernels alone so that register use,
// spills and SASS can be inspected in seconds:  nvcc ... -Idflo_b200/csrc -Xptxas -v -c scripts/probe_kernel.cu
#include "abi_impl.h"
template <class K>
__global__ void __launch_bounds__ (K::THREADS, K::MIN_BLOCKS) probe_kernel (const typename K::Args a)
{
   extern __shared__ __align__ (16) double dflo_smem[];
#pragma unroll
   for (int p = 0; p < K::NPHASE; ++p)
   {
      K::phase (p, a, dflo_smem, threadIdx.x, blockIdx.x);
      if (p + 1 < K::NPHASE) __syncthreads ();
   }
}
#define X(B,N,F) template __global__ void probe_kernel<dflo::StageKernel<B,N,F>>(const dflo::StageArgs);
#ifdef PB   // -DPB=0 -DPN=4 -DPF=3: basis, k+1, flux
X(PB, PN, PF)
#else
X(0,4,3) X(0,3,4) X(1,3,4) X(0,2,0) X(0,4,2)
#endif
