// Developer probe: compiles a few row stage kernels alone so that register use, spills and SASS can be inspected in
// seconds:  nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -Idflo_b200/csrc -Xptxas -v -c scripts/probe_row.cu
#include "abi_impl.h"
#include "p2p_halo.cuh"
#include "row_kernel.cuh"
#define X(N,F) template __global__ void dflo::row_stage_kernel<N,F>(const dflo::StageArgs);
#ifdef PN
X(PN, PF)
#else
X(4,3) X(3,4) X(4,2) X(2,0)
#endif
