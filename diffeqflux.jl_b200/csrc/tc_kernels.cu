// tc_kernels.cu -- instantiations of the persistent solver with the tcgen05 RHS engine (own translation unit
// so the two halves of the library compile in parallel).
#define B200_TC_TU 1
#include "solver.cuh"

using namespace b200;

template <int FL>
static cudaError_t launch_tc_flavor(const SolveParams& P, int smem_bytes, cudaStream_t stream) {
  cudaError_t e = cudaFuncSetAttribute(solve_kernel<FL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  if (e != cudaSuccess) return e;
  void* args[] = {const_cast<SolveParams*>(&P)};
  return cudaLaunchCooperativeKernel(reinterpret_cast<void*>(solve_kernel<FL, true>), dim3(P.nctas), dim3(TC_THREADS),
                                     args, (size_t)smem_bytes, stream);
}

cudaError_t b200_tc_init_tableau() {
  const Tableau tab = host_tableau();
  return cudaMemcpyToSymbol(c_tab, &tab, sizeof tab);
}

cudaError_t b200_tc_launch(int flavor, const SolveParams& P, int smem_bytes, cudaStream_t stream) {
  switch (flavor) {
    case FL_FWD_NODE: return launch_tc_flavor<FL_FWD_NODE>(P, smem_bytes, stream);
    case FL_ADJ_NODE_BACKSOLVE: return launch_tc_flavor<FL_ADJ_NODE_BACKSOLVE>(P, smem_bytes, stream);
    case FL_ADJ_NODE_INTERP: return launch_tc_flavor<FL_ADJ_NODE_INTERP>(P, smem_bytes, stream);
    default: return cudaErrorInvalidValue;
  }
}
