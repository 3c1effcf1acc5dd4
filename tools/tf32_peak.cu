// tf32_peak.cu -- direct measurement of the dense TF32 tensor-core peak of the whole GPU: every SM issues
// tcgen05.mma.cta_group::1.kind::tf32 (M128 x N256 x K8, both operands in shared memory, fp32 accumulators in TMEM)
// back to back.  This is the denominator of bench.py's roofline: the 3xTF32 algorithmic ceiling is this figure / 3.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/tf32_peak tools/tf32_peak.cu
//   run  : gpurun_out/tf32_peak [seconds]   -> one JSON line (burst = best single launch, sustained = back to back)
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>
#include "../diffeqflux.jl_b200/csrc/tc_common.cuh"
using namespace b200tc;

__global__ void __launch_bounds__(128, 1) mma_stream(int N, int iters, int pattern) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  uint8_t* sbase = (uint8_t*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x;
  for (int i = tid; i < (16384 * 2 + 32768 * 2) / 4; i += 128) ((float*)sbase)[i] = 0.001f * (float)((i * 7 + blockIdx.x) & 255);
  if (tid < 32) tmem_alloc(&tmem_base, 512);
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid < 32) {          // one converged warp issues (mma_tf32_elect): a single divergent lane is issue-bound at ~135 cycles per MMA
    const uint64_t da = smem_desc_k_sw128(smem_u32(sbase)), da2 = smem_desc_k_sw128(smem_u32(sbase + 16384));
    const uint64_t db = smem_desc_k_sw128(smem_u32(sbase + 32768)), db2 = smem_desc_k_sw128(smem_u32(sbase + 65536));
    const uint32_t idesc = idesc_tf32(128, N);
    for (int i = 0; i < iters; ++i) {
      const uint64_t adv = (uint64_t)((i & 3) * 2);
      const uint32_t acc = tmem_base + ((i & 4) ? 256u : 0u);
      if (pattern == 0) {
        mma_tf32_elect(acc, ((i & 1) ? da2 : da) + adv, ((i & 2) ? db2 : db) + adv, idesc, 1u);
      } else {   // the 3xTF32 issue pattern of the solver: lo*hi, hi*lo, hi*hi into one accumulator
        mma_tf32_elect(acc, da2 + adv, db + adv, idesc, 1u);
        mma_tf32_elect(acc, da + adv, db2 + adv, idesc, 1u);
        mma_tf32_elect(acc, da + adv, db + adv, idesc, 1u);
      }
    }
    tc_commit_elect(&bar);
    mbar_wait(&bar, 0);
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem_base, 512);
}

int main(int argc, char** argv) {
  const double seconds = argc > 1 ? atof(argv[1]) : 3.0;
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  const int smem = 16384 * 2 + 32768 * 2 + 1024;
  cudaFuncSetAttribute(mma_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  printf("{\"gpu\": \"%s\", \"sms\": %d", prop.name, sms);
  for (int pattern = 0; pattern < 2; ++pattern) {
    const int N = 256;
    const int per = pattern ? 3 : 1;
    const int iters = 400000 / per;                               // ~25 ms per launch
    const double flop = (double)sms * iters * per * 2.0 * 128 * N * 8;
    mma_stream<<<sms, 128, smem>>>(N, iters, pattern);            // warm-up
    cudaDeviceSynchronize();
    double best = 1e30;
    for (int r = 0; r < 10; ++r) {
      cudaEventRecord(e0);
      mma_stream<<<sms, 128, smem>>>(N, iters, pattern);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      best = std::min(best, (double)ms);
    }
    // sustained: launches back to back for `seconds`
    int launches = std::max(4, (int)(seconds * 1e3 / best));
    cudaEventRecord(e0);
    for (int r = 0; r < launches; ++r) mma_stream<<<sms, 128, smem>>>(N, iters, pattern);
    cudaEventRecord(e1);
    cudaError_t err = cudaEventSynchronize(e1);
    if (err != cudaSuccess) { printf(", \"error\": \"%s\"}\n", cudaGetErrorString(err)); return 1; }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf(", \"%s\": {\"burst_tflops\": %.1f, \"sustained_tflops\": %.1f, \"sustained_s\": %.2f, \"ms_per_launch\": %.3f}",
           pattern ? "tf32_3x_pattern" : "tf32_dense", flop / (best * 1e-3) / 1e12, flop * launches / (ms * 1e-3) / 1e12, ms * 1e-3, best);
  }
  printf(", \"shape\": \"M128 N256 K8 kind::tf32 cta_group::1, SS operands, 1 CTA/SM, converged-warp issue\"}\n");
  return 0;
}
