// cycles per tcgen05.mma kind::tf32 (M=128, SS mode, K=8) issued back to back by one thread
#include <cuda_runtime.h>
#include <stdio.h>
#include "../diffeqflux.jl_b200/csrc/tc_common.cuh"
using namespace b200tc;
__global__ void __launch_bounds__(128, 1) bench(int N, int iters, int kind, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  uint8_t* sbase = (uint8_t*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x;
  for (int i = tid; i < (16384 * 2 + 32768 * 2) / 4; i += 128) ((float*)sbase)[i] = 0.001f * (i & 255);
  if (tid < 32) tmem_alloc(&tmem_base, 512);
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    const uint64_t da = smem_desc_k_sw128(smem_u32(sbase)), da2 = smem_desc_k_sw128(smem_u32(sbase + 16384));
    const uint64_t db = smem_desc_k_sw128(smem_u32(sbase + 32768)), db2 = smem_desc_k_sw128(smem_u32(sbase + 65536));
    const uint32_t idesc = idesc_tf32(128, N);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint64_t adv = (uint64_t)((i & 3) * 2);
      if (kind == 0) mma_tf32(tmem_base, da + adv, db + adv, idesc, 1u);               // same operands
      else if (kind == 1) mma_tf32(tmem_base + ((i & 1) ? 256u : 0u), (i & 1 ? da2 : da) + adv, (i & 2 ? db2 : db) + adv, idesc, 1u);
      else {   // 3xTF32 pattern: lo*hi, hi*lo, hi*hi
        mma_tf32(tmem_base, da2 + adv, db + adv, idesc, 1u);
        mma_tf32(tmem_base, da + adv, db2 + adv, idesc, 1u);
        mma_tf32(tmem_base, da + adv, db + adv, idesc, 1u);
      }
    }
    long long t1 = clock64();
    tc_commit(&bar);
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem_base, 512);
}
int main() {
  long long* out;
  cudaMallocManaged(&out, 16);
  const int smem = 16384 * 2 + 32768 * 2 + 1024;
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int kind = 0; kind < 3; ++kind)
    for (int N : {64, 128, 256}) {
      const int iters = 4096;
      bench<<<1, 128, smem>>>(N, iters, kind, out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
      const int nm = kind == 2 ? iters * 3 : iters;
      printf("kind %d N=%3d: issue %.1f cyc/mma, complete %.1f cyc/mma  (floor %d)\n", kind, N, (double)out[0] / nm, (double)out[1] / nm, N / 2);
    }
  // all SMs busy at once (power / shared limits)
  return 0;
}
