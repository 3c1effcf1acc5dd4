// Bring-up harness for the tcgen05 3xTF32 tile GEMM: D[128][N] = A[128][K] * Bt[N][K]^T.
#include <cuda_runtime.h>
#include <stdio.h>
#include "../diffeqflux.jl_b200/csrc/tc_common.cuh"
using namespace b200tc;

__global__ void __launch_bounds__(128, 1) gemm_kernel(const float* __restrict__ A, const float* __restrict__ Bt,
                                                     float* __restrict__ D, int K, int N, int nprod) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  uint8_t* sbase = (uint8_t*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
  uint8_t* a_hi = sbase;                 // 128 rows * 128 B = 16 KB
  uint8_t* a_lo = a_hi + 16384;
  uint8_t* b_hi = a_lo + 16384;          // N rows * 128 B (<= 32 KB)
  uint8_t* b_lo = b_hi + 32768;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base;
  const uint32_t idesc = idesc_tf32(128, N);
  uint32_t phase = 0;
  for (int kc = 0; kc < K; kc += 32) {
    // A chunk: thread = row, 8 float4 along k
    for (int i = tid; i < 128 * 8; i += 128) {
      const int r = i >> 3, c4 = i & 7;
      const float4 v = *reinterpret_cast<const float4*>(A + (size_t)r * K + kc + c4 * 4);
      float4 h, l;
      split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
      const uint32_t off = sw128_off(r, c4 * 4);
      *reinterpret_cast<float4*>(a_hi + off) = h;
      *reinterpret_cast<float4*>(a_lo + off) = l;
    }
    for (int i = tid; i < N * 8; i += 128) {
      const int r = i >> 3, c4 = i & 7;
      const float4 v = *reinterpret_cast<const float4*>(Bt + (size_t)r * K + kc + c4 * 4);
      float4 h, l;
      split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
      if (nprod == 64) {   // two K=16 units, each [hi N x 64 B][lo N x 64 B] in SWIZZLE_64B
        const int unit = c4 >> 2, kk = (c4 & 3) * 4;
        uint8_t* ub = b_hi + unit * (2 * N * 64);
        const uint32_t off = sw64_off(r, kk);
        *reinterpret_cast<float4*>(ub + off) = h;
        *reinterpret_cast<float4*>(ub + N * 64 + off) = l;
      } else {
        const uint32_t off = sw128_off(r, c4 * 4);
        *reinterpret_cast<float4*>(b_hi + off) = h;
        *reinterpret_cast<float4*>(b_lo + off) = l;
      }
    }
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint64_t dah = smem_desc_k_sw128(smem_u32(a_hi)), dal = smem_desc_k_sw128(smem_u32(a_lo));
      const uint64_t dbh = smem_desc_k_sw128(smem_u32(b_hi)), dbl = smem_desc_k_sw128(smem_u32(b_lo));
      for (int ks = 0; ks < 4; ++ks) {
        const uint64_t adv = (uint64_t)(ks * 32 >> 4);   // +32 B per UMMA_K in 16-B units
        if (nprod == 64) {
          const int unit = ks >> 1;
          uint8_t* ub = b_hi + unit * (2 * N * 64);
          const uint64_t ubh = smem_desc_k_sw64(smem_u32(ub)) + (uint64_t)((ks & 1) * 2);
          const uint64_t ubl = smem_desc_k_sw64(smem_u32(ub + N * 64)) + (uint64_t)((ks & 1) * 2);
          mma_tf32(tmem, dal + adv, ubh, idesc, (kc | ks) ? 1u : 0u);
          mma_tf32(tmem, dah + adv, ubl, idesc, 1u);
          mma_tf32(tmem, dah + adv, ubh, idesc, 1u);
        } else if (nprod >= 3) {
          mma_tf32(tmem, dal + adv, dbh + adv, idesc, (kc | ks) ? 1u : 0u);
          mma_tf32(tmem, dah + adv, dbl + adv, idesc, 1u);
          mma_tf32(tmem, dah + adv, dbh + adv, idesc, 1u);
        } else {
          mma_tf32(tmem, dah + adv, dbh + adv, idesc, (kc | ks) ? 1u : 0u);
        }
      }
      tc_commit(&bar);
    }
    mbar_wait(&bar, phase);
    phase ^= 1;
    tc_fence_after();
    __syncthreads();
  }
  // epilogue: warp w reads lanes [32w, 32w+32)
  for (int c0 = 0; c0 < N; c0 += 32) {
    float v[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    const int r = warp * 32 + (tid & 31);
    for (int j = 0; j < 32; ++j) D[(size_t)r * N + c0 + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

extern "C" int tc_gemm_test(const float* A, const float* Bt, float* D, int K, int N, int nprod) {
  const int smem = 16384 * 2 + 32768 * 2 + 1024;
  cudaFuncSetAttribute(gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  gemm_kernel<<<1, 128, smem>>>(A, Bt, D, K, N, nprod);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return -1; }
  return 0;
}
