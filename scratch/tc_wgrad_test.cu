// Bring-up harness for MN-major tcgen05 operands: G[128][N] = sum_r A[r][128-block] * B[r][N]  (wgrad shape)
#include <cuda_runtime.h>
#include <stdio.h>
#include "../diffeqflux.jl_b200/csrc/tc_common.cuh"
using namespace b200tc;

__device__ __forceinline__ uint64_t smem_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes = 1024u) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;   // stride between 32-element MN atoms
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;   // stride between 8-row K atoms
  d |= (uint64_t)1u << 46;
  d |= (uint64_t)1u << 61;                             // SWIZZLE_128B_BASE32B (MN-major tf32)
  return d;
}
// element (K row r, MN feature f in [0,32)) of a 32-feature MN atom column: 128-B lines, 32-B units XOR (r & 3)
__device__ __forceinline__ uint32_t mn32_off(int r, int f) {
  return (uint32_t)(r * 128 + ((((f >> 3) ^ (r & 3)) & 3) << 5) + ((f & 7) << 2));
}

__global__ void __launch_bounds__(128, 1) wgrad_kernel(const float* __restrict__ A, const float* __restrict__ Bm,
                                                      float* __restrict__ G, int R, int KA, int N, int variant) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  uint8_t* sbase = (uint8_t*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
  const uint32_t tile = R * 128;                 // one chunk tile: R rows x 32 features
  uint8_t* a_hi = sbase;                         // 4 chunk tiles
  uint8_t* a_lo = a_hi + 4 * tile;
  uint8_t* b_hi = a_lo + 4 * tile;               // N/32 chunk tiles
  uint8_t* b_lo = b_hi + (N / 32) * tile;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  for (int i = tid; i < R * 32; i += 128) {      // A: features 0..127 of KA
    const int r = i / 32, f4 = (i % 32) * 4;
    const float4 v = *reinterpret_cast<const float4*>(A + (size_t)r * KA + f4);
    float4 h, l;
    split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
    const uint32_t off = (f4 / 32) * tile + mn32_off(r, f4 % 32);
    *reinterpret_cast<float4*>(a_hi + off) = h;
    *reinterpret_cast<float4*>(a_lo + off) = l;
  }
  for (int i = tid; i < R * (N / 4); i += 128) {
    const int r = i / (N / 4), f4 = (i % (N / 4)) * 4;
    const float4 v = *reinterpret_cast<const float4*>(Bm + (size_t)r * N + f4);
    float4 h, l;
    split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
    const uint32_t off = (f4 / 32) * tile + mn32_off(r, f4 % 32);
    *reinterpret_cast<float4*>(b_hi + off) = h;
    *reinterpret_cast<float4*>(b_lo + off) = l;
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base;
  if (tid == 0) {
    const uint32_t idesc = idesc_tf32(128, N, 1, 1);
    const uint32_t L = (variant & 1) ? 512u : tile, S = (variant & 1) ? tile : 512u;
    const uint64_t dah = smem_desc_mn_sw128(smem_u32(a_hi), L, S), dal = smem_desc_mn_sw128(smem_u32(a_lo), L, S);
    const uint64_t dbh = smem_desc_mn_sw128(smem_u32(b_hi), L, S), dbl = smem_desc_mn_sw128(smem_u32(b_lo), L, S);
    for (int ks = 0; ks < R / 8; ++ks) {
      const uint64_t adv = (uint64_t)(ks * 1024 >> 4);
      mma_tf32(tmem, dal + adv, dbh + adv, idesc, ks ? 1u : 0u);
      mma_tf32(tmem, dah + adv, dbl + adv, idesc, 1u);
      mma_tf32(tmem, dah + adv, dbh + adv, idesc, 1u);
    }
    tc_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 32) {
    float v[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    const int k = warp * 32 + (tid & 31);
    for (int j = 0; j < 32; ++j) G[(size_t)k * N + c0 + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

extern "C" int tc_wgrad_test(const float* A, const float* Bm, float* G, int R, int KA, int N, int variant) {
  const int smem = 2 * 4 * R * 128 + 2 * (N / 32) * R * 128 + 1024;
  cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  wgrad_kernel<<<1, 128, smem>>>(A, Bm, G, R, KA, N, variant);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return -1; }
  return 0;
}
