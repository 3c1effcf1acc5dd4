// mma_shape_probe.cu -- cycles per tcgen05.mma.kind::tf32 (M128 x N x K8, SS operands) on one SM for the operand layouts
// and widths the solver's engine issues: which GEMM shapes run at the tensor floor (128 N / 256 cycles) and which are
// bound by something else (issue cost, operand fetch).
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/mma_shape_probe tools/mma_shape_probe.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include "../diffeqflux.jl_b200/csrc/tc_common.cuh"
using namespace b200tc;

// layout: 0 = K-major SWIZZLE_128B, 1 = K-major SWIZZLE_64B, 2 = MN-major SWIZZLE_128B_BASE32B (wgrad operands)
__device__ uint64_t make_desc(int layout, uint32_t addr) {
  return layout == 0 ? smem_desc_k_sw128(addr) : layout == 1 ? smem_desc_k_sw64(addr) : smem_desc_mn_b32(addr, 4096u);
}
__global__ void __launch_bounds__(128, 1) probe(int N, int la, int lb, int iters, int pattern, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint64_t bar2[3];
  __shared__ uint32_t tmem_base;
  uint8_t* sbase = (uint8_t*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x;
  for (int i = tid; i < (196608) / 4; i += 128) ((float*)sbase)[i] = 0.001f * (float)((i * 7) & 255);
  if (tid < 32) tmem_alloc(&tmem_base, 512);
  if (tid == 0) { mbar_init(&bar, 1); for (int i = 0; i < 3; ++i) mbar_init(&bar2[i], 1); mbar_fence_init(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (pattern >= 4 && tid < 32) {              // converged-warp issue (mma_tf32_elect), 3x pattern
    const uint32_t sb = smem_u32(sbase);
    const uint64_t da = make_desc(la, sb), da2 = make_desc(la, sb + 16384);
    const uint64_t db = make_desc(lb, sb + 32768), db2 = make_desc(lb, sb + 32768 + 65536);
    const uint32_t idesc = idesc_tf32(128, N, la == 2, lb == 2);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint64_t adv = la == 2 ? (uint64_t)((i & 3) * 1024 >> 4) : (uint64_t)((i & 3) * 2);
      const uint32_t acc = tmem_base + ((i & 4) ? 256u : 0u);
      mma_tf32_elect(acc, da2 + adv, db + adv, idesc, 1u);
      mma_tf32_elect(acc, da + adv, db2 + adv, idesc, 1u);
      mma_tf32_elect(acc, da + adv, db + adv, idesc, 1u);
      if (pattern >= 5 && (i & 1)) {
        tc_commit_elect(&bar2[i & 2 ? 1 : 0]);
        if ((i & 3) == 3) tc_commit_elect(&bar2[2]);
      }
    }
    const long long t1 = clock64();
    tc_commit_elect(&bar);
    mbar_wait(&bar, 0);
    const long long t2 = clock64();
    if (tid == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  } else if (pattern < 4 && tid == 0) {
    const uint64_t da = make_desc(la, smem_u32(sbase)), da2 = make_desc(la, smem_u32(sbase + 16384));
    const uint64_t db = make_desc(lb, smem_u32(sbase + 32768)), db2 = make_desc(lb, smem_u32(sbase + 32768 + 65536));
    const uint32_t idesc = idesc_tf32(128, N, la == 2, lb == 2);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint64_t adv = la == 2 ? (uint64_t)((i & 3) * 1024 >> 4) : (uint64_t)((i & 3) * 2);
      const uint32_t acc = tmem_base + ((i & 4) ? 256u : 0u);
      if (pattern == 0) {
        mma_tf32(acc, ((i & 1) ? da2 : da) + adv, ((i & 2) ? db2 : db) + adv, idesc, 1u);
      } else {
        mma_tf32(acc, da2 + adv, db + adv, idesc, 1u);
        mma_tf32(acc, da + adv, db2 + adv, idesc, 1u);
        mma_tf32(acc, da + adv, db + adv, idesc, 1u);
        // pattern 2: a completion commit after every 6 MMAs and a second one after every 12 (the engine's unit_done /
        // mma_done); pattern 3: additionally poll a third barrier that is already complete (the engine's b_full wait)
        if (pattern >= 2 && (i & 1)) {
          tc_commit(&bar2[i & 2 ? 1 : 0]);
          if ((i & 3) == 3) tc_commit(&bar2[2]);
          if (pattern >= 3) { tc_fence_after(); (void)mbar_try_wait(&bar, 1); }
        }
      }
    }
    const long long t1 = clock64();
    tc_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem_base, 512);
}

int main() {
  const int smem = 196608 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long* out;
  cudaMallocManaged(&out, 16);
  const char* names[3] = {"K-major SW128", "K-major SW64", "MN-major B32"};
  const int cases[][3] = {{256, 0, 0}, {256, 0, 1}, {128, 0, 1}, {64, 0, 1}, {64, 0, 0}, {32, 0, 1}, {256, 2, 2}, {64, 2, 2}};
  for (auto& c : cases)
    for (int pattern = 0; pattern < 6; ++pattern) {
      const int iters = 3000, per = pattern ? 3 : 1;
      probe<<<1, 128, smem>>>(c[0], c[1], c[2], iters, pattern, out);   // warm-up
      cudaDeviceSynchronize();
      probe<<<1, 128, smem>>>(c[0], c[1], c[2], iters, pattern, out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
      printf("N=%3d A %-14s B %-14s %s: issue %.1f cyc/MMA, complete %.1f cyc/MMA (floor %d)\n", c[0], names[c[1]], names[c[2]],
             pattern == 0 ? "all distinct" : pattern == 1 ? "3x pattern " : pattern == 2 ? "3x + commits" : pattern == 3 ? "3x+commit+poll" : pattern == 4 ? "3x, converged warp + elect" : "3x + commits, converged", (double)out[0] / (iters * per), (double)out[1] / (iters * per), 128 * c[0] / 256);
    }
  return 0;
}
