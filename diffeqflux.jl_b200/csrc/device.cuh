// device.cuh -- persistent fused Tsit5 kernels for libb200node.so (sm_100a).
//
// One cooperative launch integrates the whole batch: every CTA owns tiles of TB trajectories, runs
// the 6 new stages of a Tsit5 step (FSAL) with the Dense/tanh MLP evaluated from shared memory,
// publishes its share of the batch-wide error norm, meets the other CTAs at a grid barrier and then
// replays the same fp32 PI controller, so all CTAs (and all ranks) take the same accept/reject branch.
// Replaces the OrdinaryDiffEq loop x Lux Dense chain reached from DiffEqFlux src/neural_de.jl:55-58
// and the SciMLSensitivity augmented backward solve selected at src/neural_de.jl:57.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>

#include "tc_common.cuh"

namespace b200 {

constexpr int MAXL = 8;
constexpr int NTHREADS = 256;
constexpr int NWARPS = NTHREADS / 32;
constexpr int MAX_CTAS = 1024;
constexpr int MAX_RANKS = 8;
constexpr int NSTEP_SLAB = 7;  // per-CTA weight-gradient slabs of a step: stages 2..6 (0..4) + the two FSAL slabs G1/G7 (5, 6; swapped on accept)
constexpr int NSLAB = 8;       // + slab 7: the CTA's running quadrature sum of the Quadrature / Gauss adjoints

enum Flavor : int {
  FL_FWD_NODE = 0,
  FL_FWD_FFJORD = 1,
  FL_ADJ_NODE_INTERP = 2,
  FL_ADJ_NODE_BACKSOLVE = 3,
  FL_ADJ_FFJORD = 4,
};

enum Act : int { ACT_ID = 0, ACT_TANH = 1, ACT_RELU = 2, ACT_SIGMOID = 3, ACT_SOFTPLUS = 4 };

struct MlpDesc {
  int nl;
  int pre;
  int dims[MAXL + 1];
  int dpad[MAXL + 1];
  int act[MAXL];
  int w_off[MAXL];   // packed [dpad[l]][dpad[l+1]] (also the layout of mu / gradient slabs)
  int b_off[MAXL];   // packed [dpad[l+1]]
  int wt_off[MAXL];  // transposed copy [dpad[l+1]][dpad[l]] in the wtp buffer
  int fw_off[MAXL];  // offsets in the caller's flat parameter vector
  int fb_off[MAXL];
  int ppad;          // packed size (multiple of 4)
  int nparams;       // true parameter count
};

struct SmemPlan {
  int act_off[MAXL + 1];  // a_l buffers [TB][dpad[l]], l = 0..nl-1
  int hd_off[MAXL + 1];   // hdot_l buffers (FFJORD adjoint), l = 0..nl-1
  int q_off[MAXL + 1];    // q_l = W_l hdot_{l-1} buffers (FFJORD adjoint), l = 1..nl-1
  int dl_off;             // seed buffer [TB][dpad[nl]] (delta_L / hbar_L)
  int dl2_off;            // second seed (hdbar_L, FFJORD adjoint) or m(z) (FFJORD forward)
  int x_off;              // raw input when a cube pre-layer is present
  int ll_off;             // lambda_logp per row (FFJORD adjoint)
  int roww_off, rowwt_off;// row-per-thread FFJORD forward: the packed parameters / transposed weights staged once per solve
  int total;              // floats
};

// Multi-GPU step mailbox (CUDA-IPC mapped into every peer): rank r writes {value, tag} into slot
// [parity][r] of every rank.  Tags are 64-bit epochs that grow monotonically over the life of the
// context (solve sequence number << 24 | exchange index), compared with >=.
constexpr int XMU_CAP = 1 << 19;             // packed parameters supported by the cross-rank mu exchange
struct Mailbox {
  volatile double val[2][2][MAX_RANKS];      // [exchange parity][component][src rank]
  volatile unsigned long long tag[2][MAX_RANKS];
  volatile unsigned long long xtag[2][MAX_RANKS][MAX_CTAS];   // per mu-slice (= per CTA) tags
  double xmu[2][MAX_RANKS][2][XMU_CAP];      // [parity][src rank][quantity][packed parameter index], fp64 so
                                             // that the sharded total rounds like the unsharded one
};

struct Ctrl {
  unsigned int bar;
  unsigned long long xflag;  // epoch of the last published cross-rank sum
  int abort_flag;
  int retcode;
  int naccept, nreject, nf;
  int nsteps;                // accepted steps recorded on the dense tape
  float t_final, dt_last, tie_dist;
  int n_out;
  double xsum[2][2];         // cross-rank sums [parity][component]
  double partial[2][MAX_CTAS][2];
};

// Tensor-core (tcgen05 / TMEM, 3xTF32) RHS engine.
//   Weight "unit" = (layer, 16 reduction indices): [hi tile: rows x 64 B][lo tile], K-major SWIZZLE_64B,
//   <= 32 KB, streamed by cp.async.bulk into a 4-deep ring.  Forward orientation: rows = outputs n,
//   reduction over inputs k (wimg / w_off).  Dgrad orientation: rows = inputs k, reduction over outputs n
//   (wimg_t / t_off).
//   Adjoint scratch (per CTA, global/L2): every activation a_l and cotangent delta_l of the tile as hi/lo
//   tf32 in the MN-major SWIZZLE_128B_BASE32B layout the weight-gradient GEMMs read:
//   [row quarter (4)][hi|lo][32-feature column][32 rows x 128 B].
constexpr int MAXL_TC = 4;
struct TcParams {
  int enabled;
  const uint8_t* wimg;
  const uint8_t* wimg_t;
  unsigned int w_off[MAXL_TC];     // byte offset of layer l's first forward unit
  unsigned int t_off[MAXL_TC];     // byte offset of layer l's first dgrad unit
  uint8_t* scratch;                // adjoint: [nctas][scratch_cta]
  unsigned int scratch_cta;
  unsigned int sa_off[MAXL_TC];    // a_l   (l = 0..L-1): fp32 stash, [chunk][column group][row][TC_CW] (16 KB per 32-feature chunk)
  unsigned int sd_off[MAXL_TC + 1];// delta_l (l = 1..L)
  unsigned short ncols_a[MAXL_TC];
  unsigned short ncols_d[MAXL_TC + 1];
};
constexpr int TC_ROWS = 128;                 // trajectories per tile = UMMA M
constexpr int TC_A_SLOT = 32768;             // hi (16 KB) + lo (16 KB) of a 128 x 32 A chunk
constexpr int TC_B_SLOT = 32768;             // one weight unit
constexpr int TC_NB = 4;                     // weight ring depth (units) = 2 K-chunks of prefetch
#ifndef B200_TC_WORKERS
#define B200_TC_WORKERS 256
#endif
constexpr int TC_WORKERS = B200_TC_WORKERS;  // epilogue/producer threads: the 256 solver threads (+ 8 helper warps)
constexpr int TC_CW = 32 / (TC_WORKERS / 128);   // columns of a 32-wide chunk owned by one worker thread
constexpr int TC_NQ = TC_CW / 4;
static_assert(TC_WORKERS == 256 || TC_WORKERS == 512, "worker threads: 8 or 16 warps");
constexpr int TC_THREADS = TC_WORKERS + 128; // worker warps + one control warpgroup (warps 8..11: MMA issuer, loader, 2 spare)
// Register reallocation (setmaxnreg): the kernel is launched at the uniform cap 65536 / TC_THREADS; the control
// warpgroup then drops to TC_REGS_CTRL and the worker warpgroups take TC_REGS_WORK each.
#if B200_TC_WORKERS == 256
constexpr int TC_REGS_CTRL = 40, TC_REGS_WORK = 232;    // 128*40 + 256*232 = 64512 <= 65536
#else
constexpr int TC_REGS_CTRL = 32, TC_REGS_WORK = 112;    // launch cap 96: 128*32 + 512*112 = 61440 = 640*96
#endif
constexpr int TC_SMEM = 2 * TC_A_SLOT + TC_NB * TC_B_SLOT + 1024;
constexpr int TC_MAX_OPS = 3 * MAXL_TC * 8;
constexpr int TC_MAX_LOADS = 2 * MAXL_TC * 16 + MAXL_TC * 12;
constexpr int TC_MAX_BIAS = MAXL_TC * 256;
struct TcOp {                // one GEMM chunk for the MMA issuer
  unsigned char kind;        // 0: K-major chunk (A slot filled by the workers + 2 weight units); 1: wgrad pair
  unsigned char first;       // first MMA overwrites the accumulator
  unsigned char abuf, bbuf;  // wgrad operand buffers
  unsigned short n;          // N of the MMA
  unsigned short acc;        // accumulator column
  unsigned char new_b, last_b, layer_end, phase_end;   // phase_end: 1 = last op of the fwd/bwd phase, 2 = last wgrad op
};
struct TcLoad {              // one bulk load (weight unit) for the loader warp
  unsigned int off, bytes;
  unsigned char kind;        // 0: forward weight unit, 1: dgrad (transposed) weight unit
  unsigned char first_f;     // first weight load of an RHS evaluation: previous wgrad phase done, evaluation started
};

struct SolveParams {
  int flavor;
  TcParams tc;
  MlpDesc mlp;
  SmemPlan sp;
  const float* wp;
  const float* wtp;
  int B, NV, TB, ntiles, nctas;
  int D;    // MLP input/output width
  int NS;   // forward state width per trajectory (D, or D+1 for FFJORD)
  long long n_norm;   // number of entries in the error norm (global batch x NV (+ P))
  float t0, t1, abstol, reltol, dt0_user;
  int maxiters;
  // integrator ring: vector v (0 = u, 1..6 = k2..k7) of slot s
  float* ring;
  int nslots;
  long long vec_elems;
  float* k1_first;
  float* final_state;  // [B][NV]
  // forward
  const float* u0;
  int d_in;
  const float* saveat;
  int nsave, save_start, save_end;
  float* out_u;   // [n_out][B][NV]
  float* out_t;
  float* step_t;  // [tape_cap + 1] knot times of the dense tape
  int dense;      // record the dense tape (ring has tape_cap + 1 slots)
  int tape_cap;
  // adjoint
  const float* f_ring;
  int f_nslots, f_nsteps;
  long long f_vec_elems;
  const float* f_k1_first;
  const float* f_step_t;
  const float* f_us;     // forward saved states [n_ts][B][NS]
  const float* f_final;  // forward end state [B][NS]
  const float* f_ts;
  int n_ts;
  const float* dL;       // [n_ts][B][NS]
  float* slabs;          // [nctas][4][ppad]
  float* mu;             // [2][ppad]
  float* mu_out;         // [ppad]
  int exact_trace;       // FFJORD monte_carlo = false: trace of the Jacobian by D unit-vector passes instead of one noise pass
  int rowthread;         // FFJORD forward, all widths <= 64: one thread per trajectory (rhs_ffjord_fwd_rows)
  int regularize;        // FFJORD regularize = true: state rows D+1, D+2 integrate sum(abs2, m(z)) and ||e^T J||_2 (src/ffjord.jl:131-133)
  int mu_in_norm;        // 1: mu is a state of the adjoint ODE (Interpolating / Backsolve); 0: quadrature (Quadrature / Gauss adjoints)
  // FFJORD
  unsigned long long seed;
  int noise_per_call;
  int b_offset;
  // multi-GPU
  int rank, world;
  unsigned long long xbase;   // solve sequence number << 24: first exchange epoch of this solve - 1
  Mailbox* peers[MAX_RANKS];
  Ctrl* ctrl;
  float* trace;      // optional [trace_cap][4] = (t, dt, EEst, accepted) per attempted step
  int trace_cap;
};

// ------------------------------------------------------------------------------------ tableau
struct Tableau {
  float c[6];
  float a[7][6];   // a[s-1][j-1] for stage s = 2..7
  float bt[7];
  float r[7][4];
};
static __constant__ Tableau c_tab;   // one copy per translation unit (see b200node.cu / tc_kernels.cu)

static inline Tableau host_tableau() {
  Tableau t{};
  const double c[6] = {0.161, 0.327, 0.9, 0.9800255409045097, 1.0, 1.0};
  for (int i = 0; i < 6; ++i) t.c[i] = (float)c[i];
  const double a[7][6] = {
      {0, 0, 0, 0, 0, 0},
      {0.161, 0, 0, 0, 0, 0},
      {-0.008480655492356989, 0.335480655492357, 0, 0, 0, 0},
      {2.8971530571054935, -6.359448489975075, 4.3622954328695815, 0, 0, 0},
      {5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525, 0, 0},
      {5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401,
       -0.028269050394068383, 0},
      {0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081,
       2.324710524099774}};
  for (int i = 0; i < 7; ++i)
    for (int j = 0; j < 6; ++j) t.a[i][j] = (float)a[i][j];
  const double bt[7] = {-0.00178001105222577714, -0.0008164344596567469, 0.007880878010261995,
                        -0.1447110071732629,     0.5823571654525552,     -0.45808210592918697,
                        0.015151515151515152};
  for (int i = 0; i < 7; ++i) t.bt[i] = (float)bt[i];
  const double r[7][4] = {{1.0, -2.763706197274826, 2.9132554618219126, -1.0530884977290216},
                          {0.0, 0.13169999999999998, -0.2234, 0.1017},
                          {0.0, 3.9302962368947516, -5.941033872131505, 2.490627285651253},
                          {0.0, -12.411077166933676, 30.33818863028232, -16.548102889244902},
                          {0.0, 37.50931341651104, -88.1789048947664, 47.37952196281928},
                          {0.0, -27.896526289197286, 65.09189467479366, -34.87065786149661},
                          {0.0, 1.5, -4.0, 2.5}};
  for (int i = 0; i < 7; ++i)
    for (int j = 0; j < 4; ++j) t.r[i][j] = (float)r[i][j];
  return t;
}

// ------------------------------------------------------------------------------------ small helpers
__device__ __forceinline__ float act_apply(int a, float x) {
  switch (a) {
    case ACT_TANH: return tanhf(x);
    case ACT_RELU: return fmaxf(x, 0.f);
    case ACT_SIGMOID: return 1.f / (1.f + expf(-x));
    case ACT_SOFTPLUS: return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x)));
    default: return x;
  }
}
// Branch-free tanh for the tensor-core epilogues (accurate tanhf costs as much as the GEMMs there):
// |x| < 0.35: odd Taylor polynomial to x^11 (truncation < 2e-8 relative); otherwise 1 - 2/(2^(2x log2 e) + 1)
// with MUFU ex2/rcp.  Max error ~4 ulp (3e-7 relative) over the whole range.
__device__ __forceinline__ float tanh_fast(float x) {
  const float t = fabsf(x);
  const float x2 = x * x;
  float p = fmaf(x2, -8.8632355e-3f, 2.1869488e-2f);
  p = fmaf(x2, p, -5.3968254e-2f);
  p = fmaf(x2, p, 1.3333334e-1f);
  p = fmaf(x2, p, -3.3333334e-1f);
  p = fmaf(x * x2, p, x);
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t * 2.8853900817779268f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.f));
  const float b = copysignf(fmaf(-2.f, r, 1.f), x);
  return t < 0.35f ? p : b;
}
// Same function with ONE MUFU op per value: the reciprocal of e + 1 (>= 2) by three Newton steps from a bit-trick seed
// (12 % -> 1.4e-2 -> 2e-4 -> 4e-8 relative) instead of MUFU.RCP.  For the tensor-core engine's epilogues, where the two
// worker warps of a scheduler hit the 4-lane MUFU unit at the same moment (-DB200_TC_TANH_NR).
__device__ __forceinline__ float tanh_fast_nr(float x) {
  const float t = fminf(fabsf(x), 10.f);          // tanh(10) rounds to 1; keeps e + 1 finite
  const float x2 = x * x;
  float p = fmaf(x2, -8.8632355e-3f, 2.1869488e-2f);
  p = fmaf(x2, p, -5.3968254e-2f);
  p = fmaf(x2, p, 1.3333334e-1f);
  p = fmaf(x2, p, -3.3333334e-1f);
  p = fmaf(x * x2, p, x);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t * 2.8853900817779268f));
  const float d = e + 1.f;
  float r = __int_as_float(0x7EF311C7 - __float_as_int(d));
  r = r * fmaf(-d, r, 2.f);
  r = r * fmaf(-d, r, 2.f);
  r = r * fmaf(-d, r, 2.f);
  const float b = copysignf(fmaf(-2.f, r, 1.f), x);
  return t < 0.35f ? p : b;
}
__device__ __forceinline__ float act_apply_fast(int a, float x) { return a == ACT_TANH ? tanh_fast(x) : act_apply(a, x); }
// act'(pre) written in terms of h = act(pre)
__device__ __forceinline__ float act_d1(int a, float h) {
  switch (a) {
    case ACT_TANH: return 1.f - h * h;
    case ACT_RELU: return h > 0.f ? 1.f : 0.f;
    case ACT_SIGMOID: return h * (1.f - h);
    case ACT_SOFTPLUS: return 1.f - expf(-h);
    default: return 1.f;
  }
}
__device__ __forceinline__ float act_d2(int a, float h) {
  switch (a) {
    case ACT_TANH: return -2.f * h * (1.f - h * h);
    case ACT_SIGMOID: return h * (1.f - h) * (1.f - 2.f * h);
    case ACT_SOFTPLUS: { float s = 1.f - expf(-h); return s * (1.f - s); }
    default: return 0.f;
  }
}

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int ld_volatile_i32(const int* p) {
  return *reinterpret_cast<const volatile int*>(p);
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Philox4x32-10 (same stream as oracle/philox.py)
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const unsigned int hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const unsigned int hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}
__device__ __forceinline__ float u01(unsigned int x) {
  return ((float)(x >> 8) + 0.5f) * 5.9604644775390625e-08f;  // 2^-24
}
// four standard normals for (trajectory b, dim quad q, rhs call, stream)
__device__ __forceinline__ float4 philox_normal4(unsigned long long seed, unsigned int stream,
                                                 unsigned int call, unsigned int b, unsigned int q) {
  const uint4 x = philox4x32_10(make_uint4(b, q, call, stream),
                                make_uint2((unsigned int)seed, (unsigned int)(seed >> 32)));
  float4 o;
  float r = sqrtf(-2.f * logf(u01(x.x)));
  float s, c;
  sincosf(6.283185307179586f * u01(x.y), &s, &c);
  o.x = r * c;
  o.y = r * s;
  r = sqrtf(-2.f * logf(u01(x.z)));
  sincosf(6.283185307179586f * u01(x.w), &s, &c);
  o.z = r * c;
  o.w = r * s;
  return o;
}

// DiffEqBase fastpow (Float32): exp2(y * fastlog2(x)); explicit _rn ops so nvcc cannot contract.
__device__ __forceinline__ float fastlog2f(float x) {
  const unsigned int ux = __float_as_uint(x);
  const unsigned int e = (ux & 0x7F800000u) >> 23;
  float sig, fexp;
  if (ux & 0x00400000u) {
    sig = __uint_as_float((ux & 0x007FFFFFu) | 0x3F000000u);
    fexp = __fsub_rn((float)e, 126.0f);
  } else {
    sig = __uint_as_float((ux & 0x007FFFFFu) | 0x3F800000u);
    fexp = __fsub_rn((float)e, 127.0f);
  }
  const float s = __fsub_rn(sig, 1.0f);
  const float num = __fmul_rn(s, __fadd_rn(__fmul_rn(0.338953f, s), 2.198599f));
  return __fadd_rn(fexp, __fdiv_rn(num, __fadd_rn(s, 1.523692f)));
}
__device__ __forceinline__ float fastpowf(float x, float y) {
  if (x == 0.f) return 0.f;
  return exp2f(__fmul_rn(y, fastlog2f(x)));
}
__device__ __forceinline__ float eps32(float x) {
  x = fabsf(x);
  if (x < 1.17549435e-38f) return 1.4e-45f;
  const float nxt = __uint_as_float(__float_as_uint(x) + 1u);
  return nxt - x;
}

__device__ __forceinline__ int pow2_quads(int nq) {  // smallest power of two >= nq, capped at 32
  int qn = 1;
  while (qn < nq && qn < 32) qn <<= 1;
  return qn;
}

// ------------------------------------------------------------------------------------ tile GEMM
// C[r][n] = sum_k A[r][k] * W[k][n];  A: shared [TB][lda] (pad columns zero), W: global row-major
// [Kpad][ldw].  Each thread owns an 8-row x 4-column register tile; a warp spans (32/qn)*8 rows and
// qn column quads.  epi(row0, col0, acc) receives the finished 8x4 block.
template <typename Epi>
__device__ __forceinline__ void tile_gemm(const float* __restrict__ A, int lda,
                                          const float* __restrict__ W, int ldw, int TB, int Kpad,
                                          int Npad, Epi&& epi) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nq = Npad >> 2;
  const int qn = pow2_quads(nq);
  const int rsub = lane / qn, ql = lane - rsub * qn;
  const int rows_per_warp = (32 / qn) * 8;
  const int rows_per_pass = rows_per_warp * NWARPS;
  for (int rb = 0; rb < TB; rb += rows_per_pass) {
    const int row0 = rb + warp * rows_per_warp + rsub * 8;
    for (int q0 = 0; q0 < nq; q0 += qn) {
      const int quad = q0 + ql;
      if (row0 < TB && quad < nq) {
        float acc[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
        const float* ap = A + row0 * lda;
        const float* wp = W + quad * 4;
#pragma unroll 2
        for (int k = 0; k < Kpad; k += 4) {
          const float4 w0 = __ldg(reinterpret_cast<const float4*>(wp + (size_t)(k + 0) * ldw));
          const float4 w1 = __ldg(reinterpret_cast<const float4*>(wp + (size_t)(k + 1) * ldw));
          const float4 w2 = __ldg(reinterpret_cast<const float4*>(wp + (size_t)(k + 2) * ldw));
          const float4 w3 = __ldg(reinterpret_cast<const float4*>(wp + (size_t)(k + 3) * ldw));
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 a = *reinterpret_cast<const float4*>(ap + i * lda + k);
            acc[i][0] = fmaf(a.x, w0.x, acc[i][0]);
            acc[i][1] = fmaf(a.x, w0.y, acc[i][1]);
            acc[i][2] = fmaf(a.x, w0.z, acc[i][2]);
            acc[i][3] = fmaf(a.x, w0.w, acc[i][3]);
            acc[i][0] = fmaf(a.y, w1.x, acc[i][0]);
            acc[i][1] = fmaf(a.y, w1.y, acc[i][1]);
            acc[i][2] = fmaf(a.y, w1.z, acc[i][2]);
            acc[i][3] = fmaf(a.y, w1.w, acc[i][3]);
            acc[i][0] = fmaf(a.z, w2.x, acc[i][0]);
            acc[i][1] = fmaf(a.z, w2.y, acc[i][1]);
            acc[i][2] = fmaf(a.z, w2.z, acc[i][2]);
            acc[i][3] = fmaf(a.z, w2.w, acc[i][3]);
            acc[i][0] = fmaf(a.w, w3.x, acc[i][0]);
            acc[i][1] = fmaf(a.w, w3.y, acc[i][1]);
            acc[i][2] = fmaf(a.w, w3.z, acc[i][2]);
            acc[i][3] = fmaf(a.w, w3.w, acc[i][3]);
          }
        }
        epi(row0, quad * 4, acc);
      }
    }
  }
}

// G[k][n] = sum_r A1[r][k] * D1[r][n] (+ sum_r A2[r][k] * D2[r][n]); thread tile 8 k x 4 n.
template <typename Out>
__device__ __forceinline__ void tile_wgrad(const float* __restrict__ A1, const float* __restrict__ D1,
                                           const float* __restrict__ A2, const float* __restrict__ D2,
                                           int lda, int ldd, int TB, int Kpad, int Npad, Out&& out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nq = Npad >> 2;
  const int qn = pow2_quads(nq);
  const int ksub = lane / qn, ql = lane - ksub * qn;
  const int kg_per_warp = 32 / qn;
  const int k_per_pass = kg_per_warp * NWARPS * 8;
  for (int kb = 0; kb < Kpad; kb += k_per_pass) {
    const int k0 = kb + (warp * kg_per_warp + ksub) * 8;
    for (int q0 = 0; q0 < nq; q0 += qn) {
      const int quad = q0 + ql;
      if (k0 < Kpad && quad < nq) {
        const bool hi = (k0 + 4) < Kpad;
        float acc[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
        for (int pass = 0; pass < 2; ++pass) {
          const float* A = pass ? A2 : A1;
          const float* D = pass ? D2 : D1;
          if (A == nullptr) break;
          const float* ap = A + k0;
          const float* dp = D + quad * 4;
#pragma unroll 4
          for (int r = 0; r < TB; ++r) {
            const float4 d = *reinterpret_cast<const float4*>(dp + r * ldd);
            const float4 a0 = *reinterpret_cast<const float4*>(ap + r * lda);
            float4 a1 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (hi) a1 = *reinterpret_cast<const float4*>(ap + r * lda + 4);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              acc[i][0] = fmaf(av[i], d.x, acc[i][0]);
              acc[i][1] = fmaf(av[i], d.y, acc[i][1]);
              acc[i][2] = fmaf(av[i], d.z, acc[i][2]);
              acc[i][3] = fmaf(av[i], d.w, acc[i][3]);
            }
          }
        }
        out(k0, quad * 4, acc, hi ? 8 : 4);
      }
    }
  }
}

// ------------------------------------------------------------------------------------ kernel state
struct InterpInfo {
  int slot;
  float dts;
  float w[7];
};

struct SlabOut {       // what a wgrad tile does with G (already signed)
  float* store;        // G -> store (plain)
  float* accA; float ca;
  float* accE; float ce;
};

struct TcArgs {        // one stage evaluation, posted by thread 0 for the helper worker warps
  const float* u;
  const float* kp[6];
  float coef[6];
  float* dst;
  float* kout;
  SlabOut so;
  InterpInfo ii;
  float dt;
  int tile, nk, acc_store, cmd;   // cmd: 1 = evaluate, 2 = exit
};

struct StepShared {
  float t, dt, tnew;
  float qold, q11;
  float EEst;
  int accept, done, retcode;
  int naccept, nreject, nf, iter;
  int step;        // index of the current step (= ring slot holding u)
  int first_step;  // step whose k1 is k1_first
  int si;          // next saveat / output index
  int ti;          // next tstop index (adjoint)
  int mu_cur;
  int g1;          // which of slabs 2/3 holds G1 (the mu-derivative of the FSAL stage)
  int aborted;
  int out_idx;
  float tie;
  unsigned int epoch;
  unsigned long long xepoch;   // last completed scalar exchange epoch
  unsigned long long mu_xe;    // last completed mu-slice exchange epoch
  unsigned int rhs_calls;
  InterpInfo ii[8];
  double red[NWARPS][2];
  double sums[2];
  // tensor-core engine: mbarriers
  unsigned long long a_full[2];          // (TC_WORKERS arrivals) the workers filled A slot s
  unsigned long long mma_done[2];        // MMAs that read A slot s have completed
  // -DB200_TC_HALF_BARRIERS: the same pair per 16-column half of a slot (half h is filled by worker warpgroup h and read by
  // the K16 group hh = h of the chunk's MMAs): the two warpgroups stop waiting for each other and run half a chunk apart
  unsigned long long a_fullh[2][2];
  unsigned long long mma_doneh[2][2];
  unsigned long long unit_done[TC_NB];   // MMAs that read weight slot s have completed
  unsigned long long b_full[TC_NB];      // weight unit landed in slot s
  unsigned long long w_full[4];          // (TC_WORKERS arrivals) wgrad operand buffer filled by the workers (A0, A1, B0, B1)
  unsigned long long w_done[4];          // MMAs that read that buffer have completed
  unsigned long long w_all_done;         // wgrad MMAs of this RHS complete
  unsigned long long wacc_done;          // a wgrad layer's accumulators are complete
  unsigned long long mu_full[2];         // mu phase: slab slices staged in the operand buffers landed
  unsigned long long rhs_go;             // (TC_WORKERS arrivals) a new RHS evaluation starts: the loader may fetch its first units
  unsigned long long wacc_free;          // (TC_WORKERS arrivals) ... and have been read
  int tc_exit;                           // workers -> control warps: the solve is over
  TcArgs targs;
  unsigned int tmem_base;
  int n_ops, n_loads;
  TcOp ops[TC_MAX_OPS];
  TcLoad loads[TC_MAX_LOADS];
  float bias_acc[4][TC_MAX_BIAS];        // per TMEM quadrant partial column sums of the cotangents
  alignas(16) float tc_bias[TC_MAX_BIAS]; // the layers' biases back to back (layer l at sum_{i<l} dims[i+1]): read once per chunk by
                                         // every worker, on the MMA's critical path (a global load there is an L2 round trip)
};

template <int FL>
struct Traits {
  static constexpr bool adj = (FL >= FL_ADJ_NODE_INTERP);
  static constexpr bool interp = (FL == FL_ADJ_NODE_INTERP || FL == FL_ADJ_FFJORD);
  static constexpr bool ffjord = (FL == FL_FWD_FFJORD || FL == FL_ADJ_FFJORD);
};

template <int FL, bool TC = false>
struct Solver {
  const SolveParams& P;
  float* sm;
  StepShared& ss;
  const int tid, cta;
  unsigned int tc_g = 0;   // running chunk counter of the tensor-core pipeline (uniform across threads)
  unsigned int tc_wl = 0;  // wgrad layers completed so far (worker side)
  unsigned int tc_wused = 0u, tc_wpar = 0u;   // wgrad operand buffers A0, A1, B0, B1: ever filled / parity of the fill count (bit per buffer)
#ifdef B200_TC_TRACE
  unsigned int tr_eval = 0u;
#define TR_W(on, row, col) { if (on) b200tc::g_tc_trace[0][(row)][(col)] = clock64(); }
#else
#define TR_W(on, row, col) {}
#endif
#ifdef B200_TC_PROF
  long long pf[20] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#define PF_T(i, ...) { long long t_ = clock64(); __VA_ARGS__; pf[i] += clock64() - t_; }
#else
#define PF_T(i, ...) { __VA_ARGS__; }
#endif

  __device__ Solver(const SolveParams& p, float* s, StepShared& st)
      : P(p), sm(s), ss(st), tid(threadIdx.x), cta(blockIdx.x) {}

  // CTA barrier of the NTHREADS solver threads (the tensor-core kernels run two extra control warps)
  __device__ __forceinline__ void sync() const {
    if constexpr (TC) asm volatile("bar.sync 1, 256;" ::: "memory");
    else __syncthreads();
  }

  // ---------------------------------------------------------------- pointers
  __device__ __forceinline__ float* slotv(int slot, int v) const {
    return P.ring + ((long long)(slot % P.nslots) * 7 + v) * P.vec_elems;
  }
  __device__ __forceinline__ const float* k1ptr(int step) const {
    return step == ss.first_step ? P.k1_first : slotv(step - 1, 6);
  }
  __device__ __forceinline__ float* k1ptr_w(int step) const {
    return step == ss.first_step ? P.k1_first : slotv(step - 1, 6);
  }
  __device__ __forceinline__ const float* fslotv(int slot, int v) const {
    return P.f_ring + ((long long)(slot % P.f_nslots) * 7 + v) * P.f_vec_elems;
  }
  __device__ __forceinline__ float* slab(int which) const {   // 0..4: G of stages 2..6, 5/6: G1 / G7 (swapped)
    return P.slabs + ((long long)cta * NSLAB + which) * P.mlp.ppad;
  }

  // ---------------------------------------------------------------- barriers / reductions
  __device__ void fail(int code) {
    if (tid == 0) {
      atomicCAS(&P.ctrl->retcode, 0, code);
      atomicExch(&P.ctrl->abort_flag, 1);
    }
  }

  // Grid barrier; thread 0 publishes this CTA's two partial sums first.  Afterwards ss.sums holds
  // the batch-wide (and cross-rank) totals, identical on every CTA and rank.
  // `cross`: also sum over the ranks (one NVLink round trip); a barrier that only orders this GPU's CTAs skips it
  __device__ void barrier_reduce(double p0, double p1, bool cross = true) {
    // block reduce (deterministic order)
    p0 = warp_sum_d(p0);
    p1 = warp_sum_d(p1);
    if ((tid & 31) == 0) {
      ss.red[tid >> 5][0] = p0;
      ss.red[tid >> 5][1] = p1;
    }
    sync();
    Ctrl* c = P.ctrl;
    const unsigned int par = ss.epoch & 1u;
    if (tid == 0) {
      double s0 = 0, s1 = 0;
      for (int w = 0; w < NWARPS; ++w) {
        s0 += ss.red[w][0];
        s1 += ss.red[w][1];
      }
      c->partial[par][cta][0] = s0;
      c->partial[par][cta][1] = s1;
      __threadfence();
      atomicAdd(&c->bar, 1u);
      const unsigned int target = (ss.epoch + 1u) * (unsigned int)P.nctas;
      long long t_start = clock64();
      while (ld_acquire_u32(&c->bar) < target) {
        if (ld_volatile_i32(&c->abort_flag)) break;
        if (clock64() - t_start > 40000000000LL) {   // ~20 s: a peer CTA never arrived
          atomicCAS(&c->retcode, 0, -5);
          atomicExch(&c->abort_flag, 1);
          break;
        }
        __nanosleep(20);
      }
      __threadfence();
    }
    sync();
    // every CTA sums all partials in the same order
    if (tid < 32) {
      double s0 = 0, s1 = 0;
      for (int i = tid; i < P.nctas; i += 32) {
        s0 += __ldcg(&c->partial[par][i][0]);
        s1 += __ldcg(&c->partial[par][i][1]);
      }
      s0 = warp_sum_d(s0);
      s1 = warp_sum_d(s1);
      if (tid == 0) {
        ss.sums[0] = s0;
        ss.sums[1] = s1;
      }
    }
    sync();
    if (P.world > 1 && cross) cross_rank_exchange();
    if (tid == 0) {
      ss.epoch += 1u;
      ss.aborted = ld_volatile_i32(&c->abort_flag);
    }
    sync();
  }

  // Per-step exchange of the local sums over NVLink peer memory: CTA 0 pushes {value, tag} into
  // every rank's mailbox (its own included), waits for all ranks' tags, adds the values in rank
  // order and publishes the total to the other CTAs of this GPU.
  __device__ void cross_rank_exchange() {
    Ctrl* c = P.ctrl;
    const unsigned long long xe = ss.xepoch + 1ull;
    const unsigned int par = (unsigned int)(xe & 1ull);
    if (cta == 0) {
      if (tid < P.world) {
        Mailbox* mb = P.peers[tid];
        mb->val[par][0][P.rank] = ss.sums[0];
        mb->val[par][1][P.rank] = ss.sums[1];
        __threadfence_system();
        mb->tag[par][P.rank] = xe;
      }
      sync();
      if (tid == 0) {
        Mailbox* me = P.peers[P.rank];
        double s0 = 0, s1 = 0;
        long long t_start = clock64();
        for (int r = 0; r < P.world; ++r) {
          while (me->tag[par][r] < xe) {
            if (ld_volatile_i32(&c->abort_flag)) break;
            if (clock64() - t_start > 40000000000LL) {
              atomicCAS(&c->retcode, 0, -5);
              atomicExch(&c->abort_flag, 1);
              break;
            }
          }
          __threadfence_system();
          s0 += me->val[par][0][r];
          s1 += me->val[par][1][r];
        }
        c->xsum[par][0] = s0;
        c->xsum[par][1] = s1;
        __threadfence();
        atomicExch(&c->xflag, xe);
      }
    }
    if (tid == 0) {
      long long t_start = clock64();
      while (ld_acquire_u64(&c->xflag) < xe) {
        if (ld_volatile_i32(&c->abort_flag)) break;
        if (clock64() - t_start > 40000000000LL) {
          atomicCAS(&c->retcode, 0, -5);
          atomicExch(&c->abort_flag, 1);
          break;
        }
        __nanosleep(20);
      }
      ss.sums[0] = __ldcg(&c->xsum[par][0]);
      ss.sums[1] = __ldcg(&c->xsum[par][1]);
      ss.xepoch = xe;
    }
    sync();
  }

  __device__ __forceinline__ bool aborted() const { return ss.aborted != 0; }   // CTA-uniform

  // ---------------------------------------------------------------- shared-memory routing
  // element (row r, state column c) of a stage input goes to the buffers the RHS reads
  __device__ __forceinline__ void route(int r, int c, float v) {
    const MlpDesc& M = P.mlp;
    const int D = P.D;
    if constexpr (FL == FL_FWD_NODE) {
      put_input(r, c, v);
    } else if constexpr (FL == FL_FWD_FFJORD) {
      if (c < D) put_input(r, c, v);
    } else if constexpr (FL == FL_ADJ_NODE_INTERP) {
      sm[P.sp.dl_off + r * M.dpad[M.nl] + c] = v;
    } else if constexpr (FL == FL_ADJ_NODE_BACKSOLVE) {
      if (c < D) sm[P.sp.dl_off + r * M.dpad[M.nl] + c] = v;
      else put_input(r, c - D, v);
    } else {  // FL_ADJ_FFJORD
      if (c < D) sm[P.sp.dl_off + r * M.dpad[M.nl] + c] = v;
      else sm[P.sp.ll_off + (c - D) * P.TB + r] = v;     // lambda of delta_logp [, of the two regulariser rows]
    }
  }
  __device__ __forceinline__ void put_input(int r, int c, float v) {
    const MlpDesc& M = P.mlp;
    if (M.pre == 1) {
      sm[P.sp.x_off + r * M.dpad[0] + c] = v;
      v = v * v * v;
    }
    sm[P.sp.act_off[0] + r * M.dpad[0] + c] = v;
  }

  // stage input x = u + dt * sum_j coef[j] * k_j for this tile; optionally also stored to `dst`
  __device__ void build_stage(int tile, const float* u, const float* const* kp, const float* coef,
                              int nk, float dt, float* dst) {
    const int NV = P.NV;
    const int base_row = tile * P.TB;
    const int n = P.TB * NV;
    const long long gbase = (long long)base_row * NV;
    for (int e = tid; e < n; e += NTHREADS) {
      const int r = e / NV, c = e - r * NV;
      float v = 0.f;
      if (base_row + r < P.B) {
        const long long g = gbase + e;
        v = u[g];
        if (nk > 0) {
          float acc = coef[0] * kp[0][g];
          for (int j = 1; j < nk; ++j) acc = fmaf(coef[j], kp[j][g], acc);
          v = fmaf(dt, acc, v);
        }
        if (dst) dst[g] = v;
      }
      route(r, c, v);
    }
  }

  // y(t) from the forward dense tape into the MLP input buffer (Interpolating flavours)
  __device__ void build_y_interp(int tile, const InterpInfo& ii) {
    const int NS = P.NS, D = P.D;
    const int base_row = tile * P.TB;
    const int n = P.TB * NS;
    const long long gbase = (long long)base_row * NS;
    const float* u = fslotv(ii.slot, 0);
    const float* k1 = ii.slot == 0 ? P.f_k1_first : fslotv(ii.slot - 1, 6);
    const float* kk[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) kk[j] = fslotv(ii.slot, 1 + j);
    for (int e = tid; e < n; e += NTHREADS) {
      const int r = e / NS, c = e - r * NS;
      if (c >= D) continue;
      float v = 0.f;
      if (base_row + r < P.B) {
        const long long g = gbase + e;
        float acc = ii.w[0] * k1[g];
#pragma unroll
        for (int j = 0; j < 6; ++j) acc = fmaf(ii.w[1 + j], kk[j][g], acc);
        v = fmaf(ii.dts, acc, u[g]);
      }
      put_input(r, c, v);
    }
  }

  // thread 0: locate t on the forward tape and fill the interpolation weights
  __device__ void locate(float t, InterpInfo& ii) {
    const float* st = P.f_step_t;
    int lo = 0, hi = P.f_nsteps - 1;
    const bool fwd = st[P.f_nsteps] >= st[0];
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      const float e = st[mid + 1];
      if (fwd ? (e >= t) : (e <= t)) hi = mid; else lo = mid + 1;
    }
    const float ta = st[lo], tb = st[lo + 1];
    const float dts = __fsub_rn(tb, ta);
    const float th = __fdiv_rn(__fsub_rn(t, ta), dts);
    ii.slot = lo;
    ii.dts = dts;
    interp_weights(th, ii.w);
  }
  __device__ static void interp_weights(float th, float* w) {
#pragma unroll
    for (int i = 0; i < 7; ++i) {
      const float* r = c_tab.r[i];
      const float inner = __fadd_rn(r[1], __fmul_rn(th, __fadd_rn(r[2], __fmul_rn(th, r[3]))));
      if (i == 0) w[i] = __fmul_rn(th, __fadd_rn(r[0], __fmul_rn(th, inner)));
      else w[i] = __fmul_rn(__fmul_rn(th, th), inner);
    }
  }

  // ---------------------------------------------------------------- MLP pieces
  // forward through all layers; hidden activations stay in shared memory; the last layer's 8x4
  // blocks go to fin(row, col, float4).
  template <typename Fin>
  __device__ void mlp_forward(Fin&& fin) {
    const MlpDesc& M = P.mlp;
    for (int l = 0; l < M.nl; ++l) {
      const float* A = sm + P.sp.act_off[l];
      const float* W = P.wp + M.w_off[l];
      const float* bias = P.wp + M.b_off[l];
      const int Kp = M.dpad[l], Np = M.dpad[l + 1], N = M.dims[l + 1], a = M.act[l];
      if (l < M.nl - 1) {
        float* O = sm + P.sp.act_off[l + 1];
        tile_gemm(A, Kp, W, Np, P.TB, Kp, Np, [&](int row0, int col0, float (&acc)[8][4]) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(bias + col0));
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float4 v;
            v.x = (col0 + 0 < N) ? act_apply(a, acc[i][0] + b.x) : 0.f;
            v.y = (col0 + 1 < N) ? act_apply(a, acc[i][1] + b.y) : 0.f;
            v.z = (col0 + 2 < N) ? act_apply(a, acc[i][2] + b.z) : 0.f;
            v.w = (col0 + 3 < N) ? act_apply(a, acc[i][3] + b.w) : 0.f;
            *reinterpret_cast<float4*>(O + (row0 + i) * Np + col0) = v;
          }
        });
      } else {
        tile_gemm(A, Kp, W, Np, P.TB, Kp, Np, [&](int row0, int col0, float (&acc)[8][4]) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(bias + col0));
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float4 v;
            v.x = act_apply(a, acc[i][0] + b.x);
            v.y = act_apply(a, acc[i][1] + b.y);
            v.z = act_apply(a, acc[i][2] + b.z);
            v.w = act_apply(a, acc[i][3] + b.w);
            fin(row0 + i, col0, v);
          }
        });
      }
      sync();
    }
  }

  __device__ __forceinline__ void slab_apply(const SlabOut& so, int idx, float4 g) {
    if (so.store) *reinterpret_cast<float4*>(so.store + idx) = g;
    if (so.accA) {
      float4 v = *reinterpret_cast<float4*>(so.accA + idx);
      v.x = fmaf(so.ca, g.x, v.x); v.y = fmaf(so.ca, g.y, v.y);
      v.z = fmaf(so.ca, g.z, v.z); v.w = fmaf(so.ca, g.w, v.w);
      *reinterpret_cast<float4*>(so.accA + idx) = v;
    }
    if (so.accE) {
      float4 v = *reinterpret_cast<float4*>(so.accE + idx);
      v.x = fmaf(so.ce, g.x, v.x); v.y = fmaf(so.ce, g.y, v.y);
      v.z = fmaf(so.ce, g.z, v.z); v.w = fmaf(so.ce, g.w, v.w);
      *reinterpret_cast<float4*>(so.accE + idx) = v;
    }
  }

  // weight + bias gradient of layer li from (A1,D1) (+ (A2,D2)); G = -sum (adjoint sign) goes to
  // this CTA's slabs.  `first_tile` false => the slabs already hold other tiles' contribution.
  __device__ void layer_wgrad(int li, const float* A1, const float* D1, const float* A2,
                              const float* D2, const SlabOut& so, bool accumulate_store) {
    const MlpDesc& M = P.mlp;
    const int Kp = M.dpad[li], Np = M.dpad[li + 1];
    SlabOut s2 = so;
    if (accumulate_store && so.store) {   // further tiles of this CTA add to the stored G
      s2.store = nullptr;
    }
    tile_wgrad(A1, D1, A2, D2, Kp, Np, P.TB, Kp, Np,
               [&](int k0, int col0, float (&acc)[8][4], int nk) {
                 for (int i = 0; i < nk; ++i) {
                   const int idx = M.w_off[li] + (k0 + i) * Np + col0;
                   const float4 g = make_float4(-acc[i][0], -acc[i][1], -acc[i][2], -acc[i][3]);
                   if (accumulate_store && so.store) {
                     float4 v = *reinterpret_cast<float4*>(so.store + idx);
                     v.x += g.x; v.y += g.y; v.z += g.z; v.w += g.w;
                     *reinterpret_cast<float4*>(so.store + idx) = v;
                   }
                   slab_apply(s2, idx, g);
                 }
               });
    // bias gradient: column sums of D (D2 carries no bias term)
    for (int n = tid; n < Np; n += NTHREADS) {
      float s = 0.f;
      for (int r = 0; r < P.TB; ++r) s += D1[r * Np + n];
      const int idx = M.b_off[li] + n;
      const float g = -s;
      if (so.store) so.store[idx] = (accumulate_store ? so.store[idx] : 0.f) + g;
      if (so.accA) so.accA[idx] = fmaf(so.ca, g, so.accA[idx]);
      if (so.accE) so.accE[idx] = fmaf(so.ce, g, so.accE[idx]);
    }
  }

  // reverse pass of a plain chain: seed delta_L in dl buffer; hidden deltas overwrite the
  // activations in place; optional weight gradients; the input cotangent goes to fin(row,col,float4)
  template <bool WGRAD, typename Fin>
  __device__ void mlp_backward(const SlabOut& so, bool accumulate_store, Fin&& fin) {
    const MlpDesc& M = P.mlp;
    for (int l = M.nl; l >= 1; --l) {
      const int li = l - 1;
      const float* Dl = (l == M.nl) ? sm + P.sp.dl_off : sm + P.sp.act_off[l];
      float* Aprev = sm + P.sp.act_off[li];
      const int Np = M.dpad[l], Kp = M.dpad[li];
      if constexpr (WGRAD) {
        layer_wgrad(li, Aprev, Dl, nullptr, nullptr, so, accumulate_store);
        sync();
      }
      const float* WT = P.wtp + M.wt_off[li];
      if (li >= 1) {
        const int a = M.act[li - 1], Kd = M.dims[li];
        tile_gemm(Dl, Np, WT, Kp, P.TB, Np, Kp, [&](int row0, int col0, float (&acc)[8][4]) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float4* hp = reinterpret_cast<float4*>(Aprev + (row0 + i) * Kp + col0);
            const float4 h = *hp;
            float4 v;
            v.x = (col0 + 0 < Kd) ? acc[i][0] * act_d1(a, h.x) : 0.f;
            v.y = (col0 + 1 < Kd) ? acc[i][1] * act_d1(a, h.y) : 0.f;
            v.z = (col0 + 2 < Kd) ? acc[i][2] * act_d1(a, h.z) : 0.f;
            v.w = (col0 + 3 < Kd) ? acc[i][3] * act_d1(a, h.w) : 0.f;
            *hp = v;
          }
        });
      } else {
        const float* X = sm + P.sp.x_off;
        const int pre = M.pre;
        tile_gemm(Dl, Np, WT, Kp, P.TB, Np, Kp, [&](int row0, int col0, float (&acc)[8][4]) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float4 v = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
            if (pre == 1) {
              const float4 x = *reinterpret_cast<const float4*>(X + (row0 + i) * Kp + col0);
              v.x *= 3.f * x.x * x.x; v.y *= 3.f * x.y * x.y;
              v.z *= 3.f * x.z * x.z; v.w *= 3.f * x.w * x.w;
            }
            fin(row0 + i, col0, v);
          }
        });
      }
      sync();
    }
  }

  // ---------------------------------------------------------------- noise
  // unit >= 0: the unit vector e_unit instead of noise (exact trace, monte_carlo = false: src/ffjord.jl:120-123)
  __device__ void fill_noise(int tile, unsigned int stream, unsigned int call, float* dst, int ld, int unit = -1) {
    const int D = P.D, nq = (D + 3) >> 2;
    const int base_row = tile * P.TB;
    for (int e = tid; e < P.TB * nq; e += NTHREADS) {
      const int r = e / nq, q = e - r * nq;
      float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      if (unit >= 0) {
        if ((unit >> 2) == q) (&z.x)[unit & 3] = 1.f;
      } else if (base_row + r < P.B) {
        z = philox_normal4(P.seed, stream, P.noise_per_call ? call : 0u,
                           (unsigned int)(P.b_offset + base_row + r), (unsigned int)q);
        if (q * 4 + 1 >= D) z.y = 0.f;
        if (q * 4 + 2 >= D) z.z = 0.f;
        if (q * 4 + 3 >= D) z.w = 0.f;
      }
      *reinterpret_cast<float4*>(dst + r * ld + q * 4) = z;
    }
  }

  // ---------------------------------------------------------------- RHS per flavour
  // Evaluates the RHS for the tile whose stage input is already routed into shared memory and
  // writes the derivative to kout ([B][NV] global vector).  Adjoint flavours also emit the
  // parameter-gradient tile through `so`.
  __device__ void rhs_tile(int tile, float* kout, const SlabOut& so, bool acc_store,
                           unsigned int call) {
    const MlpDesc& M = P.mlp;
    const int D = P.D, NV = P.NV;
    const int base_row = tile * P.TB;
    const int B = P.B;
    sync();   // stage input complete
    if constexpr (FL == FL_FWD_NODE) {
      mlp_forward([&](int row, int col, float4 v) {
        if (base_row + row < B) {
          float* o = kout + (long long)(base_row + row) * NV + col;
          if ((NV & 3) == 0) { *reinterpret_cast<float4*>(o) = v; }
          else {
            if (col + 0 < D) o[0] = v.x;
            if (col + 1 < D) o[1] = v.y;
            if (col + 2 < D) o[2] = v.z;
            if (col + 3 < D) o[3] = v.w;
          }
        }
      });
    } else if constexpr (FL == FL_FWD_FFJORD) {
      if (P.rowthread) {
        rhs_ffjord_fwd_rows(tile, kout, call);
        return;
      }
      const int Lp = M.dpad[M.nl];
      float* mz = sm + P.sp.dl2_off;
      float* e = sm + P.sp.dl_off;
      fill_noise(tile, 0u, call, e, Lp);
      mlp_forward([&](int row, int col, float4 v) {
        *reinterpret_cast<float4*>(mz + row * Lp + col) = v;
      });
      float* eJ = sm + P.sp.act_off[0];
      const int Kp0 = M.dpad[0];
      mlp_backward<false>(so, false, [&](int row, int col, float4 v) {
        *reinterpret_cast<float4*>(eJ + row * Kp0 + col) = v;
      });
      for (int r = tid; r < P.TB; r += NTHREADS) {
        if (base_row + r < B) {
          float tr = 0.f, s1 = 0.f, s2 = 0.f;
          float* o = kout + (long long)(base_row + r) * NV;
          for (int d = 0; d < D; ++d) {
            const float ej = eJ[r * Kp0 + d], m = mz[r * Lp + d];
            tr = fmaf(ej, e[r * Lp + d], tr);
            s1 = fmaf(m, m, s1);
            s2 = fmaf(ej, ej, s2);
            o[d] = m;
          }
          o[D] = -tr;
          if (P.regularize) { o[D + 1] = s1; o[D + 2] = sqrtf(s2); }
        }
      }
      sync();
    } else if constexpr (FL == FL_ADJ_NODE_INTERP || FL == FL_ADJ_NODE_BACKSOLVE) {
      // forward recompute (activations kept), f(y) is the y-part derivative for Backsolve
      mlp_forward([&](int row, int col, float4 v) {
        if constexpr (FL == FL_ADJ_NODE_BACKSOLVE) {
          if (base_row + row < B) {
            float* o = kout + (long long)(base_row + row) * NV + D + col;
            if (col + 0 < D) o[0] = v.x;
            if (col + 1 < D) o[1] = v.y;
            if (col + 2 < D) o[2] = v.z;
            if (col + 3 < D) o[3] = v.w;
          }
        }
      });
      mlp_backward<true>(so, acc_store, [&](int row, int col, float4 v) {
        if (base_row + row < B) {
          float* o = kout + (long long)(base_row + row) * NV + col;
          if (col + 0 < D) o[0] = -v.x;
          if (col + 1 < D) o[1] = -v.y;
          if (col + 2 < D) o[2] = -v.z;
          if (col + 3 < D) o[3] = -v.w;
        }
      });
    } else {  // FL_ADJ_FFJORD
      rhs_ffjord_adjoint(tile, kout, so, acc_store, call);
    }
  }

  // ---------------------------------------------------------------- FFJORD forward, one thread per trajectory
  // Small networks (every width <= 64, e.g. the CNF examples' 2 -> 64 -> 64 -> 2): the 8 x 4 register tiles of tile_gemm
  // leave most threads idle in the narrow first / last layers and pay a CTA barrier per layer.  Here thread t owns
  // trajectory t of the tile: its 64 accumulators live in registers, the weight rows are warp-uniform (broadcast)
  // 16-byte loads through L1, hidden activations go through shared memory column-major ([feature][row]: conflict free),
  // and no barrier separates the layers.  acc[n] (+)= x * W[n .. ] for NQ quads of outputs.
  // (the weights are in shared memory: stage_row_weights; a warp-uniform 16-byte LDS per four FMAs)
  __device__ void stage_row_weights() {
    const MlpDesc& M = P.mlp;
    const int L = M.nl, nwt = M.wt_off[L - 1] + M.dpad[L] * M.dpad[L - 1];
    for (int i = tid; i < M.ppad; i += NTHREADS) sm[P.sp.roww_off + i] = P.wp[i];
    for (int i = tid; i < nwt; i += NTHREADS) sm[P.sp.rowwt_off + i] = P.wtp[i];
    sync();
  }
  template <int NQ>
  __device__ __forceinline__ void row_axpy(float (&acc)[64], float x, const float* __restrict__ wrow) {
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const float4 w = *(reinterpret_cast<const float4*>(wrow) + q);
      acc[4 * q + 0] = fmaf(x, w.x, acc[4 * q + 0]);
      acc[4 * q + 1] = fmaf(x, w.y, acc[4 * q + 1]);
      acc[4 * q + 2] = fmaf(x, w.z, acc[4 * q + 2]);
      acc[4 * q + 3] = fmaf(x, w.w, acc[4 * q + 3]);
    }
  }
  // acc[0 .. 4 nq) += sum_k x_k * W[k][.]: x_k = xs[k * xstride], W row-major with ld floats per row
  __device__ __forceinline__ void row_matvec(float (&acc)[64], const float* xs, int xstride, int K, const float* __restrict__ W,
                                             int ld, int nq) {
    // (two / four reduction indices per pass: the weight loads of the next index are in flight under the FMAs of this one)
    if (nq <= 1) {
#pragma unroll 4
      for (int k = 0; k < K; ++k) row_axpy<1>(acc, xs[k * xstride], W + (size_t)k * ld);
    } else if (nq <= 4) {
#pragma unroll 4
      for (int k = 0; k < K; ++k) row_axpy<4>(acc, xs[k * xstride], W + (size_t)k * ld);
    } else if (nq <= 8) {
#pragma unroll 2
      for (int k = 0; k < K; ++k) row_axpy<8>(acc, xs[k * xstride], W + (size_t)k * ld);
    } else {
#pragma unroll 2
      for (int k = 0; k < K; ++k) row_axpy<16>(acc, xs[k * xstride], W + (size_t)k * ld);
    }
  }
  __device__ void rhs_ffjord_fwd_rows(int tile, float* kout, unsigned int call) {
    const MlpDesc& M = P.mlp;
    const int D = P.D, NV = P.NV, B = P.B, TB = P.TB, L = M.nl;
    const int base_row = tile * TB;
    const int Lp = M.dpad[L];
    float* e = sm + P.sp.dl_off;                       // noise (or the unit vector of the pass), [row][Lp] (fill_noise layout)
    // exact trace: sum_d e_d^T J e_d, one reverse pass per unit vector; with regularize one more pass with e ~ randn for
    // ||e^T J|| (src/ffjord.jl:124-125)
    const int npass = P.exact_trace ? D + (P.regularize ? 1 : 0) : 1;
    float acc[64];
    for (int pass = 0; pass < npass; ++pass) {
    if (pass > 0) sync();                              // every thread has read the previous pass' seed
    const bool unit_pass = P.exact_trace && pass < D;
    fill_noise(tile, 0u, call, e, Lp, unit_pass ? pass : -1);
    sync();
    for (int t = tid; t < TB; t += NTHREADS) {
      // ---- forward: h_{l+1} = act(h_l W_l + b_l); hidden h_l kept column-major in the act_off[l] buffer
      for (int l = 0; l < L; ++l) {
        const int K = M.dims[l], Np = M.dpad[l + 1], N = M.dims[l + 1], nq = Np >> 2;
        const float* bias = sm + P.sp.roww_off + M.b_off[l];
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
          if (q < nq) b = *(reinterpret_cast<const float4*>(bias) + q);
          acc[4 * q] = b.x; acc[4 * q + 1] = b.y; acc[4 * q + 2] = b.z; acc[4 * q + 3] = b.w;
        }
        if (l == 0) row_matvec(acc, sm + P.sp.act_off[0] + t * M.dpad[0], 1, K, sm + P.sp.roww_off + M.w_off[0], Np, nq);
        else row_matvec(acc, sm + P.sp.act_off[l] + t, TB, K, sm + P.sp.roww_off + M.w_off[l], Np, nq);
        const int a = M.act[l];
        if (l < L - 1) {
          float* H = sm + P.sp.act_off[l + 1] + t;
#pragma unroll
          for (int n = 0; n < 64; ++n)
            if (n < N) H[n * TB] = act_apply(a, acc[n]);
        } else if (pass == 0 && base_row + t < B) {
          float* o = kout + (long long)(base_row + t) * NV;
          float s1 = 0.f;
#pragma unroll
          for (int n = 0; n < 64; ++n)
            if (n < D) {
              const float m = act_apply(a, acc[n]);
              o[n] = m;
              s1 = fmaf(m, m, s1);
            }
          if (P.regularize) o[D + 1] = s1;
        }
      }
      // ---- reverse: eJ = e^T dm/dz; delta_L = e, delta_l = (delta_{l+1} W_l^T) .* act'(h_l), in place in the h buffers
      float tr = 0.f, s2 = 0.f;
      for (int l = L - 1; l >= 0; --l) {
        const int N = M.dims[l + 1], Kp = M.dpad[l], K = M.dims[l], kq = Kp >> 2;
#pragma unroll
        for (int n = 0; n < 64; ++n) acc[n] = 0.f;
        if (l == L - 1) row_matvec(acc, e + t * Lp, 1, N, sm + P.sp.rowwt_off + M.wt_off[l], Kp, kq);
        else row_matvec(acc, sm + P.sp.act_off[l + 1] + t, TB, N, sm + P.sp.rowwt_off + M.wt_off[l], Kp, kq);
        if (l > 0) {
          const int a = M.act[l - 1];
          float* H = sm + P.sp.act_off[l] + t;
#pragma unroll
          for (int k = 0; k < 64; ++k)
            if (k < K) H[k * TB] = acc[k] * act_d1(a, H[k * TB]);
        } else {
#pragma unroll
          for (int k = 0; k < 64; ++k)
            if (k < D) {
              tr = fmaf(acc[k], e[t * Lp + k], tr);
              s2 = fmaf(acc[k], acc[k], s2);
            }
        }
      }
      if (base_row + t < B) {
        float* o = kout + (long long)(base_row + t) * NV + D;
        if (!P.exact_trace || unit_pass) *o = pass == 0 ? -tr : *o - tr;
        if (P.regularize && !unit_pass) o[2] = sqrtf(s2);
      }
    }
    }
    sync();
  }

  // d/d(z,p) of  <lz, m(z)> - ll * <e, J(z) e>  through the doubled (h, hdot) network.
  __device__ void rhs_ffjord_adjoint(int tile, float* kout, const SlabOut& so, bool acc_store,
                                     unsigned int call) {
    const MlpDesc& M = P.mlp;
    const int D = P.D, NV = P.NV, B = P.B;
    const int base_row = tile * P.TB;
    const int L = M.nl, Lp = M.dpad[L];
    float* hb = sm + P.sp.dl_off;     // hbar_L = lz (already routed)
    float* hdb = sm + P.sp.dl2_off;   // hdbar_L = -ll * e
    const float* ll = sm + P.sp.ll_off;
    // Everything below is linear in the seed pair (hbar_L, hdbar_L) given the tangent direction hdot_0, so the RHS is a sum
    // of passes; the first emitting pass stores its results (derivative rows, gradient slabs), the others accumulate.
    //   kind 0 (trace):  Hutchinson: one pass, direction e ~ randn, seeds (lz, -ll e).  Exact trace (monte_carlo = false):
    //                    <lz, m(z)> - ll * sum_d <e_d, J(z) e_d>  =  D passes with direction e_d, seeds (lz / D, -ll e_d).
    //   regularize = true adds  l1 * sum(abs2, m(z)) + l2 * ||e^T J||  (l1, l2: cotangents of the two extra state rows):
    //                    the first term is 2 l1 m(z) more on hbar_L; for the second,
    //   kind 1:          seeds (e, 0): the z-cotangent of this pass is eJ = e^T J itself (nothing emitted), then
    //   kind 2:          direction w = l2 eJ / ||eJ|| held constant, seeds (0, e):  d/d(z,p) <e, J w>.
    const int nT = P.exact_trace ? D : 1;
    const int npass = nT + (P.regularize ? 2 : 0);
    if (nT > 1) {
      const float inv = 1.f / (float)nT;
      for (int e = tid; e < P.TB * Lp; e += NTHREADS) hb[e] *= inv;
      sync();
    }
    bool emitted = false;
    for (int pass = 0; pass < npass; ++pass) {
    const int kind = pass < nT ? 0 : (pass == nT ? 1 : 2);
    const bool emit = kind != 1;
    const bool acc_pass = emitted;
    const bool acc_store_p = acc_store || emitted;
    float* dir = sm + P.sp.hd_off[0];                    // hdot_0
    if (kind == 0) {
      fill_noise(tile, 1u, call, dir, M.dpad[0], P.exact_trace ? pass : -1);
      sync();
      for (int e = tid; e < P.TB * Lp; e += NTHREADS) {
        const int r = e / Lp, c = e - r * Lp;
        hdb[e] = (c < D) ? -ll[r] * dir[r * M.dpad[0] + c] : 0.f;
      }
    } else if (kind == 1) {
      fill_noise(tile, 1u, call, dir, M.dpad[0]);        // (any finite direction: hdbar = 0 makes the pass first order)
      for (int e = tid; e < P.TB * Lp; e += NTHREADS) { hb[e] = 0.f; hdb[e] = 0.f; }
      sync();
      fill_noise(tile, 1u, call, hb, Lp);
      sync();
    } else {
      for (int e = tid; e < P.TB * Lp; e += NTHREADS) { hb[e] = 0.f; hdb[e] = 0.f; }
      sync();
      fill_noise(tile, 1u, call, hdb, Lp);
      sync();
    }
    // forward of (h, hdot); hidden layers store h_l, q_l, hdot_l
    for (int l = 0; l < L - 1; ++l) {
      const int Kp = M.dpad[l], Np = M.dpad[l + 1], N = M.dims[l + 1], a = M.act[l];
      const float* W = P.wp + M.w_off[l];
      const float* bias = P.wp + M.b_off[l];
      float* H = sm + P.sp.act_off[l + 1];
      float* Q = sm + P.sp.q_off[l + 1];
      float* HD = sm + P.sp.hd_off[l + 1];
      tile_gemm(sm + P.sp.act_off[l], Kp, W, Np, P.TB, Kp, Np,
                [&](int row0, int col0, float (&acc)[8][4]) {
                  const float4 b = __ldg(reinterpret_cast<const float4*>(bias + col0));
#pragma unroll
                  for (int i = 0; i < 8; ++i) {
                    float4 v;
                    v.x = (col0 + 0 < N) ? act_apply(a, acc[i][0] + b.x) : 0.f;
                    v.y = (col0 + 1 < N) ? act_apply(a, acc[i][1] + b.y) : 0.f;
                    v.z = (col0 + 2 < N) ? act_apply(a, acc[i][2] + b.z) : 0.f;
                    v.w = (col0 + 3 < N) ? act_apply(a, acc[i][3] + b.w) : 0.f;
                    *reinterpret_cast<float4*>(H + (row0 + i) * Np + col0) = v;
                  }
                });
      sync();
      tile_gemm(sm + P.sp.hd_off[l], Kp, W, Np, P.TB, Kp, Np,
                [&](int row0, int col0, float (&acc)[8][4]) {
#pragma unroll
                  for (int i = 0; i < 8; ++i) {
                    const float4 h = *reinterpret_cast<const float4*>(H + (row0 + i) * Np + col0);
                    const float4 q = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
                    float4 hd;
                    hd.x = (col0 + 0 < N) ? act_d1(a, h.x) * q.x : 0.f;
                    hd.y = (col0 + 1 < N) ? act_d1(a, h.y) * q.y : 0.f;
                    hd.z = (col0 + 2 < N) ? act_d1(a, h.z) * q.z : 0.f;
                    hd.w = (col0 + 3 < N) ? act_d1(a, h.w) * q.w : 0.f;
                    *reinterpret_cast<float4*>(Q + (row0 + i) * Np + col0) = q;
                    *reinterpret_cast<float4*>(HD + (row0 + i) * Np + col0) = hd;
                  }
                });
      sync();
    }
    if (P.regularize && pass == 0) {
      // hbar_L += 2 l1 m(z) (/ D in the exact-trace split): m(z) = h_{L-1} W_{L-1} + b_{L-1}, the one layer the doubled
      // forward above does not evaluate
      const int Kp = M.dpad[L - 1], Np = M.dpad[L];
      const float* bias = P.wp + M.b_off[L - 1];
      const float* l1 = ll + P.TB;
      const float sc = 2.f / (float)nT;
      tile_gemm(sm + P.sp.act_off[L - 1], Kp, P.wp + M.w_off[L - 1], Np, P.TB, Kp, Np,
                [&](int row0, int col0, float (&acc)[8][4]) {
                  const float4 b = __ldg(reinterpret_cast<const float4*>(bias + col0));
#pragma unroll
                  for (int i = 0; i < 8; ++i) {
                    float4* hp = reinterpret_cast<float4*>(hb + (row0 + i) * Np + col0);
                    const float w = sc * l1[row0 + i];
                    float4 v = *hp;
                    if (col0 + 0 < D) v.x = fmaf(w, acc[i][0] + b.x, v.x);
                    if (col0 + 1 < D) v.y = fmaf(w, acc[i][1] + b.y, v.y);
                    if (col0 + 2 < D) v.z = fmaf(w, acc[i][2] + b.z, v.z);
                    if (col0 + 3 < D) v.w = fmaf(w, acc[i][3] + b.w, v.w);
                    *hp = v;
                  }
                });
      sync();
    }
    // reverse: (prebar_l, qbar_l) live in (X, Y); last layer is identity so they are the seeds
    const float* X = hb;
    const float* Y = hdb;
    for (int l = L; l >= 1; --l) {
      const int li = l - 1;
      const int Np = M.dpad[l], Kp = M.dpad[li];
      if (emit) {
        layer_wgrad(li, sm + P.sp.act_off[li], X, sm + P.sp.hd_off[li], Y, so, acc_store_p);
        sync();
      }
      const float* WT = P.wtp + M.wt_off[li];
      if (li >= 1) {
        const int a = M.act[li - 1], Kd = M.dims[li];
        float* H = sm + P.sp.act_off[li];
        float* Q = sm + P.sp.q_off[li];
        float* HD = sm + P.sp.hd_off[li];
        // hdbar_{l-1} = qbar_l W^T ; qbar_{l-1} -> HD, T = hdbar * q * s2 -> Q
        tile_gemm(Y, Np, WT, Kp, P.TB, Np, Kp, [&](int row0, int col0, float (&acc)[8][4]) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 h = *reinterpret_cast<const float4*>(H + (row0 + i) * Kp + col0);
            float4* qp = reinterpret_cast<float4*>(Q + (row0 + i) * Kp + col0);
            const float4 q = *qp;
            float4 qb, tt;
            qb.x = (col0 + 0 < Kd) ? acc[i][0] * act_d1(a, h.x) : 0.f;
            qb.y = (col0 + 1 < Kd) ? acc[i][1] * act_d1(a, h.y) : 0.f;
            qb.z = (col0 + 2 < Kd) ? acc[i][2] * act_d1(a, h.z) : 0.f;
            qb.w = (col0 + 3 < Kd) ? acc[i][3] * act_d1(a, h.w) : 0.f;
            tt.x = acc[i][0] * q.x * act_d2(a, h.x);
            tt.y = acc[i][1] * q.y * act_d2(a, h.y);
            tt.z = acc[i][2] * q.z * act_d2(a, h.z);
            tt.w = acc[i][3] * q.w * act_d2(a, h.w);
            *reinterpret_cast<float4*>(HD + (row0 + i) * Kp + col0) = qb;
            *qp = tt;
          }
        });
        sync();
        // hbar_{l-1} = prebar_l W^T ; prebar_{l-1} = hbar * s1 + T -> H
        tile_gemm(X, Np, WT, Kp, P.TB, Np, Kp, [&](int row0, int col0, float (&acc)[8][4]) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float4* hp = reinterpret_cast<float4*>(H + (row0 + i) * Kp + col0);
            const float4 h = *hp;
            const float4 tt = *reinterpret_cast<const float4*>(Q + (row0 + i) * Kp + col0);
            float4 v;
            v.x = (col0 + 0 < Kd) ? fmaf(acc[i][0], act_d1(a, h.x), tt.x) : 0.f;
            v.y = (col0 + 1 < Kd) ? fmaf(acc[i][1], act_d1(a, h.y), tt.y) : 0.f;
            v.z = (col0 + 2 < Kd) ? fmaf(acc[i][2], act_d1(a, h.z), tt.z) : 0.f;
            v.w = (col0 + 3 < Kd) ? fmaf(acc[i][3], act_d1(a, h.w), tt.w) : 0.f;
            *hp = v;
          }
        });
        sync();
        X = H;
        Y = HD;
      } else {
        tile_gemm(X, Np, WT, Kp, P.TB, Np, Kp, [&](int row0, int col0, float (&acc)[8][4]) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int row = row0 + i;
            if (!emit) {                                 // kind 1: eJ = e^T J into the direction buffer of the next pass
              *reinterpret_cast<float4*>(dir + row * Kp + col0) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
            } else if (base_row + row < B) {
              float* o = kout + (long long)(base_row + row) * NV + col0;
              if (col0 + 0 < D) o[0] = (acc_pass ? o[0] : 0.f) - acc[i][0];
              if (col0 + 1 < D) o[1] = (acc_pass ? o[1] : 0.f) - acc[i][1];
              if (col0 + 2 < D) o[2] = (acc_pass ? o[2] : 0.f) - acc[i][2];
              if (col0 + 3 < D) o[3] = (acc_pass ? o[3] : 0.f) - acc[i][3];
            }
          }
        });
        if (emit) {
          for (int r = tid; r < P.TB; r += NTHREADS)
            if (base_row + r < B)
              for (int c = D; c < NV; ++c) kout[(long long)(base_row + r) * NV + c] = 0.f;
        }
        sync();
        if (!emit) {
          // w = l2 * eJ / ||eJ||_2 per trajectory (0 where the norm vanishes: the reference's sqrt has no derivative there)
          const float* l2 = ll + 2 * P.TB;
          for (int r = tid; r < P.TB; r += NTHREADS) {
            float* w = dir + r * Kp;
            float s2 = 0.f;
            for (int d = 0; d < D; ++d) s2 = fmaf(w[d], w[d], s2);
            const float nrm = sqrtf(s2);
            for (int d = 0; d < Kp; ++d) w[d] = (d < D && nrm > 0.f) ? (l2[r] * w[d]) / nrm : 0.f;
          }
          sync();
        }
      }
    }
    if (emit) emitted = true;
    }   // pass
  }

  // ---------------------------------------------------------------- whole-vector helpers
  template <typename F>
  __device__ void for_my_tiles(F&& f) {
    for (int tile = cta; tile < P.ntiles; tile += P.nctas) f(tile);
  }

  // One RHS evaluation for one tile: stage input x = u + dt * sum_j coef[j] k_j (also stored to dst if
  // given), derivative to kout.  FFMA path: staged through shared memory; tensor-core path: fused.
  __device__ void stage_eval(int tile, const float* u, const float* const* kp, const float* coef, int nk,
                             float dt, float* dst, float* kout, const InterpInfo* ii, const SlabOut& so,
                             bool acc_store, unsigned int call) {
    if constexpr (TC) {
      if constexpr (TC_WORKERS > NTHREADS) {      // hand the stage to the helper warps
        if (tid == 0) {
          TcArgs& a = ss.targs;
          a.u = u; a.dst = dst; a.kout = kout; a.so = so; a.dt = dt;
          a.tile = tile; a.nk = nk; a.acc_store = acc_store ? 1 : 0; a.cmd = 1;
          for (int j = 0; j < 6; ++j) { a.kp[j] = j < nk ? kp[j] : u; a.coef[j] = j < nk ? coef[j] : 0.f; }
          if (ii) a.ii = *ii;
        }
        tc_sync_workers();
      }
      if constexpr (Traits<FL>::adj) tc_adj_stage_eval(tile, u, kp, coef, nk, dt, dst, kout, ii, so, acc_store);
      else tc_stage_eval(tile, u, kp, coef, nk, dt, dst, kout);
      tc_sync_workers();                          // results of every worker visible to the solver threads
      if constexpr (Traits<FL>::adj) tc_bias_epilogue(so, acc_store);
    } else {
      build_stage(tile, u, kp, coef, nk, dt, dst);
      if constexpr (Traits<FL>::interp) build_y_interp(tile, *ii);
      rhs_tile(tile, kout, so, acc_store, call);
    }
  }

  // RHS at an arbitrary state vector `u` (time info in ii for Interpolating flavours)
  __device__ void eval_rhs_vec(const float* u, float* kout, const InterpInfo* ii, const SlabOut& so) {
    bool first = true;
    for_my_tiles([&](int tile) {
      stage_eval(tile, u, nullptr, nullptr, 0, 0.f, nullptr, kout, ii, so, !first, ss.rhs_calls);
      first = false;
    });
    sync();
    if (tid == 0) { ss.rhs_calls += 1u; ss.nf += 1; }
    sync();
  }

  // ---------------------------------------------------------------- tensor-core engine (tcgen05)
#include "tc_engine.inc"


  __device__ void zero_slab(float* s) {
    for (int i = tid; i < P.mlp.ppad; i += NTHREADS) s[i] = 0.f;
  }

  // Totals over ALL trajectories (all CTAs, all ranks) of Q0 = c1*slab[wx] + c2*slab[wy] and, if
  // wz >= 0, Q1 = slab[wz], for the packed-parameter slice this CTA owns; returns the fp64 sum of
  // f(j, tot0, tot1).  Cross-CTA and cross-rank sums run in fp64 in a fixed order.  With several
  // ranks the slice totals travel through the peers' mailboxes (the parameter-gradient
  // terms of the error norm need the global sums every step).
  template <typename F>
  __device__ double mu_reduce(int wx, float c1, int wy, float c2, int wz, F&& f) {
    const int pp = P.mlp.ppad;
    int chunk = (pp + P.nctas - 1) / P.nctas;
    chunk = (chunk + 3) & ~3;
    const int lo = min(pp, cta * chunk), hi = min(pp, lo + chunk);
    double part = 0.0;
    auto local = [&](int j, double& t0, double& t1) {
      if (wx == -2) { step_local(j, wy, wz, t0, t1); return; }   // step mode: wy = g1, wz = g7
      t0 = 0.0;
      t1 = 0.0;
      for (int cidx = 0; cidx < P.nctas; ++cidx) {
        const float* base = P.slabs + (long long)cidx * NSLAB * pp + j;
        double v = (double)c1 * (double)__ldcg(base + (long long)wx * pp);
        if (wy >= 0) v += (double)c2 * (double)__ldcg(base + (long long)wy * pp);
        t0 += v;
        if (wz >= 0) t1 += (double)__ldcg(base + (long long)wz * pp);
      }
    };
    // every slice element this thread owns, with its two cross-CTA totals
    auto produce = [&](auto&& sink) {
      if constexpr (TC && Traits<FL>::adj) {
        if (wx == -2) {                               // step mode: slabs staged through shared memory by bulk copies
          for (int slo = lo; slo < hi; slo += MU_SUB) {
            const int n = min(MU_SUB, hi - slo);
            double T0[MU_SUB / NTHREADS], T1[MU_SUB / NTHREADS];
            step_staged(slo, n, wy, wz, T0, T1);
#pragma unroll
            for (int k = 0; k < MU_SUB / NTHREADS; ++k) {
              const int jj = tid + k * NTHREADS;
              if (jj < n) sink(slo + jj, T0[k], T1[k]);
            }
          }
          return;
        }
      }
      for (int j = lo + tid; j < hi; j += NTHREADS) {
        double t0, t1;
        local(j, t0, t1);
        sink(j, t0, t1);
      }
    };
    if (P.world == 1) {
      produce([&](int j, double t0, double t1) { part += (double)f(j, (float)t0, (float)t1); });
      return part;
    }
    if (lo >= hi) return 0.0;                       // same decision on every rank (same grid)
    const unsigned long long xe = ss.mu_xe + 1ull;
    const unsigned int par = (unsigned int)(xe & 1ull);
    produce([&](int j, double t0, double t1) {
      for (int r = 0; r < P.world; ++r) {
        Mailbox* mb = P.peers[r];
        mb->xmu[par][P.rank][0][j] = t0;
        if (wz >= 0) mb->xmu[par][P.rank][1][j] = t1;
      }
    });
    __threadfence_system();
    sync();
    if (tid == 0) {
      for (int r = 0; r < P.world; ++r) P.peers[r]->xtag[par][P.rank][cta] = xe;
      Mailbox* me = P.peers[P.rank];
      long long t_start = clock64();
      for (int r = 0; r < P.world; ++r) {
        while (me->xtag[par][r][cta] < xe) {
          if (ld_volatile_i32(&P.ctrl->abort_flag)) break;
          if (clock64() - t_start > 40000000000LL) {
            atomicCAS(&P.ctrl->retcode, 0, -5);
            atomicExch(&P.ctrl->abort_flag, 1);
            break;
          }
        }
      }
      __threadfence_system();
    }
    sync();
    const Mailbox* me = P.peers[P.rank];
    for (int j = lo + tid; j < hi; j += NTHREADS) {
      double t0 = 0.0, t1 = 0.0;
      for (int r = 0; r < P.world; ++r) {
        t0 += __ldcg(&me->xmu[par][r][0][j]);
        if (wz >= 0) t1 += __ldcg(&me->xmu[par][r][1][j]);
      }
      part += (double)f(j, (float)t0, (float)t1);
    }
    // every rank evaluates the same global terms (it needs mu_new locally); only rank 0 adds them
    // to the norm, otherwise the cross-rank sum would count them `world` times
    return P.rank == 0 ? part : 0.0;
  }
  // The step's two parameter-gradient combinations for slice element j, summed over all CTAs in fp64:
  //   t0 = sum_s b_s G_s  (mu update),  t1 = sum_s btilde_s G_s  (embedded error), G_1 = slab g1, G_7 = slab g7.
  __device__ __forceinline__ void step_local(int j, int g1, int g7, double& t0, double& t1) const {
    const int pp = P.mlp.ppad;
    constexpr int U = 8;                       // independent chains: 56 loads in flight per thread
    double sa[U], se[U];
#pragma unroll
    for (int u = 0; u < U; ++u) sa[u] = se[u] = 0.0;
    auto one = [&](int cidx, float& a, float& e) {
      const float* base = P.slabs + (long long)cidx * NSLAB * pp + j;
      const float k1 = __ldcg(base + (long long)g1 * pp), k7 = __ldcg(base + (long long)g7 * pp);
      float ks[5];
#pragma unroll
      for (int s5 = 0; s5 < 5; ++s5) ks[s5] = __ldcg(base + (long long)s5 * pp);
      a = c_tab.a[6][0] * k1;
      e = c_tab.bt[0] * k1;
#pragma unroll
      for (int s5 = 0; s5 < 5; ++s5) {
        a = fmaf(c_tab.a[6][1 + s5], ks[s5], a);
        e = fmaf(c_tab.bt[1 + s5], ks[s5], e);
      }
      e = fmaf(c_tab.bt[6], k7, e);
    };
    int cidx = 0;
    for (; cidx + U <= P.nctas; cidx += U) {
      float a[U], e[U];
#pragma unroll
      for (int u = 0; u < U; ++u) one(cidx + u, a[u], e[u]);
#pragma unroll
      for (int u = 0; u < U; ++u) { sa[u] += (double)a[u]; se[u] += (double)e[u]; }
    }
    for (int u = 0; cidx < P.nctas; ++cidx, ++u) {
      float a, e;
      one(cidx, a, e);
      sa[u] += (double)a;
      se[u] += (double)e;
    }
    t0 = ((sa[0] + sa[1]) + (sa[2] + sa[3])) + ((sa[4] + sa[5]) + (sa[6] + sa[7]));
    t1 = ((se[0] + se[1]) + (se[2] + se[3])) + ((se[4] + se[5]) + (se[6] + se[7]));
  }
  // Tensor-core adjoint kernels: the same totals for slice elements [slo, slo + n), n <= MU_SUB, with the slab
  // slices of MU_CB CTAs at a time staged in the (idle) operand buffers by cp.async.bulk -- the copy engine
  // keeps ~170 KB in flight per SM where register-held loads reach ~30 GB/s per SM.  Thread t gets elements
  // slo + t + k*NTHREADS.  Sums over CTAs run in fp64 in CTA order (deterministic).
  static constexpr int MU_SUB = 1024;
  static constexpr int MU_STAGE_BYTES = 2 * TC_A_SLOT + TC_NB * TC_B_SLOT;      // the operand buffers, idle in the mu phase
  unsigned int mu_use[2] = {0u, 0u};
  __device__ void step_staged(int slo, int n, int g1, int g7, double (&T0)[MU_SUB / NTHREADS], double (&T1)[MU_SUB / NTHREADS]) {
    using namespace b200tc;
    const int pp = P.mlp.ppad;
    float* stage = reinterpret_cast<float*>(tc_base());
    // two staging buffers; each holds the 7 slab slices (n floats, row pitch nst) of `cb` CTAs
#ifdef B200_MU_ADAPTIVE     // (measured slower: 31.0 vs 30.65 ms -- the 3088-byte row pitch costs more than the 25 % fewer batches save)
    const int nst = (n + 3) & ~3;
    const int cb = max(1, min(NTHREADS / NSTEP_SLAB, (MU_STAGE_BYTES / 2) / (NSTEP_SLAB * nst * 4)));
#else
    const int nst = MU_SUB, cb = 3;
#endif
    const int BUF = cb * NSTEP_SLAB * nst;            // floats per staging buffer
    const int nb = (P.nctas + cb - 1) / cb;
    auto issue = [&](int b) {                         // all threads: one copy each (a single issuing thread spent
      const int buf = b & 1, c0 = b * cb, cn = min(cb, P.nctas - c0);   // ~1/3 of the phase issuing the copies of a batch)
      uint64_t* bar = reinterpret_cast<uint64_t*>(&ss.mu_full[buf]);
      if (tid == 0) mbar_arrive_expect_tx(bar, (uint32_t)(cn * NSTEP_SLAB * n * 4));
      if (tid < cn * NSTEP_SLAB) {
        const int c = tid / NSTEP_SLAB, sl = tid - c * NSTEP_SLAB;
        bulk_g2s(stage + buf * BUF + (c * NSTEP_SLAB + sl) * nst, P.slabs + ((long long)(c0 + c) * NSLAB + sl) * pp + slo,
                 (uint32_t)(n * 4), bar);
      }
    };
    issue(0);
    if (nb > 1) issue(1);
#pragma unroll
    for (int k = 0; k < MU_SUB / NTHREADS; ++k) T0[k] = T1[k] = 0.0;
    for (int b = 0; b < nb; ++b) {
      const int buf = b & 1, cn = min(cb, P.nctas - b * cb);
      tc_wait_bar(&ss.mu_full[buf], mu_use[buf] & 1u);
      mu_use[buf] += 1u;
      const float* sb = stage + buf * BUF;
      for (int c = 0; c < cn; ++c) {
#pragma unroll
        for (int k = 0; k < MU_SUB / NTHREADS; ++k) {
          const int jj = tid + k * NTHREADS;
          if (jj < n) {
            const float* q = sb + c * NSTEP_SLAB * nst + jj;
            const float k1 = q[g1 * nst], k7 = q[g7 * nst];
            float a = c_tab.a[6][0] * k1, e = c_tab.bt[0] * k1;
#pragma unroll
            for (int s5 = 0; s5 < 5; ++s5) {
              const float ks = q[s5 * nst];
              a = fmaf(c_tab.a[6][1 + s5], ks, a);
              e = fmaf(c_tab.bt[1 + s5], ks, e);
            }
            e = fmaf(c_tab.bt[6], k7, e);
            T0[k] += (double)a;
            T1[k] += (double)e;
          }
        }
      }
      sync();                                         // buffer drained by every thread
      if (b + 2 < nb) issue(b + 2);
    }
  }
  // Quadrature / Gauss adjoints: mu is not a state of the adjoint ODE, so no batch-wide total is needed per step.  After
  // an ACCEPTED step every CTA adds dt * sum_s b_s G_s of its own slabs to its running slab 7 (local traffic only: no grid
  // barrier, no cross-CTA or cross-rank exchange); the CTAs' sums meet once, after the last step (solver.cuh epilogue).
  // (kept small on purpose -- one float4 group per pass, the passes two ahead pulled towards L2 by prefetches: a wider,
  // register-hungry version drove the register allocation of the whole FP32-FFMA adjoint kernels from 192 to 255)
  __device__ void quad_accumulate(float dt, int g1) {
    const int pp = P.mlp.ppad;
    float* base = P.slabs + (long long)cta * NSLAB * pp;
    const int stride = NTHREADS * 4;
    auto prefetch = [&](int j) {
      if (j < pp) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (long long)g1 * pp + j));
#pragma unroll
        for (int s5 = 0; s5 < 5; ++s5) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (long long)s5 * pp + j));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(base + 7ll * pp + j));
      }
    };
    prefetch(tid * 4);
    prefetch(tid * 4 + stride);
    for (int j = tid * 4; j < pp; j += stride) {
      prefetch(j + 2 * stride);
      const float4 k1 = __ldcg(reinterpret_cast<const float4*>(base + (long long)g1 * pp + j));
      float4 a = make_float4(c_tab.a[6][0] * k1.x, c_tab.a[6][0] * k1.y, c_tab.a[6][0] * k1.z, c_tab.a[6][0] * k1.w);
#pragma unroll
      for (int s5 = 0; s5 < 5; ++s5) {
        const float4 ks = __ldcg(reinterpret_cast<const float4*>(base + (long long)s5 * pp + j));
        a.x = fmaf(c_tab.a[6][1 + s5], ks.x, a.x); a.y = fmaf(c_tab.a[6][1 + s5], ks.y, a.y);
        a.z = fmaf(c_tab.a[6][1 + s5], ks.z, a.z); a.w = fmaf(c_tab.a[6][1 + s5], ks.w, a.w);
      }
      float4* q = reinterpret_cast<float4*>(base + 7ll * pp + j);
      float4 v = __ldcg(q);
      v.x = fmaf(dt, a.x, v.x); v.y = fmaf(dt, a.y, v.y); v.z = fmaf(dt, a.z, v.z); v.w = fmaf(dt, a.w, v.w);
      __stcg(q, v);
    }
  }
  // every CTA advances the mu-exchange epoch together (call after mu_reduce, CTA-uniformly)
  __device__ void mu_reduce_done() {
    sync();
    if (tid == 0 && P.world > 1) ss.mu_xe += 1ull;
    sync();
  }
};

}  // namespace b200
