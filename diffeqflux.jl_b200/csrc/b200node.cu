// b200node.cu -- C ABI of libb200node.so (see include/b200node.h).
// Host side: packs the Lux parameter vector, plans shared memory / tiles, owns device memory and
// launches ONE cooperative persistent kernel per solve.  No CPU fallback exists on purpose.
#include "../../include/b200node.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <thread>
#include <cstring>
#include <string>
#include <vector>

#include "solver.cuh"

using namespace b200;

// tc_kernels.cu
cudaError_t b200_tc_init_tableau();
cudaError_t b200_tc_launch(int flavor, const SolveParams& P, int smem_bytes, cudaStream_t stream);

// ------------------------------------------------------------------------------------------------
struct b200_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  int num_sms = 0;
  int smem_optin = 0;
  std::string err;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  Ctrl* ctrl_dev = nullptr;       // control block (device)
  Ctrl* ctrl_host = nullptr;      // pinned copy of the header part
  // multi-GPU
  int rank = 0, world = 1;
  Mailbox* mailbox = nullptr;     // cudaMalloc'd (IPC-exportable)
  Mailbox* peers[MAX_RANKS] = {};
  unsigned long long solve_seq = 0;
  // workspace cache: device blocks are reused across calls (all work is stream-ordered on ctx->stream)
  struct Block { void* ptr; size_t size; bool free; };
  std::vector<Block> blocks;
  // multiple shooting: groups are independent solves of a fraction of the GPU each; they run concurrently on
  // sub-contexts (own stream, control block, workspace cache), one host thread per sub-context
  std::vector<b200_ctx*> subs;
  bool owns_stream = false;
  std::vector<b200_tape*> tapes;  // live tapes that own blocks of this context (neutralised by b200_ctx_destroy)
  // diagnostics
  float* trace_dev = nullptr;
  int trace_cap = 0;
  int trace_n = 0;
};

struct b200_model {
  b200_ctx* ctx;
  MlpDesc m;
  int wt_total;
};

struct b200_tape {
  b200_ctx* ctx = nullptr;
  MlpDesc m{};
  int fwd_flavor = FL_FWD_NODE;
  int sensealg = B200_SENSE_NONE;
  int B = 0, NS = 0, D = 0, d_in = 0, augment = 0;
  float t0 = 0, t1 = 0;
  float* wp = nullptr;
  float* wtp = nullptr;
  uint8_t* wimg = nullptr;     // tensor-core weight images (forward + dgrad orientation units)
  TcParams tc{};
  float* ring = nullptr;
  int nslots = 0;
  long long vec_elems = 0;
  float* k1_first = nullptr;
  float* step_t = nullptr;
  int nsteps = 0;
  int dense = 0;
  float* out_u = nullptr;    // [n_out][B][NS]
  float* out_t = nullptr;
  float* final_state = nullptr;
  int n_out = 0;
  std::vector<float> ts;     // host copy of the saved times
  unsigned long long seed = 0;
  int noise_per_call = 1;
  int b_offset = 0;
  int exact_trace = 0;
  int regularize = 0;
};

// Every entry point binds the context's device for the duration of the call and restores the caller's device
// afterwards (a host framework may be working on another GPU).
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

static int set_err(b200_ctx* ctx, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (ctx) ctx->err = buf;
  return code;
}

#define CU(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess)                                                                    \
      return set_err(ctx, B200_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                     __FILE__, __LINE__);                                                     \
  } while (0)

// Caching device allocator: best-fit reuse of blocks this context already owns, cudaMalloc otherwise.
// Reuse without synchronisation is safe because every kernel / copy of the library runs on ctx->stream.
static cudaError_t dalloc_bytes(b200_ctx* ctx, void** p, size_t bytes) {
  bytes = (std::max<size_t>(bytes, 1) + 511) & ~(size_t)511;
  int best = -1;
  for (int i = 0; i < (int)ctx->blocks.size(); ++i) {
    const auto& b = ctx->blocks[i];
    if (b.free && b.size >= bytes && b.size <= bytes + bytes / 4 + (1u << 20) && (best < 0 || b.size < ctx->blocks[best].size)) best = i;
  }
  if (best >= 0) {
    ctx->blocks[best].free = false;
    *p = ctx->blocks[best].ptr;
    return cudaSuccess;
  }
  cudaError_t e = cudaMalloc(p, bytes);
  if (e == cudaErrorMemoryAllocation) {       // release every cached free block and retry once
    cudaGetLastError();
    cudaStreamSynchronize(ctx->stream);
    for (auto it = ctx->blocks.begin(); it != ctx->blocks.end();) {
      if (it->free) { cudaFree(it->ptr); it = ctx->blocks.erase(it); } else ++it;
    }
    e = cudaMalloc(p, bytes);
  }
  if (e == cudaSuccess) ctx->blocks.push_back({*p, bytes, false});
  return e;
}
template <typename Tp>
static cudaError_t dalloc(b200_ctx* ctx, Tp** p, size_t n) {
  return dalloc_bytes(ctx, reinterpret_cast<void**>(p), n * sizeof(Tp));
}
static void dfree(b200_ctx* ctx, void* p) {
  if (!p) return;
  for (auto& b : ctx->blocks)
    if (b.ptr == p) { b.free = true; return; }
}

// ------------------------------------------------------------------------------------------------
extern "C" int b200_abi_version(void) { return B200_ABI_VERSION; }

extern "C" void b200_opts_default(b200_opts* o) {
  memset(o, 0, sizeof *o);
  o->abstol = 1e-6f;
  o->reltol = 1e-3f;
  o->dt0 = 0.f;
  o->maxiters = 100000;
  o->save_start = 1;
  o->save_end = 1;
  o->sensealg = B200_SENSE_INTERPOLATING;
  o->ptr_kind = B200_PTR_HOST;
  o->noise_per_call = 1;
}

extern "C" int b200_ctx_create(int device, void* stream, b200_ctx** out) {
  if (!out) return B200_ERR_ARG;
  *out = nullptr;
  b200_ctx* ctx = new b200_ctx();
  ctx->device = device;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) {
    // no CUDA device: fail loudly -- there is no CPU path in this library
    delete ctx;
    return B200_ERR_CUDA;
  }
  DeviceGuard guard(device);
  int cur = -1;
  if (cudaGetDevice(&cur) != cudaSuccess || cur != device) { delete ctx; return B200_ERR_CUDA; }
  cudaDeviceProp prop{};
  cudaGetDeviceProperties(&prop, device);
  ctx->num_sms = prop.multiProcessorCount;
  ctx->smem_optin = (int)prop.sharedMemPerBlockOptin;
  ctx->stream = stream ? (cudaStream_t)stream : cudaStreamLegacy;
  cudaEventCreate(&ctx->ev0);
  cudaEventCreate(&ctx->ev1);
  cudaMalloc(&ctx->ctrl_dev, sizeof(Ctrl));
  cudaMallocHost(&ctx->ctrl_host, sizeof(Ctrl));
  const Tableau tab = host_tableau();
  if (cudaMemcpyToSymbol(c_tab, &tab, sizeof tab) != cudaSuccess || b200_tc_init_tableau() != cudaSuccess) { delete ctx; return B200_ERR_CUDA; }
  *out = ctx;
  return 0;
}

extern "C" void b200_ctx_destroy(b200_ctx* ctx) {
  if (!ctx) return;
  DeviceGuard guard(ctx->device);
  for (b200_ctx* sub : ctx->subs) b200_ctx_destroy(sub);
  ctx->subs.clear();
  cudaStreamSynchronize(ctx->stream);
  // tapes that outlive their context keep no device memory: b200_tape_free on them only releases the host object
  for (b200_tape* t : ctx->tapes) t->ctx = nullptr;
  ctx->tapes.clear();
  for (int r = 0; r < ctx->world; ++r)
    if (ctx->peers[r] && r != ctx->rank) cudaIpcCloseMemHandle(ctx->peers[r]);
  for (auto& b : ctx->blocks) cudaFree(b.ptr);
  if (ctx->mailbox) cudaFree(ctx->mailbox);
  if (ctx->trace_dev) cudaFree(ctx->trace_dev);
  if (ctx->ctrl_dev) cudaFree(ctx->ctrl_dev);
  if (ctx->ctrl_host) cudaFreeHost(ctx->ctrl_host);
  cudaEventDestroy(ctx->ev0);
  cudaEventDestroy(ctx->ev1);
  if (ctx->owns_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

extern "C" int b200_ctx_trace(b200_ctx* ctx, int capacity, float* out, int max_entries) {
  if (!ctx) return B200_ERR_ARG;
  DeviceGuard guard(ctx->device);
  if (capacity >= 0 && capacity != ctx->trace_cap) {
    CU(cudaStreamSynchronize(ctx->stream));
    if (ctx->trace_dev) CU(cudaFree(ctx->trace_dev));
    ctx->trace_dev = nullptr;
    ctx->trace_cap = 0;
    ctx->trace_n = 0;
    if (capacity > 0) {
      CU(cudaMalloc(&ctx->trace_dev, (size_t)capacity * 4 * sizeof(float)));
      ctx->trace_cap = capacity;
    }
  }
  const int n = std::min(ctx->trace_n, max_entries);
  if (out && n > 0) {
    CU(cudaMemcpyAsync(out, ctx->trace_dev, (size_t)n * 4 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
  }
  return n;
}

extern "C" const char* b200_last_error(b200_ctx* ctx) { return ctx ? ctx->err.c_str() : "null ctx"; }

extern "C" int b200_ctx_sync(b200_ctx* ctx) {
  if (!ctx) return B200_ERR_ARG;
  DeviceGuard guard(ctx->device);
  CU(cudaStreamSynchronize(ctx->stream));
  return 0;
}

static inline int pad4(int x) { return (x + 3) & ~3; }

extern "C" int b200_model_create(b200_ctx* ctx, int nlayers, const int* dims, const int* acts,
                                 int pre, b200_model** out) {
  if (!ctx || !out || !dims || !acts) return B200_ERR_ARG;
  if (nlayers < 1 || nlayers > MAXL)
    return set_err(ctx, B200_ERR_UNSUPPORTED, "model: %d Dense layers (supported 1..%d)", nlayers, MAXL);
  if (pre != B200_PRE_NONE && pre != B200_PRE_CUBE)
    return set_err(ctx, B200_ERR_UNSUPPORTED, "model: unknown pre-layer %d", pre);
  b200_model* mo = new b200_model();
  mo->ctx = ctx;
  MlpDesc& m = mo->m;
  memset(&m, 0, sizeof m);
  m.nl = nlayers;
  m.pre = pre;
  int off = 0, woff = 0, foff = 0;
  for (int l = 0; l <= nlayers; ++l) {
    if (dims[l] < 1) { delete mo; return set_err(ctx, B200_ERR_ARG, "model: dims[%d] = %d", l, dims[l]); }
    m.dims[l] = dims[l];
    m.dpad[l] = pad4(dims[l]);
  }
  for (int l = 0; l < nlayers; ++l) {
    if (acts[l] < B200_ACT_IDENTITY || acts[l] > B200_ACT_SOFTPLUS) {
      delete mo;
      return set_err(ctx, B200_ERR_UNSUPPORTED, "model: unknown activation %d in layer %d", acts[l], l);
    }
    m.act[l] = acts[l];
    m.w_off[l] = off;
    off += m.dpad[l] * m.dpad[l + 1];
    m.b_off[l] = off;
    off += m.dpad[l + 1];
    m.wt_off[l] = woff;
    woff += m.dpad[l + 1] * m.dpad[l];
    m.fw_off[l] = foff;
    foff += dims[l] * dims[l + 1];
    m.fb_off[l] = foff;
    foff += dims[l + 1];
  }
  m.ppad = off;
  m.nparams = foff;
  mo->wt_total = woff;
  *out = mo;
  return 0;
}
extern "C" void b200_model_destroy(b200_model* m) { delete m; }
extern "C" int b200_model_nparams(const b200_model* m) { return m ? m->m.nparams : B200_ERR_ARG; }

// ------------------------------------------------------------------------------------------------ small kernels
__global__ void pack_params_kernel(MlpDesc m, const float* __restrict__ p, float* __restrict__ wp,
                                   float* __restrict__ wtp) {
  const int l = blockIdx.y;
  const int K = m.dims[l], N = m.dims[l + 1], Kp = m.dpad[l], Np = m.dpad[l + 1];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Kp * Np; i += gridDim.x * blockDim.x) {
    const int k = i / Np, n = i - k * Np;
    const float v = (k < K && n < N) ? p[m.fw_off[l] + k * N + n] : 0.f;
    wp[m.w_off[l] + i] = v;
    wtp[m.wt_off[l] + n * Kp + k] = v;
  }
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < Np; n += gridDim.x * blockDim.x)
    wp[m.b_off[l] + n] = (n < N) ? p[m.fb_off[l] + n] : 0.f;
}

__global__ void unpack_grad_kernel(MlpDesc m, const float* __restrict__ g, float* __restrict__ dp,
                                   int accumulate) {
  const int l = blockIdx.y;
  const int K = m.dims[l], N = m.dims[l + 1], Np = m.dpad[l + 1];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < K * N; i += gridDim.x * blockDim.x) {
    const int k = i / N, n = i - k * N;
    const float v = g[m.w_off[l] + k * Np + n];
    float* o = dp + m.fw_off[l] + i;
    *o = accumulate ? *o + v : v;
  }
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
    float* o = dp + m.fb_off[l] + n;
    *o = accumulate ? *o + g[m.b_off[l] + n] : g[m.b_off[l] + n];
  }
}

// tensor-core weight image: layer l, K-chunk c -> [hi tile N x 128 B][lo tile], K-major SWIZZLE_128B
__global__ void pack_tc_kernel(MlpDesc m, TcParams tc, const float* __restrict__ p, uint8_t* __restrict__ wimg) {
  const int l = blockIdx.y;
  const int K = m.dims[l], N = m.dims[l + 1];
  const unsigned int unit_bytes = 2u * N * 64u;                   // [hi N x 64 B][lo N x 64 B]
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N * (K / 4); i += gridDim.x * blockDim.x) {
    const int n = i % N, k4 = i / N;
    const int unit = k4 >> 2, kk = (k4 & 3) * 4;                  // 16 inputs per unit
    float4 h, lo;
    const float* w = p + m.fw_off[l] + (size_t)(16 * unit + kk) * N + n;
    b200tc::split_tf32(w[0], h.x, lo.x);
    b200tc::split_tf32(w[N], h.y, lo.y);
    b200tc::split_tf32(w[2 * N], h.z, lo.z);
    b200tc::split_tf32(w[3 * N], h.w, lo.w);
    uint8_t* dst = wimg + tc.w_off[l] + (size_t)unit * unit_bytes + b200tc::sw64_off(n, kk);
    *reinterpret_cast<float4*>(dst) = h;
    *reinterpret_cast<float4*>(dst + N * 64) = lo;
  }
}

// dgrad orientation: layer l, unit u = 16 outputs n: [hi tile: K rows x 64 B][lo tile], rows = inputs k
__global__ void pack_tc_t_kernel(MlpDesc m, TcParams tc, const float* __restrict__ p, uint8_t* __restrict__ wimg_t) {
  const int l = blockIdx.y;
  const int K = m.dims[l], N = m.dims[l + 1];
  const unsigned int unit_bytes = 2u * K * 64u;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < K * (N / 4); i += gridDim.x * blockDim.x) {
    const int n4 = i % (N / 4), k = i / (N / 4);
    const int unit = n4 >> 2, nn = (n4 & 3) * 4;
    const float4 w = *reinterpret_cast<const float4*>(p + m.fw_off[l] + (size_t)k * N + 16 * unit + nn);
    float4 h, lo;
    b200tc::split_tf32(w.x, h.x, lo.x);
    b200tc::split_tf32(w.y, h.y, lo.y);
    b200tc::split_tf32(w.z, h.z, lo.z);
    b200tc::split_tf32(w.w, h.w, lo.w);
    uint8_t* dst = wimg_t + tc.t_off[l] + (size_t)unit * unit_bytes + b200tc::sw64_off(k, nn);
    *reinterpret_cast<float4*>(dst) = h;
    *reinterpret_cast<float4*>(dst + K * 64) = lo;
  }
}

// shapes the tcgen05 engine handles: K dims multiples of 32, N dims multiples of 16 and <= 256
static bool tc_eligible(const MlpDesc& m, int flavor) {
  if (flavor != FL_FWD_NODE && flavor != FL_ADJ_NODE_BACKSOLVE && flavor != FL_ADJ_NODE_INTERP) return false;
  if (m.pre != B200_PRE_NONE || m.nl < 1 || m.nl > MAXL_TC) return false;
  const bool adj = flavor != FL_FWD_NODE;
  if (adj && m.act[m.nl - 1] != ACT_ID) return false;   // the cotangent sweep seeds delta_L = lambda (identity output layer)
  for (int l = 0; l < m.nl; ++l) {
    if (m.dims[l] % 32 || m.dims[l] > 256) return false;
    if (m.dims[l + 1] % (adj ? 32 : 16) || m.dims[l + 1] > 256) return false;
  }
  return true;
}

// logpx = -sum(z^2)/2 - D/2 log(2 pi) - delta_logp      (src/ffjord.jl:171-176)
__global__ void ffjord_logpx_kernel(const float* __restrict__ zend, int B, int D, int NS, float* __restrict__ logpx) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* z = zend + (long long)b * NS;
  float s = 0.f;
  for (int d = 0; d < D; ++d) s = fmaf(z[d], z[d], s);
  logpx[b] = -s / 2.f - (float)(D / 2.0 * 1.8378770664093453) - z[D];
}
__global__ void add_into_kernel(float* __restrict__ dst, const float* __restrict__ src, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] += src[i];
}
// cotangent of the end state from the cotangent of logpx
// (g: cotangent of logpx or null; gz: cotangent of the end state itself or null)
__global__ void ffjord_seed_kernel(const float* __restrict__ zend, const float* __restrict__ g,
                                   const float* __restrict__ gz, int B, int D, int NS, float* __restrict__ dL) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* z = zend + (long long)b * NS;
  float* o = dL + (long long)b * NS;
  const float gb = g ? g[b] : 0.f;
  for (int c = 0; c < NS; ++c) {
    float v = c < D ? -z[c] * gb : (c == D ? -gb : 0.f);
    if (gz) v += gz[(long long)b * NS + c];
    o[c] = v;
  }
}
// du0 (first d_in of NV columns) from the final adjoint state
__global__ void extract_cols_kernel(const float* __restrict__ src, int B, int NV, int ncols,
                                    float* __restrict__ dst, int accumulate) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * ncols) return;
  const int b = (int)(i / ncols), c = (int)(i - (long long)b * ncols);
  const float v = src[(long long)b * NV + c];
  dst[i] = accumulate ? dst[i] + v : v;
}

// sum-of-squares loss against the data and its cotangent: loss = sum (pred - data)^2, dL = 2 (pred - data)
// (docs/src/examples/neural_ode.md:36-44 `loss = sum(abs2, ode_data .- pred)`); fp64 partial per block, fixed order
__global__ void sse_loss_kernel(const float* __restrict__ pred, const float* __restrict__ data, long long n,
                                float* __restrict__ dL, double* __restrict__ partial) {
  __shared__ double red[256];
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float d = pred[i] - data[i];
    dL[i] = 2.f * d;
    acc += (double)d * (double)d;
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}
__global__ void sum_partials_kernel(const double* __restrict__ partial, int n, float* __restrict__ out) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += partial[i];
    *out = (float)s;
  }
}
// Optimisers.jl Adam (OptimizationOptimisers.Adam, docs/src/examples/neural_ode.md:70):
//   m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2;  p -= lr * (m / (1 - b1^t)) / (sqrt(v / (1 - b2^t)) + eps)
__global__ void adam_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
                            const float* __restrict__ g, int n, float lr, float b1, float b2, float eps, float c1, float c2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gi = g[i];
  const float mi = b1 * m[i] + (1.f - b1) * gi;
  const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  p[i] -= lr * (mi / c1) / (sqrtf(vi / c2) + eps);
}
// general strided gather used by b200_dimmover: out (contiguous, <= 5 axes, axis 0 fastest) <- in[sum idx_a * stride_a]
struct PermDesc { int nd; long long dim[5]; long long stride[5]; };
__global__ void strided_copy_kernel(PermDesc d, const float* __restrict__ in, float* __restrict__ out, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    long long r = i, off = 0;
#pragma unroll
    for (int a = 0; a < 5; ++a) {
      if (a < d.nd) {
        const long long q = r / d.dim[a];
        off += (r - q * d.dim[a]) * d.stride[a];
        r = q;
      }
    }
    out[i] = in[off];
  }
}

// ------------------------------------------------------------------------------------------------ planning
static int plan_smem(const MlpDesc& m, int flavor, int TB, SmemPlan& sp, bool rowthread = false) {
  memset(&sp, 0, sizeof sp);
  int off = 0;
  auto take = [&](int floats_per_row) {
    const int o = off;
    off += floats_per_row * TB;
    return o;
  };
  for (int l = 0; l < m.nl; ++l) sp.act_off[l] = take(m.dpad[l]);
  if (flavor != FL_FWD_NODE) sp.dl_off = take(m.dpad[m.nl]);
  if (flavor == FL_FWD_FFJORD || flavor == FL_ADJ_FFJORD) sp.dl2_off = take(m.dpad[m.nl]);
  if (m.pre != B200_PRE_NONE) sp.x_off = take(m.dpad[0]);
  if (flavor == FL_ADJ_FFJORD) {
    for (int l = 0; l < m.nl; ++l) sp.hd_off[l] = take(m.dpad[l]);
    for (int l = 1; l < m.nl; ++l) sp.q_off[l] = take(m.dpad[l]);
    sp.ll_off = take(4);
  }
  if (rowthread) {                       // row-per-thread FFJORD forward: weights + biases and the transposed weights
    sp.roww_off = off;
    off += m.ppad;
    sp.rowwt_off = off;
    off += m.wt_off[m.nl - 1] + m.dpad[m.nl] * m.dpad[m.nl - 1];
    off = (off + 3) & ~3;
  }
  sp.total = off;
  return off * (int)sizeof(float);
}

template <int FL>
static cudaError_t launch_flavor(const SolveParams& P, int smem_bytes, cudaStream_t stream) {
  cudaError_t e = cudaFuncSetAttribute(solve_kernel<FL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  if (e != cudaSuccess) return e;
  void* args[] = {const_cast<SolveParams*>(&P)};
  return cudaLaunchCooperativeKernel(reinterpret_cast<void*>(solve_kernel<FL>), dim3(P.nctas), dim3(NTHREADS),
                                     args, (size_t)smem_bytes, stream);
}
template <int FL>
static cudaError_t occupancy_flavor(int smem_bytes, int* blocks_per_sm) {
  cudaError_t e = cudaFuncSetAttribute(solve_kernel<FL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  if (e != cudaSuccess) return e;
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, solve_kernel<FL>, NTHREADS, (size_t)smem_bytes);
}

static cudaError_t launch_any(int flavor, const SolveParams& P, int smem_bytes, cudaStream_t s) {
  if (P.tc.enabled) return b200_tc_launch(flavor, P, smem_bytes, s);
  switch (flavor) {
    case FL_FWD_NODE: return launch_flavor<FL_FWD_NODE>(P, smem_bytes, s);
    case FL_FWD_FFJORD: return launch_flavor<FL_FWD_FFJORD>(P, smem_bytes, s);
    case FL_ADJ_NODE_INTERP: return launch_flavor<FL_ADJ_NODE_INTERP>(P, smem_bytes, s);
    case FL_ADJ_NODE_BACKSOLVE: return launch_flavor<FL_ADJ_NODE_BACKSOLVE>(P, smem_bytes, s);
    default: return launch_flavor<FL_ADJ_FFJORD>(P, smem_bytes, s);
  }
}
template <int FL>
static cudaError_t static_smem_flavor(int* bytes) {
  cudaFuncAttributes a{};
  cudaError_t e = cudaFuncGetAttributes(&a, solve_kernel<FL>);
  *bytes = (int)a.sharedSizeBytes;
  return e;
}
// bytes of static __shared__ (StepShared) a flavour's kernel carries beside its dynamic shared memory
static cudaError_t static_smem_any(int flavor, int* bytes) {
  switch (flavor) {
    case FL_FWD_NODE: return static_smem_flavor<FL_FWD_NODE>(bytes);
    case FL_FWD_FFJORD: return static_smem_flavor<FL_FWD_FFJORD>(bytes);
    case FL_ADJ_NODE_INTERP: return static_smem_flavor<FL_ADJ_NODE_INTERP>(bytes);
    case FL_ADJ_NODE_BACKSOLVE: return static_smem_flavor<FL_ADJ_NODE_BACKSOLVE>(bytes);
    default: return static_smem_flavor<FL_ADJ_FFJORD>(bytes);
  }
}
static cudaError_t occupancy_any(int flavor, int smem_bytes, int* n) {
  switch (flavor) {
    case FL_FWD_NODE: return occupancy_flavor<FL_FWD_NODE>(smem_bytes, n);
    case FL_FWD_FFJORD: return occupancy_flavor<FL_FWD_FFJORD>(smem_bytes, n);
    case FL_ADJ_NODE_INTERP: return occupancy_flavor<FL_ADJ_NODE_INTERP>(smem_bytes, n);
    case FL_ADJ_NODE_BACKSOLVE: return occupancy_flavor<FL_ADJ_NODE_BACKSOLVE>(smem_bytes, n);
    default: return occupancy_flavor<FL_ADJ_FFJORD>(smem_bytes, n);
  }
}

static int cudaFuncAttributes_ok(cudaError_t e) { return e == cudaSuccess ? 0 : 1; }
// FFJORD forward with every layer width <= 64: the row-per-thread RHS (device.cuh: rhs_ffjord_fwd_rows)
static bool rowthread_eligible(const MlpDesc& m, int flavor) {
  if (flavor != FL_FWD_FFJORD || m.pre != B200_PRE_NONE || getenv("B200_NO_ROWTHREAD")) return false;
  for (int l = 0; l <= m.nl; ++l)
    if (m.dpad[l] > 64) return false;
  return true;
}
static int max_width(const MlpDesc& m) {
  int w = 0;
  for (int l = 0; l <= m.nl; ++l) w = std::max(w, m.dpad[l]);
  return w;
}
// every width <= 64 and enough trajectories to give each SM a 128-row tile: the tile engine at TB = 128 beats both its
// own smaller tiles and the row-per-thread RHS (218 vs 249 ms on the cfg 3 forward)
static bool narrow_full_tiles(b200_ctx* ctx, const MlpDesc& m, int flavor, int B) {
  return (flavor == FL_FWD_FFJORD || flavor == FL_ADJ_FFJORD) && max_width(m) <= 64 && B >= 128 * ctx->num_sms;
}
// choose the tile (trajectories per CTA pass) and the grid
static int plan_tiles(b200_ctx* ctx, const MlpDesc& m, int flavor, int B, int tile_opt, SolveParams& P,
                      int* smem_bytes, bool exact_trace = false) {
  int static_bytes = 0;                        // static __shared__ StepShared lives beside the dynamic part
  if (cudaFuncAttributes_ok(static_smem_any(flavor, &static_bytes)) != 0)
    return set_err(ctx, B200_ERR_CUDA, "cudaFuncGetAttributes failed");
  const int budget = ctx->smem_optin - static_bytes;
  int best = 0;
  double best_cost = 1e300;
  bool rows = false;
  SmemPlan sp;
  if (tile_opt > 0) {
    if (tile_opt % 8) return set_err(ctx, B200_ERR_ARG, "opts.tile must be a multiple of 8");
    if (plan_smem(m, flavor, tile_opt, sp) > budget)
      return set_err(ctx, B200_ERR_UNSUPPORTED, "opts.tile=%d needs %d B of shared memory (> %d)", tile_opt,
                     plan_smem(m, flavor, tile_opt, sp), budget);
    best = tile_opt;
  } else if (rowthread_eligible(m, flavor) && (exact_trace || !narrow_full_tiles(ctx, m, flavor, B))) {
    rows = true;
    // one thread per trajectory: a tile is as many rows as the CTA has threads (fewer for small batches)
    for (int TB = NTHREADS; TB >= 32; TB >>= 1) {
      if (plan_smem(m, flavor, TB, sp, true) > budget) continue;
      best = TB;
      if ((B + TB - 1) / TB >= ctx->num_sms || TB == 32) break;     // keep every SM busy before growing the tile
    }
    if (!best) return set_err(ctx, B200_ERR_UNSUPPORTED, "row-per-thread FFJORD path: shared-memory plan does not fit");
  } else {
    for (int TB = 128; TB >= 8; TB >>= 1) {
      if (plan_smem(m, flavor, TB, sp) > budget) continue;
      const int ntiles = (B + TB - 1) / TB;
      const int waves = (ntiles + ctx->num_sms - 1) / ctx->num_sms;
      double cost = (double)waves * std::max(TB, 32);
      // narrow networks with plenty of rows: a tile's GEMM is (TB/8) x (width/4) register tiles over NTHREADS threads -- below
      // one per thread the idle threads are paid for (measured, cfg 3 FFJORD 2->64->64->2 at B = 65536: TB = 64 forward
      // 323 ms / adjoint 43.1 ms, TB = 128 218 ms / 33.7 ms)
      if (narrow_full_tiles(ctx, m, flavor, B)) cost /= std::min(1.0, (TB / 8) * (max_width(m) / 4) / (double)NTHREADS);
      if (cost < best_cost) { best_cost = cost; best = TB; }
    }
    if (!best)
      return set_err(ctx, B200_ERR_UNSUPPORTED,
                     "model too wide for the shared-memory resident path (8 rows need %d B > %d B)",
                     plan_smem(m, flavor, 8, sp), budget);
  }
  P.rowthread = rows ? 1 : 0;
  *smem_bytes = plan_smem(m, flavor, best, P.sp, P.rowthread != 0);
  P.TB = best;
  P.ntiles = (B + best - 1) / best;
  int per_sm = 0;
  cudaError_t e = occupancy_any(flavor, *smem_bytes, &per_sm);
  if (e != cudaSuccess) return set_err(ctx, B200_ERR_CUDA, "occupancy query: %s", cudaGetErrorString(e));
  if (per_sm < 1) return set_err(ctx, B200_ERR_CUDA, "kernel does not fit on an SM (smem %d B)", *smem_bytes);
  P.nctas = std::min({P.ntiles, ctx->num_sms * per_sm, MAX_CTAS});
  return 0;
}

static void fill_stats(b200_ctx* ctx, const SolveParams& P, float ms, b200_stats* st) {
  if (!st) return;
  const Ctrl* c = ctx->ctrl_host;
  memset(st, 0, sizeof *st);
  st->retcode = c->retcode;
  st->naccept = c->naccept;
  st->nreject = c->nreject;
  st->nf = c->nf;
  st->t_final = c->t_final;
  st->dt_last = c->dt_last;
  st->tie_dist = c->tie_dist;
  st->kernel_ms = ms;
  st->n_out = c->n_out;
  st->grid = P.nctas;
  st->tile = P.TB;
  st->engine = P.tc.enabled;
}

// run the persistent kernel, bring the control header back
static int run_solve(b200_ctx* ctx, int flavor, SolveParams& P, int smem_bytes, b200_stats* st) {
  P.flavor = flavor;
  P.ctrl = ctx->ctrl_dev;
  P.rank = ctx->rank;
  P.world = ctx->world;
  for (int r = 0; r < MAX_RANKS; ++r) P.peers[r] = ctx->peers[r];
  ctx->solve_seq += 1;
  P.xbase = ctx->solve_seq << 24;
  if (ctx->world > 1 && P.mlp.ppad > XMU_CAP)
    return set_err(ctx, B200_ERR_UNSUPPORTED, "multi-GPU adjoint supports up to %d packed parameters (model has %d)", XMU_CAP, P.mlp.ppad);
  P.trace = ctx->trace_dev;
  P.trace_cap = ctx->trace_cap;
  const size_t header = offsetof(Ctrl, partial);
  CU(cudaMemsetAsync(ctx->ctrl_dev, 0, header, ctx->stream));
  CU(cudaEventRecord(ctx->ev0, ctx->stream));
  CU(launch_any(flavor, P, smem_bytes, ctx->stream));
  CU(cudaEventRecord(ctx->ev1, ctx->stream));
  CU(cudaMemcpyAsync(ctx->ctrl_host, ctx->ctrl_dev, header, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
  fill_stats(ctx, P, ms, st);
  ctx->trace_n = std::min(ctx->trace_cap, ctx->ctrl_host->naccept + ctx->ctrl_host->nreject);
  const int rc = ctx->ctrl_host->retcode;
  if (rc == B200_ERR_TAPE_OVERFLOW)
    return set_err(ctx, rc, "dense tape overflow after %d accepted steps: raise opts.max_dense_steps",
                   ctx->ctrl_host->nsteps);
  if (rc == B200_ERR_TIMEOUT) return set_err(ctx, rc, "grid / peer barrier timed out inside the persistent kernel");
  return rc;
}

// stage a caller buffer onto the device (no-op for device pointers)
static int to_device(b200_ctx* ctx, const float* src, size_t n, int ptr_kind, const float** out,
                     std::vector<void*>& temps) {
  if (ptr_kind == B200_PTR_DEVICE || src == nullptr || n == 0) { *out = src; return 0; }
  float* d = nullptr;
  CU(dalloc(ctx, &d, n));
  temps.push_back(d);
  CU(cudaMemcpyAsync(d, src, n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  *out = d;
  return 0;
}
static int from_device(b200_ctx* ctx, float* dst, const float* src, size_t n, int ptr_kind) {
  if (!dst || n == 0 || dst == src) return 0;
  CU(cudaMemcpyAsync(dst, src, n * sizeof(float),
                     ptr_kind == B200_PTR_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, ctx->stream));
  return 0;
}
static void free_temps(b200_ctx* ctx, std::vector<void*>& temps) {
  for (void* p : temps) dfree(ctx, p);
  temps.clear();
}

extern "C" void b200_tape_free(b200_tape* t) {
  if (!t) return;
  b200_ctx* ctx = t->ctx;
  if (!ctx) { delete t; return; }            // its context is gone (and took the device memory with it)
  DeviceGuard guard(ctx->device);
  ctx->tapes.erase(std::remove(ctx->tapes.begin(), ctx->tapes.end(), t), ctx->tapes.end());
  dfree(ctx, t->wp); dfree(ctx, t->wtp); dfree(ctx, t->wimg); dfree(ctx, t->ring); dfree(ctx, t->k1_first);
  dfree(ctx, t->step_t); dfree(ctx, t->out_u); dfree(ctx, t->out_t); dfree(ctx, t->final_state);
  delete t;
}
extern "C" int b200_tape_nsteps(const b200_tape* t) { return t ? t->nsteps : B200_ERR_ARG; }

extern "C" int b200_tape_steps(b200_ctx* ctx, const b200_tape* t, float* ts, float* us, int ptr_kind) {
  if (!ctx || !t) return B200_ERR_ARG;
  if (!t->dense) return set_err(ctx, B200_ERR_ARG, "tape holds no dense record (sensealg != Interpolating)");
  if (t->ctx != ctx) return set_err(ctx, B200_ERR_ARG, "tape belongs to another (or a destroyed) context");
  DeviceGuard guard(ctx->device);
  const int n = t->nsteps;
  if (ts) { int rc = from_device(ctx, ts, t->step_t, n + 1, ptr_kind); if (rc) return rc; }
  if (us) {
    for (int s = 0; s <= n; ++s) {
      const float* src = t->ring + ((long long)(s % t->nslots) * 7) * t->vec_elems;
      int rc = from_device(ctx, us + (long long)s * t->vec_elems, src, (size_t)t->vec_elems, ptr_kind);
      if (rc) return rc;
    }
  }
  CU(cudaStreamSynchronize(ctx->stream));
  return 0;
}

// ------------------------------------------------------------------------------------------------ forward core
// u0_dev: [B][d_in] on the device.  Produces a tape that owns every device buffer of the solve.
static int forward_core(b200_ctx* ctx, const b200_model* model, int flavor, const float* p_dev,
                        const float* u0_dev, int d_in, int augment, int B, float t0, float t1,
                        const float* saveat_dev, int nsave, const b200_opts* o, b200_stats* st,
                        b200_tape** tape_out) {
  const MlpDesc& m = model->m;
  const int D = m.dims[0];
  const int NS = (flavor == FL_FWD_FFJORD) ? D + (o->ffjord_regularize ? 3 : 1) : D;
  b200_tape* t = new b200_tape();
  ctx->tapes.push_back(t);
  t->ctx = ctx; t->m = m; t->fwd_flavor = flavor; t->sensealg = o->sensealg;
  t->B = B; t->NS = NS; t->D = D; t->d_in = d_in; t->augment = augment; t->t0 = t0; t->t1 = t1;
  t->seed = o->seed; t->noise_per_call = o->noise_per_call; t->b_offset = o->b_offset;
  t->dense = (o->sensealg == B200_SENSE_INTERPOLATING || o->sensealg == B200_SENSE_QUADRATURE || o->sensealg == B200_SENSE_GAUSS) || o->save_everystep;
  if (o->save_everystep) nsave = 0;
  t->vec_elems = (long long)B * NS;
  const int n_out = (o->save_start ? 1 : 0) + nsave + (o->save_end ? 1 : 0);
  int rc = 0;
  SolveParams P{};
  int smem_bytes = 0;
  auto fail = [&](int code) { b200_tape_free(t); return code; };

  P.mlp = m;
  const bool want_tc = o->path != 1 && tc_eligible(m, flavor) && (o->path == 2 || B >= 64);
  if (o->path == 2 && !want_tc)
    return fail(set_err(ctx, B200_ERR_UNSUPPORTED, "opts.path = 2: model shape / flavour is not eligible for the tcgen05 engine"));
  if (want_tc) {
    memset(&P.sp, 0, sizeof P.sp);
    P.TB = TC_ROWS;
    P.ntiles = (B + TC_ROWS - 1) / TC_ROWS;
    P.nctas = std::min({P.ntiles, ctx->num_sms, MAX_CTAS});
    smem_bytes = TC_SMEM;
    unsigned int off = 0;
    for (int l = 0; l < m.nl; ++l) {
      t->tc.w_off[l] = off;
      off += (unsigned int)(m.dims[l] / 16) * 2u * m.dims[l + 1] * 64u;
    }
    const unsigned int fwd_bytes = off;
    const bool want_t = o->sensealg != B200_SENSE_NONE && tc_eligible(m, FL_ADJ_NODE_BACKSOLVE);
    if (want_t)
      for (int l = 0; l < m.nl; ++l) {
        t->tc.t_off[l] = off - fwd_bytes;
        off += (unsigned int)(m.dims[l + 1] / 16) * 2u * m.dims[l] * 64u;
      }
    t->tc.enabled = 1;
    cudaError_t e2 = dalloc(ctx, &t->wimg, (size_t)off);
    if (e2 != cudaSuccess) return fail(set_err(ctx, B200_ERR_CUDA, "wimg alloc: %s", cudaGetErrorString(e2)));
    t->tc.wimg = t->wimg;
    t->tc.wimg_t = want_t ? t->wimg + fwd_bytes : nullptr;
    pack_tc_kernel<<<dim3(64, m.nl), 256, 0, ctx->stream>>>(m, t->tc, p_dev, t->wimg);
    if (want_t) pack_tc_t_kernel<<<dim3(64, m.nl), 256, 0, ctx->stream>>>(m, t->tc, p_dev, t->wimg + fwd_bytes);
  } else if ((rc = plan_tiles(ctx, m, flavor, B, o->tile, P, &smem_bytes, flavor == FL_FWD_FFJORD && o->ffjord_exact_trace)) != 0) return fail(rc);
  P.tc = t->tc;

  // dense tape capacity
  int cap = 0;
  if (t->dense) {
    cap = o->max_dense_steps;
    if (cap <= 0) {
      size_t free_b = 0, total_b = 0;
      cudaMemGetInfo(&free_b, &total_b);
      const double per_step = 7.0 * (double)t->vec_elems * sizeof(float);
      const double budget = std::min(0.5 * (double)free_b, 24e9);
      cap = (int)std::max(16.0, std::min((double)o->maxiters, std::min(4096.0, budget / per_step)));
    }
    t->nslots = cap + 1;
  } else {
    t->nslots = 2;
  }
  cudaError_t e;
#define TRY(x) if ((e = (x)) != cudaSuccess) return fail(set_err(ctx, B200_ERR_CUDA, "%s: %s", #x, cudaGetErrorString(e)))
  TRY(dalloc(ctx, &t->wp, (size_t)m.ppad));
  TRY(dalloc(ctx, &t->wtp, (size_t)model->wt_total));
  TRY(dalloc(ctx, &t->ring, (size_t)t->nslots * 7 * t->vec_elems));
  TRY(dalloc(ctx, &t->k1_first, (size_t)t->vec_elems));
  TRY(dalloc(ctx, &t->step_t, (size_t)cap + 2));
  TRY(dalloc(ctx, &t->out_u, (size_t)std::max(n_out, 1) * t->vec_elems));
  TRY(dalloc(ctx, &t->out_t, (size_t)n_out + 1));
  TRY(dalloc(ctx, &t->final_state, (size_t)t->vec_elems));
  pack_params_kernel<<<dim3(32, m.nl), 256, 0, ctx->stream>>>(m, p_dev, t->wp, t->wtp);
  TRY(cudaGetLastError());

  P.wp = t->wp; P.wtp = t->wtp;
  P.B = B; P.NV = NS; P.D = D; P.NS = NS;
  P.n_norm = (long long)B * ctx->world * NS;
  P.t0 = t0; P.t1 = t1; P.abstol = o->abstol; P.reltol = o->reltol; P.dt0_user = o->dt0;
  P.maxiters = o->maxiters;
  P.ring = t->ring; P.nslots = t->nslots; P.vec_elems = t->vec_elems; P.k1_first = t->k1_first;
  P.final_state = t->final_state;
  P.u0 = u0_dev; P.d_in = d_in;
  P.saveat = saveat_dev; P.nsave = nsave; P.save_start = o->save_start; P.save_end = o->save_end;
  P.out_u = t->out_u; P.out_t = t->out_t; P.step_t = t->step_t; P.dense = t->dense; P.tape_cap = cap;
  P.seed = o->seed; P.noise_per_call = o->noise_per_call; P.b_offset = o->b_offset;
  P.exact_trace = (flavor == FL_FWD_FFJORD && o->ffjord_exact_trace) ? 1 : 0;
  if (P.exact_trace && !P.rowthread)
    return fail(set_err(ctx, B200_ERR_UNSUPPORTED, "FFJORD monte_carlo = false (exact trace) is built for networks whose widths are all <= 64"));
  t->exact_trace = P.exact_trace;
  P.regularize = (flavor == FL_FWD_FFJORD && o->ffjord_regularize) ? 1 : 0;
  t->regularize = P.regularize;
  if (o->save_everystep) { P.save_start = 0; P.save_end = 0; }
  rc = run_solve(ctx, flavor, P, smem_bytes, st);
  if (rc < 0) return fail(rc);
  t->nsteps = ctx->ctrl_host->nsteps;
  t->n_out = ctx->ctrl_host->n_out;
  t->ts.resize(t->n_out);
  if (o->save_everystep) {
    // saved points = accepted steps (minus the ends the caller declined)
    std::vector<float> st_h(t->nsteps + 1);
    if (cudaMemcpyAsync(st_h.data(), t->step_t, st_h.size() * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
        cudaStreamSynchronize(ctx->stream) != cudaSuccess)
      return fail(set_err(ctx, B200_ERR_CUDA, "copying step times failed"));
    t->ts.clear();
    for (int s = 0; s <= t->nsteps; ++s) {
      if (s == 0 && !o->save_start) continue;
      if (s == t->nsteps && !o->save_end && t->nsteps > 0) continue;
      t->ts.push_back(st_h[s]);
    }
    t->n_out = (int)t->ts.size();
    if (st) st->n_out = t->n_out;
  } else if (t->n_out) {
    if (cudaMemcpyAsync(t->ts.data(), t->out_t, t->n_out * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
        cudaStreamSynchronize(ctx->stream) != cudaSuccess)
      return fail(set_err(ctx, B200_ERR_CUDA, "copying saved times failed"));
  }
#undef TRY
  *tape_out = t;
  return rc;
}

static int check_common(b200_ctx* ctx, const b200_model* model, const b200_opts* o, int B) {
  if (!ctx) return B200_ERR_ARG;
  if (!model || !o) return set_err(ctx, B200_ERR_ARG, "null model/opts");
  if (B < 1) return set_err(ctx, B200_ERR_ARG, "batch %d", B);
  if (!(o->abstol > 0.f) || !(o->reltol > 0.f)) return set_err(ctx, B200_ERR_ARG, "abstol/reltol must be > 0");
  if (o->maxiters < 1) return set_err(ctx, B200_ERR_ARG, "maxiters %d", o->maxiters);
  if (o->ptr_kind != B200_PTR_HOST && o->ptr_kind != B200_PTR_DEVICE) return set_err(ctx, B200_ERR_ARG, "ptr_kind");
  if (model->m.dims[0] != model->m.dims[model->m.nl])
    return set_err(ctx, B200_ERR_ARG, "model maps %d -> %d: an ODE right-hand side must preserve the state size",
                   model->m.dims[0], model->m.dims[model->m.nl]);
  return 0;
}

extern "C" int b200_node_forward(b200_ctx* ctx, const b200_model* model, const float* p, const float* u0,
                                 int d_in, int augment, int B, float t0, float t1, const float* saveat,
                                 int nsave, const b200_opts* o, float* out_u, float* out_t,
                                 b200_stats* stats, b200_tape** tape) {
  int rc = check_common(ctx, model, o, B);
  if (rc) return rc;
  DeviceGuard guard(ctx->device);
  const MlpDesc& m = model->m;
  if (!p || !u0) return set_err(ctx, B200_ERR_ARG, "null p/u0");
  if (o->save_everystep && !tape)
    return set_err(ctx, B200_ERR_ARG, "save_everystep needs a tape: the accepted steps are fetched with b200_tape_steps");
  if (d_in < 1 || augment < 0 || d_in + augment != m.dims[0])
    return set_err(ctx, B200_ERR_ARG, "state size mismatch: d_in %d + augment %d != model input %d", d_in, augment, m.dims[0]);
  if (nsave < 0 || (nsave > 0 && !saveat)) return set_err(ctx, B200_ERR_ARG, "saveat");
  if (o->sensealg < B200_SENSE_NONE || o->sensealg > B200_SENSE_GAUSS)
    return set_err(ctx, B200_ERR_UNSUPPORTED, "sensealg %d (supported: Interpolating, Backsolve, Quadrature, Gauss adjoints)", o->sensealg);
  if (o->save_everystep && o->sensealg == B200_SENSE_BACKSOLVE)
    return set_err(ctx, B200_ERR_UNSUPPORTED, "save_everystep with BacksolveAdjoint: pass saveat, or use InterpolatingAdjoint");
  std::vector<void*> temps;
  const float *p_d, *u0_d, *sv_d;
  if ((rc = to_device(ctx, p, (size_t)m.nparams, o->ptr_kind, &p_d, temps))) return rc;
  if ((rc = to_device(ctx, u0, (size_t)B * d_in, o->ptr_kind, &u0_d, temps))) return rc;
  // saveat is tiny and always needed by the kernel: copy unless already on the device
  if ((rc = to_device(ctx, saveat, (size_t)nsave, o->ptr_kind, &sv_d, temps))) return rc;
  b200_tape* t = nullptr;
  rc = forward_core(ctx, model, FL_FWD_NODE, p_d, u0_d, d_in, augment, B, t0, t1, sv_d, nsave, o, stats, &t);
  if (rc >= 0) {
    int rc2 = 0;
    if (out_u && !o->save_everystep && (rc2 = from_device(ctx, out_u, t->out_u, (size_t)t->n_out * t->vec_elems, o->ptr_kind))) rc = rc2;
    if (!rc2 && out_t && !o->save_everystep) {
      if (o->ptr_kind == B200_PTR_HOST) memcpy(out_t, t->ts.data(), t->ts.size() * sizeof(float));
      else if ((rc2 = from_device(ctx, out_t, t->out_t, (size_t)t->n_out, o->ptr_kind))) rc = rc2;
    }
    cudaStreamSynchronize(ctx->stream);
    if (tape && (o->sensealg != B200_SENSE_NONE || o->save_everystep) && rc2 == 0) *tape = t;
    else { b200_tape_free(t); if (tape) *tape = nullptr; }
  } else if (tape) *tape = nullptr;
  free_temps(ctx, temps);
  return rc;
}

// ------------------------------------------------------------------------------------------------ adjoint core
// dL_dev: [n_out][B][NS] on the device.  Leaves lambda(t0) in *lam_out ([B][NV], caller frees) and the
// packed mu(t0) in *mu_out.
static int adjoint_core(b200_ctx* ctx, b200_tape* t, int flavor, const float* dL_dev, const b200_opts* o,
                        float** lam_out, float** mu_out, b200_stats* st) {
  const MlpDesc& m = t->m;
  if (m.act[m.nl - 1] != B200_ACT_IDENTITY)
    return set_err(ctx, B200_ERR_UNSUPPORTED, "adjoint needs an identity output layer (Dense(h => d)); got act %d", m.act[m.nl - 1]);
  if (flavor == FL_ADJ_FFJORD && m.pre != B200_PRE_NONE)
    return set_err(ctx, B200_ERR_UNSUPPORTED, "FFJORD adjoint does not support a pre-layer");
  SolveParams P{};
  int smem_bytes = 0, rc;
  P.mlp = m;
  uint8_t* scratch = nullptr;
  const bool want_tc = o->path != 1 && t->tc.enabled && t->tc.wimg_t && tc_eligible(m, flavor);
  if (o->path == 2 && !want_tc)
    return set_err(ctx, B200_ERR_UNSUPPORTED, "opts.path = 2: adjoint not eligible for the tcgen05 engine (shape, or forward ran on the FFMA path)");
  if (want_tc) {
    memset(&P.sp, 0, sizeof P.sp);
    P.TB = TC_ROWS;
    P.ntiles = (t->B + TC_ROWS - 1) / TC_ROWS;
    P.nctas = std::min({P.ntiles, ctx->num_sms, MAX_CTAS});
    smem_bytes = TC_SMEM;
    P.tc = t->tc;
    unsigned int off = 0;                     // per-CTA fp32 stash: 16 KB per 32-feature chunk of a_l and delta_l
    for (int l = 0; l < m.nl; ++l) {
      P.tc.ncols_a[l] = (unsigned short)(((m.dims[l] / 32) + 3) & ~3);       // wgrad M groups of 128 features
      P.tc.sa_off[l] = off;
      off += (unsigned int)(m.dims[l] / 32) * 16384u;
    }
    for (int l = 1; l <= m.nl; ++l) {
      P.tc.ncols_d[l] = (unsigned short)(m.dims[l] / 32);
      P.tc.sd_off[l] = off;
      off += (unsigned int)(m.dims[l] / 32) * 16384u;
    }
    P.tc.scratch_cta = off;
    cudaError_t e2 = dalloc(ctx, &scratch, (size_t)off * P.nctas);
    if (e2 != cudaSuccess) return set_err(ctx, B200_ERR_CUDA, "adjoint scratch alloc: %s", cudaGetErrorString(e2));
    P.tc.scratch = scratch;
  } else if ((rc = plan_tiles(ctx, m, flavor, t->B, o->tile, P, &smem_bytes)) != 0) return rc;
  const int NS = t->NS;
  const int NV = (flavor == FL_ADJ_NODE_BACKSOLVE) ? 2 * NS : NS;
  const long long vec = (long long)t->B * NV;
  float *ring = nullptr, *k1 = nullptr, *fin = nullptr, *slabs = nullptr, *mu = nullptr, *muo = nullptr, *ts_d = nullptr;
  cudaError_t e;
  auto cleanup = [&]() { dfree(ctx, ring); dfree(ctx, k1); dfree(ctx, slabs); dfree(ctx, mu); dfree(ctx, ts_d); dfree(ctx, scratch); };
#define TRY(x) if ((e = (x)) != cudaSuccess) { cleanup(); dfree(ctx, fin); dfree(ctx, muo); return set_err(ctx, B200_ERR_CUDA, "%s: %s", #x, cudaGetErrorString(e)); }
  TRY(dalloc(ctx, &ring, (size_t)2 * 7 * vec));
  TRY(dalloc(ctx, &k1, (size_t)vec));
  TRY(dalloc(ctx, &fin, (size_t)vec));
  TRY(dalloc(ctx, &slabs, (size_t)P.nctas * NSLAB * m.ppad));
  TRY(dalloc(ctx, &mu, (size_t)2 * m.ppad));
  TRY(dalloc(ctx, &muo, (size_t)m.ppad));
  TRY(dalloc(ctx, &ts_d, t->ts.size() + 1));
  TRY(cudaMemsetAsync(slabs, 0, (size_t)P.nctas * NSLAB * m.ppad * sizeof(float), ctx->stream));
  TRY(cudaMemsetAsync(mu, 0, (size_t)2 * m.ppad * sizeof(float), ctx->stream));
  TRY(cudaMemsetAsync(muo, 0, (size_t)m.ppad * sizeof(float), ctx->stream));
  if (!t->ts.empty()) TRY(cudaMemcpyAsync(ts_d, t->ts.data(), t->ts.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));

  P.wp = t->wp; P.wtp = t->wtp;
  P.B = t->B; P.NV = NV; P.D = t->D; P.NS = NS;
  // Quadrature / Gauss adjoints: lambda is integrated on its own (mu is a quadrature of lambda^T df/dp, not a state of
  // the adjoint ODE): the parameter-gradient block stays out of the error norm, so the step sequence is lambda's alone
  P.mu_in_norm = (t->sensealg == B200_SENSE_QUADRATURE || t->sensealg == B200_SENSE_GAUSS) ? 0 : 1;
  P.n_norm = (long long)t->B * ctx->world * NV + (P.mu_in_norm ? m.nparams : 0);
  P.t0 = t->t1; P.t1 = t->t0;          // integrate backwards
  P.abstol = o->abstol; P.reltol = o->reltol; P.dt0_user = o->dt0; P.maxiters = o->maxiters;
  P.ring = ring; P.nslots = 2; P.vec_elems = vec; P.k1_first = k1; P.final_state = fin;
  P.f_ring = t->ring; P.f_nslots = t->nslots; P.f_nsteps = t->nsteps; P.f_vec_elems = t->vec_elems;
  P.f_k1_first = t->k1_first; P.f_step_t = t->step_t;
  P.f_us = t->out_u; P.f_final = t->final_state; P.f_ts = ts_d; P.n_ts = (int)t->ts.size(); P.dL = dL_dev;
  P.slabs = slabs; P.mu = mu; P.mu_out = muo;
  P.seed = t->seed; P.noise_per_call = t->noise_per_call; P.b_offset = t->b_offset;
  P.exact_trace = (flavor == FL_ADJ_FFJORD) ? t->exact_trace : 0;
  P.regularize = (flavor == FL_ADJ_FFJORD) ? t->regularize : 0;
  rc = run_solve(ctx, flavor, P, smem_bytes, st);
  cleanup();
#undef TRY
  if (rc < 0) { dfree(ctx, fin); dfree(ctx, muo); return rc; }
  *lam_out = fin;
  *mu_out = muo;
  return rc;
}

extern "C" int b200_node_adjoint(b200_ctx* ctx, const b200_model* model, b200_tape* tape, const float* dL_du,
                                 const b200_opts* o, float* du0, float* dp, b200_stats* stats) {
  int rc = check_common(ctx, model, o, tape ? tape->B : 0);
  if (rc) return rc;
  DeviceGuard guard(ctx->device);
  if (!tape || !dL_du) return set_err(ctx, B200_ERR_ARG, "null tape/dL_du");
  if (tape->ctx != ctx) return set_err(ctx, B200_ERR_ARG, "tape belongs to another (or a destroyed) context");
  if (tape->fwd_flavor != FL_FWD_NODE) return set_err(ctx, B200_ERR_ARG, "tape was not recorded by b200_node_forward");
  if (tape->sensealg < B200_SENSE_INTERPOLATING || tape->sensealg > B200_SENSE_GAUSS)
    return set_err(ctx, B200_ERR_ARG, "tape recorded with sensealg %d", tape->sensealg);
  const int flavor = tape->sensealg == B200_SENSE_BACKSOLVE ? FL_ADJ_NODE_BACKSOLVE : FL_ADJ_NODE_INTERP;
  std::vector<void*> temps;
  const float* dL_d;
  if ((rc = to_device(ctx, dL_du, (size_t)tape->n_out * tape->vec_elems, o->ptr_kind, &dL_d, temps))) return rc;
  float *lam = nullptr, *mu = nullptr;
  rc = adjoint_core(ctx, tape, flavor, dL_d, o, &lam, &mu, stats);
  if (rc >= 0) {
    const MlpDesc& m = tape->m;
    const int NV = flavor == FL_ADJ_NODE_BACKSOLVE ? 2 * tape->NS : tape->NS;
    float *du0_d = nullptr, *dp_d = nullptr;
    cudaError_t e = cudaSuccess;
    if (du0) {
      if (o->ptr_kind == B200_PTR_DEVICE) du0_d = du0; else { e = dalloc(ctx, &du0_d, (size_t)tape->B * tape->d_in); temps.push_back(du0_d); }
      if (e == cudaSuccess) {
        const long long n = (long long)tape->B * tape->d_in;
        extract_cols_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(lam, tape->B, NV, tape->d_in, du0_d, 0);
        if (o->ptr_kind == B200_PTR_HOST) from_device(ctx, du0, du0_d, (size_t)n, B200_PTR_HOST);
      }
    }
    if (dp && e == cudaSuccess) {
      if (o->ptr_kind == B200_PTR_DEVICE) dp_d = dp; else { e = dalloc(ctx, &dp_d, (size_t)m.nparams); temps.push_back(dp_d); }
      if (e == cudaSuccess) {
        unpack_grad_kernel<<<dim3(32, m.nl), 256, 0, ctx->stream>>>(m, mu, dp_d, 0);
        if (o->ptr_kind == B200_PTR_HOST) from_device(ctx, dp, dp_d, (size_t)m.nparams, B200_PTR_HOST);
      }
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = set_err(ctx, B200_ERR_CUDA, "adjoint output: %s", cudaGetErrorString(e));
    dfree(ctx, lam);
    dfree(ctx, mu);
  }
  free_temps(ctx, temps);
  return rc;
}

// ------------------------------------------------------------------------------------------------ FFJORD
extern "C" int b200_ffjord_forward(b200_ctx* ctx, const b200_model* model, const float* p, const float* x,
                                   int B, float t0, float t1, const b200_opts* o_in, float* logpx,
                                   float* z_out, b200_stats* stats, b200_tape** tape) {
  int rc = check_common(ctx, model, o_in, B);
  if (rc) return rc;
  DeviceGuard guard(ctx->device);
  const MlpDesc& m = model->m;
  if (!p || !x) return set_err(ctx, B200_ERR_ARG, "null p/x");
  if (m.pre != B200_PRE_NONE) return set_err(ctx, B200_ERR_UNSUPPORTED, "FFJORD does not support a pre-layer");
  b200_opts o = *o_in;
  o.save_start = 0;                     // src/ffjord.jl:156: save_everystep=false, save_start=false, save_end=true
  o.save_end = 1;
  if (o.sensealg == B200_SENSE_BACKSOLVE) o.sensealg = B200_SENSE_INTERPOLATING;   // hard-wired at src/ffjord.jl:143
  const int D = m.dims[0];
  std::vector<void*> temps;
  const float *p_d, *x_d;
  if ((rc = to_device(ctx, p, (size_t)m.nparams, o.ptr_kind, &p_d, temps))) return rc;
  if ((rc = to_device(ctx, x, (size_t)B * D, o.ptr_kind, &x_d, temps))) return rc;
  b200_tape* t = nullptr;
  rc = forward_core(ctx, model, FL_FWD_FFJORD, p_d, x_d, D, 1, B, t0, t1, nullptr, 0, &o, stats, &t);
  if (rc >= 0) {
    cudaError_t e = cudaSuccess;
    if (logpx) {
      float* lp_d = nullptr;
      if (o.ptr_kind == B200_PTR_DEVICE) lp_d = logpx; else { e = dalloc(ctx, &lp_d, (size_t)B); temps.push_back(lp_d); }
      if (e == cudaSuccess) {
        ffjord_logpx_kernel<<<(B + 255) / 256, 256, 0, ctx->stream>>>(t->final_state, B, D, t->NS, lp_d);
        if (o.ptr_kind == B200_PTR_HOST) from_device(ctx, logpx, lp_d, (size_t)B, B200_PTR_HOST);
      }
    }
    if (z_out) from_device(ctx, z_out, t->final_state, (size_t)B * t->NS, o.ptr_kind);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = set_err(ctx, B200_ERR_CUDA, "ffjord output: %s", cudaGetErrorString(e));
    if (tape && o.sensealg != B200_SENSE_NONE && e == cudaSuccess) *tape = t;
    else { b200_tape_free(t); if (tape) *tape = nullptr; }
  } else if (tape) *tape = nullptr;
  free_temps(ctx, temps);
  return rc;
}

extern "C" int b200_ffjord_adjoint(b200_ctx* ctx, const b200_model* model, b200_tape* tape,
                                   const float* dL_dlogpx, const b200_opts* o, float* dx, float* dp,
                                   b200_stats* stats) {
  if (ctx && !dL_dlogpx) return set_err(ctx, B200_ERR_ARG, "null tape/dL_dlogpx");
  return b200_ffjord_adjoint_state(ctx, model, tape, dL_dlogpx, nullptr, o, dx, dp, stats);
}

extern "C" int b200_ffjord_adjoint_state(b200_ctx* ctx, const b200_model* model, b200_tape* tape,
                                         const float* dL_dlogpx, const float* dL_dz, const b200_opts* o, float* dx,
                                         float* dp, b200_stats* stats) {
  int rc = check_common(ctx, model, o, tape ? tape->B : 0);
  if (rc) return rc;
  DeviceGuard guard(ctx->device);
  if (!tape || (!dL_dlogpx && !dL_dz)) return set_err(ctx, B200_ERR_ARG, "null tape/dL_dlogpx");
  if (tape->ctx != ctx) return set_err(ctx, B200_ERR_ARG, "tape belongs to another (or a destroyed) context");
  if (tape->fwd_flavor != FL_FWD_FFJORD || !tape->dense)
    return set_err(ctx, B200_ERR_ARG, "tape was not recorded by b200_ffjord_forward with InterpolatingAdjoint");
  const int B = tape->B, D = tape->D, NS = tape->NS;
  const MlpDesc& m = tape->m;
  std::vector<void*> temps;
  const float *g_d = nullptr, *gz_d = nullptr;
  if (dL_dlogpx && (rc = to_device(ctx, dL_dlogpx, (size_t)B, o->ptr_kind, &g_d, temps))) return rc;
  if (dL_dz && (rc = to_device(ctx, dL_dz, (size_t)B * NS, o->ptr_kind, &gz_d, temps))) { free_temps(ctx, temps); return rc; }
  float* dL_d = nullptr;
  CU(dalloc(ctx, &dL_d, (size_t)B * NS));
  temps.push_back(dL_d);
  ffjord_seed_kernel<<<(B + 255) / 256, 256, 0, ctx->stream>>>(tape->final_state, g_d, gz_d, B, D, NS, dL_d);
  float *lam = nullptr, *mu = nullptr;
  rc = adjoint_core(ctx, tape, FL_ADJ_FFJORD, dL_d, o, &lam, &mu, stats);
  if (rc >= 0) {
    cudaError_t e = cudaSuccess;
    if (dx) {
      float* dx_d = nullptr;
      if (o->ptr_kind == B200_PTR_DEVICE) dx_d = dx; else { e = dalloc(ctx, &dx_d, (size_t)B * D); temps.push_back(dx_d); }
      if (e == cudaSuccess) {
        const long long n = (long long)B * D;
        extract_cols_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(lam, B, NS, D, dx_d, 0);
        if (o->ptr_kind == B200_PTR_HOST) from_device(ctx, dx, dx_d, (size_t)n, B200_PTR_HOST);
      }
    }
    if (dp && e == cudaSuccess) {
      float* dp_d = nullptr;
      if (o->ptr_kind == B200_PTR_DEVICE) dp_d = dp; else { e = dalloc(ctx, &dp_d, (size_t)m.nparams); temps.push_back(dp_d); }
      if (e == cudaSuccess) {
        unpack_grad_kernel<<<dim3(32, m.nl), 256, 0, ctx->stream>>>(m, mu, dp_d, 0);
        if (o->ptr_kind == B200_PTR_HOST) from_device(ctx, dp, dp_d, (size_t)m.nparams, B200_PTR_HOST);
      }
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = set_err(ctx, B200_ERR_CUDA, "ffjord adjoint output: %s", cudaGetErrorString(e));
    dfree(ctx, lam);
    dfree(ctx, mu);
  }
  free_temps(ctx, temps);
  return rc;
}

// ------------------------------------------------------------------------------------------------ multiple shooting
extern "C" int b200_group_ranges(int datasize, int groupsize, int* first, int* last, int max_groups) {
  if (!(2 <= groupsize && groupsize <= datasize)) return B200_ERR_ARG;   // DomainError in the reference
  int g = 0;
  for (int i = 1; i <= datasize - 1; i += groupsize - 1) {
    if (g < max_groups && first && last) {
      first[g] = i - 1;
      last[g] = std::min(datasize, i + groupsize - 1) - 1;
    }
    ++g;
  }
  return g;
}

// Sub-contexts for concurrent group solves: how many 128-trajectory-tile grids of `nctas` CTAs fit on the GPU.
static int ensure_subs(b200_ctx* ctx, int want) {
  while ((int)ctx->subs.size() < want) {
    cudaStream_t st = nullptr;
    if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) break;
    b200_ctx* sub = nullptr;
    if (b200_ctx_create(ctx->device, st, &sub) != 0 || !sub) { cudaStreamDestroy(st); break; }
    sub->owns_stream = true;
    ctx->subs.push_back(sub);
  }
  return (int)ctx->subs.size();
}
static int mshoot_lanes(b200_ctx* ctx, const MlpDesc& m, int B, int G, int path) {
  if (ctx->world > 1 || G < 2) return 1;                       // sharded solves exchange through the parent's mailbox
  const bool tc = path != 1 && tc_eligible(m, FL_FWD_NODE) && (path == 2 || B >= 64);
  const int rows = tc ? TC_ROWS : 64;
  const int ctas = std::max(1, (B + rows - 1) / rows);
  int lanes = std::min({G, 8, std::max(1, ctx->num_sms / ctas)});
  if (const char* e = getenv("B200_MSHOOT_LANES")) lanes = std::max(1, std::min(lanes, atoi(e)));
  return lanes <= 1 ? 1 : std::min(lanes, ensure_subs(ctx, lanes));
}

extern "C" int b200_mshoot_forward(b200_ctx* ctx, const b200_model* model, const float* p,
                                   const float* ode_data, const float* tsteps, int T, int B, int G,
                                   const int* first, const int* last, const b200_opts* o_in, float* preds,
                                   b200_stats* stats, b200_tape** tapes) {
  int rc = check_common(ctx, model, o_in, B);
  if (rc) return rc;
  DeviceGuard guard(ctx->device);
  if (!p || !ode_data || !tsteps || !first || !last || G < 1 || T < 2) return set_err(ctx, B200_ERR_ARG, "mshoot arguments");
  if (tapes) for (int g = 0; g < G; ++g) tapes[g] = nullptr;
  const MlpDesc& m = model->m;
  const int D = m.dims[0];
  b200_opts o = *o_in;
  o.save_start = 1;
  o.save_end = 1;
  std::vector<void*> temps;
  const float *p_d, *data_d;
  if ((rc = to_device(ctx, p, (size_t)m.nparams, o.ptr_kind, &p_d, temps))) return rc;
  if ((rc = to_device(ctx, ode_data, (size_t)T * B * D, o.ptr_kind, &data_d, temps))) return rc;
  std::vector<float> ts(T);
  if (o.ptr_kind == B200_PTR_DEVICE) {
    CU(cudaMemcpyAsync(ts.data(), tsteps, T * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
  } else memcpy(ts.data(), tsteps, T * sizeof(float));
  float* ts_d = nullptr;
  CU(dalloc(ctx, &ts_d, (size_t)T));
  temps.push_back(ts_d);
  CU(cudaMemcpyAsync(ts_d, ts.data(), T * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  int worst = 0;
  std::vector<long long> row0(G + 1, 0);
  for (int g = 0; g < G; ++g) {
    const int a = first[g], b = last[g];
    if (a < 0 || b >= T || b <= a) { free_temps(ctx, temps); return set_err(ctx, B200_ERR_ARG, "group %d range [%d,%d]", g, a, b); }
    row0[g + 1] = row0[g] + (b - a + 1);
  }
  auto solve_group = [&](b200_ctx* c, int g) -> int {
    const int a = first[g], b = last[g], len = b - a + 1;
    b200_tape* t = nullptr;
    int r = forward_core(c, model, FL_FWD_NODE, p_d, data_d + (long long)a * B * D, D, 0, B, ts[a], ts[b],
                         ts_d + a + 1, len - 2, &o, stats ? stats + g : nullptr, &t);
    if (r < 0) return r;
    if (preds) from_device(c, preds + row0[g] * B * D, t->out_u, (size_t)t->n_out * B * D, o.ptr_kind);
    if (tapes && o.sensealg != B200_SENSE_NONE) tapes[g] = t; else { cudaStreamSynchronize(c->stream); b200_tape_free(t); }
    return r;
  };
  if (o.max_dense_steps <= 0) {
    // the automatic dense-tape capacity takes up to half of the free memory per solve; G tapes are alive at once
    // (and several are sized concurrently), so split that budget between the groups up front
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    const double per_step = 7.0 * (double)B * D * sizeof(float);
    const double budget = std::min(0.5 * (double)free_b / G, 24e9);
    o.max_dense_steps = (int)std::max(16.0, std::min((double)o.maxiters, std::min(4096.0, budget / per_step)));
  }
  const int lanes = mshoot_lanes(ctx, m, B, G, o.path);
  if (lanes <= 1) {
    for (int g = 0; g < G; ++g) {
      rc = solve_group(ctx, g);
      if (rc < 0) break;
      worst = std::max(worst, rc);
    }
  } else {
    cudaStreamSynchronize(ctx->stream);                       // inputs staged on the parent stream
    std::vector<int> lane_rc(lanes, 0);
    std::vector<std::thread> th;
    for (int k = 0; k < lanes; ++k)
      th.emplace_back([&, k]() {
        b200_ctx* c = ctx->subs[k];
        cudaSetDevice(c->device);
        int w = 0;
        for (int g = k; g < G; g += lanes) {
          const int r = solve_group(c, g);
          if (r < 0) { w = r; break; }
          w = std::max(w, r);
        }
        cudaStreamSynchronize(c->stream);
        lane_rc[k] = w;
      });
    for (auto& t : th) t.join();
    for (int k = 0; k < lanes; ++k) {
      if (lane_rc[k] < 0) { rc = lane_rc[k]; ctx->err = ctx->subs[k]->err; break; }
      worst = std::max(worst, lane_rc[k]);
    }
  }
  cudaStreamSynchronize(ctx->stream);
  free_temps(ctx, temps);
  if (rc < 0 && tapes)                       // a group failed: release the tapes the other groups already produced
    for (int g = 0; g < G; ++g)
      if (tapes[g]) { b200_tape_free(tapes[g]); tapes[g] = nullptr; }
  return rc < 0 ? rc : worst;
}

extern "C" int b200_mshoot_adjoint(b200_ctx* ctx, const b200_model* model, b200_tape** tapes, int G,
                                   const float* dL_dpreds, const b200_opts* o, float* dp, float* d_data,
                                   b200_stats* stats) {
  if (!ctx || !tapes || G < 1 || !tapes[0]) return B200_ERR_ARG;
  int rc = check_common(ctx, model, o, tapes[0]->B);
  if (rc) return rc;
  DeviceGuard guard(ctx->device);
  for (int g = 0; g < G; ++g) {
    if (!tapes[g]) return set_err(ctx, B200_ERR_ARG, "tapes[%d] is null", g);
    b200_ctx* owner = tapes[g]->ctx;
    if (owner != ctx && std::find(ctx->subs.begin(), ctx->subs.end(), owner) == ctx->subs.end())
      return set_err(ctx, B200_ERR_ARG, "tapes[%d] belongs to another (or a destroyed) context", g);
  }
  const MlpDesc& m = model->m;
  const int B = tapes[0]->B, D = tapes[0]->D;
  std::vector<void*> temps;
  long long rows = 0;
  for (int g = 0; g < G; ++g) rows += tapes[g]->n_out;
  const float* dL_d;
  if ((rc = to_device(ctx, dL_dpreds, (size_t)rows * B * D, o->ptr_kind, &dL_d, temps))) return rc;
  float* dp_d = nullptr;
  if (o->ptr_kind == B200_PTR_DEVICE && dp) dp_d = dp; else { CU(dalloc(ctx, &dp_d, (size_t)m.nparams)); temps.push_back(dp_d); }
  CU(cudaMemsetAsync(dp_d, 0, m.nparams * sizeof(float), ctx->stream));
  int worst = 0;
  std::vector<long long> row0(G + 1, 0);
  for (int g = 0; g < G; ++g) row0[g + 1] = row0[g] + tapes[g]->n_out;
  // one group's adjoint on the context that owns its tape; its parameter gradient is added to dp_acc
  auto adj_group = [&](b200_ctx* c, int g, float* dp_acc, std::vector<void*>& tmp) -> int {
    b200_tape* t = tapes[g];
    const int flavor = t->sensealg == B200_SENSE_BACKSOLVE ? FL_ADJ_NODE_BACKSOLVE : FL_ADJ_NODE_INTERP;
    float *lam = nullptr, *mu = nullptr;
    int r = adjoint_core(c, t, flavor, dL_d + row0[g] * B * D, o, &lam, &mu, stats ? stats + g : nullptr);
    if (r < 0) return r;
    unpack_grad_kernel<<<dim3(32, m.nl), 256, 0, c->stream>>>(m, mu, dp_acc, 1);
    if (d_data) {
      // gradient w.r.t. this group's initial condition; groups overlap by one point so the
      // caller adds it at time index first[g]: we return per-group blocks [G][B][D]
      const int NV = flavor == FL_ADJ_NODE_BACKSOLVE ? 2 * D : D;
      float* dst = nullptr;
      cudaError_t e = cudaSuccess;
      if (o->ptr_kind == B200_PTR_DEVICE) dst = d_data + (long long)g * B * D;
      else { e = dalloc(c, &dst, (size_t)B * D); tmp.push_back(dst); }
      if (e == cudaSuccess) {
        const long long n = (long long)B * D;
        extract_cols_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(lam, B, NV, D, dst, 0);
        if (o->ptr_kind == B200_PTR_HOST) from_device(c, d_data + (long long)g * B * D, dst, (size_t)n, B200_PTR_HOST);
      }
    }
    dfree(c, lam);
    dfree(c, mu);
    return r;
  };
  std::vector<b200_ctx*> lanes;                        // distinct owners of the tapes (sub-contexts of a concurrent forward)
  for (int g = 0; g < G; ++g)
    if (std::find(lanes.begin(), lanes.end(), tapes[g]->ctx) == lanes.end()) lanes.push_back(tapes[g]->ctx);
  if (lanes.size() == 1 && lanes[0] == ctx) {
    for (int g = 0; g < G; ++g) {
      rc = adj_group(ctx, g, dp_d, temps);
      if (rc < 0) break;
      worst = std::max(worst, rc);
    }
  } else {
    cudaStreamSynchronize(ctx->stream);                 // dL staged and dp zeroed on the parent stream
    const int nl = (int)lanes.size();
    std::vector<int> lane_rc(nl, 0);
    std::vector<float*> lane_dp(nl, nullptr);
    std::vector<std::vector<void*>> lane_tmp(nl);
    std::vector<std::thread> th;
    for (int k = 0; k < nl; ++k)
      th.emplace_back([&, k]() {
        b200_ctx* c = lanes[k];
        cudaSetDevice(c->device);
        int w = 0;
        if (dalloc(c, &lane_dp[k], (size_t)m.nparams) != cudaSuccess) { lane_rc[k] = B200_ERR_CUDA; return; }
        cudaMemsetAsync(lane_dp[k], 0, m.nparams * sizeof(float), c->stream);
        for (int g = 0; g < G; ++g) {
          if (tapes[g]->ctx != c) continue;
          const int r = adj_group(c, g, lane_dp[k], lane_tmp[k]);
          if (r < 0) { w = r; break; }
          w = std::max(w, r);
        }
        cudaStreamSynchronize(c->stream);
        lane_rc[k] = w;
      });
    for (auto& t : th) t.join();
    for (int k = 0; k < nl; ++k) {
      if (lane_rc[k] < 0) { if (rc >= 0) { rc = lane_rc[k]; ctx->err = lanes[k]->err; } }
      else worst = std::max(worst, lane_rc[k]);
      if (lane_dp[k] && rc >= 0) add_into_kernel<<<(m.nparams + 255) / 256, 256, 0, ctx->stream>>>(dp_d, lane_dp[k], m.nparams);
    }
    cudaStreamSynchronize(ctx->stream);
    for (int k = 0; k < nl; ++k) {
      free_temps(lanes[k], lane_tmp[k]);
      dfree(lanes[k], lane_dp[k]);
    }
  }
  if (rc >= 0 && dp && o->ptr_kind == B200_PTR_HOST) from_device(ctx, dp, dp_d, (size_t)m.nparams, B200_PTR_HOST);
  cudaStreamSynchronize(ctx->stream);
  free_temps(ctx, temps);
  return rc < 0 ? rc : worst;
}

// ------------------------------------------------------------------------------------------------ training step / output helpers
extern "C" int b200_node_train_step(b200_ctx* ctx, const b200_model* model, float* p, float* adam_m, float* adam_v,
                                    int step, float lr, float beta1, float beta2, float eps, const float* u0, int d_in,
                                    int augment, int B, float t0, float t1, const float* saveat, int nsave,
                                    const float* data, const b200_opts* o, float* loss, b200_stats* fwd_stats,
                                    b200_stats* adj_stats) {
  int rc = check_common(ctx, model, o, B);
  if (rc) return rc;
  DeviceGuard guard(ctx->device);
  const MlpDesc& m = model->m;
  if (!p || !adam_m || !adam_v || !u0 || !data) return set_err(ctx, B200_ERR_ARG, "null p/adam state/u0/data");
  if (step < 1 || !(lr > 0.f)) return set_err(ctx, B200_ERR_ARG, "step must be >= 1 and lr > 0");
  if (d_in < 1 || augment < 0 || d_in + augment != m.dims[0]) return set_err(ctx, B200_ERR_ARG, "state size mismatch");
  if (o->sensealg < B200_SENSE_INTERPOLATING || o->sensealg > B200_SENSE_GAUSS) return set_err(ctx, B200_ERR_ARG, "sensealg");
  if (o->save_everystep) return set_err(ctx, B200_ERR_UNSUPPORTED, "train step needs saveat / the end state (save_everystep = 0)");
  const int n_out = (o->save_start ? 1 : 0) + nsave + (o->save_end ? 1 : 0);
  if (n_out < 1) return set_err(ctx, B200_ERR_ARG, "no saved point to compare with the data");
  const int D = m.dims[0];
  const size_t nsaved = (size_t)n_out * B * D;
  std::vector<void*> temps;
  const bool host = o->ptr_kind == B200_PTR_HOST;
  float *p_d = p, *m_d = adam_m, *v_d = adam_v;
  const float *u0_d, *sv_d, *data_d;
  cudaError_t e = cudaSuccess;
  if (host) {
    for (float** q : {&p_d, &m_d, &v_d}) {
      float* src = *q;
      CU(dalloc(ctx, q, (size_t)m.nparams));
      temps.push_back(*q);
      CU(cudaMemcpyAsync(*q, src, m.nparams * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    }
  }
  if ((rc = to_device(ctx, u0, (size_t)B * d_in, o->ptr_kind, &u0_d, temps))) return rc;
  if ((rc = to_device(ctx, saveat, (size_t)nsave, o->ptr_kind, &sv_d, temps))) return rc;
  if ((rc = to_device(ctx, data, nsaved, o->ptr_kind, &data_d, temps))) return rc;
  b200_tape* t = nullptr;
  rc = forward_core(ctx, model, FL_FWD_NODE, p_d, u0_d, d_in, augment, B, t0, t1, sv_d, nsave, o, fwd_stats, &t);
  if (rc != 0) {                                  // error, or a retcode other than Success: no update (the caller sees the code)
    if (t) b200_tape_free(t);
    free_temps(ctx, temps);
    return rc;
  }
  float *dL = nullptr, *loss_d = nullptr;
  double* partial = nullptr;
  const int nblk = 296;
  e = dalloc(ctx, &dL, nsaved);
  if (e == cudaSuccess) e = dalloc(ctx, &loss_d, 1);
  if (e == cudaSuccess) e = dalloc(ctx, &partial, (size_t)nblk);
  if (e != cudaSuccess) { b200_tape_free(t); free_temps(ctx, temps); return set_err(ctx, B200_ERR_CUDA, "train step workspace: %s", cudaGetErrorString(e)); }
  temps.push_back(dL); temps.push_back(loss_d); temps.push_back(partial);
  sse_loss_kernel<<<nblk, 256, 0, ctx->stream>>>(t->out_u, data_d, (long long)nsaved, dL, partial);
  sum_partials_kernel<<<1, 32, 0, ctx->stream>>>(partial, nblk, loss_d);
  const int flavor = o->sensealg == B200_SENSE_BACKSOLVE ? FL_ADJ_NODE_BACKSOLVE : FL_ADJ_NODE_INTERP;
  float *lam = nullptr, *mu = nullptr;
  rc = adjoint_core(ctx, t, flavor, dL, o, &lam, &mu, adj_stats);
  if (rc == 0) {
    float* g = nullptr;
    e = dalloc(ctx, &g, (size_t)m.nparams);
    if (e == cudaSuccess) {
      temps.push_back(g);
      unpack_grad_kernel<<<dim3(32, m.nl), 256, 0, ctx->stream>>>(m, mu, g, 0);
      const float c1 = 1.f - powf(beta1, (float)step), c2 = 1.f - powf(beta2, (float)step);
      adam_kernel<<<(m.nparams + 255) / 256, 256, 0, ctx->stream>>>(p_d, m_d, v_d, g, m.nparams, lr, beta1, beta2, eps, c1, c2);
      if (host) {
        from_device(ctx, p, p_d, (size_t)m.nparams, B200_PTR_HOST);
        from_device(ctx, adam_m, m_d, (size_t)m.nparams, B200_PTR_HOST);
        from_device(ctx, adam_v, v_d, (size_t)m.nparams, B200_PTR_HOST);
      }
    } else rc = set_err(ctx, B200_ERR_CUDA, "train step gradient buffer: %s", cudaGetErrorString(e));
  }
  if (rc >= 0 && loss) from_device(ctx, loss, loss_d, 1, o->ptr_kind);
  e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess && rc >= 0) rc = set_err(ctx, B200_ERR_CUDA, "train step: %s", cudaGetErrorString(e));
  dfree(ctx, lam);
  dfree(ctx, mu);
  b200_tape_free(t);
  free_temps(ctx, temps);
  return rc;
}

extern "C" int b200_dimmover(b200_ctx* ctx, const float* x, int ndims, const int* dims, int from, int to, float* out,
                             int* out_ndims, int* out_dims, int ptr_kind) {
  if (!ctx) return B200_ERR_ARG;
  if (!x || !out || !dims || ndims < 2 || ndims > 4) return set_err(ctx, B200_ERR_ARG, "dimmover: 2 <= ndims <= 4 and non-null buffers");
  if (from == 0 || to == 0) return set_err(ctx, B200_ERR_ARG, "dimmover: from and to must be non-zero (src/neural_de.jl:407)");
  const int f = (from > 0 ? from : ndims + 1 + from) - 1, tt = (to > 0 ? to : ndims + 1 + to) - 1;   // 0-based (Julia order)
  if (f < 0 || f >= ndims || tt < 0 || tt >= ndims) return set_err(ctx, B200_ERR_ARG, "dimmover: from/to out of range");
  DeviceGuard guard(ctx->device);
  // Julia column-major strides of x
  long long st[4], n = 1;
  for (int a = 0; a < ndims; ++a) { st[a] = n; n *= dims[a]; }
  // slices drop axis f; cat(...; dims = to): to == ndims (1-based) appends a new axis, otherwise it concatenates along
  // axis `to` of the (ndims-1)-dimensional slices
  PermDesc d{};
  int sl[3], nsl = 0;
  for (int a = 0; a < ndims; ++a) if (a != f) sl[nsl++] = a;
  int od[4], ond = 0;
  if (tt == ndims - 1) {
    for (int k = 0; k < nsl; ++k) { d.dim[d.nd] = dims[sl[k]]; d.stride[d.nd++] = st[sl[k]]; od[ond++] = dims[sl[k]]; }
    d.dim[d.nd] = dims[f]; d.stride[d.nd++] = st[f]; od[ond++] = dims[f];
  } else {
    for (int k = 0; k < nsl; ++k) {
      d.dim[d.nd] = dims[sl[k]]; d.stride[d.nd++] = st[sl[k]];
      if (k == tt) { d.dim[d.nd] = dims[f]; d.stride[d.nd++] = st[f]; od[ond++] = dims[sl[k]] * dims[f]; }
      else od[ond++] = dims[sl[k]];
    }
  }
  if (out_ndims) *out_ndims = ond;
  if (out_dims) for (int a = 0; a < ond; ++a) out_dims[a] = od[a];
  std::vector<void*> temps;
  const float* x_d;
  int rc = to_device(ctx, x, (size_t)n, ptr_kind, &x_d, temps);
  if (rc) return rc;
  float* o_d = out;
  if (ptr_kind == B200_PTR_HOST) { CU(dalloc(ctx, &o_d, (size_t)n)); temps.push_back(o_d); }
  strided_copy_kernel<<<(unsigned)std::min<long long>((n + 255) / 256, 148 * 8), 256, 0, ctx->stream>>>(d, x_d, o_d, n);
  if (ptr_kind == B200_PTR_HOST) from_device(ctx, out, o_d, (size_t)n, B200_PTR_HOST);
  CU(cudaStreamSynchronize(ctx->stream));
  free_temps(ctx, temps);
  return 0;
}

// ------------------------------------------------------------------------------------------------ multi-GPU plumbing
extern "C" int b200_ctx_ipc_export(b200_ctx* ctx, unsigned char handle[B200_IPC_HANDLE_BYTES]) {
  if (!ctx || !handle) return B200_ERR_ARG;
  DeviceGuard guard(ctx->device);
  if (!ctx->mailbox) {
    CU(cudaMalloc(&ctx->mailbox, sizeof(Mailbox)));
    CU(cudaMemset(ctx->mailbox, 0, sizeof(Mailbox)));
  }
  static_assert(sizeof(cudaIpcMemHandle_t) <= B200_IPC_HANDLE_BYTES, "handle size");
  cudaIpcMemHandle_t h;
  CU(cudaIpcGetMemHandle(&h, ctx->mailbox));
  memset(handle, 0, B200_IPC_HANDLE_BYTES);
  memcpy(handle, &h, sizeof h);
  return 0;
}

extern "C" int b200_ctx_ipc_import(b200_ctx* ctx, int rank, int world, const unsigned char* handles) {
  if (!ctx || !handles) return B200_ERR_ARG;
  if (world < 1 || world > MAX_RANKS || rank < 0 || rank >= world) return set_err(ctx, B200_ERR_ARG, "rank %d / world %d", rank, world);
  if (!ctx->mailbox) return set_err(ctx, B200_ERR_ARG, "call b200_ctx_ipc_export first");
  DeviceGuard guard(ctx->device);
  for (int r = 0; r < world; ++r) {
    if (r == rank) { ctx->peers[r] = ctx->mailbox; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + (size_t)r * B200_IPC_HANDLE_BYTES, sizeof h);
    void* ptr = nullptr;
    CU(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    ctx->peers[r] = reinterpret_cast<Mailbox*>(ptr);
  }
  // Exchange epochs are (solve sequence number << 24 | exchange index) and must agree on all ranks: solves a rank ran on
  // its own before joining (world == 1: no mailbox traffic) do not count.  From here on every rank must issue the same
  // sequence of solves on this context.
  if (ctx->world == 1) ctx->solve_seq = 0;
  ctx->rank = rank;
  ctx->world = world;
  return 0;
}
