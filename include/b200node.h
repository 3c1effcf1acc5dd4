/* b200node.h -- C ABI of libb200node.so: the B200-native (sm_100a) replacement for the batched
 * NeuralODE solve + adjoint hot path of SciML/DiffEqFlux.jl.
 *
 * Every entry point names the reference interface it replaces (paths relative to the DiffEqFlux.jl
 * checkout, v4.8.1).  The reference reaches this arithmetic through un-vendored Julia packages
 * (OrdinaryDiffEq Tsit5, SciMLSensitivity adjoints, Lux Dense); the library below fuses that whole
 * stack into persistent CUDA kernels.  Host code (Julia `ccall`, Python `ctypes`, C) sees only plain
 * pointers and sizes.
 *
 * Conventions
 *   - All arrays are float32.  A Julia (D,B) column-major matrix == C array [B][D]: trajectory b's
 *     state is contiguous at u[b*D .. b*D+D).  `Array(sol)` (D,B,T) == C array [T][B][D].
 *   - Parameters `p` are the flat ComponentArray of a Lux Chain of Dense layers:
 *     [vec(W1); b1; vec(W2); b2; ...], W_l is (out,in) column-major == C [in][out].
 *   - Pointers are host pointers unless opts->ptr_kind == B200_PTR_DEVICE.
 *   - Return value: >= 0 is a SciML-style retcode (0 Success, 1 MaxIters, 2 DtLessThanMin,
 *     3 Unstable); < 0 is an argument / CUDA error with text in b200_last_error().
 *   - There is NO CPU fallback: an unsupported configuration is an error.
 *   - One b200_ctx is bound to one CUDA device and one stream and is not thread-safe.
 */
#ifndef B200NODE_H
#define B200NODE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_ABI_VERSION 1

/* retcodes (SciMLBase.ReturnCode as used at src/multiple_shooting.jl:66-67) */
#define B200_SUCCESS 0
#define B200_MAXITERS 1
#define B200_DT_LESS_THAN_MIN 2
#define B200_UNSTABLE 3
/* errors */
#define B200_ERR_ARG (-1)
#define B200_ERR_CUDA (-2)
#define B200_ERR_UNSUPPORTED (-3)
#define B200_ERR_TAPE_OVERFLOW (-4)
#define B200_ERR_TIMEOUT (-5)

/* activations of Lux.Dense layers */
#define B200_ACT_IDENTITY 0
#define B200_ACT_TANH 1
#define B200_ACT_RELU 2
#define B200_ACT_SIGMOID 3
#define B200_ACT_SOFTPLUS 4
/* input pre-layer (docs/src/examples/neural_ode.md:32 uses `x -> x.^3`) */
#define B200_PRE_NONE 0
#define B200_PRE_CUBE 1

/* sensealg (src/DiffEqFlux.jl:17-22 re-exports; src/neural_de.jl:57 default) */
#define B200_SENSE_NONE (-1)          /* forward only, keep no tape */
#define B200_SENSE_INTERPOLATING 0    /* InterpolatingAdjoint: dense forward output kept in HBM */
#define B200_SENSE_BACKSOLVE 1        /* BacksolveAdjoint(checkpointing=true): saved states only */
#define B200_SENSE_QUADRATURE 2       /* QuadratureAdjoint (src/DiffEqFlux.jl:19-20): lambda is integrated alone against the dense
                                         forward output; dp = int lambda^T df/dp dt is a quadrature, not an ODE state (it has no
                                         part in the step control).  The quadrature uses the Runge-Kutta weights of the lambda
                                         solve's own stages (5th order) instead of the reference's adaptive Gauss-Kronrod rule */
#define B200_SENSE_GAUSS 3            /* GaussAdjoint (src/DiffEqFlux.jl:19-20): same engine as B200_SENSE_QUADRATURE */

#define B200_PTR_HOST 0
#define B200_PTR_DEVICE 1

typedef struct b200_ctx b200_ctx;
typedef struct b200_model b200_model;
typedef struct b200_tape b200_tape;

/* Solver keyword arguments that DiffEqFlux splats into `solve` (src/neural_de.jl:55-58). */
typedef struct b200_opts {
  float abstol;          /* default 1e-6 */
  float reltol;          /* default 1e-3 */
  float dt0;             /* 0 = automatic (Hairer) */
  int maxiters;          /* default 100000 */
  int save_start;        /* output u(t0) first */
  int save_end;          /* output u(t1) last */
  int sensealg;          /* B200_SENSE_*: which tape the forward pass records */
  int ptr_kind;          /* B200_PTR_HOST / B200_PTR_DEVICE for every data pointer of the call */
  int tile;              /* trajectories per CTA tile, 0 = automatic */
  int max_dense_steps;   /* capacity of the dense tape in accepted steps, 0 = automatic */
  uint64_t seed;         /* FFJORD Hutchinson noise seed */
  int noise_per_call;    /* 1 (default) = fresh noise for every RHS call as src/ffjord.jl:92 */
  int b_offset;          /* global index of this shard's first trajectory (noise stream, multi-GPU) */
  int save_everystep;    /* 1: the saved points are the accepted steps (`sol.t`); needs the dense tape: `tape` must be
                            non-NULL (B200_ERR_ARG otherwise) and the points are fetched with b200_tape_steps
                            (out_u/out_t of the call stay untouched) */
  int path;              /* RHS engine: 0 = automatic, 1 = FP32-FFMA kernels only, 2 = require the tcgen05
                            3xTF32 tensor-core engine (error if the model shape is not eligible) */
  int ffjord_exact_trace;/* FFJORD: 1 = monte_carlo = false (src/ffjord.jl:120-123): the trace of the Jacobian is exact (D
                            unit-vector passes per RHS call instead of one Hutchinson noise pass); networks with widths <= 64 */
  int ffjord_regularize; /* FFJORD: 1 = regularize = true (src/ffjord.jl:131-133,149-153): the state carries two more rows,
                          * d/dt = [sum(abs2, m(z)); ||e^T J||_2]; end state / cotangents are [B][D+3] */
} b200_opts;

typedef struct b200_stats {
  int retcode;
  int naccept;
  int nreject;
  int nf;                /* RHS evaluations (per batch, as sol.stats.nf) */
  float t_final;
  float dt_last;
  float tie_dist;        /* min |EEst - 1| over the run: how close any accept/reject was to a tie */
  float kernel_ms;       /* device time of the persistent kernel (CUDA events on the ctx stream) */
  int n_out;             /* number of saved time points written */
  int grid;              /* CTAs launched */
  int tile;              /* trajectories per tile used */
  int engine;            /* 0 = FP32 FFMA, 1 = tcgen05 3xTF32 */
  int reserved[4];
} b200_stats;

/* ---- context / model ------------------------------------------------------------------------- */
int b200_abi_version(void);
/* device: CUDA ordinal; stream: a cudaStream_t to run on, or NULL for the legacy default stream
 * (the stream PyTorch / CUDA.jl use by default, so no extra synchronisation is needed; a caller that works on a
 * non-blocking stream must pass that stream here, otherwise its pending writes to u0 / p are not ordered before the
 * solve).  Every entry point binds `device` for the duration of the call and restores the caller's device. */
int b200_ctx_create(int device, void* stream, b200_ctx** out);
void b200_ctx_destroy(b200_ctx* ctx);
const char* b200_last_error(b200_ctx* ctx);
int b200_ctx_sync(b200_ctx* ctx);
void b200_opts_default(b200_opts* opts);
/* Diagnostics for step-sequence parity: capacity >= 0 (re)sizes the per-ctx trace of attempted
 * steps (0 disables, -1 leaves it); then copies up to max_entries rows (t, dt, EEst, accepted) of the
 * LAST solve into `out` (host) and returns the number of rows. */
int b200_ctx_trace(b200_ctx* ctx, int capacity, float* out, int max_entries);

/* Chain(Dense(dims[0]=>dims[1], act[0]), ..., Dense(dims[n-1]=>dims[n], act[n-1])): the model a
 * NeuralODE / FFJORD wraps (src/neural_de.jl:35-45, src/ffjord.jl:52-73). */
int b200_model_create(b200_ctx* ctx, int nlayers, const int* dims, const int* acts, int pre,
                      b200_model** out);
void b200_model_destroy(b200_model* model);
int b200_model_nparams(const b200_model* model);

/* ---- NeuralODE ------------------------------------------------------------------------------- */
/* Replaces `(n::NeuralODE)(x, p, st)` (src/neural_de.jl:47-61) = solve(ODEProblem{false}(dudt, x,
 * tspan, p), Tsit5(); saveat, abstol, reltol, save_start, save_end, sensealg).
 *   u0      [B][d_in]; the state is zero-padded by `augment` rows per trajectory
 *           (AugmentedNDELayer, src/neural_de.jl:372-386); d_in + augment == dims[0].
 *   saveat  nsave INTERIOR output times in integration order (strictly between t0 and t1).
 *   out_u   [n_out][B][D], out_t [n_out], n_out = save_start + nsave + save_end.
 *   tape    if non-NULL receives the opaque forward record the adjoint needs (opts->sensealg). */
int b200_node_forward(b200_ctx* ctx, const b200_model* model, const float* p, const float* u0,
                      int d_in, int augment, int B, float t0, float t1, const float* saveat,
                      int nsave, const b200_opts* opts, float* out_u, float* out_t,
                      b200_stats* stats, b200_tape** tape);

/* Replaces the ChainRules pullback of that solve (SciMLSensitivity `_concrete_solve_adjoint`,
 * selected at src/neural_de.jl:57): backward Tsit5 solve of [lambda; mu; (y)] with
 * lambda += dL_du[i] at every saved time.
 *   dL_du [n_out][B][D] cotangent of out_u;  du0 [B][d_in];  dp [P]. */
int b200_node_adjoint(b200_ctx* ctx, const b200_model* model, b200_tape* tape, const float* dL_du,
                      const b200_opts* opts, float* du0, float* dp, b200_stats* stats);

/* Accepted-step record of a dense (Interpolating) tape: `sol.t` / `sol.u` of a save_everystep
 * solve.  ts [nsteps+1], us [nsteps+1][B][D] (either may be NULL). */
int b200_tape_nsteps(const b200_tape* tape);
int b200_tape_steps(b200_ctx* ctx, const b200_tape* tape, float* ts, float* us, int ptr_kind);
void b200_tape_free(b200_tape* tape);

/* ---- FFJORD ---------------------------------------------------------------------------------- */
/* Replaces `__forward_ffjord` (src/ffjord.jl:140-178) with regularize=false, monte_carlo=true,
 * basedist=nothing: augment one delta_logp row, solve to t1 with the Hutchinson RHS `__ffjord`
 * (src/ffjord.jl:81-136), logpx = logpz - delta_logp.
 *   x [B][D]; logpx [B]; z_out (nullable) [B][D+1] end state ([B][D+3] with opts->ffjord_regularize: the reference's
 *   lambda_1 / lambda_2 outputs, src/ffjord.jl:164-166, are slices of it the caller takes). */
int b200_ffjord_forward(b200_ctx* ctx, const b200_model* model, const float* p, const float* x,
                        int B, float t0, float t1, const b200_opts* opts, float* logpx,
                        float* z_out, b200_stats* stats, b200_tape** tape);
/* Pullback of the above (sensealg hard-wired to InterpolatingAdjoint, src/ffjord.jl:143).
 *   dL_dlogpx [B]; dx [B][D]; dp [P]. */
int b200_ffjord_adjoint(b200_ctx* ctx, const b200_model* model, b200_tape* tape,
                        const float* dL_dlogpx, const b200_opts* opts, float* dx, float* dp,
                        b200_stats* stats);
/* Same pullback with a cotangent on the whole end state as well (what Zygote hands the solve when the loss also uses
 * lambda_1 / lambda_2 of a regularize = true layer, src/ffjord.jl:164-166,177):
 *   dL_dlogpx [B] (nullable); dL_dz [B][NS] (nullable) cotangent of z_out, NS = D+1 or D+3; at least one is given. */
int b200_ffjord_adjoint_state(b200_ctx* ctx, const b200_model* model, b200_tape* tape,
                              const float* dL_dlogpx, const float* dL_dz, const b200_opts* opts, float* dx,
                              float* dp, b200_stats* stats);

/* ---- multiple shooting ----------------------------------------------------------------------- */
/* `group_ranges(datasize, groupsize)` (src/multiple_shooting.jl:231-239): writes 0-based
 * [first, last] inclusive index pairs; returns the number of groups or B200_ERR_ARG (DomainError). */
int b200_group_ranges(int datasize, int groupsize, int* first, int* last, int max_groups);
/* The grouped solves of `multiple_shoot` (src/multiple_shooting.jl:52-63): group g starts from
 * ode_data[.., first[g]], spans (tsteps[first[g]], tsteps[last[g]]) and saves at tsteps[first..last].
 *   ode_data [T][B][D]; preds: concatenated per group, [sum_g len_g][B][D]; tapes [G] (nullable);
 *   stats [G].  Returns the worst retcode (the `Inf` gate of :66-67 is the caller's). */
int b200_mshoot_forward(b200_ctx* ctx, const b200_model* model, const float* p,
                        const float* ode_data, const float* tsteps, int T, int B, int G,
                        const int* first, const int* last, const b200_opts* opts, float* preds,
                        b200_stats* stats, b200_tape** tapes);
/* Sum of the per-group pullbacks: dL_dpreds laid out like preds; dp [P] (summed over groups);
 * d_data (nullable) [G][B][D]: block g is the cotangent of group g's initial condition ode_data[first[g]];
 * neighbouring groups share a data point, so the caller ACCUMULATES block g into time index first[g]. */
int b200_mshoot_adjoint(b200_ctx* ctx, const b200_model* model, b200_tape** tapes, int G,
                        const float* dL_dpreds, const b200_opts* opts, float* dp, float* d_data,
                        b200_stats* stats);

/* ---- the caller's side of the path: one training iteration, output re-layout ------------------- */
/* One iteration of the docs' training loop (docs/src/examples/neural_ode.md:36-44,64-70:
 * `loss = sum(abs2, ode_data .- Array(prob_neuralode(u0, p, st)[1]))`, Zygote gradient, `OptimizationOptimisers.Adam(lr)`)
 * in ONE call, everything on the device: forward solve with saveat, loss and dL/du(t_i) = 2 (pred - data), adjoint,
 * Adam update of p in place.  With opts->ptr_kind == B200_PTR_DEVICE p / adam_m / adam_v stay resident in HBM across
 * iterations (no D*B*T host round trip per iteration).
 *   p, adam_m, adam_v [P] in/out (m = v = 0 before step 1);  step: 1-based iteration number (bias correction);
 *   data [n_out][B][D] laid out like out_u of b200_node_forward;  loss: receives the loss BEFORE the update.
 * A retcode other than Success returns without touching p (the reference's loop would see an Inf/NaN loss). */
int b200_node_train_step(b200_ctx* ctx, const b200_model* model, float* p, float* adam_m, float* adam_v, int step,
                         float lr, float beta1, float beta2, float eps, const float* u0, int d_in, int augment, int B,
                         float t0, float t1, const float* saveat, int nsave, const float* data,
                         const b200_opts* opts, float* loss, b200_stats* fwd_stats, b200_stats* adj_stats);

/* `DimMover(from, to)` (src/neural_de.jl:397-412): `cat(eachslice(x; dims = from)...; dims = to)` on the device.
 * x is a Julia column-major array of `ndims` (2..4) dimensions `dims` (so `Array(sol)` (D,B,T) is dims = {D,B,T});
 * from / to are 1-based, negative values count from the end (defaults of the reference: from = -2, to = -1, which
 * turns (D,B,T) into (D,T,B)).  out receives the same number of elements; out_ndims / out_dims the result shape. */
int b200_dimmover(b200_ctx* ctx, const float* x, int ndims, const int* dims, int from, int to, float* out,
                  int* out_ndims, int* out_dims, int ptr_kind);

/* ---- multi-GPU (one process per GPU; handles travel over torch.distributed / MPI / files) ----- */
/* Each rank exports a CUDA-IPC handle of its step mailbox, all-gathers the handles with its own
 * plumbing, and imports the peers'.  Afterwards every solve on this ctx shards the batch: the
 * per-step error-norm partials are exchanged inside the persistent kernel over NVLink peer memory. */
#define B200_IPC_HANDLE_BYTES 64
int b200_ctx_ipc_export(b200_ctx* ctx, unsigned char handle[B200_IPC_HANDLE_BYTES]);
int b200_ctx_ipc_import(b200_ctx* ctx, int rank, int world,
                        const unsigned char* handles /* [world][B200_IPC_HANDLE_BYTES] */);

#ifdef __cplusplus
}
#endif
#endif /* B200NODE_H */
