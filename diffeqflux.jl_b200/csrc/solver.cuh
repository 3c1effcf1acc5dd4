// solver.cuh -- the persistent integrator loop (one cooperative launch per solve).
// Controller arithmetic mirrors oracle/tsit5.py operation by operation (fp32, no FMA contraction).
#pragma once
#include "device.cuh"

namespace b200 {

// The tensor-core translation unit (tc_kernels.cu) launches TC_THREADS threads: 2 (or 4) worker warpgroups + one
// control warpgroup; setmaxnreg moves the control group's registers to the workers right after tc_init.
#ifdef B200_TC_TU
#define B200_KERNEL_ATTR __launch_bounds__(TC_THREADS, 1)
#else
// The FFJORD forward is compiled for two CTAs per SM (128 registers, ~0.5 KB of spills): its tile engine stalls 31 % of
// the time at the per-layer CTA barriers, which a second resident CTA hides (cfg 3 forward 219 -> 147 ms); its shared-
// memory plan (72 KB at 128-row tiles of 64-wide networks) leaves room for two.  Every other flavour needs the registers.
#ifdef B200_FFMA_OCC2            // experiment: every FP32-FFMA flavour at 128 registers
#define B200_KERNEL_ATTR __launch_bounds__(NTHREADS, 2)
#else
#define B200_KERNEL_ATTR __launch_bounds__(NTHREADS, (FL == FL_FWD_FFJORD ? 2 : 1))
#endif
#endif
template <int FL, bool TC = false>
__global__ void B200_KERNEL_ATTR solve_kernel(const __grid_constant__ SolveParams P) {
  extern __shared__ __align__(16) float sm[];
  __shared__ StepShared ss;
  using T = Traits<FL>;
  Solver<FL, TC> S(P, sm, ss);
  const int tid = threadIdx.x;
  const int cta = blockIdx.x;
  const MlpDesc& M = P.mlp;
  const int NV = P.NV, TB = P.TB, B = P.B;
  Ctrl* ctrl = P.ctrl;

  for (int i = tid; i < P.sp.total; i += blockDim.x) sm[i] = 0.f;
  const float tdir = P.t1 >= P.t0 ? 1.f : -1.f;
  const float dtmax = fabsf(__fsub_rn(P.t1, P.t0));
  const float inv_qmax = __fdiv_rn(1.f, 10.f), inv_qmin = __fdiv_rn(1.f, 0.2f);
  const float beta1 = (float)(7.0 / 50.0), beta2 = (float)(2.0 / 25.0), gamma = 0.9f;
  const double n_norm = (double)P.n_norm;

  // adjoint: does the last saved time coincide with the start of the backward solve?
  const bool init_jump = T::adj && P.n_ts > 0 && P.f_ts[P.n_ts - 1] == P.t0;
  if (tid == 0) {
    ss.t = P.t0; ss.dt = 0.f; ss.tnew = P.t0;
    ss.qold = 1e-4f; ss.q11 = 1.f; ss.EEst = 0.f;
    ss.accept = 0; ss.done = 0; ss.retcode = 0;
    ss.naccept = ss.nreject = ss.nf = ss.iter = 0;
    ss.step = 0; ss.first_step = 0; ss.si = 0;
    ss.ti = T::adj ? (init_jump ? P.n_ts - 2 : P.n_ts - 1) : -1;
    ss.mu_cur = 0; ss.out_idx = 0; ss.tie = 3.0e38f;
    ss.epoch = 0u; ss.xepoch = P.xbase; ss.mu_xe = P.xbase; ss.rhs_calls = 0u;
    ss.g1 = 5; ss.aborted = 0;
  }
  __syncthreads();
  if constexpr (TC) {
    S.tc_init();
    if (tid >= NTHREADS && tid < NTHREADS + 128) {      // control warpgroup: MMA issuer + weight loader
      b200tc::reg_dec<TC_REGS_CTRL>();
      S.tc_control();
      return;
    }
    b200tc::reg_inc<TC_REGS_WORK>();
    if (tid >= NTHREADS + 128) {                        // helper worker warpgroups (TC_WORKERS = 512)
      S.tc_helper();
      return;
    }
  }

#ifdef B200_TC_PROF
  long long qf[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const long long qf_start = clock64();
#endif
  if constexpr (FL == FL_FWD_FFJORD) {
    if (P.rowthread) S.stage_row_weights();
  }
  // ------------------------------------------------------------------ initial state (slot 0)
  float* U0 = S.slotv(0, 0);
  S.for_my_tiles([&](int tile) {
    const int base_row = tile * TB;
    for (int e = tid; e < TB * NV; e += NTHREADS) {
      const int r = e / NV, c = e - r * NV;
      const int b = base_row + r;
      if (b >= B) continue;
      float v = 0.f;
      if constexpr (!T::adj) {
        if (c < P.d_in) v = P.u0[(long long)b * P.d_in + c];
      } else {
        const int NS = P.NS;
        if (c < NS) {
          if (init_jump) v = P.dL[((long long)(P.n_ts - 1) * B + b) * NS + c];
        } else {
          v = P.f_final[(long long)b * NS + (c - NS)];
        }
      }
      U0[(long long)b * NV + c] = v;
      if constexpr (!T::adj) {
        if (P.save_start) P.out_u[(long long)b * NV + c] = v;
      }
    }
  });
  if constexpr (!T::adj) {
    if (P.save_start && cta == 0 && tid == 0) P.out_t[0] = P.t0;
    if (tid == 0 && P.save_start) ss.out_idx = 1;
    if (P.dense && cta == 0 && tid == 0) P.step_t[0] = P.t0;
  }
  S.sync();

  if (P.t0 != P.t1) {
    // ---------------------------------------------------------------- FSAL init
    if constexpr (T::interp) {
      if (tid == 0) S.locate(P.t0, ss.ii[0]);
      S.sync();
    }
    {
      SlabOut so{T::adj ? S.slab(ss.g1) : nullptr, nullptr, 0.f, nullptr, 0.f};
      S.eval_rhs_vec(U0, P.k1_first, &ss.ii[0], so);
    }
    // ---------------------------------------------------------------- initial dt (Hairer)
    if (P.dt0_user != 0.f) {
      if (tid == 0) ss.dt = tdir * fabsf(P.dt0_user);
      S.sync();
    } else {
      float* F0 = S.slotv(0, 1);
      float* F1 = S.slotv(0, 2);
      float* U1 = S.slotv(1, 0);
      const int g7 = 11 - ss.g1;
      {
        SlabOut so{T::adj ? S.slab(g7) : nullptr, nullptr, 0.f, nullptr, 0.f};
        S.eval_rhs_vec(U0, F0, &ss.ii[0], so);
      }
      double p0 = 0.0, p1 = 0.0;
      S.for_my_tiles([&](int tile) {
        const int base_row = tile * TB;
        float a0 = 0.f, a1 = 0.f;
        for (int e = tid; e < TB * NV; e += NTHREADS) {
          const int r = e / NV;
          if (base_row + r >= B) continue;
          const long long g = (long long)base_row * NV + e;
          const float u = U0[g];
          const float sk = fmaf(fabsf(u), P.reltol, P.abstol);
          const float x0 = u / sk, x1 = F0[g] / sk;
          a0 = fmaf(x0, x0, a0);
          a1 = fmaf(x1, x1, a1);
        }
        p0 += (double)a0;
        p1 += (double)a1;
      });
      if constexpr (T::adj) if (P.mu_in_norm) {
        S.barrier_reduce(0.0, 0.0);   // all slabs written
        p1 += S.mu_reduce(g7, 1.f, -1, 0.f, -1, [&](int, float tot, float) {
          const float x = tot / P.abstol;       // mu(t1) = 0  =>  sk = abstol
          return x * x;
        });
        S.mu_reduce_done();
      }
      S.barrier_reduce(p0, p1);
      if (tid == 0) {
        const float d0 = (float)sqrt(ss.sums[0] / n_norm), d1 = (float)sqrt(ss.sums[1] / n_norm);
        float dt0;
        if (d0 < 1e-5f || d1 < 1e-5f) dt0 = 1e-6f;
        else dt0 = __fdiv_rn(__fdiv_rn(d0, d1), 100.f);
        dt0 = fminf(dt0, dtmax);
        ss.EEst = d1;          // stash
        ss.q11 = dt0;          // stash
        ss.accept = (dt0 < __fmul_rn(10.f, 1.1920929e-07f)) ? 0 : 1;
        if (!ss.accept) ss.dt = tdir * 1e-6f;
      }
      S.sync();
      if (ss.accept) {
        const float dt0 = ss.q11;
        const float dt0_t = tdir * dt0;
        S.for_my_tiles([&](int tile) {
          const int base_row = tile * TB;
          for (int e = tid; e < TB * NV; e += NTHREADS) {
            const int r = e / NV;
            if (base_row + r >= B) continue;
            const long long g = (long long)base_row * NV + e;
            U1[g] = fmaf(dt0_t, F0[g], U0[g]);
          }
        });
        S.sync();
        if constexpr (T::interp) {
          if (tid == 0) S.locate(__fadd_rn(P.t0, dt0_t), ss.ii[1]);
          S.sync();
        }
        {
          SlabOut so{T::adj ? S.slab(0) : nullptr, nullptr, 0.f, nullptr, 0.f};
          S.eval_rhs_vec(U1, F1, &ss.ii[1], so);
        }
        double p2 = 0.0;
        S.for_my_tiles([&](int tile) {
          const int base_row = tile * TB;
          float a2 = 0.f;
          for (int e = tid; e < TB * NV; e += NTHREADS) {
            const int r = e / NV;
            if (base_row + r >= B) continue;
            const long long g = (long long)base_row * NV + e;
            const float sk = fmaf(fabsf(U0[g]), P.reltol, P.abstol);
            const float x = (F1[g] - F0[g]) / sk;
            a2 = fmaf(x, x, a2);
          }
          p2 += (double)a2;
        });
        if constexpr (T::adj) if (P.mu_in_norm) {
          S.barrier_reduce(0.0, 0.0);
          p2 += S.mu_reduce(0, 1.f, g7, -1.f, -1, [&](int, float tot, float) {
            const float x = tot / P.abstol;
            return x * x;
          });
          S.mu_reduce_done();
        }
        S.barrier_reduce(p2, 0.0);
        if (tid == 0) {
          const float d1 = ss.EEst;
          const float d2 = __fdiv_rn((float)sqrt(ss.sums[0] / n_norm), dt0);
          const float mx = fmaxf(d1, d2);
          double dt1;
          if (mx <= 1e-15f) dt1 = (double)fmaxf(1e-6f, __fmul_rn(dt0, 1e-3f));
          else {
            const float x32 = __fdiv_rn(-__fadd_rn(2.f, log10f(mx)), 5.f);
            dt1 = pow(10.0, (double)x32);
          }
          const float dt = (float)fmin(fmin(100.0 * (double)dt0, dt1), (double)dtmax);
          ss.dt = tdir * dt;
        }
        S.sync();
      }
      if (tid == 0) { ss.EEst = 0.f; ss.q11 = 1.f; ss.accept = 0; }
      S.sync();
    }
  }

  // ------------------------------------------------------------------ main loop
  while (P.t0 != P.t1) {
    S.sync();   // nobody still reads the previous iteration's ss fields
    // ---- loop header (thread 0): tstop, dt clamps, exits
    if (tid == 0) {
      float tstop = P.t1;
      if constexpr (T::adj) {
        if (ss.ti >= 0 && P.f_ts[ss.ti] != P.t1) tstop = P.f_ts[ss.ti];
      }
      ss.tnew = tstop;   // stash the current tstop
      if (ss.t == P.t1) ss.done = 1;
      else {
        ss.iter += 1;
        if (ss.iter > P.maxiters) { ss.retcode = 1; ss.done = 1; }
        else {
          float dt = tdir * fminf(fabsf(ss.dt), dtmax);
          dt = tdir * fminf(fabsf(dt), fabsf(__fsub_rn(tstop, ss.t)));
          ss.dt = dt;
          if (!isfinite(dt)) { ss.retcode = 3; ss.done = 1; }
          else if (fabsf(dt) <= eps32(ss.t)) { ss.retcode = 2; ss.done = 1; }
          else if (P.dense && ss.step >= P.tape_cap) { ss.retcode = -4; ss.done = 1; }
        }
      }
      if (!ss.done) {
        if constexpr (T::interp) {
          for (int s = 2; s <= 7; ++s)
            S.locate(__fadd_rn(ss.t, __fmul_rn(c_tab.c[s - 2], ss.dt)), ss.ii[s - 1]);
        }
      }
    }
    S.sync();
    if (ss.done || S.aborted()) break;
    const float t = ss.t, dt = ss.dt, tstop = ss.tnew;
    const int step = ss.step;
    const int g1 = ss.g1, g7 = 11 - g1;

    float* u = S.slotv(step, 0);
    float* unew = S.slotv(step + 1, 0);
    const float* kp[7];
    kp[0] = S.k1ptr(step);
    for (int j = 1; j < 7; ++j) kp[j] = S.slotv(step, j);

    double perr = 0.0;
    bool first_tile = true;
#ifdef B200_TC_PROF
    long long q_t0 = clock64();
#endif
    S.for_my_tiles([&](int tile) {
      for (int s = 2; s <= 7; ++s) {
        SlabOut so{nullptr, nullptr, 0.f, nullptr, 0.f};
        if constexpr (T::adj) {
          so = SlabOut{S.slab(s < 7 ? s - 2 : g7), nullptr, 0.f, nullptr, 0.f};   // G_s stored once, combined in the mu phase
        }
        S.stage_eval(tile, u, kp, c_tab.a[s - 1], s - 1, dt, s == 7 ? unew : nullptr, const_cast<float*>(kp[s - 1]),
                     &ss.ii[s - 1], so, !first_tile, ss.rhs_calls + (s - 2));
        S.sync();
      }
#ifdef B200_TC_PROF
      qf[0] += clock64() - q_t0; q_t0 = clock64();
#endif
      // embedded error estimate of this tile (tile rows are contiguous in every vector; NV % 4 == 0 on the
      // tensor-core path, so 16-byte loads there)
      const int base_row = tile * TB;
      float a = 0.f;
      auto err_term = [&](float k0, float k1, float k2, float k3, float k4, float k5, float k6, float uo, float un) {
        float acc = c_tab.bt[0] * k0;
        acc = fmaf(c_tab.bt[1], k1, acc); acc = fmaf(c_tab.bt[2], k2, acc); acc = fmaf(c_tab.bt[3], k3, acc);
        acc = fmaf(c_tab.bt[4], k4, acc); acc = fmaf(c_tab.bt[5], k5, acc); acc = fmaf(c_tab.bt[6], k6, acc);
        const float ut = dt * acc;
        const float sk = fmaf(fmaxf(fabsf(uo), fabsf(un)), P.reltol, P.abstol);
        const float x = ut / sk;
        return x * x;
      };
      if constexpr (TC) {
        const int rows = min(TB, B - base_row);
        const long long g0 = (long long)base_row * NV;
        for (int e = tid * 4; e < rows * NV; e += NTHREADS * 4) {
          const long long g = g0 + e;
          float4 kk[7];
#pragma unroll
          for (int j = 0; j < 7; ++j) kk[j] = *reinterpret_cast<const float4*>(kp[j] + g);
          const float4 uo = *reinterpret_cast<const float4*>(u + g), un = *reinterpret_cast<const float4*>(unew + g);
          a += err_term(kk[0].x, kk[1].x, kk[2].x, kk[3].x, kk[4].x, kk[5].x, kk[6].x, uo.x, un.x);
          a += err_term(kk[0].y, kk[1].y, kk[2].y, kk[3].y, kk[4].y, kk[5].y, kk[6].y, uo.y, un.y);
          a += err_term(kk[0].z, kk[1].z, kk[2].z, kk[3].z, kk[4].z, kk[5].z, kk[6].z, uo.z, un.z);
          a += err_term(kk[0].w, kk[1].w, kk[2].w, kk[3].w, kk[4].w, kk[5].w, kk[6].w, uo.w, un.w);
        }
      } else {
        for (int e = tid; e < TB * NV; e += NTHREADS) {
          const int r = e / NV;
          if (base_row + r >= B) continue;
          const long long g = (long long)base_row * NV + e;
          a += err_term(kp[0][g], kp[1][g], kp[2][g], kp[3][g], kp[4][g], kp[5][g], kp[6][g], u[g], unew[g]);
        }
      }
      perr += (double)a;
      first_tile = false;
    });

#ifdef B200_TC_PROF
    qf[1] += clock64() - q_t0; q_t0 = clock64();
#endif
    if constexpr (T::adj) if (P.mu_in_norm) {
      S.barrier_reduce(0.0, 0.0, false);   // every CTA's slabs of this step complete (local ordering only)
#ifdef B200_TC_PROF
      qf[2] += clock64() - q_t0; q_t0 = clock64();
#endif
      const float* mu = P.mu + (long long)ss.mu_cur * M.ppad;
      float* mun = P.mu + (long long)(1 - ss.mu_cur) * M.ppad;
      perr += S.mu_reduce(-2, 0.f, g1, 0.f, g7, [&](int j, float totA, float totE) {
        const float m0 = mu[j];
        const float m1 = fmaf(dt, totA, m0);
        mun[j] = m1;
        const float ut = dt * totE;
        const float sk = fmaf(fmaxf(fabsf(m0), fabsf(m1)), P.reltol, P.abstol);
        const float x = ut / sk;
        return x * x;
      });
      S.mu_reduce_done();
#ifdef B200_TC_PROF
      qf[3] += clock64() - q_t0; q_t0 = clock64();
#endif
    }
    S.barrier_reduce(perr, 0.0);
#ifdef B200_TC_PROF
    qf[4] += clock64() - q_t0; q_t0 = clock64();
    qf[6] += 1;
#endif
    if (S.aborted()) break;

    // ---- controller (thread 0)
    if (tid == 0) {
      ss.nf += 6;
      ss.rhs_calls += 6u;
      const float EEst = (float)sqrt(ss.sums[0] / n_norm);
      ss.EEst = EEst;
      if (!isfinite(EEst)) { ss.retcode = 3; ss.done = 1; ss.accept = 0; }
      else {
        float q;
        if (EEst == 0.f) q = inv_qmax;
        else {
          ss.q11 = fastpowf(EEst, beta1);
          q = __fdiv_rn(ss.q11, fastpowf(ss.qold, beta2));
          q = fmaxf(inv_qmax, fminf(inv_qmin, __fdiv_rn(q, gamma)));
        }
        ss.tie = fminf(ss.tie, fabsf(EEst - 1.f));
        if (P.trace && cta == 0 && ss.iter <= P.trace_cap) {
          float* tr = P.trace + (long long)(ss.iter - 1) * 4;
          tr[0] = t; tr[1] = dt; tr[2] = EEst; tr[3] = EEst <= 1.f ? 1.f : 0.f;
        }
        if (EEst <= 1.f) {
          ss.accept = 1;
          ss.naccept += 1;
          ss.qold = fmaxf(EEst, 1e-4f);
          const float dtnew = __fdiv_rn(dt, q);
          float ttmp = __fadd_rn(t, dt);
          if (fabsf(__fsub_rn(ttmp, tstop)) < __fmul_rn(100.f, eps32(fmaxf(fabsf(t), fabsf(tstop)))))
            ttmp = tstop;
          ss.tnew = ttmp;
          // dt for the next step is set after the outputs below (they need this step's dt)
          ss.EEst = dtnew;   // stash
        } else {
          ss.accept = 0;
          ss.nreject += 1;
          ss.dt = __fdiv_rn(dt, fminf(inv_qmin, __fdiv_rn(ss.q11, gamma)));
        }
      }
    }
    S.sync();
    if (ss.done) break;
    if (!ss.accept) continue;

    // ---- accepted step
    const float tnew = ss.tnew;
    if constexpr (T::adj) if (!P.mu_in_norm) S.quad_accumulate(dt, g1);   // (its slab reads are ordered by the S.sync() above)
    if constexpr (!T::adj) {
      // saveat: 4th-order interpolant with the integrator's dt (OrdinaryDiffEq savevalues!)
      int si = ss.si, oi = ss.out_idx;
      while (si < P.nsave && tdir * P.saveat[si] <= tdir * tnew) {
        const float tsv = P.saveat[si];
        float w[7];
        const bool exact = (tsv == tnew);
        if (!exact) Solver<FL, TC>::interp_weights(__fdiv_rn(__fsub_rn(tsv, t), dt), w);
        float* o = P.out_u + (long long)oi * B * NV;
        S.for_my_tiles([&](int tile) {
          const int base_row = tile * TB;
          for (int e = tid; e < TB * NV; e += NTHREADS) {
            const int r = e / NV;
            if (base_row + r >= B) continue;
            const long long g = (long long)base_row * NV + e;
            float v;
            if (exact) v = unew[g];
            else {
              float acc = w[0] * kp[0][g];
#pragma unroll
              for (int j = 1; j < 7; ++j) acc = fmaf(w[j], kp[j][g], acc);
              v = fmaf(dt, acc, u[g]);
            }
            o[g] = v;
          }
        });
        if (cta == 0 && tid == 0) P.out_t[oi] = tsv;
        ++si;
        ++oi;
      }
      S.sync();
      if (tid == 0) { ss.si = si; ss.out_idx = oi; }
      if (P.dense && cta == 0 && tid == 0) P.step_t[step + 1] = tnew;
    }
    if (tid == 0) {
      ss.t = tnew;
      ss.dt = tdir * fminf(dtmax, fabsf(ss.EEst));
      ss.step = step + 1;
      ss.g1 = g7;
      ss.mu_cur = T::adj ? 1 - ss.mu_cur : 0;
    }
    S.sync();

    if constexpr (T::adj) {
      if (tnew == tstop) {
        // discrete-loss jump (ReverseLossCallback): lambda += dL[ti]; Backsolve also resets y
        const int ti = ss.ti;
        const bool has = ti >= 0 && P.f_ts[ti] == tnew;
        const bool is_final = (tnew == P.t1);
        if (has) {
          const int NS = P.NS;
          S.for_my_tiles([&](int tile) {
            const int base_row = tile * TB;
            for (int e = tid; e < TB * NV; e += NTHREADS) {
              const int r = e / NV, c = e - r * NV;
              const int b = base_row + r;
              if (b >= B) continue;
              const long long g = (long long)b * NV + c;
              if (c < NS) unew[g] += P.dL[((long long)ti * B + b) * NS + c];
              else unew[g] = P.f_us[((long long)ti * B + b) * NS + (c - NS)];
            }
          });
          S.sync();
          if (tid == 0) ss.ti = ti - 1;
          S.sync();
          if (!is_final) {   // FSAL re-evaluated lazily: only if the integration continues
            if constexpr (T::interp) {
              if (tid == 0) S.locate(tnew, ss.ii[0]);
              S.sync();
            }
            SlabOut so{S.slab(ss.g1), nullptr, 0.f, nullptr, 0.f};
            S.eval_rhs_vec(unew, S.k1ptr_w(ss.step), &ss.ii[0], so);
          }
        }
      }
    }
  }

  // ------------------------------------------------------------------ epilogue
  S.sync();
  if constexpr (T::adj) {
    if (!P.mu_in_norm && !S.aborted()) {
      // quadrature adjoints: the one and only cross-CTA (and cross-rank) reduction of the parameter gradient
      S.barrier_reduce(0.0, 0.0, false);
      float* mo = P.mu_out;
      S.mu_reduce(7, 1.f, -1, 0.f, -1, [&](int j, float tot, float) { mo[j] = tot; return 0.f; });
      S.mu_reduce_done();
    }
  }
  if (P.world > 1 && !S.aborted()) {
    // closing exchange: once it completes every rank has consumed all earlier mailbox slots, so the
    // next solve may overwrite them
    if (tid == 0) { ss.sums[0] = 0.0; ss.sums[1] = 0.0; }
    S.sync();
    S.cross_rank_exchange();
  }
  {
    const float* uf = S.slotv(ss.step, 0);
    const bool write_end = !T::adj && P.save_end && P.t0 != P.t1;
    float* oe = write_end ? P.out_u + (long long)ss.out_idx * B * NV : nullptr;
    S.for_my_tiles([&](int tile) {
      const int base_row = tile * TB;
      for (int e = tid; e < TB * NV; e += NTHREADS) {
        const int r = e / NV;
        if (base_row + r >= B) continue;
        const long long g = (long long)base_row * NV + e;
        const float v = uf[g];
        P.final_state[g] = v;
        if (oe) oe[g] = v;
      }
    });
    if constexpr (T::adj) {
      if (P.mu_in_norm) {
        const float* mu = P.mu + (long long)ss.mu_cur * M.ppad;
        const int pp = M.ppad;
        int chunk = (pp + P.nctas - 1) / P.nctas;
        chunk = (chunk + 3) & ~3;
        const int lo = cta * chunk, hi = min(pp, lo + chunk);
        for (int j = lo + tid; j < hi; j += NTHREADS) P.mu_out[j] = mu[j];
      }
    }
    if (cta == 0 && tid == 0) {
      int n_out = ss.out_idx;
      if (write_end) { P.out_t[n_out] = ss.t; n_out += 1; }
      if (ss.retcode != 0) atomicCAS(&ctrl->retcode, 0, ss.retcode);
      ctrl->naccept = ss.naccept;
      ctrl->nreject = ss.nreject;
      ctrl->nf = ss.nf;
      ctrl->nsteps = ss.step;
      ctrl->t_final = ss.t;
      ctrl->dt_last = ss.dt;
      ctrl->tie_dist = ss.tie;
      ctrl->n_out = n_out;
    }
  }
#ifdef B200_TC_PROF
  if (cta == 0 && tid == 0) printf("TCPROF loop adj=%d steps=%lld stages=%lld errpart=%lld bar1=%lld mu=%lld bar2=%lld total=%lld\n", (int)T::adj, qf[6], qf[0], qf[1], qf[2], qf[3], qf[4], clock64() - qf_start);
#endif
  if constexpr (TC) S.tc_fini();
}

}  // namespace b200
