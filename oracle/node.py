"""ORACLE (test infrastructure only) -- NeuralODE / FFJORD forward solve and adjoint gradient.

PARITY UNPINNED (see oracle/tsit5.py header).  Follows:
  * /root/reference/src/neural_de.jl:47-61   NeuralODE call: ODEProblem{false}(dudt, x, tspan, p),
    solve(prob, Tsit5(); sensealg = InterpolatingAdjoint(ZygoteVJP), kwargs...)
  * /root/reference/src/neural_de.jl:372-386 AugmentedNDELayer zero-padding of the state dim
  * /root/reference/src/ffjord.jl:81-136     __ffjord RHS (monte_carlo = true, AutoZygote branch)
  * /root/reference/src/ffjord.jl:140-178    __forward_ffjord (augment, end state, logpx)
  * SciMLSensitivity (compat 7.112, /root/reference/Project.toml:70) Interpolating/Backsolve
    adjoint restated per SURVEY.md section 3.2: backward Tsit5 solve of [lambda; mu; (y)] with
    discrete-loss jumps at the saved times, mu inside the error norm.

numpy view: states are (B, NS) C-order == Julia (NS, B) column-major.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import tsit5
from .mlp import MLP
from .philox import normal

f32 = np.float32
SENSE_INTERPOLATING, SENSE_BACKSOLVE, SENSE_QUADRATURE, SENSE_GAUSS = 0, 1, 2, 3


# ----------------------------------------------------------------------------- problems
class NodeProblem:
    """du/dt = model(u, p)  (/root/reference/src/neural_de.jl:50)."""

    def __init__(self, mlp: MLP):
        self.mlp = mlp
        self.ns = mlp.dims[0]
        assert mlp.dims[-1] == mlp.dims[0]

    def f(self, p, y, call, stream):
        out, cache = self.mlp.forward(p, y)
        return out, cache

    def vjp(self, p, y, lam, call, stream):
        """returns (f(y), lam^T df/dy, lam^T df/dp)."""
        out, cache = self.mlp.forward(p, y)
        ybar, pbar = self.mlp.vjp(cache, lam)
        return out, ybar, pbar


class FfjordProblem:
    """state rows [z (D); delta_logp (1)], RHS = [m(z); -e^T J e] with fresh e per RHS call
    (/root/reference/src/ffjord.jl:85-92,105-115,134).
    regularize = true (/root/reference/src/ffjord.jl:131-133,149-153): two more rows, d/dt = [sum(abs2, m(z)); ||e^T J||_2]."""

    def __init__(self, mlp: MLP, seed: int, b_offset: int = 0, monte_carlo: bool = True, regularize: bool = False,
                 noise_per_call: bool = True):
        self.mlp = mlp
        self.D = mlp.dims[0]
        self.regularize = regularize
        self.ns = self.D + (3 if regularize else 1)
        self.seed = seed
        self.b_offset = b_offset
        # False: one noise draw per solve direction instead of one per RHS call (b200_opts.noise_per_call = 0; the
        # reference always redraws, src/ffjord.jl:92) -- a smooth RHS, used by the tight-tolerance parity tests
        self.noise_per_call = noise_per_call
        # monte_carlo = false (/root/reference/src/ffjord.jl:120-123): the trace of the batched Jacobian itself, here as
        # sum_d e_d^T J e_d over the unit vectors (what `__trace_batched(Lux.batched_jacobian(...))` evaluates)
        self.monte_carlo = monte_carlo

    def noise(self, B, call, stream):
        return normal(self.seed, stream, call if self.noise_per_call else 0, B, self.D, self.b_offset)

    def units(self, B):
        for d in range(self.D):
            e = np.zeros((B, self.D), np.float32)
            e[:, d] = 1
            yield e

    def _reg_rows(self, mz, eJ):
        """rows D+1, D+2 of the RHS: sum(abs2, mz) and sqrt(sum(abs2, eJ)) per trajectory (src/ffjord.jl:78-79,132)."""
        r1 = np.sum(mz * mz, axis=1, dtype=self.mlp.dtype)
        r2 = np.sqrt(np.sum(eJ * eJ, axis=1, dtype=self.mlp.dtype)).astype(self.mlp.dtype)
        return [r1[:, None], r2[:, None]]

    def f(self, p, y, call, stream):
        z = np.ascontiguousarray(y[:, :self.D])
        mz, cache = self.mlp.forward(p, z)
        if not self.monte_carlo:
            tr = np.zeros(z.shape[0], self.mlp.dtype)
            for e in self.units(z.shape[0]):
                eJ, _ = self.mlp.vjp(cache, e, want_dp=False)
                tr = (tr + np.sum(eJ * e, axis=1, dtype=self.mlp.dtype)).astype(self.mlp.dtype)
            rows = [mz, -tr[:, None]]
            if self.regularize:                           # e ~ randn, eJ = e^T J (src/ffjord.jl:124-125)
                eJ, _ = self.mlp.vjp(cache, self.noise(z.shape[0], call, stream), want_dp=False)
                rows += self._reg_rows(mz, eJ)
            return np.concatenate(rows, axis=1).astype(self.mlp.dtype), None
        e = self.noise(z.shape[0], call, stream)
        eJ, _ = self.mlp.vjp(cache, e, want_dp=False)
        tr = np.sum(eJ * e, axis=1, dtype=self.mlp.dtype)
        rows = [mz, -tr[:, None]]
        if self.regularize:
            rows += self._reg_rows(mz, eJ)
        return np.concatenate(rows, axis=1).astype(self.mlp.dtype), None

    def vjp(self, p, y, lam, call, stream):
        """(f(y), lam^T df/dy, lam^T df/dp) of  <lz, m> - ll * tr  [+ l1 * sum(m^2) + l2 * ||e^T J||].
        Everything is a sum of second-order reverse passes through the doubled (h, hdot) network, each linear in its seed
        pair (hbar_L, hdbar_L) with tangent direction hdot_0:
          trace (Hutchinson):  direction e,    seeds (lz [+ 2 l1 m(z)], -ll e)
          trace (exact):       direction e_d,  seeds ((lz [+ 2 l1 m(z)]) / D, -ll e_d) for every unit vector
          ||e^T J||:           direction w = l2 * eJ / ||eJ|| (held constant), seeds (0, e):  d/d(z,p) <e, J w>"""
        z = np.ascontiguousarray(y[:, :self.D])
        lz = np.ascontiguousarray(lam[:, :self.D])
        ll = lam[:, self.D:self.D + 1]
        B = z.shape[0]
        zeros_tail = np.zeros((B, self.ns - self.D), np.float32)
        out, _ = self.f(p, y, call, stream)
        if self.regularize:
            l1 = lam[:, self.D + 1:self.D + 2]
            l2 = lam[:, self.D + 2:self.D + 3]
            lz = (lz + np.float32(2) * l1 * out[:, :self.D]).astype(np.float32)
        if not self.monte_carlo:
            zbar = np.zeros_like(z)
            pbar = np.zeros(p.size, np.float32)
            for e in self.units(B):          # linear in the seeds: (lz / D, -ll e_d) per unit vector
                _, _, cache = self.mlp.jvp_forward(p, z, e)
                zb, pb = self.mlp.second_order_vjp(cache, (lz / np.float32(self.D)).astype(np.float32), (-ll * e).astype(np.float32))
                zbar = (zbar + zb).astype(np.float32)
                pbar = (pbar + pb).astype(np.float32)
        else:
            e = self.noise(B, call, stream)
            hL, hdL, cache = self.mlp.jvp_forward(p, z, e)
            zbar, pbar = self.mlp.second_order_vjp(cache, lz, (-ll * e).astype(np.float32))
        if self.regularize:
            e = self.noise(B, call, stream)
            _, cache1 = self.mlp.forward(p, z)
            eJ, _ = self.mlp.vjp(cache1, e, want_dp=False)
            nrm = np.sqrt(np.sum(eJ * eJ, axis=1, dtype=np.float32)).astype(np.float32)[:, None]
            w = np.where(nrm > 0, l2 * eJ / np.where(nrm > 0, nrm, np.float32(1)), np.float32(0)).astype(np.float32)
            _, _, cache = self.mlp.jvp_forward(p, z, w)
            zb, pb = self.mlp.second_order_vjp(cache, np.zeros_like(e), e)
            zbar = (zbar + zb).astype(np.float32)
            pbar = (pbar + pb).astype(np.float32)
        ybar = np.concatenate([zbar, zeros_tail], axis=1).astype(np.float32)
        return out, ybar, pbar


# ----------------------------------------------------------------------------- forward
@dataclass
class Forward:
    ts: np.ndarray            # (T,) output times
    us: np.ndarray            # (T, B, NS)
    stats: tsit5.Stats
    sol: tsit5.Solution
    t0: np.float32
    t1: np.float32


def split_saveat(saveat, t0, t1, save_start=None, save_end=None, save_everystep=None):
    """OrdinaryDiffEq saveat conventions: the interior points go to the integrator's saveat heap,
    the ends are governed by save_start / save_end whose defaults depend on saveat."""
    t0, t1 = f32(t0), f32(t1)
    if saveat is None:
        sv = np.zeros(0, dtype=np.float32)
        scalar = False
    elif np.isscalar(saveat):
        h = f32(abs(saveat))
        n = int(np.floor(float((t1 - t0) / h) + 1e-6))
        sv = (t0 + h * np.arange(0, n + 1, dtype=np.float32)).astype(np.float32)
        scalar = True
    else:
        sv = np.asarray(saveat, dtype=np.float32)
        scalar = False
    if save_everystep is None:
        save_everystep = sv.size == 0
    if save_start is None:
        save_start = bool(save_everystep or sv.size == 0 or scalar or (t0 in sv))
    if save_end is None:
        save_end = bool(save_everystep or sv.size == 0 or scalar or (t1 in sv))
    lo, hi = min(t0, t1), max(t0, t1)
    interior = np.sort(np.array([s for s in sv if lo < s < hi], dtype=np.float32))
    if t1 < t0:
        interior = interior[::-1].copy()
    return interior, bool(save_start), bool(save_end), bool(save_everystep)


def forward(problem, p, u0, t0, t1, *, saveat=None, save_start=None, save_end=None,
            save_everystep=None, abstol=1e-6, reltol=1e-3, dense=False, dt0=None,
            maxiters=100000, record_trace=False, augment=0) -> Forward:
    u0 = np.asarray(u0, dtype=np.float32)
    if u0.ndim == 1:
        u0 = u0[None, :]
    if augment:
        u0 = np.concatenate([u0, np.zeros((u0.shape[0], augment), np.float32)], axis=1)
    interior, ss, se, sev = split_saveat(saveat, t0, t1, save_start, save_end, save_everystep)
    counter = [0]

    def rhs(state, t):
        out, _ = problem.f(p, state[0], counter[0], 0)
        counter[0] += 1
        return [out]

    sol = tsit5.solve(rhs, [u0], t0, t1, abstol=abstol, reltol=reltol, saveat=interior,
                      save_start=ss, save_end=se, save_everystep=sev, dense=dense, dt0=dt0,
                      maxiters=maxiters, record_trace=record_trace)
    ts = np.array(sol.ts, dtype=np.float32)
    us = np.stack([u[0] for u in sol.us]) if sol.us else np.zeros((0,) + u0.shape, np.float32)
    return Forward(ts=ts, us=us, stats=sol.stats, sol=sol, t0=f32(t0), t1=f32(t1))


# ----------------------------------------------------------------------------- adjoint
@dataclass
class Adjoint:
    du0: np.ndarray
    dp: np.ndarray
    stats: tsit5.Stats


def adjoint(problem, p, fwd: Forward, dL_du, *, sensealg=SENSE_INTERPOLATING, abstol=1e-6,
            reltol=1e-3, maxiters=100000, record_trace=False, augment=0, dt0=None, quad_abstol=1e-6,
            quad_reltol=1e-3) -> Adjoint:
    """Backward solve from t1 to t0 of d(lambda)/dt = -(df/dy)^T lambda, d(mu)/dt = -(df/dp)^T lambda
    (and dy/dt = f for Backsolve) with lambda += dL/du(t_i) at every saved time t_i.
    sensealg = SENSE_QUADRATURE / SENSE_GAUSS: see quadrature_adjoint."""
    if sensealg in (SENSE_QUADRATURE, SENSE_GAUSS):
        return quadrature_adjoint(problem, p, fwd, dL_du, abstol=abstol, reltol=reltol, maxiters=maxiters,
                                  record_trace=record_trace, augment=augment, dt0=dt0, quad_abstol=quad_abstol,
                                  quad_reltol=quad_reltol)
    ts = fwd.ts
    dL = np.asarray(dL_du, dtype=np.float32)
    assert dL.shape == fwd.us.shape, (dL.shape, fwd.us.shape)
    t0, t1 = fwd.t0, fwd.t1
    B, NS = fwd.us.shape[1:]
    P = p.size
    counter = [0]
    jumps = {float(t): i for i, t in enumerate(ts)}
    interior = [t for t in ts[::-1] if min(t0, t1) < t < max(t0, t1)]

    if sensealg == SENSE_INTERPOLATING:
        assert fwd.sol.dense, "InterpolatingAdjoint needs the forward dense output"

        def rhs(state, t):
            y = tsit5.dense_eval(fwd.sol.dense, t)
            _, ybar, pbar = problem.vjp(p, y, state[0], counter[0], 1)
            counter[0] += 1
            return [-ybar, -pbar]

        z0 = [np.zeros((B, NS), np.float32), np.zeros(P, np.float32)]
    else:
        def rhs(state, t):
            out, ybar, pbar = problem.vjp(p, state[2], state[0], counter[0], 1)
            counter[0] += 1
            return [-ybar, -pbar, out]

        z0 = [np.zeros((B, NS), np.float32), np.zeros(P, np.float32), fwd.sol.u_final[0].copy()]

    def jump(state, t):
        i = jumps.get(float(t))
        if i is None:
            return None
        new = list(state)
        new[0] = (state[0] + dL[i]).astype(np.float32)
        if sensealg == SENSE_BACKSOLVE:
            new[2] = fwd.us[i].copy()            # checkpointing = true: reset y to the stored state
        return new

    def init_jump(state):
        mod = jump(state, t1)
        return state if mod is None else mod

    sol = tsit5.solve(rhs, z0, t1, t0, abstol=abstol, reltol=reltol, saveat=(), save_start=False,
                      save_end=True, save_everystep=False, tstops=interior, on_tstop=jump,
                      init_jump=init_jump, maxiters=maxiters, record_trace=record_trace, dt0=dt0)
    lam, mu = sol.u_final[0], sol.u_final[1]
    if augment:
        lam = lam[:, :NS - augment]
    return Adjoint(du0=lam, dp=mu, stats=sol.stats)


# ----------------------------------------------------------------------------- QuadratureAdjoint
# Gauss-Kronrod (7, 15) rule on [-1, 1] (QuadGK.jl default order 7): Kronrod nodes x >= 0, Kronrod weights, Gauss weights
_GK_X = np.array([0.991455371120812639206854697526329, 0.949107912342758524526189684047851, 0.864864423359769072789712788640926,
                  0.741531185599394439863864773280788, 0.586087235467691130294144838258730, 0.405845151377397166906606412076961,
                  0.207784955007898467600689403773245, 0.0])
_GK_WK = np.array([0.022935322010529224963732008058970, 0.063092092629978553290700663189204, 0.104790010322250183839876322541518,
                   0.140653259715525918745189590510238, 0.169004726639267902826583426598550, 0.190350578064785409913256402421014,
                   0.204432940075298892414161999234649, 0.209482141084727828012999174891714])
_GK_WG = np.array([0.129484966168869693270611432679082, 0.279705391489276667901467771423780, 0.381830050505118944950369775488975,
                   0.417959183673469387755102040816327])     # Gauss-Legendre 7 weights for nodes _GK_X[1], [3], [5], [7]


def quadgk_vec(f, a, b, atol, rtol, maxevals=10**5):
    """QuadGK.jl `quadgk(f, a, b; atol, rtol)` for a vector-valued integrand [UPSTREAM-RECALL; QuadGK is a dependency of
    SciMLSensitivity, not vendored]: 15-point Kronrod estimate I, error E = ||I - G7|| (Euclidean norm), the segment with
    the largest error is bisected until sum(E) <= max(atol, rtol * ||sum(I)||).  float64 throughout."""
    def rule(lo, hi):
        c, h = 0.5 * (lo + hi), 0.5 * (hi - lo)
        f0 = f(c)
        ik, ig = _GK_WK[7] * f0, _GK_WG[3] * f0
        for j in range(7):
            fs = f(c - h * _GK_X[j]) + f(c + h * _GK_X[j])
            ik = ik + _GK_WK[j] * fs
            if j % 2 == 1:
                ig = ig + _GK_WG[j // 2] * fs
        return ik * h, float(np.linalg.norm((ik - ig) * h))
    segs = [(a, b) + rule(a, b)]
    nev = 15
    while True:
        tot = sum(sg[2] for sg in segs)
        err = sum(sg[3] for sg in segs)
        if err <= max(atol, rtol * float(np.linalg.norm(tot))) or nev >= maxevals:
            return tot, err
        k = max(range(len(segs)), key=lambda i: segs[i][3])
        lo, hi, _, _ = segs.pop(k)
        mid = 0.5 * (lo + hi)
        segs.append((lo, mid) + rule(lo, mid))
        segs.append((mid, hi) + rule(mid, hi))
        nev += 30


def quadrature_adjoint(problem, p, fwd: Forward, dL_du, *, abstol=1e-6, reltol=1e-3, maxiters=100000, record_trace=False,
                       augment=0, dt0=None, quad_abstol=1e-6, quad_reltol=1e-3) -> Adjoint:
    """SciMLSensitivity `QuadratureAdjoint(abstol = 1e-6, reltol = 1e-3)` (re-exported at /root/reference/src/DiffEqFlux.jl:19-20;
    used at /root/reference/test/AdvancedNeuralDE/second_order_ode_tests.jl:59,83), restated [UPSTREAM-RECALL]:
      1. lambda ALONE is integrated from t1 to t0 with the same Tsit5 loop (error norm over lambda only, the solve's
         abstol / reltol), y(t) from the forward dense output, lambda += dL/du(t_i) at every saved time, dense output kept;
      2. dp = sum over the intervals between saved times of quadgk(t -> lambda(t)^T df/dp(y(t), p), t_i, t_{i+1};
         atol = sensealg.abstol, rtol = sensealg.reltol);  du0 = lambda(t0).
    GaussAdjoint is treated alike (its Gauss-Legendre callbacks integrate the same integrand)."""
    assert fwd.sol.dense, "QuadratureAdjoint needs the forward dense output"
    ts = fwd.ts
    dL = np.asarray(dL_du, dtype=np.float32)
    t0, t1 = fwd.t0, fwd.t1
    B, NS = fwd.us.shape[1:]
    counter = [0]
    jumps = {float(t): i for i, t in enumerate(ts)}
    interior = [t for t in ts[::-1] if min(t0, t1) < t < max(t0, t1)]

    def rhs(state, t):
        y = tsit5.dense_eval(fwd.sol.dense, t)
        _, ybar, _ = problem.vjp(p, y, state[0], counter[0], 1)
        counter[0] += 1
        return [-ybar]

    def jump(state, t):
        i = jumps.get(float(t))
        return None if i is None else [(state[0] + dL[i]).astype(np.float32)]

    def init_jump(state):
        mod = jump(state, t1)
        return state if mod is None else mod

    sol = tsit5.solve(rhs, [np.zeros((B, NS), np.float32)], t1, t0, abstol=abstol, reltol=reltol, saveat=(), save_start=False,
                      save_end=True, save_everystep=False, tstops=interior, on_tstop=jump, init_jump=init_jump,
                      maxiters=maxiters, record_trace=record_trace, dt0=dt0, dense=True)
    lam_dense = sol.dense                       # steps run from t1 down to t0; the jumps sit on step boundaries

    def lam_at(t, lower_side):
        """lambda(t) from the adjoint's dense output; at a jump time the value depends on the side: quadgk never
        evaluates the end points, so only interior times occur and the bracketing step is unambiguous."""
        return tsit5.dense_eval(lam_dense, f32(t))

    def integrand(t):
        y = tsit5.dense_eval(fwd.sol.dense, f32(t))
        lam = lam_at(t, False)
        _, _, pbar = problem.vjp(p, y, lam, 0, 1)
        return pbar.astype(np.float64)

    knots = sorted(set([float(t0), float(t1)] + [float(t) for t in ts if min(t0, t1) < t < max(t0, t1)]))
    dp = np.zeros(p.size, np.float64)
    for a, b in zip(knots[:-1], knots[1:]):
        val, _ = quadgk_vec(integrand, a, b, quad_abstol, quad_reltol)
        dp += val
    # d(mu)/dt = -lambda^T df/dp integrated from t1 DOWN to t0: mu(t0) = + int_{t0}^{t1} lambda^T df/dp dt
    lam = sol.u_final[0]
    if augment:
        lam = lam[:, :NS - augment]
    return Adjoint(du0=lam, dp=dp.astype(np.float32), stats=sol.stats)


# ----------------------------------------------------------------------------- FFJORD wrapper
def ffjord_logpx(u_end, D):
    """logpx = logpz - delta_logp with a unit-normal base (/root/reference/src/ffjord.jl:171-176)."""
    z = u_end[:, :D]
    logpz = -np.sum(z * z, axis=1, dtype=np.float32) / f32(2.0) - f32(D / 2.0 * np.log(2.0 * np.pi))
    return (logpz - u_end[:, D]).astype(np.float32)


def ffjord_logpx_grad(u_end, D, g):
    """d/du_end of sum(g * logpx)."""
    out = np.empty_like(u_end)
    out[:, :D] = -u_end[:, :D] * g[:, None]
    out[:, D] = -g
    return out
