"""ORACLE (test infrastructure only) -- fp32 CPU restatement of the Tsit5 integrator loop.

PARITY UNPINNED: the arithmetic on DiffEqFlux's hot path lives in OrdinaryDiffEq.jl
(compat "7", /root/reference/Project.toml:57), which is not vendored under /root/reference and
cannot be executed here (no Julia toolchain).  The reference's own tests pin no numbers on this
path (/root/reference/test/BasicNeuralDE/neural_de_tests.jl:46-48 assert only `!iszero(grad)`).
This file restates the published algorithm; it is reached from the reference at
/root/reference/src/neural_de.jl:55-58 (`solve(prob, Tsit5(); ...)`).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product path (diffeqflux.jl_b200) never does.

Restated semantics:
  * Tsit5 tableau, 7 stages, FSAL, 4th-order embedded error, 4th-order dense interpolant.
  * shared-dt batch: the error norm is an RMS over ALL entries of the state (state x batch).
  * PI controller (beta1=7/50, beta2=2/25, gamma=9/10, qmin=1/5, qmax=10, qsteady=[1,1],
    qoldinit=1e-4) with the Float32 `fastpow`.
  * Hairer initial step (2 RHS) after FSAL initialisation (1 RHS).
  * tstops with callbacks (state jump, FSAL re-evaluated: +1 RHS) for the adjoint's
    discrete-loss jumps.
  * All scalar time arithmetic in float32.
The state is a list of float32 ndarrays so the adjoint's [lambda; mu; y] fits unchanged.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

f32 = np.float32

# ----------------------------------------------------------------------------- tableau
C = [f32(x) for x in (0.161, 0.327, 0.9, 0.9800255409045097, 1.0, 1.0)]
A = {
    (2, 1): 0.161,
    (3, 1): -0.008480655492356989, (3, 2): 0.335480655492357,
    (4, 1): 2.8971530571054935, (4, 2): -6.359448489975075, (4, 3): 4.3622954328695815,
    (5, 1): 5.325864828439257, (5, 2): -11.748883564062828, (5, 3): 7.4955393428898365,
    (5, 4): -0.09249506636175525,
    (6, 1): 5.86145544294642, (6, 2): -12.92096931784711, (6, 3): 8.159367898576159,
    (6, 4): -0.071584973281401, (6, 5): -0.028269050394068383,
    (7, 1): 0.09646076681806523, (7, 2): 0.01, (7, 3): 0.4798896504144996,
    (7, 4): 1.379008574103742, (7, 5): -3.290069515436081, (7, 6): 2.324710524099774,
}
A = {k: f32(v) for k, v in A.items()}
BTILDE = [f32(x) for x in (
    -0.00178001105222577714, -0.0008164344596567469, 0.007880878010261995,
    -0.1447110071732629, 0.5823571654525552, -0.45808210592918697, 0.015151515151515152)]
# dense output polynomial coefficients r[i] = (ri1, ri2, ri3, ri4); only b1 has a linear term
R = [
    (1.0, -2.763706197274826, 2.9132554618219126, -1.0530884977290216),
    (0.0, 0.13169999999999998, -0.2234, 0.1017),
    (0.0, 3.9302962368947516, -5.941033872131505, 2.490627285651253),
    (0.0, -12.411077166933676, 30.33818863028232, -16.548102889244902),
    (0.0, 37.50931341651104, -88.1789048947664, 47.37952196281928),
    (0.0, -27.896526289197286, 65.09189467479366, -34.87065786149661),
    (0.0, 1.5, -4.0, 2.5),
]
R = [tuple(f32(x) for x in r) for r in R]

BETA1 = f32(7.0 / 50.0)
BETA2 = f32(2.0 / 25.0)
GAMMA = f32(9.0 / 10.0)
QMIN = f32(1.0 / 5.0)
QMAX = f32(10.0)
QOLDINIT = f32(1e-4)

RET_SUCCESS, RET_MAXITERS, RET_DTMIN, RET_UNSTABLE = 0, 1, 2, 3


# ----------------------------------------------------------------------------- fastpow
def fastlog2(x) -> np.float32:
    """Goldberg rational log2 (DiffEqBase `fastlog2`, Float32). Pure fp32 mul/add/div."""
    x = f32(x)
    ux = int(np.array([x], dtype=np.float32).view(np.uint32)[0])
    e = (ux & 0x7F800000) >> 23
    a, b, c = f32(0.338953), f32(2.198599), f32(1.523692)
    if ux & 0x00400000:
        sig = np.array([(ux & 0x007FFFFF) | 0x3F000000], dtype=np.uint32).view(np.float32)[0]
        fexp = f32(e) - f32(126.0)
    else:
        sig = np.array([(ux & 0x007FFFFF) | 0x3F800000], dtype=np.uint32).view(np.float32)[0]
        fexp = f32(e) - f32(127.0)
    s = f32(sig - f32(1.0))
    num = f32(s * f32(f32(a * s) + b))
    return f32(fexp + f32(num / f32(s + c)))


def fastpow(x, y) -> np.float32:
    x = f32(x)
    if x == 0:
        return f32(0.0)
    return f32(np.exp2(f32(f32(y) * fastlog2(x))))


def eps32(x) -> np.float32:
    return f32(np.spacing(f32(abs(x))))


# ----------------------------------------------------------------------------- state helpers
def _axpy_stage(u, ks, coefs, dt):
    """u + dt*(sum_j coefs[j]*ks[j]) for every block of the state."""
    out = []
    for b in range(len(u)):
        if _is_deferred(u[b], ks[0][b]):
            out.append(u[b])            # never read by the RHS; updated at the end of the step
            continue
        acc = coefs[0] * ks[0][b]
        for j in range(1, len(coefs)):
            acc = acc + coefs[j] * ks[j][b]
        out.append((u[b] + dt * acc).astype(np.float32, copy=False))
    return out


def _fma32(a, b, c):
    """float32 fused multiply-add (the product of two float32 is exact in float64)."""
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(np.float32)


def _is_deferred(u_block, k_block) -> bool:
    """A "deferred-sum" block: its derivative arrives as per-tile partial sums [ntiles, ...] (the parameter
    gradient of the adjoint with MLP(wgrad_accum=("tile_defer", n))).  The RHS never reads such a block, so
    stage values are not formed; at the end of a step the stage combinations are taken per tile in float32
    (FMA chain in stage order) and the tiles are summed in float64 -- the accumulation order of the CUDA path
    (csrc/device.cuh: step_local / mu_reduce)."""
    return k_block.ndim == u_block.ndim + 1


def _deferred_total(k_block):
    return np.sum(k_block, axis=0, dtype=np.float64).astype(np.float32)


def _deferred_combo(ks, b, coefs):
    """sum over tiles (float64) of the per-tile float32 FMA chain  coefs[0]*k1 + coefs[1]*k2 + ..."""
    acc = (coefs[0] * ks[0][b]).astype(np.float32)
    for j in range(1, len(coefs)):
        acc = _fma32(coefs[j], ks[j][b], acc)
    return np.sum(acc, axis=0, dtype=np.float64).astype(np.float32)


def _sumsq(blocks) -> float:
    """float64 sum of fp32-squared entries (the CUDA path combines per-tile fp32 partials in
    fp64; a pure-fp32 reassociated sum as Julia's differs only at the 1e-7 level)."""
    s = 0.0
    for x in blocks:
        s += float(np.sum(np.square(x, dtype=np.float32), dtype=np.float64))
    return s


def _nelem(blocks) -> int:
    return int(sum(x.size for x in blocks))


def rms_norm(blocks) -> np.float32:
    return f32(math.sqrt(_sumsq(blocks) / _nelem(blocks)))


@dataclass
class Stats:
    nf: int = 0
    naccept: int = 0
    nreject: int = 0
    retcode: int = RET_SUCCESS
    tie_dist: float = float("inf")      # min |EEst-1| over the run (H2 diagnostic)
    trace: list = field(default_factory=list)   # (t, dt, EEst, accepted)


@dataclass
class DenseStep:
    t: np.float32          # step start
    dt: np.float32         # integrator dt of the step
    t_end: np.float32      # step end after tstop snapping
    u: list                # state at step start
    k: list                # 7 stage derivatives (each a list of blocks)


@dataclass
class Solution:
    ts: list
    us: list
    stats: Stats
    dense: list
    t_final: np.float32
    u_final: list


def interp_weights(theta):
    """b_i(Theta) of the Tsit5 4th-order interpolant."""
    th = f32(theta)
    out = []
    for i, (r1, r2, r3, r4) in enumerate(R):
        if i == 0:
            out.append(f32(th * f32(r1 + f32(th * f32(r2 + f32(th * f32(r3 + f32(th * r4))))))))
        else:
            out.append(f32(f32(th * th) * f32(r2 + f32(th * f32(r3 + f32(th * r4))))))
    return out


def interpolate(u, ks, dt, theta):
    w = interp_weights(theta)
    return _axpy_stage(u, ks, w, f32(dt))


def dense_eval(dense, t, block=0):
    """sol(t) from stored dense steps (OrdinaryDiffEq `ode_interpolation`): the interval is the
    one whose RIGHT end is the first knot >= t; dt is the difference of the stored knots."""
    t = f32(t)
    n = len(dense)
    ends = [d.t_end for d in dense]
    fwd = dense[0].t_end >= dense[0].t
    lo, hi = 0, n - 1
    while lo < hi:                      # first interval whose end is >= t (direction aware)
        mid = (lo + hi) // 2
        if (ends[mid] >= t) if fwd else (ends[mid] <= t):
            hi = mid
        else:
            lo = mid + 1
    d = dense[lo]
    dts = f32(d.t_end - d.t)
    theta = f32(f32(t - d.t) / dts)
    w = interp_weights(theta)
    acc = w[0] * d.k[0][block]
    for j in range(1, 7):
        acc = acc + w[j] * d.k[j][block]
    return (d.u[block] + dts * acc).astype(np.float32, copy=False)


# ----------------------------------------------------------------------------- initial dt
def initial_dt(rhs, u0, t0, tdir, dtmax, abstol, reltol, stats: Stats) -> np.float32:
    """Hairer's starting step (OrdinaryDiffEq `ode_determine_initdt`, out-of-place, order 5)."""
    abstol, reltol = f32(abstol), f32(reltol)
    sk = [(abstol + np.abs(x) * reltol).astype(np.float32) for x in u0]
    d0 = rms_norm([x / s for x, s in zip(u0, sk)])
    f0 = rhs(u0, t0)
    stats.nf += 1
    f0 = [_deferred_total(f) if _is_deferred(x, f) else f for x, f in zip(u0, f0)]
    d1 = rms_norm([x / s for x, s in zip(f0, sk)])
    smalldt = f32(1e-6)
    if d0 < f32(1e-5) or d1 < f32(1e-5):
        dt0 = smalldt
    else:
        dt0 = f32(f32(d0 / d1) / f32(100.0))
    dt0 = min(dt0, f32(abs(dtmax)))
    if dt0 < f32(10.0) * f32(np.finfo(np.float32).eps):
        return f32(tdir) * smalldt
    dt0_t = f32(tdir) * dt0
    u1 = [(x + dt0_t * f).astype(np.float32) for x, f in zip(u0, f0)]
    f1 = rhs(u1, f32(t0 + dt0_t))
    stats.nf += 1
    f1 = [_deferred_total(f) if _is_deferred(x, f) else f for x, f in zip(u0, f1)]
    d2 = f32(rms_norm([(a - b) / s for a, b, s in zip(f1, f0, sk)]) / dt0)
    mx = max(d1, d2)
    if mx <= f32(1e-15):
        dt1 = float(max(smalldt, f32(dt0 * f32(1e-3))))
    else:
        x32 = f32(-(f32(2.0) + f32(np.log10(mx))) / f32(5.0))
        dt1 = 10.0 ** float(x32)
    dt = f32(min(100.0 * float(dt0), dt1, float(abs(dtmax))))
    return f32(tdir) * dt


# ----------------------------------------------------------------------------- main loop
def solve(rhs, u0, t0, t1, *, abstol=1e-6, reltol=1e-3, saveat=(), save_start=True,
          save_end=True, save_everystep=False, dense=False, tstops=(), on_tstop=None,
          init_jump=None, dt0=None, maxiters=100000, record_trace=False) -> Solution:
    """Shared-dt adaptive Tsit5 on a list-of-blocks state.

    rhs(state, t) -> list of blocks.  `saveat` are INTERIOR output times (strictly inside the
    span, in integration order); the ends are governed by save_start/save_end.  `tstops` are
    interior forced step ends; on reaching one, on_tstop(state, t) may return a modified state
    (u_modified => FSAL re-evaluated, nf += 1).  init_jump(state) is applied before FSAL init
    (PresetTimeCallback at t0 of the solve).
    """
    t0, t1 = f32(t0), f32(t1)
    abstol, reltol = f32(abstol), f32(reltol)
    tdir = f32(1.0) if t1 >= t0 else f32(-1.0)
    dtmax = f32(abs(t1 - t0))
    stats = Stats()
    u = [np.asarray(x, dtype=np.float32) for x in u0]
    if init_jump is not None:
        u = init_jump(u)
    ts_out, us_out, dense_out = [], [], []
    if save_start:
        ts_out.append(t0)
        us_out.append([x.copy() for x in u])
    saveat = [f32(s) for s in saveat]
    stops = [f32(s) for s in tstops] + [t1]
    si = 0          # next saveat index
    ti = 0          # next tstop index
    n_elem = _nelem(u)

    k1 = rhs(u, t0)
    stats.nf += 1
    if dt0 is None or dt0 == 0:
        dt = initial_dt(rhs, u, t0, tdir, dtmax, abstol, reltol, stats)
    else:
        dt = f32(tdir) * f32(abs(dt0))
    t = t0
    qold = QOLDINIT
    q11 = f32(1.0)
    it = 0
    retcode = RET_SUCCESS
    if t0 == t1:
        stops = []
    while ti < len(stops):
        it += 1
        if it > maxiters:
            retcode = RET_MAXITERS
            break
        tstop = stops[ti]
        # clamp to dtmax and to the next tstop
        dt = tdir * min(abs(dt), dtmax)
        dt = tdir * min(abs(dt), abs(f32(tstop - t)))
        if not np.isfinite(dt) or abs(dt) <= eps32(t):
            retcode = RET_DTMIN if np.isfinite(dt) else RET_UNSTABLE
            break
        ks = [k1]
        for s in range(2, 8):
            us = _axpy_stage(u, ks, [A[(s, j)] for j in range(1, s)], dt)
            ks.append(rhs(us, f32(t + C[s - 2] * dt)))
            if s == 7:
                unew = us
        stats.nf += 6
        # embedded error estimate
        num = 0.0
        for b in range(len(u)):
            if _is_deferred(u[b], ks[0][b]):
                tot_a = _deferred_combo(ks, b, [A[(7, j)] for j in range(1, 7)])
                tot_e = _deferred_combo(ks, b, BTILDE)
                unew[b] = _fma32(dt, tot_a, u[b])
                ut = (dt * tot_e).astype(np.float32)
                sk = _fma32(np.maximum(np.abs(u[b]), np.abs(unew[b])), reltol, abstol)
                r = (ut / sk).astype(np.float32)
                num += float(np.sum(np.square(r, dtype=np.float32), dtype=np.float64))
                continue
            acc = BTILDE[0] * ks[0][b]
            for j in range(1, 7):
                acc = acc + BTILDE[j] * ks[j][b]
            ut = dt * acc
            sk = abstol + np.maximum(np.abs(u[b]), np.abs(unew[b])) * reltol
            r = (ut / sk).astype(np.float32)
            num += float(np.sum(np.square(r, dtype=np.float32), dtype=np.float64))
        EEst = f32(math.sqrt(num / n_elem))
        if not np.isfinite(EEst):
            retcode = RET_UNSTABLE
            break
        # PI controller
        if EEst == 0:
            q = f32(1.0) / QMAX
        else:
            q11 = fastpow(EEst, BETA1)
            q = f32(q11 / fastpow(qold, BETA2))
            q = max(f32(1.0) / QMAX, min(f32(1.0) / QMIN, f32(q / GAMMA)))
        accept = bool(EEst <= f32(1.0))
        stats.tie_dist = min(stats.tie_dist, abs(float(EEst) - 1.0))
        if record_trace:
            stats.trace.append((float(t), float(dt), float(EEst), accept))
        if not accept:
            stats.nreject += 1
            dt = f32(dt / min(f32(1.0) / QMIN, f32(q11 / GAMMA)))
            continue
        stats.naccept += 1
        # qsteady_min = qsteady_max = 1: the "snap q to 1" window is the single point q == 1
        qold = max(EEst, QOLDINIT)
        dtnew = f32(dt / q)
        ttmp = f32(t + dt)
        if abs(f32(ttmp - tstop)) < f32(100.0) * eps32(max(abs(t), abs(tstop))):
            ttmp = tstop
        tprev = t
        # saveat: interpolate with the integrator's dt (OrdinaryDiffEq `savevalues!`)
        while si < len(saveat) and tdir * saveat[si] <= tdir * ttmp:
            ts_ = saveat[si]
            if ts_ == ttmp:
                val = [x.copy() for x in unew]
            else:
                theta = f32(f32(ts_ - tprev) / dt)
                val = interpolate(u, ks, dt, theta)
            ts_out.append(ts_)
            us_out.append(val)
            si += 1
        if dense:
            dense_out.append(DenseStep(t=tprev, dt=dt, t_end=ttmp, u=u, k=ks))
        u = unew
        t = ttmp
        k1 = ks[6]
        if save_everystep:
            ts_out.append(t)
            us_out.append([x.copy() for x in u])
        dt = tdir * min(dtmax, abs(dtnew))
        if t == tstop:
            ti += 1
            if on_tstop is not None:
                mod = on_tstop(u, t)
                if mod is not None:
                    u = mod
                    if ti < len(stops):      # FSAL is re-evaluated lazily: not after the last stop
                        k1 = rhs(u, t)
                        stats.nf += 1
    if retcode == RET_SUCCESS and any(not np.all(np.isfinite(x)) for x in u):
        retcode = RET_UNSTABLE
    if save_end and not save_everystep and t0 != t1:
        ts_out.append(t)
        us_out.append([x.copy() for x in u])
    stats.retcode = retcode
    return Solution(ts=ts_out, us=us_out, stats=stats, dense=dense_out, t_final=t, u_final=u)
