#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the oracle (oracle/*.py).

The reference (DiffEqFlux.jl + OrdinaryDiffEq + SciMLSensitivity + Lux) cannot run in the build image (no
Julia, no network), so these vectors are NOT outputs of the reference itself: they freeze the oracle -- the
restatement of the reference's algorithm that the CUDA path is held to -- so that (a) a later edit of the
oracle cannot silently move the parity target (tests/test_golden.py, CPU) and (b) the GPU box, where neither
/root/reference nor a second implementation exists, still compares the CUDA path against committed numbers
(tests/test_golden.py, -m gpu).  "Parity unpinned"  therefore still applies.

Cases mirror the reference's own tests for the path:
  spiral      test/BasicNeuralDE/neural_de_tests.jl + docs spiral (2 -> 50 -> 2 tanh, 30 saveat, Interpolating)
  batched     neural_de_tests.jl batched input with BacksolveAdjoint, widths the tcgen05 engine takes
  ffjord      test/CNF/cnf_tests.jl monte-carlo branch (2 -> 16 -> 2 tanh), fixed noise seed
  mshoot      test/MultipleShooting group_ranges + grouped solves
  headline    BASELINE.json configs[1] at FULL size on bench.py's exact inputs (B = 16384, 64 -> 256 -> 256 -> 64,
              Tsit5 1e-6, BacksolveAdjoint): step statistics, per-step (t, dt, EEst) traces, dp and strided slices
              of u(t1) / du0, for the three accumulation models of the oracle (see oracle/mlp.py) -- ~3 minutes;
              only regenerated with `make_golden.py headline`

Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import mshoot as omshoot  # noqa: E402
from oracle import node as onode  # noqa: E402
from oracle.mlp import ACT_ID, ACT_TANH, MLP  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def stats_vec(st):
    return np.array([st.naccept, st.nreject, st.nf, st.retcode], np.int64)


def spiral():
    rng = np.random.default_rng(11)
    mlp = MLP([2, 50, 2], [ACT_TANH, ACT_ID])
    p = mlp.init(rng, scale=1.0)
    u0 = np.array([[2.0, 0.0]], np.float32)
    ts = np.linspace(0.0, 1.5, 30).astype(np.float32)
    prob = onode.NodeProblem(mlp)
    fw = onode.forward(prob, p, u0, 0.0, 1.5, saveat=ts, dense=True, abstol=1e-6, reltol=1e-3)
    data = np.stack([2.0 * np.cos(ts), 2.0 * np.sin(ts)], -1).astype(np.float32)[:, None, :] * np.exp(-0.1 * ts)[:, None, None].astype(np.float32)
    dL = (2.0 * (fw.us - data)).astype(np.float32)          # d/du of sum((u - data)^2)
    adj = onode.adjoint(prob, p, fw, dL, sensealg=0, abstol=1e-6, reltol=1e-3)
    np.savez(os.path.join(HERE, "spiral.npz"), p=p, u0=u0, ts=ts, us=fw.us, dL=dL, du0=adj.du0, dp=adj.dp,
             fwd_stats=stats_vec(fw.stats), adj_stats=stats_vec(adj.stats))


def batched():
    rng = np.random.default_rng(3)
    dims = [32, 64, 64, 32]
    mlp = MLP(dims, [ACT_TANH, ACT_TANH, ACT_ID])
    p = mlp.init(rng, scale=2.5)
    u0 = rng.standard_normal((192, 32)).astype(np.float32)
    prob = onode.NodeProblem(mlp)
    fw = onode.forward(prob, p, u0, 0.0, 1.0, save_everystep=False, save_start=False, abstol=1e-6, reltol=1e-6)
    dL = np.ones_like(fw.us)
    adj = onode.adjoint(prob, p, fw, dL, sensealg=1, abstol=1e-6, reltol=1e-6)
    np.savez(os.path.join(HERE, "batched.npz"), dims=np.array(dims), p=p, u0=u0, us=fw.us, du0=adj.du0, dp=adj.dp,
             fwd_stats=stats_vec(fw.stats), adj_stats=stats_vec(adj.stats))


def ffjord():
    rng = np.random.default_rng(6)
    mlp = MLP([2, 16, 16, 2], [ACT_TANH, ACT_TANH, ACT_ID])
    p = mlp.init(rng, scale=0.5)
    B = 128
    x = rng.standard_normal((B, 2)).astype(np.float32)
    prob = onode.FfjordProblem(mlp, seed=3)
    u0 = np.concatenate([x, np.zeros((B, 1), np.float32)], 1)
    fw = onode.forward(prob, p, u0, 0.0, 0.5, save_everystep=False, save_start=False, dense=True, abstol=1e-3, reltol=1e-3)
    logpx = onode.ffjord_logpx(fw.us[0], 2)
    g = np.full(B, -1.0 / B, np.float32)                       # loss = -mean(logpx)
    adj = onode.adjoint(prob, p, fw, onode.ffjord_logpx_grad(fw.us[0], 2, g)[None], abstol=1e-3, reltol=1e-3)
    np.savez(os.path.join(HERE, "ffjord.npz"), p=p, x=x, logpx=logpx, g=g, dx=adj.du0[:, :2], dp=adj.dp, seed=np.int64(3),
             fwd_stats=stats_vec(fw.stats), adj_stats=stats_vec(adj.stats))


def mshoot():
    rng = np.random.default_rng(7)
    mlp = MLP([4, 16, 4], [ACT_TANH, ACT_ID])
    teacher = mlp.init(rng, scale=2.0)
    prob = onode.NodeProblem(mlp)
    ts = np.linspace(0, 2, 9, dtype=np.float32)
    u0 = rng.standard_normal((24, 4)).astype(np.float32)
    tol = dict(abstol=1e-5, reltol=1e-5)
    data = onode.forward(prob, teacher, u0, 0.0, 2.0, saveat=ts, **tol).us          # (T, B, D)
    p = (teacher + 0.05 * rng.standard_normal(teacher.size)).astype(np.float32)
    loss, preds, dp, _ = omshoot.multiple_shoot(prob, p, data, ts, 3, **tol)
    ranges = omshoot.group_ranges(9, 3)
    np.savez(os.path.join(HERE, "mshoot.npz"), p=p, u0=u0, ts=ts, data=data, loss=np.float64(loss), dp=dp.astype(np.float32),
             pred_last=preds[-1], ranges=np.array(ranges, np.int64), group_size=np.int64(3),
             ranges_21_5=np.array(omshoot.group_ranges(21, 5), np.int64))


def headline():
    import zlib
    import bench
    p, u0 = bench.synth(bench.BATCH)
    out = dict(p_crc=np.int64(zlib.crc32(p.tobytes())), u0_crc=np.int64(zlib.crc32(u0.tobytes())), batch=np.int64(bench.BATCH))
    kw = dict(abstol=bench.TOL, reltol=bench.TOL)
    # model name -> (wgrad accumulation, layer-GEMM accumulation); "defer" is the parity target of the CUDA path
    models = {"blas": ("blas", "blas"), "defer": (("tile_defer", 128), "blas"), "exactgemm": (("tile", 128), "fp64")}
    for name, (wacc, gacc) in models.items():
        mlp = MLP(bench.DIMS, [ACT_TANH, ACT_TANH, ACT_ID], wgrad_accum=wacc, gemm_accum=gacc)
        prob = onode.NodeProblem(mlp)
        fw = onode.forward(prob, p, u0, bench.TSPAN[0], bench.TSPAN[1], save_everystep=False, save_start=False, record_trace=True, **kw)
        adj = onode.adjoint(prob, p, fw, np.ones_like(fw.us), sensealg=onode.SENSE_BACKSOLVE, record_trace=True, **kw)
        print("headline", name, stats_vec(fw.stats), stats_vec(adj.stats), flush=True)
        out[f"{name}_fwd_stats"], out[f"{name}_adj_stats"] = stats_vec(fw.stats), stats_vec(adj.stats)
        out[f"{name}_fwd_trace"] = np.array(fw.stats.trace, np.float32)
        out[f"{name}_adj_trace"] = np.array(adj.stats.trace, np.float32)
        if name == "defer":
            out.update(dp=adj.dp, du0_s=adj.du0[::64].copy(), uT_s=fw.us[-1][::64].copy())
    np.savez_compressed(os.path.join(HERE, "headline.npz"), **out)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "headline":
        headline()
        sys.exit(0)
    spiral()
    batched()
    ffjord()
    mshoot()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
