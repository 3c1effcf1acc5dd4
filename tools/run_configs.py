#!/usr/bin/env python
"""Run every BASELINE.json config once on cuda:0 and print one JSON line per config (RHS evals/s, step
counts, kernel times, engine).  cfg2 is the bench.py workload; the others are reported for BASELINE.md."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import diffeqflux_b200 as dfx  # noqa: E402


def mlp_params(dims, seed, scale):
    rng = np.random.default_rng(seed)
    p = []
    for i, o in zip(dims[:-1], dims[1:]):
        p.append((rng.standard_normal(i * o) * scale / np.sqrt(i)).astype(np.float32))
        p.append((rng.standard_normal(o) * 0.1 * scale).astype(np.float32))
    return np.concatenate(p), rng


def chain_of(dims):
    return dfx.Chain(*[dfx.Dense(i, o, dfx.tanh if k < len(dims) - 2 else dfx.identity)
                       for k, (i, o) in enumerate(zip(dims[:-1], dims[1:]))])


def flops(dims):
    return 2 * sum(i * o for i, o in zip(dims[:-1], dims[1:]))


def node_cfg(name, dims, B, tspan, kw, scale, seed, augment=0, reps=2):
    import torch
    p, rng = mlp_params(dims, seed, scale)
    u0 = rng.standard_normal((B, dims[0] - augment)).astype(np.float32)
    node = dfx.NeuralODE(chain_of(dims), tspan, dfx.Tsit5(), **kw)
    layer = dfx.AugmentedNDELayer(node, augment) if augment else node
    pt, ut = torch.from_numpy(p).cuda(), torch.from_numpy(u0).cuda()
    best = None
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        sol, back = layer.pullback(ut, pt)
        du0, dp = back(torch.ones_like(sol.array()))
        torch.cuda.synchronize(); wall = time.perf_counter() - t0
        fs, as_ = sol.stats, layer.last["adjoint_stats"]
        best = wall if best is None else min(best, wall)
    nf = fs.nf + as_.nf
    F = flops(dims)
    kms = fs.kernel_ms + as_.kernel_ms
    return {"config": name, "B": B, "dims": dims, "retcode": int(sol.retcode), "evals_per_s": B * nf / best, "wall_ms": best * 1e3,
            "kernel_ms": kms, "fwd": [fs.naccept, fs.nreject, fs.nf, round(fs.kernel_ms, 3), fs.engine],
            "adj": [as_.naccept, as_.nreject, as_.nf, round(as_.kernel_ms, 3), as_.engine],
            "tflops": B * (fs.nf * F + as_.nf * 3 * F) / (kms * 1e-3) / 1e12}


def main():
    out = []
    ts30 = np.linspace(0, 1.5, 30, dtype=np.float32)
    out.append(node_cfg("cfg1 spiral 2->50->2, 30 saveat, Interpolating", [2, 50, 2], 1, (0.0, 1.5), dict(saveat=ts30), 1.0, 1))
    out.append(node_cfg("cfg2 batched 64->256->256->64, tol 1e-6, Backsolve", [64, 256, 256, 64], 16384, (0.0, 1.0),
                        dict(abstol=1e-6, reltol=1e-6, save_everystep=False, save_start=False, sensealg=dfx.BacksolveAdjoint()), 2.0, 2))
    out.append(node_cfg("cfg2 shape with InterpolatingAdjoint", [64, 256, 256, 64], 16384, (0.0, 1.0),
                        dict(abstol=1e-6, reltol=1e-6, save_everystep=False, save_start=False, max_dense_steps=64), 2.0, 2))
    out.append(node_cfg("cfg5 augmented 128+64 -> 192->512->512->192, tol 1e-3, Interpolating (1 GPU shard of 4096)", [192, 512, 512, 192], 4096, (0.0, 1.0),
                        dict(abstol=1e-3, reltol=1e-3, save_everystep=False, save_start=False, max_dense_steps=64), 1.0, 5, augment=64, reps=1))
    # cfg3 FFJORD
    import torch
    dims = [2, 64, 64, 2]
    p, rng = mlp_params(dims, 3, 0.5)
    B = 65536
    x = rng.standard_normal((B, 2)).astype(np.float32)
    ff = dfx.FFJORD(chain_of(dims), (0.0, 1.0), (2,), dfx.Tsit5(), abstol=1e-3, reltol=1e-3, maxiters=200000)
    st = {"regularize": False, "monte_carlo": True, "seed": 3, "calls": 0}
    t0 = time.perf_counter()
    logpx, back = ff.pullback(torch.from_numpy(x).cuda(), torch.from_numpy(p).cuda(), st)
    dx, dp = back(torch.full((B,), -1.0 / B, device="cuda"))
    torch.cuda.synchronize(); wall = time.perf_counter() - t0
    fs, as_ = ff.last["stats"], ff.last["adjoint_stats"]
    F = flops(dims)
    out.append({"config": "cfg3 FFJORD 2->64->64->2, B=65536, Hutchinson noise per RHS call, tol 1e-3", "B": B, "retcode": [int(ff.last["rc"]), int(ff.last["adjoint_rc"])],
                "evals_per_s": B * (fs.nf + as_.nf) / wall, "wall_ms": wall * 1e3, "kernel_ms": fs.kernel_ms + as_.kernel_ms,
                "fwd": [fs.naccept, fs.nreject, fs.nf, round(fs.kernel_ms, 3)], "adj": [as_.naccept, as_.nreject, as_.nf, round(as_.kernel_ms, 3)],
                "tflops": B * (fs.nf * 2 * F + as_.nf * 6 * F) / ((fs.kernel_ms + as_.kernel_ms) * 1e-3) / 1e12})
    # cfg4 multiple shooting
    dims = [32, 128, 128, 32]
    p, rng = mlp_params(dims, 4, 1.5)
    B, T, gs = 4096, 65, 5
    ts = np.linspace(0, 4, T, dtype=np.float32)
    data = torch.from_numpy(rng.standard_normal((T, B, 32)).astype(np.float32)).cuda()
    pt = torch.from_numpy(p).cuda().requires_grad_(True)
    prob = dfx.ODEProblem(chain_of(dims), None, (0.0, 4.0), None)
    for _ in range(2):                       # first pass sizes the workspaces; report the second
        pt.grad = None
        dfx.multiple_shoot.last = None
        torch.cuda.synchronize(); t0 = time.perf_counter()
        loss, preds = dfx.multiple_shoot(pt, data, ts, prob, lambda u, uh: ((u - uh) ** 2).sum(), dfx.Tsit5(), gs, abstol=1e-5, reltol=1e-5)
        loss.backward()
        torch.cuda.synchronize(); wall = time.perf_counter() - t0
    run = dfx.multiple_shoot.last
    nf = sum(s["nf"] for s in run.stats) + sum(s["nf"] for s in run.adjoint_stats)
    kms = sum(s["kernel_ms"] for s in run.stats) + sum(s["kernel_ms"] for s in run.adjoint_stats)
    out.append({"config": "cfg4 multiple_shoot 32->128->128->32, 16 groups x 4096, group_size 5 (groups run concurrently, 4 at a time)", "B": B * len(run.stats), "groups": len(run.stats),
                "evals_per_s": B * nf / wall, "wall_ms": wall * 1e3, "kernel_ms": kms, "loss": float(loss), "retcode": int(run.retcode)})
    for o in out:
        if o.get("fwd") and o.get("adj"):       # rate over the kernels' own time (wall clock of one pass includes host set-up)
            o["evals_per_s_kernel_time"] = o["B"] * (o["fwd"][2] + o["adj"][2]) / (o["kernel_ms"] * 1e-3)
        print(json.dumps(o))


if __name__ == "__main__":
    main()
