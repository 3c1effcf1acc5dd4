#!/usr/bin/env python
"""One launch of every persistent-kernel flavour, in a fixed order, for `ncu -k regex:solve_kernel` (profiles/README_r2.md):
  1, 2  cfg2 (bench workload, B = 16384) on the tcgen05 engine: forward, Backsolve adjoint
  3, 4  cfg2 shape, InterpolatingAdjoint on the tcgen05 engine: forward (dense tape), adjoint
  5, 6  cfg2 shape at B = 4096 on the FP32-FFMA engine: forward, Backsolve adjoint
  7, 8  cfg1 spiral 2 -> 50 -> 2 (B = 1, 30 saveat): forward, Interpolating adjoint
  9, 10 cfg3 FFJORD 2 -> 64 -> 64 -> 2, B = 65536, abstol = reltol = 1e-3: forward, adjoint"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import diffeqflux_b200 as dfx  # noqa: E402
import torch  # noqa: E402
from tools.run_configs import chain_of, mlp_params  # noqa: E402

dev = torch.device("cuda:0")


def node(dims, B, kw, p, u0, dL=None):
    n = dfx.NeuralODE(chain_of(dims), kw.pop("tspan", (0.0, 1.0)), dfx.Tsit5(), **kw)
    sol, back = n.pullback(torch.from_numpy(u0).to(dev), torch.from_numpy(p).to(dev))
    back(torch.ones_like(sol.array()) if dL is None else dL)
    fs, as_ = n.last["stats"], n.last["adjoint_stats"]
    print(dims, B, "fwd", fs.naccept, fs.nreject, fs.nf, round(fs.kernel_ms, 3), "adj", as_.naccept, as_.nreject, as_.nf, round(as_.kernel_ms, 3), "engine", fs.engine, flush=True)


which = sys.argv[1] if len(sys.argv) > 1 else "all"
tol6 = dict(abstol=1e-6, reltol=1e-6, save_everystep=False, save_start=False)
p, u0 = bench.synth(16384)
if which in ("all", "tc"):
    node(bench.DIMS, 16384, dict(tol6, sensealg=dfx.BacksolveAdjoint()), p, u0)
    node(bench.DIMS, 16384, dict(tol6, max_dense_steps=64), p, u0)
if which in ("all", "ffma"):
    node(bench.DIMS, 4096, dict(tol6, sensealg=dfx.BacksolveAdjoint(), path=1), p, u0[:4096])
    p1, rng = mlp_params([2, 50, 2], 1, 1.0)
    ts30 = np.linspace(0, 1.5, 30, dtype=np.float32)
    node([2, 50, 2], 1, dict(saveat=ts30, tspan=(0.0, 1.5)), p1, rng.standard_normal((1, 2)).astype(np.float32))
if which in ("all", "ffjord"):
    dims = [2, 64, 64, 2]
    p3, rng = mlp_params(dims, 3, 0.5)
    B = 65536
    x = rng.standard_normal((B, 2)).astype(np.float32)
    ff = dfx.FFJORD(chain_of(dims), (0.0, 1.0), (2,), dfx.Tsit5(), abstol=1e-3, reltol=1e-3, maxiters=200000)   # (abstol 1e-6 with fresh noise per RHS call needs > 4096 steps: see DESIGN.md)
    st = {"regularize": False, "monte_carlo": True, "seed": 3, "calls": 0}
    logpx, back = ff.pullback(torch.from_numpy(x).to(dev), torch.from_numpy(p3).to(dev), st)
    back(torch.full((B,), -1.0 / B, device=dev))
    fs, as_ = ff.last["stats"], ff.last["adjoint_stats"]
    print("ffjord fwd", fs.naccept, fs.nreject, fs.nf, round(fs.kernel_ms, 3), "adj", as_.naccept, as_.nreject, as_.nf, round(as_.kernel_ms, 3), flush=True)
