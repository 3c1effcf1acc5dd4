import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import diffeqflux_b200 as dfx
from tools.run_configs import chain_of, mlp_params
dims = [2, 64, 64, 2]
p, rng = mlp_params(dims, 3, 0.5)
B = 65536
x = rng.standard_normal((B, 2)).astype(np.float32)
ff = dfx.FFJORD(chain_of(dims), (0.0, 1.0), (2,), dfx.Tsit5(), abstol=1e-3, reltol=1e-3, maxiters=200000)
st = {"regularize": False, "monte_carlo": True, "seed": 3, "calls": 0}
for r in range(2):
    logpx, back = ff.pullback(torch.from_numpy(x).cuda(), torch.from_numpy(p).cuda(), st)
    dx, dp = back(torch.full((B,), -1.0 / B, device="cuda"))
    fs, as_ = ff.last["stats"], ff.last["adjoint_stats"]
    F = 2 * sum(i * o for i, o in zip(dims[:-1], dims[1:]))
    print("ffjord fwd", fs.naccept, fs.nreject, fs.nf, round(fs.kernel_ms, 2), "ms tile", fs.tile, "grid", fs.grid, "| adj", as_.naccept, as_.nf, round(as_.kernel_ms, 2),
          "| fwd TFLOP/s", round(B * fs.nf * 2 * F / (fs.kernel_ms * 1e-3) / 1e12, 2), "logpx sum", float(logpx.double().sum()), "dp sum", float(dp.double().sum()), flush=True)
