"""FFJORD adjoint (tile engine) time against the forced tile size; forward shown for the tile engine too."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import diffeqflux_b200 as dfx
from tools.run_configs import chain_of, mlp_params
dims = [2, 64, 64, 2]
p, rng = mlp_params(dims, 3, 0.5)
B = 65536
x = rng.standard_normal((B, 2)).astype(np.float32)
st = {"regularize": False, "monte_carlo": True, "seed": 3, "calls": 0}
for tile in (0, 128, 256):
    kw = dict(tile=tile) if tile else {}
    ff = dfx.FFJORD(chain_of(dims), (0.0, 1.0), (2,), dfx.Tsit5(), abstol=1e-3, reltol=1e-3, maxiters=200000, **kw)
    for r in range(2):
        logpx, back = ff.pullback(torch.from_numpy(x).cuda(), torch.from_numpy(p).cuda(), st)
        dx, dp = back(torch.full((B,), -1.0 / B, device="cuda"))
    fs, as_ = ff.last["stats"], ff.last["adjoint_stats"]
    print("tile opt", tile, "| fwd", fs.nf, round(fs.kernel_ms, 2), "ms tile", fs.tile, "grid", fs.grid, "| adj", as_.nf, round(as_.kernel_ms, 2), "ms tile", as_.tile, "grid", as_.grid,
          "dp sum", float(dp.double().sum()), flush=True)
