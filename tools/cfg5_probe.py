import os, sys
import numpy as np
sys.path.insert(0, "/root/repo")
import torch
import diffeqflux_b200 as dfx
from tools.run_configs import chain_of, mlp_params
dims = [192, 512, 512, 192]
p, rng = mlp_params(dims, 5, 1.0)
u0 = rng.standard_normal((4096, 128)).astype(np.float32)
for tile in (0, 8, 16, 32):
    node = dfx.NeuralODE(chain_of(dims), (0.0, 1.0), dfx.Tsit5(), abstol=1e-3, reltol=1e-3, save_everystep=False, save_start=False, max_dense_steps=64, tile=tile)
    layer = dfx.AugmentedNDELayer(node, 64)
    pt, ut = torch.from_numpy(p).cuda(), torch.from_numpy(u0).cuda()
    for r in range(2):
        sol, back = layer.pullback(ut, pt)
        du0, dp = back(torch.ones_like(sol.array()))
        fs, as_ = sol.stats, layer.last["adjoint_stats"]
        print("tile opt", tile, "rep", r, "fwd", round(fs.kernel_ms, 2), fs.grid, fs.tile, "adj", round(as_.kernel_ms, 2), as_.grid, as_.tile, as_.naccept, flush=True)
