"""Single-GPU timing of the headline model at shard sizes (strong-scaling shards of B = 16384): engine / tile choice."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import diffeqflux_b200 as dfx
import torch
dev = torch.device("cuda:0")
chain = dfx.Chain(dfx.Dense(64, 256, dfx.tanh), dfx.Dense(256, 256, dfx.tanh), dfx.Dense(256, 64))
for B in (2048, 4096, 8192):
    p, u0 = bench.synth(B)
    p_d, u0_d = torch.from_numpy(p).to(dev), torch.from_numpy(u0).to(dev)
    dL = torch.ones((1, B, 64), device=dev)
    for path, tile in ((2, 0), (1, 0), (1, 16), (1, 32), (1, 64)):
        try:
            node = dfx.NeuralODE(chain, bench.TSPAN, dfx.Tsit5(), abstol=bench.TOL, reltol=bench.TOL, save_everystep=False, save_start=False,
                                 sensealg=dfx.BacksolveAdjoint(), path=path, tile=tile)
            best = 1e9
            for r in range(3):
                sol, back = node.pullback(u0_d, p_d)
                du0, dp = back(dL)
                fs, as_ = node.last["stats"], node.last["adjoint_stats"]
                best = min(best, fs.kernel_ms + as_.kernel_ms)
            print(f"B {B} path {path} tile {tile}: {best:8.2f} ms  grid {fs.grid}/{as_.grid} tile {fs.tile}/{as_.tile} steps {fs.naccept} {as_.naccept}", flush=True)
        except Exception as e:
            print(f"B {B} path {path} tile {tile}: {e!r}"[:150])
