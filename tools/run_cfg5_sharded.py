#!/usr/bin/env python
"""BASELINE.json configs[4]: augmented NeuralODE (state 128 + 64 augment, 192 -> 512 -> 512 -> 192, batch 32768) sharded
over the ranks of one node (torchrun, one rank per GPU; 4096 trajectories per rank at 8 GPUs), InterpolatingAdjoint,
tolerances 1e-3 (docs/src/examples/augmented_neural_ode.md:47-50).  Prints one JSON line on rank 0."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import diffeqflux_b200 as dfx  # noqa: E402
from diffeqflux_jl_b200 import layers  # noqa: E402
from tools.run_configs import chain_of, flops, mlp_params  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
        layers.attach_process_group(local)
    rank = dist.get_rank() if world > 1 else 0
    dims = [192, 512, 512, 192]
    total = 32768
    B = total // world
    p, _ = mlp_params(dims, 5, 1.0)
    u0 = np.random.default_rng(1005).standard_normal((total, 128)).astype(np.float32)[rank * B:(rank + 1) * B]
    node = dfx.NeuralODE(chain_of(dims), (0.0, 1.0), dfx.Tsit5(), abstol=1e-3, reltol=1e-3, save_everystep=False, save_start=False, max_dense_steps=64)
    layer = dfx.AugmentedNDELayer(node, 64)
    pt, ut = torch.from_numpy(p).cuda(), torch.from_numpy(u0).cuda()
    best = None
    for _ in range(3):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        sol, back = layer.pullback(ut, pt)
        du0, dp = back(torch.ones_like(sol.array()))
        torch.cuda.synchronize(); wall = time.perf_counter() - t0
        best = wall if best is None else min(best, wall)
    t = torch.tensor([best], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    fs, as_ = sol.stats, layer.last["adjoint_stats"]
    if rank == 0:
        F = flops(dims)
        wall = float(t.item())
        print(json.dumps({"config": "cfg5 augmented 128+64 -> 192->512->512->192, batch 32768 sharded", "n_gpus": world, "batch_per_gpu": B,
                          "wall_ms": wall * 1e3, "evals_per_s": total * (fs.nf + as_.nf) / wall, "fwd": [fs.naccept, fs.nreject, fs.nf, round(fs.kernel_ms, 2), fs.engine],
                          "adj": [as_.naccept, as_.nreject, as_.nf, round(as_.kernel_ms, 2), as_.engine],
                          "tflops_per_gpu": B * (fs.nf * F + as_.nf * 3 * F) / wall / 1e12, "dp_checksum": float(dp.double().sum())}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
