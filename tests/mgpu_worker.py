"""Worker of tests/test_multigpu.py (launched by torchrun, one rank per GPU): a batch-sharded forward +
adjoint must follow the step sequence of the unsharded solve and return the global parameter gradient."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import diffeqflux_b200 as dfx  # noqa: E402
from diffeqflux_jl_b200 import layers  # noqa: E402


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    rank, world = dist.get_rank(), dist.get_world_size()
    sense = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    tc = len(sys.argv) > 2 and sys.argv[2] == "tc"          # tensor-core engine shape
    rng = np.random.default_rng(3)
    dims = [64, 128, 128, 64] if tc else [16, 64, 64, 16]
    chain = dfx.Chain(dfx.Dense(dims[0], dims[1], dfx.tanh), dfx.Dense(dims[1], dims[2], dfx.tanh), dfx.Dense(dims[2], dims[3]))
    p = np.concatenate([np.concatenate([(rng.standard_normal(i * o) * 3.0 / np.sqrt(i)), rng.standard_normal(o) * 0.3])
                        for i, o in zip(dims[:-1], dims[1:])]).astype(np.float32)
    B = 512
    u0 = rng.standard_normal((B, dims[0])).astype(np.float32)
    dL = np.cos(np.arange(B * dims[0], dtype=np.float32).reshape(1, B, dims[0]) * 0.01)
    sa = {0: dfx.InterpolatingAdjoint(), 1: dfx.BacksolveAdjoint(), 2: dfx.QuadratureAdjoint()}[sense]
    kw = dict(abstol=1e-4, reltol=1e-4, save_everystep=False, save_start=False, sensealg=sa)
    kw.update(dict(path=2) if tc else dict(tile=32))
    node = dfx.NeuralODE(chain, (0.0, 1.0), dfx.Tsit5(), **kw)
    # 1) unsharded reference run on this GPU (before attaching the peers)
    sol1, back1 = node.pullback(u0, p)
    du1, dp1 = back1(dL)
    st1 = (sol1.stats.naccept, sol1.stats.nreject, node.last["adjoint_stats"].naccept, node.last["adjoint_stats"].nreject)
    # 2) sharded run
    layers.attach_process_group(local)
    lo, hi = layers.shard_range(B, rank, world)
    sol, back = node.pullback(u0[lo:hi], p)
    du, dp = back(dL[:, lo:hi])
    st = (sol.stats.naccept, sol.stats.nreject, node.last["adjoint_stats"].naccept, node.last["adjoint_stats"].nreject)

    def rel(a, b):
        return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))

    res = {"rank": rank, "engine": [sol.stats.engine, node.last["adjoint_stats"].engine], "steps_sharded": st, "steps_single": st1,
           "state_err": rel(sol.array(), sol1.array()[:, lo:hi]), "du0_err": rel(du, du1[lo:hi]), "dp_err": rel(dp, dp1)}
    out = [None] * world
    dist.all_gather_object(out, res)
    if rank == 0:
        print("MGPU_RESULT " + json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
