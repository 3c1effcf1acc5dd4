"""CPU (gloo, world_size 2) tests of the host-side multi-GPU logic: handle all_gather plumbing, shard
bounds, and the property the in-kernel exchange relies on -- per-shard fp64 partial sums of the error
norm add up to the unsharded EEst, so every rank replays the same controller decision."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import diffeqflux_b200  # noqa: F401  (the shim registers diffeqflux_jl_b200)
    from diffeqflux_jl_b200 import layers
    handle = bytes([rank + 1]) * 64
    allh = layers.gather_ipc_handles(handle)
    ok = allh == b"".join(bytes([r + 1]) * 64 for r in range(world))
    lo, hi = layers.shard_range(100, rank, world)
    # error-norm partials: each rank sums its shard in fp64, all ranks add the partials in rank order
    rng = np.random.default_rng(0)
    r_full = rng.standard_normal((100, 8)).astype(np.float32)
    part = float(np.sum(np.square(r_full[lo:hi], dtype=np.float32), dtype=np.float64))
    parts = [None] * world
    dist.all_gather_object(parts, part)
    eest = np.float32(np.sqrt(sum(parts) / r_full.size))
    q.put((rank, ok, (lo, hi), float(eest)))
    dist.destroy_process_group()


def test_gloo_two_ranks_host_logic():
    import diffeqflux_b200  # noqa: F401  (registers the package for the workers' import path)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29533, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert [r[1] for r in res] == [True, True]
    assert [r[2] for r in res] == [(0, 50), (50, 100)]
    rng = np.random.default_rng(0)
    r_full = rng.standard_normal((100, 8)).astype(np.float32)
    want = np.float32(np.sqrt(float(np.sum(np.square(r_full, dtype=np.float32), dtype=np.float64)) / r_full.size))
    assert res[0][3] == res[1][3]                       # identical on every rank => identical branch
    assert abs(res[0][3] - float(want)) <= 1e-7 * float(want)


def test_shard_range_requires_equal_shards():
    import diffeqflux_b200  # noqa: F401
    from diffeqflux_jl_b200 import layers
    assert layers.shard_range(16384, 3, 8) == (6144, 8192)
    with pytest.raises(ValueError):
        layers.shard_range(100, 0, 8)
