"""2-GPU parity (needs >= 2 GPUs; skipped otherwise): sharded == unsharded step sequence and gradients."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("engine", ["ffma", "tc"])
@pytest.mark.parametrize("sense", [1, 0, 2])          # Backsolve, Interpolating, Quadrature (one parameter-gradient exchange per solve)
def test_two_gpu_shards_follow_the_single_gpu_step_sequence(sense, engine):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(29511 + sense + (3 if engine == "tc" else 0)), os.path.join(ROOT, "tests", "mgpu_worker.py"), str(sense), engine]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    line = [l for l in r.stdout.splitlines() if l.startswith("MGPU_RESULT ")]
    assert line, r.stdout[-2000:] + r.stderr[-4000:]
    for res in json.loads(line[0][len("MGPU_RESULT "):]):
        assert res["engine"] == ([1, 1] if engine == "tc" else [0, 0]), res
        assert res["steps_sharded"] == res["steps_single"], res
        # fp64 partial sums in a fixed order: the sharded run is bitwise the unsharded run
        assert res["state_err"] <= 1e-6 and res["du0_err"] <= 1e-6 and res["dp_err"] <= 1e-6, res
