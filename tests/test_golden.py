"""Golden fixtures (tests/golden/*.npz, written by tests/golden/make_golden.py from the oracle).

CPU half (not gpu): the oracle still reproduces the committed vectors -- the parity target cannot drift
silently.  GPU half (-m gpu): the CUDA path, through the layer API / C ABI, against the same committed
numbers; nothing under /root/reference or oracle/ is needed for those comparisons beyond building the
problem description.  Tolerances: states / gradients rtol 1e-4 at solver tolerance 1e-6 (BASELINE.json
north_star), 3 x reltol for looser solves; step counts as in tests/test_gpu_parity.py."""
import os

import numpy as np
import pytest

from oracle import mshoot as omshoot
from oracle import node as onode
from oracle.mlp import ACT_ID, ACT_TANH, MLP

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def close(a, b, rtol, what=""):
    """element-wise |a - b| <= rtol * |b| + rtol * max|b|"""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    scale = max(np.abs(b).max(), 1e-30)
    np.testing.assert_allclose(a, b, rtol=rtol, atol=rtol * scale, err_msg=what)


# ------------------------------------------------------------------------------------------ CPU: oracle vs golden
def test_oracle_reproduces_spiral():
    g = load("spiral")
    mlp = MLP([2, 50, 2], [ACT_TANH, ACT_ID])
    prob = onode.NodeProblem(mlp)
    fw = onode.forward(prob, g["p"], g["u0"], 0.0, 1.5, saveat=g["ts"], dense=True, abstol=1e-6, reltol=1e-3)
    adj = onode.adjoint(prob, g["p"], fw, g["dL"], sensealg=0, abstol=1e-6, reltol=1e-3)
    close(fw.us, g["us"], 1e-6, "us")
    close(adj.dp, g["dp"], 1e-5, "dp")
    close(adj.du0, g["du0"], 1e-5, "du0")
    assert [fw.stats.naccept, fw.stats.nreject, fw.stats.nf] == list(g["fwd_stats"][:3])
    assert [adj.stats.naccept, adj.stats.nreject, adj.stats.nf] == list(g["adj_stats"][:3])


def test_oracle_reproduces_batched():
    g = load("batched")
    mlp = MLP(list(g["dims"]), [ACT_TANH, ACT_TANH, ACT_ID])
    prob = onode.NodeProblem(mlp)
    fw = onode.forward(prob, g["p"], g["u0"], 0.0, 1.0, save_everystep=False, save_start=False, abstol=1e-6, reltol=1e-6)
    adj = onode.adjoint(prob, g["p"], fw, np.ones_like(fw.us), sensealg=1, abstol=1e-6, reltol=1e-6)
    close(fw.us, g["us"], 1e-6, "us")
    close(adj.dp, g["dp"], 1e-5, "dp")
    assert [fw.stats.naccept, fw.stats.nreject] == list(g["fwd_stats"][:2])


def test_oracle_reproduces_mshoot_and_ranges():
    g = load("mshoot")
    assert np.array_equal(np.array(omshoot.group_ranges(9, 3)), g["ranges"])
    assert np.array_equal(np.array(omshoot.group_ranges(21, 5)), g["ranges_21_5"])
    mlp = MLP([4, 16, 4], [ACT_TANH, ACT_ID])
    loss, preds, dp, _ = omshoot.multiple_shoot(onode.NodeProblem(mlp), g["p"], g["data"], g["ts"], 3, abstol=1e-5, reltol=1e-5)
    assert abs(loss - float(g["loss"])) <= 1e-6 * abs(float(g["loss"]))
    close(dp, g["dp"], 1e-5, "dp")


def test_headline_golden_is_consistent_with_bench_inputs():
    """tests/golden/headline.npz (full-size oracle runs, `make_golden.py headline`) belongs to bench.py's current
    inputs and records the three accumulation models in the documented order blas >= defer >= exactgemm."""
    import zlib
    import bench
    g = load("headline")
    p, u0 = bench.synth(int(g["batch"]))
    assert zlib.crc32(p.tobytes()) == int(g["p_crc"]) and zlib.crc32(u0.tobytes()) == int(g["u0_crc"])
    assert int(g["batch"]) == bench.BATCH == 16384
    for m in ("blas", "defer", "exactgemm"):
        f, a = g[m + "_fwd_stats"], g[m + "_adj_stats"]
        assert f[1] == 0 and a[1] == 0 and f[3] == 0 and a[3] == 0
        assert f[2] == 3 + 6 * f[0] and a[2] == 3 + 6 * a[0]
        assert len(g[m + "_adj_trace"]) == a[0]
    assert g["blas_adj_stats"][0] >= g["defer_adj_stats"][0] >= g["exactgemm_adj_stats"][0]
    assert list(g["defer_fwd_stats"][:3]) == [19, 0, 117] and list(g["defer_adj_stats"][:3]) == [36, 0, 219]
    assert g["dp"].shape == (p.size,)


def test_oracle_accumulation_models_agree_at_small_batch():
    """The oracle's accumulation switches (oracle/mlp.py: wgrad_accum / gemm_accum) change only rounding: at B = 256
    the gradients of all models agree to 2e-6 and the per-tile deferred sums reproduce the plain totals."""
    rng = np.random.default_rng(5)
    dims = [32, 64, 32]
    base = MLP(dims, [ACT_TANH, ACT_ID])
    p = base.init(rng, scale=2.0)
    u0 = rng.standard_normal((256, 32)).astype(np.float32)
    kw = dict(abstol=1e-6, reltol=1e-6)
    outs = {}
    for name, mk in (("blas", {}), ("fp64", dict(wgrad_accum="fp64")), ("tile", dict(wgrad_accum=("tile", 128))),
                     ("defer", dict(wgrad_accum=("tile_defer", 128))), ("3xtf32", dict(wgrad_accum=("tile_defer", 128), gemm_accum="3xtf32"))):
        prob = onode.NodeProblem(MLP(dims, [ACT_TANH, ACT_ID], **mk))
        fw = onode.forward(prob, p, u0, 0.0, 1.0, save_everystep=False, save_start=False, **kw)
        adj = onode.adjoint(prob, p, fw, np.ones_like(fw.us), sensealg=1, **kw)
        assert adj.dp.shape == p.shape and adj.stats.retcode == 0
        outs[name] = adj
    for name, adj in outs.items():
        close(adj.dp, outs["fp64"].dp, 2e-6 if name != "3xtf32" else 1e-5, f"dp {name}")
        close(adj.du0, outs["fp64"].du0, 2e-5, f"du0 {name}")
        assert abs(adj.stats.naccept - outs["fp64"].stats.naccept) <= 2


# ------------------------------------------------------------------------------------------ GPU: CUDA path vs golden
@pytest.mark.gpu
def test_gpu_spiral_matches_golden():
    import diffeqflux_b200 as dfx
    g = load("spiral")
    chain = dfx.Chain(dfx.Dense(2, 50, dfx.tanh), dfx.Dense(50, 2))
    node = dfx.NeuralODE(chain, (0.0, 1.5), dfx.Tsit5(), saveat=g["ts"], abstol=1e-6, reltol=1e-3)
    sol, back = node.pullback(g["u0"], g["p"])
    assert sol.retcode == 0
    close(sol.array(), g["us"], 3e-3, "us")           # 3 x reltol: two valid fp32 step sequences
    du0, dp = back(g["dL"])
    close(dp, g["dp"], 3e-3, "dp")
    close(np.atleast_2d(du0), g["du0"], 3e-3, "du0")
    n_g, n_o = sol.stats.naccept + sol.stats.nreject, int(g["fwd_stats"][0] + g["fwd_stats"][1])
    assert abs(n_g - n_o) <= max(3, n_o // 4), (n_g, n_o)


@pytest.mark.gpu
@pytest.mark.parametrize("path", [1, 2])              # 1: FP32-FFMA engine, 2: tcgen05 3xTF32 engine
def test_gpu_batched_backsolve_matches_golden(path):
    import diffeqflux_b200 as dfx
    g = load("batched")
    d = list(g["dims"])
    chain = dfx.Chain(dfx.Dense(d[0], d[1], dfx.tanh), dfx.Dense(d[1], d[2], dfx.tanh), dfx.Dense(d[2], d[3]))
    node = dfx.NeuralODE(chain, (0.0, 1.0), dfx.Tsit5(), save_everystep=False, save_start=False, abstol=1e-6, reltol=1e-6,
                         sensealg=dfx.BacksolveAdjoint(), path=path)
    sol, back = node.pullback(g["u0"], g["p"])
    assert sol.retcode == 0 and sol.stats.engine == (1 if path == 2 else 0)
    close(sol.array(), g["us"], 1e-4, "us")
    du0, dp = back(np.ones_like(g["us"]))
    close(dp, g["dp"], 1e-4, "dp")
    close(du0, g["du0"], 1e-4, "du0")
    # exact accepted / rejected / nf on BOTH engines, forward and adjoint (this configuration's controller is not in
    # the rounding-noise regime for long enough to move a step: see DESIGN.md section 2)
    ast = node.last["adjoint_stats"]
    assert [sol.stats.naccept, sol.stats.nreject, sol.stats.nf] == [int(x) for x in g["fwd_stats"][:3]]
    assert [ast.naccept, ast.nreject, ast.nf] == [int(x) for x in g["adj_stats"][:3]]


@pytest.mark.gpu
def test_gpu_ffjord_matches_golden():
    import diffeqflux_b200 as dfx
    g = load("ffjord")
    chain = dfx.Chain(dfx.Dense(2, 16, dfx.tanh), dfx.Dense(16, 16, dfx.tanh), dfx.Dense(16, 2))
    ff = dfx.FFJORD(chain, (0.0, 0.5), (2,), dfx.Tsit5(), abstol=1e-3, reltol=1e-3)
    st = {"regularize": False, "monte_carlo": True, "seed": int(g["seed"]), "calls": 0}
    logpx, back = ff.pullback(g["x"], g["p"], st)
    close(logpx, g["logpx"], 1e-4, "logpx")
    dx, dp = back(g["g"])
    close(dp, g["dp"], 1e-3, "dp")
    close(dx, g["dx"], 1e-3, "dx")
    assert ff.last["stats"].naccept == int(g["fwd_stats"][0])


@pytest.mark.gpu
def test_gpu_mshoot_matches_golden():
    import torch
    import diffeqflux_b200 as dfx
    g = load("mshoot")
    # the layer API returns inclusive (first, last) pairs, the oracle half-open ranges
    assert np.array_equal(np.array(dfx.group_ranges(9, 3)), g["ranges"] - np.array([0, 1]))
    assert np.array_equal(np.array(dfx.group_ranges(21, 5)), g["ranges_21_5"] - np.array([0, 1]))
    chain = dfx.Chain(dfx.Dense(4, 16, dfx.tanh), dfx.Dense(16, 4))
    pt = torch.tensor(g["p"], requires_grad=True)
    prob = dfx.ODEProblem(chain, g["u0"], (0.0, 2.0), g["p"])
    loss, preds = dfx.multiple_shoot(pt, torch.tensor(g["data"]), g["ts"], prob, lambda u, uh: ((u - uh) ** 2).sum(), dfx.Tsit5(), 3,
                                     abstol=1e-5, reltol=1e-5)
    assert abs(loss.item() - float(g["loss"])) <= 1e-4 * abs(float(g["loss"]))
    close(preds[-1].detach().numpy(), g["pred_last"], 1e-4, "last group")
    loss.backward()
    close(pt.grad.numpy(), g["dp"], 2e-4, "dp")
