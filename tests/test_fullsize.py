"""BASELINE.json configs[1] at its full size (batch 16384, 64 -> 256 -> 256 -> 64 tanh, Tsit5 1e-6, BacksolveAdjoint).

test_headline_config_against_the_oracle_golden compares both RHS engines with the oracle on bench.py's exact inputs
(tests/golden/headline.npz, written by `tests/golden/make_golden.py headline`: the oracle needs ~1.5 minutes at this
size, so its outputs are committed instead of recomputed on the GPU box).  The other tests check size-independent
properties:

  * determinism: two runs of the tensor-core path return bitwise equal states and gradients;
  * engine agreement: the tcgen05 3xTF32 kernels and the FP32-FFMA kernels (which the oracle comparisons of
    tests/test_gpu_parity.py pin at small batches) take the same steps and agree to rtol 1e-4;
  * the parameter gradient is the derivative of the loss: a central finite difference of the forward solve along a
    random direction matches <dp, v> (fp32 forward differences: 1 % tolerance);
  * time reversal: the Backsolve adjoint's own reconstruction of the trajectory feeds d(u0); comparing du0 between
    the Backsolve and Interpolating adjoints (the latter reads the stored forward tape) bounds that reconstruction."""
import numpy as np
import pytest

import diffeqflux_b200 as dfx

pytestmark = pytest.mark.gpu

DIMS = [64, 256, 256, 64]
B = 16384


def setup_problem():
    rng = np.random.default_rng(2)
    p = []
    for i, o in zip(DIMS[:-1], DIMS[1:]):
        p.append((rng.standard_normal(i * o) * 2.0 / np.sqrt(i)).astype(np.float32))
        p.append((rng.standard_normal(o) * 0.2).astype(np.float32))
    p = np.concatenate(p)
    u0 = rng.standard_normal((B, DIMS[0])).astype(np.float32)
    chain = dfx.Chain(dfx.Dense(64, 256, dfx.tanh), dfx.Dense(256, 256, dfx.tanh), dfx.Dense(256, 64))
    return rng, p, u0, chain


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def run(chain, p, u0, path, sense=None, dL=None):
    kw = dict(abstol=1e-6, reltol=1e-6, save_everystep=False, save_start=False, path=path,
              sensealg=sense or dfx.BacksolveAdjoint(), max_dense_steps=64)
    node = dfx.NeuralODE(chain, (0.0, 1.0), dfx.Tsit5(), **kw)
    sol, back = node.pullback(u0, p)
    assert sol.retcode == 0
    du0, dp = back(np.ones((1, B, DIMS[-1]), np.float32) if dL is None else dL)
    return sol, du0, dp, node.last["adjoint_stats"]


def test_headline_config_against_the_oracle_golden():
    """bench.py's inputs (seed 2, W_SCALE 2, B = 16384), both engines, against the oracle run at full size.

    Golden models (oracle/mlp.py switches): "defer" = the CUDA path's documented accumulation order of the parameter
    gradient (fp32 per 128-trajectory tile, stages combined per tile, tiles summed in fp64) -- the parity target;
    "blas" = one fp32 sgemm over all 16384 trajectories (a CPU BLAS); "exactgemm" = layer GEMMs accumulated exactly.
    Forward: 19 accepted / 0 rejected / nf 117 in every model and on both engines.  Adjoint: the first ~15 steps ramp
    dt up from Hairer's first step (1.6e-3: mu(t1) = 0 against abstol 1e-6) with an embedded error estimate of
    4e-3..1e-2 that is pure rounding noise (the truncation error of those steps is ~1e-8), so the ramp length follows
    the arithmetic's noise floor: blas 38, defer 36, exactgemm 34 accepted steps.  The FP32-FFMA engine has the
    "defer" model's arithmetic and must reproduce it exactly (36 / 0 / nf 219, per-step error estimates within
    10 %); the tcgen05 engine's 3xTF32 GEMMs round less often than an FMA chain and land between "defer" and
    "exactgemm" (35).  States and gradients: element-wise rtol 1e-4 on both engines."""
    import os
    import zlib
    import bench
    from diffeqflux_jl_b200 import layers
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "headline.npz"))
    p, u0 = bench.synth(bench.BATCH)
    assert zlib.crc32(p.tobytes()) == int(g["p_crc"]) and zlib.crc32(u0.tobytes()) == int(g["u0_crc"]), "bench.synth changed: regenerate the golden"
    chain = dfx.Chain(dfx.Dense(64, 256, dfx.tanh), dfx.Dense(256, 256, dfx.tanh), dfx.Dense(256, 64))
    ctx = layers._context(0)
    ctx.trace(capacity=256)
    lo, hi = int(g["exactgemm_adj_stats"][0]), int(g["defer_adj_stats"][0])
    assert lo <= hi <= int(g["blas_adj_stats"][0])
    for path in (1, 2):
        node = dfx.NeuralODE(chain, bench.TSPAN, dfx.Tsit5(), abstol=bench.TOL, reltol=bench.TOL, save_everystep=False, save_start=False,
                             sensealg=dfx.BacksolveAdjoint(), path=path)
        sol, back = node.pullback(u0, p)
        tr_f = ctx.trace().copy()
        du0, dp = back(np.ones_like(sol.array()))
        tr_a = ctx.trace().copy()
        fs, as_ = node.last["stats"], node.last["adjoint_stats"]
        assert fs.engine == path - 1 and as_.engine == path - 1
        assert (fs.naccept, fs.nreject, fs.nf) == tuple(int(x) for x in g["defer_fwd_stats"][:3]), f"forward counts, path {path}"
        if path == 1:
            assert (as_.naccept, as_.nreject, as_.nf) == tuple(int(x) for x in g["defer_adj_stats"][:3]), "adjoint counts, FFMA engine"
            for tr, ref, what in ((tr_f, g["defer_fwd_trace"], "forward"), (tr_a, g["defer_adj_trace"], "adjoint")):
                assert len(tr) == len(ref)
                # (the last step is the remainder to the end of the span: its dt and estimate are not comparable)
                np.testing.assert_allclose(tr[:-1, 1], ref[:-1, 1], rtol=0.03, atol=0, err_msg=f"{what} step sizes")
                np.testing.assert_allclose(tr[:-1, 2], ref[:-1, 2], rtol=0.10, atol=0, err_msg=f"{what} per-step error estimates")
        else:
            assert lo <= as_.naccept <= hi and as_.nreject == 0 and as_.nf == 3 + 6 * as_.naccept, f"adjoint counts, tcgen05 engine: {as_}"
        scale = float(np.abs(g["dp"]).max())
        np.testing.assert_allclose(dp, g["dp"], rtol=1e-4, atol=1e-4 * scale, err_msg=f"dp path {path}")
        np.testing.assert_allclose(du0[::64], g["du0_s"], rtol=1e-4, atol=1e-4 * float(np.abs(g["du0_s"]).max()), err_msg=f"du0 path {path}")
        np.testing.assert_allclose(sol.array()[-1][::64], g["uT_s"], rtol=1e-4, atol=1e-4 * float(np.abs(g["uT_s"]).max()), err_msg=f"u(t1) path {path}")
    ctx.trace(capacity=0)


def test_full_size_determinism_and_engine_agreement():
    rng, p, u0, chain = setup_problem()
    s1, du1, dp1, a1 = run(chain, p, u0, 2)
    s2, du2, dp2, a2 = run(chain, p, u0, 2)
    assert s1.stats.engine == 1 and a1.engine == 1
    assert np.array_equal(s1.array(), s2.array()) and np.array_equal(dp1, dp2) and np.array_equal(du1, du2)
    sf, duf, dpf, af = run(chain, p, u0, 1)
    assert sf.stats.engine == 0
    assert (s1.stats.naccept, s1.stats.nreject) == (sf.stats.naccept, sf.stats.nreject)
    assert abs(a1.naccept - af.naccept) <= 1 and a1.nreject == af.nreject
    assert rel(s1.array(), sf.array()) <= 1e-4
    assert rel(dp1, dpf) <= 1e-4
    assert rel(du1, duf) <= 1e-4


def test_full_size_gradient_is_the_derivative_of_the_loss():
    rng, p, u0, chain = setup_problem()
    _, _, dp, _ = run(chain, p, u0, 2)                         # loss = sum of the end state
    v = rng.standard_normal(p.size).astype(np.float32)
    v /= np.linalg.norm(v)
    eps = 2e-3
    node = dfx.NeuralODE(chain, (0.0, 1.0), dfx.Tsit5(), abstol=1e-6, reltol=1e-6, save_everystep=False, save_start=False, path=2,
                         sensealg=dfx.BacksolveAdjoint())

    def loss(q):
        sol, _ = node(u0, q.astype(np.float32), {})
        return float(np.sum(sol.array(), dtype=np.float64))

    fd = (loss(p + eps * v) - loss(p - eps * v)) / (2 * eps)
    an = float(np.dot(dp.astype(np.float64), v.astype(np.float64)))
    assert abs(fd - an) <= 1e-2 * abs(an), (fd, an)


def test_full_size_backsolve_matches_interpolating():
    rng, p, u0, chain = setup_problem()
    _, du_b, dp_b, _ = run(chain, p, u0, 2)
    _, du_i, dp_i, _ = run(chain, p, u0, 2, sense=dfx.InterpolatingAdjoint())
    # two different adjoint formulations at tolerance 1e-6: agreement bounded by the solver tolerance, not by fp32
    assert rel(dp_b, dp_i) <= 2e-4
    assert rel(du_b, du_i) <= 2e-4


def test_more_tiles_than_sms_tensor_core_path():
    """B = 19200 > 148 * 128: the tensor-core kernels walk several 128-trajectory tiles per CTA (the per-CTA gradient slabs
    then accumulate over the CTA's tiles).  Small eligible model so the oracle stays cheap; both engines against it."""
    from oracle import node as onode
    from oracle.mlp import ACT_ID, ACT_TANH, MLP
    rng = np.random.default_rng(12)
    dims = [32, 64, 32]
    mlp = MLP(dims, [ACT_TANH, ACT_ID], wgrad_accum=("tile_defer", 128))
    p = mlp.init(rng, scale=2.0)
    nb = 19200
    u0 = rng.standard_normal((nb, 32)).astype(np.float32)
    chain = dfx.Chain(dfx.Dense(32, 64, dfx.tanh), dfx.Dense(64, 32))
    kw = dict(abstol=1e-6, reltol=1e-6, save_everystep=False, save_start=False)
    prob = onode.NodeProblem(mlp)
    fw = onode.forward(prob, p, u0, 0.0, 1.0, **kw)
    adj = onode.adjoint(prob, p, fw, np.ones_like(fw.us), sensealg=1, abstol=1e-6, reltol=1e-6)
    for path in (1, 2):
        node = dfx.NeuralODE(chain, (0.0, 1.0), dfx.Tsit5(), sensealg=dfx.BacksolveAdjoint(), path=path, **kw)
        sol, back = node.pullback(u0, p)
        du0, dp = back(np.ones_like(sol.array()))
        ast = node.last["adjoint_stats"]
        assert sol.stats.engine == path - 1 and ast.engine == path - 1
        if path == 2:
            assert sol.stats.grid == 148 and sol.stats.tile == 128          # 150 tiles on 148 CTAs
        assert abs(sol.stats.naccept - fw.stats.naccept) <= (0 if path == 1 else 1) and sol.stats.nreject == fw.stats.nreject
        assert abs(ast.naccept - adj.stats.naccept) <= max(2, -(-8 * adj.stats.naccept // 100))
        for got, want, what in ((sol.array(), fw.us, "u(t1)"), (dp, adj.dp, "dp"), (du0, adj.du0, "du0")):
            np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-4 * float(np.abs(want).max()), err_msg=f"{what} path {path}")
