"""GPU parity tests: the CUDA path (through the C ABI / the layer API) against the oracle on the same
seeded inputs.  Bar (BASELINE.json north_star): states and gradients within rtol 1e-4 (fp32) and an
identical accepted/rejected step count.

Step counts are compared exactly whenever the oracle's own count survives few-ulp perturbations of
its inputs.  Where it does not (error estimate = rounding noise: tiny dt while the controller ramps
up from Hairer's initial step), no two fp32 implementations share a step sequence -- that includes
Julia with a different BLAS -- and the test then only bounds the difference (DESIGN.md "step-count
parity")."""
import numpy as np
import pytest

import diffeqflux_b200 as dfx
from oracle import mshoot as omshoot
from oracle import node as onode
from oracle.mlp import ACT_ID, ACT_TANH, MLP, PRE_CUBE

pytestmark = pytest.mark.gpu

RTOL = 1e-4


def close(a, b, rtol=RTOL, what=""):
    """Element-wise |a - b| <= rtol * |b| + rtol * max|b| (numpy assert_allclose with atol = rtol * scale: entries
    much smaller than the array's scale carry the absolute rounding error of the sums they come from)."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    scale = max(np.abs(b).max(), 1e-30)
    np.testing.assert_allclose(a, b, rtol=rtol, atol=rtol * scale, err_msg=what)


def make(dims, seed, scale=1.0, pre=False):
    rng = np.random.default_rng(seed)
    acts = [ACT_TANH] * (len(dims) - 2) + [ACT_ID]
    mlp = MLP(dims, acts, pre=PRE_CUBE if pre else 0)
    p = mlp.init(rng, scale=scale)
    layers = [dfx.Dense(i, o, dfx.tanh if a == ACT_TANH else dfx.identity) for i, o, a in zip(dims[:-1], dims[1:], acts)]
    chain = dfx.Chain(*([dfx.cube] if pre else []), *layers)
    return rng, mlp, p, chain


def perturbed(p, seed, k=4):
    r = np.random.default_rng(seed).integers(-k, k + 1, p.size)
    return (p * (1 + r * np.float32(6e-8))).astype(np.float32)


def counts(st):
    return (st.naccept, st.nreject, st.nf)


def check_steps(gpu, orc, what, exact, spread=None):
    """exact: identical (accepted, rejected, nf).  Otherwise the configuration is in the rounding-noise / tstop-remainder
    regime where the count follows the arithmetic's noise floor: within 15 % of the oracle's, or --
    where `spread` holds the oracle's own counts under few-ulp perturbations of the weights -- within that spread
    (widened by 25 %: a handful of perturbations does not exhaust it)."""
    g, o = counts(gpu), counts(orc)
    if exact:
        assert g == o, f"{what}: gpu {g} oracle {o}"
        return
    n_g, n_o = g[0] + g[1], o[0] + o[1]
    if spread:
        lo, hi = min(spread + [n_o]), max(spread + [n_o])
        assert 0.75 * lo - 2 <= n_g <= 1.25 * hi + 2, f"{what}: gpu {g} oracle {o}, oracle spread under 4-ulp perturbations {sorted(spread)}"
    else:
        assert abs(n_g - n_o) <= max(2, -(-15 * n_o // 100)), f"{what}: gpu {g} oracle {o} (noise-dominated controller)"


def node_case(chain, mlp, p, u0, tspan, kw, dL_fn=None, augment=0, noisy=False, rtol=None, adj_exact=True):
    """Run the layer API and the oracle on the same inputs; compare states, step counts, gradients.

    noisy=True marks configurations whose embedded error estimate is rounding noise (a handful of
    trajectories, dt still ramping up from Hairer's initial step): the oracle's own step sequence
    then changes under few-ulp input perturbations, so only the results are compared tightly."""
    sense = kw.get("sensealg", dfx.InterpolatingAdjoint()).code
    okw = {k: v for k, v in kw.items() if k not in ("sensealg", "tile", "dt", "path")}
    if "dt" in kw:
        okw["dt0"] = kw["dt"]
    tol = dict(abstol=kw.get("abstol", 1e-6), reltol=kw.get("reltol", 1e-3), dt0=kw.get("dt"))
    if rtol is None:
        # the bar is 1e-4 at the north star's tolerance (1e-6).  Two valid fp32 step sequences are each
        # only reltol-accurate, so a looser solver tolerance bounds how close they can be.
        rtol = max(RTOL, (10.0 if noisy else 3.0) * tol["reltol"]) if (noisy or "reltol" in kw) else RTOL
    prob = onode.NodeProblem(mlp)
    node = dfx.NeuralODE(chain, tspan, dfx.Tsit5(), **kw)
    layer = dfx.AugmentedNDELayer(node, augment) if augment else node
    sol, back = layer.pullback(u0, p)
    vec = np.asarray(u0).ndim == 1
    fw = onode.forward(prob, p, u0, tspan[0], tspan[1], dense=(sense == 0), augment=augment, **okw)
    dL = np.ones_like(fw.us) if dL_fn is None else dL_fn(fw.us)
    adj = onode.adjoint(prob, p, fw, dL, sensealg=sense, augment=augment, **tol)
    assert sol.retcode == 0 and fw.stats.retcode == 0
    spread_f, spread_a = [], []
    if noisy:          # how far the oracle's own step counts move under 4-ulp perturbations of the weights
        for seed in (1, 2, 3, 4):
            q = perturbed(p, seed)
            f2 = onode.forward(prob, q, u0, tspan[0], tspan[1], dense=(sense == 0), augment=augment, **okw)
            a2 = onode.adjoint(prob, q, f2, np.ones_like(f2.us) if dL_fn is None else dL_fn(f2.us), sensealg=sense, augment=augment, **tol)
            spread_f.append(f2.stats.naccept + f2.stats.nreject)
            spread_a.append(a2.stats.naccept + a2.stats.nreject)
    check_steps(sol.stats, fw.stats, "forward", not noisy, spread_f)
    arr = sol.array()[:, None, :] if vec else sol.array()
    if arr.shape != fw.us.shape:          # save_everystep with a different (noisy) step count
        assert noisy and kw.get("saveat") is None
        return sol, fw, adj
    every = kw.get("saveat") is None and kw.get("save_everystep", True)
    if every and not noisy:
        # accepted-step times of two fp32 runs agree only to the controller's sensitivity (~1e-4)
        np.testing.assert_allclose(np.asarray(sol.t), fw.ts, rtol=1e-3, atol=1e-6)
        close(arr, fw.us, rtol=max(rtol, 1e-3), what="states")
    elif not every:
        np.testing.assert_allclose(np.asarray(sol.t), fw.ts, rtol=1e-6, atol=1e-7)
        close(arr, fw.us, rtol=rtol, what="states")
    du0, dp = back(dL[:, 0, :] if vec else dL)
    check_steps(layer.last["adjoint_stats"], adj.stats, "adjoint", adj_exact and not noisy, spread_a)
    if every and not noisy:
        rtol = max(rtol, 1e-3)
    if not (noisy and every):
        close(dp, adj.dp, rtol=rtol, what="dp")
        close(np.atleast_2d(du0), adj.du0, rtol=rtol, what="du0")
    assert np.abs(dp).max() > 0 and np.abs(du0).max() > 0        # the reference's own assertion
    return sol, fw, adj


KW = [
    dict(save_everystep=False, save_start=False),
    dict(save_everystep=False, save_start=False, sensealg=dfx.BacksolveAdjoint()),
    dict(saveat=np.arange(0, 1.01, 0.1, dtype=np.float32)),
    dict(saveat=0.1),
    dict(abstol=1e-6, reltol=1e-6, saveat=[0.3, 0.7]),
    dict(saveat=np.arange(0, 1.01, 0.1, dtype=np.float32), sensealg=dfx.BacksolveAdjoint()),
]


@pytest.mark.parametrize("kw", KW)
@pytest.mark.parametrize("vector", [True, False])
def test_reference_kwarg_matrix(kw, vector):
    """test/BasicNeuralDE/neural_de_tests.jl:27-42 with u0 a vector and the 2x3 matrix, Dense(2,50,tanh)."""
    rng, mlp, p, chain = make([2, 50, 2], 1)
    u0 = np.array([2.0, 0.0], np.float32) if vector else np.array([[0, 0], [1, 0], [2, 0]], np.float32)
    node_case(chain, mlp, p, u0, (0.0, 1.0), kw, noisy=True)


def test_save_everystep_returns_the_step_sequence():
    rng, mlp, p, chain = make([16, 64, 64, 16], 3, scale=3.0)
    u0 = rng.standard_normal((100, 16)).astype(np.float32)
    sol, fw, _ = node_case(chain, mlp, p, u0, (0.0, 1.0), dict(abstol=1e-4, reltol=1e-4), adj_exact=False)
    assert len(sol) == sol.stats.naccept + 1 == fw.stats.naccept + 1


def test_cube_prelayer_docs_example():
    """docs/src/examples/neural_ode.md:19-34: Chain(x -> x.^3, Dense(2, 50, tanh), Dense(50, 2)),
    tspan (0, 1.5), 30 saveat points, loss sum(abs2, data - pred)."""
    rng, mlp, p, chain = make([2, 50, 2], 2, pre=True)
    u0 = np.array([2.0, 0.0], np.float32)
    ts = np.linspace(0, 1.5, 30, dtype=np.float32)
    node_case(chain, mlp, p, u0, (0.0, 1.5), dict(saveat=ts), dL_fn=lambda us: 2 * (us - 1.0), noisy=True)


@pytest.mark.parametrize("sense", [0, 1])
@pytest.mark.parametrize("B,tile", [(100, 0), (1000, 32), (1000, 0)])
def test_batched_node_mid_size(sense, B, tile):
    """ragged batch (not a multiple of the tile), several tiles per CTA, 3 Dense layers; weights scaled
    so that the truncation error (not rounding noise) drives the controller."""
    rng, mlp, p, chain = make([16, 64, 64, 16], 3, scale=3.0)
    u0 = rng.standard_normal((B, 16)).astype(np.float32)
    sa = dfx.InterpolatingAdjoint() if sense == 0 else dfx.BacksolveAdjoint()
    kw = dict(abstol=1e-4, reltol=1e-4, save_everystep=False, save_start=False, sensealg=sa, tile=tile)
    node_case(chain, mlp, p, u0, (0.0, 1.0), kw, dL_fn=lambda us: np.cos(us).astype(np.float32))


def test_rejected_steps_and_user_dt():
    """a user dt0 far too large: the first attempts are rejected and the reject branch of the PI
    controller (dt /= min(1/qmin, q11/gamma)) is exercised."""
    rng, mlp, p, chain = make([16, 64, 64, 16], 3, scale=3.0)
    u0 = rng.standard_normal((200, 16)).astype(np.float32)
    kw = dict(abstol=1e-4, reltol=1e-4, saveat=[0.5], dt=1.0)
    sol, fw, adj = node_case(chain, mlp, p, u0, (0.0, 1.0), kw)
    assert fw.stats.nreject >= 1 and sol.stats.nreject == fw.stats.nreject


def test_headline_shape_small_batch():
    """BASELINE config 2 shape (64 -> 256 -> 256 -> 64 tanh, tol 1e-6, BacksolveAdjoint, end state only)
    at a batch the oracle finishes in seconds."""
    rng, mlp, p, chain = make([64, 256, 256, 64], 3, scale=2.5)
    u0 = rng.standard_normal((256, 64)).astype(np.float32)
    kw = dict(abstol=1e-6, reltol=1e-6, save_everystep=False, save_start=False, sensealg=dfx.BacksolveAdjoint(), path=1)
    node_case(chain, mlp, p, u0, (0.0, 1.0), kw)


def test_headline_shape_interpolating():
    rng, mlp, p, chain = make([64, 256, 256, 64], 3, scale=2.5)
    u0 = rng.standard_normal((200, 64)).astype(np.float32)
    kw = dict(abstol=1e-5, reltol=1e-5, saveat=[0.25, 0.5, 1.0])
    node_case(chain, mlp, p, u0, (0.0, 1.0), kw, dL_fn=lambda us: np.sin(us).astype(np.float32), adj_exact=False)


def test_augmented_node():
    """AugmentedNDELayer (src/neural_de.jl:372-386; test/BasicNeuralDE/neural_de_tests.jl:50-55)."""
    rng, mlp, p, chain = make([4, 50, 4], 4)
    u0 = np.array([[0, 0], [1, 0], [2, 0]], np.float32)
    sol, fw, adj = node_case(chain, mlp, p, u0, (0.0, 1.0), dict(saveat=0.1), augment=2, noisy=True)
    assert sol.array().shape == (11, 3, 4) and adj.du0.shape == (3, 2)


def test_reverse_time_span():
    """tspan (1, 0): the direction FFJORD sampling uses (src/ffjord.jl:205)."""
    rng, mlp, p, chain = make([8, 32, 8], 5, scale=2.0)
    u0 = rng.standard_normal((64, 8)).astype(np.float32)
    node_case(chain, mlp, p, u0, (1.0, 0.0), dict(abstol=1e-5, reltol=1e-5, saveat=[0.75, 0.25]))


def test_torch_autograd_device_pointers():
    """CUDA tensors travel as device pointers; loss.backward() runs the adjoint kernel."""
    import torch
    rng, mlp, p, chain = make([8, 32, 8], 5, scale=2.0)
    u0 = rng.standard_normal((64, 8)).astype(np.float32)
    ts = np.arange(0, 1.01, 0.25, dtype=np.float32)
    node = dfx.NeuralODE(chain, (0.0, 1.0), dfx.Tsit5(), saveat=ts, abstol=1e-5, reltol=1e-5)
    pt = torch.tensor(p, device="cuda", requires_grad=True)
    ut = torch.tensor(u0, device="cuda", requires_grad=True)
    sol, _ = node(ut, pt, {})
    loss = dfx.Array(sol).sum()
    loss.backward()
    prob = onode.NodeProblem(mlp)
    fw = onode.forward(prob, p, u0, 0.0, 1.0, saveat=ts, dense=True, abstol=1e-5, reltol=1e-5)
    adj = onode.adjoint(prob, p, fw, np.ones_like(fw.us), abstol=1e-5, reltol=1e-5)
    close(pt.grad.cpu().numpy(), adj.dp, what="dp")
    close(ut.grad.cpu().numpy(), adj.du0, what="du0")
    assert abs(loss.item() - float(fw.us.sum())) <= 1e-4 * abs(float(fw.us.sum()))


def test_retcodes():
    """maxiters -> MaxIters (1); the `Inf` gate of multiple_shoot keys on it (src/multiple_shooting.jl:66-67)."""
    rng, mlp, p, chain = make([8, 32, 8], 5, scale=2.0)
    u0 = rng.standard_normal((16, 8)).astype(np.float32)
    node = dfx.NeuralODE(chain, (0.0, 1.0), dfx.Tsit5(), abstol=1e-6, reltol=1e-6, maxiters=3, save_everystep=False)
    sol, _ = node(u0, p, {})
    fw = onode.forward(onode.NodeProblem(mlp), p, u0, 0.0, 1.0, abstol=1e-6, reltol=1e-6, maxiters=3, save_everystep=False)
    assert sol.retcode == dfx.ReturnCode.MaxIters == fw.stats.retcode
    bad = p.copy()
    bad[0] = np.nan
    sol, _ = dfx.NeuralODE(chain, (0.0, 1.0), dfx.Tsit5(), save_everystep=False)(u0, bad, {})
    assert sol.retcode == dfx.ReturnCode.Unstable


def test_ffjord_forward_and_adjoint():
    """FFJORD monte_carlo branch (src/ffjord.jl:81-136,140-178) with the shared Philox noise stream."""
    rng, mlp, p, chain = make([2, 16, 16, 2], 6, scale=0.5)
    B = 200
    x = rng.standard_normal((B, 2)).astype(np.float32)
    ff = dfx.FFJORD(chain, (0.0, 0.5), (2,), dfx.Tsit5(), abstol=1e-3, reltol=1e-3)
    st = {"regularize": False, "monte_carlo": True, "seed": 3, "calls": 0}
    logpx, back = ff.pullback(x, p, st)
    prob = onode.FfjordProblem(mlp, seed=3)
    u0 = np.concatenate([x, np.zeros((B, 1), np.float32)], 1)
    fw = onode.forward(prob, p, u0, 0.0, 0.5, save_everystep=False, save_start=False, dense=True, abstol=1e-3, reltol=1e-3)
    check_steps(ff.last["stats"], fw.stats, "forward", True)
    close(logpx, onode.ffjord_logpx(fw.us[0], 2), what="logpx")
    g = np.full(B, -1.0 / B, np.float32)                       # loss = -mean(logpx)
    dx, dp = back(g)
    adj = onode.adjoint(prob, p, fw, onode.ffjord_logpx_grad(fw.us[0], 2, g)[None], abstol=1e-3, reltol=1e-3)
    check_steps(ff.last["adjoint_stats"], adj.stats, "adjoint", False)
    close(dp, adj.dp, rtol=1e-3, what="dp")
    close(dx, adj.du0[:, :2], rtol=1e-3, what="dx")


def test_ffjord_layer_call_protocol():
    rng, mlp, p, chain = make([2, 16, 16, 2], 6, scale=0.5)
    x = rng.standard_normal((50, 2)).astype(np.float32)
    ff = dfx.FFJORD(chain, (0.0, 0.3), (2,), dfx.Tsit5(), abstol=1e-2, reltol=1e-2)
    ps, st = dfx.setup(np.random.default_rng(0), ff)
    (logpx, l1, l2), st2 = ff(x, p, st)
    assert logpx.shape == (1, 50) and st2["calls"] == 1 and st2["monte_carlo"] is True
    (logpx2, _, _), _ = ff(x, p, st2)
    assert not np.array_equal(logpx, logpx2)                   # fresh noise on every call


def test_multiple_shoot():
    """multiple_shoot (src/multiple_shooting.jl:36-84): grouped solves, loss and dp."""
    import torch
    rng, mlp, teacher, chain = make([4, 16, 4], 7, scale=2.0)
    prob_o = onode.NodeProblem(mlp)
    ts = np.linspace(0, 2, 9, dtype=np.float32)
    u0 = rng.standard_normal((40, 4)).astype(np.float32)
    tol = dict(abstol=1e-5, reltol=1e-5)
    data = onode.forward(prob_o, teacher, u0, 0.0, 2.0, saveat=ts, **tol).us          # (T, B, D)
    p = (teacher + 0.05 * rng.standard_normal(teacher.size)).astype(np.float32)
    want_loss, want_preds, want_dp, _ = omshoot.multiple_shoot(prob_o, p, data, ts, 3, **tol)
    pt = torch.tensor(p, requires_grad=True)
    dt = torch.tensor(data)
    prob = dfx.ODEProblem(chain, u0, (0.0, 2.0), p)
    loss, preds = dfx.multiple_shoot(pt, dt, ts, prob, lambda u, uh: ((u - uh) ** 2).sum(), dfx.Tsit5(), 3, **tol)
    assert len(preds) == 4 and tuple(preds[0].shape) == (3, 40, 4)
    for a, b in zip(preds, want_preds):
        close(a.detach().numpy(), b, what="preds")
    assert abs(loss.item() - want_loss) <= 1e-4 * abs(want_loss)
    loss.backward()
    close(pt.grad.numpy(), want_dp, rtol=2e-4, what="dp")
    # failed group solve => Inf (src/multiple_shooting.jl:66-67)
    loss_inf, _ = dfx.multiple_shoot(p, data, ts, prob, lambda u, uh: ((u - uh) ** 2).sum(), dfx.Tsit5(), 3, maxiters=2, **tol)
    assert loss_inf == float("inf")
    with pytest.raises(ValueError):
        dfx.multiple_shoot(p, data, ts, prob, lambda u, uh: 0.0, dfx.Tsit5(), 1)


@pytest.mark.parametrize("B", [200, 1000])
def test_tensor_core_engine_forward(B):
    """tcgen05 3xTF32 RHS engine (opts.path = 2) against the oracle and against the FP32-FFMA kernels:
    same accepted-step sequence, states within rtol 1e-4 (observed ~5e-6)."""
    rng, mlp, p, chain = make([64, 256, 256, 64], 3, scale=2.5)
    u0 = rng.standard_normal((B, 64)).astype(np.float32)
    kw = dict(abstol=1e-6, reltol=1e-6, saveat=[0.5])
    sols = {}
    for path in (1, 2):
        node = dfx.NeuralODE(chain, (0.0, 1.0), dfx.Tsit5(), path=path, **kw)
        sols[path], _ = node(u0, p, {})
        assert sols[path].stats.engine == path - 1 and sols[path].retcode == 0
    fw = onode.forward(onode.NodeProblem(mlp), p, u0, 0.0, 1.0, **kw)
    check_steps(sols[1].stats, fw.stats, "forward FFMA", True)
    # 3xTF32 rounding noise is ~4x fp32's: a step whose proposed dt lands within ~1e-3 of the end of the
    # span can flip, so the tensor-core count is held to +-1 of the oracle's
    assert abs(sols[2].stats.naccept - fw.stats.naccept) <= 1 and sols[2].stats.nreject == fw.stats.nreject
    for path in (1, 2):
        close(sols[path].array(), fw.us, what=f"states path {path}")
    close(sols[2].array(), sols[1].array(), rtol=2e-5, what="tensor-core vs FFMA")


def test_tensor_core_engine_rejects_ineligible_shapes():
    rng, mlp, p, chain = make([2, 50, 2], 1)
    node = dfx.NeuralODE(chain, (0.0, 1.0), dfx.Tsit5(), path=2, save_everystep=False)
    with pytest.raises(dfx.B200Error):
        node(np.zeros((4, 2), np.float32), p, {})


@pytest.mark.parametrize("sense", [1, 0])
def test_tensor_core_engine_adjoint(sense):
    """tcgen05 adjoint (forward recompute + dgrad + MN-major wgrad GEMMs, opts.path = 2) against the oracle
    and the FP32-FFMA kernels at the headline shape: states and gradients element-wise within rtol 1e-4.
    The oracle accumulates the parameter gradient in the CUDA path's order (per 128-trajectory tile, tiles in fp64).
    Step counts: forward exact on FFMA and within 1 on the tensor cores; the adjoint's first ~15 steps ramp dt up
    from Hairer's tiny first step (mu(t1) = 0 against abstol 1e-6) with an error estimate that is pure rounding
    noise, so its count follows the arithmetic's noise floor: FFMA within 1 of the oracle,
    tensor cores within 8 %."""
    rng, mlp, p, chain = make([64, 256, 256, 64], 3, scale=2.5)
    mlp.wgrad_accum = ("tile_defer", 128)
    B = 200                                                    # ragged: 1 full + 1 partial 128-row tile
    u0 = rng.standard_normal((B, 64)).astype(np.float32)
    sa = dfx.BacksolveAdjoint() if sense else dfx.InterpolatingAdjoint()
    kw = dict(abstol=1e-6, reltol=1e-6, sensealg=sa)
    kw.update(dict(save_everystep=False, save_start=False) if sense else dict(saveat=[0.5, 1.0]))
    prob = onode.NodeProblem(mlp)
    okw = {k: v for k, v in kw.items() if k != "sensealg"}
    fw = onode.forward(prob, p, u0, 0.0, 1.0, dense=(sense == 0), **okw)
    dL = np.cos(np.arange(fw.us.size, dtype=np.float32).reshape(fw.us.shape) * 0.01)
    adj = onode.adjoint(prob, p, fw, dL, sensealg=sense, abstol=1e-6, reltol=1e-6)
    out = {}
    for path in (1, 2):
        node = dfx.NeuralODE(chain, (0.0, 1.0), dfx.Tsit5(), path=path, **kw)
        sol, back = node.pullback(u0, p)
        du0, dp = back(dL)
        ast = node.last["adjoint_stats"]
        assert sol.stats.engine == path - 1 and ast.engine == path - 1
        assert abs(sol.stats.naccept - fw.stats.naccept) <= (0 if path == 1 else 1), "forward steps"
        assert sol.stats.nreject == fw.stats.nreject and ast.nreject == adj.stats.nreject
        lim = 1 if path == 1 else max(2, -(-8 * adj.stats.naccept // 100))
        assert abs(ast.naccept - adj.stats.naccept) <= lim, f"adjoint steps path {path}: {counts(ast)} vs oracle {counts(adj.stats)}"
        assert ast.nf == 3 + 6 * (ast.naccept + ast.nreject) + (0 if sense else 1)      # FSAL + one re-evaluation per interior jump
        close(sol.array(), fw.us, what=f"states path {path}")
        close(dp, adj.dp, what=f"dp path {path}")
        close(du0, adj.du0, what=f"du0 path {path}")
        out[path] = (du0, dp)
    close(out[2][1], out[1][1], rtol=1e-4, what="dp tensor-core vs FFMA")


@pytest.mark.parametrize("sense", [0, 1])
def test_augmented_node_tight_tolerance(sense):
    """AugmentedNDELayer (src/neural_de.jl:372-386) at abstol = reltol = 1e-6 with interior loss jumps: forward step
    count exact, states / dp / d(input) element-wise within rtol 1e-4 (observed ~2e-6); no gradient reaches the
    zero-padded rows.  (The adjoint count is not compared: see test_tensor_core_engine_adjoint.)"""
    rng, mlp, p, chain = make([8, 64, 8], 5, scale=3.0)
    u0 = rng.standard_normal((256, 6)).astype(np.float32)
    sa = dfx.BacksolveAdjoint() if sense else dfx.InterpolatingAdjoint()
    kw = dict(abstol=1e-6, reltol=1e-6, saveat=[0.25, 0.5, 1.0], sensealg=sa)
    okw = {k: v for k, v in kw.items() if k != "sensealg"}
    prob = onode.NodeProblem(mlp)
    fw = onode.forward(prob, p, u0, 0.0, 1.0, dense=(sense == 0), augment=2, **okw)
    dL = np.sin(np.arange(fw.us.size, dtype=np.float32).reshape(fw.us.shape) * 0.37)
    adj = onode.adjoint(prob, p, fw, dL, sensealg=sense, augment=2, abstol=1e-6, reltol=1e-6)
    layer = dfx.AugmentedNDELayer(dfx.NeuralODE(chain, (0.0, 1.0), dfx.Tsit5(), **kw), 2)
    sol, back = layer.pullback(u0, p)
    du0, dp = back(dL)
    check_steps(sol.stats, fw.stats, "forward", True)
    assert fw.stats.tie_dist > 0.5                     # truncation-dominated: no accept/reject decision near a tie
    assert sol.array().shape == (3, 256, 8) and du0.shape == (256, 6)
    close(sol.array(), fw.us, what="states")
    close(dp, adj.dp, what="dp")
    close(du0, adj.du0, what="du0")


def test_ffjord_gradients_large_batch():
    """FFJORD (src/ffjord.jl:81-178, monte-carlo branch) at B = 4096 with fresh Hutchinson noise per RHS call: step
    counts of both passes exact (rejected steps included), logpx and d(x) element-wise within 1e-4 and the parameter
    gradient within 1e-4 (it is a sum over 4096 trajectories of second-order terms; observed 4e-5)."""
    rng = np.random.default_rng(6)
    mlp = MLP([2, 16, 16, 2], [ACT_TANH, ACT_TANH, ACT_ID])
    p = mlp.init(rng, scale=0.5)
    B = 4096
    x = rng.standard_normal((B, 2)).astype(np.float32)
    tol = dict(abstol=1e-3, reltol=1e-3)
    prob = onode.FfjordProblem(mlp, seed=3)
    fw = onode.forward(prob, p, np.concatenate([x, np.zeros((B, 1), np.float32)], 1), 0.0, 0.5, save_everystep=False,
                       save_start=False, dense=True, **tol)
    g = np.full(B, -1.0 / B, np.float32)
    adj = onode.adjoint(prob, p, fw, onode.ffjord_logpx_grad(fw.us[0], 2, g)[None], **tol)
    chain = dfx.Chain(dfx.Dense(2, 16, dfx.tanh), dfx.Dense(16, 16, dfx.tanh), dfx.Dense(16, 2))
    ff = dfx.FFJORD(chain, (0.0, 0.5), (2,), dfx.Tsit5(), **tol)
    logpx, back = ff.pullback(x, p, {"regularize": False, "monte_carlo": True, "seed": 3, "calls": 0})
    dx, dp = back(g)
    check_steps(ff.last["stats"], fw.stats, "ffjord forward", True)
    check_steps(ff.last["adjoint_stats"], adj.stats, "ffjord adjoint", True)
    assert fw.stats.nreject > 0
    close(np.asarray(logpx).reshape(-1), onode.ffjord_logpx(fw.us[0], 2), what="logpx")
    close(dx, adj.du0[:, :2], what="dx")
    close(dp, adj.dp, what="dp")


def test_multiple_shoot_data_gradient():
    """b200_mshoot_adjoint's d_data blocks ([G][B][D], one per group, accumulated at time index first[g]): with a loss
    that reads only the predictions and no continuity term, d(loss)/d(ode_data) is exactly the groups' initial-
    condition cotangents, which the NeuralODE pullback of each group returns independently."""
    import torch
    rng, mlp, p, chain = make([4, 16, 4], 7, scale=2.0)
    ts = np.linspace(0, 2, 9, dtype=np.float32)
    data = rng.standard_normal((9, 24, 4)).astype(np.float32)
    tol = dict(abstol=1e-6, reltol=1e-6)
    pt = torch.tensor(p, requires_grad=True)
    dt = torch.tensor(data, requires_grad=True)
    prob = dfx.ODEProblem(chain, data[0], (0.0, 2.0), p)
    loss, preds = dfx.multiple_shoot(pt, dt, ts, prob, lambda u, uh: (uh ** 2).sum(), dfx.Tsit5(), 3, continuity_term=0, **tol)
    loss.backward()
    want = np.zeros_like(data)
    dp_want = np.zeros_like(p)
    for (a, b), pr in zip(dfx.group_ranges(9, 3), preds):
        node = dfx.NeuralODE(chain, (float(ts[a]), float(ts[b])), dfx.Tsit5(), saveat=ts[a:b + 1], **tol)
        sol, back = node.pullback(data[a], p)
        close(sol.array(), pr.detach().numpy(), rtol=1e-6, what="group prediction")
        du0, dpg = back(2 * sol.array())
        want[a] += du0
        dp_want += dpg
    close(dt.grad.numpy(), want, rtol=1e-5, what="d(ode_data)")
    close(pt.grad.numpy(), dp_want, rtol=1e-5, what="dp")
    assert np.abs(want[[0, 2, 4, 6]]).max() > 0 and np.abs(want[[1, 3, 5, 7, 8]]).max() == 0


def test_wide_model_picks_a_tile_that_fits_beside_the_static_shared_memory():
    """the tile planner must budget the kernel's static __shared__ block (21 KB) beside the dynamic
    part.  128 -> 768 -> 768 -> 128 needs 229 KB of dynamic shared memory for 32-row tiles, which passes a naive
    "opt-in minus 2 KB" check and then fails at launch; the planner now falls back to 16-row tiles."""
    rng, mlp, p, chain = make([128, 768, 768, 128], 9, scale=1.0)
    u0 = rng.standard_normal((48, 128)).astype(np.float32)
    kw = dict(abstol=1e-4, reltol=1e-4, save_everystep=False, save_start=False, sensealg=dfx.BacksolveAdjoint())
    node = dfx.NeuralODE(chain, (0.0, 0.5), dfx.Tsit5(), **kw)
    sol, back = node.pullback(u0, p)
    du0, dp = back(np.ones_like(sol.array()))
    assert sol.retcode == 0 and node.last["adjoint_stats"].tile in (8, 16)
    prob = onode.NodeProblem(mlp)
    fw = onode.forward(prob, p, u0, 0.0, 0.5, abstol=1e-4, reltol=1e-4, save_everystep=False, save_start=False)
    adj = onode.adjoint(prob, p, fw, np.ones_like(fw.us), sensealg=1, abstol=1e-4, reltol=1e-4)
    close(sol.array(), fw.us, rtol=3e-4, what="states")
    close(dp, adj.dp, rtol=3e-4, what="dp")


def test_c_abi_argument_errors_and_tape_ownership():
    """save_everystep without a tape is an argument error (it used to return success with untouched outputs); a
    tape outlives its context safely; the scalar saveat of a reversed time span is t0:-h:t1."""
    import ctypes as C
    from diffeqflux_jl_b200 import _native as N
    from diffeqflux_jl_b200 import layers
    lib = N.lib()
    ctx = N.Context(0)
    dims, acts = (C.c_int * 3)(2, 8, 2), (C.c_int * 2)(N.ACT_TANH, N.ACT_IDENTITY)
    mh = C.c_void_p()
    assert lib.b200_model_create(ctx._h, 2, dims, acts, 0, C.byref(mh)) == 0
    p = np.zeros(lib.b200_model_nparams(mh), np.float32)
    u0 = np.ones((3, 2), np.float32)
    o = N.default_opts()
    o.save_everystep = 1
    st = N.Stats()
    rc = lib.b200_node_forward(ctx._h, mh, p.ctypes.data, u0.ctypes.data, 2, 0, 3, 0.0, 1.0, None, 0, C.byref(o), None, None, C.byref(st), None)
    assert rc == -1 and b"tape" in lib.b200_last_error(ctx._h)
    tape = C.c_void_p()
    rc = lib.b200_node_forward(ctx._h, mh, p.ctypes.data, u0.ctypes.data, 2, 0, 3, 0.0, 1.0, None, 0, C.byref(o), None, None, C.byref(st), C.byref(tape))
    assert rc == 0 and tape.value and lib.b200_tape_nsteps(tape) == st.naccept
    other = N.Context(0)
    assert lib.b200_tape_steps(other._h, tape, None, None, 0) == -1          # a tape is bound to the context that recorded it
    ctx.close()                                                              # context first ...
    lib.b200_tape_free(tape)                                                 # ... then the tape: must not touch freed memory
    other.close()
    iv, ss, se, sev = layers._split_saveat(0.25, 1.0, 0.0, None, None, None)
    assert np.allclose(iv, [0.75, 0.5, 0.25]) and ss and se and not sev


def test_fused_training_step_matches_the_oracle_loop():
    """N2: b200_node_train_step (forward + sum-of-squares loss + adjoint + Adam in one call, parameters and optimiser
    state resident on the device) against the oracle's loop of docs/src/examples/neural_ode.md:36-70 for 4 iterations."""
    import torch
    from oracle import train as otrain
    rng, mlp, p0, chain = make([2, 50, 2], 2, pre=True)
    u0 = np.array([[2.0, 0.0]], np.float32)
    ts = np.linspace(0, 1.5, 30, dtype=np.float32)
    A = np.array([[-0.1, 2.0], [-2.0, -0.1]], np.float32)
    data = np.zeros((30, 1, 2), np.float32)                        # a crude spiral-like target (any fixed data does)
    for i, t in enumerate(ts):
        data[i, 0] = 2.0 * np.exp(-0.1 * t) * np.array([np.cos(2 * t), -np.sin(2 * t)], np.float32)
    prob = onode.NodeProblem(mlp)
    node = dfx.NeuralODE(chain, (0.0, 1.5), dfx.Tsit5(), saveat=ts)
    for device in ("host", "cuda"):
        p_o, m_o, v_o = p0.copy(), np.zeros_like(p0), np.zeros_like(p0)
        if device == "cuda":
            p_g, u_g, d_g = torch.tensor(p0).cuda(), torch.tensor(u0).cuda(), torch.tensor(data).cuda()
        else:
            p_g, u_g, d_g = p0.copy(), u0, data
        opt = dfx.AdamState(lr=0.05)
        solid = np.ones(p0.size, bool)
        for it in range(1, 5):
            loss_o, p_o, m_o, v_o, g_o = otrain.train_step(prob, p_o, u0, data, 0.0, 1.5, it, m_o, v_o, saveat=ts, lr=0.05)
            loss_g, rc = dfx.train_step(node, u_g, p_g, d_g, opt)
            assert rc == 0 and abs(loss_g - loss_o) <= 3e-3 * abs(loss_o), (device, it, loss_g, loss_o)
            pg = p_g.cpu().numpy() if device == "cuda" else p_g
            # Adam normalises the gradient (each parameter moves by ~lr per step whatever the size of its gradient), so
            # parameters whose gradient is numerically zero are ill-conditioned (lr * g / (|g| + 1e-8)); the others must
            # agree to a fraction of lr between two valid fp32 gradient evaluations (rtol 3e-3 at reltol 1e-3)
            solid &= np.abs(g_o) > 1e-3 * np.abs(g_o).max()
            assert solid.sum() > p0.size // 3
            assert np.abs(pg - p_o)[solid].max() <= 0.05 * 0.05 * it, (device, it, np.abs(pg - p_o)[solid].max())
        assert opt.step == 4 and np.abs(pg - p0).max() > 0.05


def test_dimmover_and_ffjord_sampling():
    """N3 / N4: DimMover (src/neural_de.jl:397-412) on the device against the oracle's eachslice/cat, and
    `__backward_ffjord` (src/ffjord.jl:183-214): the reverse-time flow of the forward flow's end points returns the data."""
    import torch
    from oracle import train as otrain
    rng = np.random.default_rng(0)
    x = rng.standard_normal((5, 7, 3)).astype(np.float32)            # C-order view of a Julia (3, 7, 5) = (D, B, T) array
    xj = np.transpose(x, (2, 1, 0))
    for frm, to in ((-2, -1), (1, 3), (3, 1), (2, 2), (-1, 1)):
        want = np.transpose(otrain.dimmover(xj, frm, to))           # back to the C-order view
        for arr in (x, torch.tensor(x).cuda()):
            got, _ = dfx.DimMover(frm, to)(arr, None, {})
            got = got.cpu().numpy() if hasattr(got, "cpu") else got
            assert got.shape == want.shape and np.array_equal(got, want), (frm, to, got.shape, want.shape)
    y, _ = dfx.DimMover()(x, None, {})
    assert y.shape == (7, 5, 3) and np.array_equal(y, np.transpose(x, (1, 0, 2)))
    # sampling: z = flow(x) forward, then flow^-1(z) through backward() must return x (noise does not enter the z rows)
    rng2, mlp, p, chain = make([2, 16, 16, 2], 6, scale=0.5)
    xs = rng2.standard_normal((512, 2)).astype(np.float32)
    ff = dfx.FFJORD(chain, (0.0, 0.5), (2,), dfx.Tsit5(), abstol=1e-6, reltol=1e-6, noise_per_call=0)
    st = {"regularize": False, "monte_carlo": True, "seed": 3, "calls": 0}
    ff(xs, p, st)
    z = ff.last["z"][:, :2].copy()
    back = ff.backward(512, p, st, z=z)
    close(back, xs, rtol=1e-4, what="flow^-1(flow(x))")
    d = dfx.FFJORDDistribution(ff, p, st)
    assert len(d) == 2 and d.rand(64, np.random.default_rng(1)).shape == (64, 2) and d.logpdf(xs).shape == (512,)


@pytest.mark.parametrize("sa", ["quadrature", "gauss"])
def test_quadrature_adjoint(sa):
    """N1: QuadratureAdjoint / GaussAdjoint (src/DiffEqFlux.jl:19-20; test/AdvancedNeuralDE/second_order_ode_tests.jl:59,83).
    The oracle restates the reference: lambda alone integrated against the dense forward output, dp by adaptive
    Gauss-Kronrod (QuadGK 7/15) between the saved times.  The CUDA path integrates lambda alone as well (mu is not in the
    error norm: fewer steps than InterpolatingAdjoint on the same problem) and forms dp with the Runge-Kutta weights of
    those steps: du0 and dp element-wise within rtol 1e-4."""
    rng, mlp, p, chain = make([16, 64, 16], 3, scale=3.0)
    u0 = rng.standard_normal((64, 16)).astype(np.float32)
    kw = dict(abstol=1e-6, reltol=1e-6, saveat=[0.3, 0.7, 1.0])
    prob = onode.NodeProblem(mlp)
    fw = onode.forward(prob, p, u0, 0.0, 1.0, dense=True, **kw)
    dL = np.sin(fw.us).astype(np.float32)
    adj = onode.adjoint(prob, p, fw, dL, sensealg=onode.SENSE_QUADRATURE, abstol=1e-6, reltol=1e-6, quad_abstol=1e-8, quad_reltol=1e-6)
    ref_i = onode.adjoint(prob, p, fw, dL, sensealg=onode.SENSE_INTERPOLATING, abstol=1e-6, reltol=1e-6)
    sens = dfx.QuadratureAdjoint(abstol=1e-8, reltol=1e-6) if sa == "quadrature" else dfx.GaussAdjoint()
    node = dfx.NeuralODE(chain, (0.0, 1.0), dfx.Tsit5(), sensealg=sens, **kw)
    sol, back = node.pullback(u0, p)
    du0, dp = back(dL)
    check_steps(sol.stats, fw.stats, "forward", True)
    ast = node.last["adjoint_stats"]
    # lambda's own step sequence: within one step of the oracle's (two interior tstops: whether the step before a stop
    # lands a few 1e-4 short of it decides about one extra remainder step), no rejections
    assert abs(ast.naccept - adj.stats.naccept) <= 1 and ast.nreject == adj.stats.nreject == 0, (counts(ast), counts(adj.stats))
    assert ast.nf == 3 + 6 * ast.naccept + 2
    assert adj.stats.naccept < ref_i.stats.naccept          # without mu's rounding noise in the norm the ramp-up is shorter
    close(du0, adj.du0, what="du0")
    close(dp, adj.dp, what="dp")
    close(dp, ref_i.dp, what="dp vs InterpolatingAdjoint")


def test_cfg5_shape_augmented_192_512_512_192():
    """BASELINE.json configs[4] at its own widths (state 128 + 64 zero-augment rows, MLP 192 -> 512 -> 512 -> 192; FP32-FFMA
    engine: the widths exceed the tcgen05 engine's 256) on a small batch: forward counts exact, states and gradients
    element-wise within 3 x reltol at the docs' tolerance 1e-3 (docs/src/examples/augmented_neural_ode.md:47-50) and
    within 1e-4 at 1e-6."""
    rng, mlp, p, chain = make([192, 512, 512, 192], 5, scale=1.0)
    u0 = rng.standard_normal((96, 128)).astype(np.float32)
    prob = onode.NodeProblem(mlp)
    for tol, rtol in ((1e-3, 3e-3), (1e-6, 1e-4)):
        kw = dict(abstol=tol, reltol=tol, save_everystep=False, save_start=False)
        fw = onode.forward(prob, p, u0, 0.0, 1.0, dense=True, augment=64, **kw)
        adj = onode.adjoint(prob, p, fw, np.ones_like(fw.us), sensealg=0, augment=64, abstol=tol, reltol=tol)
        layer = dfx.AugmentedNDELayer(dfx.NeuralODE(chain, (0.0, 1.0), dfx.Tsit5(), **kw), 64)
        sol, back = layer.pullback(u0, p)
        du0, dp = back(np.ones_like(sol.array()))
        assert sol.stats.engine == 0 and sol.array().shape == (1, 96, 192) and du0.shape == (96, 128)
        check_steps(sol.stats, fw.stats, f"forward tol {tol}", True)
        close(sol.array(), fw.us, rtol=rtol, what=f"states tol {tol}")
        close(dp, adj.dp, rtol=rtol, what=f"dp tol {tol}")
        close(du0, adj.du0, rtol=rtol, what=f"du0 tol {tol}")


def test_multiple_shoot_ensemble():
    """N4: the EnsembleProblem method of multiple_shoot (src/multiple_shooting.jl:131-195): every trajectory of every group
    is an independent solve (own step sequence); loss, predictions and dp against the oracle's restatement."""
    import torch
    rng, mlp, teacher, chain = make([4, 16, 4], 7, scale=2.0)
    prob_o = onode.NodeProblem(mlp)
    ts = np.linspace(0, 2, 9, dtype=np.float32)
    ntraj = 3
    u0 = rng.standard_normal((ntraj, 4)).astype(np.float32)
    tol = dict(abstol=1e-6, reltol=1e-6)
    data = np.stack([onode.forward(prob_o, teacher, u0[i], 0.0, 2.0, saveat=ts, **tol).us[:, 0, :] for i in range(ntraj)])   # (ntraj, T, D)
    p = (teacher + 0.05 * rng.standard_normal(teacher.size)).astype(np.float32)
    want_loss, want_preds, want_dp = omshoot.multiple_shoot_ensemble(prob_o, p, data, ts, 3, **tol)
    pt = torch.tensor(p, requires_grad=True)
    eprob = dfx.EnsembleProblem(dfx.ODEProblem(chain, u0[0], (0.0, 2.0), p))
    loss, preds = dfx.multiple_shoot(pt, torch.tensor(data), ts, eprob, dfx.EnsembleThreads(), lambda u, uh: ((u - uh) ** 2).sum(),
                                     dfx.Tsit5(), 3, trajectories=ntraj, **tol)
    assert len(preds) == 4 and tuple(preds[0].shape) == (ntraj, 3, 4)
    for a, b in zip(preds, want_preds):
        close(a.detach().numpy(), b, what="ensemble preds")
    assert abs(loss.item() - want_loss) <= 1e-4 * abs(want_loss)
    loss.backward()
    close(pt.grad.numpy(), want_dp, rtol=2e-4, what="ensemble dp")
    with pytest.raises(AssertionError):
        dfx.multiple_shoot(p, data, ts, eprob, dfx.EnsembleThreads(), lambda u, uh: 0.0, dfx.Tsit5(), 3, trajectories=ntraj + 1)


def test_ffjord_exact_trace():
    """N4: FFJORD with monte_carlo = false (src/ffjord.jl:120-123, test/AdvancedNeuralDE/cnf_tests.jl:34-38): the exact
    trace of the Jacobian by D unit-vector passes.  Deterministic RHS: step counts of both passes exact, logpx / dx / dp
    element-wise within 1e-4; for a linear model logpx has the closed form logN(x e^{tW}) + t tr(W)."""
    rng = np.random.default_rng(8)
    mlp = MLP([3, 16, 16, 3], [ACT_TANH, ACT_TANH, ACT_ID])
    p = mlp.init(rng, scale=1.0)
    B = 300
    x = rng.standard_normal((B, 3)).astype(np.float32)
    tol = dict(abstol=1e-6, reltol=1e-6)
    prob = onode.FfjordProblem(mlp, seed=0, monte_carlo=False)
    fw = onode.forward(prob, p, np.concatenate([x, np.zeros((B, 1), np.float32)], 1), 0.0, 0.5, save_everystep=False,
                       save_start=False, dense=True, **tol)
    g = np.cos(np.arange(B, dtype=np.float32))
    adj = onode.adjoint(prob, p, fw, onode.ffjord_logpx_grad(fw.us[0], 3, g)[None], **tol)
    chain = dfx.Chain(dfx.Dense(3, 16, dfx.tanh), dfx.Dense(16, 16, dfx.tanh), dfx.Dense(16, 3))
    ff = dfx.FFJORD(chain, (0.0, 0.5), (3,), dfx.Tsit5(), **tol)
    st = {"regularize": False, "monte_carlo": False, "seed": 0, "calls": 0}
    logpx, back = ff.pullback(x, p, st)
    dx, dp = back(g)
    check_steps(ff.last["stats"], fw.stats, "exact-trace forward", True)
    assert abs(ff.last["adjoint_stats"].naccept - adj.stats.naccept) <= 1
    close(np.asarray(logpx).reshape(-1), onode.ffjord_logpx(fw.us[0], 3), what="logpx")
    close(dx, adj.du0[:, :3], what="dx")
    close(dp, adj.dp, what="dp")
    (lp2, _, _), _ = ff(x, p, st)                                  # the layer call protocol takes the same branch
    assert np.array_equal(np.asarray(lp2).reshape(-1), np.asarray(logpx).reshape(-1))


@pytest.mark.parametrize("monte_carlo", [True, False])
def test_ffjord_regularize(monte_carlo):
    """N4: FFJORD with regularize = true (src/ffjord.jl:131-133,149-153,164-166; test/AdvancedNeuralDE/cnf_tests.jl:34-38):
    the state carries sum(abs2, m(z)) and ||e^T J||_2 integrals.  Forward state, logpx and the gradient of a loss that uses
    logpx AND both regulariser rows (cotangent on the whole end state) against the oracle, with one noise draw per solve
    (noise_per_call = 0: a smooth RHS, so tolerance 1e-6 and exact step counts; with the reference's fresh noise per RHS call
    the ||e^T J|| row is a random process and the controller takes ~400 noise-limited steps).  Then the layer call with fresh
    noise: (logpx, lambda_1, lambda_2) sliced as the reference slices them, autograd through all three."""
    import torch
    rng = np.random.default_rng(21)
    D = 3
    mlp = MLP([D, 16, 16, D], [ACT_TANH, ACT_TANH, ACT_ID])
    p = mlp.init(rng, scale=1.0)
    B = 300
    x = rng.standard_normal((B, D)).astype(np.float32)
    tol = dict(abstol=1e-6, reltol=1e-6)
    prob = onode.FfjordProblem(mlp, seed=5, monte_carlo=monte_carlo, regularize=True, noise_per_call=False)
    fw = onode.forward(prob, p, np.concatenate([x, np.zeros((B, 3), np.float32)], 1), 0.0, 0.5, save_everystep=False,
                       save_start=False, dense=True, **tol)
    g = np.cos(np.arange(B, dtype=np.float32))
    gz = np.zeros((B, D + 3), np.float32)
    gz[:, D + 1] = 0.3 + 0.1 * np.sin(np.arange(B))                # weights of the two regularisers in the loss
    gz[:, D + 2] = 0.7 - 0.2 * np.cos(np.arange(B))
    dL = onode.ffjord_logpx_grad(fw.us[0], D, g)
    dL[:, D + 1:] = 0
    adj = onode.adjoint(prob, p, fw, (dL + gz)[None], **tol)
    chain = dfx.Chain(dfx.Dense(D, 16, dfx.tanh), dfx.Dense(16, 16, dfx.tanh), dfx.Dense(16, D))
    ff = dfx.FFJORD(chain, (0.0, 0.5), (D,), dfx.Tsit5(), noise_per_call=0, **tol)
    st = {"regularize": True, "monte_carlo": monte_carlo, "seed": 5, "calls": 0}
    logpx, back = ff.pullback(x, p, st)
    dx, dp = back(g, gz)
    z = np.asarray(ff.last["z"])
    assert z.shape == (B, D + 3)
    check_steps(ff.last["stats"], fw.stats, "regularize forward", True)
    # (the backward solve spends most of its ~12 steps ramping up from Hairer's tiny initial dt, where the error estimate
    # is rounding noise -- DESIGN.md "step-count parity": the count follows the arithmetic's noise floor; gradients below
    # agree to 1e-6)
    assert abs(ff.last["adjoint_stats"].naccept - adj.stats.naccept) <= 3
    close(z, fw.us[0], what="end state")
    assert np.all(z[:, D + 1] > 0) and np.all(z[:, D + 2] > 0)
    close(np.asarray(logpx).reshape(-1), onode.ffjord_logpx(fw.us[0], D), what="logpx")
    close(dx, adj.du0[:, :D], what="dx")
    close(dp, adj.dp, what="dp")
    # layer protocol with the reference's fresh noise + autograd through (logpx, lambda_1, lambda_2)
    ff2 = dfx.FFJORD(chain, (0.0, 0.2), (D,), dfx.Tsit5(), abstol=1e-2, reltol=1e-2)
    xt = torch.from_numpy(x).cuda().requires_grad_(True)
    pt = torch.from_numpy(p).cuda().requires_grad_(True)
    (lp, l1, l2), st2 = ff2(xt, pt, st)
    assert lp.shape == (1, B) and l1.shape == (1, D + 3) and l2.shape == (1, D + 3) and st2["calls"] == 1
    zt = ff2.last["z"]
    assert torch.equal(l1.detach(), zt[D + 1:D + 2]) and torch.equal(l2.detach(), zt[D + 2:D + 3])
    loss = -(lp.mean()) + 0.5 * l1.sum() + 0.25 * l2.sum()
    loss.backward()
    _, back2 = ff2.pullback(x, p, st)                              # same seed, same solve
    gz2 = np.zeros((B, D + 3), np.float32)
    gz2[D + 1, :] = 0.5
    gz2[D + 2, :] = 0.25
    dx2, dp2 = back2(np.full(B, -1.0 / B, np.float32), gz2)
    close(pt.grad.cpu().numpy(), np.asarray(dp2), rtol=1e-5, what="autograd dp")
    close(xt.grad.cpu().numpy(), np.asarray(dx2), rtol=1e-5, what="autograd dx")
    smp = ff2.backward(64, p, st, rng=np.random.default_rng(0))    # sampling takes the same branch (src/ffjord.jl:194-213)
    assert np.asarray(smp).shape == (64, D) and np.isfinite(np.asarray(smp)).all()


def test_ffjord_many_rows_uses_full_tiles():
    """FFJORD with a narrow network and >= 128 trajectories per SM: the planner gives forward and adjoint 128-row tiles (a
    tile's GEMM is (TB/8) x (width/4) register tiles over 256 threads: smaller tiles leave threads idle) instead of the
    row-per-thread RHS / 64-row tiles.  Same noise stream (it is indexed by the global trajectory), so the oracle
    comparison of test_ffjord_gradients_large_batch carries over: exact step counts, logpx / dx / dp within 1e-4."""
    import torch
    nsm = torch.cuda.get_device_properties(0).multi_processor_count
    rng = np.random.default_rng(16)
    mlp = MLP([2, 16, 16, 2], [ACT_TANH, ACT_TANH, ACT_ID])
    p = mlp.init(rng, scale=0.5)
    B = 128 * nsm + 37
    x = rng.standard_normal((B, 2)).astype(np.float32)
    tol = dict(abstol=1e-3, reltol=1e-3)
    prob = onode.FfjordProblem(mlp, seed=9)
    fw = onode.forward(prob, p, np.concatenate([x, np.zeros((B, 1), np.float32)], 1), 0.0, 0.5, save_everystep=False,
                       save_start=False, dense=True, **tol)
    g = np.full(B, -1.0 / B, np.float32)
    adj = onode.adjoint(prob, p, fw, onode.ffjord_logpx_grad(fw.us[0], 2, g)[None], **tol)
    chain = dfx.Chain(dfx.Dense(2, 16, dfx.tanh), dfx.Dense(16, 16, dfx.tanh), dfx.Dense(16, 2))
    ff = dfx.FFJORD(chain, (0.0, 0.5), (2,), dfx.Tsit5(), **tol)
    logpx, back = ff.pullback(x, p, {"regularize": False, "monte_carlo": True, "seed": 9, "calls": 0})
    dx, dp = back(g)
    assert ff.last["stats"].tile == 128 and ff.last["adjoint_stats"].tile == 128
    check_steps(ff.last["stats"], fw.stats, "ffjord forward", True)
    check_steps(ff.last["adjoint_stats"], adj.stats, "ffjord adjoint", True)
    close(np.asarray(logpx).reshape(-1), onode.ffjord_logpx(fw.us[0], 2), what="logpx")
    close(dx, adj.du0[:, :2], what="dx")
    close(dp, adj.dp, what="dp")
