"""GPU parity tests: the CUDA path (through the C ABI / the layer API) against the oracle on the same
seeded inputs.  Tolerance: rtol 1e-4 (fp32) on states and gradients, identical accepted/rejected
step counts (BASELINE.json north_star)."""
import numpy as np
import pytest

import diffeqflux_b200 as dfx
from oracle import mshoot as omshoot
from oracle import node as onode
from oracle.mlp import ACT_ID, ACT_TANH, MLP, PRE_CUBE

pytestmark = pytest.mark.gpu

RTOL = 1e-4


def close(a, b, rtol=RTOL, what=""):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    scale = max(np.abs(b).max(), 1e-30)
    err = np.abs(a - b).max() / scale
    assert err <= rtol, f"{what}: max err {err:.3e} of scale {scale:.3e}"


def make(dims, acts_o, acts_d, seed, scale=1.0, pre=False):
    rng = np.random.default_rng(seed)
    mlp = MLP(dims, acts_o, pre=PRE_CUBE if pre else 0)
    p = mlp.init(rng, scale=scale)
    layers = [dfx.Dense(i, o, a) for i, o, a in zip(dims[:-1], dims[1:], acts_d)]
    chain = dfx.Chain(*([dfx.cube] if pre else []), *layers)
    return rng, mlp, p, chain


def check_steps(stats_gpu, stats_or, what):
    assert (stats_gpu.naccept, stats_gpu.nreject, stats_gpu.nf) == (stats_or.naccept, stats_or.nreject, stats_or.nf), \
        f"{what}: gpu {(stats_gpu.naccept, stats_gpu.nreject, stats_gpu.nf)} oracle {(stats_or.naccept, stats_or.nreject, stats_or.nf)} tie {stats_or.tie_dist:.2e}"


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
def test_spiral_node_forward_and_gradient(kw, vector):
    """The kwarg matrix of test/BasicNeuralDE/neural_de_tests.jl:27-42 with u0 a vector and a 2x3 matrix."""
    rng, mlp, p, chain = make([2, 50, 2], [ACT_TANH, ACT_ID], [dfx.tanh, dfx.identity], 1)
    u0 = np.array([2.0, 0.0], np.float32) if vector else np.array([[0, 0], [1, 0], [2, 0]], np.float32)
    node = dfx.NeuralODE(chain, (0.0, 1.0), dfx.Tsit5(), **kw)
    sol, back = node.pullback(u0, p)
    okw = {k: v for k, v in kw.items() if k != "sensealg"}
    sense = kw.get("sensealg", dfx.InterpolatingAdjoint()).code
    prob = onode.NodeProblem(mlp)
    fw = onode.forward(prob, p, u0, 0.0, 1.0, dense=(sense == 0), **okw)
    assert sol.retcode == 0
    check_steps(sol.stats, fw.stats, "forward")
    np.testing.assert_array_equal(np.asarray(sol.t), fw.ts)
    arr = sol.array() if not vector else sol.array()[:, None, :]
    close(arr, fw.us, what="states")
    dL = np.ones_like(arr)                                  # sum o first o node
    du0, dp = back(dL if not vector else dL[:, 0, :])
    adj = onode.adjoint(prob, p, fw, dL, sensealg=sense, abstol=kw.get("abstol", 1e-6), reltol=kw.get("reltol", 1e-3))
    check_steps(node.last["adjoint_stats"], adj.stats, "adjoint")
    close(dp, adj.dp, what="dp")
    close(np.atleast_2d(du0), adj.du0, what="du0")
    assert np.abs(dp).max() > 0 and np.abs(du0).max() > 0       # the reference's own assertion


def test_save_everystep_returns_the_step_sequence():
    rng, mlp, p, chain = make([2, 50, 2], [ACT_TANH, ACT_ID], [dfx.tanh, dfx.identity], 1)
    u0 = np.array([[2.0, 0.0], [1.0, 1.0]], np.float32)
    node = dfx.NeuralODE(chain, (0.0, 1.5), dfx.Tsit5())
    sol, back = node.pullback(u0, p)
    fw = onode.forward(onode.NodeProblem(mlp), p, u0, 0.0, 1.5, dense=True)
    assert len(sol) == fw.stats.naccept + 1
    close(np.asarray(sol.t), fw.ts, rtol=1e-6, what="step times")
    close(sol.array(), fw.us, what="step states")
    du0, dp = back(np.ones_like(sol.array()))
    adj = onode.adjoint(onode.NodeProblem(mlp), p, fw, np.ones_like(fw.us))
    close(dp, adj.dp, what="dp")


def test_cube_prelayer_docs_example():
    """docs/src/examples/neural_ode.md:32: Chain(x -> x.^3, Dense(2, 50, tanh), Dense(50, 2))."""
    rng, mlp, p, chain = make([2, 50, 2], [ACT_TANH, ACT_ID], [dfx.tanh, dfx.identity], 2, pre=True)
    u0 = np.array([2.0, 0.0], np.float32)
    ts = np.linspace(0, 1.5, 30, dtype=np.float32)
    node = dfx.NeuralODE(chain, (0.0, 1.5), dfx.Tsit5(), saveat=ts)
    sol, back = node.pullback(u0, p)
    prob = onode.NodeProblem(mlp)
    fw = onode.forward(prob, p, u0, 0.0, 1.5, saveat=ts, dense=True)
    check_steps(sol.stats, fw.stats, "forward")
    close(sol.array()[:, None, :], fw.us, what="states")
    dL = 2 * (fw.us - 1.0)
    du0, dp = back(dL[:, 0, :])
    adj = onode.adjoint(prob, p, fw, dL)
    close(dp, adj.dp, what="dp")
    close(du0[None], adj.du0, what="du0")


@pytest.mark.parametrize("sense", [0, 1])
@pytest.mark.parametrize("B", [100, 1000])
def test_batched_node_mid_size(sense, B):
    """ragged batch (not a multiple of the tile), several tiles per CTA, 3 Dense layers."""
    rng, mlp, p, chain = make([16, 64, 64, 16], [ACT_TANH, ACT_TANH, ACT_ID], [dfx.tanh, dfx.tanh, dfx.identity], 3, scale=0.7)
    u0 = rng.standard_normal((B, 16)).astype(np.float32)
    sa = dfx.InterpolatingAdjoint() if sense == 0 else dfx.BacksolveAdjoint()
    node = dfx.NeuralODE(chain, (0.0, 1.0), dfx.Tsit5(), abstol=1e-6, reltol=1e-6, save_everystep=False, save_start=False,
                         sensealg=sa, tile=32 if B == 1000 else 0)
    sol, back = node.pullback(u0, p)
    prob = onode.NodeProblem(mlp)
    fw = onode.forward(prob, p, u0, 0.0, 1.0, abstol=1e-6, reltol=1e-6, save_everystep=False, save_start=False,
                       dense=(sense == 0))
    check_steps(sol.stats, fw.stats, "forward")
    close(sol.array(), fw.us, what="end state")
    dL = rng.standard_normal(fw.us.shape).astype(np.float32)
    du0, dp = back(dL)
    adj = onode.adjoint(prob, p, fw, dL, sensealg=sense, abstol=1e-6, reltol=1e-6)
    check_steps(node.last["adjoint_stats"], adj.stats, "adjoint")
    close(dp, adj.dp, what="dp")
    close(du0, adj.du0, what="du0")


def test_headline_shape_small_batch():
    """BASELINE config 2 shape (64 -> 256 -> 256 -> 64 tanh, tol 1e-6, Backsolve) at a batch the oracle
    finishes in seconds."""
    rng, mlp, p, chain = make([64, 256, 256, 64], [ACT_TANH, ACT_TANH, ACT_ID], [dfx.tanh, dfx.tanh, dfx.identity], 2, scale=0.5)
    B = 512
    u0 = rng.standard_normal((B, 64)).astype(np.float32)
    node = dfx.NeuralODE(chain, (0.0, 1.0), dfx.Tsit5(), abstol=1e-6, reltol=1e-6, save_everystep=False, save_start=False,
                         sensealg=dfx.BacksolveAdjoint())
    sol, back = node.pullback(u0, p)
    prob = onode.NodeProblem(mlp)
    fw = onode.forward(prob, p, u0, 0.0, 1.0, abstol=1e-6, reltol=1e-6, save_everystep=False, save_start=False)
    check_steps(sol.stats, fw.stats, "forward")
    close(sol.array(), fw.us, what="end state")
    dL = np.ones_like(fw.us)
    du0, dp = back(dL)
    adj = onode.adjoint(prob, p, fw, dL, sensealg=1, abstol=1e-6, reltol=1e-6)
    check_steps(node.last["adjoint_stats"], adj.stats, "adjoint")
    close(dp, adj.dp, what="dp")
    close(du0, adj.du0, what="du0")


def test_augmented_node():
    """AugmentedNDELayer (src/neural_de.jl:372-386; test/BasicNeuralDE/neural_de_tests.jl:50-55)."""
    rng, mlp, p, chain = make([4, 50, 4], [ACT_TANH, ACT_ID], [dfx.tanh, dfx.identity], 4)
    u0 = np.array([[0, 0], [1, 0], [2, 0]], np.float32)
    anode = dfx.AugmentedNDELayer(dfx.NeuralODE(chain, (0.0, 1.0), dfx.Tsit5(), saveat=0.1), 2)
    sol, back = anode.pullback(u0, p)
    prob = onode.NodeProblem(mlp)
    fw = onode.forward(prob, p, u0, 0.0, 1.0, saveat=0.1, dense=True, augment=2)
    check_steps(sol.stats, fw.stats, "forward")
    assert sol.array().shape == (11, 3, 4)
    close(sol.array(), fw.us, what="states")
    du0, dp = back(np.ones_like(fw.us))
    adj = onode.adjoint(prob, p, fw, np.ones_like(fw.us), augment=2)
    assert du0.shape == (3, 2)
    close(du0, adj.du0, what="du0")
    close(dp, adj.dp, what="dp")


def test_torch_autograd_device_pointers():
    """CUDA tensors travel as device pointers; loss.backward() runs the adjoint kernel."""
    import torch
    rng, mlp, p, chain = make([2, 50, 2], [ACT_TANH, ACT_ID], [dfx.tanh, dfx.identity], 1)
    u0 = np.array([[0, 0], [1, 0], [2, 0]], np.float32)
    ts = np.arange(0, 1.01, 0.1, dtype=np.float32)
    node = dfx.NeuralODE(chain, (0.0, 1.0), dfx.Tsit5(), saveat=ts)
    pt = torch.tensor(p, device="cuda", requires_grad=True)
    ut = torch.tensor(u0, device="cuda", requires_grad=True)
    sol, _ = node(ut, pt, {})
    loss = dfx.Array(sol).sum()
    loss.backward()
    prob = onode.NodeProblem(mlp)
    fw = onode.forward(prob, p, u0, 0.0, 1.0, saveat=ts, dense=True)
    adj = onode.adjoint(prob, p, fw, np.ones_like(fw.us))
    close(pt.grad.cpu().numpy(), adj.dp, what="dp")
    close(ut.grad.cpu().numpy(), adj.du0, what="du0")
    assert abs(float(loss) - float(fw.us.sum())) <= 1e-4 * abs(float(fw.us.sum()))


def test_ffjord_forward_and_adjoint():
    """FFJORD monte_carlo branch (src/ffjord.jl:81-136,140-178) with the shared Philox noise stream."""
    rng, mlp, p, chain = make([2, 16, 16, 2], [ACT_TANH, ACT_TANH, ACT_ID], [dfx.tanh, dfx.tanh, dfx.identity], 6, scale=0.5)
    B = 200
    x = rng.standard_normal((B, 2)).astype(np.float32)
    ff = dfx.FFJORD(chain, (0.0, 0.5), (2,), dfx.Tsit5(), abstol=1e-3, reltol=1e-3)
    st = {"regularize": False, "monte_carlo": True, "seed": 3, "calls": 0}
    logpx, back = ff.pullback(x, p, st)
    prob = onode.FfjordProblem(mlp, seed=3)
    u0 = np.concatenate([x, np.zeros((B, 1), np.float32)], 1)
    fw = onode.forward(prob, p, u0, 0.0, 0.5, save_everystep=False, save_start=False, dense=True, abstol=1e-3, reltol=1e-3)
    check_steps(ff.last["stats"], fw.stats, "forward")
    want = onode.ffjord_logpx(fw.us[0], 2)
    close(logpx, want, what="logpx")
    g = np.full(B, -1.0 / B, np.float32)                       # loss = -mean(logpx)
    dx, dp = back(g)
    adj = onode.adjoint(prob, p, fw, onode.ffjord_logpx_grad(fw.us[0], 2, g)[None], abstol=1e-3, reltol=1e-3)
    check_steps(ff.last["adjoint_stats"], adj.stats, "adjoint")
    close(dp, adj.dp, rtol=5e-4, what="dp")
    close(dx, adj.du0[:, :2], rtol=5e-4, what="dx")


def test_multiple_shoot():
    """multiple_shoot (src/multiple_shooting.jl:36-84): grouped solves, loss and dp."""
    import torch
    rng, mlp, teacher, chain = make([4, 16, 4], [ACT_TANH, ACT_ID], [dfx.tanh, dfx.identity], 7)
    prob_o = onode.NodeProblem(mlp)
    ts = np.linspace(0, 2, 9, dtype=np.float32)
    u0 = rng.standard_normal((6, 4)).astype(np.float32)
    data = onode.forward(prob_o, teacher, u0, 0.0, 2.0, saveat=ts).us          # (T, B, D)
    p = (teacher + 0.05 * rng.standard_normal(teacher.size)).astype(np.float32)
    want_loss, want_preds, want_dp, _ = omshoot.multiple_shoot(prob_o, p, data, ts, 3)
    pt = torch.tensor(p, requires_grad=True)
    dt = torch.tensor(data)
    prob = dfx.ODEProblem(chain, u0, (0.0, 2.0), p)
    loss, preds = dfx.multiple_shoot(pt, dt, ts, prob, lambda u, uh: ((u - uh) ** 2).sum(), dfx.Tsit5(), 3)
    assert len(preds) == 4 and tuple(preds[0].shape) == (3, 6, 4)
    for a, b in zip(preds, want_preds):
        close(a.detach().numpy(), b, what="preds")
    assert abs(float(loss) - want_loss) <= 1e-4 * abs(want_loss)
    loss.backward()
    close(pt.grad.numpy(), want_dp, what="dp")
