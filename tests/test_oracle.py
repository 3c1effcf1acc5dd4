"""CPU tests pinning the oracle against independent truths (scipy fp64, torch autograd, finite
differences, Random123 known-answer vectors).  The reference itself cannot be executed here
(no Julia), so these are the strongest available anchors: "parity unpinned" w.r.t. Julia."""
import math

import numpy as np
import pytest

from oracle import mshoot, node, philox, tsit5
from oracle.mlp import ACT_ID, ACT_TANH, MLP


def test_tableau_order_conditions():
    c = [0.0] + [float(x) for x in tsit5.C]
    b = [float(tsit5.A[(7, j)]) for j in range(1, 7)] + [0.0]
    a = np.zeros((7, 7))
    for (i, j), v in tsit5.A.items():
        a[i - 1, j - 1] = float(v)
    assert abs(sum(b) - 1) < 1e-6
    for i in range(1, 7):
        assert abs(a[i].sum() - c[i]) < 1e-6
    cb = np.array(c)
    bb = np.array(b)
    assert abs(bb @ cb - 1 / 2) < 1e-6
    assert abs(bb @ cb**2 - 1 / 3) < 1e-6
    assert abs(bb @ (a @ cb) - 1 / 6) < 1e-6
    assert abs(bb @ cb**3 - 1 / 4) < 1e-6
    assert abs(bb @ cb**4 - 1 / 5) < 1e-6
    assert abs(sum(float(x) for x in tsit5.BTILDE)) < 1e-6
    # dense output reproduces b at theta = 1 and 0 at theta = 0
    w1 = tsit5.interp_weights(1.0)
    for j in range(7):
        assert abs(float(w1[j]) - b[j]) < 2e-6
    assert all(float(x) == 0 for x in tsit5.interp_weights(0.0))


def test_fastpow_accuracy():
    for x in (1e-4, 0.03, 0.5, 0.99, 1.0, 1.7, 30.0):
        for y in (7 / 50, 2 / 25):
            assert abs(float(tsit5.fastpow(x, y)) - x**y) / x**y < 2e-4
    assert float(tsit5.fastpow(0.0, 0.14)) == 0.0


def test_philox_known_answers():
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
        ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
        ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
         (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
    ]
    for ctr, key, want in kat:
        got = philox.philox4x32_10(*[np.uint32(c) for c in ctr], key[0], key[1])
        assert tuple(int(g) for g in got) == want


def test_philox_normal_moments_and_sharding():
    e = philox.normal(3, 0, 7, 4096, 6)
    assert e.shape == (4096, 6) and e.dtype == np.float32
    assert abs(e.mean()) < 0.03 and abs(e.std() - 1) < 0.03
    part = philox.normal(3, 0, 7, 1000, 6, b_offset=2000)
    np.testing.assert_array_equal(part, e[2000:3000])


def _spiral():
    rng = np.random.default_rng(1)
    mlp = MLP([2, 50, 2], [ACT_TANH, ACT_ID])
    return mlp, mlp.init(rng)


def test_forward_matches_scipy_fp64():
    from scipy.integrate import solve_ivp
    mlp, p = _spiral()
    prob = node.NodeProblem(mlp)
    u0 = np.array([[2.0, 0.0], [1.0, -1.0], [0.0, 0.5]], np.float32)
    m64 = MLP(mlp.dims, mlp.acts, dtype=np.float64)
    p64 = p.astype(np.float64)
    errs = []
    for tol in (1e-3, 1e-5):
        fw = node.forward(prob, p, u0, 0.0, 1.5, abstol=tol, reltol=tol, saveat=[0.5, 1.0, 1.5],
                          save_start=False)
        assert fw.stats.retcode == 0
        assert fw.stats.nf == 3 + 6 * (fw.stats.naccept + fw.stats.nreject)
        ref = solve_ivp(lambda t, y: m64(p64, y.reshape(3, 2)).reshape(-1), (0, 1.5),
                        u0.reshape(-1).astype(np.float64), method="DOP853", rtol=1e-12, atol=1e-12,
                        t_eval=[0.5, 1.0, 1.5])
        want = ref.y.T.reshape(3, 3, 2)
        errs.append(np.abs(fw.us - want).max())
    assert errs[0] < 5e-3 and errs[1] < 5e-5 and errs[1] < errs[0]


def test_saveat_conventions():
    mlp, p = _spiral()
    prob = node.NodeProblem(mlp)
    u0 = np.array([2.0, 0.0], np.float32)
    fw = node.forward(prob, p, u0, 0.0, 1.0, saveat=0.1)            # scalar: 0:0.1:1
    assert fw.ts.shape == (11,) and fw.ts[0] == 0 and fw.ts[-1] == 1
    fw = node.forward(prob, p, u0, 0.0, 1.0, save_everystep=False, save_start=False)
    assert fw.ts.tolist() == [1.0] and fw.us.shape == (1, 1, 2)
    fw = node.forward(prob, p, u0, 0.0, 1.0)                        # every step
    assert fw.ts.shape[0] == fw.stats.naccept + 1
    fw2 = node.forward(prob, p, u0, 0.0, 1.0, saveat=[0.25, 0.5])   # ends not in saveat
    assert fw2.ts.tolist() == [0.25, 0.5]
    # saveat does not change the step sequence
    assert (fw2.stats.naccept, fw2.stats.nreject) == (fw.stats.naccept, fw.stats.nreject)


def _loss_fp64(mlp, p64, u0, ts, t1):
    from scipy.integrate import solve_ivp
    m64 = MLP(mlp.dims, mlp.acts, pre=mlp.pre, dtype=np.float64)
    B, D = u0.shape
    sol = solve_ivp(lambda t, y: m64(p64, y.reshape(B, D)).reshape(-1), (0, t1),
                    u0.reshape(-1).astype(np.float64), method="DOP853", rtol=1e-11, atol=1e-11,
                    t_eval=ts)
    return sol.y.sum()


@pytest.mark.parametrize("sense", [node.SENSE_INTERPOLATING, node.SENSE_BACKSOLVE])
def test_adjoint_matches_finite_differences(sense):
    rng = np.random.default_rng(5)
    mlp = MLP([2, 6, 2], [ACT_TANH, ACT_ID])
    p = mlp.init(rng)
    prob = node.NodeProblem(mlp)
    u0 = rng.standard_normal((3, 2)).astype(np.float32)
    ts = np.array([0.0, 0.4, 0.8, 1.0], np.float32)
    fw = node.forward(prob, p, u0, 0.0, 1.0, saveat=ts, abstol=1e-7, reltol=1e-6,
                      dense=(sense == node.SENSE_INTERPOLATING))
    adj = node.adjoint(prob, p, fw, np.ones_like(fw.us), sensealg=sense, abstol=1e-7, reltol=1e-6)
    assert adj.stats.retcode == 0
    # nf = 3 (init) + 6/step + one FSAL re-evaluation per interior jump
    assert adj.stats.nf == 3 + 6 * (adj.stats.naccept + adj.stats.nreject) + 2
    p64 = p.astype(np.float64)
    h = 1e-5
    gp = np.zeros_like(p64)
    for i in range(p64.size):
        e = np.zeros_like(p64)
        e[i] = h
        gp[i] = (_loss_fp64(mlp, p64 + e, u0, ts, 1.0) - _loss_fp64(mlp, p64 - e, u0, ts, 1.0)) / (2 * h)
    assert np.abs(adj.dp - gp).max() <= 2e-4 * np.abs(gp).max()
    gu = np.zeros(u0.shape)
    for i in range(u0.size):
        e = np.zeros(u0.size)
        e[i] = h
        e = e.reshape(u0.shape)
        gu.flat[i] = (_loss_fp64(mlp, p64, u0 + e, ts, 1.0) - _loss_fp64(mlp, p64, u0 - e, ts, 1.0)) / (2 * h)
    assert np.abs(adj.du0 - gu).max() <= 2e-4 * np.abs(gu).max()


def test_interpolating_and_backsolve_agree():
    mlp, p = _spiral()
    prob = node.NodeProblem(mlp)
    u0 = np.array([[2.0, 0.0]], np.float32)
    ts = np.linspace(0, 1.5, 30, dtype=np.float32)
    fw = node.forward(prob, p, u0, 0.0, 1.5, saveat=ts, dense=True)
    a0 = node.adjoint(prob, p, fw, np.ones_like(fw.us), sensealg=node.SENSE_INTERPOLATING)
    a1 = node.adjoint(prob, p, fw, np.ones_like(fw.us), sensealg=node.SENSE_BACKSOLVE)
    assert np.abs(a0.dp - a1.dp).max() < 1e-3 * np.abs(a0.dp).max()
    assert np.abs(a0.du0 - a1.du0).max() < 1e-3 * np.abs(a0.du0).max()


def test_linear_ode_against_expm():
    from scipy.linalg import expm
    mlp = MLP([3, 3], [ACT_ID])
    rng = np.random.default_rng(2)
    p = mlp.init(rng)
    p[9:] = 0
    Wn = p[:9].reshape(3, 3).astype(np.float64)     # y' = y @ Wn
    u0 = rng.standard_normal((4, 3)).astype(np.float32)
    fw = node.forward(node.NodeProblem(mlp), p, u0, 0.0, 2.0, abstol=1e-7, reltol=1e-6,
                      save_everystep=False, save_start=False)
    want = u0.astype(np.float64) @ expm(2.0 * Wn)
    assert np.abs(fw.us[0] - want).max() < 2e-5 * max(1, np.abs(want).max())


def test_cube_prelayer_gradient():
    import torch
    from oracle.mlp import PRE_CUBE
    rng = np.random.default_rng(3)
    mlp = MLP([2, 5, 2], [ACT_TANH, ACT_ID], pre=PRE_CUBE, dtype=np.float64)
    p = mlp.init(rng).astype(np.float64)
    x = rng.standard_normal((4, 2))
    yb = rng.standard_normal((4, 2))
    out, cache = mlp.forward(p, x)
    xbar, pbar = mlp.vjp(cache, yb)
    pt = torch.tensor(p, requires_grad=True)
    xt = torch.tensor(x, requires_grad=True)
    W1, b1 = pt[:10].reshape(2, 5), pt[10:15]
    W2, b2 = pt[15:25].reshape(5, 2), pt[25:27]
    o = torch.tanh((xt**3) @ W1 + b1) @ W2 + b2
    (o * torch.tensor(yb)).sum().backward()
    np.testing.assert_allclose(out, o.detach().numpy(), rtol=1e-12)
    np.testing.assert_allclose(xbar, xt.grad.numpy(), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(pbar, pt.grad.numpy(), rtol=1e-10, atol=1e-12)


def test_ffjord_second_order_vjp_against_torch():
    import torch
    rng = np.random.default_rng(4)
    mlp = MLP([2, 7, 7, 2], [ACT_TANH, ACT_TANH, ACT_ID], dtype=np.float64)
    p = mlp.init(rng).astype(np.float64)
    z = rng.standard_normal((5, 2))
    e = rng.standard_normal((5, 2))
    lz = rng.standard_normal((5, 2))
    ll = rng.standard_normal((5, 1))
    hL, hdL, cache = mlp.jvp_forward(p, z, e)
    zbar, pbar = mlp.second_order_vjp(cache, lz, -ll * e)
    pt = torch.tensor(p, requires_grad=True)
    zt = torch.tensor(z, requires_grad=True)
    et = torch.tensor(e)

    def model(zz):
        h = zz
        o = 0
        for i, n, a in zip(mlp.dims[:-1], mlp.dims[1:], mlp.acts):
            W = pt[o:o + i * n].reshape(i, n)
            b = pt[o + i * n:o + i * n + n]
            o += i * n + n
            h = h @ W + b
            if a == ACT_TANH:
                h = torch.tanh(h)
        return h

    mz = model(zt)
    eJ, = torch.autograd.grad(mz, zt, et, create_graph=True)       # the reference's VJP form
    tr = (eJ * et).sum(dim=1)
    np.testing.assert_allclose(hL, mz.detach().numpy(), rtol=1e-12)
    np.testing.assert_allclose((hdL * e).sum(1), tr.detach().numpy(), rtol=1e-10, atol=1e-12)
    phi = (torch.tensor(lz) * mz).sum() - (torch.tensor(ll[:, 0]) * tr).sum()
    gz, gp = torch.autograd.grad(phi, (zt, pt))
    np.testing.assert_allclose(zbar, gz.numpy(), rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(pbar, gp.numpy(), rtol=1e-9, atol=1e-11)
    # the forward RHS uses the VJP form; both forms give the same trace
    prob = node.FfjordProblem(MLP(mlp.dims, mlp.acts), seed=11)
    y = np.concatenate([z, np.zeros((5, 1))], 1).astype(np.float32)
    out, _ = prob.f(p.astype(np.float32), y, 3, 0)
    e32 = prob.noise(5, 3, 0)
    _, hd32, _ = prob.mlp.jvp_forward(p.astype(np.float32), z.astype(np.float32), e32)
    np.testing.assert_allclose(out[:, 2], -(hd32 * e32).sum(1), rtol=2e-4, atol=1e-5)


def test_ffjord_forward_and_adjoint_run():
    rng = np.random.default_rng(6)
    mlp = MLP([2, 8, 8, 2], [ACT_TANH, ACT_TANH, ACT_ID])
    p = mlp.init(rng, scale=0.5)
    prob = node.FfjordProblem(mlp, seed=3)
    x = rng.standard_normal((16, 2)).astype(np.float32)
    u0 = np.concatenate([x, np.zeros((16, 1), np.float32)], 1)
    fw = node.forward(prob, p, u0, 0.0, 0.5, save_everystep=False, save_start=False, dense=True,
                      abstol=1e-3, reltol=1e-3)
    assert fw.stats.retcode == 0 and fw.us.shape == (1, 16, 3)
    logpx = node.ffjord_logpx(fw.us[0], 2)
    assert np.all(np.isfinite(logpx))
    g = np.full(16, -1 / 16, np.float32)
    adj = node.adjoint(prob, p, fw, node.ffjord_logpx_grad(fw.us[0], 2, g)[None], abstol=1e-3, reltol=1e-3)
    assert adj.stats.retcode == 0 and np.isfinite(adj.dp).all() and np.abs(adj.dp).max() > 0


def test_group_ranges_and_mshoot():
    assert mshoot.group_ranges(10, 5) == [(0, 5), (4, 9), (8, 10)]     # Julia docstring: 1:5, 5:9, 9:10
    with pytest.raises(ValueError):
        mshoot.group_ranges(10, 1)
    with pytest.raises(ValueError):
        mshoot.group_ranges(4, 5)
    rng = np.random.default_rng(7)
    mlp = MLP([2, 8, 2], [ACT_TANH, ACT_ID])
    teacher = mlp.init(rng)
    prob = node.NodeProblem(mlp)
    ts = np.linspace(0, 2, 9, dtype=np.float32)
    u0 = rng.standard_normal((4, 2)).astype(np.float32)
    data = node.forward(prob, teacher, u0, 0.0, 2.0, saveat=ts).us            # (T, B, D)
    loss0, preds, dp0, _ = mshoot.multiple_shoot(prob, teacher, data, ts, 3)
    assert len(preds) == 4 and preds[0].shape == (3, 4, 2)
    assert loss0 < 1e-3                                   # teacher reproduces its own data
    p = teacher + 0.05 * rng.standard_normal(teacher.size).astype(np.float32)
    loss, _, dp, _ = mshoot.multiple_shoot(prob, p, data, ts, 3)
    assert loss > loss0
    # directional derivative check of dp
    d = rng.standard_normal(p.size).astype(np.float32)
    h = 1e-3
    lp, *_ = mshoot.multiple_shoot(prob, p + h * d, data, ts, 3, want_grad=False)
    lm, *_ = mshoot.multiple_shoot(prob, p - h * d, data, ts, 3, want_grad=False)
    fd = (lp - lm) / (2 * h)
    assert abs(float(dp @ d) - fd) < 2e-2 * max(1.0, abs(fd))


def test_training_step_and_dimmover_restatement():
    """oracle/train.py: Adam's first step moves every parameter by lr * sign(g) (bias-corrected m / sqrt(v) = +-1), the
    loss decreases over a few iterations of the docs' loop, and DimMover reproduces Julia's eachslice/cat shapes."""
    from oracle import train as otrain
    g = np.array([0.5, -2.0, 1e-3], np.float32)
    p, m, v = otrain.adam_update(np.zeros(3, np.float32), np.zeros(3, np.float32), np.zeros(3, np.float32), g, 1, lr=0.05)
    np.testing.assert_allclose(p, -0.05 * np.sign(g), rtol=1e-4)
    x = np.arange(24, dtype=np.float32).reshape(2, 3, 4)                # Julia (2, 3, 4)
    assert otrain.dimmover(x).shape == (2, 4, 3) and otrain.dimmover(x, 3, 1).shape == (8, 3) and otrain.dimmover(x, 2, 2).shape == (2, 12)
    assert np.array_equal(otrain.dimmover(x)[:, :, 1], x[:, 1, :])
    rng = np.random.default_rng(2)
    mlp = MLP([2, 16, 2], [ACT_TANH, ACT_ID])
    p = mlp.init(rng)
    prob = node.NodeProblem(mlp)
    ts = np.linspace(0, 1.0, 6, dtype=np.float32)
    u0 = np.array([[1.0, 0.0]], np.float32)
    data = np.stack([np.stack([np.cos(ts), -np.sin(ts)], -1)], 1).astype(np.float32)
    m, v, losses = np.zeros_like(p), np.zeros_like(p), []
    for it in range(1, 6):
        loss, p, m, v, _ = otrain.train_step(prob, p, u0, data, 0.0, 1.0, it, m, v, saveat=ts, lr=0.05)
        losses.append(loss)
    assert losses[-1] < losses[0]


def test_quadrature_adjoint_agrees_with_interpolating():
    """The QuadratureAdjoint restatement (lambda-only dense solve + adaptive Gauss-Kronrod) and the InterpolatingAdjoint
    restatement (mu inside the ODE state) compute the same gradient; the 7/15 rule integrates a degree-20 polynomial
    exactly and the adaptive driver reaches the requested tolerance on a peaked integrand."""
    val, err = node.quadgk_vec(lambda t: np.array([t ** 20, 1.0]), 0.0, 2.0, 0.0, 1e-13)
    np.testing.assert_allclose(val, [2.0 ** 21 / 21, 2.0], rtol=1e-12)
    val, err = node.quadgk_vec(lambda t: np.array([1.0 / (1e-3 + (t - 0.3) ** 2)]), 0.0, 1.0, 1e-10, 1e-10)
    exact = (np.arctan(0.7 / np.sqrt(1e-3)) + np.arctan(0.3 / np.sqrt(1e-3))) / np.sqrt(1e-3)
    assert abs(val[0] - exact) <= 1e-8 * exact
    rng = np.random.default_rng(3)
    mlp = MLP([8, 32, 8], [ACT_TANH, ACT_ID])
    p = mlp.init(rng, scale=2.5)
    u0 = rng.standard_normal((16, 8)).astype(np.float32)
    prob = node.NodeProblem(mlp)
    kw = dict(abstol=1e-6, reltol=1e-6)
    fw = node.forward(prob, p, u0, 0.0, 1.0, saveat=[0.4, 1.0], dense=True, **kw)
    dL = np.cos(fw.us).astype(np.float32)
    ai = node.adjoint(prob, p, fw, dL, sensealg=node.SENSE_INTERPOLATING, **kw)
    aq = node.adjoint(prob, p, fw, dL, sensealg=node.SENSE_QUADRATURE, quad_abstol=1e-8, quad_reltol=1e-6, **kw)
    assert np.abs(aq.dp - ai.dp).max() <= 2e-5 * np.abs(ai.dp).max()
    assert np.abs(aq.du0 - ai.du0).max() <= 2e-5 * np.abs(ai.du0).max()
    assert aq.stats.nf == 3 + 6 * (aq.stats.naccept + aq.stats.nreject) + 1      # one FSAL re-evaluation at the interior jump


def test_ffjord_exact_trace_restatement():
    """monte_carlo = false: the RHS's last row is minus the exact trace of dm/dz (checked against a finite-difference
    Jacobian of the fp64 model), and the adjoint RHS is the derivative of <lam, f> (finite differences in z)."""
    rng = np.random.default_rng(9)
    mlp64 = MLP([3, 8, 3], [ACT_TANH, ACT_ID], dtype=np.float64)
    mlp = MLP([3, 8, 3], [ACT_TANH, ACT_ID])
    p = mlp.init(rng, scale=1.5)
    z = rng.standard_normal((5, 3)).astype(np.float32)
    prob = node.FfjordProblem(mlp, seed=1, monte_carlo=False)
    y = np.concatenate([z, np.zeros((5, 1), np.float32)], 1)
    out, _ = prob.f(p, y, 0, 0)
    eps = 1e-6
    for b in range(5):
        J = np.zeros((3, 3))
        for k in range(3):
            zp, zm = z[b:b + 1].astype(np.float64).copy(), z[b:b + 1].astype(np.float64).copy()
            zp[0, k] += eps; zm[0, k] -= eps
            J[:, k] = (mlp64(p.astype(np.float64), zp) - mlp64(p.astype(np.float64), zm))[0] / (2 * eps)
        assert abs(-out[b, 3] - np.trace(J)) <= 1e-4 * max(1.0, abs(np.trace(J)))
    lam = rng.standard_normal((5, 4)).astype(np.float32)
    _, ybar, pbar = prob.vjp(p, y, lam, 0, 1)
    prob64 = node.FfjordProblem(MLP([3, 8, 3], [ACT_TANH, ACT_ID], dtype=np.float64), seed=1, monte_carlo=False)

    def phi(zz):
        yy = np.concatenate([zz, np.zeros((5, 1))], 1)
        o, _ = prob.f(p, yy.astype(np.float32), 0, 0)
        return float(np.sum(lam.astype(np.float64) * o.astype(np.float64)))
    h = 1e-2
    for (b, k) in ((0, 0), (3, 2)):
        zp, zm = z.astype(np.float64).copy(), z.astype(np.float64).copy()
        zp[b, k] += h; zm[b, k] -= h
        fd = (phi(zp) - phi(zm)) / (2 * h)
        assert abs(fd - ybar[b, k]) <= 2e-2 * max(1.0, abs(fd)), (fd, ybar[b, k])
    assert pbar.shape == p.shape and np.abs(pbar).max() > 0


@pytest.mark.parametrize("monte_carlo", [True, False])
def test_ffjord_regularize_restatement(monte_carlo):
    """regularize = true (src/ffjord.jl:131-133): rows D+1, D+2 of the RHS are sum(abs2, m(z)) and ||e^T J||_2 (checked
    against an fp64 finite-difference Jacobian), and the adjoint RHS is the derivative of <lam, f> with all four cotangent
    blocks live (finite differences in z and in p, noise held fixed)."""
    rng = np.random.default_rng(19)
    dims, acts = [3, 8, 3], [ACT_TANH, ACT_ID]
    mlp, mlp64 = MLP(dims, acts), MLP(dims, acts, dtype=np.float64)
    p = mlp.init(rng, scale=1.5)
    B = 5
    z = rng.standard_normal((B, 3)).astype(np.float32)
    prob = node.FfjordProblem(mlp, seed=1, monte_carlo=monte_carlo, regularize=True)
    assert prob.ns == 6
    y = np.concatenate([z, np.zeros((B, 3), np.float32)], 1)
    out, _ = prob.f(p, y, 4, 0)
    assert out.shape == (B, 6)
    e = prob.noise(B, 4, 0).astype(np.float64)
    eps = 1e-6
    for b in range(B):
        J = np.zeros((3, 3))
        for k in range(3):
            zp, zm = z[b:b + 1].astype(np.float64).copy(), z[b:b + 1].astype(np.float64).copy()
            zp[0, k] += eps; zm[0, k] -= eps
            J[:, k] = (mlp64(p.astype(np.float64), zp) - mlp64(p.astype(np.float64), zm))[0] / (2 * eps)
        tr = np.trace(J) if not monte_carlo else e[b] @ J @ e[b]
        assert abs(-out[b, 3] - tr) <= 1e-4 * max(1.0, abs(tr))
        assert abs(out[b, 4] - np.sum(out[b, :3].astype(np.float64) ** 2)) <= 1e-5 * max(1.0, out[b, 4])
        assert abs(out[b, 5] - np.linalg.norm(e[b] @ J)) <= 1e-4 * max(1.0, out[b, 5])
    lam = rng.standard_normal((B, 6)).astype(np.float32)
    o2, ybar, pbar = prob.vjp(p, y, lam, 4, 0)
    assert np.array_equal(o2, out) and np.all(ybar[:, 3:] == 0)
    prob64 = node.FfjordProblem(mlp64, seed=1, monte_carlo=monte_carlo, regularize=True)

    def phi(zz, pp):
        yy = np.concatenate([zz, np.zeros((B, 3))], 1)
        o, _ = prob64.f(pp, yy, 4, 0)
        return float(np.sum(lam.astype(np.float64) * o))
    h = 1e-5
    z64, p64 = z.astype(np.float64), p.astype(np.float64)
    for (b, k) in ((0, 0), (3, 2), (4, 1)):
        zp, zm = z64.copy(), z64.copy()
        zp[b, k] += h; zm[b, k] -= h
        fd = (phi(zp, p64) - phi(zm, p64)) / (2 * h)
        assert abs(fd - ybar[b, k]) <= 2e-3 * max(1.0, abs(fd)), (fd, ybar[b, k])
    for i in (0, 7, 20, p.size - 1):
        pp, pm = p64.copy(), p64.copy()
        pp[i] += h; pm[i] -= h
        fd = (phi(z64, pp) - phi(z64, pm)) / (2 * h)
        assert abs(fd - pbar[i]) <= 2e-3 * max(1.0, abs(fd)), (i, fd, pbar[i])
