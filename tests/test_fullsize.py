"""BASELINE.json configs[1] at its full size (batch 16384, 64 -> 256 -> 256 -> 64 tanh, Tsit5 1e-6, BacksolveAdjoint),
where the oracle would need minutes: size-independent properties instead of a pointwise comparison.

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
