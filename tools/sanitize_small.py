"""Small end-to-end cases for compute-sanitizer (memcheck): FFMA and tcgen05 NeuralODE forward+adjoint, FFJORD (row path, exact
trace), quadrature adjoint, train step, dimmover."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import diffeqflux_b200 as dfx
rng = np.random.default_rng(0)

def params(dims, scale=1.5):
    return np.concatenate([np.concatenate([(rng.standard_normal(i * o) * scale / np.sqrt(i)), rng.standard_normal(o) * 0.1]) for i, o in zip(dims[:-1], dims[1:])]).astype(np.float32)

def chain(dims):
    return dfx.Chain(*[dfx.Dense(i, o, dfx.tanh if k < len(dims) - 2 else dfx.identity) for k, (i, o) in enumerate(zip(dims[:-1], dims[1:]))])

for dims, B, path, sa in (([6, 20, 6], 37, 1, dfx.InterpolatingAdjoint()), ([32, 64, 32], 150, 2, dfx.BacksolveAdjoint()), ([32, 64, 32], 150, 2, dfx.QuadratureAdjoint())):
    p = params(dims); u0 = rng.standard_normal((B, dims[0])).astype(np.float32)
    node = dfx.NeuralODE(chain(dims), (0.0, 0.3), dfx.Tsit5(), abstol=1e-4, reltol=1e-4, saveat=[0.1, 0.3], sensealg=sa, path=path)
    sol, back = node.pullback(u0, p)
    du0, dp = back(np.ones_like(sol.array()))
    print("node", dims, path, type(sa).__name__, sol.stats.naccept, node.last["adjoint_stats"].naccept, float(np.abs(dp).sum()), flush=True)
dims = [2, 16, 2]
p = params(dims, 0.5); x = rng.standard_normal((300, 2)).astype(np.float32)
for mc in (True, False):
    ff = dfx.FFJORD(chain(dims), (0.0, 0.3), (2,), dfx.Tsit5(), abstol=1e-3, reltol=1e-3)
    lp, back = ff.pullback(x, p, {"regularize": False, "monte_carlo": mc, "seed": 1, "calls": 0})
    dx, dp = back(np.full(300, -1 / 300, np.float32))
    print("ffjord mc", mc, float(np.sum(lp)), float(np.abs(dp).sum()), flush=True)
node = dfx.NeuralODE(chain([2, 16, 2]), (0.0, 0.5), dfx.Tsit5(), saveat=np.linspace(0, 0.5, 4, dtype=np.float32))
pp = params([2, 16, 2]); opt = dfx.AdamState()
print("train", dfx.train_step(node, np.ones((3, 2), np.float32), pp, np.zeros((4, 3, 2), np.float32), opt))
print("dimmover", dfx.DimMover()(rng.standard_normal((4, 3, 2)).astype(np.float32))[0].shape)

# FFJORD regularize (both trace modes) and the many-rows tile path (two CTAs per SM)
dims = [3, 16, 16, 3]
p = params(dims, 1.0)
for mc in (True, False):
    x = rng.standard_normal((77, 3)).astype(np.float32)
    ff = dfx.FFJORD(chain(dims), (0.0, 0.3), (3,), dfx.Tsit5(), abstol=1e-3, reltol=1e-3, noise_per_call=0)
    lp, back = ff.pullback(x, p, {"regularize": True, "monte_carlo": mc, "seed": 1, "calls": 0})
    gz = np.zeros((77, 6), np.float32); gz[:, 4] = 0.5; gz[:, 5] = 0.25
    dx, dp = back(np.ones(77, np.float32), gz)
    print("ffjord regularize", mc, float(np.abs(dp).sum()), flush=True)
x = rng.standard_normal((128 * 148 + 5, 3)).astype(np.float32)
ff = dfx.FFJORD(chain(dims), (0.0, 0.2), (3,), dfx.Tsit5(), abstol=1e-2, reltol=1e-2)
lp, back = ff.pullback(x, p, {"regularize": False, "monte_carlo": True, "seed": 1, "calls": 0})
dx, dp = back(np.ones(x.shape[0], np.float32))
print("ffjord many rows", ff.last["stats"].tile, ff.last["stats"].grid, ff.last["adjoint_stats"].tile, float(np.abs(dp).sum()), flush=True)
