"""Probe: step counts / errors of the CUDA path vs the oracle for the cases the parity tests tighten (run on the GPU box)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import diffeqflux_b200 as dfx
from oracle import node as onode
from oracle.mlp import ACT_ID, ACT_TANH, MLP

def make(dims, seed, scale=1.0, **mk):
    rng = np.random.default_rng(seed)
    acts = [ACT_TANH] * (len(dims) - 2) + [ACT_ID]
    mlp = MLP(dims, acts, **mk)
    p = mlp.init(rng, scale=scale)
    layers = [dfx.Dense(i, o, dfx.tanh if a == ACT_TANH else dfx.identity) for i, o, a in zip(dims[:-1], dims[1:], acts)]
    return rng, mlp, p, dfx.Chain(*layers)

def relerr(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    s = np.abs(b).max()
    return float((np.abs(a - b) / (np.abs(b) + s)).max()), float(np.abs(a - b).max() / s)

def cnt(s): return (s.naccept, s.nreject, s.nf)

# 1. TC engine adjoint test shapes
for sense in (1, 0):
    for B in (200, 1000):
        rng, mlp, p, chain = make([64, 256, 256, 64], 3, scale=2.5, wgrad_accum=("tile_defer", 128))
        u0 = rng.standard_normal((B, 64)).astype(np.float32)
        sa = dfx.BacksolveAdjoint() if sense else dfx.InterpolatingAdjoint()
        kw = dict(abstol=1e-6, reltol=1e-6, sensealg=sa)
        kw.update(dict(save_everystep=False, save_start=False) if sense else dict(saveat=[0.5, 1.0]))
        okw = {k: v for k, v in kw.items() if k != "sensealg"}
        prob = onode.NodeProblem(mlp)
        fw = onode.forward(prob, p, u0, 0.0, 1.0, dense=(sense == 0), **okw)
        dL = np.cos(np.arange(fw.us.size, dtype=np.float32).reshape(fw.us.shape) * 0.01)
        adj = onode.adjoint(prob, p, fw, dL, sensealg=sense, abstol=1e-6, reltol=1e-6)
        mlp_b = MLP(mlp.dims, mlp.acts)
        adj_b = onode.adjoint(onode.NodeProblem(mlp_b), p, fw, dL, sensealg=sense, abstol=1e-6, reltol=1e-6)
        line = f"TCadj sense {sense} B {B}: oracle fwd {cnt(fw.stats)} adj defer {cnt(adj.stats)} blas {cnt(adj_b.stats)} tie {adj.stats.tie_dist:.3f}"
        for path in (1, 2):
            node = dfx.NeuralODE(chain, (0.0, 1.0), dfx.Tsit5(), path=path, **kw)
            sol, back = node.pullback(u0, p)
            du0, dp = back(dL)
            a = node.last["adjoint_stats"]
            line += f" | path {path} fwd {cnt(sol.stats)} adj {cnt(a)} dp {relerr(dp, adj.dp)[1]:.1e}/{relerr(dp, adj.dp)[0]:.1e} du0 {relerr(du0, adj.du0)[1]:.1e} u {relerr(sol.array(), fw.us)[1]:.1e}"
        print(line, flush=True)

# 2. golden batched
g = np.load(os.path.join(ROOT, "tests/golden/batched.npz"))
dims = [int(x) for x in g["dims"]]
chain = dfx.Chain(*[dfx.Dense(i, o, dfx.tanh if k < len(dims) - 2 else dfx.identity) for k, (i, o) in enumerate(zip(dims[:-1], dims[1:]))])
for path in (1, 2):
    node = dfx.NeuralODE(chain, (0.0, 1.0), dfx.Tsit5(), abstol=1e-6, reltol=1e-6, save_everystep=False, save_start=False, sensealg=dfx.BacksolveAdjoint(), path=path)
    sol, back = node.pullback(g["u0"], g["p"])
    du0, dp = back(np.ones_like(sol.array()))
    print("golden batched path", path, "fwd", cnt(sol.stats), "golden", g["fwd_stats"][:3], "adj", cnt(node.last["adjoint_stats"]), "golden", g["adj_stats"][:3], "dp", relerr(dp, g["dp"]), flush=True)

# 3. augmented at 1e-6, truncation dominated
rng, mlp, p, chain = make([8, 64, 8], 5, scale=3.0)
u0 = rng.standard_normal((256, 6)).astype(np.float32)
for sense in (0, 1):
    sa = dfx.BacksolveAdjoint() if sense else dfx.InterpolatingAdjoint()
    kw = dict(abstol=1e-6, reltol=1e-6, saveat=[0.25, 0.5, 1.0], sensealg=sa)
    okw = {k: v for k, v in kw.items() if k != "sensealg"}
    prob = onode.NodeProblem(mlp)
    fw = onode.forward(prob, p, u0, 0.0, 1.0, dense=(sense == 0), augment=2, **okw)
    dL = np.sin(np.arange(fw.us.size, dtype=np.float32).reshape(fw.us.shape) * 0.37)
    adj = onode.adjoint(prob, p, fw, dL, sensealg=sense, augment=2, abstol=1e-6, reltol=1e-6)
    node = dfx.NeuralODE(chain, (0.0, 1.0), dfx.Tsit5(), **kw)
    layer = dfx.AugmentedNDELayer(node, 2)
    sol, back = layer.pullback(u0, p)
    du0, dp = back(dL)
    print("augmented sense", sense, "oracle", cnt(fw.stats), cnt(adj.stats), "tie", fw.stats.tie_dist, adj.stats.tie_dist, "gpu", cnt(sol.stats), cnt(layer.last["adjoint_stats"]),
          "u", relerr(sol.array(), fw.us), "dp", relerr(dp, adj.dp), "du0", relerr(du0, adj.du0), flush=True)

# 4. FFJORD B = 4096
rng = np.random.default_rng(6)
mlp = MLP([2, 16, 16, 2], [ACT_TANH, ACT_TANH, ACT_ID])
p = mlp.init(rng, scale=0.5)
B = 4096
x = rng.standard_normal((B, 2)).astype(np.float32)
for tol in (1e-3, 1e-6):
    prob = onode.FfjordProblem(mlp, seed=3)
    u0f = np.concatenate([x, np.zeros((B, 1), np.float32)], 1)
    fw = onode.forward(prob, p, u0f, 0.0, 0.5, save_everystep=False, save_start=False, dense=True, abstol=tol, reltol=tol)
    gg = np.full(B, -1.0 / B, np.float32)
    adj = onode.adjoint(prob, p, fw, onode.ffjord_logpx_grad(fw.us[0], 2, gg)[None], abstol=tol, reltol=tol)
    chain = dfx.Chain(dfx.Dense(2, 16, dfx.tanh), dfx.Dense(16, 16, dfx.tanh), dfx.Dense(16, 2))
    ff = dfx.FFJORD(chain, (0.0, 0.5), (2,), dfx.Tsit5(), abstol=tol, reltol=tol)
    st = {"regularize": False, "monte_carlo": True, "seed": 3, "calls": 0}
    try:
        (logpx, back) = ff.pullback(x, p, st)[:2]
        dx, dp = back(gg)
        print("ffjord tol", tol, "oracle", cnt(fw.stats), cnt(adj.stats), "gpu", cnt(ff.last["stats"]), cnt(ff.last["adjoint_stats"]), "logpx", relerr(np.asarray(logpx).reshape(-1), onode.ffjord_logpx(fw.us[0], 2)),
              "dp", relerr(dp, adj.dp), "dx", relerr(dx, adj.du0[:, :2]), flush=True)
    except Exception as e:
        print("ffjord probe failed:", repr(e))
