"""debug: FFJORD regularize adjoint blocks against the oracle."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import diffeqflux_b200 as dfx
from oracle import node as onode
from oracle.mlp import MLP, ACT_TANH, ACT_ID

rng = np.random.default_rng(21); D = 3
mlp = MLP([D, 16, 16, D], [ACT_TANH, ACT_TANH, ACT_ID]); p = mlp.init(rng, scale=1.0); B = 300
x = rng.standard_normal((B, D)).astype(np.float32)
tol = dict(abstol=1e-6, reltol=1e-6)
chain = dfx.Chain(dfx.Dense(D, 16, dfx.tanh), dfx.Dense(16, 16, dfx.tanh), dfx.Dense(16, D))
for mc in (True, False):
    prob = onode.FfjordProblem(mlp, seed=5, monte_carlo=mc, regularize=True, noise_per_call=False)
    fw = onode.forward(prob, p, np.concatenate([x, np.zeros((B, 3), np.float32)], 1), 0.0, 0.5, save_everystep=False, save_start=False, dense=True, **tol)
    ff = dfx.FFJORD(chain, (0.0, 0.5), (D,), dfx.Tsit5(), noise_per_call=0, **tol)
    st = {"regularize": True, "monte_carlo": mc, "seed": 5, "calls": 0}
    logpx, back = ff.pullback(x, p, st)
    for blk in range(3):
        g = np.zeros(B, np.float32); gz = np.zeros((B, D + 3), np.float32)
        if blk == 0: g[:] = np.cos(np.arange(B))
        else: gz[:, D + blk] = 0.5
        dL = onode.ffjord_logpx_grad(fw.us[0], D, g); dL[:, D + 1:] = 0
        adj = onode.adjoint(prob, p, fw, (dL + gz)[None], **tol)
        dx, dp = back(g, gz)
        a = ff.last["adjoint_stats"]
        print(mc, blk, "steps", a.naccept, adj.stats.naccept, "dx err", np.abs(dx - adj.du0[:, :D]).max() / np.abs(adj.du0[:, :D]).max(),
              "dp err", np.abs(dp - adj.dp).max() / np.abs(adj.dp).max(), flush=True)
