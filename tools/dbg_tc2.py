import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import diffeqflux_b200 as dfx
rng = np.random.default_rng(3)
dims = [64, 256, 256, 64]
p = []
for i, o in zip(dims[:-1], dims[1:]):
    p.append((rng.standard_normal(i * o) * 2.5 / np.sqrt(i)).astype(np.float32)); p.append((rng.standard_normal(o) * 0.25).astype(np.float32))
p = np.concatenate(p)
chain = dfx.Chain(dfx.Dense(64, 256, dfx.tanh), dfx.Dense(256, 256, dfx.tanh), dfx.Dense(256, 64))
from diffeqflux_jl_b200 import layers
ctx = layers._context(0)
ctx.trace(capacity=256)
for B in (200, 1000):
    u0 = rng.standard_normal((B, 64)).astype(np.float32)
    for kw in (dict(save_everystep=False, save_start=False),):
        ref = None
        for path in (1, 2, 2, 2, 2):
            node = dfx.NeuralODE(chain, (0.0, 1.0), dfx.Tsit5(), path=path, abstol=1e-6, reltol=1e-6, **kw)
            sol, _ = node(u0, p, {})
            tr = ctx.trace()
            a = sol.array()[-1]
            if ref is None: ref = a; tr_ref = tr.copy()
            n = min(len(tr), len(tr_ref))
            bad = np.where(np.abs(tr[:n, 2] - tr_ref[:n, 2]) > 0.2 * np.abs(tr_ref[:n, 2]))[0]
            print(B, list(kw), "path", path, "steps", sol.stats.naccept, sol.stats.nreject, "maxdiff vs ffma", float(np.abs(a - ref).max()), "first bad step", bad[:3], tr[bad[:2], :] if len(bad) else "")
