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
for B in (64, 200):
    u0 = rng.standard_normal((B, 64)).astype(np.float32)
    out = {}
    for path in (1, 2):
        # one fixed tiny step: u(t1) - u0 = dt * (RK combination of RHS evaluations)
        node = dfx.NeuralODE(chain, (0.0, 1e-3), dfx.Tsit5(), path=path, abstol=1e-6, reltol=1e-6, dt=1e-3, save_everystep=False, save_start=False)
        sol, _ = node(u0, p, {})
        out[path] = (sol.array()[-1] - u0) / 1e-3
        print(B, path, sol.stats.naccept, sol.stats.nreject, sol.retcode)
    d = np.abs(out[2] - out[1])
    print("B", B, "max abs diff of RHS-ish", d.max(), "scale", np.abs(out[1]).max(), "bad rows", np.unique(np.where(d > 1e-2)[0])[:20], "bad cols", np.unique(np.where(d > 1e-2)[1])[:64])
