#!/usr/bin/env python
"""Step-sequence diagnostic at bench.py's exact inputs (configs[1], B = 16384): runs the forward solve and the
Backsolve adjoint on both RHS engines with the per-step trace enabled (b200_ctx_trace) and stores
(t, dt, EEst, accepted) per attempted step plus dp and strided slices of u(t1) / du0 in gpurun_out/.
Compared on the CPU side with the oracle's record_trace (tests/golden/make_golden.py headline case)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import diffeqflux_b200 as dfx  # noqa: E402
from diffeqflux_jl_b200 import layers  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else bench.BATCH
    out = {}
    p, u0 = bench.synth(B)
    chain = dfx.Chain(dfx.Dense(64, 256, dfx.tanh), dfx.Dense(256, 256, dfx.tanh), dfx.Dense(256, 64))
    ctx = layers._context(0)
    ctx.trace(capacity=512)
    for name, path in (("ffma", 1), ("tc", 2)):
        node = dfx.NeuralODE(chain, bench.TSPAN, dfx.Tsit5(), abstol=bench.TOL, reltol=bench.TOL, save_everystep=False,
                             save_start=False, sensealg=dfx.BacksolveAdjoint(), path=path)
        sol, back = node.pullback(u0, p)
        tr_f = ctx.trace().copy()
        du0, dp = back(np.ones_like(sol.array()))
        tr_a = ctx.trace().copy()
        fs, as_ = node.last["stats"], node.last["adjoint_stats"]
        print(name, "fwd", fs.naccept, fs.nreject, fs.nf, "adj", as_.naccept, as_.nreject, as_.nf,
              "ms", round(fs.kernel_ms, 2), round(as_.kernel_ms, 2), flush=True)
        out[name] = dict(trace_f=tr_f, trace_a=tr_a, dp=dp, du0=du0[::64], uT=sol.array()[-1][::64],
                         stats=np.array([fs.naccept, fs.nreject, fs.nf, as_.naccept, as_.nreject, as_.nf]))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    np.savez(os.path.join(ROOT, "gpurun_out", f"trace_headline_{B}.npz"),
             **{f"{k}_{kk}": v for k, d in out.items() for kk, v in d.items()})


if __name__ == "__main__":
    main()
