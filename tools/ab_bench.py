#!/usr/bin/env python
"""A/B timing of kernel variants at bench.py's workload: B200NODE_LIB=variants/libb200node_X.so python tools/ab_bench.py [reps] [sense]
prints the forward / adjoint kernel times (library CUDA events, best and median), step counts and a checksum of dp."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import diffeqflux_b200 as dfx  # noqa: E402


def main():
    import torch
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    sense = sys.argv[2] if len(sys.argv) > 2 else "backsolve"
    B = int(sys.argv[3]) if len(sys.argv) > 3 else bench.BATCH
    p, u0 = bench.synth(B)
    dev = torch.device("cuda:0")
    p_d, u0_d = torch.from_numpy(p).to(dev), torch.from_numpy(u0).to(dev)
    dL = torch.ones((1, B, 64), device=dev)
    chain = dfx.Chain(dfx.Dense(64, 256, dfx.tanh), dfx.Dense(256, 256, dfx.tanh), dfx.Dense(256, 64))
    sa = {"backsolve": dfx.BacksolveAdjoint(), "interp": dfx.InterpolatingAdjoint(), "quad": dfx.QuadratureAdjoint()}[sense]
    node = dfx.NeuralODE(chain, bench.TSPAN, dfx.Tsit5(), abstol=bench.TOL, reltol=bench.TOL, save_everystep=False,
                         save_start=False, sensealg=sa, max_dense_steps=64)
    tf, ta = [], []
    for r in range(reps + 2):
        sol, back = node.pullback(u0_d, p_d)
        du0, dp = back(dL)
        fs, as_ = node.last["stats"], node.last["adjoint_stats"]
        if r >= 2:
            tf.append(fs.kernel_ms)
            ta.append(as_.kernel_ms)
    flop = B * (fs.nf * bench.F_FWD + as_.nf * bench.F_ADJ)
    print(f"{os.environ.get('B200NODE_LIB', 'default'):40s} {sense} fwd {min(tf):7.3f}/{np.median(tf):7.3f} ms adj {min(ta):7.3f}/{np.median(ta):7.3f} ms  "
          f"steps {fs.naccept}/{fs.nreject} {as_.naccept}/{as_.nreject} nf {fs.nf}+{as_.nf}  {flop / ((np.median(tf) + np.median(ta)) * 1e-3) / 1e12:6.1f} TFLOP/s  "
          f"dp_sum {float(dp.double().sum()):.6e} du0_sum {float(du0.double().sum()):.6e}", flush=True)


if __name__ == "__main__":
    main()
