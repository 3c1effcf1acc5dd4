#!/usr/bin/env python
"""Pin the oracle (and the CUDA path's golden) against a dump of the UNMODIFIED Julia reference.

    python tools/dump_inputs.py DIR                        # bench.py's inputs as raw Float32
    julia -t auto baseline/run_reference.jl DIR            # wherever a Julia depot exists: writes final_state.f32, du0.f32,
                                                           # dp.f32 and stats.txt (nf / naccept / nreject of both solves)
    python tools/compare_reference_dump.py DIR             # this script (CPU only)

Compares the reference's u(t1), du0, dp and step counts with tests/golden/headline.npz (the oracle at full size on the
same inputs): element-wise rtol 1e-4 (atol = rtol * max|.|) and equality of the forward step count; reports which of the
oracle's accumulation models (blas / defer / exactgemm) the reference's adjoint count matches.
Exit status 0 = the oracle is pinned to the reference on this path."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    d = sys.argv[1]
    g = np.load(os.path.join(ROOT, "tests", "golden", "headline.npz"))
    B, D = bench.BATCH, bench.DIMS[0]
    need = ["final_state.f32", "du0.f32", "dp.f32"]
    missing = [f for f in need if not os.path.exists(os.path.join(d, f))]
    if missing:
        print("no reference dump in", d, "- missing", missing, "(run baseline/run_reference.jl there first)")
        return 2
    uT = np.fromfile(os.path.join(d, "final_state.f32"), "<f4").reshape(B, D)
    du0 = np.fromfile(os.path.join(d, "du0.f32"), "<f4").reshape(B, D)
    dp = np.fromfile(os.path.join(d, "dp.f32"), "<f4")
    ok = True

    def check(name, got, want):
        nonlocal ok
        scale = float(np.abs(want).max())
        err = float((np.abs(got.astype(np.float64) - want) / (np.abs(want) + scale)).max())
        good = err <= 1e-4
        ok &= good
        print(f"{name:6s} max element-wise error {err:.3e} (rtol 1e-4): {'ok' if good else 'MISMATCH'}")

    check("u(t1)", uT[::64], g["uT_s"])
    check("du0", du0[::64], g["du0_s"])
    check("dp", dp, g["dp"])
    st = os.path.join(d, "stats.txt")
    if os.path.exists(st):
        kv = dict(l.strip().split("=") for l in open(st) if "=" in l)
        fwd = [int(kv.get(k, -1)) for k in ("fwd_naccept", "fwd_nreject", "fwd_nf")]
        adj = [int(kv.get(k, -1)) for k in ("adj_naccept", "adj_nreject", "adj_nf")]
        same_f = fwd == [int(x) for x in g["defer_fwd_stats"][:3]]
        ok &= same_f
        print("forward steps", fwd, "oracle", list(g["defer_fwd_stats"][:3]), "ok" if same_f else "MISMATCH")
        models = [m for m in ("blas", "defer", "exactgemm") if adj[0] == int(g[m + "_adj_stats"][0])]
        print("adjoint steps", adj, "oracle models", {m: int(g[m + "_adj_stats"][0]) for m in ("blas", "defer", "exactgemm")},
              "-> matches", models or "none (outside the oracle's noise bracket)" )
    else:
        print("stats.txt absent: step counts not compared")
    print("PINNED" if ok else "NOT PINNED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
