#!/usr/bin/env python
"""Write the bench.py workload (BASELINE.json configs[1]) as raw Float32 files for baseline/run_reference.jl:
p.f32 (flat parameters, the C ABI's order = Julia's ComponentArray order), u0.f32 ([B][D] = Julia (D, B)),
meta.txt.  usage: python tools/dump_inputs.py OUTDIR [batch]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    out = sys.argv[1]
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else bench.BATCH
    os.makedirs(out, exist_ok=True)
    p, u0 = bench.synth(batch)
    p.astype("<f4").tofile(os.path.join(out, "p.f32"))
    u0.astype("<f4").tofile(os.path.join(out, "u0.f32"))
    with open(os.path.join(out, "meta.txt"), "w") as f:
        f.write("dims=" + ",".join(str(d) for d in bench.DIMS) + "\n")
        f.write(f"batch={batch}\n")
        f.write("tol=1e-6\n")
    print("wrote", out, "params", p.size, "u0", u0.shape)


if __name__ == "__main__":
    main()
