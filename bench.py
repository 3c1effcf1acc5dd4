#!/usr/bin/env python
"""bench.py -- NeuralODE forward+adjoint RHS evaluations per second (BASELINE.json metric).

A "step" is one pass of the hot path over one batch of synthetic input: one forward solve
(b200_node_forward) plus one adjoint gradient (b200_node_adjoint) of BASELINE.json configs[1]:
batch 16384, state 64, MLP 64->256->256->64 tanh, Tsit5 abstol=reltol=1e-6, BacksolveAdjoint,
end state only, dL/du(t1) = ones.

  value : B * (nf_fwd + nf_adj) / time with inputs resident in HBM (device pointers), CUDA events
  e2e   : the same through the layer API with HOST buffers (H2D of u0 and p, D2H of u(t1), du0, dp
          inside the timed region)
  roofline    : algorithmic FLOPs of the two persistent kernels / their device time (library events)
  cpu_baseline: the restated CPU path (oracle, numpy + threaded BLAS) on a bounded sample
  --impl reference : times that CPU path alone (Julia is not installable here: see DESIGN.md)

Launch: python bench.py [--gpus N --steps K --warmup W]; for N > 1 under torchrun (one rank per GPU).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DIMS = [64, 256, 256, 64]
BATCH = 16384
TSPAN = (0.0, 1.0)
TOL = 1e-6
SEED = 2
W_SCALE = 2.0          # W ~ N(0, 1/in) * W_SCALE: truncation error (not rounding noise) drives the controller
F_FWD = 2 * sum(i * o for i, o in zip(DIMS[:-1], DIMS[1:]))     # 196,608 FLOP / trajectory / forward RHS
F_ADJ = 3 * F_FWD                                               # recompute + dgrad + wgrad


def synth(batch, seed=SEED):
    """Synthetic inputs shared by every arm (numpy default_rng, fp32)."""
    rng = np.random.default_rng(seed)
    p = []
    for i, o in zip(DIMS[:-1], DIMS[1:]):
        p.append((rng.standard_normal(i * o) * W_SCALE / np.sqrt(i)).astype(np.float32))
        p.append((rng.standard_normal(o) * 0.1 * W_SCALE).astype(np.float32))
    p = np.concatenate(p)
    u0 = np.random.default_rng(seed + 1000).standard_normal((batch, DIMS[0])).astype(np.float32)
    return p, u0


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "sm_max_mhz": 1965.0}, "fallback"


# dram__bytes_read.sum + dram__bytes_write.sum of ONE adjoint solve-kernel launch at this workload (ncu --set full capture
# of the final round-1 build, profiles/solve_kernels_r1_final_raw.csv); algorithmic bytes are ~0 (the path is tensor bound)
ADJ_DRAM_BYTES_PER_LAUNCH = 30.955821e9 + 29.046970e9


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.rows, self.proc, self.index, self.start_row = [], None, index, 0

    def mark(self):
        self.start_row = max(0, len(self.rows) - 1)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            time.sleep(0.12)                  # let the last sample of the timed region land
            self.proc.terminate()
        rows = self.rows[self.start_row:] or self.rows
        self.rows = rows
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        mx = max((float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()), default=0.0)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------- CPU arm
def cpu_pass(p, u0):
    """One forward + Backsolve adjoint of the restated CPU path; returns (seconds, nf_fwd, nf_adj)."""
    from oracle import node as onode
    from oracle.mlp import ACT_ID, ACT_TANH, MLP
    mlp = MLP(DIMS, [ACT_TANH, ACT_TANH, ACT_ID])
    prob = onode.NodeProblem(mlp)
    t0 = time.perf_counter()
    fw = onode.forward(prob, p, u0, TSPAN[0], TSPAN[1], abstol=TOL, reltol=TOL, save_everystep=False, save_start=False)
    adj = onode.adjoint(prob, p, fw, np.ones_like(fw.us), sensealg=onode.SENSE_BACKSOLVE, abstol=TOL, reltol=TOL)
    return time.perf_counter() - t0, fw.stats.nf, adj.stats.nf, fw, adj


def cpu_threads():
    try:
        import threadpoolctl
        infos = threadpoolctl.threadpool_info()
        return max([i.get("num_threads", 1) for i in infos] + [1])
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    """The reference arm: DiffEqFlux's CPU path cannot be installed (Julia absent, no network), so the
    restated oracle stands in for it -- numpy with all host BLAS threads on a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = 256
    p, u0 = synth(sample)
    for _ in range(min(args.warmup, 1)):
        cpu_pass(p, u0)
    times, evals = [], 0
    for _ in range(max(1, min(args.steps, 3))):
        dt, nf, na, _, _ = cpu_pass(p, u0)
        times.append(dt)
        evals = sample * (nf + na)
    sec = float(np.mean(times))
    val = evals / sec
    cores = cpu_threads()
    line = {"impl": "reference", "metric": "NeuralODE fwd+adjoint RHS evals/sec", "value": val, "unit": "RHS evals/s",
            "n_gpus": args.gpus, "steps": len(times), "warmup": min(args.warmup, 1), "ms_per_step": sec * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[1] batched NeuralODE 64->256->256->64 tanh, Tsit5 1e-6, BacksolveAdjoint",
                       "batch": BATCH, "sample_batch": sample, "w_scale": W_SCALE},
            "cpu_baseline": {"value": val, "unit": "RHS evals/s", "cores": cores, "kind": "port",
                             "sample": f"{sample} of {BATCH} trajectories, 1 forward+adjoint pass per step (numpy + threaded BLAS; Julia not installable here)"},
            "e2e": {"value": val, "unit": "RHS evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist
    import diffeqflux_b200 as dfx
    from diffeqflux_jl_b200 import layers

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    dev = torch.device(f"cuda:{local}")
    B = BATCH                                     # weak scaling: every rank integrates BATCH trajectories
    _, u0_h = synth(B, SEED + rank * 7919)        # every rank integrates its own trajectories
    p_h = synth(8, SEED)[0]                       # same weights on every rank
    chain = dfx.Chain(dfx.Dense(64, 256, dfx.tanh), dfx.Dense(256, 256, dfx.tanh), dfx.Dense(256, 64))
    node = dfx.NeuralODE(chain, TSPAN, dfx.Tsit5(), abstol=TOL, reltol=TOL, save_everystep=False, save_start=False,
                         sensealg=dfx.BacksolveAdjoint())
    if world > 1:
        layers.attach_process_group(local)        # IPC mailboxes for the in-kernel error-norm exchange
    p_d = torch.from_numpy(p_h).to(dev)
    u0_d = torch.from_numpy(u0_h).to(dev)
    dL_d = torch.ones((1, B, 64), device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)    # > 126 MB L2

    def step_device():
        sol, back = node.pullback(u0_d, p_d)
        du0, dp = back(dL_d)
        fw = node.last
        return fw["stats"], fw["adjoint_stats"], dp

    # pinned host buffers for the end-to-end arm
    u0_pin = torch.from_numpy(u0_h).pin_memory()
    p_pin = torch.from_numpy(p_h).pin_memory()

    def step_e2e():
        u = u0_pin.to(dev, non_blocking=True)
        pp = p_pin.to(dev, non_blocking=True)
        sol, back = node.pullback(u, pp)
        du0, dp = back(dL_d)
        out = (sol.array().cpu(), du0.cpu(), dp.cpu())        # D2H of u(t1), du0, dp
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                       # nvidia-smi needs ~0.5 s to start: launch it before the warm-up
    for _ in range(args.warmup):
        fs, as_, _ = step_device()
    barrier()
    if rank == 0:
        sampler.mark()                        # only samples taken from here on count (timed region)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k_ms = 0.0
    e0.record()
    for _ in range(args.steps):
        flush.fill_(1)                                         # L2 flush between timed iterations
        fs, as_, _ = step_device()
        k_ms += fs.kernel_ms + as_.kernel_ms
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    nf = fs.nf + as_.nf
    evals_per_step = B * world * nf
    value = evals_per_step * args.steps / (ms * 1e-3)

    # end-to-end (host buffers)
    for _ in range(min(args.warmup, 2)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 5))
    for _ in range(e2e_steps):
        step_e2e()
    barrier()
    t_e2e = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_val = evals_per_step * e2e_steps / float(t_e2e.item())

    if rank == 0:
        pk, pk_src = peaks()
        flop = B * (fs.nf * F_FWD + as_.nf * F_ADJ)            # algorithmic, per step, per GPU
        kernel_s = k_ms / args.steps * 1e-3
        achieved = flop / kernel_s / 1e12
        tensor_peak = pk["bf16_tflops_sustained"] / 2.0 / 3.0   # 3xTF32 algorithmic ceiling
        ffma_peak = 148 * 128 * 2 * pk.get("sm_max_mhz", 1965.0) * 1e6 / 1e12
        cpu = None
        if True:
            sample = 128
            ps, us = synth(sample)
            dt, nfc, nac, _, _ = cpu_pass(ps, us)
            cpu = {"value": sample * (nfc + nac) / dt, "unit": "RHS evals/s", "cores": cpu_threads(), "kind": "port",
                   "sample": f"{sample} of {BATCH} trajectories, one forward+adjoint pass ({dt:.1f} s), numpy + threaded BLAS restatement (Julia not installable)"}
        line = {
            "metric": "NeuralODE fwd+adjoint RHS evals/sec", "value": value, "unit": "RHS evals/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[1] batched NeuralODE 64->256->256->64 tanh, Tsit5 abstol=reltol=1e-6, BacksolveAdjoint, end state only",
                       "batch_per_gpu": B, "state": 64, "w_scale": W_SCALE, "tspan": list(TSPAN), "l2_flush_between_steps": True,
                       "nf_fwd": fs.nf, "nf_adj": as_.nf, "naccept_fwd": fs.naccept, "nreject_fwd": fs.nreject,
                       "naccept_adj": as_.naccept, "nreject_adj": as_.nreject, "grid": fs.grid, "tile": fs.tile,
                       "engine_fwd": fs.engine, "engine_adj": as_.engine,
                       "parallelism": f"batch-sharded x{world}"},
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": tensor_peak, "unit": "TFLOP/s", "frac": achieved / tensor_peak,
                         "traffic": ADJ_DRAM_BYTES_PER_LAUNCH, "traffic_source": "ncu --set full, adjoint solve kernel (dominant), profiles/solve_kernels_r1_final_raw.csv", "peak_source": f"{pk_src} bf16 sustained / 2 (TF32) / 3 (3xTF32), BASELINE.md section 2",
                         "kernel_ms_per_step": k_ms / args.steps, "algorithmic_flop_per_step": flop,
                         "ffma_peak_tflops": ffma_peak, "frac_of_ffma_peak": achieved / ffma_peak,
                         "engine": "tcgen05 3xTF32" if fs.engine else "fp32 ffma"},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_val, "unit": "RHS evals/s", "h2d_bytes_per_step": int(u0_h.nbytes + p_h.nbytes),
                    "d2h_bytes_per_step": int(2 * u0_h.nbytes + p_h.nbytes), "steps": e2e_steps},
            "gpu_launches": 5 * args.steps,
            "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
