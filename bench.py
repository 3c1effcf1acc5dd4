#!/usr/bin/env python
"""bench.py -- NeuralODE forward+adjoint RHS evaluations per second (BASELINE.json metric).

A "step" is one pass of the hot path over one batch of synthetic input: one forward solve
(b200_node_forward) plus one adjoint gradient (b200_node_adjoint) of BASELINE.json configs[1]:
batch 16384, state 64, MLP 64->256->256->64 tanh, Tsit5 abstol=reltol=1e-6, BacksolveAdjoint,
end state only, dL/du(t1) = ones.

  value : B * (nf_fwd + nf_adj) / time with inputs resident in HBM (device pointers), CUDA events
  e2e   : the same through the layer API with HOST buffers and the C ABI's B200_PTR_HOST path (the library copies
          u0 and p to the device and u(t1), du0, dp back inside the timed region; the inputs sit in pinned memory)
  roofline    : algorithmic FLOPs of the two persistent kernels / their device time (library events) against the
          directly measured TF32 tensor-core peak / 3 (profiles/tf32_peak_r2.json, tools/tf32_peak.cu)
  cpu_baseline: the restated CPU path (oracle, numpy + threaded BLAS, all host cores) on the full batch
  --impl reference : times that CPU path alone (Julia is not installable here: see DESIGN.md)
  --scaling strong : the 16384-trajectory batch is split over the ranks (default: weak, 16384 per rank)

Launch: python bench.py [--gpus N --steps K --warmup W]; for N > 1 under torchrun (one rank per GPU).  At N > 1 the
run first checks, untimed, that a batch sharded over the ranks takes the same steps and returns the same states and
gradients as the same batch on one GPU (`config.sharded_equals_single`; the run fails if it does not).
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
WORKLOAD = "configs[1] batched NeuralODE 64->256->256->64 tanh, Tsit5 abstol=reltol=1e-6, BacksolveAdjoint, end state only"


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


def tf32_peak():
    """Directly measured dense TF32 tensor-core rate of the whole GPU (tools/tf32_peak.cu: every SM issues
    tcgen05.mma kind::tf32 M128 N256 K8 back to back for seconds; committed under profiles/).  The solver's 3xTF32
    GEMMs spend three such MMAs per algorithmic product, so the algorithmic ceiling is this figure / 3.  The step is a
    long-running kernel: the SUSTAINED figure of the kernel's own issue pattern is the denominator (the burst one is
    reported beside it)."""
    path = os.path.join(ROOT, "profiles", "tf32_peak_r2.json")
    if os.path.exists(path):
        j = json.load(open(path))
        # the higher of the two issue patterns each way: the all-distinct pattern bursts higher (1110) but settles lower
        # under the 660 W it draws; the 3xTF32 pattern holds 1010
        sus = max(j["tf32_3x_pattern"]["sustained_tflops"], j["tf32_dense"]["sustained_tflops"])
        bur = max(j["tf32_3x_pattern"]["burst_tflops"], j["tf32_dense"]["burst_tflops"])
        return sus, bur, "measured: profiles/tf32_peak_r2.json (tools/tf32_peak.cu, tcgen05 kind::tf32 M128xN256xK8 on all 148 SMs, converged-warp issue)"
    pk, src = peaks()
    return pk["bf16_tflops_sustained"] / 2.0, pk["bf16_tflops"] / 2.0, f"derived: {src} bf16 / 2"


# dram__bytes_read.sum + dram__bytes_write.sum of ONE adjoint solve-kernel launch at this workload (ncu --set full capture of
# this round's build, profiles/solve_kernels_r2_tc_raw.csv, second row); algorithmic bytes are ~0 (the path is tensor bound)
ADJ_DRAM_BYTES_PER_LAUNCH = 29.147705e9 + 28.179562e9


def golden_steps():
    """Step statistics of the full-size oracle runs on exactly these inputs (tests/golden/headline.npz)."""
    path = os.path.join(ROOT, "tests", "golden", "headline.npz")
    if not os.path.exists(path):
        return None
    g = np.load(path)
    return {m: {"fwd": [int(x) for x in g[m + "_fwd_stats"][:3]], "adj": [int(x) for x in g[m + "_adj_stats"][:3]]} for m in ("blas", "defer", "exactgemm")}


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
    """One forward + Backsolve adjoint of the restated CPU path with every host core the BLAS may use (pinned
    explicitly: torchrun exports OMP_NUM_THREADS=1); returns (seconds, nf_fwd, nf_adj, forward, adjoint, threads)."""
    from oracle import node as onode
    from oracle.mlp import ACT_ID, ACT_TANH, MLP
    mlp = MLP(DIMS, [ACT_TANH, ACT_TANH, ACT_ID])
    prob = onode.NodeProblem(mlp)
    ncores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        import threadpoolctl
        limiter = threadpoolctl.threadpool_limits(limits=ncores)
    except Exception:
        limiter = None
    try:
        used = cpu_threads()
        t0 = time.perf_counter()
        fw = onode.forward(prob, p, u0, TSPAN[0], TSPAN[1], abstol=TOL, reltol=TOL, save_everystep=False, save_start=False)
        adj = onode.adjoint(prob, p, fw, np.ones_like(fw.us), sensealg=onode.SENSE_BACKSOLVE, abstol=TOL, reltol=TOL)
        dt = time.perf_counter() - t0
    finally:
        if limiter is not None:
            limiter.restore_original_limits()
    return dt, fw.stats.nf, adj.stats.nf, fw, adj, used


def cpu_threads():
    try:
        import threadpoolctl
        infos = threadpoolctl.threadpool_info()
        return max([i.get("num_threads", 1) for i in infos] + [1])
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    """The reference arm: DiffEqFlux's CPU path cannot be installed (Julia absent, no network), so the restated
    oracle stands in for it -- numpy with all host BLAS threads on the FULL batch of the GPU arm's workload (one
    forward + adjoint pass per step, ~20 s on 8 cores; steps / warm-up are capped so the run ends within minutes)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    p, u0 = synth(BATCH)
    warm = min(args.warmup, 1)
    cores = 1
    for _ in range(warm):
        cpu_pass(*synth(1024))                       # BLAS thread pool / page faults; a full pass costs ~20 s
    times, evals = [], 0
    for _ in range(max(1, min(args.steps, 3))):
        dt, nf, na, _, _, cores = cpu_pass(p, u0)
        times.append(dt)
        evals = BATCH * (nf + na)
    sec = float(np.mean(times))
    val = evals / sec
    line = {"impl": "reference", "metric": "NeuralODE fwd+adjoint RHS evals/sec", "value": val, "unit": "RHS evals/s",
            "n_gpus": args.gpus, "steps": len(times), "warmup": warm, "ms_per_step": sec * 1e3,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch": BATCH, "state": 64, "w_scale": W_SCALE, "tspan": list(TSPAN), "nf_fwd": nf, "nf_adj": na},
            "cpu_baseline": {"value": val, "unit": "RHS evals/s", "cores": cores, "kind": "port",
                             "sample": f"the full {BATCH}-trajectory batch, one forward+adjoint pass per step (numpy + OpenBLAS on {cores} threads; Julia is not installable here)"},
            "e2e": {"value": val, "unit": "RHS evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------- GPU arm
def make_node(dfx):
    chain = dfx.Chain(dfx.Dense(64, 256, dfx.tanh), dfx.Dense(256, 256, dfx.tanh), dfx.Dense(256, 64))
    return dfx.NeuralODE(chain, TSPAN, dfx.Tsit5(), abstol=TOL, reltol=TOL, save_everystep=False, save_start=False,
                         sensealg=dfx.BacksolveAdjoint())


def stats3(s):
    return [s.naccept, s.nreject, s.nf]


def sharded_check(torch, dist, dfx, layers, dev, rank, world, per_rank=2048):
    """Untimed multi-GPU proof: a (per_rank * world)-trajectory batch solved (a) on rank 0 alone, before the process
    group is attached, and (b) sharded over all ranks with the in-kernel cross-rank exchange.  Returns a dict for
    `config`; `sharded_equals_single` requires identical step statistics on every rank and states / gradients equal to
    1e-6 relative (they are in fact bitwise equal whenever the fp64 cross-rank sums round alike: `sharded_bitwise`)."""
    nb = per_rank * world
    p_h, u0_h = synth(nb, SEED + 17)
    node = make_node(dfx)
    p_d = torch.from_numpy(p_h).to(dev)
    single = None
    if rank == 0:
        u_all = torch.from_numpy(u0_h).to(dev)
        sol, back = node.pullback(u_all, p_d)
        du0, dp = back(torch.ones((1, nb, 64), device=dev))
        single = dict(u=sol.array()[-1].clone(), du0=du0.clone(), dp=dp.clone(), st=stats3(node.last["stats"]) + stats3(node.last["adjoint_stats"]))
    dist.barrier()
    layers.attach_process_group(dev.index)                      # from here on every solve on this device is one shard
    lo, hi = rank * per_rank, (rank + 1) * per_rank
    u_sh = torch.from_numpy(u0_h[lo:hi]).to(dev)
    sol, back = node.pullback(u_sh, p_d)
    du0, dp = back(torch.ones((1, per_rank, 64), device=dev))
    st = torch.tensor(stats3(node.last["stats"]) + stats3(node.last["adjoint_stats"]), device=dev, dtype=torch.int64)
    st_all = [torch.empty_like(st) for _ in range(world)]
    dist.all_gather(st_all, st)
    u_all = [torch.empty_like(sol.array()[-1]) for _ in range(world)]
    dist.all_gather(u_all, sol.array()[-1].contiguous())
    d_all = [torch.empty_like(du0) for _ in range(world)]
    dist.all_gather(d_all, du0.contiguous())
    dp_all = [torch.empty_like(dp) for _ in range(world)]
    dist.all_gather(dp_all, dp.contiguous())
    out = None
    if rank == 0:
        def rel(a, b):
            return float((a.double() - b.double()).abs().max() / b.double().abs().max())
        u_s, d_s = torch.cat(u_all), torch.cat(d_all)
        steps_same = all(x.tolist() == single["st"] for x in st_all)
        dp_same_ranks = all(torch.equal(x, dp_all[0]) for x in dp_all)
        errs = {"u": rel(u_s, single["u"]), "du0": rel(d_s, single["du0"]), "dp": rel(dp_all[0], single["dp"])}
        bitwise = bool(torch.equal(u_s, single["u"]) and torch.equal(d_s, single["du0"]) and torch.equal(dp_all[0], single["dp"]))
        ok = bool(steps_same and dp_same_ranks and max(errs.values()) <= 1e-6)
        out = {"sharded_equals_single": ok, "sharded_bitwise": bitwise, "sharded_check_batch": nb, "sharded_steps": st_all[0].tolist(),
               "single_steps": single["st"], "sharded_max_rel_diff": errs}
    return out


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
    shard_info = None
    if world > 1:
        shard_info = sharded_check(torch, dist, dfx, layers, dev, rank, world)   # also attaches the IPC mailboxes
    strong = args.scaling == "strong"
    if strong and BATCH % world:
        raise SystemExit(f"--scaling strong: {BATCH} trajectories do not split over {world} ranks")
    B = BATCH // world if strong else BATCH       # weak: every rank integrates BATCH trajectories; strong: BATCH in total
    if strong:
        u0_h = synth(BATCH, SEED)[1][rank * B:(rank + 1) * B]
    else:
        _, u0_h = synth(B, SEED + rank * 7919 if rank else SEED)   # rank 0 = the golden inputs; the others their own trajectories
    p_h = synth(8, SEED)[0]                       # same weights on every rank
    node = make_node(dfx)
    p_d = torch.from_numpy(p_h).to(dev)
    u0_d = torch.from_numpy(u0_h).to(dev)
    dL_d = torch.ones((1, B, 64), device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)    # > 126 MB L2

    def step_device():
        sol, back = node.pullback(u0_d, p_d)
        du0, dp = back(dL_d)
        fw = node.last
        return fw["stats"], fw["adjoint_stats"], dp

    # end-to-end arm: HOST buffers through the C ABI's B200_PTR_HOST path (what a Julia `Array` caller hits): the
    # library itself copies u0 / p host -> device and u(t1) / du0 / dp device -> host; the inputs are pinned
    u0_pin = torch.from_numpy(u0_h).pin_memory().numpy()
    p_pin = torch.from_numpy(p_h).pin_memory().numpy()
    dL_h = np.ones((1, B, 64), np.float32)

    def step_e2e():
        sol, back = node.pullback(u0_pin, p_pin)
        du0, dp = back(dL_h)
        return sol.array(), du0, dp

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                       # nvidia-smi needs ~0.5 s to start: launch it before the warm-up
    for _ in range(max(args.warmup, 3)):
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
    for _ in range(2):
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
        tf32_sus, tf32_burst, tf32_src = tf32_peak()
        tensor_peak = tf32_sus / 3.0                            # 3xTF32: three tensor-core products per algorithmic one
        ffma_peak = 148 * 128 * 2 * pk.get("sm_max_mhz", 1965.0) * 1e6 / 1e12
        gold = golden_steps()
        steps_now = {"fwd": stats3(fs), "adj": stats3(as_)}
        match = None
        bracket = None
        if gold and not strong and world == 1:
            match = steps_now == gold["defer"]
            bracket = steps_now["fwd"] == gold["defer"]["fwd"] and gold["exactgemm"]["adj"][0] <= as_.naccept <= gold["defer"]["adj"][0] and as_.nreject == 0
        cpu = None
        if world == 1:
            ps, us = synth(BATCH)
            dt, nfc, nac, _, _, cores = cpu_pass(ps, us)
            cpu = {"value": BATCH * (nfc + nac) / dt, "unit": "RHS evals/s", "cores": cores, "kind": "port",
                   "sample": f"the full {BATCH}-trajectory batch, one forward+adjoint pass ({dt:.1f} s), numpy + OpenBLAS restatement on {cores} threads (Julia not installable)"}
        config = {"workload": WORKLOAD, "batch_per_gpu": B, "batch_total": B * world, "state": 64, "w_scale": W_SCALE, "tspan": list(TSPAN),
                  "l2_flush_between_steps": True,
                  "nf_fwd": fs.nf, "nf_adj": as_.nf, "naccept_fwd": fs.naccept, "nreject_fwd": fs.nreject,
                  "naccept_adj": as_.naccept, "nreject_adj": as_.nreject, "grid": fs.grid, "tile": fs.tile,
                  "engine_fwd": fs.engine, "engine_adj": as_.engine,
                  # oracle runs on exactly these inputs (tests/golden/headline.npz): "defer" is the parity target (the
                  # FP32-FFMA engine reproduces it exactly, tests/test_fullsize.py); the adjoint count of the tensor-core
                  # engine lies between "exactgemm" and "defer" because its ramp-up steps see a lower rounding-noise floor
                  "golden_steps": gold, "steps_match_golden": match, "steps_within_golden_noise_bracket": bracket,
                  "parallelism": f"batch-sharded x{world}"}
        if shard_info:
            config.update(shard_info)
        line = {
            "metric": "NeuralODE fwd+adjoint RHS evals/sec", "value": value, "unit": "RHS evals/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": tensor_peak, "unit": "TFLOP/s", "frac": achieved / tensor_peak,
                         "traffic": ADJ_DRAM_BYTES_PER_LAUNCH, "traffic_source": "ncu --set full, adjoint solve kernel (dominant), profiles/solve_kernels_r2_tc_raw.csv",
                         "peak_source": f"{tf32_src}: sustained {tf32_sus:.1f} TFLOP/s TF32 / 3 (burst {tf32_burst:.1f} / 3 = {tf32_burst / 3:.1f})",
                         "frac_of_burst_peak": achieved / (tf32_burst / 3.0),
                         "kernel_ms_per_step": k_ms / args.steps, "algorithmic_flop_per_step": flop,
                         "ffma_peak_tflops": ffma_peak, "frac_of_ffma_peak": achieved / ffma_peak,
                         "engine": "tcgen05 3xTF32" if fs.engine else "fp32 ffma"},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_val, "unit": "RHS evals/s", "h2d_bytes_per_step": int(u0_h.nbytes + p_h.nbytes + dL_h.nbytes),
                    "d2h_bytes_per_step": int(2 * u0_h.nbytes + p_h.nbytes), "steps": e2e_steps,
                    "path": "layer API -> b200_node_forward / b200_node_adjoint with B200_PTR_HOST (library-side copies)"},
            "gpu_launches": 7 * args.steps,      # per step: pack_tc, pack_tc_t, pack_params, solve (forward), solve (adjoint), extract_cols, unpack_grad
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    ok = True
    if world > 1:
        flag = torch.tensor([1 if (rank != 0 or (shard_info and shard_info["sharded_equals_single"])) else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok = bool(flag.item())
        dist.destroy_process_group()
    if not ok:
        raise SystemExit("sharded solve differs from the single-GPU solve (config.sharded_equals_single = false)")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
