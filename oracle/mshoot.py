"""ORACLE (test infrastructure only) -- `multiple_shoot` (single-problem and ensemble variants).

PARITY UNPINNED (see oracle/tsit5.py header; the reference's multiple-shooting test file is a
skipped stub, /root/reference/test/BasicNeuralDE/multiple_shoot_tests.jl:5-10).  Follows
/root/reference/src/multiple_shooting.jl:36-84 (grouped solves + loss) and :231-243
(`group_ranges`, default continuity loss), and :131-195 (EnsembleProblem method: every trajectory of every group is an
independent solve with its own step sequence).

numpy view of `ode_data`: (T, B, D) C-order == Julia (D, B, T) column-major.
"""
from __future__ import annotations

import numpy as np

from . import node, tsit5

f32 = np.float32


def group_ranges(datasize: int, groupsize: int):
    """0-based half-open (first, last_exclusive) ranges; Julia: [i:min(datasize, i+groupsize-1)
    for i in 1:groupsize-1:datasize-1]."""
    if not (2 <= groupsize <= datasize):
        raise ValueError("datasize must be positive and groupsize must to be within [2, datasize]")
    return [(i - 1, min(datasize, i + groupsize - 1)) for i in range(1, datasize, groupsize - 1)]


def multiple_shoot(problem, p, ode_data, tsteps, group_size, *, continuity_term=100.0,
                   abstol=1e-6, reltol=1e-3, sensealg=node.SENSE_INTERPOLATING, want_grad=True):
    """loss = sum_g sum(abs2, data_g - pred_g) + continuity_term * sum_g>1 sum(abs, pred_{g-1}[end] - data_g[1])
    Returns (loss, preds, dp, stats) ; loss is +inf if any group solve failed."""
    ode_data = np.asarray(ode_data, dtype=np.float32)
    tsteps = np.asarray(tsteps, dtype=np.float32)
    T = ode_data.shape[0]
    if group_size < 2 or group_size > T:
        raise ValueError("group_size can't be < 2 or > number of data points")
    ranges = group_ranges(T, group_size)
    fwds, preds = [], []
    for (a, b) in ranges:
        fw = node.forward(problem, p, ode_data[a], tsteps[a], tsteps[b - 1], saveat=tsteps[a:b],
                          abstol=abstol, reltol=reltol, dense=(sensealg == node.SENSE_INTERPOLATING))
        fwds.append(fw)
        preds.append(fw.us)
    stats = [fw.stats for fw in fwds]
    if any(s.retcode != tsit5.RET_SUCCESS for s in stats):
        return float("inf"), preds, None, stats
    loss = 0.0
    dLs = []
    for g, (a, b) in enumerate(ranges):
        diff = (preds[g] - ode_data[a:b]).astype(np.float32)
        loss += float(np.sum(diff * diff, dtype=np.float64))
        dLs.append(2 * diff)
    for g in range(1, len(ranges)):
        a, _ = ranges[g]
        d = (preds[g - 1][-1] - ode_data[a]).astype(np.float32)
        loss += float(continuity_term) * float(np.sum(np.abs(d), dtype=np.float64))
        dLs[g - 1][-1] += f32(continuity_term) * np.sign(d)
    if not want_grad:
        return loss, preds, None, stats
    dp = np.zeros(p.size, np.float64)
    adj_stats = []
    for g, fw in enumerate(fwds):
        adj = node.adjoint(problem, p, fw, dLs[g].astype(np.float32), sensealg=sensealg,
                           abstol=abstol, reltol=reltol)
        dp += adj.dp
        adj_stats.append(adj.stats)
    return loss, preds, dp.astype(np.float32), (stats, adj_stats)


def multiple_shoot_ensemble(problem, p, ode_data, tsteps, group_size, *, continuity_term=100.0, abstol=1e-6, reltol=1e-3,
                            sensealg=node.SENSE_INTERPOLATING):
    """Ensemble method (/root/reference/src/multiple_shooting.jl:131-195).  ode_data: (trajectories, T, D) C-order == Julia
    (D, T, trajectories).  loss = sum over groups of sum(abs2, u - u_hat) over all trajectories + continuity terms; every
    (group, trajectory) solve is independent, i.e. a batch-1 run of the shared-dt integrator.  Returns (loss, preds, dp)
    with preds[g] of shape (trajectories, len_g, D); loss is +inf if any solve failed."""
    ode_data = np.asarray(ode_data, dtype=np.float32)
    ntraj = ode_data.shape[0]
    total, dp = 0.0, np.zeros(p.size, np.float64)
    per = []
    for i in range(ntraj):
        loss_i, preds_i, dp_i, _ = multiple_shoot(problem, p, ode_data[i][:, None, :], tsteps, group_size, continuity_term=continuity_term,
                                                  abstol=abstol, reltol=reltol, sensealg=sensealg)
        if not np.isfinite(loss_i):
            return float("inf"), None, None
        total += loss_i
        dp += dp_i
        per.append(preds_i)
    preds = [np.concatenate([per[i][g] for i in range(ntraj)], axis=1).transpose(1, 0, 2) for g in range(len(per[0]))]
    return total, preds, dp.astype(np.float32)
