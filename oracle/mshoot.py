"""ORACLE (test infrastructure only) -- `multiple_shoot` (single-problem variant).

PARITY UNPINNED (see oracle/tsit5.py header; the reference's multiple-shooting test file is a
skipped stub, /root/reference/test/BasicNeuralDE/multiple_shoot_tests.jl:5-10).  Follows
/root/reference/src/multiple_shooting.jl:36-84 (grouped solves + loss) and :231-243
(`group_ranges`, default continuity loss).

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
