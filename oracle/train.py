"""ORACLE (test infrastructure only) -- the caller's side of the hot path: one iteration of the training loop of
/root/reference/docs/src/examples/neural_ode.md:36-44,64-70 (`loss = sum(abs2, ode_data .- pred)`, Zygote gradient through the
NeuralODE, `OptimizationOptimisers.Adam(lr)`) and the `DimMover` layer (/root/reference/src/neural_de.jl:397-412).

PARITY UNPINNED (see oracle/tsit5.py).  Adam follows Optimisers.jl's `Adam` rule (the dependency is not vendored under
/root/reference; OptimizationOptimisers compat in /root/reference/docs/Project.toml): defaults beta = (0.9, 0.999),
eps = 1e-8;  m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2;  p -= lr * m / (1 - b1^t) / (sqrt(v / (1 - b2^t)) + eps).
"""
from __future__ import annotations

import numpy as np

from . import node as onode

f32 = np.float32


def adam_update(p, m, v, g, step, lr=0.05, beta1=0.9, beta2=0.999, eps=1e-8):
    b1, b2, lr, eps = f32(beta1), f32(beta2), f32(lr), f32(eps)
    m = (b1 * m + (f32(1) - b1) * g).astype(np.float32)
    v = (b2 * v + (f32(1) - b2) * g * g).astype(np.float32)
    c1 = f32(1) - f32(np.power(b1, f32(step)))
    c2 = f32(1) - f32(np.power(b2, f32(step)))
    p = (p - lr * (m / c1) / (np.sqrt(v / c2) + eps)).astype(np.float32)
    return p, m, v


def train_step(problem, p, u0, data, t0, t1, step, m, v, *, sensealg=onode.SENSE_INTERPOLATING, lr=0.05, augment=0, **kw):
    """returns (loss before the update, p_new, m_new, v_new, gradient)."""
    tol = dict(abstol=kw.get("abstol", 1e-6), reltol=kw.get("reltol", 1e-3))
    fw = onode.forward(problem, p, u0, t0, t1, dense=(sensealg == onode.SENSE_INTERPOLATING), augment=augment, **kw)
    d = (fw.us - np.asarray(data, np.float32).reshape(fw.us.shape)).astype(np.float32)
    loss = float(np.sum(d.astype(np.float64) ** 2))
    adj = onode.adjoint(problem, p, fw, (f32(2) * d).astype(np.float32), sensealg=sensealg, augment=augment, **tol)
    return (loss,) + adam_update(p, m, v, adj.dp, step, lr=lr) + (adj.dp,)


def dimmover(x_julia, from_=-2, to=-1):
    """`cat(eachslice(x; dims = from)...; dims = to)` on an array indexed in JULIA dimension order (a numpy array whose
    axis k is Julia's dimension k + 1)."""
    nd = x_julia.ndim
    f = (from_ if from_ > 0 else nd + 1 + from_) - 1
    t = (to if to > 0 else nd + 1 + to) - 1
    slices = [np.take(x_julia, i, axis=f) for i in range(x_julia.shape[f])]
    if t == nd - 1:
        return np.stack(slices, axis=t)
    return np.concatenate(slices, axis=t)
