"""ORACLE (test infrastructure only) -- Lux `Chain(Dense...)` forward / VJP / second-order VJP.

PARITY UNPINNED (see oracle/tsit5.py header).  Restates what the reference's RHS closure
`dudt(u,p,t) = model(u,p)` (/root/reference/src/neural_de.jl:48-50) evaluates through Lux
`Dense`: y = act.(W*x .+ b) with W (out,in) column-major and the flat ComponentArray parameter
vector ordered [vec(W1); b1; vec(W2); b2; ...] (/root/reference/test/BasicNeuralDE/
neural_de_tests.jl:44-45).

Memory convention: a Julia (D,B) column-major matrix has the same bytes as a numpy (B,D) C-order
array; a Julia (out,in) column-major W has the same bytes as a numpy (in,out) C-order array.  All
arrays here use the numpy view, so y = x @ W + b.
"""
from __future__ import annotations

import numpy as np

ACT_ID, ACT_TANH, ACT_RELU, ACT_SIGMOID, ACT_SOFTPLUS = 0, 1, 2, 3, 4
PRE_NONE, PRE_CUBE = 0, 1


def _act(a, x):
    if a == ACT_ID:
        return x
    if a == ACT_TANH:
        return np.tanh(x)
    if a == ACT_RELU:
        return np.maximum(x, 0)
    if a == ACT_SIGMOID:
        return (1 / (1 + np.exp(-x))).astype(x.dtype)
    if a == ACT_SOFTPLUS:
        return (np.maximum(x, 0) + np.log1p(np.exp(-np.abs(x)))).astype(x.dtype)
    raise ValueError(a)


def _dact(a, pre, h):
    """act'(pre) given pre-activation and h = act(pre)."""
    if a == ACT_ID:
        return np.ones_like(h)
    if a == ACT_TANH:
        return 1 - h * h
    if a == ACT_RELU:
        return (pre > 0).astype(h.dtype)
    if a == ACT_SIGMOID:
        return h * (1 - h)
    if a == ACT_SOFTPLUS:
        return (1 / (1 + np.exp(-pre))).astype(h.dtype)
    raise ValueError(a)


def _d2act(a, pre, h):
    if a in (ACT_ID, ACT_RELU):
        return np.zeros_like(h)
    if a == ACT_TANH:
        return -2 * h * (1 - h * h)
    if a == ACT_SIGMOID:
        return h * (1 - h) * (1 - 2 * h)
    if a == ACT_SOFTPLUS:
        s = (1 / (1 + np.exp(-pre))).astype(h.dtype)
        return s * (1 - s)
    raise ValueError(a)


class MLP:
    def __init__(self, dims, acts, pre=PRE_NONE, dtype=np.float32):
        self.dims = list(dims)
        self.acts = list(acts)
        assert len(self.acts) == len(self.dims) - 1
        self.pre = pre
        self.dtype = dtype
        self.offsets = []
        o = 0
        for i, n in zip(self.dims[:-1], self.dims[1:]):
            self.offsets.append((o, o + i * n))
            o += i * n + n
        self.nparams = o

    def unpack(self, p):
        out = []
        for (ow, ob), i, n in zip(self.offsets, self.dims[:-1], self.dims[1:]):
            out.append((p[ow:ob].reshape(i, n), p[ob:ob + n]))
        return out

    def init(self, rng, scale=1.0):
        """Glorot-ish deterministic init for tests / synthetic benchmarks."""
        p = np.zeros(self.nparams, dtype=self.dtype)
        for (ow, ob), i, n in zip(self.offsets, self.dims[:-1], self.dims[1:]):
            p[ow:ob] = (rng.standard_normal(i * n) * scale / np.sqrt(i)).astype(self.dtype)
            p[ob:ob + n] = (rng.standard_normal(n) * 0.1 * scale).astype(self.dtype)
        return p

    # forward: returns output and the cache needed by vjp
    def forward(self, p, x):
        Wb = self.unpack(p)
        x = np.asarray(x, dtype=self.dtype)
        h0 = x * x * x if self.pre == PRE_CUBE else x
        hs, pres = [h0], []
        h = h0
        for (W, b), a in zip(Wb, self.acts):
            pre = (h @ W + b).astype(self.dtype, copy=False)
            h = _act(a, pre).astype(self.dtype, copy=False)
            pres.append(pre)
            hs.append(h)
        return h, (x, hs, pres, Wb)

    def __call__(self, p, x):
        return self.forward(p, x)[0]

    def vjp(self, cache, ybar, want_dp=True):
        """(xbar, pbar) = ybar^T d(model)/d(x, p); pbar summed over the batch."""
        x, hs, pres, Wb = cache
        pbar = np.zeros(self.nparams, dtype=self.dtype) if want_dp else None
        hbar = np.asarray(ybar, dtype=self.dtype)
        for l in range(len(Wb) - 1, -1, -1):
            W, _ = Wb[l]
            d = (hbar * _dact(self.acts[l], pres[l], hs[l + 1])).astype(self.dtype, copy=False)
            if want_dp:
                ow, ob = self.offsets[l]
                pbar[ow:ob] = (hs[l].T @ d).reshape(-1)
                pbar[ob:ob + self.dims[l + 1]] = d.sum(axis=0)
            hbar = (d @ W.T).astype(self.dtype, copy=False)
        if self.pre == PRE_CUBE:
            hbar = hbar * (3 * x * x)
        return hbar.astype(self.dtype, copy=False), pbar

    # ---- second order: gradient of Phi = <lz, m(z)> - ll * <e, J(z) e> wrt (z, p) -------------
    def jvp_forward(self, p, z, e):
        """h_L, hdot_L with hdot_0 = e (directional derivative of the model along e)."""
        assert self.pre == PRE_NONE, "second-order path supports plain Dense chains only"
        Wb = self.unpack(p)
        h = np.asarray(z, dtype=self.dtype)
        hd = np.asarray(e, dtype=self.dtype)
        hs, hds, pres, qs = [h], [hd], [], []
        for (W, b), a in zip(Wb, self.acts):
            pre = (h @ W + b).astype(self.dtype, copy=False)
            q = (hd @ W).astype(self.dtype, copy=False)
            h = _act(a, pre).astype(self.dtype, copy=False)
            hd = (_dact(a, pre, h) * q).astype(self.dtype, copy=False)
            pres.append(pre)
            qs.append(q)
            hs.append(h)
            hds.append(hd)
        return h, hd, (hs, hds, pres, qs, Wb)

    def second_order_vjp(self, cache, hbar_L, hdbar_L, want_dp=True):
        """Reverse pass through the doubled (h, hdot) network with seeds on h_L and hdot_L."""
        hs, hds, pres, qs, Wb = cache
        pbar = np.zeros(self.nparams, dtype=self.dtype) if want_dp else None
        hbar = np.asarray(hbar_L, dtype=self.dtype)
        hdbar = np.asarray(hdbar_L, dtype=self.dtype)
        for l in range(len(Wb) - 1, -1, -1):
            W, _ = Wb[l]
            a = self.acts[l]
            s1 = _dact(a, pres[l], hs[l + 1])
            s2 = _d2act(a, pres[l], hs[l + 1])
            qbar = (hdbar * s1).astype(self.dtype, copy=False)
            prebar = (hbar * s1 + hdbar * qs[l] * s2).astype(self.dtype, copy=False)
            if want_dp:
                ow, ob = self.offsets[l]
                pbar[ow:ob] = (hs[l].T @ prebar + hds[l].T @ qbar).reshape(-1)
                pbar[ob:ob + self.dims[l + 1]] = prebar.sum(axis=0)
            hbar = (prebar @ W.T).astype(self.dtype, copy=False)
            hdbar = (qbar @ W.T).astype(self.dtype, copy=False)
        return hbar, pbar
