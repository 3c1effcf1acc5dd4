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


def _tf32_rn(x):
    """round-to-nearest (ties away) to the 10-bit tf32 mantissa: csrc/tc_common.cuh tf32_rn."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    return ((u + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def _split_tf32(x):
    hi = _tf32_rn(x)
    lo = _tf32_rn((np.asarray(x, np.float32) - hi).astype(np.float32))
    return hi.astype(np.float64), lo.astype(np.float64)


def _mm_3xtf32(a, b):
    """Model of the tcgen05 engine's 3xTF32 GEMM: operands split into hi/lo tf32, lo*hi + hi*lo + hi*hi summed without
    intermediate rounding (fp64 here; the tensor core's fp32 accumulator rounds far less often than an FMA chain),
    one rounding to fp32 at the end.  The lo*lo term is dropped as on the GPU (relative error ~2^-22 per product)."""
    ah, al = _split_tf32(a)
    bh, bl = _split_tf32(b)
    return (al @ bh + ah @ bl + ah @ bh).astype(np.float32)


class MLP:
    def __init__(self, dims, acts, pre=PRE_NONE, dtype=np.float32, wgrad_accum="blas", gemm_accum="blas"):
        self.dims = list(dims)
        self.acts = list(acts)
        assert len(self.acts) == len(self.dims) - 1
        self.pre = pre
        self.dtype = dtype
        # How the batch contraction of the parameter gradient (K = batch) is accumulated.  The reference
        # leaves this to its BLAS (sgemm blocking/threading is implementation-defined), so it is a named
        # switch:
        #   "blas"  one fp32 sgemm over the whole batch (what Zygote's `a' * d` does on the CPU)
        #   "fp64"  exact: fp64 accumulation, rounded to fp32 once
        #   ("tile", n)  fp32 sgemm per n-trajectory tile, tiles summed in fp64 in order, rounded once
        #   ("tile_defer", n)  the accumulation order of the CUDA path: vjp returns the per-tile fp32 partials
        #           [ntiles, P]; the integrator combines the stages per tile in fp32 and only then sums the
        #           tiles in fp64 (oracle/tsit5.py, "deferred-sum blocks")
        self.wgrad_accum = wgrad_accum
        # Accumulation of the layer GEMMs (forward h W and cotangent d W^T): "blas" = fp32 sgemm (CPU reference and
        # the FP32-FFMA engine); "fp64" = products and sums exact to fp64, rounded to fp32 once per output -- the
        # model of the tcgen05 engine, whose 3xTF32 products carry 2^-22 relative error but whose accumulator does
        # not round after every FMA (less rounding NOISE between neighbouring stage evaluations than an fp32 chain).
        self.gemm_accum = gemm_accum
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
            pre = (self._mm(h, W) + b).astype(self.dtype, copy=False)
            h = _act(a, pre).astype(self.dtype, copy=False)
            pres.append(pre)
            hs.append(h)
        return h, (x, hs, pres, Wb)

    def _mm(self, a, b):
        if self.dtype != np.float32 or self.gemm_accum == "blas":
            return a @ b
        if self.gemm_accum == "fp64":
            return (a.astype(np.float64) @ b.astype(np.float64)).astype(np.float32)
        assert self.gemm_accum == "3xtf32"
        return _mm_3xtf32(a, b)

    def __call__(self, p, x):
        return self.forward(p, x)[0]

    def vjp(self, cache, ybar, want_dp=True):
        """(xbar, pbar) = ybar^T d(model)/d(x, p); pbar summed over the batch."""
        x, hs, pres, Wb = cache
        defer = isinstance(self.wgrad_accum, tuple) and self.wgrad_accum[0] == "tile_defer"
        if defer and want_dp:
            nt = -(-x.shape[0] // self.wgrad_accum[1])
            pbar = np.zeros((nt, self.nparams), dtype=self.dtype)
        else:
            pbar = np.zeros(self.nparams, dtype=self.dtype) if want_dp else None
        hbar = np.asarray(ybar, dtype=self.dtype)
        for l in range(len(Wb) - 1, -1, -1):
            W, _ = Wb[l]
            d = (hbar * _dact(self.acts[l], pres[l], hs[l + 1])).astype(self.dtype, copy=False)
            if want_dp:
                ow, ob = self.offsets[l]
                gw, gb = self._wgrad(hs[l], d)
                pbar[..., ow:ob] = gw.reshape(gw.shape[:-2] + (-1,))
                pbar[..., ob:ob + self.dims[l + 1]] = gb
            hbar = self._mm(d, W.T).astype(self.dtype, copy=False)
        if self.pre == PRE_CUBE:
            hbar = hbar * (3 * x * x)
        return hbar.astype(self.dtype, copy=False), pbar

    def _wgrad(self, a, d):
        """(a^T d, column sums of d) over the batch with the configured accumulation (see __init__)."""
        mode = self.wgrad_accum
        if mode == "blas" or self.dtype != np.float32:
            return a.T @ d, d.sum(axis=0)
        if mode == "fp64":
            a64, d64 = a.astype(np.float64), d.astype(np.float64)
            return (a64.T @ d64).astype(np.float32), d64.sum(axis=0).astype(np.float32)
        kind, n = mode
        if kind == "tile_defer":
            nt = -(-a.shape[0] // n)
            mm = _mm_3xtf32 if self.gemm_accum == "3xtf32" else (lambda x, y: x @ y)
            gw = np.stack([mm(np.ascontiguousarray(a[i * n:(i + 1) * n].T), d[i * n:(i + 1) * n]) for i in range(nt)])
            gb = np.stack([d[i * n:(i + 1) * n].sum(axis=0, dtype=np.float32) for i in range(nt)])
            return gw, gb
        assert kind == "tile"
        gw = np.zeros((a.shape[1], d.shape[1]), np.float64)
        gb = np.zeros(d.shape[1], np.float64)
        for r0 in range(0, a.shape[0], n):
            at, dt = a[r0:r0 + n], d[r0:r0 + n]
            gw += (at.T @ dt).astype(np.float64)
            gb += dt.sum(axis=0, dtype=np.float32).astype(np.float64)
        return gw.astype(np.float32), gb.astype(np.float32)

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
