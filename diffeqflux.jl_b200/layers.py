"""Host-side mirror of DiffEqFlux's Lux-layer API for the NeuralODE hot path.

Same names, argument meaning and error behaviour as the reference (cited per class); every call
does one C-ABI call into libb200node.so.  Julia is not available in this image, so this Python
layer plays the role of the `ccall` glue (diffeqflux.jl_b200/julia/B200DiffEqFlux.jl shows the
Julia side a maintainer would add).

Array convention: a Julia (D, B) column-major matrix is passed here as a (B, D) row-major array --
the bytes are identical.  `Array(sol)` (D, B, T) is therefore (T, B, D) here.  Parameters are the
flat ComponentArray [vec(W1); b1; vec(W2); b2; ...] with W (out, in) column-major.

Tensors may be numpy arrays (host pointers, no autograd; use `.pullback`) or torch tensors
(CUDA tensors go to the library as device pointers; autograd is wired through
torch.autograd.Function, the stand-in for the ChainRules `rrule` Zygote calls in the reference).
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field

import numpy as np

from . import _native as N

try:  # torch is plumbing here (device memory, autograd tape, torch.distributed), not the product
    import torch
except Exception:  # pragma: no cover
    torch = None


# --------------------------------------------------------------------------------------- activations
def identity(x):
    return x


def tanh(x):
    return np.tanh(x)


def relu(x):
    return np.maximum(x, 0)


def sigmoid(x):
    return 1 / (1 + np.exp(-x))


def softplus(x):
    return np.log1p(np.exp(x))


def cube(x):
    """`x -> x .^ 3`, the pre-layer of docs/src/examples/neural_ode.md:32."""
    return x ** 3


_ACT_CODES = {identity: N.ACT_IDENTITY, tanh: N.ACT_TANH, relu: N.ACT_RELU, sigmoid: N.ACT_SIGMOID,
              softplus: N.ACT_SOFTPLUS, None: N.ACT_IDENTITY,
              "identity": N.ACT_IDENTITY, "tanh": N.ACT_TANH, "relu": N.ACT_RELU,
              "sigmoid": N.ACT_SIGMOID, "softplus": N.ACT_SOFTPLUS}


class UnsupportedError(NotImplementedError):
    """Raised for configurations outside the B200 hot path (never a silent fallback)."""


# --------------------------------------------------------------------------------------- Lux subset
class Dense:
    """Lux `Dense(in => out, activation)`: y = act.(W*x .+ b)."""

    def __init__(self, in_dims: int, out_dims: int, activation=identity):
        if activation not in _ACT_CODES:
            raise UnsupportedError(f"activation {activation!r}: supported are identity, tanh, relu, sigmoid, softplus")
        self.in_dims, self.out_dims, self.activation = int(in_dims), int(out_dims), activation

    @property
    def act_code(self):
        return _ACT_CODES[self.activation]


class Chain:
    """Lux `Chain(layers...)` restricted to [cube,] Dense, Dense, ... (the models of the hot path)."""

    def __init__(self, *layers):
        layers = list(layers)
        self.pre = N.PRE_NONE
        if layers and layers[0] is cube:
            self.pre = N.PRE_CUBE
            layers = layers[1:]
        if not layers or not all(isinstance(l, Dense) for l in layers):
            raise UnsupportedError("model must be Chain([x -> x.^3,] Dense, Dense, ...): other Lux layers are outside the B200 path")
        for a, b in zip(layers[:-1], layers[1:]):
            if a.out_dims != b.in_dims:
                raise ValueError(f"DimensionMismatch: Dense({a.in_dims} => {a.out_dims}) then Dense({b.in_dims} => {b.out_dims})")
        self.layers = layers
        self.dims = [layers[0].in_dims] + [l.out_dims for l in layers]
        self.acts = [l.act_code for l in layers]

    @property
    def nparams(self):
        return sum(i * o + o for i, o in zip(self.dims[:-1], self.dims[1:]))

    def initialparameters(self, rng) -> np.ndarray:
        """Flat ComponentArray: glorot-uniform weights, zero biases (Lux defaults)."""
        out = []
        for i, o in zip(self.dims[:-1], self.dims[1:]):
            lim = math.sqrt(6.0 / (i + o))
            out.append(rng.uniform(-lim, lim, size=i * o).astype(np.float32))
            out.append(np.zeros(o, np.float32))
        return np.concatenate(out)

    def axes(self):
        """ComponentArray axes: {layer_k: {weight: slice, bias: slice}}."""
        ax, off = {}, 0
        for k, (i, o) in enumerate(zip(self.dims[:-1], self.dims[1:]), 1):
            ax[f"layer_{k}"] = {"weight": (slice(off, off + i * o), (o, i)), "bias": (slice(off + i * o, off + i * o + o), (o,))}
            off += i * o + o
        return ax

    def view(self, ps, layer: int, name: str):
        """`ps.layer_k.weight` as an (out, in) view / `ps.layer_k.bias`."""
        sl, shape = self.axes()[f"layer_{layer}"][name]
        v = ps[sl]
        return v.reshape(shape[1], shape[0]).T if name == "weight" else v


def setup(rng, layer):
    """`Lux.setup(rng, layer) -> (ps, st)`; wrapper layers expose the wrapped model's parameters
    (NeuralDELayer <: AbstractLuxWrapperLayer{:model}, src/neural_de.jl:1)."""
    model = layer.model if hasattr(layer, "model") else layer
    ps = model.initialparameters(rng)
    st = layer.initialstates(rng) if hasattr(layer, "initialstates") else {}
    return ps, st


# --------------------------------------------------------------------------------------- algorithms
class Tsit5:
    """OrdinaryDiffEq `Tsit5()` -- the only solver of the B200 path."""


@dataclass
class InterpolatingAdjoint:
    autojacvec: object = None
    checkpointing: bool = False
    code: int = N.SENSE_INTERPOLATING


@dataclass
class BacksolveAdjoint:
    autojacvec: object = None
    checkpointing: bool = True
    code: int = N.SENSE_BACKSOLVE


@dataclass
class QuadratureAdjoint:
    """`QuadratureAdjoint(; abstol = 1e-6, reltol = 1e-3)` (re-exported at src/DiffEqFlux.jl:19-20): lambda alone is
    integrated backwards against the dense forward output; dp = int lambda^T df/dp dt is a quadrature.  On the B200 path
    the quadrature uses the Runge-Kutta weights of the lambda solve's own stages (error O(solver tolerance)); `abstol` /
    `reltol` of the reference's adaptive Gauss-Kronrod rule are accepted and ignored."""
    abstol: float = 1e-6
    reltol: float = 1e-3
    autojacvec: object = None
    code: int = N.SENSE_QUADRATURE


@dataclass
class GaussAdjoint:
    """`GaussAdjoint()` (src/DiffEqFlux.jl:19-20): same engine as QuadratureAdjoint on the B200 path."""
    autojacvec: object = None
    checkpointing: bool = False
    code: int = N.SENSE_GAUSS


class ZygoteVJP:
    pass


class ReturnCode:
    Success, MaxIters, DtLessThanMin, Unstable = 0, 1, 2, 3


def successful_retcode(rc) -> bool:
    return rc == 0


# --------------------------------------------------------------------------------------- runtime plumbing
_ctx_cache: dict = {}
_model_cache: dict = {}


def _context(device: int, stream: int = 0) -> N.Context:
    key = (device, stream)
    if key not in _ctx_cache:
        _ctx_cache[key] = N.Context(device, stream or None)
    return _ctx_cache[key]


def _model_handle(ctx: N.Context, chain: Chain):
    key = (id(ctx), tuple(chain.dims), tuple(chain.acts), chain.pre)
    if key not in _model_cache:
        h = C.c_void_p()
        dims = (C.c_int * len(chain.dims))(*chain.dims)
        acts = (C.c_int * len(chain.acts))(*chain.acts)
        ctx.check(N.lib().b200_model_create(ctx._h, len(chain.acts), dims, acts, chain.pre, C.byref(h)))
        _model_cache[key] = h
    return _model_cache[key]


def _is_torch(x):
    return torch is not None and isinstance(x, torch.Tensor)


class _Buf:
    """A float32 buffer the C ABI can read: numpy (host) or torch CUDA (device)."""

    def __init__(self, x, like=None):
        if _is_torch(x):
            if x.dtype != torch.float32:
                raise UnsupportedError(f"eltype {x.dtype}: the B200 path is Float32 only")
            self.t = x.detach().contiguous()
            self.dev = self.t.is_cuda
            if not self.dev:
                self.a = self.t.numpy()
            self.ptr = self.t.data_ptr()
        else:
            a = np.asarray(x)
            if a.dtype == np.float64 and like is None:
                raise UnsupportedError("eltype Float64: the B200 path is Float32 only")
            self.a = np.ascontiguousarray(a, dtype=np.float32)
            self.dev = False
            self.ptr = self.a.ctypes.data


def _pick_device(*xs):
    for x in xs:
        if _is_torch(x) and x.is_cuda:
            return x.device.index if x.device.index is not None else torch.cuda.current_device()
    if torch is not None and torch.cuda.is_available():
        return torch.cuda.current_device()       # host (numpy) buffers: the process's current CUDA device
    return 0


def _alloc(shape, device_tensor_like, on_device, device):
    if on_device:
        return torch.empty(shape, dtype=torch.float32, device=f"cuda:{device}")
    return np.empty(shape, np.float32)


def _ptr(x):
    if x is None:
        return None
    return x.data_ptr() if _is_torch(x) else x.ctypes.data


@dataclass
class SolveStats:
    nf: int = 0
    naccept: int = 0
    nreject: int = 0
    kernel_ms: float = 0.0
    tie_dist: float = float("inf")
    grid: int = 0
    tile: int = 0
    engine: int = 0


class ODESolution:
    """Minimal `ODESolution`: `.t`, `.u` (list of (B, D) states), `.retcode`, `.stats`,
    `Array(sol)` = `sol.array()` of shape (T, B, D), `sol[i]`, `sol[-1]`, `len(sol)`."""

    def __init__(self, t, arr, retcode, stats, vector_input=False):
        self.t = t
        self._arr = arr
        self.retcode = retcode
        self.stats = stats
        self._vec = vector_input

    def array(self):
        return self._arr[:, 0, :] if self._vec else self._arr

    @property
    def u(self):
        a = self.array()
        return [a[i] for i in range(a.shape[0])]

    def __getitem__(self, i):
        return self.array()[i]

    def __len__(self):
        return self._arr.shape[0]

    def __iter__(self):
        return iter(self.u)


def Array(sol: ODESolution):
    return sol.array()


_SOLVE_KW = {"saveat", "abstol", "reltol", "save_everystep", "save_start", "save_end", "sensealg",
             "dt", "maxiters", "tile", "max_dense_steps", "dense", "path"}


def _split_saveat(saveat, t0, t1, save_start, save_end, save_everystep):
    """OrdinaryDiffEq conventions (SURVEY.md section 8(c) 'dense output / saveat')."""
    f32 = np.float32
    t0, t1 = f32(t0), f32(t1)
    scalar = False
    if saveat is None:
        sv = np.zeros(0, np.float32)
    elif np.isscalar(saveat):
        h = f32(abs(saveat))
        n = int(np.floor(float(abs(t1 - t0) / h) + 1e-6))
        sv = (t0 + (h if t1 >= t0 else -h) * np.arange(0, n + 1, dtype=np.float32)).astype(np.float32)   # t0:h:t1 in either direction
        scalar = True
    else:
        sv = np.asarray(saveat.detach().cpu().numpy() if _is_torch(saveat) else saveat, dtype=np.float32).reshape(-1)
    if save_everystep is None:
        save_everystep = sv.size == 0
    if save_start is None:
        save_start = bool(save_everystep or sv.size == 0 or scalar or (t0 in sv))
    if save_end is None:
        save_end = bool(save_everystep or sv.size == 0 or scalar or (t1 in sv))
    lo, hi = min(t0, t1), max(t0, t1)
    interior = np.sort(np.array([s for s in sv if lo < s < hi], dtype=np.float32))   # saveat is a heap upstream
    if t1 < t0:
        interior = interior[::-1].copy()
    return interior, bool(save_start), bool(save_end), bool(save_everystep) and sv.size == 0


class _Solve:
    """One forward solve through the C ABI plus (optionally) its tape."""

    def __init__(self, chain: Chain, tspan, args, kwargs, augment=0):
        if len(args) == 0 or args[0] is None:
            raise UnsupportedError("alg = nothing (DifferentialEquations default algorithm): pass Tsit5()")
        if not isinstance(args[0], Tsit5) and args[0] is not Tsit5:
            raise UnsupportedError(f"solver {type(args[0]).__name__}: only Tsit5() runs on the B200 path")
        if len(args) > 1:
            raise UnsupportedError("extra positional solve arguments are not supported")
        bad = set(kwargs) - _SOLVE_KW
        if bad:
            raise UnsupportedError(f"unsupported solve keyword(s) {sorted(bad)}")
        self.chain, self.tspan, self.kw, self.augment = chain, tspan, dict(kwargs), augment
        sa = self.kw.get("sensealg", InterpolatingAdjoint(autojacvec=ZygoteVJP()))   # src/neural_de.jl:57
        if isinstance(sa, type):
            sa = sa()
        if not isinstance(sa, (InterpolatingAdjoint, BacksolveAdjoint, QuadratureAdjoint, GaussAdjoint)):
            raise UnsupportedError(f"sensealg {type(sa).__name__}: supported are Interpolating, Backsolve, Quadrature and Gauss adjoints")
        self.sense = sa.code

    def opts(self, on_device, want_tape):
        o = N.default_opts()
        o.abstol = float(self.kw.get("abstol", 1e-6))
        o.reltol = float(self.kw.get("reltol", 1e-3))
        o.dt0 = float(self.kw.get("dt", 0.0))
        o.maxiters = int(self.kw.get("maxiters", 100000))
        o.tile = int(self.kw.get("tile", 0))
        o.max_dense_steps = int(self.kw.get("max_dense_steps", 0))
        o.path = int(self.kw.get("path", 0))
        o.ptr_kind = N.PTR_DEVICE if on_device else N.PTR_HOST
        o.sensealg = self.sense if want_tape else N.SENSE_NONE
        return o

    def forward(self, x, p, want_tape):
        t0, t1 = self.tspan
        xb, pb = _Buf(x), _Buf(p)
        vec = (xb.t.dim() if _is_torch(x) else xb.a.ndim) == 1
        shape = tuple(xb.t.shape) if _is_torch(x) else xb.a.shape
        B, d_in = (1, shape[0]) if vec else shape
        D = d_in + self.augment
        if D != self.chain.dims[0]:
            raise ValueError(f"DimensionMismatch: state has {D} rows, model expects {self.chain.dims[0]}")
        on_dev = xb.dev
        if pb.dev != on_dev:
            raise ValueError("u0 and p must live on the same device")
        dev = _pick_device(x, p)
        ctx = _context(dev)
        mh = _model_handle(ctx, self.chain)
        interior, ss, se, sev = _split_saveat(self.kw.get("saveat"), t0, t1, self.kw.get("save_start"),
                                              self.kw.get("save_end"), self.kw.get("save_everystep"))
        o = self.opts(on_dev, want_tape)
        o.save_start, o.save_end, o.save_everystep = int(ss), int(se), int(sev)
        if sev and o.sensealg == N.SENSE_BACKSOLVE:
            raise UnsupportedError("save_everystep=true with BacksolveAdjoint: pass saveat or use InterpolatingAdjoint")
        st = N.Stats()
        tape = C.c_void_p()
        sv = interior
        if on_dev:
            sv_t = torch.from_numpy(interior).to(f"cuda:{dev}") if interior.size else None
            sv_ptr = sv_t.data_ptr() if sv_t is not None else None
        else:
            sv_ptr = sv.ctypes.data if sv.size else None
        n_out = int(ss) + len(interior) + int(se)
        out_u = None if sev else _alloc((n_out, B, D), x, on_dev, dev)
        out_t = None if sev else _alloc((n_out,), x, on_dev, dev)
        need_tape = want_tape or sev
        rc = ctx.check(N.lib().b200_node_forward(
            ctx._h, mh, pb.ptr, xb.ptr, d_in, self.augment, B, float(t0), float(t1), sv_ptr, len(interior),
            C.byref(o), _ptr(out_u), _ptr(out_t), C.byref(st), C.byref(tape) if need_tape else None))
        if sev:
            n = N.lib().b200_tape_nsteps(tape)
            ts = _alloc((n + 1,), x, on_dev, dev)
            us = _alloc((n + 1, B, D), x, on_dev, dev)
            ctx.check(N.lib().b200_tape_steps(ctx._h, tape, _ptr(ts), _ptr(us), o.ptr_kind))
            lo = 0 if ss else 1
            hi = n + 1 if (se or n == 0) else n
            out_u, out_t = us[lo:hi], ts[lo:hi]
            if not want_tape:
                N.lib().b200_tape_free(tape)
                tape = C.c_void_p()
        elif st.n_out < n_out:
            # an early exit (MaxIters / DtLessThanMin / Unstable, or t0 == t1) wrote fewer points than requested:
            # only the rows the kernel wrote are returned, as `sol.u` of an aborted solve upstream
            out_u, out_t = out_u[:st.n_out], out_t[:st.n_out]
        stats = SolveStats(st.nf, st.naccept, st.nreject, st.kernel_ms, st.tie_dist, st.grid, st.tile, st.engine)
        return dict(ctx=ctx, mh=mh, tape=tape, out_u=out_u, out_t=out_t, rc=rc, stats=stats, vec=vec,
                    B=B, D=D, d_in=d_in, on_dev=on_dev, dev=dev, opts=o)

    def adjoint(self, fw, dL):
        """(du0, dp) from the cotangent of the saved states; consumes nothing (tape reusable)."""
        ctx, o = fw["ctx"], fw["opts"]
        dLb = _Buf(dL)
        B, d_in = fw["B"], fw["d_in"]
        du0 = _alloc((B, d_in), None, fw["on_dev"], fw["dev"])
        dp = _alloc((self.chain.nparams,), None, fw["on_dev"], fw["dev"])
        st = N.Stats()
        rc = ctx.check(N.lib().b200_node_adjoint(ctx._h, fw["mh"], fw["tape"], dLb.ptr, C.byref(o), _ptr(du0),
                                                 _ptr(dp), C.byref(st)))
        stats = SolveStats(st.nf, st.naccept, st.nreject, st.kernel_ms, st.tie_dist, st.grid, st.tile, st.engine)
        if fw["vec"]:
            du0 = du0[0]
        return du0, dp, rc, stats


class _TapeOwner:
    """Frees the C tape when the last Python reference disappears."""

    def __init__(self, tape):
        self.tape = tape

    def __del__(self):
        try:
            if self.tape:
                N.lib().b200_tape_free(self.tape)
        except Exception:
            pass


if torch is not None:
    class _NodeFn(torch.autograd.Function):
        """The rrule of `solve` (reference: SciMLSensitivity via src/DiffEqFlux.jl:17-22)."""

        @staticmethod
        def forward(ctx, x, p, solve, holder):
            fw = solve.forward(x, p, want_tape=True)
            holder.update(fw)
            ctx.solve, ctx.fw, ctx.owner = solve, fw, _TapeOwner(fw["tape"])
            ctx.x_is_host = not x.is_cuda
            out = fw["out_u"]
            if not _is_torch(out):
                out = torch.from_numpy(out)
            return out

        @staticmethod
        def backward(ctx, g):
            fw = ctx.fw
            g = g.contiguous()
            if not fw["on_dev"]:
                g = g.cpu().numpy()
            du0, dp, rc, stats = ctx.solve.adjoint(fw, g)
            fw["adjoint_stats"] = stats
            fw["adjoint_rc"] = rc
            if not _is_torch(du0):
                du0, dp = torch.from_numpy(np.ascontiguousarray(du0)), torch.from_numpy(dp)
            return du0, dp, None, None


# --------------------------------------------------------------------------------------- NeuralODE
class NeuralODE:
    """`NeuralODE(model, tspan, alg, args...; kwargs...)` (src/neural_de.jl:35-45) called as
    `(x, p, st) -> (sol, st)` (src/neural_de.jl:47-61)."""

    def __init__(self, model, tspan, *args, **kwargs):
        if not isinstance(model, Chain):
            raise UnsupportedError("model must be a Chain of Dense layers")
        self.model, self.tspan, self.args, self.kwargs = model, tuple(tspan), args, kwargs
        self._solve = _Solve(model, self.tspan, args, kwargs, augment=getattr(self, "_augment", 0))
        self.last = {}

    def __call__(self, x, p, st=None):
        st = {} if st is None else st
        needs_grad = _is_torch(x) and _is_torch(p) and torch.is_grad_enabled() and (x.requires_grad or p.requires_grad)
        if needs_grad:
            self.last = {}
            arr = _NodeFn.apply(x, p, self._solve, self.last)
            fw = self.last
        else:
            fw = self._solve.forward(x, p, want_tape=False)
            self.last = fw
            arr = fw["out_u"]
        t = fw["out_t"]
        return ODESolution(t, arr, fw["rc"], fw["stats"], fw["vec"]), st

    def pullback(self, x, p, st=None):
        """Zygote.pullback-style: returns (sol, back) with back(dL_du) -> (du0, dp)."""
        fw = self._solve.forward(x, p, want_tape=True)
        owner = _TapeOwner(fw["tape"])
        self.last = fw
        sol = ODESolution(fw["out_t"], fw["out_u"], fw["rc"], fw["stats"], fw["vec"])

        def back(dL):
            _ = owner
            if fw["vec"]:
                dL = dL.reshape((dL.shape[0], 1, -1))
            du0, dp, rc, stats = self._solve.adjoint(fw, dL)
            fw["adjoint_stats"], fw["adjoint_rc"] = stats, rc
            return du0, dp

        return sol, back


class AugmentedNDELayer:
    """`AugmentedNDELayer(nde, adim)` (src/neural_de.jl:372-386): zero-pad `adim` state rows before the
    NDE (the padding is fused into the kernel's initial-state load; no gradient flows into the zeros)."""

    def __init__(self, nde: NeuralODE, adim: int):
        if not isinstance(nde, NeuralODE):
            raise UnsupportedError("AugmentedNDELayer wraps a NeuralODE on the B200 path")
        self.nde, self.adim, self.model = nde, int(adim), nde.model
        self._inner = NeuralODE.__new__(NeuralODE)
        self._inner._augment = self.adim
        NeuralODE.__init__(self._inner, nde.model, nde.tspan, *nde.args, **nde.kwargs)

    def __call__(self, x, p, st=None):
        return self._inner(x, p, st)

    def pullback(self, x, p, st=None):
        return self._inner.pullback(x, p, st)

    @property
    def last(self):
        return self._inner.last


# --------------------------------------------------------------------------------------- FFJORD
if torch is not None:
    class _FfjordFn(torch.autograd.Function):
        """(logpx, end state) of the FFJORD solve; both outputs carry gradients (lambda_1 / lambda_2 of a regularize = true
        layer are slices of the end state)."""

        @staticmethod
        def forward(ctx, x, p, layer, seed, holder):
            seed, exact, regularize = seed
            fw = layer._forward(x, p, seed, want_tape=True, exact=exact, regularize=regularize)
            holder.update(fw)
            ctx.layer, ctx.fw, ctx.owner = layer, fw, _TapeOwner(fw["tape"])
            out, z = fw["logpx"], fw["z"]
            if not _is_torch(out):
                out, z = torch.from_numpy(out), torch.from_numpy(z)
            return out, z.clone()

        @staticmethod
        def backward(ctx, g, gz):
            fw = ctx.fw
            g = g.contiguous().reshape(-1)
            gz = gz.contiguous()
            if not fw["on_dev"]:
                g, gz = g.cpu().numpy(), gz.cpu().numpy()
            dx, dp = ctx.layer._adjoint(fw, g, gz)
            if not _is_torch(dx):
                dx, dp = torch.from_numpy(dx), torch.from_numpy(dp)
            return dx, dp, None, None, None


class FFJORD:
    """`FFJORD(model, tspan, input_dims, args...; ad, basedist, kwargs...)` (src/ffjord.jl:52-73),
    called as `(x, ps, st) -> ((logpx, l1, l2), st)` (src/ffjord.jl:138-178).  `st["monte_carlo"]` picks the Hutchinson
    estimator or the exact trace, `st["regularize"]` adds the two regulariser rows to the state (src/ffjord.jl:61-66);
    ad in (nothing, AutoZygote) and basedist = nothing; anything else raises."""

    def __init__(self, model, tspan, input_dims, *args, ad=None, basedist=None, **kwargs):
        if not isinstance(model, Chain):
            raise UnsupportedError("model must be a Chain of Dense layers")
        if basedist is not None:
            raise UnsupportedError("basedist: only the unit normal (basedist = nothing) is supported")
        if ad not in (None, "AutoZygote"):
            raise UnsupportedError(f"ad = {ad}: only nothing / AutoZygote (VJP Hutchinson estimator)")
        self.model, self.tspan, self.input_dims, self.args = model, tuple(tspan), input_dims, args
        self.kwargs = kwargs
        bad = set(kwargs) - (_SOLVE_KW | {"seed", "noise_per_call"})
        if bad:
            raise UnsupportedError(f"unsupported solve keyword(s) {sorted(bad)}")
        kw = {k: v for k, v in kwargs.items() if k in _SOLVE_KW}
        self._solve = _Solve(model, self.tspan, args, kw)
        self.last = {}

    def initialstates(self, rng):
        return {"model": {}, "regularize": False, "monte_carlo": True,
                "seed": int(rng.integers(0, 2**62)) if hasattr(rng, "integers") else 0, "calls": 0}

    def _forward(self, x, p, seed, want_tape, exact=False, regularize=False):
        xb, pb = _Buf(x), _Buf(p)
        shape = tuple(xb.t.shape) if _is_torch(x) else xb.a.shape
        if len(shape) != 2 or shape[1] != self.model.dims[0]:
            raise ValueError(f"DimensionMismatch: x must be (B, {self.model.dims[0]})")
        B, D = shape
        on_dev = xb.dev
        dev = _pick_device(x, p)
        ctx = _context(dev)
        mh = _model_handle(ctx, self.model)
        o = self._solve.opts(on_dev, want_tape)
        o.seed = int(seed) & (2**64 - 1)
        o.noise_per_call = int(self.kwargs.get("noise_per_call", 1))
        o.ffjord_exact_trace = 1 if exact else 0        # monte_carlo = false (src/ffjord.jl:120-123)
        o.ffjord_regularize = 1 if regularize else 0    # src/ffjord.jl:131-133,149-153
        o.b_offset = _dist_state["rank"] * B            # global trajectory index of this shard (noise stream)
        if want_tape:
            o.sensealg = N.SENSE_INTERPOLATING      # hard-wired at src/ffjord.jl:143
        st = N.Stats()
        tape = C.c_void_p()
        NS = D + (3 if regularize else 1)
        logpx = _alloc((B,), x, on_dev, dev)
        z = _alloc((B, NS), x, on_dev, dev)
        t0, t1 = self.tspan
        rc = ctx.check(N.lib().b200_ffjord_forward(ctx._h, mh, pb.ptr, xb.ptr, B, float(t0), float(t1), C.byref(o),
                                                   _ptr(logpx), _ptr(z), C.byref(st),
                                                   C.byref(tape) if want_tape else None))
        stats = SolveStats(st.nf, st.naccept, st.nreject, st.kernel_ms, st.tie_dist, st.grid, st.tile, st.engine)
        return dict(ctx=ctx, mh=mh, tape=tape, logpx=logpx, z=z, rc=rc, stats=stats, B=B, D=D, NS=NS, on_dev=on_dev,
                    dev=dev, opts=o, regularize=bool(regularize))

    def _adjoint(self, fw, g, gz=None):
        """g: cotangent of logpx [B] (or None); gz: cotangent of the end state [B][NS] (or None)."""
        ctx, o = fw["ctx"], fw["opts"]
        gb = _Buf(g) if g is not None else None
        gzb = _Buf(gz) if gz is not None else None
        dx = _alloc((fw["B"], fw["D"]), None, fw["on_dev"], fw["dev"])
        dp = _alloc((self.model.nparams,), None, fw["on_dev"], fw["dev"])
        st = N.Stats()
        rc = ctx.check(N.lib().b200_ffjord_adjoint_state(ctx._h, fw["mh"], fw["tape"], gb.ptr if gb else None,
                                                         gzb.ptr if gzb else None, C.byref(o), _ptr(dx), _ptr(dp), C.byref(st)))
        fw["adjoint_stats"] = SolveStats(st.nf, st.naccept, st.nreject, st.kernel_ms, st.tie_dist, st.grid, st.tile, st.engine)
        fw["adjoint_rc"] = rc
        return dx, dp

    @staticmethod
    def _lambdas(z, D, regularize):
        """(delta_logp, lambda_1, lambda_2) out of the end state z [B][NS], literally as src/ffjord.jl:161-169 takes them:
        delta_logp = row D of every trajectory; with regularize the reference selects along the LAST (batch) axis,
        `selectdim(pred, N, (L-1):(L-1))` and `selectdim(pred, N, L:L)` with L = D + 3 -- the full state of trajectories
        L-1 and L (1-based), Julia shape (L, 1) = C-order (1, L) here.  (The regulariser integrals themselves are
        z[:, D+1] and z[:, D+2].)"""
        dlp = z[:, D].reshape(1, -1)
        if not regularize:
            return dlp, dlp, dlp
        L = D + 3
        if z.shape[0] < L:
            raise IndexError(f"BoundsError: regularize = true reads trajectories {L - 1} and {L} of a batch of {z.shape[0]}")
        return dlp, z[L - 2:L - 1, :], z[L - 1:L, :]

    def __call__(self, x, ps, st):
        regularize = bool(st.get("regularize", False))
        exact = not st.get("monte_carlo", True)
        seed = (int(st.get("seed", 0)) + 0x9E3779B97F4A7C15 * int(st.get("calls", 0))) & (2**64 - 1)
        needs_grad = _is_torch(x) and _is_torch(ps) and torch.is_grad_enabled() and (x.requires_grad or ps.requires_grad)
        if needs_grad:
            self.last = {}
            logpx, z = _FfjordFn.apply(x, ps, self, (seed, exact, regularize), self.last)
        else:
            self.last = self._forward(x, ps, seed, want_tape=False, exact=exact, regularize=regularize)
            logpx, z = self.last["logpx"], self.last["z"]
        logpx = logpx.reshape(1, -1)
        _, l1, l2 = self._lambdas(z, self.model.dims[0], regularize)
        st2 = dict(st)
        st2["calls"] = int(st.get("calls", 0)) + 1
        return (logpx, l1, l2), st2

    def pullback(self, x, ps, st):
        """-> (logpx, back);  back(g_logpx, g_z = None) -> (dx, dp), g_z the cotangent of the end state `self.last["z"]`
        ([B][D+1], or [B][D+3] with regularize: columns D+1, D+2 are the regulariser integrals)."""
        seed = (int(st.get("seed", 0)) + 0x9E3779B97F4A7C15 * int(st.get("calls", 0))) & (2**64 - 1)
        fw = self._forward(x, ps, seed, want_tape=True, exact=not st.get("monte_carlo", True),
                           regularize=bool(st.get("regularize", False)))
        owner = _TapeOwner(fw["tape"])
        self.last = fw

        def back(g, gz=None):
            _ = owner
            if g is not None:
                g = np.asarray(g, np.float32).reshape(-1) if not _is_torch(g) else g.reshape(-1)
            if gz is not None:
                gz = np.ascontiguousarray(gz, np.float32) if not _is_torch(gz) else gz.contiguous()
            return self._adjoint(fw, g, gz)

        return fw["logpx"], back

    def backward(self, n_samples: int, ps, st, rng=None, z=None):
        """`__backward_ffjord(T, n, n_samples, ps, st, rng)` (src/ffjord.jl:183-214): draw z ~ N(0, I) (or take `z`),
        integrate the same augmented RHS over reverse(tspan) and return the first D rows of the end state -- samples of
        the learned density.  One b200_ffjord_forward call with the time span reversed."""
        D = self.model.dims[0]
        if z is None:
            rng = np.random.default_rng() if rng is None else rng
            z = rng.standard_normal((int(n_samples), D)).astype(np.float32)
            if _is_torch(ps):
                z = torch.from_numpy(z).to(ps.device)
        seed = (int(st.get("seed", 0)) + 0x9E3779B97F4A7C15 * int(st.get("calls", 0))) & (2**64 - 1)
        rev = FFJORD(self.model, (self.tspan[1], self.tspan[0]), self.input_dims, *self.args, **self.kwargs)
        fw = rev._forward(z, ps, seed, want_tape=False, exact=not st.get("monte_carlo", True),
                          regularize=bool(st.get("regularize", False)))
        self.last = fw
        return fw["z"][:, :D]


class FFJORDDistribution:
    """`FFJORDDistribution(model, ps, st)` (src/ffjord.jl:215-245): `logpdf(d, x)` = first output of the forward pass,
    `rand(rng, d, n)` = `__backward_ffjord`."""

    def __init__(self, model: FFJORD, ps, st):
        self.model, self.ps, self.st = model, ps, st

    def __len__(self):
        return int(np.prod(self.model.input_dims))

    def logpdf(self, x):
        x2 = x.reshape(1, -1) if (x.dim() if _is_torch(x) else np.asarray(x).ndim) == 1 else x
        (logpx, _, _), _ = self.model(x2, self.ps, self.st)
        return logpx.reshape(-1)

    def rand(self, n: int = 1, rng=None):
        return self.model.backward(n, self.ps, self.st, rng)


# --------------------------------------------------------------------------------------- training step / DimMover
@dataclass
class AdamState:
    """State of `OptimizationOptimisers.Adam(lr)` (Optimisers.jl defaults beta = (0.9, 0.999), eps = 1e-8)."""
    lr: float = 0.05
    beta1: float = 0.9
    beta2: float = 0.999
    eps: float = 1e-8
    step: int = 0
    m: object = None
    v: object = None


def train_step(node: "NeuralODE", u0, p, data, opt: AdamState):
    """One iteration of the docs' training loop (docs/src/examples/neural_ode.md:36-44,64-70) as ONE C-ABI call
    (b200_node_train_step): forward solve at the layer's saveat, loss = sum(abs2, data .- pred), adjoint gradient and the
    Adam update of `p` IN PLACE.  Returns (loss before the update, retcode).  With CUDA tensors nothing but the loss
    scalar leaves the device."""
    solve = node._solve
    t0, t1 = solve.tspan
    xb, pb, db = _Buf(u0), _Buf(p), _Buf(data)
    if not (xb.dev == pb.dev == db.dev):
        raise ValueError("u0, p and data must live on the same device")
    if not _is_torch(p) and not (isinstance(p, np.ndarray) and p.dtype == np.float32 and p.flags.c_contiguous):
        raise ValueError("p is updated in place: pass a contiguous float32 array / tensor")
    if _is_torch(p) and not p.is_contiguous():
        raise ValueError("p is updated in place: pass a contiguous tensor")
    vec = (xb.t.dim() if _is_torch(u0) else xb.a.ndim) == 1
    shape = tuple(xb.t.shape) if _is_torch(u0) else xb.a.shape
    B, d_in = (1, shape[0]) if vec else shape
    on_dev = xb.dev
    dev = _pick_device(u0, p)
    ctx = _context(dev)
    mh = _model_handle(ctx, solve.chain)
    interior, ss, se, sev = _split_saveat(solve.kw.get("saveat"), t0, t1, solve.kw.get("save_start"), solve.kw.get("save_end"),
                                          solve.kw.get("save_everystep"))
    if sev:
        raise UnsupportedError("train_step needs saveat (or save_everystep = false): the data must have a fixed time grid")
    o = solve.opts(on_dev, True)
    o.save_start, o.save_end, o.save_everystep = int(ss), int(se), 0
    n_out = int(ss) + len(interior) + int(se)
    D = d_in + solve.augment
    dshape = tuple(db.t.shape) if _is_torch(data) else db.a.shape
    if int(np.prod(dshape)) != n_out * B * D:
        raise ValueError(f"DimensionMismatch: data has {int(np.prod(dshape))} entries, the prediction {n_out} x {B} x {D}")
    nP = solve.chain.nparams
    if opt.m is None:
        opt.m = torch.zeros(nP, dtype=torch.float32, device=p.device) if _is_torch(p) else np.zeros(nP, np.float32)
        opt.v = torch.zeros(nP, dtype=torch.float32, device=p.device) if _is_torch(p) else np.zeros(nP, np.float32)
    opt.step += 1
    if on_dev:
        sv_t = torch.from_numpy(interior).to(f"cuda:{dev}") if interior.size else None
        sv_ptr = sv_t.data_ptr() if sv_t is not None else None
    else:
        sv_ptr = interior.ctypes.data if interior.size else None
    loss = _alloc((1,), None, on_dev, dev)
    fst, ast = N.Stats(), N.Stats()
    rc = ctx.check(N.lib().b200_node_train_step(
        ctx._h, mh, _ptr(p), _ptr(opt.m), _ptr(opt.v), opt.step, float(opt.lr), float(opt.beta1), float(opt.beta2), float(opt.eps),
        xb.ptr, d_in, solve.augment, B, float(t0), float(t1), sv_ptr, len(interior), db.ptr, C.byref(o), _ptr(loss),
        C.byref(fst), C.byref(ast)))
    node.last = dict(stats=SolveStats(fst.nf, fst.naccept, fst.nreject, fst.kernel_ms, fst.tie_dist, fst.grid, fst.tile, fst.engine),
                     adjoint_stats=SolveStats(ast.nf, ast.naccept, ast.nreject, ast.kernel_ms, ast.tie_dist, ast.grid, ast.tile, ast.engine), rc=rc)
    return float(loss[0]), rc


class DimMover:
    """`DimMover(; from = -2, to = -1)` (src/neural_de.jl:397-412): `cat(eachslice(x; dims = from)...; dims = to)`.
    Arrays here are the C-order views of Julia arrays (shape reversed: `Array(sol)` (D,B,T) is (T,B,D)), and the layer
    returns the C-order view of the Julia result, so the default turns (T,B,D) into (B,T,D).  Runs on the device
    through b200_dimmover."""

    def __init__(self, from_: int = -2, to: int = -1, **kw):
        from_ = kw.pop("from", from_)
        if kw:
            raise TypeError(f"unexpected keyword(s) {sorted(kw)}")
        if from_ == 0 or to == 0:
            raise AssertionError("from !== 0 && to !== 0")
        self.from_, self.to = int(from_), int(to)

    def __call__(self, x, ps=None, st=None):
        xb = _Buf(x)
        shape = tuple(xb.t.shape) if _is_torch(x) else xb.a.shape
        nd = len(shape)
        jdims = (C.c_int * nd)(*shape[::-1])                     # Julia dimension order
        dev = _pick_device(x)
        ctx = _context(dev)
        ond, odims = C.c_int(), (C.c_int * 4)()
        out = _alloc((int(np.prod(shape)),), None, xb.dev, dev)
        ctx.check(N.lib().b200_dimmover(ctx._h, xb.ptr, nd, jdims, self.from_, self.to, _ptr(out), C.byref(ond), odims,
                                        N.PTR_DEVICE if xb.dev else N.PTR_HOST))
        oshape = tuple(odims[i] for i in range(ond.value))[::-1]
        return out.reshape(oshape), ({} if st is None else st)


# --------------------------------------------------------------------------------------- multiple shooting
def group_ranges(datasize: int, groupsize: int):
    """`group_ranges(datasize, groupsize)` (src/multiple_shooting.jl:231-239) as 0-based inclusive
    (first, last) pairs; raises DomainError-like ValueError on a bad groupsize."""
    if not (2 <= groupsize <= datasize):
        raise ValueError("DomainError: datasize must be positive and groupsize must to be within [2, datasize]")
    n = N.lib().b200_group_ranges(datasize, groupsize, None, None, 0)
    first = (C.c_int * n)()
    last = (C.c_int * n)()
    N.lib().b200_group_ranges(datasize, groupsize, first, last, n)
    return [(first[i], last[i]) for i in range(n)]


def _default_continuity_loss(u_end, u_0):
    """`sum(abs, u_end - u_0)` (src/multiple_shooting.jl:243)."""
    return (u_end - u_0).abs().sum() if _is_torch(u_end) else np.abs(u_end - u_0).sum()


@dataclass
class ODEProblem:
    """`ODEProblem(f, u0, tspan, p)` where f is a Chain (the NN right-hand side of a NeuralODE)."""
    f: Chain
    u0: object
    tspan: tuple
    p: object = None
    kwargs: dict = field(default_factory=dict)


if torch is not None:
    class _MshootFn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, p, data, run):
            preds, tapes, stats, rc = run.forward(p, data)
            ctx.run, ctx.tapes = run, tapes
            ctx.owner = [_TapeOwner(t) for t in tapes]
            run.retcode, run.stats = rc, stats
            return preds if _is_torch(preds) else torch.from_numpy(preds)

        @staticmethod
        def backward(ctx, g):
            g = g.contiguous()
            run = ctx.run
            if not run.on_dev:
                g = g.cpu().numpy()
            want_data = ctx.needs_input_grad[1]
            dp, dblocks = run.adjoint(ctx.tapes, g, want_data)
            if not _is_torch(dp):
                dp = torch.from_numpy(dp)
            ddata = None
            if want_data:
                # group g starts from ode_data[first[g]] (src/multiple_shooting.jl:52-63): its initial-condition
                # cotangent lands on that time index (neighbouring groups share a point, hence the accumulation)
                if not _is_torch(dblocks):
                    dblocks = torch.from_numpy(dblocks)
                ddata = torch.zeros((len(run.tsteps),) + tuple(dblocks.shape[1:]), dtype=dblocks.dtype, device=dblocks.device)
                for i, (a, _) in enumerate(run.ranges):
                    ddata[a] += dblocks[i]
            return dp, ddata, None


class _MshootRun:
    def __init__(self, chain, tsteps, ranges, solve: _Solve):
        self.chain, self.ranges, self.solve = chain, ranges, solve
        self.tsteps = np.ascontiguousarray(np.asarray(tsteps.detach().cpu().numpy() if _is_torch(tsteps) else tsteps, np.float32))
        self.retcode, self.stats = 0, None

    def forward(self, p, data):
        pb, db = _Buf(p), _Buf(data)
        shape = tuple(db.t.shape) if _is_torch(data) else db.a.shape
        T, B, D = shape
        self.on_dev, self.dev = db.dev, _pick_device(data, p)
        if pb.dev != db.dev:
            raise ValueError("p and ode_data must live on the same device")
        ctx = _context(self.dev)
        mh = _model_handle(ctx, self.chain)
        self.ctx, self.mh, self.B, self.D = ctx, mh, B, D
        G = len(self.ranges)
        o = self.solve.opts(self.on_dev, True)
        self.opts = o
        first = (C.c_int * G)(*[a for a, _ in self.ranges])
        last = (C.c_int * G)(*[b for _, b in self.ranges])
        rows = sum(b - a + 1 for a, b in self.ranges)
        preds = _alloc((rows, B, D), None, self.on_dev, self.dev)
        stats = (N.Stats * G)()
        tapes = (C.c_void_p * G)()
        if self.on_dev:
            ts_dev = torch.from_numpy(self.tsteps).to(f"cuda:{self.dev}")
            ts_ptr = ts_dev.data_ptr()
        else:
            ts_ptr = self.tsteps.ctypes.data
        rc = ctx.check(N.lib().b200_mshoot_forward(ctx._h, mh, pb.ptr, db.ptr, ts_ptr, T, B, G, first, last, C.byref(o),
                                                   _ptr(preds), stats, tapes))
        return preds, [C.c_void_p(t) for t in tapes], [s.as_dict() for s in stats], rc

    def adjoint(self, tapes, g, want_data=False):
        """dp (summed over the groups) and, on request, the per-group initial-condition cotangents [G][B][D]."""
        gb = _Buf(g)
        G = len(tapes)
        arr = (C.c_void_p * G)(*[t.value for t in tapes])
        dp = _alloc((self.chain.nparams,), None, self.on_dev, self.dev)
        dblocks = _alloc((G, self.B, self.D), None, self.on_dev, self.dev) if want_data else None
        stats = (N.Stats * G)()
        self.ctx.check(N.lib().b200_mshoot_adjoint(self.ctx._h, self.mh, arr, G, gb.ptr, C.byref(self.opts), _ptr(dp),
                                                   _ptr(dblocks), stats))
        self.adjoint_stats = [s.as_dict() for s in stats]
        return dp, dblocks


@dataclass
class EnsembleProblem:
    """`EnsembleProblem(prob)` (SciMLBase) as the ensemble `multiple_shoot` takes it (src/multiple_shooting.jl:131-195);
    `prob_func` / `output_func` / `reduction` other than the defaults are outside the B200 path."""
    prob: ODEProblem


class EnsembleThreads:
    pass


class EnsembleSerial:
    pass


def _multiple_shoot_ensemble(p, ode_data, tsteps, eprob, ensemblealg, loss_function, rest, continuity_term, kwargs):
    """Ensemble variant (src/multiple_shooting.jl:131-195): `ode_data` is Julia (D, T, trajectories) = C-order
    (trajectories, T, D); every (group, trajectory) pair is an INDEPENDENT solve with its own adaptive step sequence
    (that is what an EnsembleProblem means), so each trajectory runs as its own batch-1 grouped solve -- the groups of
    one trajectory in one C-ABI call, concurrently on the library's lanes.  group_predictions[i] is (trajectories, len, D)."""
    if len(rest) == 3:
        continuity_loss, solver, group_size = rest
    elif len(rest) == 2:
        continuity_loss = _default_continuity_loss
        solver, group_size = rest
    else:
        raise TypeError("multiple_shoot(p, ode_data, tsteps, ensembleprob, ensemblealg, loss_function, [continuity_loss], solver, group_size)")
    kwargs = dict(kwargs)
    nd = ode_data.dim() if _is_torch(ode_data) else np.asarray(ode_data).ndim
    if nd < 3:
        raise AssertionError("ode_data must have at least three dimension: `size(ode_data) = (problem_dimension,length(tsteps),trajectories)")
    ntraj, datasize = ode_data.shape[0], ode_data.shape[1]
    if group_size < 2 or group_size > datasize:
        raise ValueError("DomainError: group_size can't be < 2 or > number of data points")
    if datasize != len(tsteps):
        raise AssertionError("datasize == length(tsteps)")
    if kwargs.pop("trajectories", None) != ntraj:
        raise AssertionError("ntraj == kwargs[:trajectories]")
    ranges = group_ranges(datasize, group_size)
    per_traj, worst = [], 0
    for i in range(ntraj):
        data_i = ode_data[i][:, None, :]                           # (T, 1, D): one trajectory = one shared-dt batch of 1
        loss_i, preds_i = multiple_shoot(p, data_i, tsteps, eprob.prob, lambda u, uh: 0.0, lambda a, b: 0.0, solver, group_size,
                                         continuity_term=0, **kwargs)
        worst = max(worst, multiple_shoot.last.retcode)
        per_traj.append(preds_i)
    cat = (lambda xs: torch.cat(xs, dim=1)) if _is_torch(per_traj[0][0]) else (lambda xs: np.concatenate(xs, axis=1))
    group_predictions = []
    for g in range(len(ranges)):
        gp = cat([per_traj[i][g] for i in range(ntraj)])           # (len, trajectories, D)
        group_predictions.append(gp.permute(1, 0, 2) if _is_torch(gp) else np.transpose(gp, (1, 0, 2)))
    if not successful_retcode(worst):                               # `any(.!convergeds) && return Inf, group_predictions`
        return float("inf"), group_predictions
    loss = 0
    for g, (a, b) in enumerate(ranges):
        u = ode_data[:, a:b + 1]
        loss = loss + loss_function(u, group_predictions[g])
        if g > 0:
            loss = loss + continuity_term * continuity_loss(group_predictions[g - 1][:, -1], u[:, 0])
    return loss, group_predictions


def multiple_shoot(p, ode_data, tsteps, prob, loss_function, *rest, continuity_term=100, **kwargs):
    """`multiple_shoot(p, ode_data, tsteps, prob, loss_function, [continuity_loss], solver, group_size;
    continuity_term = 100, kwargs...)` (src/multiple_shooting.jl:36-94) -> (loss, group_predictions), and the ensemble
    method `multiple_shoot(p, ode_data, tsteps, ensembleprob, ensemblealg, loss_function, [continuity_loss], solver,
    group_size; trajectories, ...)` (src/multiple_shooting.jl:131-195).

    `ode_data` is (T, B, D) (Julia (D, B, T)) or (T, D).  All groups are solved by ONE C-ABI call;
    the loss callbacks run on the host framework (torch), whose autograd chains into the adjoint."""
    if isinstance(prob, EnsembleProblem):
        # positional layout of the ensemble method: (ensembleprob, ensemblealg, loss_function, [continuity_loss], solver, group_size)
        return _multiple_shoot_ensemble(p, ode_data, tsteps, prob, loss_function, rest[0], rest[1:], continuity_term, kwargs)
    if len(rest) == 3:
        continuity_loss, solver, group_size = rest
    elif len(rest) == 2:
        continuity_loss = _default_continuity_loss
        solver, group_size = rest
    else:
        raise TypeError("multiple_shoot(p, ode_data, tsteps, prob, loss_function, [continuity_loss], solver, group_size)")
    nd = ode_data.dim() if _is_torch(ode_data) else np.asarray(ode_data).ndim
    squeeze = nd == 2
    data3 = ode_data[:, None, :] if squeeze else ode_data
    datasize = data3.shape[0]
    if group_size < 2 or group_size > datasize:
        raise ValueError("DomainError: group_size can't be < 2 or > number of data points")
    ranges = group_ranges(datasize, group_size)
    kw = dict(prob.kwargs)
    kw.update(kwargs)
    solve = _Solve(prob.f, prob.tspan, (solver,), kw)
    run = _MshootRun(prob.f, tsteps, ranges, solve)
    use_torch = _is_torch(p) and _is_torch(data3)
    if use_torch:
        preds = _MshootFn.apply(p, data3, run)
    else:
        preds, tapes, stats, rc = run.forward(p, data3)
        run.retcode, run.stats = rc, stats
        for t in tapes:
            N.lib().b200_tape_free(t)
    group_predictions, row = [], 0
    for a, b in ranges:
        n = b - a + 1
        g = preds[row:row + n]
        group_predictions.append(g[:, 0, :] if squeeze else g)
        row += n
    multiple_shoot.last = run
    if not successful_retcode(run.retcode):          # src/multiple_shooting.jl:66-67
        return float("inf"), group_predictions
    loss = 0
    for i, (a, b) in enumerate(ranges):
        u = ode_data[a:b + 1]
        uh = group_predictions[i]
        loss = loss + loss_function(u, uh)
        if i > 0:
            loss = loss + continuity_term * continuity_loss(group_predictions[i - 1][-1], u[0])
    return loss, group_predictions


# --------------------------------------------------------------------------------------- multi-GPU plumbing
_dist_state = {"rank": 0, "world": 1}


def shard_range(n: int, rank: int, world: int):
    """Contiguous trajectory block of rank `rank`: equal shards are required
    because every rank must launch the same grid for the in-kernel slice exchange."""
    if n % world:
        raise ValueError(f"batch {n} is not divisible by the {world} ranks (equal shards are required)")
    per = n // world
    return rank * per, (rank + 1) * per


def gather_ipc_handles(handle: bytes, device=None) -> bytes:
    """all_gather of the 64-byte CUDA-IPC handles over the default torch.distributed group."""
    import torch.distributed as dist
    world = dist.get_world_size()
    t = torch.frombuffer(bytearray(handle), dtype=torch.uint8).clone()
    if dist.get_backend() == "nccl":
        t = t.cuda(device)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return b"".join(bytes(o.cpu().numpy().tobytes()) for o in out)


def attach_process_group(device: int = None):
    """Wire this process's b200 context to its peers: export the step mailbox (CUDA IPC), all_gather the
    handles, import.  Afterwards every solve on this device is one shard of a batch-sharded solve whose
    per-step error-norm terms are exchanged inside the persistent kernel over NVLink peer memory."""
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    if device is None:
        device = torch.cuda.current_device()
    ctx = _context(device)
    buf = C.create_string_buffer(N.IPC_HANDLE_BYTES)
    ctx.check(N.lib().b200_ctx_ipc_export(ctx._h, buf))
    handles = gather_ipc_handles(buf.raw, device)
    ctx.check(N.lib().b200_ctx_ipc_import(ctx._h, rank, world, handles))
    _dist_state.update(rank=rank, world=world)
    dist.barrier()
    return rank, world
