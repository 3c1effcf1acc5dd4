"""ctypes binding of libb200node.so -- the only way host code reaches CUDA in this package.

The declarations below mirror include/b200node.h one to one; tests/test_abi.py checks that every
symbol the header declares is exported.  There is no CPU fallback: if the shared library is missing
or no CUDA device is present the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB_PATH = os.path.join(HERE, "libb200node.so")
HEADER = os.path.join(ROOT, "include", "b200node.h")
SOURCES = [os.path.join(HERE, "csrc", f) for f in ("b200node.cu", "tc_kernels.cu", "device.cuh", "solver.cuh", "tc_common.cuh", "tc_engine.inc")]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC"]


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into libb200node.so (in-tree so it travels to the GPU box)."""
    newest = max(os.path.getmtime(p) for p in SOURCES + [HEADER])
    if not force and os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= newest:
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cwd = os.path.join(HERE, "csrc")
    objs, procs = [], []
    for src in SOURCES[:2]:                  # the two translation units compile in parallel
        obj = os.path.join(HERE, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        procs.append(subprocess.Popen(cmd, cwd=cwd))
    for pr in procs:
        if pr.wait() != 0:
            raise RuntimeError("nvcc failed")
    subprocess.run([nvcc, "-shared", "-o", LIB_PATH] + objs, check=True, cwd=cwd)
    return LIB_PATH


class Opts(C.Structure):
    _fields_ = [("abstol", C.c_float), ("reltol", C.c_float), ("dt0", C.c_float),
                ("maxiters", C.c_int), ("save_start", C.c_int), ("save_end", C.c_int),
                ("sensealg", C.c_int), ("ptr_kind", C.c_int), ("tile", C.c_int),
                ("max_dense_steps", C.c_int), ("seed", C.c_uint64), ("noise_per_call", C.c_int),
                ("b_offset", C.c_int), ("save_everystep", C.c_int), ("path", C.c_int), ("ffjord_exact_trace", C.c_int),
                ("ffjord_regularize", C.c_int)]


class Stats(C.Structure):
    _fields_ = [("retcode", C.c_int), ("naccept", C.c_int), ("nreject", C.c_int), ("nf", C.c_int),
                ("t_final", C.c_float), ("dt_last", C.c_float), ("tie_dist", C.c_float),
                ("kernel_ms", C.c_float), ("n_out", C.c_int), ("grid", C.c_int), ("tile", C.c_int),
                ("engine", C.c_int), ("reserved", C.c_int * 4)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


SUCCESS, MAXITERS, DT_LESS_THAN_MIN, UNSTABLE = 0, 1, 2, 3
RETCODE_NAMES = {0: "Success", 1: "MaxIters", 2: "DtLessThanMin", 3: "Unstable"}
ACT_IDENTITY, ACT_TANH, ACT_RELU, ACT_SIGMOID, ACT_SOFTPLUS = range(5)
PRE_NONE, PRE_CUBE = 0, 1
SENSE_NONE, SENSE_INTERPOLATING, SENSE_BACKSOLVE, SENSE_QUADRATURE, SENSE_GAUSS = -1, 0, 1, 2, 3
PTR_HOST, PTR_DEVICE = 0, 1
IPC_HANDLE_BYTES = 64

_vp, _fp, _ip = C.c_void_p, C.c_void_p, C.POINTER(C.c_int)   # data pointers travel as raw addresses
_SIGNATURES = {
    "b200_abi_version": (C.c_int, []),
    "b200_ctx_create": (C.c_int, [C.c_int, _vp, C.POINTER(_vp)]),
    "b200_ctx_destroy": (None, [_vp]),
    "b200_last_error": (C.c_char_p, [_vp]),
    "b200_ctx_sync": (C.c_int, [_vp]),
    "b200_opts_default": (None, [C.POINTER(Opts)]),
    "b200_ctx_trace": (C.c_int, [_vp, C.c_int, _fp, C.c_int]),
    "b200_model_create": (C.c_int, [_vp, C.c_int, _ip, _ip, C.c_int, C.POINTER(_vp)]),
    "b200_model_destroy": (None, [_vp]),
    "b200_model_nparams": (C.c_int, [_vp]),
    "b200_node_forward": (C.c_int, [_vp, _vp, _fp, _fp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float,
                                    _fp, C.c_int, C.POINTER(Opts), _fp, _fp, C.POINTER(Stats),
                                    C.POINTER(_vp)]),
    "b200_node_adjoint": (C.c_int, [_vp, _vp, _vp, _fp, C.POINTER(Opts), _fp, _fp, C.POINTER(Stats)]),
    "b200_tape_nsteps": (C.c_int, [_vp]),
    "b200_tape_steps": (C.c_int, [_vp, _vp, _fp, _fp, C.c_int]),
    "b200_tape_free": (None, [_vp]),
    "b200_ffjord_forward": (C.c_int, [_vp, _vp, _fp, _fp, C.c_int, C.c_float, C.c_float, C.POINTER(Opts),
                                      _fp, _fp, C.POINTER(Stats), C.POINTER(_vp)]),
    "b200_ffjord_adjoint": (C.c_int, [_vp, _vp, _vp, _fp, C.POINTER(Opts), _fp, _fp, C.POINTER(Stats)]),
    "b200_ffjord_adjoint_state": (C.c_int, [_vp, _vp, _vp, _fp, _fp, C.POINTER(Opts), _fp, _fp, C.POINTER(Stats)]),
    "b200_group_ranges": (C.c_int, [C.c_int, C.c_int, _ip, _ip, C.c_int]),
    "b200_mshoot_forward": (C.c_int, [_vp, _vp, _fp, _fp, _fp, C.c_int, C.c_int, C.c_int, _ip, _ip,
                                      C.POINTER(Opts), _fp, C.POINTER(Stats), C.POINTER(_vp)]),
    "b200_mshoot_adjoint": (C.c_int, [_vp, _vp, C.POINTER(_vp), C.c_int, _fp, C.POINTER(Opts), _fp, _fp,
                                      C.POINTER(Stats)]),
    "b200_node_train_step": (C.c_int, [_vp, _vp, _fp, _fp, _fp, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, _fp,
                                       C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, _fp, C.c_int, _fp, C.POINTER(Opts),
                                       _fp, C.POINTER(Stats), C.POINTER(Stats)]),
    "b200_dimmover": (C.c_int, [_vp, _fp, C.c_int, _ip, C.c_int, C.c_int, _fp, _ip, _ip, C.c_int]),
    "b200_ctx_ipc_export": (C.c_int, [_vp, C.c_char_p]),
    "b200_ctx_ipc_import": (C.c_int, [_vp, C.c_int, C.c_int, C.c_char_p]),
}

_lib = None


def lib():
    """The loaded shared library; raises if it has not been built (no fallback by design)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (nvcc, sm_100a). This package has no CPU or PyTorch fallback.")
        L = C.CDLL(os.environ.get("B200NODE_LIB") or LIB_PATH)   # the override is for kernel-variant experiments
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def header_symbols():
    """Names of the functions include/b200node.h declares (used by the ABI test)."""
    import re
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", text)))


class B200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libb200node error {code}: {msg}")
        self.code = code


class Context:
    """One CUDA device + stream (b200_ctx)."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self._h = _vp()
        rc = lib().b200_ctx_create(int(device), _vp(stream) if stream else None, C.byref(self._h))
        if rc != 0:
            raise B200Error(rc, "b200_ctx_create failed: no usable CUDA device (there is no CPU fallback)")
        self.device = device

    def check(self, rc):
        if rc < 0:
            raise B200Error(rc, lib().b200_last_error(self._h).decode())
        return rc

    def trace(self, capacity: int = -1, max_entries: int = 1 << 16):
        """Enable (capacity > 0) / read the (t, dt, EEst, accepted) trace of the last solve."""
        import numpy as np
        buf = np.zeros((max_entries, 4), np.float32)
        n = self.check(lib().b200_ctx_trace(self._h, capacity, buf.ctypes.data, max_entries))
        return buf[:n]

    def close(self):
        if self._h:
            lib().b200_ctx_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def default_opts() -> Opts:
    o = Opts()
    lib().b200_opts_default(C.byref(o))
    return o
