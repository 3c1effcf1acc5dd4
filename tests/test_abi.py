"""CPU-only checks of the drop-in boundary: the shared library loads and exports every symbol that
include/b200node.h declares, the host-side API mirrors the reference's argument handling, and the
product path refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os

import numpy as np
import pytest

import diffeqflux_b200 as dfx
from diffeqflux_b200 import _native as N


def test_library_exports_every_header_symbol():
    assert os.path.exists(N.LIB_PATH), "run `python -c 'import __graft_entry__ as g; g.build()'`"
    lib = N.lib()
    names = N.header_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
        assert n in N._SIGNATURES, f"{n} has no ctypes signature"
    assert lib.b200_abi_version() == 1


def test_struct_sizes_match_header():
    # b200_opts: 3 floats + 7 ints (40) -> u64 at 40, + 2 ints + 4 ints = 72; stats: 16 ints/floats = 64
    assert C.sizeof(N.Opts) == 72
    assert C.sizeof(N.Stats) == 64
    o = N.default_opts()
    assert abs(o.abstol - 1e-6) < 1e-12 and abs(o.reltol - 1e-3) < 1e-9
    assert (o.maxiters, o.save_start, o.save_end, o.sensealg, o.noise_per_call) == (100000, 1, 1, 0, 1)


def test_group_ranges_matches_reference_docstring():
    # src/multiple_shooting.jl:225-229: group_ranges(10, 5) == [1:5, 5:9, 9:10]
    assert dfx.group_ranges(10, 5) == [(0, 4), (4, 8), (8, 9)]
    assert dfx.group_ranges(65, 5)[-1] == (60, 64) and len(dfx.group_ranges(65, 5)) == 16
    for bad in ((10, 1), (4, 5)):
        with pytest.raises(ValueError):
            dfx.group_ranges(*bad)
    from oracle import mshoot
    for T, g in ((10, 5), (65, 5), (7, 2), (9, 9)):
        assert dfx.group_ranges(T, g) == [(a, b - 1) for a, b in mshoot.group_ranges(T, g)]


def test_unsupported_configurations_raise():
    chain = dfx.Chain(dfx.Dense(2, 8, dfx.tanh), dfx.Dense(8, 2))
    with pytest.raises(dfx.UnsupportedError):
        dfx.NeuralODE(chain, (0.0, 1.0))                       # alg = nothing
    with pytest.raises(dfx.UnsupportedError):
        dfx.NeuralODE(chain, (0.0, 1.0), object())             # not Tsit5
    with pytest.raises(dfx.UnsupportedError):
        dfx.NeuralODE(chain, (0.0, 1.0), dfx.Tsit5(), callback=None)
    with pytest.raises(dfx.UnsupportedError):
        dfx.NeuralODE(chain, (0.0, 1.0), dfx.Tsit5(), sensealg="TrackerAdjoint")
    with pytest.raises(dfx.UnsupportedError):
        dfx.Chain(dfx.Dense(2, 8, lambda x: x))
    with pytest.raises(ValueError):
        dfx.Chain(dfx.Dense(2, 8), dfx.Dense(9, 2))
    with pytest.raises(dfx.UnsupportedError):
        dfx.FFJORD(chain, (0.0, 1.0), (2,), dfx.Tsit5(), basedist="MvNormal")
    ff = dfx.FFJORD(chain, (0.0, 1.0), (2,), dfx.Tsit5())
    st = ff.initialstates(np.random.default_rng(0))
    assert st["regularize"] is False and st["monte_carlo"] is True     # src/ffjord.jl:61-66
    with pytest.raises(dfx.UnsupportedError):
        dfx.FFJORD(chain, (0.0, 1.0), (2,), dfx.Tsit5(), ad="AutoForwardDiff")
    # lambda_1 / lambda_2 of a regularize = true layer are taken out of the end state exactly as src/ffjord.jl:164-166 does
    z = np.arange(7 * 5, dtype=np.float32).reshape(7, 5)               # B = 7 trajectories, D = 2 -> L = 5 rows each
    dlp, l1, l2 = ff._lambdas(z, 2, True)
    assert np.array_equal(dlp, z[:, 2].reshape(1, -1)) and np.array_equal(l1, z[3:4]) and np.array_equal(l2, z[4:5])
    with pytest.raises(IndexError):
        ff._lambdas(z[:4], 2, True)


def test_setup_and_parameter_layout():
    chain = dfx.Chain(dfx.Dense(2, 50, dfx.tanh), dfx.Dense(50, 2))
    node = dfx.NeuralODE(chain, (0.0, 1.5), dfx.Tsit5(), saveat=0.1)
    ps, st = dfx.setup(np.random.default_rng(0), node)
    assert ps.shape == (252,) and ps.dtype == np.float32 and st == {}
    W1 = chain.view(ps, 1, "weight")
    assert W1.shape == (50, 2)
    # Julia (out,in) column-major: element (o, i) sits at o + out*i
    assert W1[3, 1] == ps[3 + 50 * 1]
    assert np.all(chain.view(ps, 2, "bias") == 0)


def test_saveat_splitting_follows_the_oracle():
    from diffeqflux_jl_b200.layers import _split_saveat
    from oracle.node import split_saveat
    cases = [dict(saveat=None), dict(saveat=0.1), dict(saveat=np.linspace(0, 1, 11)),
             dict(saveat=[0.25, 0.5]), dict(saveat=None, save_everystep=False, save_start=False)]
    for kw in cases:
        a = _split_saveat(kw.get("saveat"), 0.0, 1.0, kw.get("save_start"), kw.get("save_end"), kw.get("save_everystep"))
        b = split_saveat(kw.get("saveat"), 0.0, 1.0, kw.get("save_start"), kw.get("save_end"), kw.get("save_everystep"))
        np.testing.assert_array_equal(a[0], b[0])
        assert a[1:3] == b[1:3]


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    chain = dfx.Chain(dfx.Dense(2, 8, dfx.tanh), dfx.Dense(8, 2))
    node = dfx.NeuralODE(chain, (0.0, 1.0), dfx.Tsit5(), saveat=0.1)
    with pytest.raises(dfx.B200Error):
        node(np.zeros((3, 2), np.float32), np.zeros(chain.nparams, np.float32), {})


def test_product_path_does_not_import_the_oracle():
    import re
    root = os.path.dirname(N.HERE)
    for dirpath, _, files in os.walk(N.HERE):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".jl")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), f


def test_header_compiles_as_c_and_links(tmp_path):
    """The boundary is plain C: tests/c_abi_smoke.c (gcc -std=c99 -Wall -Werror) includes include/b200node.h, links
    against libb200node.so and calls the entry points that need no device.  Without a GPU b200_ctx_create must
    fail loudly (B200_ERR_CUDA) -- there is no CPU path."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.dirname(here)
    libdir = os.path.dirname(N.LIB_PATH)
    exe = str(tmp_path / "c_abi_smoke")
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-I" + os.path.join(root, "include"), os.path.join(here, "c_abi_smoke.c"),
                    "-L" + libdir, "-lb200node", "-Wl,-rpath," + libdir, "-o", exe], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    assert "c abi ok, sizeof(b200_opts)=72 sizeof(b200_stats)=64" in out
