"""Import shim: loads the package directory `diffeqflux.jl_b200/` (not a legal identifier) under the
module name `diffeqflux_jl_b200` and re-exports its public API."""
import importlib.util
import os
import sys

_NAME = "diffeqflux_jl_b200"
_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "diffeqflux.jl_b200")

if _NAME not in sys.modules:
    _spec = importlib.util.spec_from_file_location(_NAME, os.path.join(_DIR, "__init__.py"),
                                                   submodule_search_locations=[_DIR])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_NAME] = _mod
    _spec.loader.exec_module(_mod)
_pkg = sys.modules[_NAME]
from diffeqflux_jl_b200 import *  # noqa: E402,F401,F403
from diffeqflux_jl_b200 import _native, layers  # noqa: E402,F401

__all__ = _pkg.__all__
