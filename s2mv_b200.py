"""Importable alias of the package directory `stereo-to-multiview-cuda_b200/`
(the hyphens in the mandated directory name are not valid in a Python
identifier): `import s2mv_b200` loads that package under this name."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stereo-to-multiview-cuda_b200")
_name = "s2mv_b200_pkg"
if _name not in sys.modules:
    _spec = importlib.util.spec_from_file_location(_name, os.path.join(_dir, "__init__.py"),
                                                   submodule_search_locations=[_dir])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_name] = _mod
    _spec.loader.exec_module(_mod)
_pkg = sys.modules[_name]
from s2mv_b200_pkg import *  # noqa: F401,F403,E402
from s2mv_b200_pkg import (COMPAT_SYMBOLS, EXPORTED_SYMBOLS, LIB_PATH, Params, Pipeline, S2mvError,  # noqa: E402,F401
                           build, default_params, lib)
