"""Import shim: the package lives in the directory ``cosmo.jl_b200/`` (the name the
project layout prescribes), which Python cannot import by name.  ``import
cosmo_b200`` loads that directory as the package ``cosmo_b200``."""
import importlib.util
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_pkg = os.path.join(_here, "cosmo.jl_b200")
_spec = importlib.util.spec_from_file_location("cosmo_b200", os.path.join(_pkg, "__init__.py"),
                                               submodule_search_locations=[_pkg])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["cosmo_b200"] = _mod
_spec.loader.exec_module(_mod)
