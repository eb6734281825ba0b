"""Import shim: the package directory is named ``covariancefunctions.jl_b200`` (not an identifier), so this module
loads it and registers it as ``covfn_b200``.  ``import covfn_b200`` therefore yields the package itself."""
import importlib.util as _u
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "covariancefunctions.jl_b200")
_spec = _u.spec_from_file_location("covfn_b200", _os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = _u.module_from_spec(_spec)
_sys.modules["covfn_b200"] = _mod
_spec.loader.exec_module(_mod)
