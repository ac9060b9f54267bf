"""Import shim: ``import lsnf_b200`` loads the package that lives in the directory
``latent-space-normalizing-flow_b200/`` (the repository layout the project brief names; a
hyphenated directory cannot be imported by name, so this file registers it under a valid one).
"""
import importlib.util as _ilu
import os as _os
import sys as _sys

_pkg_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "latent-space-normalizing-flow_b200")
_spec = _ilu.spec_from_file_location(
    "lsnf_b200", _os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir]
)
_mod = _ilu.module_from_spec(_spec)
_sys.modules["lsnf_b200"] = _mod
_spec.loader.exec_module(_mod)
