"""Import shim: the product package lives in ``practical-multi-view_b200/`` (the name the
build contract fixes; the hyphen makes it un-importable by name), so this module loads it
under the importable alias ``pmv_b200``.  ``import pmv_b200`` == the package."""
import importlib.util as _u
import sys as _s
from pathlib import Path as _P

_dir = _P(__file__).resolve().parent / "practical-multi-view_b200"
_spec = _u.spec_from_file_location("pmv_b200", _dir / "__init__.py",
                                   submodule_search_locations=[str(_dir)])
_mod = _u.module_from_spec(_spec)
_s.modules["pmv_b200"] = _mod
_spec.loader.exec_module(_mod)
