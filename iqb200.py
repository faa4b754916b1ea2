"""`import iqb200` -- import alias for the package directory `imagequilting.jl_b200/` (whose
name, fixed by the project layout, contains a dot and so cannot be imported directly)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "imagequilting.jl_b200")
_spec = importlib.util.spec_from_file_location(
    "iqb200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["iqb200"] = _mod
_spec.loader.exec_module(_mod)
