"""Import shim: `import cgoptim_b200` loads the package directory
`conjugategradientoptim.jl_b200/` (whose name, with a dot, is not importable directly)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "conjugategradientoptim.jl_b200")
_spec = importlib.util.spec_from_file_location(
    "cgoptim_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["cgoptim_b200"] = _mod
_spec.loader.exec_module(_mod)
