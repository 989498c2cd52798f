"""Import shim: the package directory is `incrementalinference.jl_b200/` (a name Python cannot
import directly because of the dot), so `import iifb200` loads it under this alias."""
import importlib.util
import os
import sys

_d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "incrementalinference.jl_b200")
_spec = importlib.util.spec_from_file_location("iifb200", os.path.join(_d, "__init__.py"),
                                               submodule_search_locations=[_d])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["iifb200"] = _mod
_spec.loader.exec_module(_mod)
