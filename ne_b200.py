"""Import shim: the package directory is `numericalearth.jl_b200/` (a dot is not importable), so
load it under the module name `numericalearth_jl_b200` and re-export it as `ne_b200`."""
import importlib.util
import os
import sys

_NAME = "numericalearth_jl_b200"
if _NAME not in sys.modules:
    _dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "numericalearth.jl_b200")
    _spec = importlib.util.spec_from_file_location(_NAME, os.path.join(_dir, "__init__.py"),
                                                   submodule_search_locations=[_dir])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_NAME] = _mod
    _spec.loader.exec_module(_mod)
_pkg = sys.modules[_NAME]
globals().update({k: v for k, v in vars(_pkg).items() if not k.startswith("__")})
package = _pkg
