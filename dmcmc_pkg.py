"""Import helper: the package directory is named `diffusionmcmc.jl_b200` (a dot is not a valid
Python identifier), so it is loaded under the module name `diffusionmcmc_jl_b200`."""
import importlib.util
import os
import sys

_NAME = "diffusionmcmc_jl_b200"
_ROOT = os.path.dirname(os.path.abspath(__file__))


def load():
    if _NAME in sys.modules:
        return sys.modules[_NAME]
    path = os.path.join(_ROOT, "diffusionmcmc.jl_b200")
    spec = importlib.util.spec_from_file_location(
        _NAME, os.path.join(path, "__init__.py"), submodule_search_locations=[path])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[_NAME] = mod
    spec.loader.exec_module(mod)
    return mod
