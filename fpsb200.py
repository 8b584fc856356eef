"""Import alias: `import fpsb200` loads the package that lives in the directory
`fletcherpenaltysolver.jl_b200/` (not a valid Python identifier because of the dot)."""
import importlib.util
import os
import sys

_d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fletcherpenaltysolver.jl_b200")
_spec = importlib.util.spec_from_file_location(
    __name__, os.path.join(_d, "__init__.py"), submodule_search_locations=[_d])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
