"""Import shim: the product package lives in `dsp.jl_b200/` (a directory name Python cannot import
directly); this module loads it under the name `dspb200`."""
import importlib.util
import os
import sys

_pkg_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dsp.jl_b200")
_spec = importlib.util.spec_from_file_location("dspb200", os.path.join(_pkg_dir, "__init__.py"),
                                               submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["dspb200"] = _mod
_spec.loader.exec_module(_mod)
