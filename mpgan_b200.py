"""Import shim: exposes the package directory ``multi-pass-gan_b200/`` as module ``mpgan_b200``."""
import importlib.util
import os
import sys

_root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "multi-pass-gan_b200")
_spec = importlib.util.spec_from_file_location(
    "mpgan_b200", os.path.join(_root, "__init__.py"), submodule_search_locations=[_root])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["mpgan_b200"] = _mod
_spec.loader.exec_module(_mod)
