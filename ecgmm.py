"""Import shim: exposes the package in ``ecg-multimodal-model_b200/`` (not a valid Python
identifier) under the importable name ``ecgmm``."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ecg-multimodal-model_b200")
_spec = importlib.util.spec_from_file_location(
    "ecgmm", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir]
)
_mod = importlib.util.module_from_spec(_spec)
sys.modules["ecgmm"] = _mod
_spec.loader.exec_module(_mod)
