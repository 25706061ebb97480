"""Import shim: the package lives in ``options-model_b200/`` (not an importable name)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "options-model_b200")
_spec = importlib.util.spec_from_file_location(
    "options_model_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir]
)
_mod = importlib.util.module_from_spec(_spec)
sys.modules["options_model_b200"] = _mod
_spec.loader.exec_module(_mod)
