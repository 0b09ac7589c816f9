"""Import shim: the package directory is named ``cu-sdr-collection_b200`` (not a valid Python
identifier), so ``import cu_sdr_collection_b200`` lands here and this module replaces itself
with the real package loaded from that directory."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cu-sdr-collection_b200")
_spec = importlib.util.spec_from_file_location(
    "cu_sdr_collection_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["cu_sdr_collection_b200"] = _mod
_spec.loader.exec_module(_mod)
