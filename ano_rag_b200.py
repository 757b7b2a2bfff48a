"""Import shim: the package lives in the directory ``ano-rag_b200/`` (the layout the
build contract names); a hyphen is not importable, so this module loads that
directory under the name ``ano_rag_b200``."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ano-rag_b200")
_spec = importlib.util.spec_from_file_location(
    "ano_rag_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["ano_rag_b200"] = _mod
_spec.loader.exec_module(_mod)
