"""Import shim: the package directory is named `joint-multimodal-transformer-6th-abaw_b200` (not a valid
Python identifier), so it is loaded here under the importable name `jmt_b200`."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "joint-multimodal-transformer-6th-abaw_b200")
_NAME = "jmt_b200"

_spec = importlib.util.spec_from_file_location(_NAME, os.path.join(_PKG_DIR, "__init__.py"),
                                               submodule_search_locations=[_PKG_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[_NAME] = _mod
_spec.loader.exec_module(_mod)
