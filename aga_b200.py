"""Import shim: ``import aga_b200`` loads the package stored in the hyphenated directory
``attention-guided-adaptation-for-code-switching-speech-recognition_b200/`` (not a valid identifier)."""
import importlib.util
import os
import sys

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                    "attention-guided-adaptation-for-code-switching-speech-recognition_b200")
_spec = importlib.util.spec_from_file_location("aga_b200", os.path.join(_DIR, "__init__.py"),
                                               submodule_search_locations=[_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["aga_b200"] = _mod
_spec.loader.exec_module(_mod)
