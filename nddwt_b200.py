"""Importable alias of the package directory `non-decimated_wavelets_b200/` (whose name, taken
from the reference repo, is not a Python identifier)."""
import importlib as _importlib
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.abspath(__file__))
if _root not in _sys.path:
    _sys.path.insert(0, _root)
_pkg = _importlib.import_module("non-decimated_wavelets_b200")
globals().update({k: getattr(_pkg, k) for k in _pkg.__all__})
api = _importlib.import_module("non-decimated_wavelets_b200.api")
_lib = _importlib.import_module("non-decimated_wavelets_b200._lib")
__all__ = list(_pkg.__all__) + ["api"]
