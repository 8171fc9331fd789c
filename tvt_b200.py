"""Import alias: ``import tvt_b200`` loads the package directory ``data-efficient-video-transformers_b200``
(whose name, mirroring the reference repository, is not a valid Python identifier)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("data-efficient-video-transformers_b200")
sys.modules[__name__] = _pkg
