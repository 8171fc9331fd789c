"""Import alias: ``import tvt_b200`` (and ``tvt_b200.<submodule>``) resolves to the package directory
``data-efficient-video-transformers_b200`` — whose name, mirroring the reference repository, is not a valid
Python identifier — without ever creating a second copy of a module."""
import importlib
import importlib.abc
import importlib.util
import os
import sys

_REAL = "data-efficient-video-transformers_b200"
_ALIAS = __name__
_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)


class _AliasLoader(importlib.abc.Loader):
    def __init__(self, mod):
        self.mod = mod

    def create_module(self, spec):
        return self.mod            # hand back the already-imported real module

    def exec_module(self, module):
        pass


class _AliasFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.startswith(_ALIAS + "."):
            real = importlib.import_module(_REAL + fullname[len(_ALIAS):])
            return importlib.util.spec_from_loader(fullname, _AliasLoader(real))
        return None


if not any(isinstance(f, _AliasFinder) for f in sys.meta_path):
    sys.meta_path.insert(0, _AliasFinder())
_pkg = importlib.import_module(_REAL)
sys.modules[_ALIAS] = _pkg
