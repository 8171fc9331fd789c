"""Import the UNMODIFIED reference model files from /root/reference under ``oracle.shims``.

Only usable where /root/reference exists (this container, not the GPU box).  Test infrastructure only.
Returns module objects whose classes are the reference's own code, executed as-is:
  ``transformer``   src/models/transformer.py   (imports cleanly once pytorch_lightning is shimmed)
  ``tpn``           src/models/TPN.py           (file has no import statements; exec'd with nn/torch/pl injected)
  ``vit``           src/models/vit.py           (imports as-is)
  ``collab``        src/models/collabgating.py  (no imports; exec'd with nn/torch/F/pl injected)
"""
import importlib.util
import os
import sys
import types

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import shims

REF_ROOT = os.environ.get("TVT_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "src/models/transformer.py"))


def _exec_file(relpath, name, inject):
    path = os.path.join(REF_ROOT, relpath)
    mod = types.ModuleType(name)
    mod.__dict__.update(inject)
    with open(path) as f:
        code = compile(f.read(), path, "exec")
    exec(code, mod.__dict__)
    return mod


def load():
    if not available():
        raise FileNotFoundError(f"reference not found under {REF_ROOT}")
    shims.install()
    import pytorch_lightning as pl

    out = types.SimpleNamespace()
    spec = importlib.util.spec_from_file_location("ref_transformer", os.path.join(REF_ROOT, "src/models/transformer.py"))
    out.transformer = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(out.transformer)
    spec = importlib.util.spec_from_file_location("ref_vit", os.path.join(REF_ROOT, "src/models/vit.py"))
    out.vit = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(out.vit)
    out.tpn = _exec_file("src/models/TPN.py", "ref_tpn", {"nn": nn, "torch": torch, "pl": pl, "custom_resnet": None})
    out.collab = _exec_file("src/models/collabgating.py", "ref_collab", {"nn": nn, "torch": torch, "F": F, "pl": pl})
    return out
