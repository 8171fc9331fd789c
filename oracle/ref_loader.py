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


def load_frame_transformer():
    """The UNMODIFIED src/models/frame_transformer.py as a module.  Its VidResNet asks torchvision for pretrained
    R(2+1)D-18 weights (a download; there is no network), so ``torchvision.models.video.r2plus1d_18`` and
    ``torchvision.models.resnet18`` are wrapped to construct the same architectures with ``weights=None`` — the only
    intervention, outside the reference file.  ``from models import custom_resnet`` resolves against the reference's own
    ``src/`` directory."""
    if not available():
        raise FileNotFoundError(f"reference not found under {REF_ROOT}")
    shims.install()
    import torchvision.models as tvm
    if not getattr(tvm.video.r2plus1d_18, "_tvt_offline", False):
        def offline(fn):
            def build(pretrained=False, **kw):
                kw.pop("weights", None)
                return fn(weights=None, **kw)
            build._tvt_offline = True
            return build
        tvm.video.r2plus1d_18 = offline(tvm.video.r2plus1d_18)
        tvm.resnet18 = offline(tvm.resnet18)
    src = os.path.join(REF_ROOT, "src")
    # the reference's src/models has no __init__.py (a namespace package), which loses to ANY regular package named
    # ``models`` anywhere on sys.path — the repo's own src/models shim in particular: take that directory off the path
    # while the reference file is executed
    here = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "src")
    saved_path = list(sys.path)
    sys.path[:] = [src] + [p for p in sys.path if os.path.abspath(p or ".") != here]
    saved_models = {k: sys.modules.pop(k) for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]}
    try:
        spec = importlib.util.spec_from_file_location("ref_frame_transformer", os.path.join(src, "models/frame_transformer.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
            del sys.modules[k]
        sys.modules.update(saved_models)
        sys.path[:] = saved_path
    return mod
