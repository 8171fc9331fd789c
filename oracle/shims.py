"""Stand-ins for third-party packages the reference imports but this image lacks.

``install()`` registers minimal ``pytorch_lightning`` / ``torchmetrics`` / ``pytorch_grad_cam`` /
``wandb`` modules in ``sys.modules`` when the real ones are not importable.  ``LightningModule`` is
``nn.Module`` plus the handful of attributes the reference touches (``save_hyperparameters``,
``hparams``, ``log``, ``device``).  Test infrastructure only.
"""
import importlib
import inspect
import sys
import types

import torch
import torch.nn as nn


class _HParams(dict):
    """Attribute-style dict, like Lightning's AttributeDict."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


class LightningModule(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        self._hparams = _HParams()
        self.logged = {}

    @property
    def hparams(self):
        return self._hparams

    def save_hyperparameters(self, *args, **kwargs):
        # Collect the caller's **kwargs the way Lightning does (reference passes everything as kwargs).
        frame = inspect.currentframe().f_back
        local = frame.f_locals
        init_kwargs = local.get("kwargs", {})
        for k, v in init_kwargs.items():
            self._hparams[k] = v
        for k, v in local.items():
            if k in ("self", "kwargs", "__class__") or k.startswith("_"):
                continue
            if isinstance(v, (int, float, str, bool, list, tuple, type(None))):
                self._hparams.setdefault(k, v)

    def log(self, name, value, *a, **k):
        self.logged[name] = value

    @property
    def device(self):
        try:
            return next(self.parameters()).device
        except StopIteration:  # pragma: no cover
            return torch.device("cpu")


class LightningDataModule:
    def __init__(self, *a, **k):
        pass


class Callback:
    pass


class _Metric(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()

    def forward(self, *a, **k):
        return torch.tensor(0.0)


def _missing(name):
    try:
        importlib.import_module(name)
        return False
    except Exception:
        return True


def install():
    if _missing("pytorch_lightning"):
        pl = types.ModuleType("pytorch_lightning")
        pl.LightningModule = LightningModule
        pl.LightningDataModule = LightningDataModule
        pl.Callback = Callback
        cb = types.ModuleType("pytorch_lightning.callbacks")
        cb.Callback = Callback
        pl.callbacks = cb
        sys.modules["pytorch_lightning"] = pl
        sys.modules["pytorch_lightning.callbacks"] = cb
    if _missing("torchmetrics"):
        tm = types.ModuleType("torchmetrics")
        tm.AUROC = tm.F1 = tm.AveragePrecision = _Metric
        sys.modules["torchmetrics"] = tm
    if _missing("pytorch_grad_cam"):
        gc = types.ModuleType("pytorch_grad_cam")
        for n in ("GradCAM", "ScoreCAM", "GradCAMPlusPlus", "AblationCAM", "XGradCAM", "EigenCAM", "FullGrad"):
            setattr(gc, n, object)
        u = types.ModuleType("pytorch_grad_cam.utils")
        mt = types.ModuleType("pytorch_grad_cam.utils.model_targets")
        mt.ClassifierOutputTarget = object
        im = types.ModuleType("pytorch_grad_cam.utils.image")
        im.show_cam_on_image = lambda *a, **k: None
        sys.modules.update({"pytorch_grad_cam": gc, "pytorch_grad_cam.utils": u,
                            "pytorch_grad_cam.utils.model_targets": mt, "pytorch_grad_cam.utils.image": im})
    if _missing("wandb"):  # present in this image, stub kept for completeness
        sys.modules["wandb"] = types.ModuleType("wandb")
