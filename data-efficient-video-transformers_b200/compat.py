"""Minimal stand-in for pytorch_lightning (absent from this image): the reference's model classes derive
from ``pl.LightningModule`` and use ``save_hyperparameters`` / ``hparams`` / ``log`` / ``device``
(src/models/transformer.py:28-34,143; src/models/frame_transformer.py:84-88,253-258).  When the real
package is importable it is used instead, so ``src/main.py``'s Trainer drives these modules unchanged."""
import torch
import torch.nn as nn

try:  # pragma: no cover - not installed in the build image
    import pytorch_lightning as pl
    LightningModule = pl.LightningModule
    HAVE_LIGHTNING = True
except Exception:
    HAVE_LIGHTNING = False

    class _HParams(dict):
        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError as e:
                raise AttributeError(k) from e

        def __setattr__(self, k, v):
            self[k] = v

    class LightningModule(nn.Module):
        def __init__(self):
            super().__init__()
            self._hparams = _HParams()
            self.logged = {}

        @property
        def hparams(self):
            return self._hparams

        def save_hyperparameters(self, **kwargs):
            self._hparams.update(kwargs)

        def log(self, name, value, *args, **kwargs):
            self.logged[name] = value.detach() if torch.is_tensor(value) else value

        @property
        def device(self):
            for p in self.parameters():
                return p.device
            return torch.device("cpu")
