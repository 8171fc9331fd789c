"""Minimal stand-ins for pytorch_lightning / torchmetrics (absent from this image): the reference's model classes derive
from ``pl.LightningModule`` and use ``save_hyperparameters`` / ``hparams`` / ``log`` / ``device``
(src/models/transformer.py:28-34,143; src/models/frame_transformer.py:84-88,253-258), its FrameTransformer holds two
``torchmetrics.AveragePrecision`` objects (:114,117), and src/main.py:87-111 drives everything through ``pl.Trainer``.
When the real packages are importable they are used instead, so ``src/main.py``'s Trainer drives these modules unchanged.
``Trainer`` here is the smallest loop with Lightning's hook order (configure_optimizers -> training_step -> backward ->
optimizer step; validation_step + ``on_validation_epoch_end`` callbacks; test_step + ``on_test_epoch_end``) — enough to
exercise the drop-in classes the way main.py does, not a re-implementation of Lightning."""
import inspect

import torch
import torch.nn as nn

try:  # pragma: no cover - not installed in the build image
    import pytorch_lightning as pl
    LightningModule = pl.LightningModule
    Callback = pl.Callback
    Trainer = pl.Trainer
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

        def save_hyperparameters(self, *names, **kwargs):
            """Lightning semantics: with no arguments, every argument of the calling ``__init__`` (its ``**kwargs``
            included) becomes a hyper-parameter.  Explicit keyword arguments are stored as given."""
            if not names and not kwargs:
                frame = inspect.currentframe().f_back
                code = frame.f_code
                nargs = code.co_argcount + code.co_kwonlyargcount
                names_ = list(code.co_varnames[:nargs])
                if code.co_flags & inspect.CO_VARARGS:
                    nargs += 1
                if code.co_flags & inspect.CO_VARKEYWORDS:
                    v = frame.f_locals.get(code.co_varnames[nargs])
                    if isinstance(v, dict):
                        self._hparams.update(v)
                for k in names_:
                    if k != "self":
                        self._hparams.setdefault(k, frame.f_locals.get(k))
            self._hparams.update(kwargs)

        def log(self, name, value, *args, **kwargs):
            self.logged[name] = value.detach() if torch.is_tensor(value) else value

        @property
        def device(self):
            for p in self.parameters():
                return p.device
            return torch.device("cpu")

    class Callback:
        pass

    class Trainer:
        """fit / validate / test with Lightning's hook names and order, single device, no logging backends."""

        def __init__(self, max_epochs=1, max_steps=-1, callbacks=None, limit_val_batches=None, **_ignored):
            self.max_epochs, self.max_steps = max_epochs, max_steps
            self.callbacks = list(callbacks or [])
            self.limit_val_batches = limit_val_batches
            self.global_step = 0
            self.current_epoch = 0

        @staticmethod
        def _loader(dm, kind):
            if dm is None:
                return None
            fn = getattr(dm, f"{kind}_dataloader", None)
            return fn() if fn is not None else (dm if kind == "train" else None)

        def _hook(self, name, module):
            for cb in self.callbacks:
                fn = getattr(cb, name, None)
                if fn is not None:
                    fn(self, module)

        def _eval_loop(self, module, loader, step_name, end_hook):
            if loader is None:
                return
            module.eval()
            with torch.no_grad():
                for i, batch in enumerate(loader):
                    if self.limit_val_batches is not None and i >= self.limit_val_batches:
                        break
                    getattr(module, step_name)(batch, i)
            self._hook(end_hook, module)

        def fit(self, module, datamodule=None, train_dataloaders=None, val_dataloaders=None):
            train = train_dataloaders if train_dataloaders is not None else self._loader(datamodule, "train")
            val = val_dataloaders if val_dataloaders is not None else self._loader(datamodule, "val")
            opt = module.configure_optimizers()
            opt = opt[0] if isinstance(opt, (list, tuple)) else opt
            for epoch in range(self.max_epochs):
                self.current_epoch = epoch
                module.train()
                for i, batch in enumerate(train):
                    opt.zero_grad(set_to_none=True)
                    loss = module.training_step(batch, i)
                    loss.backward()
                    opt.step()
                    self.global_step += 1
                    if 0 < self.max_steps <= self.global_step:
                        break
                self._eval_loop(module, val, "validation_step", "on_validation_epoch_end")
                if 0 < self.max_steps <= self.global_step:
                    break

        def validate(self, module, datamodule=None, dataloaders=None):
            self._eval_loop(module, dataloaders if dataloaders is not None else self._loader(datamodule, "val"),
                            "validation_step", "on_validation_epoch_end")

        def test(self, module, datamodule=None, dataloaders=None, ckpt_path=None):
            self._eval_loop(module, dataloaders if dataloaders is not None else self._loader(datamodule, "test"),
                            "test_step", "on_test_epoch_end")


try:  # pragma: no cover - not installed in the build image
    from torchmetrics import AveragePrecision
except Exception:
    class AveragePrecision(nn.Module):
        """Stateless stand-in for ``torchmetrics.AveragePrecision(num_classes=...)``: the reference only calls it on
        (logits, int targets) and logs the object (frame_transformer.py:277-281); sklearn computes the reported metrics
        in the callback.  Keeps no state, so it adds no ``state_dict`` keys (neither does torchmetrics')."""

        def __init__(self, num_classes=None, **_ignored):
            super().__init__()
            self.num_classes = num_classes
            self.updates = 0

        def forward(self, preds, target):
            self.updates += 1
            return torch.zeros((), device=preds.device)
