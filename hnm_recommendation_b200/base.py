"""Base class for the model mirrors: pytorch_lightning.LightningModule when it is
installed (as in the reference, src/models/lightgcn.py:13), otherwise torch.nn.Module
with the two Lightning conveniences the reference's classes rely on."""
from __future__ import annotations

import inspect
from types import SimpleNamespace

import torch

try:  # pragma: no cover - not installed in the build image
    import pytorch_lightning as pl

    ModelBase = pl.LightningModule
    HAVE_LIGHTNING = True
except Exception:  # noqa: BLE001
    HAVE_LIGHTNING = False

    class ModelBase(torch.nn.Module):
        """nn.Module + save_hyperparameters()/hparams/log()/device, as Lightning provides them."""

        def save_hyperparameters(self, *args, **kwargs) -> None:
            frame = inspect.currentframe().f_back
            init = type(self).__init__
            names = [p for p in inspect.signature(init).parameters if p != "self"]
            self.hparams = SimpleNamespace(**{n: frame.f_locals[n] for n in names if n in frame.f_locals})

        @property
        def hyper_parameters(self) -> dict:
            return dict(vars(self.hparams))

        def log(self, *args, **kwargs) -> None:
            pass

        @property
        def device(self) -> torch.device:
            try:
                return next(self.parameters()).device
            except StopIteration:
                return torch.device("cpu")
