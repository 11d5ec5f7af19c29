"""`LightningModule` base for the drop-in classes.

The reference subclasses `pytorch_lightning.LightningModule` (tqdne/edm.py:55, tqdne/autoencoder.py:7).  When
pytorch_lightning is importable we subclass it too, so `isinstance` checks, `Trainer` plumbing and
`load_from_checkpoint` behave exactly as upstream.  When it is not (this image), the small class below supplies
the parts of that interface the sampling path touches: `save_hyperparameters`, `log`, `device`, `dtype`,
`hparams` and a `load_from_checkpoint` that reads Lightning's on-disk `.ckpt` layout
(`state_dict`, `hyper_parameters`; reference call sites generate_waveforms.py:114-124,161-174).
"""

from __future__ import annotations

import inspect

import torch
from torch import nn

try:  # pragma: no cover - not installed in the build image
    import pytorch_lightning as _pl

    LightningModule = _pl.LightningModule
    HAVE_LIGHTNING = True
except Exception:  # noqa: BLE001
    HAVE_LIGHTNING = False

    class LightningModule(nn.Module):
        def __init__(self):
            super().__init__()
            self._hparams = {}

        # -- hyper-parameters ----------------------------------------------------------------------
        def save_hyperparameters(self, *args, ignore=None, **kwargs):
            ignore = {ignore} if isinstance(ignore, str) else set(ignore or ())
            frame = inspect.currentframe().f_back
            init = getattr(type(self), "__init__")
            names = [p for p in inspect.signature(init).parameters if p != "self"]
            local = frame.f_locals
            self._hparams = {k: local[k] for k in names if k in local and k not in ignore}

        @property
        def hparams(self):
            return self._hparams

        def log(self, *args, **kwargs):
            return None

        @property
        def device(self) -> torch.device:
            try:
                return next(self.parameters()).device
            except StopIteration:
                return torch.device("cpu")

        @property
        def dtype(self) -> torch.dtype:
            try:
                return next(self.parameters()).dtype
            except StopIteration:
                return torch.float32

        # -- checkpoints ---------------------------------------------------------------------------
        @classmethod
        def load_from_checkpoint(cls, checkpoint_path, map_location=None, strict=True, **kwargs):
            """Read a Lightning `.ckpt` (torch.save dict with `state_dict` and `hyper_parameters`)."""
            ckpt = torch.load(str(checkpoint_path), map_location=map_location or "cpu", weights_only=False)
            hp = dict(ckpt.get("hyper_parameters", {}))
            hp.update(kwargs)
            accepted = inspect.signature(cls.__init__).parameters
            model = cls(**{k: v for k, v in hp.items() if k in accepted})
            model.load_state_dict(ckpt["state_dict"], strict=strict)
            return model

        def on_save_checkpoint_dict(self) -> dict:
            """The dict Lightning would write for this module (used by tests and the CLI's --save)."""
            return {"state_dict": self.state_dict(), "hyper_parameters": dict(self._hparams),
                    "pytorch-lightning_version": "2.5.1"}
