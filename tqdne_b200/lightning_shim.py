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
import pickle

import torch
from torch import nn


class _RemapUnpickler(pickle.Unpickler):
    """`tqdne.<module>.<name>` -> `tqdne_b200.<module>.<name>` when the reference package is absent."""

    def find_class(self, module, name):
        if module == "tqdne" or module.startswith("tqdne."):
            try:
                return super().find_class(module, name)
            except (ImportError, AttributeError):
                module = "tqdne_b200" + module[len("tqdne"):]
        return super().find_class(module, name)


class _RemapPickle:
    """The slice of the `pickle` module interface torch.load(pickle_module=...) uses."""

    __name__ = "pickle"
    Unpickler = _RemapUnpickler
    Pickler = pickle.Pickler
    HIGHEST_PROTOCOL = pickle.HIGHEST_PROTOCOL

    @staticmethod
    def load(f, **kw):
        return _RemapUnpickler(f, **kw).load()

    @staticmethod
    def loads(b, **kw):
        import io

        return _RemapUnpickler(io.BytesIO(b), **kw).load()


def read_checkpoint(checkpoint_path, map_location=None) -> dict:
    """torch.load of a Lightning checkpoint with the `tqdne.*` -> `tqdne_b200.*` class remap."""
    return torch.load(str(checkpoint_path), map_location=map_location or "cpu", weights_only=False,
                      pickle_module=_RemapPickle)

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
        def load_from_checkpoint(cls, checkpoint_path, map_location=None, strict=True, use_ema=False, **kwargs):
            """Read a Lightning `.ckpt` written by the reference (torch.save dict with `state_dict`,
            `hyper_parameters` -- which pickle `tqdne.edm.EDM`, reference generate_waveforms.py:170 -- and, when the
            EMA callback was active, `ema_state`, reference ema.py:50-54).  Objects pickled under `tqdne.*` resolve to
            their `tqdne_b200.*` counterparts when the reference package is not importable.  `use_ema=True` (engine
            extension) overlays the EMA weights like the reference does for validation (ema.py:30-32)."""
            ckpt = read_checkpoint(checkpoint_path, map_location)
            hp = dict(ckpt.get("hyper_parameters", {}))
            hp.update(kwargs)
            accepted = inspect.signature(cls.__init__).parameters
            model = cls(**{k: v for k, v in hp.items() if k in accepted})
            model.load_state_dict(ckpt["state_dict"], strict=strict)
            if use_ema:
                if "ema_state" not in ckpt:
                    raise KeyError(f"{checkpoint_path}: no ema_state in this checkpoint")
                model.load_state_dict(ckpt["ema_state"], strict=False)
            return model

        def on_save_checkpoint_dict(self) -> dict:
            """The dict Lightning would write for this module (used by tests and the CLI's --save)."""
            return {"state_dict": self.state_dict(), "hyper_parameters": dict(self._hparams),
                    "pytorch-lightning_version": "2.5.1"}
