"""Import the UNMODIFIED reference package from /root/reference (build container only).

pytorch_lightning / pathos are not installed here, so two tiny stand-in modules are registered first:
`pytorch_lightning.LightningModule` (an nn.Module with no-op save_hyperparameters/log and device/dtype
properties) and `pathos.multiprocessing.Pool`.  Used by oracle/make_golden.py (fixture generation) and by
`bench.py --impl reference` when the path exists; never by the product, never on the GPU box.
"""

from __future__ import annotations

import sys
import types
from pathlib import Path

REFERENCE_ROOT = Path("/root/reference")


def available() -> bool:
    return (REFERENCE_ROOT / "tqdne" / "edm.py").exists()


def _install_stubs() -> None:
    import torch
    from torch import nn

    if "pytorch_lightning" not in sys.modules:
        pl = types.ModuleType("pytorch_lightning")

        class LightningModule(nn.Module):
            def save_hyperparameters(self, *a, **k):
                pass

            def log(self, *a, **k):
                pass

            @property
            def device(self):
                return next(self.parameters()).device

            @property
            def dtype(self):
                return next(self.parameters()).dtype

        class Callback:
            pass

        pl.LightningModule = LightningModule
        pl.Callback = Callback
        sys.modules["pytorch_lightning"] = pl
    if "pathos" not in sys.modules:
        pathos = types.ModuleType("pathos")
        mp = types.ModuleType("pathos.multiprocessing")

        class Pool:
            def map(self, fn, it):
                return [fn(x) for x in it]

            def close(self):
                pass

        mp.Pool = Pool
        pathos.multiprocessing = mp
        sys.modules["pathos"] = pathos
        sys.modules["pathos.multiprocessing"] = mp


def load():
    """Returns the imported reference `tqdne` package (edm, unet, blocks, autoencoder, architectures)."""
    if not available():
        raise FileNotFoundError(f"reference not found under {REFERENCE_ROOT}")
    _install_stubs()
    if str(REFERENCE_ROOT) not in sys.path:
        sys.path.insert(0, str(REFERENCE_ROOT))
    import tqdne  # noqa: F401
    import tqdne.architectures
    import tqdne.autoencoder
    import tqdne.blocks
    import tqdne.edm
    import tqdne.unet

    return tqdne
