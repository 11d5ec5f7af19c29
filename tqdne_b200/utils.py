"""Boundary glue (reference: tqdne/utils.py:11-43,93-101)."""

from __future__ import annotations

import logging
from collections.abc import Mapping, Sequence
from pathlib import Path

import torch


def get_device() -> str:
    """This engine is CUDA-only; anything else is an error rather than a silent CPU run."""
    if not torch.cuda.is_available():
        raise RuntimeError("tqdne_b200 needs a CUDA device (sm_100a); there is no CPU / MPS path")
    return "cuda"


def to_numpy(x):
    if isinstance(x, (str, bytes)):
        return x
    if isinstance(x, Sequence):
        return x.__class__(to_numpy(v) for v in x)
    if isinstance(x, Mapping):
        return x.__class__((k, to_numpy(v)) for k, v in x.items())
    return x.numpy(force=True) if isinstance(x, torch.Tensor) else x


def get_last_checkpoint(dirpath):
    ckpts = sorted(Path(dirpath).glob("*.ckpt"))
    if not ckpts:
        logging.info("No checkpoint found. Returning None.")
        return None
    logging.info(f"Last checkpoint is : {ckpts[-1]}")
    return ckpts[-1]


def load_model(type, path: Path, **kwargs):
    if not Path(path).exists():
        logging.info("Model not found. Returning None.")
        return None
    return type.load_from_checkpoint(path, **kwargs)
