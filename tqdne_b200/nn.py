"""Parameter holders shared by the network definitions.

The classes here only *own* parameters under the same names, shapes and default initialisation as the
reference primitives (tqdne/nn.py) so that reference checkpoints load unchanged; the arithmetic is done by
the CUDA kernels the lowering in `tqdne_b200.lowering` emits, never by these modules' own forward.
"""

from __future__ import annotations

import torch
from torch import nn


class EngineOnly(nn.Module):
    """A module whose maths lives in the engine: calling it directly is an error, not a fallback."""

    def forward(self, *args, **kwargs):  # pragma: no cover - guard
        raise RuntimeError(
            f"{type(self).__name__} is lowered into a tqdne_b200 kernel plan by its parent network "
            "(UNetModel / Encoder / Decoder); it has no stand-alone PyTorch forward."
        )


class GroupNorm32(nn.GroupNorm):
    """32-group normalisation, statistics in fp32 (reference: tqdne/nn.py:11-13)."""

    def forward(self, x):  # pragma: no cover - guard
        raise RuntimeError("GroupNorm32 runs inside the tqdne_b200 GroupNorm+SiLU kernel")


def normalization(channels: int) -> GroupNorm32:
    """reference: tqdne/nn.py:90-105 -> GroupNorm32(32, channels), eps 1e-5, affine."""
    return GroupNorm32(32, channels)


_CONV = {1: nn.Conv1d, 2: nn.Conv2d, 3: nn.Conv3d}


def conv_nd(dims: int, *args, **kwargs) -> nn.Module:
    """reference: tqdne/nn.py:16-24 (same ValueError for unsupported dims)."""
    if dims not in _CONV:
        raise ValueError(f"unsupported dimensions: {dims}")
    if dims == 3:
        raise ValueError("tqdne_b200 lowers 1D and 2D convolutions only (no shipped config uses dims=3)")
    return _CONV[dims](*args, **kwargs)


def zero_module(module: nn.Module) -> nn.Module:
    """reference: tqdne/nn.py:59-63."""
    with torch.no_grad():
        for p in module.parameters():
            p.zero_()
    return module


def append_dims(x: torch.Tensor, target_dims: int) -> torch.Tensor:
    """reference: tqdne/nn.py:78-83."""
    extra = target_dims - x.ndim
    if extra < 0:
        raise ValueError(f"input has {x.ndim} dims but target_dims is {target_dims}, which is less")
    return x[(...,) + (None,) * extra]
