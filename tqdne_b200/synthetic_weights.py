"""Deterministic synthetic weights, independent of module construction order.

No checkpoints ship with the reference (weights/ is an empty LFS placeholder), and its zero-initialised
modules (zero_module: tqdne/unet.py:102,357, tqdne/blocks.py:134,249) make a default-initialised network
output exactly 0.  Every tensor is therefore drawn from its own generator seeded by (seed, key), so the same
state dict can be rebuilt anywhere (build container, GPU box) and loaded into the reference modules, the
oracle restatement and the engine alike.
"""

from __future__ import annotations

import zlib

import torch


def _gen(seed: int, key: str) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((seed * 2654435761 + zlib.crc32(key.encode())) % (2**63 - 1))
    return g


def seeded_state_dict(shapes: dict, seed: int = 0) -> dict:
    """shapes: {key: torch.Size}.  Conv / linear weights ~ N(0, 1/fan_in) (activations stay O(1) through
    ~80 layers), norm weights ~ 1 + 0.1 N(0,1), biases ~ 0.1 N(0,1), Fourier frequencies ~ 0.02 N(0,1)."""
    out = {}
    for key, shape in shapes.items():
        shape = tuple(shape)
        g = _gen(seed, key)
        r = torch.randn(shape, generator=g, dtype=torch.float32)
        if key.endswith("time_embed.W") or key.endswith("cond_embed.W"):
            t = 0.02 * r
        elif key.endswith(".bias"):
            t = 0.1 * r
        elif len(shape) == 1:  # normalisation scale
            t = 1.0 + 0.1 * r
        else:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            t = r / fan_in**0.5
        out[key] = t
    return out


def shapes_of(module) -> dict:
    return {k: v.shape for k, v in module.state_dict().items()}
