"""Network building blocks: parameter containers with the reference's attribute names.

State-dict keys, shapes and default initialisation follow tqdne/blocks.py so that reference `.ckpt`
files load with `strict=True`; Encoder / Decoder run through `tqdne_b200.lowering` (CUDA kernel plans).
"""

from __future__ import annotations

import torch
from torch import nn

from .nn import EngineOnly, conv_nd, normalization, zero_module


class GaussianFourierProjection(EngineOnly):
    """Frozen random Fourier features of the noise level (reference: tqdne/blocks.py:15-26)."""

    def __init__(self, channels: int, scale: float = 0.02) -> None:
        super().__init__()
        self.W = nn.Parameter(torch.randn(channels // 2) * scale, requires_grad=False)


class Upsample(EngineOnly):
    """Nearest x2 followed by a conv (reference: tqdne/blocks.py:29-66)."""

    def __init__(self, channels, use_conv, dims=2, out_channels=None, kernel_size=3):
        super().__init__()
        self.channels, self.out_channels = channels, out_channels or channels
        self.use_conv, self.dims = use_conv, dims
        if use_conv:
            self.conv = conv_nd(dims, self.channels, self.out_channels, kernel_size, padding="same")


class Downsample(EngineOnly):
    """Stride-2 conv, k=3 (reference: tqdne/blocks.py:69-108)."""

    def __init__(self, channels, use_conv, dims=2, out_channels=None, kernel_size=3):
        super().__init__()
        self.channels, self.out_channels = channels, out_channels or channels
        self.use_conv, self.dims = use_conv, dims
        if use_conv:
            self.op = conv_nd(dims, self.channels, self.out_channels, kernel_size, stride=2, padding=kernel_size // 2)
        else:   # conv_resample=False: parameter-free average pool (reference: blocks.py:102-104)
            assert self.channels == self.out_channels
            self.op = nn.AvgPool1d(2, 2) if dims == 1 else nn.AvgPool2d(2, 2)


class QKVAttention(EngineOnly):
    """Marker for the attention core (reference: tqdne/blocks.py:148-190); no parameters."""

    def __init__(self, n_heads, use_causal_mask=False):
        super().__init__()
        self.n_heads = n_heads
        self.use_causal_mask = use_causal_mask


class AttentionBlock(EngineOnly):
    """GroupNorm -> 1x1 qkv -> multi-head attention -> zero-init 1x1 projection, residual
    (reference: tqdne/blocks.py:111-145)."""

    def __init__(self, channels, num_heads=1, use_checkpoint=False, flash_attention=True, dims=2, use_causal_mask=False):
        super().__init__()
        self.channels, self.num_heads = channels, num_heads
        self.use_checkpoint = use_checkpoint
        self.norm = normalization(channels)
        self.qkv = conv_nd(dims, channels, channels * 3, 1)
        # the reference's flash_attention=True branch needs flash-attn v1 and is dead code; both settings
        # lower to the same kernels here.  Like the reference, only the flash_attention=False core takes the causal mask
        # (blocks.py:136-139: QKVFlashAttention is built without it).
        self.attention = QKVAttention(num_heads, use_causal_mask=use_causal_mask and not flash_attention)
        self.proj_out = zero_module(conv_nd(dims, channels, channels, 1))


class ResBlock(EngineOnly):
    """Embedding-free residual block of the autoencoder (reference: tqdne/blocks.py:233-260)."""

    def __init__(self, channels, dropout, out_channels=None, kernel_size=3, dims=2):
        super().__init__()
        out_channels = out_channels or channels
        self.in_layers = nn.Sequential(
            normalization(channels), nn.SiLU(), conv_nd(dims, channels, out_channels, kernel_size, padding="same"))
        self.out_layers = nn.Sequential(
            normalization(out_channels), nn.SiLU(), nn.Dropout(p=dropout),
            zero_module(conv_nd(dims, out_channels, out_channels, kernel_size, padding="same")))
        self.skip_connection = nn.Identity() if out_channels == channels else conv_nd(dims, channels, out_channels, 1)


def _resample_levels(channel_mult):
    return len(channel_mult) - 1


class _Coder(nn.Module):
    """Shared engine plumbing of Encoder and Decoder: plan cache keyed by (batch, spatial, dtype)."""

    dims: int

    def _run(self, x: torch.Tensor, kind: str) -> torch.Tensor:
        from . import lowering  # local import: lowering imports this module

        return lowering.run_coder(self, x, kind)

    def _apply(self, fn, *args, **kwargs):
        # parameters moved / cast: cached plans and packed weights point at stale storage
        self.__dict__.pop("_tq_cache", None)
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self.__dict__.pop("_tq_cache", None)
        return super().load_state_dict(*args, **kwargs)


class Encoder(_Coder):
    """reference: tqdne/blocks.py:263-348 (same constructor signature and attribute names)."""

    def __init__(self, in_channels, model_channels, out_channels, num_res_blocks, attention_resolutions=(8, 16, 32),
                 dropout=0, channel_mult=(1, 2, 4, 8), conv_kernel_size=3, conv_resample=True, dims=2, num_heads=1,
                 flash_attention=True):
        super().__init__()
        self.dims = dims
        ch = int(channel_mult[0] * model_channels)
        self.input_layer = conv_nd(dims, in_channels, ch, conv_kernel_size, padding="same")
        stages, ds = [], 1
        for level, mult in enumerate(channel_mult):
            width = int(mult * model_channels)
            for _ in range(num_res_blocks):
                stages.append(ResBlock(ch, dropout, out_channels=width, kernel_size=conv_kernel_size, dims=dims))
                ch = width
                if ds in attention_resolutions:
                    stages.append(AttentionBlock(ch, num_heads=num_heads, dims=dims, flash_attention=flash_attention))
            if level != _resample_levels(channel_mult):
                stages.append(Downsample(ch, conv_resample, dims=dims, out_channels=ch))
                ds *= 2
        self.down_blocks = nn.Sequential(*stages)
        self.output_layer = conv_nd(dims, ch, out_channels, conv_kernel_size, padding="same")

    def forward(self, x):
        return self._run(x, "encoder")


class Decoder(_Coder):
    """reference: tqdne/blocks.py:351-436 (same constructor signature and attribute names)."""

    def __init__(self, in_channels, model_channels, out_channels, num_res_blocks, attention_resolutions=(8, 16, 32),
                 dropout=0, channel_mult=(1, 2, 4, 8), conv_kernel_size=3, conv_resample=True, dims=2, num_heads=1,
                 flash_attention=True):
        super().__init__()
        self.dims = dims
        ch = int(channel_mult[-1] * model_channels)
        self.input_layer = conv_nd(dims, in_channels, ch, conv_kernel_size, padding="same")
        stages, ds = [], 2 ** _resample_levels(channel_mult)
        for level in range(len(channel_mult) - 1, -1, -1):
            width = int(channel_mult[level] * model_channels)
            if level != _resample_levels(channel_mult):
                # note: like the reference, the resampler keeps its default k=3
                stages.append(Upsample(ch, conv_resample, dims=dims, out_channels=ch))
                ds //= 2
            for _ in range(num_res_blocks):
                stages.append(ResBlock(ch, dropout, out_channels=width, kernel_size=conv_kernel_size, dims=dims))
                ch = width
                if ds in attention_resolutions:
                    stages.append(AttentionBlock(ch, num_heads=num_heads, dims=dims, flash_attention=flash_attention))
        self.up_blocks = nn.Sequential(*stages)
        self.output_layer = conv_nd(dims, ch, out_channels, conv_kernel_size, padding="same")

    def forward(self, x):
        return self._run(x, "decoder")
