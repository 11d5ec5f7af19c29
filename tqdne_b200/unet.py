"""Conditional UNet denoiser: same constructor, attribute tree and state-dict keys as tqdne/unet.py,
forward pass executed by the sm_100a kernel plan built in `tqdne_b200.lowering`."""

from __future__ import annotations

import torch
from torch import nn

from .blocks import AttentionBlock, Downsample, GaussianFourierProjection, Upsample
from .nn import EngineOnly, conv_nd, normalization, zero_module


class TimestepBlock(EngineOnly):
    """Marker: a block that consumes the timestep/conditioning embedding (reference: unet.py:15-24)."""


class TimestepEmbedSequential(nn.Sequential, TimestepBlock):
    """Container whose children may take the embedding (reference: unet.py:27-39)."""

    def forward(self, *args, **kwargs):  # pragma: no cover - guard
        return EngineOnly.forward(self, *args, **kwargs)


class ResBlock(TimestepBlock):
    """GN-SiLU-conv, + Linear(SiLU(emb)), GN-SiLU-(dropout)-conv(zero init), + skip
    (reference: tqdne/unet.py:42-143)."""

    def __init__(self, channels, emb_channels, dropout, out_channels=None, kernel_size=3, use_conv=False,
                 use_scale_shift_norm=False, dims=2, use_checkpoint=False):
        super().__init__()
        out_channels = out_channels or channels
        self.use_conv, self.use_checkpoint, self.use_scale_shift_norm = use_conv, use_checkpoint, use_scale_shift_norm
        self.in_layers = nn.Sequential(
            normalization(channels), nn.SiLU(), conv_nd(dims, channels, out_channels, kernel_size, padding="same"))
        self.emb_layers = nn.Sequential(
            nn.SiLU(), nn.Linear(emb_channels, 2 * out_channels if use_scale_shift_norm else out_channels))
        self.out_layers = nn.Sequential(
            normalization(out_channels), nn.SiLU(), nn.Dropout(p=dropout),
            zero_module(conv_nd(dims, out_channels, out_channels, kernel_size, padding="same")))
        if out_channels == channels:
            self.skip_connection = nn.Identity()
        elif use_conv:
            self.skip_connection = conv_nd(dims, channels, out_channels, kernel_size, padding="same")
        else:
            self.skip_connection = conv_nd(dims, channels, out_channels, 1)


class UNetModel(nn.Module):
    """reference: tqdne/unet.py:146-398 -- identical signature; `forward(x, timesteps, cond)` takes and returns
    [N, C, ...] tensors on a CUDA device."""

    def __init__(self, in_channels, model_channels, out_channels, num_res_blocks, attention_resolutions=(8, 16, 32),
                 dropout=0, channel_mult=(1, 2, 4, 8), conv_kernel_size=3, conv_resample=True, dims=2,
                 cond_features=None, cond_emb_scale=None, use_checkpoint=False, num_heads=1,
                 use_scale_shift_norm=False, flash_attention=True, use_causal_mask=False):
        super().__init__()
        self.in_channels, self.out_channels, self.model_channels = in_channels, out_channels, model_channels
        self.dims, self.num_heads = dims, num_heads
        emb_dim = model_channels * 4
        self.time_embed = GaussianFourierProjection(model_channels)
        self.time_mlp = nn.Sequential(nn.Linear(model_channels, emb_dim), nn.SiLU(), nn.Linear(emb_dim, emb_dim))
        self.cond_features = cond_features
        if cond_features is not None:
            self.cond_embed = None
            if cond_emb_scale is not None:
                # reference unet.py:217-219, 386-387.  GaussianFourierProjection.forward broadcasts x[:, None] * W[None, :]
                # (blocks.py:23): with a [N, F] conditioning tensor that only runs for F = 1, so that is what is lowered.
                if cond_features != 1:
                    raise NotImplementedError("tqdne_b200: cond_emb_scale with more than one conditioning feature does not run "
                                              "in the reference either (GaussianFourierProjection broadcasts [N, 1, F] * [1, C/2])")
                self.cond_embed = GaussianFourierProjection(model_channels, cond_emb_scale)
                cond_features = cond_features * model_channels
            self.cond_mlp = nn.Sequential(nn.Linear(cond_features, emb_dim), nn.SiLU(), nn.Linear(emb_dim, emb_dim))

        def res(cin, cout):
            return ResBlock(cin, emb_dim, dropout, out_channels=cout, kernel_size=conv_kernel_size, dims=dims,
                            use_checkpoint=use_checkpoint, use_scale_shift_norm=use_scale_shift_norm)

        def attn(c):
            return AttentionBlock(c, num_heads=num_heads, dims=dims, use_checkpoint=use_checkpoint,
                                  flash_attention=flash_attention, use_causal_mask=use_causal_mask)

        ch = stem = int(channel_mult[0] * model_channels)
        self.input_blocks = nn.ModuleList(
            [TimestepEmbedSequential(conv_nd(dims, in_channels, ch, conv_kernel_size, padding="same"))])
        skip_widths, ds = [ch], 1
        last = len(channel_mult) - 1
        for level, mult in enumerate(channel_mult):
            width = int(mult * model_channels)
            for _ in range(num_res_blocks):
                layers = [res(ch, width)]
                ch = width
                if ds in attention_resolutions:
                    layers.append(attn(ch))
                self.input_blocks.append(TimestepEmbedSequential(*layers))
                skip_widths.append(ch)
            if level != last:
                # like the reference, Downsample keeps its default k=3 even when conv_kernel_size=5
                self.input_blocks.append(TimestepEmbedSequential(Downsample(ch, conv_resample, dims=dims, out_channels=ch)))
                skip_widths.append(ch)
                ds *= 2
        self.middle_block = TimestepEmbedSequential(res(ch, ch), attn(ch), res(ch, ch))
        self.output_blocks = nn.ModuleList([])
        for level in range(last, -1, -1):
            width = int(model_channels * channel_mult[level])
            for i in range(num_res_blocks + 1):
                layers = [res(ch + skip_widths.pop(), width)]
                ch = width
                if ds in attention_resolutions:
                    layers.append(attn(ch))
                if level and i == num_res_blocks:
                    layers.append(Upsample(ch, conv_resample, dims=dims, out_channels=ch, kernel_size=conv_kernel_size))
                    ds //= 2
                self.output_blocks.append(TimestepEmbedSequential(*layers))
        self.out = nn.Sequential(
            normalization(ch), nn.SiLU(),
            zero_module(conv_nd(dims, stem, out_channels, conv_kernel_size, padding="same")))

    # -- engine plumbing ---------------------------------------------------------------------------
    def _apply(self, fn, *args, **kwargs):
        self.__dict__.pop("_tq_cache", None)
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self.__dict__.pop("_tq_cache", None)
        return super().load_state_dict(*args, **kwargs)

    def forward(self, x, timesteps, cond=None):
        """x:[N, C, ...] , timesteps:[N], cond:[N, cond_features] or None -> [N, C_out, ...] (x.dtype)."""
        assert (cond is not None) == (self.cond_features is not None), \
            "must specify cond if and only if the model is conditioned"
        from . import lowering

        return lowering.unet_forward(self, x, timesteps, cond)
