"""Walk a tqdne network (UNetModel / Encoder / Decoder parameter tree) and emit its kernel plan.

Reference forward passes being lowered: UNetModel.forward (tqdne/unet.py:360-398), ResBlock._forward
(unet.py:131-143), AttentionBlock._forward (blocks.py:139-145), blocks.ResBlock.forward (blocks.py:256-260),
Encoder/Decoder.forward (blocks.py:344-348, 432-436).
"""

from __future__ import annotations

import torch
from torch import nn

from . import blocks as B
from . import unet as U
from .engine import Act, Plan, device_guard, nchw_to_nhwc, nhwc_to_nchw, pack_conv, require_cuda

MODES = {torch.bfloat16: "bf16", torch.float32: "fp32"}


def _cache(model: nn.Module) -> dict:
    c = model.__dict__.get("_tq_cache")
    if c is None:
        c = {"packed": {}, "plans": {}}
        model.__dict__["_tq_cache"] = c
    return c


def invalidate(model: nn.Module) -> None:
    model.__dict__.pop("_tq_cache", None)


def _param_stamp(model: nn.Module) -> int:
    return sum(p._version for p in model.parameters())


class Lowerer:
    """Emits ops into a Plan; owns the packed-weight cache of one model for one activation dtype."""

    def __init__(self, model: nn.Module, plan: Plan, dims: int):
        self.model, self.plan, self.dims = model, plan, dims
        self.packed = _cache(model)["packed"].setdefault(plan.act_dtype, {})
        self.flops = 0

    # -- weights --------------------------------------------------------------------------------
    def _conv_flops(self, conv: nn.Module, out: Act) -> None:
        w = conv.weight
        self.flops += 2 * out.N * out.H * out.W * w.shape[0] * w[0].numel()

    # -- blocks ---------------------------------------------------------------------------------
    def conv(self, conv: nn.Module, srcs: list[Act], *, segments=None, shortcut: nn.Module | None = None,
             shortcut_srcs: list[Act] | None = None, **kw) -> Act:
        """`segments`: real input channels per source (the stem source is zero padded to 64)."""
        stride = conv.stride[0]
        segments = list(segments) if segments is not None else [a.C for a in srcs]
        sc_segments = [a.C for a in shortcut_srcs] if shortcut is not None else None
        key = (id(conv), tuple(segments), id(shortcut) if shortcut is not None else None,
               tuple(sc_segments) if sc_segments else None)
        pc = self.packed.get(key)
        if pc is None:
            sc = (shortcut.weight, shortcut.bias, sc_segments) if shortcut is not None else None
            pc = pack_conv(conv.weight, conv.bias, segments, self.plan.act_dtype, shortcut=sc)
            self.packed[key] = pc
        out = self.plan.conv(pc, srcs, stride=stride, dims=self.dims,
                             shortcut_srcs=shortcut_srcs if shortcut is not None else None, **kw)
        self._conv_flops(conv, out)
        if shortcut is not None:
            self._conv_flops(shortcut, out)
        return out

    def resblock(self, blk: nn.Module, srcs: list[Act], emb: torch.Tensor | None, emb_off: int, emb_ld: int,
                 free_inputs: bool) -> Act:
        """GN+SiLU -> conv (+emb) -> GN+SiLU -> conv (+ identity residual | fused 1x1 shortcut)."""
        plan = self.plan
        gn1, conv1 = blk.in_layers[0], blk.in_layers[2]
        gn2, conv2 = blk.out_layers[0], blk.out_layers[3]
        h0 = plan.groupnorm(srcs, gn1.weight, gn1.bias, silu=True)
        e = emb[:, emb_off:] if emb is not None else None
        if getattr(blk, "use_scale_shift_norm", False):
            # FiLM (unet.py:135-139): h = out_norm(conv1(h0)) * (1 + scale) + shift, (scale, shift) = the two halves of the
            # block's 2 * C embedding columns -- folded into the second norm's per-sample affine, not added after conv1
            h1 = self.conv(conv1, [h0], stats=True)
            plan.release(h0)
            h2 = plan.groupnorm([h1], gn2.weight, gn2.bias, silu=True, film=e, film_ld=emb_ld)
        else:
            h1 = self.conv(conv1, [h0], emb=e, emb_ld=emb_ld, stats=True)
            plan.release(h0)
            h2 = plan.groupnorm([h1], gn2.weight, gn2.bias, silu=True)
        plan.release(h1)
        skip = blk.skip_connection
        if isinstance(skip, nn.Identity):
            assert len(srcs) == 1
            out = self.conv(conv2, [h2], residual=srcs[0], stats=True)
        else:
            if skip.weight[0, 0].numel() != 1:
                raise NotImplementedError("tqdne_b200: ResBlock(use_conv=True) spatial shortcut is not lowered")
            # out = conv2(h2) + skip_1x1(cat(srcs)): one GEMM, K = taps*Cout + sum(Cin segments)
            out = self.conv(conv2, [h2], shortcut=skip, shortcut_srcs=srcs, stats=True)
        plan.release(h2)
        if free_inputs:
            for a in srcs:
                plan.release(a)
        return out

    def attention(self, blk: B.AttentionBlock, x: Act, free_input: bool) -> Act:
        plan = self.plan
        g = plan.groupnorm([x], blk.norm.weight, blk.norm.bias, silu=False)
        qkv = self.conv(blk.qkv, [g])
        plan.release(g)
        heads = blk.num_heads
        a = plan.attention(qkv, heads, causal=bool(getattr(blk.attention, "use_causal_mask", False)))
        self.flops += 4 * x.N * x.P * x.P * x.C  # QK^T and PV, 2*MAC each
        plan.release(qkv)
        out = self.conv(blk.proj_out, [a], residual=x, stats=True)
        plan.release(a)
        if free_input:
            plan.release(x)
        return out

    def downsample(self, blk: B.Downsample, x: Act) -> Act:
        if not blk.use_conv:   # conv_resample=False: nn.AvgPool{1,2}d(2, 2) (blocks.py:104)
            return self.plan.resample2(x, "avg_pool")
        return self.conv(blk.op, [x], stats=True)

    def upsample(self, blk: B.Upsample, x: Act, free_input: bool) -> Act:
        if not blk.use_conv:   # conv_resample=False: F.interpolate(scale_factor=2, mode="nearest") alone (blocks.py:59-62)
            out = self.plan.resample2(x, "nearest")
            if free_input:
                self.plan.release(x)
            return out
        out = self.conv(blk.conv, [x], upsample=True, stats=True)
        if free_input:
            self.plan.release(x)
        return out


# ------------------------------------------------------------------------------------------------
# UNet
# ------------------------------------------------------------------------------------------------
class UNetPlan:
    """One denoiser call for a fixed (batch, spatial size, dtype): F = UNet(xin, t, cond)."""

    def __init__(self, model: U.UNetModel, N: int, spatial: tuple, act_dtype: torch.dtype, uniform_t: bool,
                 cond_channels: int = 0):
        """`cond_channels` > 0: the trailing input channels of the UNet are a conditioning signal that stays fixed over
        the denoiser calls of a sample() (LightningEDM.forward: th.cat((sample_in, cond_sample), dim=1), edm.py:109).  It
        lives in a tensor of its own (`xcond`) and the stem convolution reads it as a second concat segment, so the
        sampler kernels keep writing a state-only `xin`."""
        dev = next(model.parameters()).device
        require_cuda(next(model.parameters()), "UNetModel parameters")
        self.N, self.spatial, self.act_dtype, self.uniform_t = N, tuple(spatial), act_dtype, uniform_t
        H, W = (spatial if len(spatial) == 2 else (1, spatial[0]))
        self.H, self.W = H, W
        plan = Plan(dev, act_dtype)
        low = Lowerer(model, plan, model.dims)
        self.plan, self.low = plan, low
        mc = model.model_channels
        E = 4 * mc
        f32 = dict(device=dev, dtype=torch.float32)
        rows_t = 1 if uniform_t else N
        self.t = torch.zeros(rows_t, **f32)
        assert 0 <= cond_channels < model.in_channels
        self.cond_channels = cond_channels
        state_channels = model.in_channels - cond_channels
        self.cin_pad = (state_channels + 63) // 64 * 64
        self.xin = Act(torch.zeros(N * H * W * self.cin_pad, device=dev, dtype=act_dtype), N, H, W, self.cin_pad)
        self.xcond = None
        stem_srcs, stem_segments = [self.xin], [state_channels]
        if cond_channels:
            cpad = (cond_channels + 63) // 64 * 64
            self.xcond = Act(torch.zeros(N * H * W * cpad, device=dev, dtype=act_dtype), N, H, W, cpad)
            stem_srcs, stem_segments = [self.xin, self.xcond], [state_channels, cond_channels]

        # ---- conditioning prologue (constant over the NFE calls of one sample()): cond_mlp ----
        self.cond = None
        self.cond_plan = None
        cemb = None
        if model.cond_features is not None:
            self.cond = torch.zeros(N, model.cond_features, **f32)
            cp = Plan(dev, act_dtype)
            c1 = torch.empty(N, E, **f32)
            cemb = torch.empty(N, E, **f32)
            l0, l2 = model.cond_mlp[0], model.cond_mlp[2]
            cin = self.cond
            if getattr(model, "cond_embed", None) is not None:
                # Fourier-embedded conditioning (unet.py:386-387): one feature per sample -> [N, model_channels]
                cin = torch.empty(N, mc, **f32)
                cp.fourier(self.cond.view(-1), model.cond_embed.W.detach().float().contiguous(), N, cin)
            cp.linear(cin, l0.weight.detach().float().contiguous(), l0.bias.detach().float().contiguous(), N, y=c1)
            cp.linear(c1, l2.weight.detach().float().contiguous(), l2.bias.detach().float().contiguous(), N, act_in=True,
                      y=cemb)
            self.cond_plan = cp
            low.flops += 2 * N * (model.cond_features * E + E * E)
        # ---- per-call embedding: Fourier -> time_mlp -> (+cond) -> SiLU -> all emb_layers in one GEMM ----
        feat = torch.empty(rows_t, mc, **f32)
        h1 = torch.empty(rows_t, E, **f32)
        emb_act = torch.empty(N, E, device=dev, dtype=act_dtype)
        plan.fourier(self.t, model.time_embed.W.detach().float().contiguous(), rows_t, feat)
        t0, t2 = model.time_mlp[0], model.time_mlp[2]
        plan.linear(feat, t0.weight.detach().float().contiguous(), t0.bias.detach().float().contiguous(), rows_t, y=h1)
        plan.linear(h1, t2.weight.detach().float().contiguous(), t2.bias.detach().float().contiguous(), N, x_rows=rows_t,
                    act_in=True, add=cemb, add_rows=N, y_act=emb_act)
        low.flops += 2 * N * (mc * E + E * E)
        res_blocks = [m for m in model.modules() if isinstance(m, U.ResBlock)]
        key = ("emb_all",)
        pc = low.packed.get(key)
        if pc is None:
            Wall = torch.cat([r.emb_layers[1].weight.detach().float() for r in res_blocks], dim=0)
            ball = torch.cat([r.emb_layers[1].bias.detach().float() for r in res_blocks], dim=0)
            pc = pack_conv(Wall, ball, [E], act_dtype)
            low.packed[key] = pc
        self.emb_off = {}
        off = 0
        for r in res_blocks:
            self.emb_off[id(r)] = off
            off += r.emb_layers[1].weight.shape[0]
        emb_in = Act(emb_act, N, 1, 1, E)
        emb_all = plan.conv(pc, [emb_in], out_dtype=torch.float32)
        low.flops += 2 * N * E * off
        self.emb_all = emb_all.t.view(N, off)
        emb_ld = off

        def run_seq(seq, srcs: list[Act], free_inputs: bool) -> Act:
            cur = srcs
            first = True
            for layer in seq:
                fi = free_inputs or not first
                if isinstance(layer, U.ResBlock):
                    o = low.resblock(layer, cur, self.emb_all, self.emb_off[id(layer)], emb_ld, fi)
                elif isinstance(layer, B.AttentionBlock):
                    o = low.attention(layer, cur[0], fi)
                elif isinstance(layer, B.Downsample):
                    o = low.downsample(layer, cur[0])
                    if fi:
                        plan.release(cur[0])
                elif isinstance(layer, B.Upsample):
                    o = low.upsample(layer, cur[0], fi)
                elif isinstance(layer, (nn.Conv1d, nn.Conv2d)):   # the stem
                    o = low.conv(layer, cur, segments=stem_segments, stats=True)
                else:
                    raise NotImplementedError(f"tqdne_b200: cannot lower {type(layer).__name__}")
                cur = [o]
                first = False
            return cur[0]

        # ---- input blocks: every output is a skip, kept alive until its output block consumes it ----
        hs: list[Act] = []
        h = None
        for i, blk in enumerate(model.input_blocks):
            h = run_seq(blk, stem_srcs if i == 0 else [h], free_inputs=False)
            hs.append(h)
        h = run_seq(model.middle_block, [h], free_inputs=False)
        for blk in model.output_blocks:
            skip = hs.pop()
            # th.cat([h, skip], dim=1): h first, skip second
            prev = h
            h = run_seq(blk, [prev, skip], free_inputs=False)
            plan.release(prev)
            plan.release(skip)
        gn_o, conv_o = model.out[0], model.out[2]
        g = plan.groupnorm([h], gn_o.weight, gn_o.bias, silu=True)
        plan.release(h)
        out = low.conv(conv_o, [g], out_dtype=torch.float32)
        plan.release(g)
        self.out = out  # Act [N,H,W,Cout] fp32
        self.flops = low.flops

    def set_cond(self, cond: torch.Tensor | None) -> None:
        if self.cond is not None:
            self.cond.copy_(cond.to(torch.float32))
            self.cond_plan.run()

    def set_cond_sample(self, cond_sample: torch.Tensor) -> None:
        """[N, cond_channels, ...] -> the channels-last second stem source (once per sample(), not per call)."""
        assert self.xcond is not None and cond_sample.shape[1] == self.cond_channels
        self.xcond.t.copy_(nchw_to_nhwc(cond_sample.to(torch.float32), self.act_dtype, self.xcond.C).view(-1))

    def set_t(self, t: torch.Tensor) -> None:
        self.t.copy_(t.to(torch.float32).reshape(-1)[: self.t.numel()])

    def run(self) -> None:
        self.plan.run()

    @property
    def launches_per_call(self) -> int:
        return self.plan.num_ops


def _act_dtype_of(model: nn.Module) -> torch.dtype:
    dt = next(model.parameters()).dtype
    if dt not in MODES:
        raise RuntimeError(f"tqdne_b200: parameters must be float32 (fp32 parity mode) or bfloat16 (tensor mode), got {dt}")
    return dt


def get_unet_plan(model: U.UNetModel, N: int, spatial: tuple, uniform_t: bool, act_dtype: torch.dtype | None = None,
                  cond_channels: int = 0) -> UNetPlan:
    c = _cache(model)
    stamp = _param_stamp(model)
    if c.get("stamp") != stamp:
        invalidate(model)
        c = _cache(model)
        c["stamp"] = stamp
    act_dtype = act_dtype or getattr(model, "engine_dtype", None) or _act_dtype_of(model)
    key = ("unet", N, tuple(spatial), act_dtype, uniform_t, cond_channels)
    p = c["plans"].get(key)
    if p is None:
        p = UNetPlan(model, N, spatial, act_dtype, uniform_t, cond_channels)
        c["plans"][key] = p
    return p


def unet_forward(model: U.UNetModel, x: torch.Tensor, timesteps: torch.Tensor, cond: torch.Tensor | None) -> torch.Tensor:
    """Drop-in UNetModel.forward: NCHW/NCL in, NCHW/NCL out (same dtype as x)."""
    require_cuda(x, "x")
    N, Cc = x.shape[0], x.shape[1]
    spatial = tuple(x.shape[2:])
    assert len(spatial) == model.dims, f"expected {model.dims} spatial dims"
    with device_guard(x.device):
        p = get_unet_plan(model, N, spatial, uniform_t=False)
        xin = nchw_to_nhwc(x.to(torch.float32), p.act_dtype, p.cin_pad)
        p.xin.t.copy_(xin.view(-1))
        p.set_t(timesteps)
        p.set_cond(cond)
        p.run()
        return nhwc_to_nchw(p.out.t, N, model.out_channels, spatial, model.out_channels, x.dtype)


# ------------------------------------------------------------------------------------------------
# Encoder / Decoder
# ------------------------------------------------------------------------------------------------
class CoderPlan:
    def __init__(self, net: nn.Module, kind: str, N: int, spatial: tuple, act_dtype: torch.dtype):
        dev = next(net.parameters()).device
        require_cuda(next(net.parameters()), f"{kind} parameters")
        self.N, self.spatial, self.act_dtype = N, tuple(spatial), act_dtype
        H, W = (spatial if len(spatial) == 2 else (1, spatial[0]))
        plan = Plan(dev, act_dtype)
        low = Lowerer(net, plan, net.dims)
        cin = net.input_layer.weight.shape[1]
        self.cin, self.cin_pad = cin, (cin + 63) // 64 * 64
        self.xin = Act(torch.zeros(N * H * W * self.cin_pad, device=dev, dtype=act_dtype), N, H, W, self.cin_pad)
        h = low.conv(net.input_layer, [self.xin], segments=[cin], stats=True)
        seq = net.down_blocks if kind == "encoder" else net.up_blocks
        for layer in seq:
            if isinstance(layer, B.ResBlock):
                h = low.resblock(layer, [h], None, 0, 0, free_inputs=True)
            elif isinstance(layer, B.AttentionBlock):
                h = low.attention(layer, h, free_input=True)
            elif isinstance(layer, B.Downsample):
                o = low.downsample(layer, h)
                plan.release(h)
                h = o
            elif isinstance(layer, B.Upsample):
                h = low.upsample(layer, h, free_input=True)
            else:
                raise NotImplementedError(f"tqdne_b200: cannot lower {type(layer).__name__}")
        out = low.conv(net.output_layer, [h], out_dtype=torch.float32)
        plan.release(h)
        self.out, self.plan, self.flops = out, plan, low.flops
        self.cout = net.output_layer.weight.shape[0]

    def run(self) -> None:
        self.plan.run()


def get_coder_plan(net: nn.Module, kind: str, N: int, spatial: tuple, act_dtype: torch.dtype | None = None) -> CoderPlan:
    c = _cache(net)
    stamp = _param_stamp(net)
    if c.get("stamp") != stamp:
        invalidate(net)
        c = _cache(net)
        c["stamp"] = stamp
    act_dtype = act_dtype or getattr(net, "engine_dtype", None) or _act_dtype_of(net)
    key = (kind, N, tuple(spatial), act_dtype)
    p = c["plans"].get(key)
    if p is None:
        p = CoderPlan(net, kind, N, spatial, act_dtype)
        c["plans"][key] = p
    return p


def run_coder(net: nn.Module, x: torch.Tensor, kind: str) -> torch.Tensor:
    """Drop-in Encoder/Decoder.forward: [N, C, ...] -> [N, C_out, ...] (x.dtype)."""
    require_cuda(x, "x")
    N = x.shape[0]
    spatial = tuple(x.shape[2:])
    with device_guard(x.device):
        p = get_coder_plan(net, kind, N, spatial)
        xin = nchw_to_nhwc(x.to(torch.float32), p.act_dtype, p.cin_pad)
        p.xin.t.copy_(xin.view(-1))
        p.run()
        so = (p.out.H, p.out.W) if len(spatial) == 2 else (p.out.W,)
        return nhwc_to_nchw(p.out.t, N, p.cout, so, p.cout, x.dtype)
