"""Training step of the 1D EDM UNet on the B200 engine (BASELINE.json configs[4]).

Reference: `LightningEDM.step` + `configure_optimizers` (tqdne/edm.py:115-134,240-251), the EMA callback
(tqdne/ema.py:24-28) and what `loss.backward()` does through `UNetModel.forward` (tqdne/unet.py:360-398).

The step is a static tape built once per (model, batch, length): forward ops (the same tcgen05 / GroupNorm / attention
kernels as sampling, every activation kept), then the backward ops in reverse order:

  convolution     input gradient = the forward igemm over dY with tap-flipped transposed weights (second gradient paths
                  are added in its epilogue); weight / bias gradient = `tq_conv1d_wgrad` (one call per concat source);
                  stride-2 (Downsample): dY zero-stuffed onto the stride-1 grid; upsampled input: weight gradient
                  against the materialised nearest-x2 input, input gradient pair-summed
  GroupNorm+SiLU  `tq_gn_silu_backward` over the virtual concat, second gradient path fused (`dx_add`)
  attention       `tq_attention_backward`
  embeddings      per-sample channel sums of the first convolution's output gradient, then `tq_linear_backward`
  loss            `tq_edm_noise`, `tq_edm_loss`;  optimiser: `tq_adam_ema_step` over ONE flat fp32 parameter buffer

Parameters live in a flat fp32 master buffer in the engine's layout (conv weights [cout, taps, cin], channels padded to
64 where the layer's are: stem input, head output); the bf16 operand copies the kernels read are refreshed from it after
every update.  Gradients are one flat buffer too, so the data-parallel all-reduce is a single NCCL call.
PyTorch is used for allocation, the random draws of the step (sigma, noise) and pure re-layout copies of the weights.
"""

from __future__ import annotations

import ctypes as C
import math
import os

import torch
from torch import nn

from . import _lib
from . import blocks as B
from . import unet as U
from .engine import Act, Plan, current_stream_ptr, pack_conv, tq_dtype

BF16 = torch.bfloat16


def _pad64(v: int) -> int:
    return (v + 63) // 64 * 64


# the gradient all-reduce of everything from this input block on overlaps the backward pass of the blocks before it (index
# 3 = the first Downsample of the shipped UNets: stem + two top-level ResBlocks remain, the longest-running part of the
# backward with the fewest parameters)
TAIL_FROM_INPUT_BLOCK = 3


class _Store:
    """Flat fp32 parameter / gradient / Adam / EMA buffers with engine-layout views per parameter."""

    def __init__(self, model: U.UNetModel, device):
        self.items = {}   # id(param) -> (offset, engine shape, kind)
        self.order = []
        off = 0

        def add(p: nn.Parameter, shape, kind):
            nonlocal off
            n = math.prod(shape)
            self.items[id(p)] = (off, tuple(shape), kind)
            self.order.append(p)
            off += (n + 3) // 4 * 4   # 16 B aligned views

        res_blocks = [m for m in model.modules() if isinstance(m, U.ResBlock)]
        for r in res_blocks:          # emb_layers first and adjacent: ONE dense layer over the concatenated rows
            add(r.emb_layers[1].weight, r.emb_layers[1].weight.shape, "dense")
        self.emb_rows = sum(r.emb_layers[1].weight.shape[0] for r in res_blocks)
        self.emb_w_off = self.items[id(res_blocks[0].emb_layers[1].weight)][0]
        for r in res_blocks:
            add(r.emb_layers[1].bias, r.emb_layers[1].bias.shape, "dense")
        self.emb_b_off = self.items[id(res_blocks[0].emb_layers[1].bias)][0]
        # Everything else in two buckets, each in forward order (convolutions, then the remaining parameters).  The LATE
        # bucket = the layers whose gradients are final early in the backward pass (from the first Downsample on: the deep
        # input blocks, the middle block, the output blocks, the head); it sits at the end of the flat buffers so that its
        # gradient all-reduce is ONE contiguous range that can run while the backward of the first blocks is still going.
        first = min(TAIL_FROM_INPUT_BLOCK, len(model.input_blocks) - 1)
        tail_mods = list(model.input_blocks)[first:] + [model.middle_block, model.output_blocks, model.out]
        tail_ids = {id(p) for m in tail_mods for p in m.parameters()}
        for tail in (False, True):
            if tail:
                self.tail_off = off
            for mod in model.modules():
                if isinstance(mod, nn.Conv1d) and (id(mod.weight) in tail_ids) == tail:
                    O, I, k = mod.weight.shape
                    add(mod.weight, (_pad64(O), k, _pad64(I)), "conv_w")
                    add(mod.bias, (_pad64(O),), "conv_b")
            for p in model.parameters():
                if p.requires_grad and id(p) not in self.items and (id(p) in tail_ids) == tail:
                    add(p, p.shape, "dense")
        self.n = off
        z = lambda: torch.zeros(self.n, device=device, dtype=torch.float32)  # noqa: E731
        self.P, self.G, self.M, self.V, self.EMA = z(), z(), z(), z(), z()
        # training progress lives HERE, not in the shape-specific tapes: optimiser steps taken (the cosine schedule's
        # position and Adam's bias correction), forward passes run (dropout counter), and a version stamp of the masters
        # that tells a tape whether its bf16 operand copies are stale
        self.step_count = 0
        self.pass_count = 0
        self.version = 0

    def view(self, buf: torch.Tensor, p: nn.Parameter) -> torch.Tensor:
        off, shape, _ = self.items[id(p)]
        return buf[off:off + math.prod(shape)].view(shape)

    @torch.no_grad()
    def load_from_module(self):
        for p in self.order:
            off, shape, kind = self.items[id(p)]
            v = self.view(self.P, p)
            v.zero_()
            if kind == "conv_w":
                O, I, k = p.shape
                v[:O, :, :I] = p.detach().float().permute(0, 2, 1)
            elif kind == "conv_b":
                v[:p.shape[0]] = p.detach().float()
            else:
                v.copy_(p.detach().float())
        self.EMA.copy_(self.P)

    @torch.no_grad()
    def to_module_layout(self, buf: torch.Tensor, p: nn.Parameter) -> torch.Tensor:
        """A tensor shaped like the module parameter (PyTorch layout) from an engine-layout buffer."""
        off, shape, kind = self.items[id(p)]
        v = self.view(buf, p)
        if kind == "conv_w":
            O, I, k = p.shape
            return v[:O, :, :I].permute(0, 2, 1).contiguous()
        if kind == "conv_b":
            return v[:p.shape[0]].clone()
        return v.clone()

    @torch.no_grad()
    def store_to_module(self, buf: torch.Tensor | None = None):
        for p in self.order:
            p.copy_(self.to_module_layout(self.P if buf is None else buf, p).to(p.dtype))


class TrainStep1D:
    """One optimisation step of `LightningEDM` with a 1D UNet: loss, gradients, Adam + EMA update."""

    def __init__(self, edm, N: int, L: int, *, lr: float = 1e-4, max_steps: int = 100000, eta_min: float = 0.0,
                 ema_decay: float = 0.999, dropout: float | None = None, betas=(0.9, 0.999), eps: float = 1e-8,
                 store: "_Store | None" = None):
        """`store`: the parameter / Adam / EMA state of an earlier TrainStep1D of the SAME model (another batch size or
        length): the new tape trains on from it instead of restarting from the nn.Module's parameters."""
        model = edm.unet
        assert model.dims == 1, "TrainStep1D: the 1D UNet (config 5); the 2-D weight gradient is not built yet"
        dev = next(model.parameters()).device
        assert dev.type == "cuda", "tqdne_b200: training needs the module on a CUDA device (no CPU path)"
        assert L % 8 == 0, "sequence length must be divisible by 8 (three resampling levels)"
        self.edm, self.model, self.dev, self.N, self.L = edm, model, dev, N, L
        self.lib = _lib.lib()
        self.lr0, self.max_steps, self.eta_min, self.ema_decay, self.betas, self.eps = lr, max_steps, eta_min, ema_decay, betas, eps
        first_res = next(m for m in model.modules() if isinstance(m, U.ResBlock))
        if first_res.use_scale_shift_norm:
            raise NotImplementedError("tqdne_b200: the training tape covers the shipped ResBlock (embedding added after conv1); "
                                      "use_scale_shift_norm models sample on the engine and train with the reference package")
        self.p_drop = float(first_res.out_layers[2].p) if dropout is None else float(dropout)
        self.sigma_data = float(edm.edm.sigma_data)
        # dropout inside the GroupNorm kernels of both passes (default) or as separate multiply passes (A/B switch)
        self.fused_dropout = os.environ.get("TQ_TRAIN_FUSED_DROPOUT", "1") != "0"
        if store is None:
            self.store = _Store(model, dev)
            self.store.load_from_module()
        else:
            self.store = store
        self.copies_version = -1   # store.version the bf16 operand copies of THIS tape were made from
        self._graph = None         # CUDA graph of the tape (None: not captured yet, False: capture failed -> eager)
        self.repack: list = []     # (fp32 master view [Op, k, Ip], forward bf16 copy | None, input-gradient bf16 copy | None, ci_off, Cs)
        self.fwd: list = []
        self.bwd: list = []
        self.masks: list = []
        self._cur: Plan | None = None
        self._keep: list = []
        self._build()
        self.refresh_weights()

    # ------------------------------------------------------------------------------------------ op lists
    def _plan(self, ops: list) -> Plan:
        if self._cur is None or self._cur_list is not ops:
            self._cur = Plan(self.dev, BF16)
            self._cur_list = ops
            ops.append(self._cur.run)
        return self._cur

    def _direct(self, ops: list, fn) -> None:
        self._cur = None
        ops.append(fn)

    def _st(self):
        return current_stream_ptr()

    def _host_seed(self, site: int) -> int:
        return ((self.pass_count << 16) + site + 1) * 0x9E3779B1 & 0xFFFFFFFFFFFF

    def _new(self, N, L, Cc, dtype=BF16) -> Act:
        return Act(torch.empty(N * L * Cc, device=self.dev, dtype=dtype), N, 1, L, Cc)

    # ------------------------------------------------------------------------------------------ forward builders
    def _conv(self, mod: nn.Conv1d, srcs: list[Act], real_in: list[int], **kw) -> Act:
        st = self.store
        pc = pack_conv(mod.weight, mod.bias, real_in, BF16)
        wv, bv = st.view(st.P, mod.weight), st.view(st.P, mod.bias)
        pc.bias = bv
        self.repack.append((wv, pc.weights, None, 0, 0))
        out = self._plan(self.fwd).conv(pc, srcs, dims=1, **kw)
        self._keep.append(pc)
        return out

    def _gn(self, gn: nn.GroupNorm, srcs: list[Act], silu: bool, drop_site: int | None = None) -> Act:
        st = self.store
        kw = dict(drop_seed=self.drop_seed, drop_p=self.p_drop, drop_site=drop_site) if drop_site is not None else {}
        return self._plan(self.fwd).groupnorm(srcs, st.view(st.P, gn.weight), st.view(st.P, gn.bias), silu=silu, **kw)

    def _build(self):
        model, st, N, L, dev = self.model, self.store, self.N, self.L, self.dev
        f32 = dict(device=dev, dtype=torch.float32)
        mc = model.model_channels
        E = 4 * mc
        self.nodes: list = []
        self.drop_of_act: dict = {}  # id(GroupNorm output) -> dropout site fused into it (or None)
        self.drop_seed = torch.zeros(1, device=dev, dtype=torch.int64)
        self.emb_of_act: dict = {}   # id(conv1 output) -> column offset of its ResBlock in e_all / de_all
        self.cin = model.in_channels
        self.cin_pad = _pad64(self.cin)
        # ---- step inputs
        self.y = torch.zeros(N, L, self.cin, **f32)          # clean sample, channels-last
        self.noise = torch.zeros(N, L, self.cin, **f32)
        self.xn = torch.zeros(N, L, self.cin, **f32)
        self.sigma = torch.ones(N, **f32)
        self.t = torch.zeros(N, **f32)
        self.loss = torch.zeros(1, **f32)
        self.xin = Act(torch.zeros(N * L * self.cin_pad, device=dev, dtype=BF16), N, 1, L, self.cin_pad)
        self._direct(self.fwd, lambda: _lib.check(self.lib.tq_edm_noise(
            self.y.data_ptr(), self.noise.data_ptr(), self.sigma.data_ptr(), self.xn.data_ptr(), self.xin.t.data_ptr(), N, L,
            self.cin, self.cin_pad, self.sigma_data, self._st()), "edm_noise"))
        # ---- embeddings: cond_mlp, Fourier -> time_mlp (+cond), all emb_layers as one dense layer
        plan = self._plan(self.fwd)
        pv = lambda p: st.view(st.P, p)  # noqa: E731
        self.cond = None
        cemb = None
        if model.cond_features is not None:
            self.cond = torch.zeros(N, model.cond_features, **f32)
            self.c1, cemb = torch.empty(N, E, **f32), torch.empty(N, E, **f32)
            l0, l2 = model.cond_mlp[0], model.cond_mlp[2]
            plan.linear(self.cond, pv(l0.weight), pv(l0.bias), N, y=self.c1)
            plan.linear(self.c1, pv(l2.weight), pv(l2.bias), N, act_in=True, y=cemb)
        self.feat, self.t1, self.emb = torch.empty(N, mc, **f32), torch.empty(N, E, **f32), torch.empty(N, E, **f32)
        self.Wf = model.time_embed.W.detach().float().contiguous()
        plan.fourier(self.t, self.Wf, N, self.feat)
        t0, t2 = model.time_mlp[0], model.time_mlp[2]
        plan.linear(self.feat, pv(t0.weight), pv(t0.bias), N, y=self.t1)
        plan.linear(self.t1, pv(t2.weight), pv(t2.bias), N, act_in=True, add=cemb, add_rows=N, y=self.emb)
        R = st.emb_rows
        self.Wall = st.P[st.emb_w_off:st.emb_w_off + R * E].view(R, E)
        self.ball = st.P[st.emb_b_off:st.emb_b_off + R]
        self.e_all = torch.empty(N, R, **f32)
        self.de_all = torch.zeros(N, R, **f32)
        plan.linear(self.emb, self.Wall, self.ball, N, act_in=True, y=self.e_all)
        res_blocks = [m for m in model.modules() if isinstance(m, U.ResBlock)]
        self.emb_off, off = {}, 0
        for r in res_blocks:
            self.emb_off[id(r)] = off
            off += r.emb_layers[1].weight.shape[0]

        # ---- the network
        def resblock(blk, srcs: list[Act]) -> Act:
            gn1, conv1 = blk.in_layers[0], blk.in_layers[2]
            gn2, conv2 = blk.out_layers[0], blk.out_layers[3]
            cins = [a.C for a in srcs]
            h0 = self._gn(gn1, srcs, True)
            self.nodes.append(("gn", gn1, srcs, h0, True))
            eo = self.emb_off[id(blk)]
            h1 = self._conv(conv1, [h0], [h0.C], emb=self.e_all[:, eo:], emb_ld=R, stats=True)
            self.nodes.append(("conv", conv1, [h0], h1, {}))
            self.emb_of_act[id(h1)] = eo
            site = None
            if self.p_drop > 0:       # nn.Dropout behind the activation
                site = len(self.masks)
                self.masks.append(site)
            h2 = self._gn(gn2, [h1], True, drop_site=site if self.fused_dropout else None)
            self.nodes.append(("gn", gn2, [h1], h2, True))
            if site is not None and self.fused_dropout:
                self.drop_of_act[id(h2)] = site
            elif site is not None:
                h2d = self._new(N, h2.W, h2.C)
                self._direct(self.fwd, lambda a=h2, o=h2d, site=site: _lib.check(self.lib.tq_dropout_apply(
                    a.t.data_ptr(), o.t.data_ptr(), a.t.numel(), self._host_seed(site), self.p_drop, self._st()), "dropout"))
                self.nodes.append(("mul", None, [h2], h2d, site))
                h2 = h2d
            skip = blk.skip_connection
            if isinstance(skip, nn.Identity):
                res = srcs[0]
            else:
                res = self._conv(skip, srcs, cins, stats=False)
                self.nodes.append(("conv", skip, srcs, res, {}))
            out = self._conv(conv2, [h2], [h2.C], residual=res, stats=True)
            self.nodes.append(("conv", conv2, [h2], out, dict(residual=res)))
            return out

        def attention(blk, x: Act) -> Act:
            if blk.attention.use_causal_mask:
                raise NotImplementedError("tqdne_b200: the attention backward has no causal mask; train this model with "
                                          "the reference package")
            g = self._gn(blk.norm, [x], False)
            self.nodes.append(("gn", blk.norm, [x], g, False))
            qkv = self._conv(blk.qkv, [g], [g.C], stats=False)
            self.nodes.append(("conv", blk.qkv, [g], qkv, {}))
            # scratch of the backward kernels (row log-sum-exp, D_i): the forward kernel fills the first half when it can
            ws = torch.empty(2 * qkv.N * blk.num_heads * qkv.W, device=self.dev, dtype=torch.float32)
            self._keep.append(ws)
            a = self._plan(self.fwd).attention(qkv, blk.num_heads, lse=ws)
            self.nodes.append(("attn", blk, [qkv], a, ws))
            out = self._conv(blk.proj_out, [a], [a.C], residual=x, stats=True)
            self.nodes.append(("conv", blk.proj_out, [a], out, dict(residual=x)))
            return out

        def run_seq(seq, srcs: list[Act]) -> Act:
            cur = srcs
            for layer in seq:
                if isinstance(layer, U.ResBlock):
                    o = resblock(layer, cur)
                elif isinstance(layer, B.AttentionBlock):
                    o = attention(layer, cur[0])
                elif isinstance(layer, (B.Downsample, B.Upsample)) and not layer.use_conv:
                    raise NotImplementedError("tqdne_b200: the training tape covers conv_resample=True models only")
                elif isinstance(layer, B.Downsample):
                    o = self._conv(layer.op, [cur[0]], [cur[0].C], stride=2, stats=True)
                    self.nodes.append(("conv", layer.op, [cur[0]], o, dict(stride=2)))
                elif isinstance(layer, B.Upsample):
                    o = self._conv(layer.conv, [cur[0]], [cur[0].C], upsample=True, stats=True)
                    self.nodes.append(("conv", layer.conv, [cur[0]], o, dict(upsample=True)))
                elif isinstance(layer, nn.Conv1d):   # the stem
                    o = self._conv(layer, cur, [self.cin], stats=True)
                    self.nodes.append(("conv", layer, cur, o, dict(no_dgrad=True)))
                else:
                    raise NotImplementedError(f"tqdne_b200: cannot train through {type(layer).__name__}")
                cur = [o]
            return cur[0]

        hs, h = [], None
        first_tail = min(TAIL_FROM_INPUT_BLOCK, len(model.input_blocks) - 1)
        for i, blk in enumerate(model.input_blocks):
            if i == first_tail:
                self.tail_node0 = len(self.nodes)   # nodes [tail_node0, end) own the parameters of the late bucket
            h = run_seq(blk, [self.xin] if i == 0 else [h])
            hs.append(h)
        h = run_seq(model.middle_block, [h])
        for blk in model.output_blocks:
            h = run_seq(blk, [h, hs.pop()])
        gn_o, conv_o = model.out[0], model.out[2]
        g = self._gn(gn_o, [h], True)
        self.nodes.append(("gn", gn_o, [h], g, True))
        self.out = self._conv(conv_o, [g], [g.C], out_dtype=torch.float32, stats=False)
        self.cout_pad = _pad64(conv_o.weight.shape[0])
        self.dF = self._new(N, L, self.cout_pad)
        self.nodes.append(("conv", conv_o, [g], self.out, dict(dy=self.dF)))
        self._build_backward()

    # ------------------------------------------------------------------------------------------ backward builders
    def _build_backward(self):
        st, N, lib = self.store, self.N, self.lib
        grad: dict[int, Act] = {}
        ops = self.bwd
        E = 4 * self.model.model_channels
        R = st.emb_rows

        def rows_op(src: Act, dst: Act, mode: int, aux: Act | None = None):
            self._direct(ops, lambda: _lib.check(lib.tq_rows_op(src.t.data_ptr(), aux.t.data_ptr() if aux is not None else None,
                                                                dst.t.data_ptr(), mode, N, dst.W, dst.C, self._st()), "rows_op"))

        # Bias gradients ride on the GroupNorm backward where they can: dL/db of a convolution is the sum of its output's
        # gradient over samples and positions, and when the FIRST consumer (in forward order) of that output is a GroupNorm,
        # the norm's backward kernel -- the last contributor in backward order, with the other paths folded in through
        # dx_add -- holds exactly that gradient in registers (tq_gn_bwd_desc.dbias0/1).  Saves the separate reduction pass of
        # tq_conv1d_wgrad over dY for all but the attention qkv convolutions and the output head.
        first_use: dict[int, int] = {}
        producer_gb: dict[int, torch.Tensor] = {}
        for i, (kind, mod, srcs, out, extra) in enumerate(self.nodes):
            used = list(srcs)
            if kind == "conv":
                producer_gb[id(out)] = st.view(st.G, mod.bias)
                if extra.get("residual") is not None:
                    used.append(extra["residual"])
            for t in used:
                first_use.setdefault(id(t), i)
        bias_done: set[int] = set()

        self.bwd_split = 0
        for idx in range(len(self.nodes) - 1, -1, -1):
            if idx == self.tail_node0 - 1:
                # every gradient of the late bucket is final once the ops recorded so far have run
                self._cur = None
                self.bwd_split = len(ops)
            kind, mod, srcs, out, extra = self.nodes[idx]
            if kind == "conv":
                dy = extra.get("dy") or grad.get(id(out))
                assert dy is not None, "gradient of a convolution output is missing"
                res = extra.get("residual")
                if res is not None:
                    if id(res) in grad:                     # a skip tensor that an output block has already visited
                        both = self._new(N, dy.W, dy.C)
                        rows_op(grad[id(res)], both, 4, aux=dy)
                        grad[id(res)] = both
                    else:
                        grad[id(res)] = dy                  # the skip path's gradient IS dy; its other consumer adds it
                O, I, k = mod.weight.shape
                Op, Ip = _pad64(O), _pad64(I)
                gw, gb = st.view(st.G, mod.weight), st.view(st.G, mod.bias)
                dy_eff, x_eff = dy, list(srcs)
                if extra.get("stride") == 2:
                    dy_eff = self._new(N, srcs[0].W, dy.C)
                    rows_op(dy, dy_eff, 0)
                if extra.get("upsample"):
                    xu = self._new(N, 2 * srcs[0].W, srcs[0].C)
                    rows_op(srcs[0], xu, 1)
                    x_eff = [xu]
                # weight / bias gradient, one call per concat source
                coff = 0
                for si, xs in enumerate(x_eff):
                    want_gb = si == 0 and id(out) not in bias_done   # else: already summed by the consumer norm's backward
                    self._direct(ops, lambda xs=xs, dy_eff=dy_eff, gw=gw, gb=gb, want_gb=want_gb, coff=coff, k=k, Ip=Ip: _lib.check(
                        lib.tq_conv1d_wgrad(xs.t.data_ptr(), dy_eff.t.data_ptr(), gw.data_ptr(),
                                            gb.data_ptr() if want_gb else None,
                                            N, xs.W, xs.C, dy_eff.C, k, Ip, coff, self._st()), "conv1d_wgrad"))
                    coff += xs.C
                if extra.get("no_dgrad"):
                    continue
                # input gradient per source: forward igemm over dY with flipped, transposed weights
                coff = 0
                wv = st.view(st.P, mod.weight)                       # [Op, k, Ip]
                for si, xs in enumerate(srcs):
                    Cs = xs.C
                    wt = mod.weight.detach()[:, coff:coff + Cs, :].permute(1, 0, 2).flip(-1)   # [Cs, O, k]
                    pcb = pack_conv(wt, None, [O], BF16)
                    assert tuple(pcb.weights.shape) == (Cs, k * Op), (pcb.weights.shape, Cs, k, Op)
                    self.repack.append((wv, None, pcb.weights, coff, Cs))
                    self._keep.append(pcb)
                    pending = grad.get(id(xs)) if not extra.get("upsample") else None
                    dx = self._plan(ops).conv(pcb, [dy_eff], dims=1, residual=pending)
                    if extra.get("upsample"):
                        assert id(xs) not in grad
                        dxs = self._new(N, xs.W, Cs)
                        rows_op(dx, dxs, 2)
                        dx = dxs
                    grad[id(xs)] = dx
                    coff += Cs
            elif kind == "gn":
                dy = grad[id(out)]
                x0, x1 = srcs[0], (srcs[1] if len(srcs) > 1 else None)
                a0, a1 = grad.get(id(x0)), (grad.get(id(x1)) if x1 is not None else None)
                dx0 = self._new(N, x0.W, x0.C)
                dx1 = self._new(N, x1.W, x1.C) if x1 is not None else None
                Ct = x0.C + (x1.C if x1 is not None else 0)
                ws = torch.empty(N * Ct * 2, device=self.dev, dtype=torch.float32)
                d = _lib.TqGnBwdDesc()
                d.dtype = tq_dtype(BF16)
                d.N, d.P, d.C0, d.C1 = N, x0.W, x0.C, (x1.C if x1 is not None else 0)
                d.x0, d.x1, d.dy = x0.t.data_ptr(), (x1.t.data_ptr() if x1 is not None else None), dy.t.data_ptr()
                d.gamma, d.beta = st.view(st.P, mod.weight).data_ptr(), st.view(st.P, mod.bias).data_ptr()
                d.eps, d.silu = 1e-5, 1 if extra else 0
                assert x0.stats is not None and (x1 is None or x1.stats is not None), "forward statistics missing"
                d.stats0, d.stats1 = x0.stats.data_ptr(), (x1.stats.data_ptr() if x1 is not None else None)
                d.parts0, d.parts1 = x0.stats_parts, (x1.stats_parts if x1 is not None else 1)
                d.ws, d.dx0, d.dx1 = ws.data_ptr(), dx0.t.data_ptr(), (dx1.t.data_ptr() if dx1 is not None else None)
                d.dgamma, d.dbeta = st.view(st.G, mod.weight).data_ptr(), st.view(st.G, mod.bias).data_ptr()
                d.dx_add0 = a0.t.data_ptr() if a0 is not None else None
                d.dx_add1 = a1.t.data_ptr() if a1 is not None else None
                site = self.drop_of_act.get(id(out))
                if site is not None:
                    d.drop_seed, d.drop_p, d.drop_site = self.drop_seed.data_ptr(), self.p_drop, site
                for which, xs in ((0, x0), (1, x1)):
                    if xs is not None and first_use.get(id(xs)) == idx and id(xs) in producer_gb:
                        setattr(d, f"dbias{which}", producer_gb[id(xs)].data_ptr())
                        bias_done.add(id(xs))
                if id(x0) in self.emb_of_act:   # x0 = conv1(h0) + e: de[n][c] = sum over positions of d(x0), fused here
                    d.dx_sum = self.de_all[:, self.emb_of_act[id(x0)]:].data_ptr()
                    d.dx_sum_ld = R
                self._keep += [ws, d]
                self._direct(ops, lambda d=d: _lib.check(lib.tq_gn_silu_backward(C.byref(d), self._st()), "gn_silu_backward"))
                grad[id(x0)] = dx0
                if x1 is not None:
                    grad[id(x1)] = dx1
            elif kind == "attn":
                qkv, da = srcs[0], grad[id(out)]
                dqkv = self._new(N, qkv.W, qkv.C)
                heads = mod.num_heads
                ws, given = extra, 1 if getattr(out, "lse_written", False) else 0
                self._direct(ops, lambda qkv=qkv, out=out, da=da, dqkv=dqkv, ws=ws, heads=heads, given=given: _lib.check(
                    lib.tq_attention_backward(qkv.t.data_ptr(), out.t.data_ptr(), da.t.data_ptr(), dqkv.t.data_ptr(), ws.data_ptr(),
                                              N, qkv.W, heads, qkv.C // 3 // heads, given, self._st()), "attention_backward"))
                grad[id(qkv)] = dqkv
            elif kind == "mul":
                dy = grad[id(out)]
                dx = self._new(N, dy.W, dy.C)
                self._direct(ops, lambda dy=dy, dx=dx, site=extra: _lib.check(lib.tq_dropout_apply(
                    dy.t.data_ptr(), dx.t.data_ptr(), dy.t.numel(), self._host_seed(site), self.p_drop, self._st()), "dropout backward"))
                grad[id(srcs[0])] = dx
        # ---- embedding MLPs (all ResBlocks have deposited their de into de_all by now)
        f32 = dict(device=self.dev, dtype=torch.float32)
        gv = lambda p: st.view(st.G, p)  # noqa: E731
        pv = lambda p: st.view(st.P, p)  # noqa: E731
        self.demb, self.dt1 = torch.empty(N, E, **f32), torch.empty(N, E, **f32)
        gWall = st.G[st.emb_w_off:st.emb_w_off + R * E]
        gball = st.G[st.emb_b_off:st.emb_b_off + R]

        def lin_bwd(dy, x, W, act, dx, dW, db, M, K, Nout):
            self._direct(ops, lambda: _lib.check(lib.tq_linear_backward(
                dy.data_ptr(), x.data_ptr(), W.data_ptr(), act, dx.data_ptr() if dx is not None else None, dW.data_ptr(),
                db.data_ptr(), M, K, Nout, self._st()), "linear_backward"))

        model = self.model
        lin_bwd(self.de_all, self.emb, self.Wall, 1, self.demb, gWall, gball, N, E, R)
        t0, t2 = model.time_mlp[0], model.time_mlp[2]
        lin_bwd(self.demb, self.t1, pv(t2.weight), 1, self.dt1, gv(t2.weight), gv(t2.bias), N, E, E)
        lin_bwd(self.dt1, self.feat, pv(t0.weight), 0, None, gv(t0.weight), gv(t0.bias), N, self.feat.shape[1], E)
        if self.cond is not None:
            l0, l2 = model.cond_mlp[0], model.cond_mlp[2]
            self.dc1 = torch.empty(N, E, **f32)
            lin_bwd(self.demb, self.c1, pv(l2.weight), 1, self.dc1, gv(l2.weight), gv(l2.bias), N, E, E)
            lin_bwd(self.dc1, self.cond, pv(l0.weight), 0, None, gv(l0.weight), gv(l0.bias), N, self.cond.shape[1], E)

    # ------------------------------------------------------------------------------------------ running
    @torch.no_grad()
    def refresh_weights(self):
        """bf16 operand copies (forward, and tap-flipped / transposed for the input gradient) from the fp32 masters: ONE
        launch over a device-resident job table (`tq_repack_batch_run`; the ~170 per-convolution launches it replaces were
        a 0.68 ms serial chain even as a CUDA graph).  `TQ_REPACK_BATCH=0` keeps the per-convolution calls."""
        if os.environ.get("TQ_REPACK_BATCH", "1") == "0":
            st = self._st()
            for wv, fwd, bwd, coff, Cs in self.repack:
                Op, k, Ip = wv.shape
                _lib.check(self.lib.tq_repack_conv_weights(wv.data_ptr(), fwd.data_ptr() if fwd is not None else None,
                                                           bwd.data_ptr() if bwd is not None else None, Op, k, Ip, coff, Cs, st),
                           "repack_conv_weights")
        else:
            tab = self.__dict__.get("_repack_table")
            if tab is None:
                entries = []
                for wv, fwd, bwd, coff, Cs in self.repack:
                    Op, k, Ip = wv.shape
                    if fwd is not None:
                        entries.append((wv.data_ptr(), fwd.data_ptr(), None, Op, k, Ip, 0, 0))
                    if bwd is not None:
                        entries.append((wv.data_ptr(), None, bwd.data_ptr(), Op, k, Ip, coff, Cs))
                jobs = (_lib.TqRepackJob * len(entries))()
                for j, (m, f, b, Op, k, Ip, coff, Cs) in zip(jobs, entries):
                    j.master, j.fwd, j.bwd, j.Op, j.k, j.Ip, j.ci_off, j.Cs = m, f, b, Op, k, Ip, coff, Cs
                total = self.lib.tq_repack_batch_prepare(jobs, len(entries))
                if total <= 0:
                    raise RuntimeError("repack_batch_prepare: " + self.lib.tq_last_error().decode(errors="replace"))
                dev_jobs = torch.frombuffer(bytearray(bytes(jobs)), dtype=torch.uint8).to(self.dev)
                tab = self._repack_table = (dev_jobs, len(entries), int(total))
            dev_jobs, n, total = tab
            _lib.check(self.lib.tq_repack_batch_run(dev_jobs.data_ptr(), n, total, self._st()), "repack_batch_run")
        self.copies_version = self.store.version

    # training progress is shared by every tape of the model (see _Store)
    @property
    def step_count(self) -> int:
        return self.store.step_count

    @step_count.setter
    def step_count(self, v: int) -> None:
        self.store.step_count = v

    @property
    def pass_count(self) -> int:
        return self.store.pass_count

    @pass_count.setter
    def pass_count(self, v: int) -> None:
        self.store.pass_count = v

    def lr(self) -> float:
        """CosineAnnealingLR stepped every optimiser step (edm.py:242-251)."""
        t = min(self.step_count, self.max_steps)
        return self.eta_min + 0.5 * (self.lr0 - self.eta_min) * (1 + math.cos(math.pi * t / self.max_steps))

    @torch.no_grad()
    def forward_backward(self, signal: torch.Tensor, cond: torch.Tensor | None, *, sigma: torch.Tensor | None = None,
                         noise: torch.Tensor | None = None, world_size: int = 1) -> torch.Tensor:
        """Loss (0-dim CUDA tensor) and gradients (flat buffer `store.G`) of LightningEDM.step on `signal` [N, C, L].
        `sigma` / `noise` may be given explicitly (tests); otherwise they are drawn as in edm.py:125-127.
        `world_size` > 1 (data-parallel step): the all-reduce of the late gradient bucket is started as soon as the backward
        pass has produced it and runs beside the rest of the backward; optimizer_step(world_size) reduces the remainder and
        waits for both (TQ_TRAIN_OVERLAP=0: one all-reduce of the whole buffer in optimizer_step, as before)."""
        N, L = self.N, self.L
        assert tuple(signal.shape) == (N, self.cin, L), f"expected signal of shape {(N, self.cin, L)}"
        if self.copies_version != self.store.version:
            self.refresh_weights()   # another tape of this model stepped the optimiser since this one last ran
        self.y.copy_(signal.to(torch.float32).permute(0, 2, 1))
        if sigma is None:
            sigma = self.edm.edm.sigma(torch.randn(N, device=self.dev))
        self.sigma.copy_(sigma.to(torch.float32))
        if noise is None:
            self.noise.normal_()
        else:
            self.noise.copy_(noise.to(torch.float32).permute(0, 2, 1))
        self.t.copy_(0.25 * torch.log(self.sigma))          # EDM.noise_conditioning
        if self.cond is not None:
            assert cond is not None, "must specify cond if and only if the model is conditioned"
            self.cond.copy_(cond.to(torch.float32))
        self.pass_count += 1
        self.drop_seed.fill_((self.pass_count * 0x9E3779B1) & 0x7FFFFFFFFFFF)   # fresh dropout decisions every pass
        stale = self.__dict__.get("_early_work")
        if stale is not None:
            stale.wait()   # a pass whose optimizer_step never came: its all-reduce must not run into this pass's G.zero_()
        self._early_work = None
        between = None
        if (world_size > 1 and os.environ.get("TQ_TRAIN_OVERLAP", "1") != "0" and self.bwd_split > 0
                and 0 < self.store.tail_off < self.store.n):
            import torch.distributed as dist

            def between():
                # enqueued behind the first part of the tape (NCCL's stream waits for this stream's work so far), beside the second
                self._early_work = dist.all_reduce(self.store.G[self.store.tail_off:], async_op=True)
        self._run_tape(between)
        return self.loss[0]

    def _tape(self, part: int | None = None):
        """Everything between the step's inputs and its gradients: ~550 launches with static arguments.  `part` 0 = up to
        the point where the gradients of the late bucket (store.tail_off ..) are final, 1 = the rest of the backward."""
        N, L = self.N, self.L
        if part in (None, 0):
            self.store.G.zero_()
            self.de_all.zero_()
            st = self._st()
            for f in self.fwd:
                f()
            _lib.check(self.lib.tq_edm_loss(self.out.t.data_ptr(), self.out.C, self.xn.data_ptr(), self.y.data_ptr(),
                                            self.sigma.data_ptr(), self.dF.t.data_ptr(), self.loss.data_ptr(), N, L, self.cin,
                                            self.cout_pad, self.sigma_data, st), "edm_loss")
        lo = 0 if part in (None, 0) else self.bwd_split
        hi = len(self.bwd) if part in (None, 1) else self.bwd_split
        for b in self.bwd[lo:hi]:
            b()

    def _run_tape(self, between=None):
        """The tape as CUDA graphs (captured on the third pass, after two eager ones): the launches all have static
        arguments -- the dropout decisions come from a counter in device memory -- so a replay replaces ~550 host-side ctypes
        calls.  Two graphs, split where the late bucket's gradients are final: `between()` (the early gradient all-reduce of a
        data-parallel step) is called between their launches.  TQ_TRAIN_GRAPH=0, separate (host-seeded) dropout passes or a
        capture failure fall back to eager launches."""
        def eager():
            self._tape(0)
            if between is not None:
                between()
            self._tape(1)

        use = os.environ.get("TQ_TRAIN_GRAPH", "1") != "0" and (self.fused_dropout or self.p_drop == 0.0)
        if not use or self._graph is False:
            return eager()
        if self._graph is None:
            self._eager_passes = getattr(self, "_eager_passes", 0) + 1
            if self._eager_passes <= 2:
                return eager()
            try:
                graphs = []
                torch.cuda.synchronize(self.dev)
                for part in (0, 1):
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        self._tape(part)
                    graphs.append(g)
                self._graph = graphs      # capturing does not execute: the replays below run this pass
            except Exception as e:  # noqa: BLE001
                self._graph = False
                import warnings

                warnings.warn(f"tqdne_b200: CUDA-graph capture of the training tape failed ({e}); running eager")
                torch.cuda.synchronize(self.dev)
                return eager()
        self._graph[0].replay()
        if between is not None:
            between()
        self._graph[1].replay()

    @torch.no_grad()
    def optimizer_step(self, world_size: int = 1):
        """Gradient all-reduce (mean over ranks), Adam, EMA, refreshed operand copies."""
        s = self.store
        if world_size > 1:
            import torch.distributed as dist

            early = self.__dict__.get("_early_work")
            if early is not None:
                dist.all_reduce(s.G[: s.tail_off])   # the early bucket: what the end of the backward pass produced
                early.wait()
                self._early_work = None
            else:
                dist.all_reduce(s.G)
        lr = self.lr()            # the k-th update runs at the schedule's value after k - 1 scheduler steps (Lightning order)
        self.step_count += 1
        _lib.check(self.lib.tq_adam_ema_step(s.P.data_ptr(), s.G.data_ptr(), s.M.data_ptr(), s.V.data_ptr(), s.EMA.data_ptr(), s.n,
                                             lr, self.betas[0], self.betas[1], self.eps, self.step_count, self.ema_decay,
                                             1.0 / world_size, self._st()), "adam_ema_step")
        s.version += 1
        self.refresh_weights()

    def training_step(self, batch: dict, world_size: int = 1) -> torch.Tensor:
        loss = self.forward_backward(batch["signal"], batch.get("cond"), world_size=world_size)
        self.optimizer_step(world_size)
        return loss

    def grads_by_name(self) -> dict:
        """{parameter name: gradient in the module's (PyTorch) layout} -- for parity checks."""
        names = {id(p): n for n, p in self.model.named_parameters()}
        return {names[id(p)]: self.store.to_module_layout(self.store.G, p) for p in self.store.order}

    def sync_module(self, ema: bool = False):
        """Write the master (or EMA) parameters back into the nn.Module (state_dict export, sampling)."""
        self.store.store_to_module(self.store.EMA if ema else None)
