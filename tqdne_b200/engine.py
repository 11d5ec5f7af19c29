"""Host-side lowering of the tqdne networks onto the C-ABI kernels.

PyTorch is used here for device memory and streams only.  A network (UNetModel / Decoder / Encoder) is
lowered ONCE per (batch, spatial size, precision) into a `tq_plan` -- a straight-line list of kernel
launches with pre-encoded TMA tensor maps -- and replayed (optionally as a CUDA graph) for every
denoiser call.  Activations are channels-last ([N, H, W, C]; 1D signals use H = 1).

Lowering rules (reference call sites in brackets):
  * conv_nd(k, padding="same")            -> K-slices, one per (tap, 64-channel chunk)   [nn.py:16-24]
  * th.cat([h, skip]) -> GroupNorm -> conv -> two-source GroupNorm, concat never stored   [unet.py:396]
  * ResBlock 1x1 skip_connection           -> extra K-slices of the second conv           [unet.py:108-112,143]
  * identity skip / attention residual     -> conv epilogue residual add                  [unet.py:143, blocks.py:145]
  * emb_layers (22 x SiLU -> Linear)       -> one GEMM over concatenated weights, added per
                                              sample in the first conv's epilogue         [unet.py:91-99,141]
  * Downsample (k=3, s=2, p=1)             -> four parity views of the input              [blocks.py:93-101]
  * Upsample (nearest x2 -> conv)          -> four output-parity classes on the low-res
                                              grid; the 4x tensor is never stored         [blocks.py:59-65]
"""

from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field

import torch

from . import _lib
from ._lib import TQ_BF16, TQ_F32, TQ_F64

GN_EPS = 1e-5  # nn.GroupNorm default, reference nn.py:105


def _pad64(c: int) -> int:
    return (c + 63) // 64 * 64


def tq_dtype(t: torch.dtype) -> int:
    return {torch.bfloat16: TQ_BF16, torch.float32: TQ_F32, torch.float64: TQ_F64}[t]


def require_cuda(t: torch.Tensor, name: str = "tensor") -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f"tqdne_b200: {name} must live on a CUDA device (got {t.device}); this engine has no CPU path."
        )


def current_stream_ptr() -> int:
    """Raw handle of the CURRENT device's current stream: every public entry point runs under `device_guard`, so the
    current device is the one that owns the tensors / plans being used."""
    return torch.cuda.current_stream().cuda_stream


def device_guard(device):
    """Make `device` (of the module's parameters or of the input tensor) the current CUDA device for the duration of a
    call: streams, CUDA-graph capture, the C-ABI launches and the SM count used by the plan builders all follow the
    current device, not the device of the pointers they are handed."""
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError(f"tqdne_b200: expected a CUDA device, got {device}; this engine has no CPU path.")
    return torch.cuda.device(device)


# ------------------------------------------------------------------------------------------------
# packed convolution weights
# ------------------------------------------------------------------------------------------------
@dataclass
class PackedConv:
    """Weight matrix [cout_pad, ktot] (K contiguous) + the K-block index of every (tap, segment, chunk)."""

    weights: torch.Tensor
    bias: torch.Tensor
    cout: int
    cout_pad: int
    ktot: int
    taps: list            # list of spatial offsets (ky, kx) in kernel coordinates
    seg_pad: list         # padded channel count per input segment
    kb_of: dict           # (tap_index, seg_index, chunk) -> K block
    kernel: tuple
    extra_kb: dict = field(default_factory=dict)  # fused 1x1 shortcut: (seg_index, chunk) -> K block
    sc_seg_pad: list = field(default_factory=list)  # padded channel count per shortcut segment
    macs_per_out: int = 0  # real (unpadded) multiply-adds per output element: taps*Cin (+ shortcut Cin)


def pack_conv(weight: torch.Tensor, bias: torch.Tensor | None, segments: list[int], dtype: torch.dtype,
              shortcut: tuple | None = None) -> PackedConv:
    """Repack a conv / linear weight [O, I, *k] for the implicit GEMM.

    `segments` splits the I input channels over the sources that feed the conv (skip concat);
    every segment is zero-padded to a multiple of 64 channels.  `shortcut` = (weight_1x1, bias,
    shortcut_segments) appends the ResBlock's 1x1 skip_connection -- which reads the block INPUT, split
    over `shortcut_segments` -- as extra K columns, and folds its bias in.
    """
    w = weight.detach().to(torch.float32)
    O, I = w.shape[0], w.shape[1]
    assert sum(segments) == I, (segments, I)
    ks = tuple(w.shape[2:])
    if len(ks) == 0:
        taps = [(0, 0)]
        w = w.reshape(O, I, 1, 1)
        kernel = (1, 1)
    elif len(ks) == 1:
        taps = [(0, kx) for kx in range(ks[0])]
        w = w.reshape(O, I, 1, ks[0])
        kernel = (1, ks[0])
    else:
        taps = [(ky, kx) for ky in range(ks[0]) for kx in range(ks[1])]
        kernel = ks
    seg_pad = [_pad64(s) for s in segments]
    cin_pad = sum(seg_pad)
    cout_pad = _pad64(O)
    sc_seg_pad = [_pad64(s) for s in shortcut[2]] if shortcut is not None else []
    ktot = len(taps) * cin_pad + sum(sc_seg_pad)
    B = torch.zeros(cout_pad, ktot, dtype=torch.float32, device=w.device)
    kb_of = {}
    col = 0
    for ti, (ky, kx) in enumerate(taps):
        off = 0
        for si, (s, sp) in enumerate(zip(segments, seg_pad)):
            B[:O, col:col + s] = w[:, off:off + s, ky, kx]
            for ch in range(sp // 64):
                kb_of[(ti, si, ch)] = (col // 64) + ch
            col += sp
            off += s
    extra_kb = {}
    b = torch.zeros(cout_pad, dtype=torch.float32, device=w.device)
    if bias is not None:
        b[:O] = bias.detach().to(torch.float32)
    if shortcut is not None:
        ws, bs, sc_segments = shortcut
        ws = ws.detach().to(torch.float32).reshape(O, sum(sc_segments))
        off = 0
        for si, (s, sp) in enumerate(zip(sc_segments, sc_seg_pad)):
            B[:O, col:col + s] = ws[:, off:off + s]
            for ch in range(sp // 64):
                extra_kb[(si, ch)] = (col // 64) + ch
            col += sp
            off += s
        if bs is not None:
            b[:O] += bs.detach().to(torch.float32)
    assert col == ktot
    macs = len(taps) * I + (sum(shortcut[2]) if shortcut is not None else 0)
    return PackedConv(B.to(dtype).contiguous(), b.contiguous(), O, cout_pad, ktot, taps, seg_pad, kb_of, kernel, extra_kb,
                      sc_seg_pad, macs)


# ------------------------------------------------------------------------------------------------
# activations
# ------------------------------------------------------------------------------------------------
@dataclass
class Act:
    """A channels-last activation [N, H, W, C] (dense)."""

    t: torch.Tensor
    N: int
    H: int
    W: int
    C: int
    stats: torch.Tensor | None = None   # [N, parts, C, 2] fp32 per-(sample, tile, channel) sum / sum of squares (conv epilogue)
    stats_parts: int = 1
    lse_written: bool = False           # attention output: the kernel also left the rows' log-sum-exp in the caller's buffer

    @property
    def P(self) -> int:
        return self.H * self.W


class Pool:
    """Size-keyed free list so that dead activations are reused (keeps the working set small)."""

    def __init__(self, device, dtype):
        self.device, self.dtype = device, dtype
        self.free: dict[int, list[torch.Tensor]] = {}
        self.all: list[torch.Tensor] = []

    def get(self, numel: int, dtype=None) -> torch.Tensor:
        dtype = dtype or self.dtype
        key = (numel, dtype)
        lst = self.free.get(key)
        if lst:
            return lst.pop()
        t = torch.empty(numel, device=self.device, dtype=dtype)
        self.all.append(t)
        return t

    def put(self, t: torch.Tensor) -> None:
        self.free.setdefault((t.numel(), t.dtype), []).append(t)

    def bytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self.all)


class Plan:
    """Python owner of a tq_plan plus every tensor its kernels point at."""

    def __init__(self, device, act_dtype: torch.dtype):
        self.lib = _lib.lib()
        self.h = self.lib.tq_plan_create()
        if not self.h:
            raise RuntimeError("tq_plan_create failed")
        self.device = device
        self.act_dtype = act_dtype
        self.tq_dtype = tq_dtype(act_dtype)
        self.pool = Pool(device, act_dtype)
        self.keep: list = []
        self._stats_block = None   # bump-allocated arena of conv-epilogue GroupNorm statistics
        self._stats_used = 0
        # one entry per kernel launch of the plan: (kind, algorithmic FLOPs, algorithmic HBM bytes)
        self.op_meta: list[tuple] = []

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.tq_plan_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # -- buffers --------------------------------------------------------------------------------
    def new_act(self, N, H, W, Cc, dtype=None) -> Act:
        t = self.pool.get(N * H * W * Cc, dtype)
        return Act(t, N, H, W, Cc)

    def release(self, a: Act) -> None:
        self.pool.put(a.t)

    STATS_BLOCK_FLOATS = 4 << 20  # 16 MiB per arena block

    def _stats_alloc(self, nfloats: int) -> torch.Tensor:
        """Statistics buffers are unique per producing conv (never pooled) and live in a few zero-filled arena blocks.
        Every slot a conv geometry writes is overwritten (plain stores) on every run and the slots it never writes stay
        zero, so nothing has to be cleared between runs."""
        nfloats = (nfloats + 3) // 4 * 4
        if self._stats_block is None or self._stats_used + nfloats > self._stats_block.numel():
            size = max(self.STATS_BLOCK_FLOATS, nfloats)
            self._stats_block = torch.zeros(size, device=self.device, dtype=torch.float32)
            self._stats_used = 0
            self.keep.append(self._stats_block)
        v = self._stats_block[self._stats_used:self._stats_used + nfloats]
        self._stats_used += nfloats
        return v

    # -- ops ------------------------------------------------------------------------------------
    def conv(self, pc: PackedConv, srcs: list[Act], *, out: Act | None = None, out_dtype=None, stride: int = 1,
             upsample: bool = False, emb: torch.Tensor | None = None, emb_ld: int = 0, residual: Act | None = None,
             shortcut_srcs: list[Act] | None = None, block_n: int = 0, dims: int = 2, stats: bool = False,
             cta_group: int = 0) -> Act:
        """Append one implicit-GEMM conv.  `srcs` are the concat segments of the conv input.  `stats`: also
        accumulate the per-(sample, channel) GroupNorm sums of the output in the epilogue."""
        N, H, W = srcs[0].N, srcs[0].H, srcs[0].W
        assert len(srcs) == len(pc.seg_pad)
        kh, kw = pc.kernel
        pad_y, pad_x = kh // 2, kw // 2
        d = _lib.TqConvDesc()
        d.dtype = self.tq_dtype
        d.cout, d.cout_pad, d.ktot = pc.cout, pc.cout_pad, pc.ktot
        slices: list[tuple] = []  # (src, dx, dy, c0, kb)
        src_views: list[tuple] = []  # (tensor_ptr_offset_elems, tensor, N, H, W, C, sn, sy, sx)

        def add_src(a: Act, off=0, Hh=None, Ww=None, sy=None, sx=None):
            Cc = a.C
            src_views.append((a.t, off, a.N, Hh if Hh is not None else a.H, Ww if Ww is not None else a.W, Cc,
                              a.H * a.W * Cc, sy if sy is not None else a.W * Cc, sx if sx is not None else Cc))
            return len(src_views) - 1

        classes = 1
        class_off = [0, 0, 0, 0]
        if stride == 1 and not upsample:
            ids = [add_src(a) for a in srcs]
            for ti, (ky, kx) in enumerate(pc.taps):
                for si, a in enumerate(srcs):
                    for ch in range(pc.seg_pad[si] // 64):
                        slices.append((ids[si], kx - pad_x, ky - pad_y, ch * 64, pc.kb_of[(ti, si, ch)]))
            Ho, Wo = H, W
            gH, gW = H, W
            o_sy, o_sx = None, None
        elif stride == 2:
            # Downsample (blocks.py:93-101): out(y,x) = sum W[ky,kx] in(2y+ky-1, 2x+kx-1): parity views
            assert len(srcs) == 1 and not upsample
            a = srcs[0]
            assert W % 2 == 0 and (H % 2 == 0 or H == 1), "stride-2 conv needs even extents"
            Ho, Wo = (H // 2 if H > 1 else 1), W // 2
            views = {}
            for ti, (ky, kx) in enumerate(pc.taps):
                s_y, s_x = ky - pad_y, kx - pad_x
                py, px = (s_y % 2 if H > 1 else 0), s_x % 2
                dy, dx = ((s_y - py) // 2 if H > 1 else 0), (s_x - px) // 2
                if (py, px) not in views:
                    views[(py, px)] = add_src(a, off=(py * W + px) * a.C, Hh=Ho, Ww=Wo,
                                              sy=2 * W * a.C if H > 1 else W * a.C, sx=2 * a.C)
                for ch in range(pc.seg_pad[0] // 64):
                    slices.append((views[(py, px)], dx, dy, ch * 64, pc.kb_of[(ti, 0, ch)]))
            gH, gW = Ho, Wo
            o_sy, o_sx = None, None
        else:
            # Upsample (blocks.py:59-65): nearest x2 then conv; output parity classes on the low-res grid
            assert upsample and len(srcs) == 1
            a = srcs[0]
            ids = [add_src(a)]
            two_d = H > 1
            Ho, Wo = (2 * H if two_d else 1), 2 * W
            cls_list = [(py, px) for py in ((0, 1) if two_d else (0,)) for px in (0, 1)]
            classes = len(cls_list)
            per_class = []
            for ci, (py, px) in enumerate(cls_list):
                sl = []
                for ti, (ky, kx) in enumerate(pc.taps):
                    dy = (py + ky - pad_y) // 2 if two_d else 0
                    dx = (px + kx - pad_x) // 2
                    for ch in range(pc.seg_pad[0] // 64):
                        sl.append((0, dx, dy, ch * 64, pc.kb_of[(ti, 0, ch)]))
                per_class.append(sl)
                class_off[ci] = (py * Wo + px) * pc.cout
            slices = [s for sl in per_class for s in sl]
            gH, gW = H, W
            o_sy, o_sx = 2 * Wo * pc.cout if two_d else Wo * pc.cout, 2 * pc.cout
        if shortcut_srcs is not None:
            assert stride == 1 and not upsample and pc.extra_kb
            base = len(src_views)
            for a in shortcut_srcs:
                add_src(a)
            assert len(shortcut_srcs) == len(pc.sc_seg_pad)
            for si, a in enumerate(shortcut_srcs):
                for ch in range(pc.sc_seg_pad[si] // 64):
                    slices.append((base + si, 0, 0, ch * 64, pc.extra_kb[(si, ch)]))
        assert len(src_views) <= 4, "at most 4 conv sources"

        odt = out_dtype or self.act_dtype
        if out is None:
            out = self.new_act(N, Ho, Wo, pc.cout, odt)
        assert out.C == pc.cout and out.N == N and out.H == Ho and out.W == Wo
        n_per_class = len(slices) // classes
        arr = (_lib.TqSlice * len(slices))()
        for i, (s, dx, dy, c0, kb) in enumerate(slices):
            arr[i].src, arr[i].dx, arr[i].dy, arr[i].c0, arr[i].kb = s, dx, dy, c0, kb
        d.N, d.H, d.W = N, gH, gW
        d.num_srcs, d.num_classes, d.num_slices = len(src_views), classes, n_per_class
        esz = srcs[0].t.element_size()
        for i, (t, off, n_, h_, w_, c_, sn, sy, sx) in enumerate(src_views):
            d.srcs[i].ptr = t.data_ptr() + off * esz
            assert c_ % 64 == 0, "conv sources must carry a multiple of 64 channels (pad the stem input)"
            d.srcs[i].N, d.srcs[i].H, d.srcs[i].W, d.srcs[i].C = n_, h_, w_, c_
            d.srcs[i].sn, d.srcs[i].sy, d.srcs[i].sx = sn, sy, sx
        d.slices = C.cast(arr, C.POINTER(_lib.TqSlice))
        d.weights = pc.weights.data_ptr()
        d.bias = pc.bias.data_ptr()
        d.emb = emb.data_ptr() if emb is not None else None
        d.emb_ld = emb_ld
        d.residual = residual.t.data_ptr() if residual is not None else None
        if residual is not None:
            assert residual.C == pc.cout and residual.t.dtype == self.act_dtype and classes == 1
        d.out = out.t.data_ptr()
        d.out_dtype = tq_dtype(odt)
        d.out_sn = Ho * Wo * pc.cout
        d.out_sy = o_sy if o_sy is not None else Wo * pc.cout
        d.out_sx = o_sx if o_sx is not None else pc.cout
        for i in range(4):
            d.out_class_off[i] = class_off[i]
        d.block_n = block_n
        d.cta_group = cta_group
        tensor_path = self.act_dtype == torch.bfloat16
        if stats and ((tensor_path and odt == torch.bfloat16 and pc.cout % 64 == 0) or (not tensor_path and pc.cout % 32 == 0)):
            parts = int(self.lib.tq_conv_stats_parts(C.byref(d)))
            if parts < 1:
                raise RuntimeError("tq_conv_stats_parts failed")
            out.stats = self._stats_alloc(N * parts * pc.cout * 2)
            out.stats_parts = parts
            d.stats = out.stats.data_ptr()
            d.stats_parts = parts
        _lib.check(self.lib.tq_plan_add_conv(self.h, C.byref(d)), "plan_add_conv")
        # algorithmic bytes: every source once, weights once, output once (+ residual once)
        io_bytes = sum(n_ * h_ * w_ * c_ for (_, _, n_, h_, w_, c_, _, _, _) in src_views) * esz
        io_bytes += pc.weights.numel() * pc.weights.element_size() + out.t.numel() * out.t.element_size()
        if residual is not None:
            io_bytes += residual.t.numel() * residual.t.element_size()
        self.op_meta.append(("conv", 2 * N * Ho * Wo * pc.cout * pc.macs_per_out, io_bytes))
        self.keep += [pc.weights, pc.bias, emb, out.t] + [v[0] for v in src_views]
        if residual is not None:
            self.keep.append(residual.t)
        return out

    def groupnorm(self, srcs: list[Act], gamma: torch.Tensor, beta: torch.Tensor, silu: bool, *, drop_seed: torch.Tensor | None = None,
                  drop_p: float = 0.0, drop_site: int = 0, film: torch.Tensor | None = None, film_ld: int = 0) -> Act:
        """`film`: fp32 view whose row n holds [scale | shift] of sample n (2 * C values, row stride `film_ld`):
        y = act(GroupNorm(x) * (1 + scale) + shift), ResBlock(use_scale_shift_norm=True)."""
        a0 = srcs[0]
        a1 = srcs[1] if len(srcs) > 1 else None
        Ct = a0.C + (a1.C if a1 else 0)
        out = self.new_act(a0.N, a0.H, a0.W, Ct)
        d = _lib.TqGnDesc()
        d.dtype = self.tq_dtype
        d.N, d.P, d.C0, d.C1 = a0.N, a0.P, a0.C, (a1.C if a1 else 0)
        d.x0 = a0.t.data_ptr()
        d.x1 = a1.t.data_ptr() if a1 else None
        g = gamma.detach().to(torch.float32).contiguous()
        b = beta.detach().to(torch.float32).contiguous()
        d.gamma, d.beta = g.data_ptr(), b.data_ptr()
        d.eps = GN_EPS
        d.silu = 1 if silu else 0
        d.y = out.t.data_ptr()
        if film is not None:
            assert film.dtype == torch.float32 and film_ld >= 2 * Ct
            d.film, d.film_ld = film.data_ptr(), film_ld
            self.keep.append(film)
        if drop_seed is not None and drop_p > 0:      # training only: dropout fused behind the activation
            d.drop_seed, d.drop_p, d.drop_site = drop_seed.data_ptr(), drop_p, drop_site
            self.keep.append(drop_seed)
        fused = a0.stats is not None and (a1 is None or a1.stats is not None)
        if fused:
            d.stats0 = a0.stats.data_ptr()
            d.stats1 = a1.stats.data_ptr() if a1 else None
            d.parts0, d.parts1 = a0.stats_parts, (a1.stats_parts if a1 else 1)
        need = int(self.lib.tq_groupnorm_ws_floats(C.byref(d)))
        if need < 0:
            raise RuntimeError("tq_groupnorm_ws_floats failed")
        if need > 0:
            # per-op scratch (the stand-alone statistics pass / the per-sample group reduction of many-part tensors)
            ws = torch.zeros(need, device=self.device, dtype=torch.float32)
            self.keep.append(ws)
            d.ws = ws.data_ptr()
        before = self.num_ops
        _lib.check(self.lib.tq_plan_add_groupnorm(self.h, C.byref(d)), "plan_add_groupnorm")
        nops = self.num_ops - before   # [gn_stats] [gn_finalize] gn_apply
        esz = a0.t.element_size()
        nel = a0.N * a0.P * Ct
        if not fused:
            self.op_meta.append(("gn_stats", 0, nel * esz))      # one read
        if nops == (2 if fused else 3):
            self.op_meta.append(("gn_finalize", 0, 0))
        self.op_meta.append(("gn_apply", 0, 2 * nel * esz))      # one read + one write
        self.keep += [g, b, a0.t, out.t] + ([a1.t] if a1 else [])
        return out

    def attention(self, qkv: Act, heads: int, causal: bool = False, lse: torch.Tensor | None = None) -> Act:
        """`lse` (training): fp32 [N * heads * T] the kernel fills with the rows' log-sum-exp for the backward pass, when the
        kernel chosen for this shape can (`out.lse_written`)."""
        Cc = qkv.C // 3
        out = self.new_act(qkv.N, qkv.H, qkv.W, Cc)
        d = _lib.TqAttnDesc()
        d.dtype = self.tq_dtype
        d.N, d.T, d.heads, d.d = qkv.N, qkv.P, heads, Cc // heads
        d.qkv, d.out = qkv.t.data_ptr(), out.t.data_ptr()
        d.causal = 1 if causal else 0
        if lse is not None and int(self.lib.tq_attention_writes_lse(C.byref(d))) == 1:
            assert lse.dtype == torch.float32 and lse.numel() >= qkv.N * heads * qkv.P
            d.lse = lse.data_ptr()
            out.lse_written = True
            self.keep.append(lse)
        _lib.check(self.lib.tq_plan_add_attention(self.h, C.byref(d)), "plan_add_attention")
        self.op_meta.append(("attention", 4 * qkv.N * qkv.P * qkv.P * Cc, 0))
        self.keep += [qkv.t, out.t]
        return out

    def linear(self, x, W, b, M, *, x_rows=None, act_in=False, add=None, add_rows=0, y=None, y_act=None):
        d = _lib.TqLinearDesc()
        d.M, d.K, d.Nout = M, W.shape[1], W.shape[0]
        d.x_rows = x_rows if x_rows is not None else M
        d.x, d.W = x.data_ptr(), W.data_ptr()
        d.b = b.data_ptr() if b is not None else None
        d.act_in = 1 if act_in else 0
        d.add = add.data_ptr() if add is not None else None
        d.add_rows = add_rows
        d.y = y.data_ptr() if y is not None else None
        d.y_act = y_act.data_ptr() if y_act is not None else None
        d.y_act_dtype = tq_dtype(y_act.dtype) if y_act is not None else TQ_F32
        _lib.check(self.lib.tq_plan_add_linear(self.h, C.byref(d)), "plan_add_linear")
        self.op_meta.append(("linear", 2 * M * W.shape[0] * W.shape[1], 0))
        self.keep += [x, W, b, add, y, y_act]

    def fourier(self, t, Wf, M, feat):
        _lib.check(self.lib.tq_plan_add_fourier(self.h, t.data_ptr(), Wf.data_ptr(), M, Wf.numel(), feat.data_ptr()),
                   "plan_add_fourier")
        self.op_meta.append(("fourier", 0, 0))
        self.keep += [t, Wf, feat]

    def resample2(self, x: Act, mode: str) -> Act:
        """'avg_pool' (nn.AvgPool{1,2}d(2, 2)) or 'nearest' (F.interpolate x2): the conv-less resamplers of conv_resample=False.
        The result carries no GroupNorm statistics: the consuming norm runs its own statistics pass."""
        up = mode == "nearest"
        Ho = 1 if x.H == 1 else (2 * x.H if up else x.H // 2)
        Wo = 2 * x.W if up else x.W // 2
        out = self.new_act(x.N, Ho, Wo, x.C)
        _lib.check(self.lib.tq_plan_add_resample2(self.h, self.tq_dtype, x.t.data_ptr(), out.t.data_ptr(), x.N, x.H, x.W, x.C,
                                                  1 if up else 0), "plan_add_resample2")
        esz = x.t.element_size()
        self.op_meta.append(("resample2", 0, (x.N * x.P + out.N * out.P) * x.C * esz))
        self.keep += [x.t, out.t]
        return out

    def spatial_mean(self, x: torch.Tensor, N: int, P: int, Cc: int, ld: int, y: torch.Tensor):
        _lib.check(self.lib.tq_plan_add_spatial_mean(self.h, x.data_ptr(), N, P, Cc, ld, y.data_ptr()), "plan_add_spatial_mean")
        self.op_meta.append(("spatial_mean", 0, 4 * N * P * Cc))
        self.keep += [x, y]

    # -- execution ------------------------------------------------------------------------------
    def enable_graph(self, on: bool = True) -> None:
        _lib.check(self.lib.tq_plan_enable_graph(self.h, 1 if on else 0), "plan_enable_graph")

    def run(self) -> None:
        _lib.check(self.lib.tq_plan_run(self.h, current_stream_ptr()), "plan_run")

    def run_range(self, first: int, last: int) -> None:
        _lib.check(self.lib.tq_plan_run_range(self.h, first, last, current_stream_ptr()), "plan_run_range")

    @property
    def num_ops(self) -> int:
        return self.lib.tq_plan_num_ops(self.h)

    def op_names(self) -> list[str]:
        return [self.lib.tq_plan_op_name(self.h, i).decode() for i in range(self.num_ops)]


# ------------------------------------------------------------------------------------------------
# layout helpers (C-ABI kernels, not torch permutes)
# ------------------------------------------------------------------------------------------------
def nchw_to_nhwc(src: torch.Tensor, dst_dtype: torch.dtype, cpad: int | None = None) -> torch.Tensor:
    """[N, C, *spatial] -> channels-last [N, P, cpad] (zero padded channels)."""
    require_cuda(src, "input")
    src = src.contiguous()
    N, Cc = src.shape[0], src.shape[1]
    P = src[0, 0].numel()
    cpad = cpad or Cc
    dst = torch.empty(N, P, cpad, device=src.device, dtype=dst_dtype)
    _lib.check(_lib.lib().tq_nchw_to_nhwc(src.data_ptr(), tq_dtype(src.dtype), dst.data_ptr(), tq_dtype(dst_dtype), N, Cc,
                                          P, cpad, current_stream_ptr()), "nchw_to_nhwc")
    return dst


def nhwc_to_nchw(src: torch.Tensor, N: int, Cc: int, spatial: tuple, cld: int, dst_dtype: torch.dtype) -> torch.Tensor:
    """channels-last [N, P, cld] (first Cc channels) -> [N, Cc, *spatial]."""
    P = math.prod(spatial)
    dst = torch.empty(N, Cc, *spatial, device=src.device, dtype=dst_dtype)
    _lib.check(_lib.lib().tq_nhwc_to_nchw(src.data_ptr(), tq_dtype(src.dtype), cld, dst.data_ptr(), tq_dtype(dst_dtype), N,
                                          Cc, P, current_stream_ptr()), "nhwc_to_nchw")
    return dst
