"""Stand-alone front-ends of the training-step kernels.

What autograd computes for the layers of the 1D EDM UNet (`LightningEDM.step`, tqdne/edm.py:115-134, config 5 of
BASELINE.json) expressed on the engine's kernels, channels-last bf16 activations / gradients, fp32 parameter gradients.
`tqdne_b200/training.py` strings the same kernels into the static tape of a whole step (with preallocated buffers); the
functions here allocate per call and exist for kernel-level parity tests and benchmarks:

  * `conv1d_input_grad`  -- dX of a stride-1 "same" convolution IS a convolution of dY with the tap-flipped,
    in/out-transposed weights: the forward tcgen05 implicit GEMM (`tq_plan_add_conv`) runs it unchanged, including the
    epilogue add that accumulates the gradient of a tensor with two consumers;
  * `conv1d_weight_grad` -- `tq_conv1d_wgrad` (tcgen05, K = positions, MN-major operands, taps by descriptor row offset);
  * `groupnorm_silu_backward` -- `tq_gn_silu_backward` (two streaming passes: group reductions, then dX / dgamma / dbeta);
  * `attention_backward` -- `tq_attention_backward` (two tcgen05 kernels: dQ + row statistics, transposed dK / dV);
  * `sample_channel_sums` -- the gradient of the per-sample embedding term.
"""

from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .engine import Act, Plan, current_stream_ptr, pack_conv, require_cuda


def conv1d_input_grad(plan: Plan, weight: torch.Tensor, dy: Act, *, accumulate_into: Act | None = None) -> Act:
    """Append dX = conv(dY, flip(W)^T) to `plan`.  weight: [cout, cin, k] (the forward layer's), dy: [N, 1, L, cout_pad].
    `accumulate_into`: a gradient already written by another consumer of X; it is added in the epilogue."""
    w = weight.detach()
    assert w.dim() == 3 and w.shape[2] % 2 == 1, "stride-1 'same' 1-D convolutions only"
    wt = w.permute(1, 0, 2).flip(-1).contiguous()           # [cin, cout, k]
    pc = pack_conv(wt, None, [w.shape[0]], plan.act_dtype)
    return plan.conv(pc, [dy], dims=1, residual=accumulate_into)


def conv1d_weight_grad(x: torch.Tensor, dy: torch.Tensor, taps: int, dw: torch.Tensor | None = None,
                       db: torch.Tensor | None = None, *, dw_ld: int = 0, ci_off: int = 0, bias: bool = True):
    """x: [N, L, cin] bf16, dy: [N, L, cout] bf16 (channels-last, channel counts multiples of 64) ->
    dw [cout, taps, cin] fp32 (+= if given), db [cout] fp32 (+= if given).  Reference layout: dw.permute(0, 2, 1)."""
    require_cuda(x, "x")
    N, L, cin = x.shape
    cout = dy.shape[2]
    assert dy.shape[:2] == (N, L) and x.dtype == dy.dtype == torch.bfloat16 and x.is_contiguous() and dy.is_contiguous()
    if dw is None:
        dw = torch.zeros(cout, taps, cin, device=x.device, dtype=torch.float32)
    if db is None and bias:
        db = torch.zeros(cout, device=x.device, dtype=torch.float32)
    _lib.check(_lib.lib().tq_conv1d_wgrad(x.data_ptr(), dy.data_ptr(), dw.data_ptr(), db.data_ptr() if bias else None, N, L, cin,
                                          cout, taps, dw_ld, ci_off, current_stream_ptr()), "conv1d_wgrad")
    return dw, db


def groupnorm_silu_backward(x0: Act, dy: Act, gamma: torch.Tensor, beta: torch.Tensor, *, silu: bool = True, x1: Act | None = None,
                            eps: float = 1e-5, dgamma: torch.Tensor | None = None, dbeta: torch.Tensor | None = None,
                            add0: Act | None = None, add1: Act | None = None, drop_seed: torch.Tensor | None = None,
                            drop_p: float = 0.0, drop_site: int = 0):
    """dX (one tensor per source), dgamma, dbeta of y = [SiLU](GroupNorm32(cat[x0, x1])).  x0 / x1 carry the forward
    per-(sample, channel) statistics (`Act.stats`, written by the producing conv's epilogue)."""
    from .engine import tq_dtype

    assert x0.stats is not None and (x1 is None or x1.stats is not None), "forward statistics missing"
    N, P = x0.N, x0.H * x0.W
    C0, C1 = x0.C, (x1.C if x1 is not None else 0)
    Ct = C0 + C1
    assert dy.C == Ct and dy.N == N and dy.H * dy.W == P
    dev = x0.t.device
    dx0 = torch.empty_like(x0.t)
    dx1 = torch.empty_like(x1.t) if x1 is not None else None
    dgamma = torch.zeros(Ct, device=dev, dtype=torch.float32) if dgamma is None else dgamma
    dbeta = torch.zeros(Ct, device=dev, dtype=torch.float32) if dbeta is None else dbeta
    ws = torch.empty(N * Ct * 2, device=dev, dtype=torch.float32)
    g32, b32 = gamma.detach().float().contiguous(), beta.detach().float().contiguous()
    d = _lib.TqGnBwdDesc()
    d.dtype = tq_dtype(x0.t.dtype)
    d.N, d.P, d.C0, d.C1 = N, P, C0, C1
    d.x0, d.x1, d.dy = x0.t.data_ptr(), (x1.t.data_ptr() if x1 is not None else None), dy.t.data_ptr()
    d.gamma, d.beta, d.eps, d.silu = g32.data_ptr(), b32.data_ptr(), eps, 1 if silu else 0
    d.stats0, d.stats1 = x0.stats.data_ptr(), (x1.stats.data_ptr() if x1 is not None else None)
    d.parts0, d.parts1 = x0.stats_parts, (x1.stats_parts if x1 is not None else 1)
    d.ws, d.dx0, d.dx1 = ws.data_ptr(), dx0.data_ptr(), (dx1.data_ptr() if dx1 is not None else None)
    d.dgamma, d.dbeta = dgamma.data_ptr(), dbeta.data_ptr()
    d.dx_sum, d.dx_sum_ld = None, 0
    if drop_seed is not None and drop_p > 0:
        d.drop_seed, d.drop_p, d.drop_site = drop_seed.data_ptr(), drop_p, drop_site
    d.dx_add0 = add0.t.data_ptr() if add0 is not None else None
    d.dx_add1 = add1.t.data_ptr() if add1 is not None else None
    _lib.check(_lib.lib().tq_gn_silu_backward(C.byref(d), current_stream_ptr()), "gn_silu_backward")
    out0 = Act(dx0, N, x0.H, x0.W, C0)
    out1 = Act(dx1, N, x1.H, x1.W, C1) if x1 is not None else None
    return out0, out1, dgamma, dbeta


def sample_channel_sums(dy: Act, out: torch.Tensor | None = None, out_ld: int = 0) -> torch.Tensor:
    """[N, C] fp32 sums of dy over the positions of each sample (+= into `out`): gradient of the embedding term."""
    if out is None:
        out = torch.zeros(dy.N, dy.C, device=dy.t.device, dtype=torch.float32)
    assert dy.t.dtype == torch.bfloat16
    _lib.check(_lib.lib().tq_sample_channel_sums(dy.t.data_ptr(), out.data_ptr(), out_ld, dy.N, dy.H * dy.W, dy.C, current_stream_ptr()),
               "sample_channel_sums")
    return out


def attention_backward(qkv: Act, out: Act, dout: Act, heads: int, lse: torch.Tensor | None = None) -> Act:
    """d(qkv) [N, T, 3C] of the attention core (QKVAttention.forward, blocks.py:156-190) from the forward's input / output
    and the gradient of its output; bf16, head dim 64, 32 < T <= 512.  `lse`: the 2 * N * heads * T float scratch whose first
    half the FORWARD kernel filled with the rows' log-sum-exp (Plan.attention(lse=...) with out.lse_written): the backward
    then skips recomputing it."""
    N, T, C3 = qkv.N, qkv.H * qkv.W, qkv.C
    Cc = C3 // 3
    assert out.C == Cc and dout.C == Cc and qkv.t.dtype == torch.bfloat16
    dqkv = torch.empty_like(qkv.t)
    given = lse is not None
    ws = lse if given else torch.empty(2 * N * heads * T, device=qkv.t.device, dtype=torch.float32)
    assert ws.numel() >= 2 * N * heads * T and ws.dtype == torch.float32
    _lib.check(_lib.lib().tq_attention_backward(qkv.t.data_ptr(), out.t.data_ptr(), dout.t.data_ptr(), dqkv.data_ptr(),
                                                ws.data_ptr(), N, T, heads, Cc // heads, 1 if given else 0, current_stream_ptr()),
               "attention_backward")
    return Act(dqkv, N, qkv.H, qkv.W, C3)


def conv2d_weight_grad(x: torch.Tensor, dy: torch.Tensor, kh: int, kw: int, dw: torch.Tensor | None = None) -> torch.Tensor:
    """x: [N, H, W, cin] bf16, dy: [N, H, W, cout] bf16 (channels-last, channel counts multiples of 64) ->
    dw [cout, kh*kw, cin] fp32 (+= if given).  Reference layout: dw.view(cout, kh, kw, cin).permute(0, 3, 1, 2)."""
    require_cuda(x, "x")
    N, H, W, cin = x.shape
    cout = dy.shape[3]
    assert dy.shape[:3] == (N, H, W) and x.dtype == dy.dtype == torch.bfloat16 and x.is_contiguous() and dy.is_contiguous()
    if dw is None:
        dw = torch.zeros(cout, kh * kw, cin, device=x.device, dtype=torch.float32)
    _lib.check(_lib.lib().tq_conv2d_wgrad(x.data_ptr(), dy.data_ptr(), dw.data_ptr(), N, H, W, cin, cout, kh, kw, current_stream_ptr()),
               "conv2d_wgrad")
    return dw
