"""GPU parity of every kernel behind the C-ABI against plain PyTorch fp32 ops on the same inputs.

Tolerances (rel-L2 against an fp32 torch reference fed the same, already-rounded operands):
  bf16 tensor path : 5e-3 with bf16 output (one bf16 rounding of the result), 2e-5 with fp32 output
  fp32 FFMA path   : 1e-5 (K up to 2304 products summed in a different order than cuDNN)
"""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests.conftest import rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def _plan(dtype):
    from tqdne_b200.engine import Plan

    return Plan(torch.device("cuda"), dtype)


def _act(x_nchw, dtype, cpad=None):
    """[N,C,*sp] fp32 -> engine Act (channels-last, optional zero channel padding)."""
    from tqdne_b200.engine import Act

    N, C = x_nchw.shape[:2]
    sp = x_nchw.shape[2:]
    H, W = (sp if len(sp) == 2 else (1, sp[0]))
    xl = x_nchw.reshape(N, C, H * W).permute(0, 2, 1)
    if cpad and cpad > C:
        xl = F.pad(xl, (0, cpad - C))
    t = xl.contiguous().to(dtype).reshape(-1)
    return Act(t, N, H, W, cpad or C)


def _to_nchw(act, sp_dims):
    t = act.t.float().reshape(act.N, act.H * act.W, act.C).permute(0, 2, 1)
    shape = (act.N, act.C, act.H, act.W) if sp_dims == 2 else (act.N, act.C, act.W)
    return t.reshape(shape).contiguous()


def _rt(x, dtype):
    """round-trip through the activation dtype so the reference sees the same operands"""
    return x.to(dtype).float()


def _ref_conv(x, w, b, stride=1, upsample=False):
    if upsample:
        x = F.interpolate(x, scale_factor=2, mode="nearest")
    fn = F.conv1d if w.dim() == 3 else F.conv2d
    return fn(x, w, b, stride=stride, padding=w.shape[-1] // 2)


CONV_CASES = [
    # name, N, spatial, cin, cout, k, stride, upsample
    ("3x3_32", 2, (32, 32), 128, 128, 3, 1, False),
    ("3x3_16", 3, (16, 16), 256, 256, 3, 1, False),
    ("3x3_8_oddN", 3, (8, 8), 128, 192, 3, 1, False),
    ("3x3_4_raggedN", 5, (4, 4), 256, 128, 3, 1, False),
    ("1x1_16", 2, (16, 16), 192, 64, 1, 1, False),
    ("k5_1d_partial", 2, (500,), 64, 64, 5, 1, False),
    ("k5_1d_long", 1, (4064,), 64, 128, 5, 1, False),
    ("s2_2d", 2, (16, 16), 128, 128, 3, 2, False),
    ("s2_1d", 2, (256,), 64, 64, 3, 2, False),
    ("up_2d", 2, (8, 8), 128, 128, 3, 1, True),
    ("up_1d_k5", 2, (64,), 64, 64, 5, 1, True),
    ("3x3_128", 1, (128, 128), 64, 64, 3, 1, False),
    ("many_tiles", 48, (32, 32), 128, 256, 3, 1, False),
]


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32], ids=["bf16", "fp32"])
@pytest.mark.parametrize("case", CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_conv_matches_torch(case, dtype):
    from tqdne_b200.engine import pack_conv

    name, N, sp, cin, cout, k, stride, up = case
    if dtype == torch.float32 and name == "many_tiles":
        N = 4
    g = torch.Generator(device="cuda").manual_seed(hash(name) % 1000)
    dims = len(sp)
    x = torch.randn(N, cin, *sp, device="cuda", generator=g)
    w = torch.randn(cout, cin, *([k] * dims), device="cuda", generator=g) / math.sqrt(cin * k**dims)
    b = torch.randn(cout, device="cuda", generator=g) * 0.1
    plan = _plan(dtype)
    pc = pack_conv(w, b, [cin], dtype)
    out = plan.conv(pc, [_act(x, dtype)], stride=stride, upsample=up, dims=dims)
    plan.run()
    torch.cuda.synchronize()
    ref = _ref_conv(_rt(x, dtype), _rt(w, dtype), b, stride, up)
    tol = 5e-3 if dtype == torch.bfloat16 else 1e-5
    assert rel_l2(_to_nchw(out, dims), ref) < tol


@pytest.mark.parametrize("cta_group", [1, 2])
@pytest.mark.parametrize("block_n", [64, 128, 256])
def test_conv_block_n_variants_fp32_out(block_n, cta_group):
    from tqdne_b200.engine import pack_conv

    g = torch.Generator(device="cuda").manual_seed(block_n)
    x = torch.randn(4, 128, 16, 16, device="cuda", generator=g)
    w = torch.randn(256, 128, 3, 3, device="cuda", generator=g) / 34.0
    b = torch.randn(256, device="cuda", generator=g) * 0.1
    plan = _plan(torch.bfloat16)
    out = plan.conv(pack_conv(w, b, [128], torch.bfloat16), [_act(x, torch.bfloat16)], out_dtype=torch.float32,
                    block_n=block_n, cta_group=cta_group)
    plan.run()
    torch.cuda.synchronize()
    ref = _ref_conv(_rt(x, torch.bfloat16), _rt(w, torch.bfloat16), b)
    assert rel_l2(_to_nchw(out, 2), ref) < 2e-5


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32], ids=["bf16", "fp32"])
def test_resblock_tail_concat_emb_shortcut(dtype):
    """conv over a two-source concat with per-sample embedding; then conv + fused 1x1 shortcut on the raw concat."""
    from tqdne_b200.engine import pack_conv

    g = torch.Generator(device="cuda").manual_seed(7)
    N, H = 3, 16
    a = torch.randn(N, 128, H, H, device="cuda", generator=g)
    s = torch.randn(N, 64, H, H, device="cuda", generator=g)
    w1 = torch.randn(128, 192, 3, 3, device="cuda", generator=g) / 41.0
    b1 = torch.randn(128, device="cuda", generator=g) * 0.1
    emb = torch.randn(N, 384, device="cuda", generator=g)
    plan = _plan(dtype)
    aa, sa = _act(a, dtype), _act(s, dtype)
    h = plan.conv(pack_conv(w1, b1, [128, 64], dtype), [aa, sa], emb=emb[:, 128:], emb_ld=384)
    w2 = torch.randn(128, 128, 3, 3, device="cuda", generator=g) / 34.0
    b2 = torch.randn(128, device="cuda", generator=g) * 0.1
    wsk = torch.randn(128, 192, 1, 1, device="cuda", generator=g) / 14.0
    bsk = torch.randn(128, device="cuda", generator=g) * 0.1
    pc2 = pack_conv(w2, b2, [128], dtype, shortcut=(wsk, bsk, [128, 64]))
    out = plan.conv(pc2, [h], shortcut_srcs=[aa, sa])
    plan.run()
    torch.cuda.synchronize()
    cat = torch.cat([_rt(a, dtype), _rt(s, dtype)], dim=1)
    h_ref = _ref_conv(cat, _rt(w1, dtype), b1) + emb[:, 128:256, None, None]
    tol = 6e-3 if dtype == torch.bfloat16 else 1e-5
    assert rel_l2(_to_nchw(h, 2), h_ref) < tol
    h_in = _to_nchw(h, 2)  # the engine's own (rounded) h feeds the second conv
    out_ref = _ref_conv(h_in, _rt(w2, dtype), b2) + _ref_conv(cat, _rt(wsk, dtype), bsk)
    assert rel_l2(_to_nchw(out, 2), out_ref) < tol


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32], ids=["bf16", "fp32"])
def test_conv_residual_and_ragged_fp32_output(dtype):
    from tqdne_b200.engine import pack_conv

    g = torch.Generator(device="cuda").manual_seed(9)
    x = torch.randn(2, 128, 32, 32, device="cuda", generator=g)
    w = torch.randn(128, 128, 3, 3, device="cuda", generator=g) / 34.0
    b = torch.randn(128, device="cuda", generator=g) * 0.1
    plan = _plan(dtype)
    xa = _act(x, dtype)
    y = plan.conv(pack_conv(w, b, [128], dtype), [xa], residual=xa)
    # UNet output conv: 128 -> 8 channels, fp32 result; stem: 8 (padded to 64) -> 128
    w8 = torch.randn(8, 128, 3, 3, device="cuda", generator=g) / 34.0
    b8 = torch.randn(8, device="cuda", generator=g) * 0.1
    y8 = plan.conv(pack_conv(w8, b8, [128], dtype), [xa], out_dtype=torch.float32)
    x8 = torch.randn(2, 8, 32, 32, device="cuda", generator=g)
    ws = torch.randn(128, 8, 3, 3, device="cuda", generator=g) / 8.5
    ys = plan.conv(pack_conv(ws, None, [8], dtype), [_act(x8, dtype, cpad=64)])
    plan.run()
    torch.cuda.synchronize()
    tol = 5e-3 if dtype == torch.bfloat16 else 1e-5
    xr = _rt(x, dtype)
    assert rel_l2(_to_nchw(y, 2), _ref_conv(xr, _rt(w, dtype), b) + xr) < tol
    assert rel_l2(_to_nchw(y8, 2), _ref_conv(xr, _rt(w8, dtype), b8)) < (2e-5 if dtype == torch.bfloat16 else 1e-5)
    assert rel_l2(_to_nchw(ys, 2), _ref_conv(_rt(x8, dtype), _rt(ws, dtype), None)) < tol


@pytest.mark.parametrize("cta_group", [1, 2])
@pytest.mark.parametrize("block_n", [64, 128, 256])
@pytest.mark.parametrize("shape", [(5, (8, 8)), (2, (32, 32)), (9, (4, 4)), (3, (300,))], ids=["8x8_oddtiles", "32x32", "4x4", "1d_ragged"])
def test_conv_tile_configs_bf16_epilogue(block_n, cta_group, shape):
    """Every (BN, CTA-group) instantiation of the tensor-core kernel with the full bf16 epilogue: bias + per-sample
    embedding + residual (TMA in), TMA store out, fused GroupNorm statistics; odd tile counts leave the second CTA
    of the last pair without rows."""
    from tqdne_b200.engine import pack_conv

    N, sp = shape
    dims = len(sp)
    g = torch.Generator(device="cuda").manual_seed(block_n + cta_group + N)
    cin = cout = 256
    x = torch.randn(N, cin, *sp, device="cuda", generator=g)
    w = torch.randn(cout, cin, *([3] * dims), device="cuda", generator=g) / math.sqrt(cin * 3**dims)
    b = torch.randn(cout, device="cuda", generator=g) * 0.2
    emb = torch.randn(N, 512, device="cuda", generator=g)
    plan = _plan(torch.bfloat16)
    xa = _act(x, torch.bfloat16)
    y = plan.conv(pack_conv(w, b, [cin], torch.bfloat16), [xa], residual=xa, emb=emb[:, 128:], emb_ld=512, dims=dims,
                  block_n=block_n, cta_group=cta_group, stats=True)
    plan.run()
    torch.cuda.synchronize()
    assert f"BN={block_n},CG={cta_group}" in plan.op_names()[-1]
    xr = _rt(x, torch.bfloat16)
    e = emb[:, 128:128 + cout]
    ref = _ref_conv(xr, _rt(w, torch.bfloat16), b) + e[(...,) + (None,) * dims] + xr
    yn = _to_nchw(y, dims)
    assert rel_l2(yn, ref) < 5e-3
    flat = yn.reshape(N, cout, -1).double()
    st = y.stats.reshape(N, y.stats_parts, cout, 2).double().sum(1)   # per-tile partial sums -> per sample
    assert rel_l2(st[..., 0], flat.sum(-1)) < 1e-4
    assert rel_l2(st[..., 1], (flat * flat).sum(-1)) < 1e-4
    # plain stores into single-writer slots, fixed summation order: a second run reproduces every bit
    first, out1 = y.stats.clone(), y.t.clone()
    plan.run()
    torch.cuda.synchronize()
    assert torch.equal(first, y.stats) and torch.equal(out1, y.t)


def test_linear_as_1x1_conv_on_tensor_path():
    """emb_layers GEMM: [M, 512] x [1536, 512]^T with fp32 output."""
    from tqdne_b200.engine import Act, pack_conv

    g = torch.Generator(device="cuda").manual_seed(3)
    M = 200
    x = torch.randn(M, 512, device="cuda", generator=g)
    w = torch.randn(1536, 512, device="cuda", generator=g) / 22.0
    b = torch.randn(1536, device="cuda", generator=g)
    plan = _plan(torch.bfloat16)
    out = plan.conv(pack_conv(w, b, [512], torch.bfloat16), [Act(x.to(torch.bfloat16).reshape(-1), M, 1, 1, 512)],
                    out_dtype=torch.float32)
    plan.run()
    torch.cuda.synchronize()
    ref = _rt(x, torch.bfloat16) @ _rt(w, torch.bfloat16).t() + b
    assert rel_l2(out.t.reshape(M, 1536), ref) < 2e-5


def test_simt_and_tensor_paths_agree_on_bf16_operands(monkeypatch):
    """Same bf16 operands through the FFMA kernel (TQ_FORCE_SIMT=1) and the tcgen05 kernel."""
    from tqdne_b200.engine import pack_conv

    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(3, 256, 8, 8, device="cuda", generator=g)
    w = torch.randn(256, 256, 3, 3, device="cuda", generator=g) / 48.0
    outs = []
    for force in ("0", "1"):
        monkeypatch.setenv("TQ_FORCE_SIMT", force)
        plan = _plan(torch.bfloat16)
        o = plan.conv(pack_conv(w, None, [256], torch.bfloat16), [_act(x, torch.bfloat16)], out_dtype=torch.float32)
        plan.run()
        torch.cuda.synchronize()
        outs.append(o.t.clone())
        names = plan.op_names()
        assert ("simt" in names[0]) == (force == "1")
    assert rel_l2(outs[0], outs[1]) < 1e-5


GN_CASES = [(2, 1024, 128, 0, True), (3, 256, 512, 256, True), (2, 64, 128, 64, False), (5, 16, 512, 512, True),
            (1, 500, 64, 0, True), (2, 16384, 64, 0, True), (2, 1016, 256, 128, True)]


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32], ids=["bf16", "fp32"])
@pytest.mark.parametrize("case", GN_CASES, ids=[f"N{c[0]}_P{c[1]}_C{c[2]}+{c[3]}" for c in GN_CASES])
def test_groupnorm_silu_matches_torch(case, dtype):
    from tqdne_b200.engine import Act

    N, P, C0, C1, silu = case
    g = torch.Generator(device="cuda").manual_seed(P + C0)
    x0 = torch.randn(N, P, C0, device="cuda", generator=g) * 1.5 + 0.3
    x1 = torch.randn(N, P, C1, device="cuda", generator=g) * 0.7 - 0.2 if C1 else None
    gamma = 1 + 0.1 * torch.randn(C0 + C1, device="cuda", generator=g)
    beta = 0.1 * torch.randn(C0 + C1, device="cuda", generator=g)
    plan = _plan(dtype)
    srcs = [Act(x0.to(dtype).reshape(-1), N, 1, P, C0)]
    if C1:
        srcs.append(Act(x1.to(dtype).reshape(-1), N, 1, P, C1))
    out = plan.groupnorm(srcs, gamma, beta, silu)
    plan.run()
    torch.cuda.synchronize()
    cat = _rt(x0, dtype) if not C1 else torch.cat([_rt(x0, dtype), _rt(x1, dtype)], dim=2)
    ref = F.group_norm(cat.permute(0, 2, 1), 32, gamma, beta, eps=1e-5)
    if silu:
        ref = F.silu(ref)
    got = out.t.float().reshape(N, P, C0 + C1).permute(0, 2, 1)
    assert rel_l2(got, ref) < (4e-3 if dtype == torch.bfloat16 else 3e-6)


FUSED_GN_CASES = [
    # name, N, spatial, cin, cout, k, upsample  -- rows of one sample per epilogue warp: 32, 32, 16, 4, 32 (1D ragged)
    ("32x32", 3, (32, 32), 64, 128, 3, False),
    ("8x8", 5, (8, 8), 128, 256, 3, False),
    ("4x4", 9, (4, 4), 128, 512, 3, False),
    ("2x2", 7, (2, 2), 64, 64, 1, False),
    ("1d_ragged", 2, (500,), 64, 64, 5, False),
    ("up_2d", 3, (8, 8), 64, 128, 3, True),
    # two 128-row tiles per sample, four epilogue warps per sample inside a tile (cross-warp statistics pass)
    ("16x16", 3, (16, 16), 64, 128, 3, False),
    ("up_16x16", 2, (16, 16), 64, 128, 3, True),
    # 128 parts per sample: the per-sample group reduction gets its own launch (gn_finalize)
    ("128x128_many_parts", 2, (128, 128), 64, 128, 3, False),
]


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32], ids=["bf16", "fp32"])
@pytest.mark.parametrize("case", FUSED_GN_CASES, ids=[c[0] for c in FUSED_GN_CASES])
def test_conv_epilogue_statistics_feed_groupnorm(case, dtype):
    """The conv epilogue's per-(sample, channel) sums replace the GroupNorm statistics pass: statistics checked
    directly, then GN(+SiLU) of the conv output -- alone and as the first half of a concat -- against torch."""
    from tqdne_b200.engine import pack_conv

    name, N, sp, cin, cout, k, up = case
    g = torch.Generator(device="cuda").manual_seed(len(name) + cout)
    dims = len(sp)
    x = torch.randn(N, cin, *sp, device="cuda", generator=g)
    w = torch.randn(cout, cin, *([k] * dims), device="cuda", generator=g) / math.sqrt(cin * k**dims)
    b = torch.randn(cout, device="cuda", generator=g) * 0.5
    plan = _plan(dtype)
    y = plan.conv(pack_conv(w, b, [cin], dtype), [_act(x, dtype)], upsample=up, dims=dims, stats=True)
    assert y.stats is not None
    gamma = 1 + 0.1 * torch.randn(cout, device="cuda", generator=g)
    beta = 0.1 * torch.randn(cout, device="cuda", generator=g)
    o1 = plan.groupnorm([y], gamma, beta, silu=True)
    # second source of a concat: another conv output with its own statistics buffer
    w2 = torch.randn(64, cin, *([1] * dims), device="cuda", generator=g) / math.sqrt(cin)
    y2 = plan.conv(pack_conv(w2, None, [cin], dtype), [_act(x, dtype)], upsample=up, dims=dims, stats=True)
    gamma2 = 1 + 0.1 * torch.randn(cout + 64, device="cuda", generator=g)
    beta2 = 0.1 * torch.randn(cout + 64, device="cuda", generator=g)
    o2 = plan.groupnorm([y, y2], gamma2, beta2, silu=False)
    names = plan.op_names()
    assert not any("memset" in n or "gn_stats" in n for n in names)
    plan.run()
    torch.cuda.synchronize()
    first = [t.clone() for t in (y.stats, o1.t, o2.t)]
    plan.run()   # nothing is accumulated: the second run overwrites every slot with the same bits
    torch.cuda.synchronize()
    assert all(torch.equal(a, b) for a, b in zip(first, (y.stats, o1.t, o2.t)))
    yn = _to_nchw(y, dims)           # the stored (rounded) conv output
    flat = yn.reshape(N, cout, -1).double()
    st = y.stats.reshape(N, y.stats_parts, cout, 2).double().sum(1)
    tol_s = 1e-5 if dtype == torch.float32 else 1e-4
    assert rel_l2(st[..., 0], flat.sum(-1)) < 10 * tol_s   # sums of zero-mean data: looser relative bound
    assert rel_l2(st[..., 1], (flat * flat).sum(-1)) < tol_s
    tol = 4e-3 if dtype == torch.bfloat16 else 3e-6
    ref1 = F.silu(F.group_norm(yn, 32, gamma, beta, eps=1e-5))
    assert rel_l2(_to_nchw(o1, dims), ref1) < tol
    ref2 = F.group_norm(torch.cat([yn, _to_nchw(y2, dims)], dim=1), 32, gamma2, beta2, eps=1e-5)
    assert rel_l2(_to_nchw(o2, dims), ref2) < tol


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32], ids=["bf16", "fp32"])
@pytest.mark.parametrize("N,T,heads,d", [(3, 16, 4, 128), (2, 100, 4, 64), (1, 508, 4, 64), (2, 256, 2, 32),
                                         (2, 256, 4, 128), (3, 512, 4, 64), (1, 300, 2, 64), (2, 33, 1, 128),
                                         (1, 129, 3, 64), (1, 600, 2, 64), (5, 32, 3, 64), (9, 16, 2, 64), (2, 20, 2, 64)])
def test_attention_matches_reference_formula(N, T, heads, d, dtype):
    """bf16 with d in {64, 128}: T in {16, 32} packed tcgen05 kernel (128 / T pairs per tile; 12, 15 and 18 pairs leave a
    ragged last CTA), 32 < T <= 128: per-block tcgen05 kernel, 128 < T <= 512: multi-block kernel with P in tensor
    memory (tq_attn_sm100.cu; ragged T exercises the TMA zero fill and the key mask, T = 300 the 256 + 128 key split
    and O at column Tk / 4, 129 the half-empty softmax split and a second query block of one row, 256 x d = 128 O beside
    S); everything else (fp32, d = 32, T > 512): FFMA kernel."""
    from tqdne_b200.engine import Act

    g = torch.Generator(device="cuda").manual_seed(T)
    C = heads * d
    qkv = torch.randn(N, T, 3 * C, device="cuda", generator=g)
    plan = _plan(dtype)
    out = plan.attention(Act(qkv.to(dtype).reshape(-1), N, 1, T, 3 * C), heads)
    tc = dtype == torch.bfloat16 and d in (64, 128) and (32 < T <= 512 or T in (16, 32))
    assert ("attention_tc" in plan.op_names()[-1]) == tc, plan.op_names()
    plan.run()
    torch.cuda.synchronize()
    # reference formula (tqdne/blocks.py:156-190) on [N, 3C, T]
    x = _rt(qkv, dtype).permute(0, 2, 1)
    q, k, v = x.chunk(3, dim=1)
    s = 1 / math.sqrt(math.sqrt(d))
    w = torch.einsum("bct,bcs->bts", (q * s).reshape(N * heads, d, T), (k * s).reshape(N * heads, d, T))
    w = torch.softmax(w.float(), dim=-1)
    a = torch.einsum("bts,bcs->bct", w, v.reshape(N * heads, d, T)).reshape(N, C, T)
    got = out.t.float().reshape(N, T, C).permute(0, 2, 1)
    assert rel_l2(got, a) < (4e-3 if dtype == torch.bfloat16 else 3e-6)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32], ids=["bf16", "fp32"])
@pytest.mark.parametrize("N,T,heads,d", [(2, 16, 2, 128), (1, 27, 2, 64), (2, 100, 2, 64), (1, 508, 2, 64), (1, 130, 1, 32)])
def test_causal_attention_matches_reference_formula(N, T, heads, d, dtype):
    """use_causal_mask (tqdne/blocks.py:181-186): scores of keys behind the query are filled with -inf before the fp32
    softmax.  Off in every shipped config; both modes run it on the FFMA kernels (tq_attn.cu)."""
    from tqdne_b200.engine import Act

    g = torch.Generator(device="cuda").manual_seed(100 + T)
    C = heads * d
    qkv = torch.randn(N, T, 3 * C, device="cuda", generator=g)
    plan = _plan(dtype)
    out = plan.attention(Act(qkv.to(dtype).reshape(-1), N, 1, T, 3 * C), heads, causal=True)
    assert "causal" in plan.op_names()[-1] and "attention_tc" not in plan.op_names()[-1]
    plan.run()
    torch.cuda.synchronize()
    x = _rt(qkv, dtype).permute(0, 2, 1)
    q, k, v = x.chunk(3, dim=1)
    s = 1 / math.sqrt(math.sqrt(d))
    w = torch.einsum("bct,bcs->bts", (q * s).reshape(N * heads, d, T), (k * s).reshape(N * heads, d, T))
    mask = torch.tril(torch.ones(T, T, device="cuda")).unsqueeze(0).expand(w.size(0), -1, -1)
    w = torch.softmax(w.masked_fill(mask == 0, -torch.inf).float(), dim=-1)
    a = torch.einsum("bts,bcs->bct", w, v.reshape(N * heads, d, T)).reshape(N, C, T)
    got = out.t.float().reshape(N, T, C).permute(0, 2, 1)
    assert rel_l2(got, a) < (4e-3 if dtype == torch.bfloat16 else 3e-6)
    # the first query sees only itself: its output is v[0]
    assert rel_l2(got[:, :, 0], v.reshape(N, C, T)[:, :, 0]) < (4e-3 if dtype == torch.bfloat16 else 1e-6)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32], ids=["bf16", "fp32"])
@pytest.mark.parametrize("N,H,W,C", [(2, 1, 64, 64), (3, 1, 37, 128), (2, 16, 16, 64), (1, 7, 10, 72)])
def test_conv_less_resamplers_match_torch(N, H, W, C, dtype):
    """conv_resample=False: nn.AvgPool{1,2}d(2, 2) (odd sizes floor) and F.interpolate(scale_factor=2, "nearest")."""
    from tqdne_b200.engine import Act

    g = torch.Generator(device="cuda").manual_seed(N * 100 + W)
    x = _rt(torch.randn(N, H, W, C, device="cuda", generator=g), dtype)
    plan = _plan(dtype)
    xa = Act(x.to(dtype).reshape(-1), N, H, W, C)
    down, up = plan.resample2(xa, "avg_pool"), plan.resample2(xa, "nearest")
    plan.run()
    torch.cuda.synchronize()
    xn = x.permute(0, 3, 1, 2)   # [N, C, H, W]
    if H == 1:
        rd = F.avg_pool1d(xn[:, :, 0], 2, 2)[:, :, None]
        ru = F.interpolate(xn[:, :, 0], scale_factor=2, mode="nearest")[:, :, None]
    else:
        rd, ru = F.avg_pool2d(xn, 2, 2), F.interpolate(xn, scale_factor=2, mode="nearest")
    assert (down.H, down.W) == tuple(rd.shape[2:]) and (up.H, up.W) == tuple(ru.shape[2:])
    got_d = down.t.float().reshape(N, down.H, down.W, C).permute(0, 3, 1, 2)
    got_u = up.t.float().reshape(N, up.H, up.W, C).permute(0, 3, 1, 2)
    assert torch.equal(got_u, ru)
    assert rel_l2(got_d, rd) < (3e-3 if dtype == torch.bfloat16 else 1e-7)


def test_embedding_mlp_kernels():
    plan = _plan(torch.bfloat16)
    g = torch.Generator(device="cuda").manual_seed(1)
    M, mc, E = 6, 128, 512
    t = torch.randn(1, device="cuda", generator=g)
    Wf = torch.randn(mc // 2, device="cuda", generator=g) * 0.02
    W0 = torch.randn(E, mc, device="cuda", generator=g) / 11.0
    b0 = torch.randn(E, device="cuda", generator=g) * 0.1
    W2 = torch.randn(E, E, device="cuda", generator=g) / 22.0
    b2 = torch.randn(E, device="cuda", generator=g) * 0.1
    cemb = torch.randn(M, E, device="cuda", generator=g)
    feat = torch.empty(1, mc, device="cuda")
    h1 = torch.empty(1, E, device="cuda")
    emb = torch.empty(M, E, device="cuda")
    act = torch.empty(M, E, device="cuda", dtype=torch.bfloat16)
    plan.fourier(t, Wf, 1, feat)
    plan.linear(feat, W0, b0, 1, y=h1)
    plan.linear(h1, W2, b2, M, x_rows=1, act_in=True, add=cemb, add_rows=M, y=emb, y_act=act)
    plan.run()
    torch.cuda.synchronize()
    h = t[:, None] * Wf[None, :] * 2 * torch.pi
    f_ref = torch.cat([torch.sin(h), torch.cos(h)], dim=-1)
    assert rel_l2(feat, f_ref) < 1e-6
    e_ref = F.linear(F.silu(F.linear(f_ref, W0, b0)), W2, b2) + cemb
    assert rel_l2(emb, e_ref) < 2e-6
    assert rel_l2(act.float(), F.silu(e_ref)) < 4e-3


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32], ids=["bf16", "fp32"])
def test_edm_update_kernels(dtype):
    """precondition / euler / heun against the reference expressions (tqdne/edm.py:105-113,176-194)."""
    from tqdne_b200 import _lib
    from tqdne_b200.engine import tq_dtype

    lib = _lib.lib()
    g = torch.Generator(device="cuda").manual_seed(2)
    NP, C, Cpad, Cf = 1000, 8, 64, 8
    x = torch.randn(NP, C, device="cuda", dtype=torch.float64, generator=g) * 5
    Fo = torch.randn(NP, Cf, device="cuda", generator=g)
    F2 = torch.randn(NP, Cf, device="cuda", generator=g)
    xin = torch.full((NP, Cpad), 7.0, device="cuda", dtype=dtype)
    d = torch.empty_like(x)
    x1 = torch.empty_like(x)
    sd = 0.5
    sig, sig_n = torch.tensor(3.0), torch.tensor(1.7)
    c = lambda s: (float(1 / (s**2 + sd**2) ** 0.5), float(s * sd / (s**2 + sd**2) ** 0.5), float(sd**2 / (s**2 + sd**2)))  # noqa
    (ci, co, cs), (ci2, co2, cs2) = c(sig), c(sig_n)
    dt = float(sig_n - sig)
    st = torch.cuda.current_stream().cuda_stream
    tnext = torch.zeros(1, device="cuda")     # the plan's time input: the sampler kernels also store the next c_noise
    _lib.check(lib.tq_edm_precondition(x.data_ptr(), xin.data_ptr(), tq_dtype(dtype), NP, C, Cpad, ci, tnext.data_ptr(), 0.25, st))
    torch.cuda.synchronize()
    assert float(tnext) == 0.25
    ref_in = (x.float() * ci).to(dtype)
    assert torch.equal(xin[:, :C], ref_in) and float(xin[:, C:].abs().max()) == 0.0
    _lib.check(lib.tq_edm_euler(x.data_ptr(), Fo.data_ptr(), Cf, d.data_ptr(), x1.data_ptr(), xin.data_ptr(), tq_dtype(dtype),
                                NP, C, Cpad, co, cs, float(sig), dt, ci2, 1, tnext.data_ptr(), -1.5, st))
    torch.cuda.synchronize()
    assert float(tnext) == -1.5
    D = (Fo * co + cs * x.float()).double()
    d_ref = (x - D) / float(sig)
    x1_ref = x + d_ref * dt
    assert rel_l2(d, d_ref) < 1e-12 and rel_l2(x1, x1_ref) < 1e-12
    assert rel_l2(xin[:, :C].float(), (x1_ref.float() * ci2)) < (4e-3 if dtype == torch.bfloat16 else 1e-6)
    xs = x.clone()
    _lib.check(lib.tq_edm_heun(xs.data_ptr(), x1.data_ptr(), d.data_ptr(), F2.data_ptr(), Cf, xin.data_ptr(), tq_dtype(dtype),
                               NP, C, Cpad, co2, cs2, float(sig_n), dt, ci2, 1, None, 0.0, st))
    torch.cuda.synchronize()
    assert float(tnext) == -1.5               # NULL t_next: left alone
    D2 = (F2 * co2 + cs2 * x1.float()).double()
    x_ref = x + dt * (0.5 * d_ref + 0.5 * (x1 - D2) / float(sig_n))
    assert rel_l2(xs, x_ref) < 1e-12


def test_layout_kernels_roundtrip():
    from tqdne_b200.engine import nchw_to_nhwc, nhwc_to_nchw

    x = torch.randn(3, 8, 32, 32, device="cuda", dtype=torch.float64)
    y = nchw_to_nhwc(x, torch.float64)
    assert torch.equal(y, x.reshape(3, 8, 1024).permute(0, 2, 1))
    yp = nchw_to_nhwc(x.float(), torch.bfloat16, 64)
    assert yp.shape == (3, 1024, 64) and float(yp[..., 8:].abs().max()) == 0
    assert torch.equal(yp[..., :8], x.float().reshape(3, 8, 1024).permute(0, 2, 1).to(torch.bfloat16))
    back = nhwc_to_nchw(y, 3, 8, (32, 32), 8, torch.float32)
    assert torch.equal(back, x.float())
    x1 = torch.randn(2, 6, 4064, device="cuda")
    assert torch.equal(nhwc_to_nchw(nchw_to_nhwc(x1, torch.float32), 2, 6, (4064,), 6, torch.float32), x1)


def test_mavg_envelope_inverse_matches_golden():
    from tests.helpers import golden
    from tqdne_b200.representation import MovingAverageEnvelope

    g = golden("mavg_inverse")
    w = MovingAverageEnvelope().invert_representation(g["rep"].cuda())
    assert w.shape == tuple(g["wave"].shape) and rel_l2(w, g["wave"]) < 1e-6


def test_griffinlim_kernel_matches_oracle(monkeypatch):
    """Batched Griffin-Lim kernels vs the NumPy restatement, identical initial phases (full 128 iterations would take
    the oracle ~1 s per item; 16 iterations on 6 items here, 128 iterations on 1 item below).  fp64 (the default: the
    reference's locked NumPy 2 runs complex128) is the parity mode: fused kernel, and the independent unfused kernel,
    to 1e-9.  fp32 is the opt-in fast mode: it is held to NumPy's OWN complex64-vs-complex128 deviation on the same
    input (Griffin-Lim amplifies rounding), not to a parity bound."""
    from oracle import griffinlim_ref
    from tqdne_b200.representation import LogSpectrogram

    g = torch.Generator().manual_seed(4)
    rep = torch.tanh(torch.randn(2, 3, 128, 128, generator=g) * 0.5)
    ref = griffinlim_ref.logspec_inverse(rep.numpy(), n_iter=16, precision="fp64")
    ref32 = griffinlim_ref.logspec_inverse(rep.numpy(), n_iter=16, precision="fp32")
    numpy32 = rel_l2(ref32.astype(np.float64), ref)
    ls = LogSpectrogram(stft_channels=256, hop_size=32)
    assert ls.precision == "fp64"
    ls.n_iter = 16
    w = ls.invert_representation(rep.cuda())
    assert w.shape == (2, 3, 4064) and w.dtype == np.float64
    assert rel_l2(w, ref) < 1e-9
    monkeypatch.setenv("TQ_GL_LEGACY", "1")
    assert rel_l2(ls.invert_representation(rep.cuda()), ref) < 1e-9
    monkeypatch.delenv("TQ_GL_LEGACY")
    ls32 = LogSpectrogram(stft_channels=256, hop_size=32, precision="fp32")
    ls32.n_iter = 16
    w32 = ls32.invert_representation(rep.cuda())
    e32 = rel_l2(w32.astype(np.float64), ref)
    print(f"16 iterations: fp32 kernel vs fp64 oracle {e32:.2e}; NumPy complex64 vs complex128 {numpy32:.2e}")
    assert w32.dtype == np.float32 and e32 < max(20 * numpy32, 1e-4)


@pytest.mark.parametrize("frames", [6, 12, 40, 128])
def test_griffinlim_other_frame_counts_and_both_cta_shapes(frames, monkeypatch):
    """Frame counts besides the reference's 128: 6 frames = a signal shorter than two windows (the kernel then keeps one
    window-sum entry per sample), 12 / 40 = the compact window-sum table with a short interior; each through the two-CTA 8-warp
    kernel (the fp64 default) and the one-CTA 16-warp kernel (TQ_GL_WARPS16=1), against the NumPy oracle and each other."""
    from oracle import griffinlim_ref
    from tqdne_b200.representation import LogSpectrogram

    g = torch.Generator().manual_seed(frames)
    rep = torch.tanh(torch.randn(2, 3, 128, frames, generator=g) * 0.5)
    ref = griffinlim_ref.logspec_inverse(rep.numpy(), n_iter=12, precision="fp64")
    ls = LogSpectrogram(stft_channels=256, hop_size=32)
    ls.n_iter = 12
    w8 = ls.invert_representation(rep.cuda())
    monkeypatch.setenv("TQ_GL_WARPS16", "1")
    w16 = ls.invert_representation(rep.cuda())
    assert w8.shape == ref.shape == (2, 3, 32 * (frames - 1))
    assert rel_l2(w8, ref) < 1e-9 and rel_l2(w16, ref) < 1e-9
    # same arithmetic per frame, and a sample's overlapping frames are added in round order in both shapes: bit-identical
    assert np.array_equal(w8, w16)


def test_griffinlim_full_iterations_and_golden():
    from oracle import griffinlim_ref
    from tests.helpers import golden
    from tqdne_b200.representation import LogSpectrogram

    g = golden("logspec_inverse_iter8")
    ls = LogSpectrogram(stft_channels=256, hop_size=32, precision="fp64")
    ls.n_iter = int(g["n_iter"])
    assert rel_l2(ls.invert_representation(g["rep"].cuda()), g["wave"]) < 1e-9
    rep = g["rep"][:, :1]
    ls.n_iter = 128
    ref = griffinlim_ref.logspec_inverse(rep.numpy(), n_iter=128, precision="fp64")
    w64 = ls.invert_representation(rep.cuda())
    assert rel_l2(w64, ref) < 1e-7          # 128 iterations amplify the last-bit differences of the fp64 arithmetic ~1e3x
    assert np.array_equal(w64, ls.invert_representation(rep.cuda()))   # bit-reproducible
    ref32 = griffinlim_ref.logspec_inverse(rep.numpy(), n_iter=128, precision="fp32")
    numpy32 = rel_l2(ref32.astype(np.float64), ref)
    ls32 = LogSpectrogram(stft_channels=256, hop_size=32, precision="fp32")
    w32 = ls32.invert_representation(rep.cuda())
    e32 = rel_l2(w32.astype(np.float64), ref)
    print(f"128 iterations: fp32 kernel vs fp64 oracle {e32:.2e}; NumPy complex64 vs complex128 {numpy32:.2e}")
    assert w32.dtype == np.float32 and e32 < max(20 * numpy32, 1e-3)
    # domain property at full size: kernel and oracle reach the same spectral inconsistency
    S = np.exp((rep.numpy()[0, 0].astype(np.float64) + 1) / 2 * (3 - np.log(1e-8)) + np.log(1e-8))

    def inconsistency(wave):
        return np.linalg.norm(np.abs(griffinlim_ref.stft(wave.astype(np.float64)))[:-1] - S) / np.linalg.norm(S)

    assert abs(inconsistency(w32[0, 0]) - inconsistency(ref[0, 0])) < 2e-2
    assert abs(inconsistency(w64[0, 0]) - inconsistency(ref[0, 0])) < 1e-9


# ---- forward representations ---------------------------------------------------------------
def test_mavg_envelope_forward_matches_golden_and_oracle():
    import tqdne_b200 as tq
    from oracle import griffinlim_ref
    from tests.helpers import golden

    rep = tq.MovingAverageEnvelope()
    g = golden("mavg_forward")
    out = rep.get_representation(g["wave"].cuda())
    assert out.shape == tuple(g["rep"].shape) and out.dtype == np.float32
    assert rel_l2(out, g["rep"]) < 1e-6
    # ragged length, window longer than the signal's edge regions, and the round trip through the inverse kernel
    rng = np.random.default_rng(7)
    w = rng.standard_normal((3, 3, 777)).astype(np.float32)
    out = rep.get_representation(w)
    assert rel_l2(out, griffinlim_ref.mavg_forward(w)) < 1e-6
    assert rel_l2(rep.invert_representation(out), w) < 1e-5


@pytest.mark.parametrize("precision", ["fp32", "fp64"])
def test_logspec_forward_matches_golden_and_oracle(precision):
    import tqdne_b200 as tq
    from oracle import griffinlim_ref
    from tests.helpers import golden

    ls = tq.LogSpectrogram(stft_channels=256, hop_size=32, precision=precision)  # SpectrogramConfig
    g = golden("logspec_forward")
    out = ls.get_representation(g["wave"].cuda())
    assert out.shape == (1, 3, 128, 128)
    ref = g["rep64" if precision == "fp64" else "rep32"].numpy()
    # values live in [-1, 1]; a float32 STFT of O(1) magnitudes leaves ~1e-6 on the log-magnitudes
    assert np.abs(out - ref).max() < (1e-9 if precision == "fp64" else 2e-5)
    rng = np.random.default_rng(8)
    w = rng.standard_normal((5, 3, 4064)).astype(np.float32)
    w[0, 0] = 0.0            # silent channel: every bin clips to the floor -> -1
    out = ls.get_representation(w)
    ref = griffinlim_ref.logspec_forward(w.astype(np.float64) if precision == "fp64" else w)
    assert np.abs(out - ref).max() < (1e-9 if precision == "fp64" else 2e-5)
    assert np.all(out[0, 0] == -1.0)


def test_logspec_forward_full_batch_scaling_property():
    """Full bench size (256 x 3 items): scaling a waveform by a shifts every unclipped log-magnitude by
    2 log(a) / (log_max - log_clip) -- a size-independent check of all 768 CTAs."""
    import tqdne_b200 as tq

    ls = tq.LogSpectrogram(stft_channels=256, hop_size=32)
    gen = torch.Generator(device="cuda").manual_seed(9)
    w = torch.randn(256, 3, 4064, device="cuda", generator=gen)
    a = 3.5
    r1 = ls.get_representation_device(w)
    r2 = ls.get_representation_device(w * a)
    shift = 2 * np.log(a) / (ls.log_max - ls.log_clip)
    assert r1.shape == (256, 3, 128, 128)
    err = (r2 - r1 - shift).abs()
    # fp32 STFT: bins where the frame nearly cancels carry a larger relative error, which the log amplifies
    assert float(err.max()) < 5e-3 and float((err > 1e-4).float().mean()) < 1e-4


@pytest.mark.parametrize("N,L,cin,cout,taps", [(3, 200, 128, 192, 5), (2, 64, 64, 64, 1), (2, 1016, 256, 256, 5),
                                                (1, 77, 64, 128, 3), (4, 508, 512, 256, 5), (2, 130, 64, 64, 7)])
def test_conv1d_wgrad_matches_autograd(N, L, cin, cout, taps):
    """tq_conv1d_wgrad (tcgen05, MN-major operands, taps through descriptor row offsets of one halo buffer) against
    the weight / bias gradient torch autograd computes for F.conv1d(padding='same') in fp32 on the same bf16-rounded
    tensors; ragged L exercises the TMA zero fill at both ends of a sample."""
    from tqdne_b200 import _lib
    from tqdne_b200.engine import current_stream_ptr

    g = torch.Generator(device="cuda").manual_seed(L + taps)
    x = torch.randn(N, L, cin, device="cuda", generator=g).to(torch.bfloat16)
    dy = torch.randn(N, L, cout, device="cuda", generator=g).to(torch.bfloat16)
    dw = torch.zeros(cout, taps, cin, device="cuda")
    db = torch.zeros(cout, device="cuda")
    _lib.check(_lib.lib().tq_conv1d_wgrad(x.data_ptr(), dy.data_ptr(), dw.data_ptr(), db.data_ptr(), N, L, cin, cout, taps,
                                          0, 0, current_stream_ptr()), "conv1d_wgrad")
    torch.cuda.synchronize()
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        w = torch.zeros(cout, cin, taps, device="cuda", dtype=torch.float64, requires_grad=True)
        b = torch.zeros(cout, device="cuda", dtype=torch.float64, requires_grad=True)
        y = F.conv1d(x.double().permute(0, 2, 1), w, b, padding=taps // 2)
        y.backward(dy.double().permute(0, 2, 1))
    finally:
        torch.backends.cudnn.allow_tf32 = old
    assert rel_l2(dw, w.grad.permute(0, 2, 1)) < 2e-5
    assert rel_l2(db, b.grad) < 2e-5


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32], ids=["bf16", "fp32"])
@pytest.mark.parametrize("N,P,C0,C1,silu", [(3, 200, 64, 0, True), (2, 1016, 256, 0, True), (2, 300, 256, 128, True),
                                            (2, 64, 128, 64, False), (5, 16, 512, 512, True), (1, 4064, 64, 128, True)])
def test_groupnorm_silu_backward_matches_autograd(N, P, C0, C1, silu, dtype):
    """tq_gn_silu_backward against autograd through silu(group_norm(cat[x0, x1])) on the same (rounded) tensors; C = 384
    and 192 have groups that straddle the concat boundary."""
    from tqdne_b200.backward import groupnorm_silu_backward
    from tqdne_b200.engine import Act

    g = torch.Generator(device="cuda").manual_seed(P + C0)
    Ct = C0 + C1
    x = _rt(torch.randn(N, P, Ct, device="cuda", generator=g) * 1.3 + 0.4, dtype)
    dy = _rt(torch.randn(N, P, Ct, device="cuda", generator=g), dtype)
    gamma = 1 + 0.2 * torch.randn(Ct, device="cuda", generator=g)
    beta = 0.3 * torch.randn(Ct, device="cuda", generator=g)

    def act(t):
        st = torch.stack([t.sum(1), (t * t).sum(1)], dim=-1).contiguous()   # [N, C, 2] like a conv epilogue writes
        return Act(t.contiguous().to(dtype).reshape(-1), N, 1, P, t.shape[2], stats=st)

    a0 = act(x[..., :C0])
    a1 = act(x[..., C0:]) if C1 else None
    dx0, dx1, dgam, dbet = groupnorm_silu_backward(a0, Act(dy.to(dtype).reshape(-1), N, 1, P, Ct), gamma, beta, silu=silu, x1=a1)
    torch.cuda.synchronize()
    xr = x.double().permute(0, 2, 1).clone().requires_grad_(True)
    gr, br = gamma.double().clone().requires_grad_(True), beta.double().clone().requires_grad_(True)
    y = F.group_norm(xr, 32, gr, br, eps=1e-5)
    if silu:
        y = F.silu(y)
    y.backward(dy.double().permute(0, 2, 1))
    want = xr.grad.permute(0, 2, 1)
    tol = 5e-3 if dtype == torch.bfloat16 else 2e-5
    assert rel_l2(dx0.t.float().reshape(N, P, C0), want[..., :C0]) < tol
    if C1:
        assert rel_l2(dx1.t.float().reshape(N, P, C1), want[..., C0:]) < tol
    ptol = 2e-3 if dtype == torch.bfloat16 else 2e-5
    assert rel_l2(dgam, gr.grad) < ptol and rel_l2(dbet, br.grad) < ptol


@pytest.mark.parametrize("N,L,cin,cout,k", [(2, 200, 128, 192, 5), (3, 508, 256, 256, 5), (2, 100, 64, 128, 1)])
def test_conv1d_input_grad_matches_autograd(N, L, cin, cout, k):
    """dX of a stride-1 'same' conv1d = the forward tcgen05 implicit GEMM over dY with tap-flipped, transposed weights
    (tqdne_b200.backward.conv1d_input_grad), with and without accumulation into an existing gradient."""
    from tqdne_b200.backward import conv1d_input_grad

    dtype = torch.bfloat16
    g = torch.Generator(device="cuda").manual_seed(L + k)
    w = _rt(torch.randn(cout, cin, k, device="cuda", generator=g) / math.sqrt(cin * k), dtype)
    dy = _rt(torch.randn(N, cout, L, device="cuda", generator=g), dtype)
    prev = _rt(torch.randn(N, cin, L, device="cuda", generator=g), dtype)
    plan = _plan(dtype)
    dx = conv1d_input_grad(plan, w, _act(dy, dtype))
    dx_acc = conv1d_input_grad(plan, w, _act(dy, dtype), accumulate_into=_act(prev, dtype))
    plan.run()
    torch.cuda.synchronize()
    x = torch.zeros(N, cin, L, device="cuda", dtype=torch.float64, requires_grad=True)
    F.conv1d(x, w.double(), padding=k // 2).backward(dy.double())
    assert rel_l2(_to_nchw(dx, 1), x.grad) < 4e-3
    assert rel_l2(_to_nchw(dx_acc, 1), x.grad + prev) < 4e-3


@pytest.mark.parametrize("N,L,C", [(3, 508, 128), (2, 1016, 256)])
def test_resblock1d_forward_backward_matches_autograd(N, L, C):
    """One 1D UNet ResBlock (tqdne/unet.py:42-143, identity skip, k = 5, embedding add) forward through the plan kernels
    and backward through tqdne_b200.backward -- input gradient, both convolutions' weight / bias gradients, both
    GroupNorm affine gradients and the embedding gradient -- against autograd on the fp32 reference block (bf16 mode)."""
    from tqdne_b200 import backward as bw
    from tqdne_b200.engine import Act, pack_conv

    dt = torch.bfloat16
    g = torch.Generator(device="cuda").manual_seed(C + L)
    rn = lambda *s: torch.randn(*s, device="cuda", generator=g)  # noqa: E731
    w0 = _rt(rn(C, 64, 1) / 8, dt)          # stem 1x1 conv: produces x WITH epilogue statistics
    u = _rt(rn(N, 64, L), dt)
    w1, b1 = _rt(rn(C, C, 5) / math.sqrt(5 * C), dt), 0.1 * rn(C)
    w2, b2 = _rt(rn(C, C, 5) / math.sqrt(5 * C), dt), 0.1 * rn(C)
    g1, be1, g2, be2 = 1 + 0.1 * rn(C), 0.1 * rn(C), 1 + 0.1 * rn(C), 0.1 * rn(C)
    e = 0.5 * rn(N, C)
    dout = _rt(rn(N, C, L), dt)

    # ---- forward on the engine
    plan = _plan(dt)
    x = plan.conv(pack_conv(w0, None, [64], dt), [_act(u, dt)], dims=1, stats=True)
    h0 = plan.groupnorm([x], g1, be1, silu=True)
    h1 = plan.conv(pack_conv(w1, b1, [C], dt), [h0], emb=e.contiguous(), emb_ld=C, dims=1, stats=True)
    h2 = plan.groupnorm([h1], g2, be2, silu=True)
    out = plan.conv(pack_conv(w2, b2, [C], dt), [h2], residual=x, dims=1)
    plan.run()
    # ---- backward on the engine
    d_out = _act(dout, dt)
    p1 = _plan(dt)
    dh2 = bw.conv1d_input_grad(p1, w2, d_out)
    p1.run()
    as3 = lambda a: a.t.reshape(a.N, a.W, a.C)  # noqa: E731
    dw2, db2 = bw.conv1d_weight_grad(as3(h2), as3(d_out), 5)
    dh1, _, dg2, dbe2 = bw.groupnorm_silu_backward(h1, dh2, g2, be2)
    de = bw.sample_channel_sums(dh1)
    p2 = _plan(dt)
    dh0 = bw.conv1d_input_grad(p2, w1, dh1)
    p2.run()
    dw1, db1 = bw.conv1d_weight_grad(as3(h0), as3(dh1), 5)
    dx_main, _, dg1, dbe1 = bw.groupnorm_silu_backward(x, dh0, g1, be1)
    torch.cuda.synchronize()
    dx = _to_nchw(dx_main, 1) + dout                      # identity skip: the two gradient paths of x add

    # ---- fp32 reference block under autograd, fed with the engine's (rounded) block input
    xr = _to_nchw(x, 1).double().requires_grad_(True)
    P = [t.double().clone().requires_grad_(True) for t in (w1, b1, w2, b2, g1, be1, g2, be2, e)]
    rw1, rb1, rw2, rb2, rg1, rbe1, rg2, rbe2, re_ = P
    hh = F.conv1d(F.silu(F.group_norm(xr, 32, rg1, rbe1, eps=1e-5)), rw1, rb1, padding=2) + re_[..., None]
    ref = xr + F.conv1d(F.silu(F.group_norm(hh, 32, rg2, rbe2, eps=1e-5)), rw2, rb2, padding=2)
    ref.backward(dout.double())
    assert rel_l2(_to_nchw(out, 1), ref.detach()) < 5e-3
    tol = 2e-2   # gradients pass through two bf16 activations-gradient tensors and two recomputed SiLUs
    assert rel_l2(dx, xr.grad) < tol
    assert rel_l2(dw2.permute(0, 2, 1), rw2.grad) < tol and rel_l2(db2, rb2.grad) < tol
    assert rel_l2(dw1.permute(0, 2, 1), rw1.grad) < tol and rel_l2(db1, rb1.grad) < tol
    assert rel_l2(dg2, rg2.grad) < tol and rel_l2(dbe2, rbe2.grad) < tol
    assert rel_l2(dg1, rg1.grad) < tol and rel_l2(dbe1, rbe1.grad) < tol
    assert rel_l2(de, re_.grad) < tol


@pytest.mark.parametrize("N,T,heads", [(2, 508, 4), (3, 128, 2), (1, 300, 4), (2, 512, 1), (1, 77, 3)])
def test_attention_backward_matches_autograd(N, T, heads):
    """tq_attention_backward (two tcgen05 kernels: dQ + row statistics, then dK / dV with everything transposed) against
    autograd through the reference formula (tqdne/blocks.py:156-190) on the same bf16-rounded tensors; ragged T
    exercises the key / query masks of both kernels."""
    from tqdne_b200.backward import attention_backward
    from tqdne_b200.engine import Act

    d, dt = 64, torch.bfloat16
    C = heads * d
    g = torch.Generator(device="cuda").manual_seed(T + heads)
    qkv = _rt(torch.randn(N, T, 3 * C, device="cuda", generator=g), dt)
    dout = _rt(torch.randn(N, T, C, device="cuda", generator=g), dt)
    plan = _plan(dt)
    qa = Act(qkv.to(dt).reshape(-1), N, 1, T, 3 * C)
    ws = torch.full((2 * N * heads * T,), float("nan"), device="cuda")
    out = plan.attention(qa, heads, lse=ws)
    # the multi-block forward kernel (T > 128) leaves the rows' log-sum-exp for the backward; other shapes recompute it
    assert out.lse_written == (T > 128)
    plan.run()
    da = Act(dout.to(dt).reshape(-1), N, 1, T, C)
    dqkv = attention_backward(qa, out, da, heads, lse=ws if out.lse_written else None)
    torch.cuda.synchronize()
    if out.lse_written:
        # the handed-over statistic equals the one the backward computes for itself (exact fp32 sums on both sides), and
        # both routes give the same gradient
        from tqdne_b200 import _lib as L
        from tqdne_b200.engine import current_stream_ptr

        own = attention_backward(qa, out, da, heads, lse=None)
        ws2 = torch.empty_like(ws)
        L.check(L.lib().tq_attention_backward(qa.t.data_ptr(), out.t.data_ptr(), da.t.data_ptr(), torch.empty_like(qa.t).data_ptr(),
                                              ws2.data_ptr(), N, T, heads, d, 0, current_stream_ptr()), "attention_backward")
        torch.cuda.synchronize()
        assert float((ws[: N * heads * T] - ws2[: N * heads * T]).abs().max()) < 2e-4
        assert rel_l2(own.t.float(), dqkv.t.float()) < 2e-3
    x = qkv.double().permute(0, 2, 1).clone().requires_grad_(True)   # [N, 3C, T] like the reference
    q, k, v = x.chunk(3, dim=1)
    s = 1 / math.sqrt(math.sqrt(d))
    w = torch.einsum("bct,bcs->bts", (q * s).reshape(N * heads, d, T), (k * s).reshape(N * heads, d, T))
    a = torch.einsum("bts,bcs->bct", torch.softmax(w, dim=-1), v.reshape(N * heads, d, T)).reshape(N, C, T)
    a.backward(dout.double().permute(0, 2, 1))
    want = x.grad.permute(0, 2, 1)                                    # [N, T, 3C]
    got = dqkv.t.float().reshape(N, T, 3 * C)
    for name, sl in (("dq", slice(0, C)), ("dk", slice(C, 2 * C)), ("dv", slice(2 * C, 3 * C))):
        err = rel_l2(got[..., sl], want[..., sl])
        assert err < 1e-2, (name, err)


def test_training_helper_kernels_match_torch():
    """tq_rows_op (zero-stuff / nearest upsample / pair-sum / mul / add), tq_linear_backward, tq_edm_noise + tq_edm_loss,
    tq_dropout_apply and tq_repack_conv_weights against their torch definitions."""
    from tqdne_b200 import _lib
    from tqdne_b200.engine import current_stream_ptr

    lib, st = _lib.lib(), current_stream_ptr()
    g = torch.Generator(device="cuda").manual_seed(77)
    bf = torch.bfloat16
    N, L, C = 3, 50, 64
    src = torch.randn(N, L, C, device="cuda", generator=g).to(bf)
    aux = torch.randn(N, L, C, device="cuda", generator=g).to(bf)
    dst2 = torch.empty(N, 2 * L, C, device="cuda", dtype=bf)
    _lib.check(lib.tq_rows_op(src.data_ptr(), None, dst2.data_ptr(), 0, N, 2 * L, C, st), "zero_stuff")
    want = torch.zeros_like(dst2)
    want[:, 0::2] = src
    assert torch.equal(dst2, want)
    _lib.check(lib.tq_rows_op(src.data_ptr(), None, dst2.data_ptr(), 1, N, 2 * L, C, st), "upsample")
    assert torch.equal(dst2, src.repeat_interleave(2, dim=1))
    half = torch.empty(N, L // 2, C, device="cuda", dtype=bf)
    _lib.check(lib.tq_rows_op(src.data_ptr(), None, half.data_ptr(), 2, N, L // 2, C, st), "pair_sum")
    assert torch.equal(half, (src[:, 0::2].float() + src[:, 1::2].float()).to(bf))
    out = torch.empty_like(src)
    _lib.check(lib.tq_rows_op(src.data_ptr(), aux.data_ptr(), out.data_ptr(), 3, N, L, C, st), "mul")
    assert torch.equal(out, (src.float() * aux.float()).to(bf))
    _lib.check(lib.tq_rows_op(src.data_ptr(), aux.data_ptr(), out.data_ptr(), 4, N, L, C, st), "add")
    assert torch.equal(out, (src.float() + aux.float()).to(bf))
    # dropout: same decisions for the same seed, kept fraction ~ 1 - p, kept values scaled by 1 / (1 - p)
    big = torch.ones(1 << 16, device="cuda", dtype=bf)
    d1, d2, d3 = torch.empty_like(big), torch.empty_like(big), torch.empty_like(big)
    for d, seed in ((d1, 11), (d2, 11), (d3, 12)):
        _lib.check(lib.tq_dropout_apply(big.data_ptr(), d.data_ptr(), big.numel(), seed, 0.1, st), "dropout_apply")
    assert torch.equal(d1, d2) and not torch.equal(d1, d3)
    kept = (d1 != 0).float().mean().item()
    assert abs(kept - 0.9) < 0.01 and torch.allclose(d1[d1 != 0].float(), torch.tensor(1 / 0.9, device="cuda"), rtol=1e-2)
    # dense layer backward
    M, K, No = 5, 48, 37
    x, W, dy = (torch.randn(s, device="cuda", generator=g) for s in ((M, K), (No, K), (M, No)))
    for act in (0, 1):
        xr, Wr = x.clone().requires_grad_(True), W.clone().requires_grad_(True)
        br = torch.zeros(No, device="cuda", requires_grad=True)
        (F.linear(F.silu(xr) if act else xr, Wr, br)).backward(dy)
        dx, dW, db = torch.empty(M, K, device="cuda"), torch.zeros(No, K, device="cuda"), torch.zeros(No, device="cuda")
        _lib.check(lib.tq_linear_backward(dy.data_ptr(), x.data_ptr(), W.data_ptr(), act, dx.data_ptr(), dW.data_ptr(), db.data_ptr(),
                                          M, K, No, st), "linear_backward")
        assert rel_l2(dx, xr.grad) < 1e-5 and rel_l2(dW, Wr.grad) < 1e-5 and rel_l2(db, br.grad) < 1e-5
    # EDM noising + loss against LightningEDM.step's formulas (edm.py:105-134)
    Nn, P, Cc, Cp = 3, 40, 6, 64
    y, nz = torch.randn(Nn, P, Cc, device="cuda", generator=g), torch.randn(Nn, P, Cc, device="cuda", generator=g)
    sig = torch.tensor([0.3, 1.0, 4.0], device="cuda")
    xn, xin = torch.empty_like(y), torch.empty(Nn, P, Cp, device="cuda", dtype=bf)
    _lib.check(lib.tq_edm_noise(y.data_ptr(), nz.data_ptr(), sig.data_ptr(), xn.data_ptr(), xin.data_ptr(), Nn, P, Cc, Cp, 0.5, st), "noise")
    s3 = sig[:, None, None]
    assert rel_l2(xn, y + s3 * nz) < 1e-6
    assert rel_l2(xin[..., :Cc].float(), (xn / (s3**2 + 0.25).sqrt()).to(bf).float()) < 1e-6 and float(xin[..., Cc:].abs().max()) == 0
    Fo = torch.randn(Nn, P, Cc, device="cuda", generator=g).requires_grad_(True)
    c_out, c_skip, w = s3 * 0.5 / (s3**2 + 0.25).sqrt(), 0.25 / (s3**2 + 0.25), (s3**2 + 0.25) / (s3 * 0.5) ** 2
    ref = (w * (c_out * Fo + c_skip * xn - y) ** 2).mean()
    ref.backward()
    dF, loss = torch.empty(Nn, P, Cp, device="cuda", dtype=bf), torch.zeros(1, device="cuda")
    _lib.check(lib.tq_edm_loss(Fo.detach().data_ptr(), Cc, xn.data_ptr(), y.data_ptr(), sig.data_ptr(), dF.data_ptr(), loss.data_ptr(),
                               Nn, P, Cc, Cp, 0.5, st), "loss")
    assert abs(float(loss) - float(ref.detach())) < 1e-5 * abs(float(ref.detach()))
    assert rel_l2(dF[..., :Cc].float(), Fo.grad) < 4e-3 and float(dF[..., Cc:].abs().max()) == 0
    # operand copies of a convolution's fp32 master [Op, k, Ip]
    Op, k, Ip, off, Cs = 128, 5, 192, 64, 128
    m = torch.randn(Op, k, Ip, device="cuda", generator=g)
    fwd, bwd = torch.empty(Op, k * Ip, device="cuda", dtype=bf), torch.empty(Cs, k * Op, device="cuda", dtype=bf)
    _lib.check(lib.tq_repack_conv_weights(m.data_ptr(), fwd.data_ptr(), bwd.data_ptr(), Op, k, Ip, off, Cs, st), "repack")
    assert torch.equal(fwd, m.reshape(Op, -1).to(bf))
    assert torch.equal(bwd, m[:, :, off:off + Cs].flip(1).permute(2, 1, 0).reshape(Cs, -1).to(bf))
    # the same copies (and two more shapes: ragged 32 x 32 tiles, k = 1) through the one-launch job table
    import ctypes as C
    shapes = [(Op, k, Ip, off, Cs), (64, 3, 64, 0, 64), (192, 1, 320, 64, 72)]
    ms = [torch.randn(o, kk, i, device="cuda", generator=g) for o, kk, i, _, _ in shapes]
    fws = [torch.zeros(o, kk * i, device="cuda", dtype=bf) for o, kk, i, _, _ in shapes]
    bws = [torch.zeros(cs, kk * o, device="cuda", dtype=bf) for o, kk, _, _, cs in shapes]
    jobs = (_lib.TqRepackJob * (2 * len(shapes)))()
    for n, ((o, kk, i, of, cs), mm, fw, bw) in enumerate(zip(shapes, ms, fws, bws)):
        a, b2 = jobs[2 * n], jobs[2 * n + 1]
        a.master, a.fwd, a.bwd, a.Op, a.k, a.Ip, a.ci_off, a.Cs = mm.data_ptr(), fw.data_ptr(), None, o, kk, i, 0, 0
        b2.master, b2.fwd, b2.bwd, b2.Op, b2.k, b2.Ip, b2.ci_off, b2.Cs = mm.data_ptr(), None, bw.data_ptr(), o, kk, i, of, cs
    total = lib.tq_repack_batch_prepare(jobs, len(jobs))
    assert total == sum(j.nblocks for j in jobs) > 0 and jobs[0].block0 == 0
    tab = torch.frombuffer(bytearray(bytes(jobs)), dtype=torch.uint8).cuda()
    _lib.check(lib.tq_repack_batch_run(tab.data_ptr(), len(jobs), total, st), "repack batch")
    for (o, kk, i, of, cs), mm, fw, bw in zip(shapes, ms, fws, bws):
        assert torch.equal(fw, mm.reshape(o, -1).to(bf))
        assert torch.equal(bw, mm[:, :, of:of + cs].flip(1).permute(2, 1, 0).reshape(cs, -1).to(bf))
    bad = (_lib.TqRepackJob * 1)()
    assert lib.tq_repack_batch_prepare(bad, 1) == -1   # no master, neither copy requested


def test_groupnorm_fused_dropout_forward_and_backward_agree():
    """Training-mode GroupNorm+SiLU with nn.Dropout fused behind the activation (tq_gn_desc.drop_*): the forward's dropped
    elements are recovered from its output, and the backward kernel -- which regenerates the decisions from (seed, element
    index) -- must match autograd through silu(group_norm(x)) * mask with exactly that mask."""
    from tqdne_b200.backward import groupnorm_silu_backward
    from tqdne_b200.engine import Act

    dt, N, P, C, p = torch.bfloat16, 3, 300, 128, 0.25
    g = torch.Generator(device="cuda").manual_seed(4)
    x = _rt(torch.randn(N, P, C, device="cuda", generator=g) + 0.3, dt)
    dy = _rt(torch.randn(N, P, C, device="cuda", generator=g), dt)
    gamma, beta = 1 + 0.1 * torch.randn(C, device="cuda", generator=g), 0.1 * torch.randn(C, device="cuda", generator=g)
    st = torch.stack([x.sum(1), (x * x).sum(1)], dim=-1).contiguous()
    xa = Act(x.to(dt).reshape(-1), N, 1, P, C, stats=st)
    seed = torch.tensor([123456789], device="cuda", dtype=torch.int64)
    plan = _plan(dt)
    y = plan.groupnorm([xa], gamma, beta, silu=True, drop_seed=seed, drop_p=p, drop_site=3)
    y0 = plan.groupnorm([xa], gamma, beta, silu=True)
    plan.run()
    torch.cuda.synchronize()
    yv, y0v = y.t.float().reshape(N, P, C), y0.t.float().reshape(N, P, C)
    dropped = (yv == 0) & (y0v != 0)
    frac = dropped.float().mean().item()
    assert abs(frac - p) < 0.01, frac
    keep = 1 / (1 - p)
    assert rel_l2(yv[~dropped], y0v[~dropped] * keep) < 5e-3          # kept elements: scaled by 1 / (1 - p)
    dx, _, dgam, dbet = groupnorm_silu_backward(xa, Act(dy.to(dt).reshape(-1), N, 1, P, C), gamma, beta, drop_seed=seed,
                                                drop_p=p, drop_site=3)
    torch.cuda.synchronize()
    mask = (~dropped).double().permute(0, 2, 1) * keep
    xr = x.double().permute(0, 2, 1).clone().requires_grad_(True)
    gr, br = gamma.double().clone().requires_grad_(True), beta.double().clone().requires_grad_(True)
    (F.silu(F.group_norm(xr, 32, gr, br, eps=1e-5)) * mask).backward(dy.double().permute(0, 2, 1))
    assert rel_l2(dx.t.float().reshape(N, P, C), xr.grad.permute(0, 2, 1)) < 5e-3
    assert rel_l2(dgam, gr.grad) < 2e-3 and rel_l2(dbet, br.grad) < 2e-3
    # another site or seed draws other decisions
    plan2 = _plan(dt)
    y2 = plan2.groupnorm([xa], gamma, beta, silu=True, drop_seed=seed, drop_p=p, drop_site=4)
    plan2.run()
    torch.cuda.synchronize()
    assert not torch.equal(y2.t, y.t)


@pytest.mark.parametrize("N,H,W,cin,cout,k", [(4, 32, 32, 128, 128, 3), (3, 4, 4, 512, 256, 3), (2, 16, 16, 256, 128, 3),
                                              (1, 128, 128, 64, 64, 3), (5, 8, 8, 192, 64, 1), (2, 24, 40, 64, 128, 3)])
def test_conv2d_wgrad_matches_autograd(N, H, W, cin, cout, k):
    """tq_conv2d_wgrad (tcgen05, K = boxes of 64 positions, one shifted TMA box per tap, tap groups over the grid) against
    autograd's weight gradient of F.conv2d(padding='same'); 4 x 4 images pack several samples into a K-step (3 samples
    leave the box ragged), 24 x 40 is ragged in both image dimensions."""
    from tqdne_b200.backward import conv2d_weight_grad

    g = torch.Generator(device="cuda").manual_seed(H + W + k)
    x = torch.randn(N, H, W, cin, device="cuda", generator=g).to(torch.bfloat16)
    dy = torch.randn(N, H, W, cout, device="cuda", generator=g).to(torch.bfloat16)
    dw = conv2d_weight_grad(x, dy, k, k)
    torch.cuda.synchronize()
    w = torch.zeros(cout, cin, k, k, device="cuda", dtype=torch.float64, requires_grad=True)
    F.conv2d(x.double().permute(0, 3, 1, 2), w, padding=k // 2).backward(dy.double().permute(0, 3, 1, 2))
    assert rel_l2(dw.view(cout, k, k, cin).permute(0, 3, 1, 2), w.grad) < 2e-5


def test_conv_random_shape_sweep_bf16():
    """Seeded sweep over shapes nobody hand-picked: odd widths / heights / batches (ragged tiles in every direction),
    1-D lengths that are not multiples of the tile, every channel count of the networks, k in {1, 3, 5}, with the whole
    bf16 epilogue (bias, residual, fused GroupNorm statistics) -- output against torch, statistics against the stored
    output, and the GroupNorm that consumes them against torch."""
    from tqdne_b200.engine import pack_conv

    rng = np.random.default_rng(2024)
    g = torch.Generator(device="cuda").manual_seed(2024)
    cases = []
    for _ in range(14):
        N = int(rng.integers(1, 10))
        if rng.random() < 0.5:
            sp = (int(rng.integers(2, 41)), int(rng.integers(2, 41)))
            k = int(rng.choice([1, 3]))
        else:
            sp = (int(rng.integers(8, 700)),)
            k = int(rng.choice([1, 3, 5]))
        cin, cout = int(rng.choice([64, 128, 192, 256])), int(rng.choice([64, 128, 256, 320]))
        cases.append((N, sp, cin, cout, k, bool(rng.random() < 0.5)))
    for N, sp, cin, cout, k, res in cases:
        dims = len(sp)
        x = torch.randn(N, cin, *sp, device="cuda", generator=g)
        w = torch.randn(cout, cin, *([k] * dims), device="cuda", generator=g) / math.sqrt(cin * k**dims)
        b = torch.randn(cout, device="cuda", generator=g) * 0.3
        plan = _plan(torch.bfloat16)
        xa = _act(x, torch.bfloat16)
        use_res = res and cin == cout
        y = plan.conv(pack_conv(w, b, [cin], torch.bfloat16), [xa], residual=xa if use_res else None, dims=dims, stats=True)
        gamma = 1 + 0.1 * torch.randn(cout, device="cuda", generator=g)
        beta = 0.1 * torch.randn(cout, device="cuda", generator=g)
        o = plan.groupnorm([y], gamma, beta, silu=True)
        plan.run()
        torch.cuda.synchronize()
        xr = _rt(x, torch.bfloat16)
        ref = _ref_conv(xr, _rt(w, torch.bfloat16), b) + (xr if use_res else 0)
        yn = _to_nchw(y, dims)
        tag = f"N={N} sp={sp} {cin}->{cout} k{k} res={use_res} [{plan.op_names()[0]}]"
        assert rel_l2(yn, ref) < 5e-3, tag
        flat = yn.reshape(N, cout, -1).double()
        st = y.stats.reshape(N, y.stats_parts, cout, 2).double().sum(1)
        assert rel_l2(st[..., 1], (flat * flat).sum(-1)) < 1e-4, tag
        assert float((st[..., 0] - flat.sum(-1)).abs().max()) < 1e-3 * float(flat.abs().sum(-1).max()), tag
        assert rel_l2(_to_nchw(o, dims), F.silu(F.group_norm(yn, 32, gamma, beta, eps=1e-5))) < 4e-3, tag
