"""CPU: host-side lowering logic with a recording fake of the C library (no compute)."""
import pytest
import torch

from tests.fake_lib import FakeLib
from tests.helpers import CFG, arch, unet_cfg


@pytest.fixture
def fake(monkeypatch):
    from tqdne_b200 import _lib, engine, lowering

    f = FakeLib()
    monkeypatch.setattr(_lib, "_lib", f)
    monkeypatch.setattr(engine, "require_cuda", lambda *a, **k: None)
    monkeypatch.setattr(lowering, "require_cuda", lambda *a, **k: None)
    return f


def _count(fake, name):
    return sum(1 for c in fake.calls if c == name)


def test_latent_unet_lowering_op_counts_and_flops(fake):
    import tqdne_b200 as tq
    from tqdne_b200.lowering import UNetPlan

    net = tq.UNetModel(**unet_cfg("latent2d")).eval()
    p = UNetPlan(net, 4, (32, 32), torch.bfloat16, uniform_t=True)
    # 78 convs per call; 14 1x1 shortcuts are fused into their ResBlock's second conv, +1 emb GEMM
    assert _count(fake, "tq_plan_add_conv") == 78 - 14 + 1
    assert _count(fake, "tq_plan_add_groupnorm") == 51
    assert _count(fake, "tq_plan_add_attention") == 6
    # 16.979 GFLOP per sample per call
    assert abs(p.flops / 4 / 16.979e9 - 1) < 0.01
    assert p.out.C == 8 and p.out.t.dtype == torch.float32 and p.cin_pad == 64
    # every GroupNorm input is a conv output: its statistics come from that conv's epilogue (no stats pass), each from
    # its own buffer of per-tile partial sums (plain stores, nothing to clear: no memset op, no atomics)
    produced = {c["stats"] for c in fake.convs if c["stats"]}
    assert len(produced) == sum(1 for c in fake.convs if c["stats"])
    assert all(c["stats_parts"] == 3 for c in fake.convs if c["stats"])   # what the (fake) tq_conv_stats_parts said
    for g in fake.gns:
        assert g["stats0"] in produced and (g["C1"] == 0 or g["stats1"] in produced) and not g["ws"]
        assert g["parts0"] == 3 and (g["C1"] == 0 or g["parts1"] == 3)
    assert _count(fake, "tq_plan_add_memset") == 0
    for c in fake.convs:
        assert c["ktot"] % 64 == 0 and c["cout_pad"] % 64 == 0
        kbs = sorted(s[4] for s in c["slices"][: c["num_slices"]])
        assert kbs == list(range(c["ktot"] // 64)), "every weight K block is used exactly once per class"
        for s in c["slices"]:
            assert 0 <= s[0] < c["num_srcs"] and s[3] % 64 == 0 and s[3] + 64 <= c["srcs"][s[0]][3]


def test_stride2_and_upsample_slice_tables(fake):
    from tqdne_b200.engine import Act, Plan, pack_conv

    plan = Plan(torch.device("cpu"), torch.bfloat16)
    x = Act(torch.zeros(2 * 16 * 16 * 64, dtype=torch.bfloat16), 2, 16, 16, 64)
    w = torch.randn(64, 64, 3, 3)
    pc = pack_conv(w, None, [64], torch.bfloat16)
    out = plan.conv(pc, [x], stride=2)
    c = fake.convs[-1]
    assert (out.H, out.W) == (8, 8) and c["num_srcs"] == 4 and c["num_slices"] == 9
    # tap (ky,kx) reads in(2y+ky-1, 2x+kx-1) = view[(ky-1)%2,(kx-1)%2](y+dy, x+dx)
    for (src, dx, dy, c0, kb), (ky, kx) in zip(c["slices"], pc.taps):
        assert dy == (-1 if ky == 0 else 0) and dx == (-1 if kx == 0 else 0)
    assert all(s[1:4] == (8, 8, 64) and s[5] == 2 * 16 * 64 and s[6] == 128 for s in c["srcs"])
    up = plan.conv(pc, [x], upsample=True)
    c = fake.convs[-1]
    assert (up.H, up.W) == (32, 32) and c["num_classes"] == 4 and c["num_slices"] == 9
    assert c["class_off"] == [0, 64, 32 * 64, 32 * 64 + 64] and c["out_strides"] == (32 * 32 * 64, 2 * 32 * 64, 128)
    # class (py=1,px=0), tap ky=2 -> dy = (1+2-1)//2 = 1 ; tap ky=0 -> dy = 0
    cls2 = c["slices"][2 * 9:3 * 9]
    assert [s[2] for s in cls2] == [0, 0, 0, 0, 0, 0, 1, 1, 1]
    assert [s[1] for s in cls2] == [-1, 0, 0] * 3


def test_1d_unet_and_decoder_lowering(fake):
    import tqdne_b200 as tq
    from tqdne_b200.lowering import CoderPlan, UNetPlan

    net = tq.UNetModel(**unet_cfg("1d")).eval()
    p = UNetPlan(net, 2, (4064,), torch.float32, uniform_t=False)
    assert abs(p.flops / 2 / 28.436e9 - 1) < 0.01
    enc_cfg, dec_cfg = arch().get_2d_autoencoder_configs(CFG)
    dec = tq.Decoder(**dec_cfg).eval()
    n0 = len(fake.convs)
    d = CoderPlan(dec, "decoder", 2, (32, 32), torch.bfloat16)
    assert abs(d.flops / 2 / 27.206e9 - 1) < 0.01
    assert len(fake.convs) - n0 == 18 - 2                # 18 convs, two 1x1 shortcuts fused
    assert (d.out.H, d.out.W, d.out.C) == (128, 128, 3)


def test_option_lowering(fake):
    """The module options no shipped config uses, as the lowering hands them to the library: FiLM -> tq_gn_desc.film on the
    second norm of every ResBlock (and no embedding add on conv1), use_causal_mask -> tq_attn_desc.causal, conv_resample=False
    -> resample ops instead of convolutions (their consumers' norms run a statistics pass: no conv epilogue behind them),
    cond_sample -> a second source of the stem convolution, cond_emb_scale -> one more Fourier op."""
    import tqdne_b200 as tq
    from tqdne_b200.lowering import UNetPlan

    def lower(kind, **kw):
        fake.calls.clear(); fake.convs.clear(); fake.gns.clear(); fake.attns.clear(); fake.resamples.clear()
        net = tq.UNetModel(**unet_cfg(kind)).eval()
        return net, UNetPlan(net, 2, (32, 32) if "2d" in kind else (512,), torch.bfloat16, uniform_t=True, **kw)

    net, _ = lower("latent2d")
    base_convs, base_gns = len(fake.convs), len(fake.gns)
    assert not any(g["film"] for g in fake.gns) and not any(a["causal"] for a in fake.attns) and not fake.resamples
    n_res = sum(1 for m in net.modules() if type(m).__name__ == "ResBlock")

    lower("latent2d_film")
    assert sum(1 for g in fake.gns if g["film"]) == n_res and len(fake.gns) == base_gns
    assert all(g["film_ld"] >= 2 * (g["C0"] + g["C1"]) for g in fake.gns if g["film"])
    assert sum(1 for c in fake.convs if c["has_emb"]) == 0            # FiLM: conv1 no longer adds the embedding

    lower("latent2d_causal")
    assert len(fake.attns) == 6 and all(a["causal"] == 1 for a in fake.attns)

    lower("latent2d_pool")
    assert len(fake.convs) == base_convs - 6                          # 3 stride-2 and 3 post-upsample convolutions are gone
    assert sorted(r["mode"] for r in fake.resamples) == [0, 0, 0, 1, 1, 1]
    assert sum(1 for g in fake.gns if not g["stats0"]) >= 6           # norms behind a resampler compute their own statistics

    _, p = lower("latent2d_condsample", cond_channels=8)
    stem = fake.convs[1]                                               # convs[0] = the embedding GEMM
    assert stem["num_srcs"] == 2 and p.xcond is not None and p.cin_pad == 64 and p.xcond.C == 64
    assert len(fake.convs) == base_convs

    lower("1d_condembed")
    assert sum(1 for c in fake.calls if c == "tq_plan_add_fourier") == 2   # the noise level and the conditioning feature
    with pytest.raises(NotImplementedError):
        tq.UNetModel(**(unet_cfg("1d") | {"cond_emb_scale": 0.5}))     # five features: the reference cannot run it either


def test_pack_conv_layout():
    from tqdne_b200.engine import pack_conv

    w = torch.arange(2 * 3 * 3 * 3, dtype=torch.float32).reshape(2, 3, 3, 3)
    pc = pack_conv(w, torch.tensor([1.0, 2.0]), [2, 1], torch.float32)
    assert pc.weights.shape == (64, 9 * 128) and pc.cout == 2 and pc.seg_pad == [64, 64]
    # tap (ky=1,kx=2) = index 5; segment 1 (input channel 2) starts at column 5*128 + 64
    assert float(pc.weights[1, 5 * 128 + 64]) == float(w[1, 2, 1, 2])
    assert float(pc.weights[0, 5 * 128 + 1]) == float(w[0, 1, 1, 2])
    assert pc.kb_of[(5, 1, 0)] == (5 * 128 + 64) // 64 and float(pc.bias[1]) == 2.0 and float(pc.weights[2:].abs().sum()) == 0
    ws = torch.ones(2, 3, 1, 1)
    pc2 = pack_conv(w[:, :2], None, [2], torch.float32, shortcut=(ws, torch.tensor([0.5, 0.5]), [2, 1]))
    assert pc2.ktot == 9 * 64 + 128 and pc2.extra_kb == {(0, 0): 9, (1, 0): 10} and float(pc2.bias[0]) == 0.5
    assert float(pc2.weights[0, 9 * 64 + 64]) == 1.0 and float(pc2.weights[0, 9 * 64 + 2]) == 0.0


def test_training_parameter_store_layout_round_trip():
    """tqdne_b200.training._Store (host logic, no kernels): every trainable parameter of the 1D UNet gets an engine-layout
    view (conv weights [cout_pad, taps, cin_pad]); module -> store -> module is the identity, the emb_layers rows are
    adjacent (one dense layer), padded stem / head channels are zero."""
    import torch

    import tqdne_b200 as tq
    from tests.helpers import seeded, unet_cfg
    from tqdne_b200.training import _Store

    net = seeded(tq.UNetModel(**unet_cfg("1d")), 3)
    st = _Store(net, "cpu")
    st.load_from_module()
    trainable = [p for p in net.parameters() if p.requires_grad]
    assert len(st.order) == len(trainable) and st.n >= sum(p.numel() for p in trainable)
    for p in trainable:
        assert torch.equal(st.to_module_layout(st.P, p), p.detach())
    stem = net.input_blocks[0][0]
    v = st.view(st.P, stem.weight)
    assert v.shape == (64, 5, 64) and float(v[:, :, 6:].abs().max()) == 0.0 and torch.equal(v[:, 2, :6], stem.weight[:, :, 2])
    head = net.out[2]
    assert st.view(st.P, head.weight).shape == (64, 5, 64) and float(st.view(st.P, head.weight)[6:].abs().max()) == 0.0
    res = [m for m in net.modules() if type(m).__name__ == "ResBlock"]
    rows = sum(r.emb_layers[1].weight.shape[0] for r in res)
    wall = st.P[st.emb_w_off:st.emb_w_off + rows * 256].view(rows, 256)
    assert st.emb_rows == rows and torch.equal(wall[:64], res[0].emb_layers[1].weight.detach())
    assert torch.equal(wall[-res[-1].emb_layers[1].weight.shape[0]:], res[-1].emb_layers[1].weight.detach())
    assert torch.equal(st.EMA, st.P)
