"""Which bf16 roundings dominate the per-call error of the latent UNet?  (CPU emulation on the oracle.)"""
import sys, itertools
sys.path.insert(0, str(__import__('pathlib').Path(__file__).resolve().parents[2]))
import torch, torch.nn.functional as F
from oracle import torch_ref as R
from oracle.weights import seeded_state_dict, shapes_of
from tests.helpers import golden, unet_cfg
import tqdne_b200 as tq

FLAGS = dict(w=True, act_in=True, branch=True, stream=True, gn_out=True)
def r(x, on=True): return x.bfloat16().float() if on else x
orig_conv, orig_gn = R._conv, R._gn

def conv(x, w, b, stride=1, round_out=True):
    y = orig_conv(r(x, FLAGS['act_in']), r(w, FLAGS['w']), b, stride)
    return y
def res_block(sd, p, x, emb):
    h = r(F.silu(orig_gn(x, sd[p+"in_layers.0.weight"], sd[p+"in_layers.0.bias"])), FLAGS['gn_out'])
    h = conv(h, sd[p+"in_layers.2.weight"], sd[p+"in_layers.2.bias"])
    if emb is not None:
        e = F.linear(r(F.silu(emb), True), r(sd[p+"emb_layers.1.weight"], FLAGS['w']), sd[p+"emb_layers.1.bias"])
        h = h + e[(...,)+(None,)*(h.dim()-2)]
    h = r(h, FLAGS['branch'])
    h = r(F.silu(orig_gn(h, sd[p+"out_layers.0.weight"], sd[p+"out_layers.0.bias"])), FLAGS['gn_out'])
    h = conv(h, sd[p+"out_layers.3.weight"], sd[p+"out_layers.3.bias"])
    if p+"skip_connection.weight" in sd:
        x = conv(x, sd[p+"skip_connection.weight"], sd[p+"skip_connection.bias"])
    return r(x + h, FLAGS['stream'])
def attention(sd, p, x, heads):
    import math
    b, c = x.shape[:2]; spatial = x.shape[2:]
    g = r(orig_gn(x, sd[p+"norm.weight"], sd[p+"norm.bias"]), FLAGS['gn_out'])
    qkv = r(conv(g, sd[p+"qkv.weight"], sd[p+"qkv.bias"]), FLAGS['branch']).reshape(b, 3*c, -1)
    t = qkv.shape[-1]; d = c//heads
    q, k, v = qkv.chunk(3, dim=1); s = 1/math.sqrt(math.sqrt(d))
    q = (q*s).reshape(b*heads, d, t); k = (k*s).reshape(b*heads, d, t)
    w = torch.softmax(torch.einsum("bct,bcs->bts", q, k).float(), dim=-1)
    a = r(torch.einsum("bts,bcs->bct", r(w, FLAGS['branch']), v.reshape(b*heads, d, t)).reshape(b, c, *spatial), FLAGS['branch'])
    return r(x + conv(a, sd[p+"proj_out.weight"], sd[p+"proj_out.bias"]), FLAGS['stream'])
def upsample(sd, p, x):
    x = F.interpolate(x, scale_factor=2, mode="nearest")
    return r(conv(x, sd[p+"conv.weight"], sd[p+"conv.bias"]), FLAGS['stream'])

def emulated(sd, cfg, x, t, cond):
    R._res_block, R._attention, R._upsample = res_block, attention, upsample
    oc = R._conv
    def top_conv(x, w, b, stride=1):
        return r(conv(x, w, b, stride), FLAGS['stream'])
    R._conv = top_conv
    try:
        with torch.no_grad():
            return R.unet_forward(sd, cfg, x, t, cond)
    finally:
        R._conv = oc
        import importlib; importlib.reload(R)

def rel(a, b): return float((a.double()-b.double()).norm()/b.double().norm())
for name, kind in (("unet_latent2d","latent2d"),("unet_1d","1d")):
    g = golden(name)
    cfg = unet_cfg(kind)
    net = tq.UNetModel(**cfg)
    sd = seeded_state_dict(shapes_of(net), g["seed"])
    print(name)
    for desc, fl in (("all roundings (engine)", {}), ("fp32 stream", dict(stream=False)), ("fp32 branch+gn_out", dict(branch=False, gn_out=False)),
                     ("fp32 weights", dict(w=False)), ("only stream rounding", dict(w=False, act_in=False, branch=False, gn_out=False)),
                     ("only weights", dict(act_in=False, branch=False, gn_out=False, stream=False)),
                     ("none", dict(w=False, act_in=False, branch=False, gn_out=False, stream=False))):
        FLAGS.update(dict(w=True, act_in=True, branch=True, stream=True, gn_out=True)); FLAGS.update(fl)
        y = emulated(sd, cfg, g["x"], g["t"], g["cond"])
        print("  %-28s rel-L2 vs reference golden = %.3e" % (desc, rel(y, g["y"])))
