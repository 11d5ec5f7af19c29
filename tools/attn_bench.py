"""Attention core alone: tcgen05 kernel (tq_attn_sm100.cu) against the FFMA kernel (TQ_ATTN_SIMT=1) on the shapes of
the 1D UNet (T = 508, 4 heads x 64) and of the pixel-space 2D UNet (T = 256, 4 heads x 128).

    python tools/attn_bench.py [N]
"""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from tqdne_b200.engine import Act, Plan  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
s = torch.cuda.Stream()


def build(T, heads, d, simt):
    os.environ["TQ_ATTN_SIMT"] = "1" if simt else "0"
    C = heads * d
    g = torch.Generator(device="cuda").manual_seed(T)
    qkv = torch.randn(N, T, 3 * C, device="cuda", generator=g).to(torch.bfloat16)
    plan = Plan(torch.device("cuda"), torch.bfloat16)
    out = plan.attention(Act(qkv.reshape(-1), N, 1, T, 3 * C), heads)
    return plan, out, qkv


def timed(plan, iters=20):
    with torch.cuda.stream(s):
        for _ in range(3):
            plan.run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(iters):
            plan.run()
        e1.record(s)
        s.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


for T, heads, d in [(508, 4, 64), (256, 4, 128), (16, 4, 128)]:
    flops = 4.0 * N * heads * T * T * d
    res = {}
    for simt in (True, False):
        plan, out, _ = build(T, heads, d, simt)
        us = timed(plan)
        res[simt] = (us, out.t.float().clone(), plan.op_names()[-1])
    a, b = res[True][1], res[False][1]
    diff = float((a - b).norm() / a.norm())
    print(f"N={N} T={T} heads={heads} d={d}: {res[True][2]} {res[True][0]:8.1f} us ({flops / res[True][0] / 1e6:7.1f} TF/s) | "
          f"{res[False][2]} {res[False][0]:8.1f} us ({flops / res[False][0] / 1e6:7.1f} TF/s) | rel diff {diff:.2e}", flush=True)
