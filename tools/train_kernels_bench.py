"""Throughput of the training-step kernels built so far (SURVEY 8(f) rank 1 groundwork) on the 1D UNet's top conv shapes
at the per-GPU batch of BASELINE.json configs[4] (64 x [C, L]): forward igemm, input gradient (the same igemm over dY),
weight gradient (tq_conv1d_wgrad) in TFLOP/s, GroupNorm+SiLU forward / backward in GB/s.  CUDA-event times, 20 runs.

    python tools/train_kernels_bench.py [N]
"""
import math
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from tqdne_b200 import backward as bw  # noqa: E402
from tqdne_b200.engine import Act, Plan, pack_conv  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dt = torch.bfloat16
dev = torch.device("cuda")
s = torch.cuda.Stream()


def timed(fn, iters=20):
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(iters):
            fn()
        e1.record(s)
        s.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


for cin, cout, L in [(256, 256, 1016), (256, 256, 508), (512, 256, 1016), (128, 128, 2032), (64, 64, 4064)]:
    k = 5
    x = torch.randn(N, L, cin, device=dev).to(dt)
    dy = torch.randn(N, L, cout, device=dev).to(dt)
    w = (torch.randn(cout, cin, k, device=dev) / math.sqrt(cin * k)).to(dt).float()
    flops = 2.0 * N * L * cin * cout * k
    with torch.cuda.stream(s):
        pf = Plan(dev, dt)
        xa = Act(x.reshape(-1), N, 1, L, cin)
        y = pf.conv(pack_conv(w, None, [cin], dt), [xa], dims=1, stats=True)
        pb = Plan(dev, dt)
        bw.conv1d_input_grad(pb, w, Act(dy.reshape(-1), N, 1, L, cout))
        dw = torch.zeros(cout, k, cin, device=dev)
        db = torch.zeros(cout, device=dev)
        gam, bet = torch.ones(cout, device=dev), torch.zeros(cout, device=dev)
        pg = Plan(dev, dt)
        pg.groupnorm([y], gam, bet, silu=True)
        pf.run()
    t_f = timed(pf.run)
    t_d = timed(pb.run)
    t_w = timed(lambda: bw.conv1d_weight_grad(x, dy, k, dw, db))
    t_g = timed(pg.run)
    dya = Act(dy.reshape(-1), N, 1, L, cout)
    t_gb = timed(lambda: bw.groupnorm_silu_backward(y, dya, gam, bet))
    el = N * L * cout
    print(f"conv1d {cin:4d}->{cout:4d} k5 L={L:5d} N={N}: fwd {t_f * 1e6:7.1f} us {flops / t_f / 1e12:7.1f} TF/s | dgrad {t_d * 1e6:7.1f} us "
          f"{flops / t_d / 1e12:7.1f} TF/s | wgrad(+bias) {t_w * 1e6:7.1f} us {flops / t_w / 1e12:7.1f} TF/s | GN+SiLU fwd "
          f"{t_g * 1e6:6.1f} us {4 * el / t_g / 1e9:6.0f} GB/s | bwd {t_gb * 1e6:6.1f} us {10 * el / t_gb / 1e9:6.0f} GB/s", flush=True)

# ---- 2-D weight gradient on the latent UNet's top shapes (batch 256), against the forward igemm of the same layer
for cin, cout, H in [(128, 128, 32), (256, 256, 16), (512, 512, 8), (512, 512, 4)]:
    Nb = 256
    x = torch.randn(Nb, H, H, cin, device=dev).to(dt)
    dy = torch.randn(Nb, H, H, cout, device=dev).to(dt)
    w = (torch.randn(cout, cin, 3, 3, device=dev) / math.sqrt(cin * 9)).to(dt).float()
    flops = 2.0 * Nb * H * H * cin * cout * 9
    with torch.cuda.stream(s):
        pf = Plan(dev, dt)
        pf.conv(pack_conv(w, None, [cin], dt), [Act(x.reshape(-1), Nb, H, H, cin)], dims=2, stats=True)
        dw = torch.zeros(cout, 9, cin, device=dev)
        pf.run()
    t_f = timed(pf.run)
    t_w = timed(lambda: bw.conv2d_weight_grad(x, dy, 3, 3, dw))
    print(f"conv2d {cin:4d}->{cout:4d} 3x3 @{H}x{H} N={Nb}: fwd {t_f * 1e6:7.1f} us {flops / t_f / 1e12:7.1f} TF/s | wgrad {t_w * 1e6:7.1f} us "
          f"{flops / t_w / 1e12:7.1f} TF/s", flush=True)
