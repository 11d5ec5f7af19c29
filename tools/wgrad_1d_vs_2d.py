"""A/B of the two weight-gradient kernels on the 1D UNet shapes: tq_conv1d_wgrad (one X halo buffer, taps by descriptor row
offset) against tq_conv2d_wgrad run as a 1 x 5 kernel (one shifted TMA box per tap).  Measured equal within 2 % (763 TF/s on
256->256 k5 @1016 x 64 without the bias kernel): the weight gradient is not bound by the operand loads.
"""
import sys, math
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from tqdne_b200 import backward as bw
dev = torch.device("cuda"); dt = torch.bfloat16
s = torch.cuda.Stream()
def timed(fn, iters=20):
    with torch.cuda.stream(s):
        for _ in range(3): fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(iters): fn()
        e1.record(s); s.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3
N = 64
for cin, cout, L in [(256, 256, 1016), (256, 256, 508), (512, 256, 1016), (128, 128, 2032), (64, 64, 4064)]:
    x = torch.randn(N, L, cin, device=dev).to(dt); dy = torch.randn(N, L, cout, device=dev).to(dt)
    dw1 = torch.zeros(cout, 5, cin, device=dev); db = torch.zeros(cout, device=dev); dw2 = torch.zeros(cout, 5, cin, device=dev)
    fl = 2.0 * N * L * cin * cout * 5
    t1 = timed(lambda: bw.conv1d_weight_grad(x, dy, 5, dw1, db, bias=False))
    t2 = timed(lambda: bw.conv2d_weight_grad(x.view(N, 1, L, cin), dy.view(N, 1, L, cout), 1, 5, dw2))
    dw1.zero_(); dw2.zero_()
    bw.conv1d_weight_grad(x, dy, 5, dw1, db, bias=False); bw.conv2d_weight_grad(x.view(N, 1, L, cin), dy.view(N, 1, L, cout), 1, 5, dw2)
    torch.cuda.synchronize()
    print(f"{cin}->{cout} L={L}: halo kernel {t1*1e6:7.1f} us {fl/t1/1e12:6.1f} TF/s | per-tap-box kernel {t2*1e6:7.1f} us {fl/t2/1e12:6.1f} TF/s | diff {float((dw1-dw2).norm()/dw1.norm()):.1e}", flush=True)
