"""Micro-benchmark of single implicit-GEMM convs through the C-ABI plan (CUDA events, graph replay of 20 launches).

    python tools/conv_bench.py            # the standard sweep
Each line: shape, options, us per launch, TFLOP/s.  Used to separate mainloop from epilogue cost.
"""
import itertools
import math
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from tqdne_b200.engine import Act, Plan, pack_conv  # noqa: E402

dev = torch.device("cuda")


def bench(N, sp, cin, cout, k, *, res=False, emb=False, stats=False, block_n=0, cta_group=0, reps=20, f32_out=False):
    dims = len(sp)
    H, W = (sp if dims == 2 else (1, sp[0]))
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(N * H * W * cin, device=dev, generator=g).to(torch.bfloat16)
    w = torch.randn(cout, cin, *([k] * dims), device=dev, generator=g) / math.sqrt(cin * k**dims)
    b = torch.randn(cout, device=dev, generator=g)
    plan = Plan(dev, torch.bfloat16)
    xa = Act(x, N, H, W, cin)
    e = torch.randn(N, cout, device=dev, generator=g) if emb else None
    pc = pack_conv(w, b, [cin], torch.bfloat16)
    for _ in range(reps):
        plan.conv(pc, [xa], residual=xa if (res and cin == cout) else None, emb=e, emb_ld=cout if emb else 0, dims=dims,
                  stats=stats, block_n=block_n, cta_group=cta_group, out_dtype=torch.float32 if f32_out else None)
    name = [n for n in plan.op_names() if "igemm" in n][0]
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        plan.enable_graph(True)
        for _ in range(2):
            plan.run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(3):
            plan.run()
        e1.record(s)
    s.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (3 * reps)
    fl = 2 * N * H * W * cout * cin * k**dims
    opts = "".join(c if f else "-" for c, f in zip("res", (res, emb, stats)))
    print(f"N={N:4d} {str(sp):12s} {cin:4d}->{cout:4d} k{k} [{opts}] {us:8.1f} us {fl / us / 1e6:8.1f} TF/s  {name}", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        # one N H W cin cout k res emb stats bn cg [f32_out]   (H = 1 -> 1D)
        a = [int(v) for v in sys.argv[2:]]
        sp = (a[1], a[2]) if a[1] > 1 else (a[2],)
        bench(a[0], sp, a[3], a[4], a[5], res=bool(a[6]), emb=bool(a[7]), stats=bool(a[8]), block_n=a[9], cta_group=a[10], reps=6,
              f32_out=len(a) > 11 and bool(a[11]))
        sys.exit(0)
    shapes = [(256, (32, 32), 128, 128, 3), (64, (128, 128), 64, 64, 3), (256, (16, 16), 256, 256, 3),
              (256, (4, 4), 512, 512, 3), (256, (8, 8), 512, 512, 3)]
    for sh in shapes:
        for res, emb, stats in [(False, False, False), (True, False, False), (False, True, False), (False, False, True),
                                (True, False, True)]:
            bench(*sh, res=res, emb=emb, stats=stats)
    # tile-config sweep on the two shapes that matter most
    for bn, cg in itertools.product((64, 128, 256), (1, 2)):
        bench(256, (32, 32), 128, 128, 3, stats=True, block_n=bn, cta_group=cg)
    for bn, cg in itertools.product((128, 256), (1, 2)):
        bench(256, (4, 4), 512, 512, 3, stats=True, block_n=bn, cta_group=cg)
        bench(256, (8, 8), 512, 512, 3, stats=True, block_n=bn, cta_group=cg)
