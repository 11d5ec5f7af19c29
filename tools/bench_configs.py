"""Secondary measurements for the BASELINE.json configs that are NOT the bench.py headline (they are parity-test
cases there): waveforms/s of the 1D EDM UNet pipeline (configs[0] shape: [B, 6, 4064], 18 Heun steps, moving-average
envelope inverse) and of the pixel-space 2D EDM UNet pipeline (configs[3]: [B, 3, 128, 128], 32 Heun steps,
Griffin-Lim), on one B200, bf16, inputs resident, CUDA-event time.  One JSON line per config plus (with `ops`) the
per-op breakdown of one denoiser call.

    python tools/bench_configs.py 1d 64 [ops]
    python tools/bench_configs.py pixel 128 [ops]
"""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

import tqdne_b200 as tq  # noqa: E402
from bench import cond_grid  # noqa: E402
from tqdne_b200.synthetic_weights import seeded_state_dict, shapes_of  # noqa: E402
from tqdne_b200.config import MovingAverageEnvelopeConfig, SpectrogramConfig  # noqa: E402
from tqdne_b200.lowering import get_unet_plan  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "1d"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
ops = len(sys.argv) > 3 and sys.argv[3] == "ops"

if what == "1d":
    cfg = MovingAverageEnvelopeConfig()
    unet_cfg = tq.get_1d_unet_config(cfg, 6, 6)
    steps, shape, spatial = 18, (B, 6, cfg.t), (cfg.t,)
    flop_per_wave = 35 * 28.436e9
    name = "1D EDM UNet (train_1d_edm config): 18 Heun steps (35 NFE) + moving-average-envelope inverse"
else:
    cfg = SpectrogramConfig()
    unet_cfg = tq.get_2d_unet_config(cfg, 3, 3)
    steps, shape, spatial = 32, (B, 3, 128, 128), (128, 128)
    flop_per_wave = 63 * 271.89e9
    name = "pixel-space 2D log-spectrogram EDM UNet (train_edm config): 32 Heun steps (63 NFE) + Griffin-Lim"
edm = tq.LightningEDM(unet_cfg, {}, num_sampling_steps=steps)
sd = seeded_state_dict(shapes_of(edm), 0)
for k in ("unet.out.2.weight", "unet.out.2.bias"):
    sd[k] = sd[k] * 0.05   # keep the random-init output inside the representation's range
edm.load_state_dict(sd)
edm.eval().cuda().set_engine_precision("bf16")
cond = torch.from_numpy(cond_grid(B)).cuda()
noise = torch.randn(shape, device="cuda", dtype=torch.float64, generator=torch.Generator(device="cuda").manual_seed(1))
rep_inv = cfg.representation


def step():
    rep = edm.sample(shape, cond=cond, noise=noise)
    if what == "1d":
        return rep_inv.invert_representation_device(rep) if hasattr(rep_inv, "invert_representation_device") else rep
    return rep_inv.invert_representation_device(torch.tanh(rep))


for _ in range(2):
    w = step()
torch.cuda.synchronize()
assert bool(torch.isfinite(w).all())
K = 3
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(K):
    w = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
print(json.dumps({"config": name, "batch": B, "ms_per_step": ms, "waveforms_per_s": B / ms * 1e3,
                  "achieved_tflops": B / ms * 1e3 * flop_per_wave / 1e12, "output": list(w.shape), "dtype": "bf16",
                  "micro_batch": max(1, min(B, edm.max_positions_per_pass // (spatial[0] if len(spatial) == 1 else 128 * 128)))}),
      flush=True)

if ops:
    mb = max(1, min(B, edm.max_positions_per_pass // (spatial[0] if len(spatial) == 1 else 128 * 128)))
    p = get_unet_plan(edm.unet, mb, spatial, uniform_t=True).plan
    n, names, meta = p.num_ops, p.op_names(), p.op_meta
    s = torch.cuda.Stream()
    iters = 3
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(n + 1)] for _ in range(iters)]
    with torch.cuda.stream(s):
        p.run_range(0, n)
        for it in range(iters):
            ev[it][0].record(s)
            for i in range(n):
                p.run_range(i, i + 1)
                ev[it][i + 1].record(s)
    s.synchronize()
    tot, agg = 0.0, {}
    for i in range(n):
        t = min(ev[it][i].elapsed_time(ev[it][i + 1]) for it in range(iters))
        tot += t
        kind, fl, by = meta[i]
        rate = f"{fl / t / 1e9:8.1f} TF/s" if fl else (f"{by / t / 1e6:8.1f} GB/s" if by else " " * 13)
        print(f"{i:4d} {t * 1e3:9.1f} us {rate}  {names[i]}")
        a = agg.setdefault(names[i].split(" ")[0], [0.0, 0, 0, 0])
        a[0] += t; a[1] += 1; a[2] += fl; a[3] += by
    print(f"# micro-batch {mb}: total {tot:.3f} ms over {n} ops")
    for k, (t, cnt, fl, by) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        rate = f"{fl / t / 1e9:8.1f} TF/s" if fl else (f"{by / t / 1e6:8.1f} GB/s" if by else "")
        print(f"#   {t:8.3f} ms {100 * t / tot:5.1f}% n={cnt:3d} {rate}  {k}")
