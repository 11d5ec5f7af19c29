"""Tile-configuration sweep over the REAL convolutions of one denoiser call: every `Plan.conv` call the lowering makes is
recorded (same sources, weights, epilogue options) and replayed alone with each forced (block_n, cta_group), graph replay
of `reps` launches, CUDA events.  Prints the time of the heuristic's choice (`build_conv_sm100`) beside the best forced
one -- the evidence behind the cost model; a static table, never a run-time autotuner (tile shapes change the statistics
partition and with it the low bits, so the choice has to be a pure function of the shape).

    python tools/tile_sweep.py [latent|1d|pixel] [B]
"""
import itertools
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

import tqdne_b200 as tq  # noqa: E402
from tqdne_b200 import engine  # noqa: E402
from tqdne_b200.config import LatentSpectrogramConfig, MovingAverageEnvelopeConfig, SpectrogramConfig  # noqa: E402
from tqdne_b200.lowering import get_unet_plan  # noqa: E402
from tqdne_b200.synthetic_weights import seeded_state_dict, shapes_of  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "latent"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
if what == "latent":
    cfg = LatentSpectrogramConfig()
    unet_cfg, spatial = tq.get_2d_unet_config(cfg, 8, 8), (32, 32)
elif what == "1d":
    cfg = MovingAverageEnvelopeConfig()
    unet_cfg, spatial = tq.get_1d_unet_config(cfg, 6, 6), (cfg.t,)
else:
    cfg = SpectrogramConfig()
    unet_cfg, spatial = tq.get_2d_unet_config(cfg, 3, 3), (128, 128)
edm = tq.LightningEDM(unet_cfg, {})
edm.load_state_dict(seeded_state_dict(shapes_of(edm), 0))
edm.eval().cuda().set_engine_precision("bf16")

records = []
orig = engine.Plan.conv


def recording_conv(self, pc, srcs, **kw):
    out = orig(self, pc, srcs, **kw)
    records.append((pc, srcs, dict(kw), self.op_names()[-1]))
    return out


engine.Plan.conv = recording_conv
get_unet_plan(edm.unet, B, spatial, uniform_t=True)
engine.Plan.conv = orig
dev = torch.device("cuda")
stream = torch.cuda.Stream()


def time_conv(pc, srcs, kw, bn, cg, reps=8):
    plan = engine.Plan(dev, torch.bfloat16)
    kw = dict(kw)
    kw.pop("out", None)   # a private output per replay
    kw.update(block_n=bn, cta_group=cg)
    try:
        for _ in range(reps):
            plan.conv(pc, srcs, **kw)
    except Exception as e:  # noqa: BLE001  (a forced shape the kernel refuses)
        return None, str(e).splitlines()[0][:60]
    name = [n for n in plan.op_names() if "igemm" in n]
    if not name:
        return None, "not on the tensor path"
    with torch.cuda.stream(stream):
        plan.enable_graph(True)
        for _ in range(2):
            plan.run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(3):
            plan.run()
        e1.record(stream)
    stream.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (3 * reps), name[0]


seen = {}
tot_auto = tot_best = 0.0
for pc, srcs, kw, name in records:
    if "igemm_sm100" not in name:
        continue
    key = (tuple((s.N, s.H, s.W, s.C) for s in srcs), pc.cout, tuple(pc.kernel), kw.get("stride", 1), kw.get("upsample", False),
           kw.get("emb") is not None, kw.get("residual") is not None, kw.get("stats", False), kw.get("shortcut_srcs") is not None,
           str(kw.get("out_dtype")))
    if key in seen:
        seen[key][0] += 1
        continue
    auto_us, auto_name = time_conv(pc, srcs, kw, 0, 0)
    res = {}
    for bn, cg in itertools.product((64, 128, 256), (1, 2)):
        us, nm = time_conv(pc, srcs, kw, bn, cg)
        if us is not None:
            res[(bn, cg)] = us
    seen[key] = [1, auto_us, auto_name, res]

print(f"# {what} UNet, batch {B}: one line per distinct convolution (count, sources, cout, kernel, options)")
for key, (cnt, auto_us, auto_name, res) in seen.items():
    best = min(res, key=res.get)
    tot_auto += cnt * auto_us
    tot_best += cnt * min(auto_us, res[best])
    srcs_s = "+".join(f"{c}" for (_, _, _, c) in key[0])
    n, h, w, _ = key[0][0]
    flags = "".join(f for f, on in zip("sUerSx", (key[3] == 2, key[4], key[5], key[6], key[7], key[8])) if on)
    line = " ".join(f"{bn}/{cg}:{us:6.1f}" for (bn, cg), us in sorted(res.items()))
    mark = "" if res[best] > 0.97 * auto_us else f"   <-- {best[0]}/{best[1]} saves {cnt * (auto_us - res[best]):.1f} us"
    print(f"x{cnt:2d} {h:3d}x{w:<4d} {srcs_s:>9s}->{key[1]:4d} k{key[2][-1]} [{flags:5s}] auto {auto_us:6.1f} us ({auto_name.split(' ')[0][12:]}) | {line}{mark}")
print(f"# sum over the call: heuristic {tot_auto:.1f} us, best forced {tot_best:.1f} us")
