"""Per-op CUDA-event timing of one kernel plan (eager replay, op by op) + phase timing of one bench step.

    python tools/op_breakdown.py [B] [unet|decoder|phases]

Prints one line per op: index, ms, TFLOP/s or GB/s (from the plan's algorithmic op_meta), kernel name.
Not a bench value: eager per-op events include launch gaps; use it to rank ops, not to quote throughput.
"""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

import tqdne_b200 as tq  # noqa: E402
from bench import build_state_dict, cond_grid  # noqa: E402
from tqdne_b200.config import LatentSpectrogramConfig  # noqa: E402
from tqdne_b200.lowering import get_coder_plan, get_unet_plan  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
what = sys.argv[2] if len(sys.argv) > 2 else "unet"
cfg = LatentSpectrogramConfig()
enc_cfg, dec_cfg = tq.get_2d_autoencoder_configs(cfg)
edm = tq.LightningEDM(tq.get_2d_unet_config(cfg, 8, 8), {}, autoencoder=tq.LightningAutoencoder(enc_cfg, dec_cfg, {}))
edm.load_state_dict(build_state_dict(edm))
edm.eval().cuda().set_engine_precision("bf16")


def per_op(p, iters=5):
    n = p.num_ops
    names = p.op_names()
    meta = p.op_meta
    s = torch.cuda.Stream()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(n + 1)] for _ in range(iters)]
    with torch.cuda.stream(s):
        for _ in range(2):
            p.run_range(0, n)
        for it in range(iters):
            ev[it][0].record(s)
            for i in range(n):
                p.run_range(i, i + 1)
                ev[it][i + 1].record(s)
    s.synchronize()
    tot = 0.0
    agg = {}
    for i in range(n):
        ms = min(ev[it][i].elapsed_time(ev[it][i + 1]) for it in range(iters))
        tot += ms
        kind, fl, by = meta[i]
        rate = f"{fl / ms / 1e9:8.1f} TF/s" if fl else (f"{by / ms / 1e6:8.1f} GB/s" if by else " " * 13)
        print(f"{i:4d} {ms * 1e3:9.1f} us {rate}  {names[i]}")
        a = agg.setdefault(names[i].split(" ")[0], [0.0, 0, 0, 0])
        a[0] += ms; a[1] += 1; a[2] += fl; a[3] += by
    print(f"# total {tot:.3f} ms over {n} ops")
    for k, (ms, cnt, fl, by) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        rate = f"{fl / ms / 1e9:8.1f} TF/s" if fl else (f"{by / ms / 1e6:8.1f} GB/s" if by else "")
        print(f"#   {ms:8.3f} ms {100 * ms / tot:5.1f}% n={cnt:3d} {rate}  {k}")
    # whole plan, graph replay
    p.enable_graph(True)
    with torch.cuda.stream(s):
        for _ in range(3):
            p.run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(10):
            p.run()
        e1.record(s)
    s.synchronize()
    print(f"# graph replay: {e0.elapsed_time(e1) / 10:.3f} ms per plan run")


if what == "unet":
    p = get_unet_plan(edm.unet, B, (32, 32), uniform_t=True)
    p.xin.t.normal_()
    p.set_cond(torch.from_numpy(cond_grid(B)).cuda())
    p.t.fill_(0.3)
    per_op(p.plan)
elif what == "decoder":
    mb = min(B, edm.decode_micro_batch)
    p = get_coder_plan(edm.autoencoder.decoder, "decoder", mb, (32, 32))
    p.xin.t.normal_()
    print(f"# decoder micro-batch {mb}")
    per_op(p.plan)
else:
    cond = torch.from_numpy(cond_grid(B)).cuda()
    noise = torch.randn(B, 8, 32, 32, device="cuda", dtype=torch.float64)
    rep_inv = cfg.representation

    def sync_time(fn, n=3):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            out = fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / n * 1e3, out

    ae = edm.autoencoder
    edm.autoencoder = None
    edm.unet  # noqa: B018
    ms_s, lat = sync_time(lambda: edm.sample((B, 8, 32, 32), cond=cond, noise=noise))
    edm.autoencoder = ae
    ms_d, rep = sync_time(lambda: ae.decode(lat))
    ms_g, wav = sync_time(lambda: rep_inv.invert_representation_device(rep))
    ms_all, _ = sync_time(lambda: rep_inv.invert_representation_device(edm.sample((B, 3, 128, 128), cond=cond, noise=noise)))
    print(f"# phases B={B}: sampler {ms_s:.1f} ms, decode {ms_d:.1f} ms, griffin-lim {ms_g:.1f} ms, whole step {ms_all:.1f} ms")
