#!/usr/bin/env python
"""Accumulated-error probe at BASELINE.json configs[1] size: 25 Heun steps (49 NFE), batch 256.

  (1) bf16 tensor path vs the engine's own fp32 parity mode (pinned to the reference at 1e-5) on the latent, the
      decoded spectrogram and the waveforms; per-sample distribution of the rel-L2 error
  (2) run-to-run reproducibility of the bf16 path (same inputs twice)
  (3) a few rows against the CPU oracle (fp32 PyTorch restatement of the reference)
  (4) Griffin-Lim precision on the 768 decoded spectrograms: fp32 fused kernel / fp32 unfused kernel / fp64 kernel,
      against each other and (a few items) against the NumPy oracle in fp64 and fp32; kernel times

    python tests/probes/parity_probe.py [--batch 256] [--steps 25] [--oracle-rows 2] [--out gpurun_out/parity_probe.json]
"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

import bench  # noqa: E402
import tqdne_b200 as tq  # noqa: E402
from tqdne_b200 import sharding  # noqa: E402
from tqdne_b200.config import LatentSpectrogramConfig  # noqa: E402


def rel(a, b):
    a, b = a.double().flatten(1), b.double().flatten(1)
    return ((a - b).norm(dim=1) / b.norm(dim=1).clamp_min(1e-300)).cpu().numpy()


def rel_all(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm())


def dist(v):
    v = np.asarray(v)
    return {"median": float(np.median(v)), "p90": float(np.quantile(v, 0.9)), "max": float(v.max()), "mean": float(v.mean())}


def timed(fn, n=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--steps", type=int, default=25)
    ap.add_argument("--oracle-rows", type=int, default=2)
    ap.add_argument("--gl-oracle-items", type=int, default=4)
    ap.add_argument("--out", default="gpurun_out/parity_probe.json")
    ap.add_argument("--gl-only", action="store_true", help="skip the sampler sections (needs the fp32-mode run only)")
    args = ap.parse_args()
    B = args.batch
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    cfg = LatentSpectrogramConfig()
    enc_cfg, dec_cfg = tq.get_2d_autoencoder_configs(cfg)
    ucfg = tq.get_2d_unet_config(cfg, cfg.latent_channels, cfg.latent_channels)
    edm = tq.LightningEDM(ucfg, {}, num_sampling_steps=args.steps, autoencoder=tq.LightningAutoencoder(enc_cfg, dec_cfg, {}))
    sd = bench.build_state_dict(edm)
    edm.load_state_dict(sd)
    edm.eval().to(dev)
    cond = torch.from_numpy(bench.cond_grid(B)).to(dev)
    noise = sharding.global_noise((8, 32, 32), 0, B, seed=1, device="cpu").to(dev)
    res = {"batch": B, "heun_steps": args.steps, "nfe": 2 * args.steps - 1}

    def run(mode):
        edm.set_engine_precision(mode)
        sig = edm.edm.sampling_sigmas(args.steps)
        lat = edm.sample_deterministically(noise * sig[0].double().to(dev), sig, None, cond)
        rep = edm.autoencoder.decode(lat.float())
        return lat, rep

    t0 = time.time()
    lat32, rep32 = run("fp32")
    torch.cuda.synchronize()
    res["fp32_mode_seconds"] = time.time() - t0
    lat16, rep16 = run("bf16")
    lat16b, rep16b = run("bf16") if not args.gl_only else (lat16, rep16)
    res["bf16_vs_fp32"] = {"latent": dist(rel(lat16, lat32)), "latent_whole": rel_all(lat16, lat32),
                           "decoded": dist(rel(rep16, rep32)), "decoded_whole": rel_all(rep16, rep32)}
    res["bf16_run_to_run"] = {"latent_equal": bool(torch.equal(lat16, lat16b)), "latent": dist(rel(lat16b, lat16)),
                              "decoded_equal": bool(torch.equal(rep16, rep16b)), "decoded": dist(rel(rep16b, rep16))}
    lat32b, rep32b = run("fp32") if not args.gl_only else (lat32, rep32)
    res["fp32_run_to_run"] = {"latent_equal": bool(torch.equal(lat32, lat32b)), "latent": dist(rel(lat32b, lat32)),
                              "decoded": dist(rel(rep32b, rep32))}
    print(json.dumps({k: res[k] for k in ("bf16_vs_fp32", "bf16_run_to_run", "fp32_run_to_run")}), flush=True)

    # error growth along the ladder: latent after k steps, bf16 vs fp32
    growth = {}
    for k in (2, 4, 8, 16):
        if k >= args.steps or args.gl_only:
            continue
        sig = edm.edm.sampling_sigmas(args.steps)[: k + 1].clone()
        outs = {}
        for mode in ("fp32", "bf16"):
            edm.set_engine_precision(mode)
            edm.num_sampling_steps = k
            # first k rungs of the real ladder; the last rung is made the Euler-only final step by appending sigma = 0
            s = torch.cat([sig[:k], torch.zeros(1)])
            outs[mode] = edm.sample_deterministically(noise[:64] * s[0].double().to(dev), s, None, cond[:64])
        growth[k] = dist(rel(outs["bf16"], outs["fp32"]))
    edm.num_sampling_steps = args.steps
    res["latent_error_vs_steps_truncated_ladder"] = growth
    print(json.dumps({"growth": growth}), flush=True)

    # ---- (3) oracle rows
    if args.oracle_rows > 0 and not args.gl_only:
        from oracle import torch_ref

        pick = [0, B - 1, B // 2][: args.oracle_rows]
        t0 = time.time()
        sig = torch_ref.sampling_sigmas(args.steps)
        with torch.no_grad():
            nz = noise[pick].cpu()
            lat_o = torch_ref.heun_sample(sd, ucfg, nz * sig[0], sig, cond[pick].cpu())
            rep_o = torch_ref.decoder_forward(sd, dec_cfg, lat_o.float(), prefix="autoencoder.decoder.")
        res["oracle_rows"] = {"rows": pick, "seconds": time.time() - t0,
                              "fp32_latent": rel(lat32[pick].cpu(), lat_o).tolist(), "fp32_decoded": rel(rep32[pick].cpu(), rep_o).tolist(),
                              "bf16_latent": rel(lat16[pick].cpu(), lat_o).tolist(), "bf16_decoded": rel(rep16[pick].cpu(), rep_o).tolist()}
        print(json.dumps({"oracle_rows": res["oracle_rows"]}), flush=True)

    # ---- (4) Griffin-Lim
    from oracle import griffinlim_ref
    from tqdne_b200.representation import LogSpectrogram

    gl = {}
    waves = {}
    for name, prec, env in (("fp64", "fp64", None), ("fp64_unfused", "fp64", "1"), ("fp32_fused", "fp32", None),
                            ("fp32_unfused", "fp32", "1")):
        if env:
            os.environ["TQ_GL_LEGACY"] = env
        else:
            os.environ.pop("TQ_GL_LEGACY", None)
        ls = LogSpectrogram(stft_channels=256, hop_size=32, precision=prec)
        ms, w = timed(lambda: ls.invert_representation_device(rep32), n=2)
        waves[name] = w.double()
        gl[name + "_ms_per_%d_items" % (3 * B)] = ms
    os.environ.pop("TQ_GL_LEGACY", None)
    for name in ("fp64_unfused", "fp32_fused", "fp32_unfused"):
        gl[name + "_vs_fp64_kernel"] = dist(rel(waves[name].flatten(0, 1), waves["fp64"].flatten(0, 1)))
    # bf16 pipeline end to end in the waveform domain (fp64 Griffin-Lim on both spectrograms)
    ls64 = LogSpectrogram(stft_channels=256, hop_size=32, precision="fp64")
    w16 = ls64.invert_representation_device(rep16).double()
    gl["waveform_bf16_vs_fp32_pipeline_fp64gl"] = dist(rel(w16.flatten(0, 1), waves["fp64"].flatten(0, 1)))
    # spectral convergence of every variant: || |STFT(w)| - S || / ||S||
    S = torch.exp((rep32.double() + 1) / 2 * (3 - np.log(1e-8)) + np.log(1e-8))

    def sc(w):
        lsf = LogSpectrogram(stft_channels=256, hop_size=32, precision="fp64")
        r = lsf.get_representation_device(w.float()).double()
        mag = torch.exp((r + 1) / 2 * (3 - np.log(1e-8)) + np.log(1e-8))
        return dist(((mag - S).flatten(2).norm(dim=2) / S.flatten(2).norm(dim=2)).flatten().cpu().numpy())

    for name in waves:
        gl[name + "_spectral_convergence"] = sc(waves[name])
    items = rep32.flatten(0, 1)[: args.gl_oracle_items].cpu().numpy()
    t0 = time.time()
    o64 = griffinlim_ref.logspec_inverse(items, n_iter=128, precision="fp64")
    o32 = griffinlim_ref.logspec_inverse(items, n_iter=128, precision="fp32")
    gl["numpy_seconds_per_item_fp64+fp32"] = (time.time() - t0) / len(items)
    n = len(items)
    o64t, o32t = torch.from_numpy(o64).double(), torch.from_numpy(o32.astype(np.float64))
    gl["numpy_fp32_vs_numpy_fp64"] = rel(o32t, o64t).tolist()
    for name in waves:
        gl[name + "_vs_numpy_fp64"] = rel(waves[name].flatten(0, 1)[:n].cpu(), o64t).tolist()
    res["griffinlim"] = gl
    print(json.dumps({"griffinlim": gl}), flush=True)
    Path(args.out).parent.mkdir(parents=True, exist_ok=True)
    Path(args.out).write_text(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
