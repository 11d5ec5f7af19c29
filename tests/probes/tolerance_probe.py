#!/usr/bin/env python
"""Measured values behind every bf16 bound of tests/test_engine_gpu.py that is not the plain north_star 1e-2:
prints rel-L2 of the engine against the reference goldens for the short-ladder cases (4-step latent, 3-step 1D), per
domain (sampler state, decoded spectrogram, waveform) and per precision, so that the bounds can be stated with their
derivation instead of a safety factor.

    python tests/probes/tolerance_probe.py > gpurun_out/tolerance_probe.json
"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

import tqdne_b200 as tq  # noqa: E402
from tests.conftest import rel_l2  # noqa: E402
from tests.helpers import CFG, arch, golden, seeded, unet_cfg  # noqa: E402


def latent_edm(seed, steps, mode):
    enc_cfg, dec_cfg = arch().get_2d_autoencoder_configs(CFG)
    ae = tq.LightningAutoencoder(enc_cfg, dec_cfg, {})
    edm = tq.LightningEDM(unet_cfg("latent2d"), {}, num_sampling_steps=steps, autoencoder=ae)
    seeded(edm, seed).cuda()
    return edm.set_engine_precision(mode)


def main():
    out = {}
    g = golden("edm_heun4_latent")
    for mode in ("fp32", "bf16"):
        edm = latent_edm(g["seed"], 4, mode)
        lat = edm.sample_deterministically(g["eps"].cuda(), g["sigmas"], None, g["cond"].cuda())
        dec = edm.autoencoder.decode(lat.float())
        dec_of_ref = edm.autoencoder.decode(g["latent"].float().cuda())
        out[f"heun4_latent_{mode}"] = {
            "latent": rel_l2(lat.cpu(), g["latent"]), "decoded": rel_l2(dec.cpu(), g["decoded"]),
            "decoder_alone_on_reference_latent": rel_l2(dec_of_ref.cpu(), g["decoded"]),
            "per_sample_latent": [rel_l2(lat[i].cpu(), g["latent"][i]) for i in range(lat.shape[0])],
        }
    g = golden("edm_heun3_1d")
    for mode in ("fp32", "bf16"):
        edm = tq.LightningEDM(unet_cfg("1d"), {}, num_sampling_steps=3)
        seeded(edm, g["seed"]).cuda()
        edm.set_engine_precision(mode)
        s = edm.sample_deterministically(g["eps"].cuda(), g["sigmas"], None, g["cond"].cuda())
        env = tq.MovingAverageEnvelope()
        wave = env.invert_representation(s.float())
        wave_of_ref = env.invert_representation(g["sample"].float().cuda())
        # first-order amplification of the representation error by the envelope inverse: w = x * (exp(l + c) - eps)
        ref = g["sample"].double()
        d = (s.cpu().double() - ref)
        C = ref.shape[1] // 2
        out[f"heun3_1d_{mode}"] = {
            "sample": rel_l2(s.cpu(), g["sample"]), "waveform": rel_l2(wave, g["waveform"]),
            "inverse_alone_on_reference_sample": rel_l2(wave_of_ref, g["waveform"]),
            "sample_signal_channels": rel_l2(s[:, :C].cpu(), g["sample"][:, :C]),
            "sample_envelope_channels": rel_l2(s[:, C:].cpu(), g["sample"][:, C:]),
            "abs_err_envelope_channels_rms": float(d[:, C:].pow(2).mean().sqrt()),
            "abs_err_envelope_channels_max": float(d[:, C:].abs().max()),
        }
    for name in ("unet_latent2d", "unet_1d", "unet_pixel2d"):
        gg = golden(name)
        kind = {"unet_latent2d": "latent2d", "unet_1d": "1d", "unet_pixel2d": "pixel2d"}[name]
        net = seeded(tq.UNetModel(**unet_cfg(kind)), gg["seed"]).cuda()
        for mode, dt in (("fp32", torch.float32), ("bf16", torch.bfloat16)):
            net.engine_dtype = dt
            y = net(gg["x"].cuda(), gg["t"].cuda(), gg["cond"].cuda())
            out[f"{name}_{mode}"] = rel_l2(y.cpu(), gg["y"])
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
