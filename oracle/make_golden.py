"""Generate tests/golden/*.npz by executing the UNMODIFIED reference (build container only).

    python -m oracle.make_golden

Each fixture holds seeded inputs and the reference's outputs; weights are NOT stored -- they are rebuilt from
oracle.weights.seeded_state_dict(shapes, seed) wherever the fixture is checked.  Reference entry points run:
  tqdne.unet.UNetModel.forward, tqdne.blocks.Decoder/Encoder.forward, tqdne.edm.LightningEDM.forward /
  .sample_deterministically / .sample, tqdne.autoencoder.LightningAutoencoder.decode,
  tqdne.representation.MovingAverageEnvelope.invert_representation,
  tqdne.representation.LogSpectrogram.invert_representation (with oracle.griffinlim_ref.griffinlim standing in
  for the absent librosa -- see oracle/__init__.py "PARITY UNPINNED").
"""

from __future__ import annotations

import sys
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import griffinlim_ref, reference_loader  # noqa: E402
from oracle.weights import seeded_state_dict, shapes_of  # noqa: E402

GOLD = ROOT / "tests" / "golden"
FEATURES = ("hypocentral_distance", "magnitude", "vs30", "hypocentre_depth", "azimuthal_gap")
CFG = SimpleNamespace(features_keys=FEATURES, channels=3, latent_channels=8)


def rand(shape, seed, dtype=torch.float32, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g, dtype=dtype) * scale


def load_seeded(module, seed):
    module.load_state_dict(seeded_state_dict(shapes_of(module), seed))
    return module.eval()


def save(name, **arrays):
    GOLD.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(GOLD / f"{name}.npz", **{k: (v.numpy() if isinstance(v, torch.Tensor) else np.asarray(v))
                                                  for k, v in arrays.items()})
    print(f"wrote {name}.npz", {k: tuple(np.shape(v)) for k, v in arrays.items()})


@torch.no_grad()
def main():
    tq = reference_loader.load()
    arch = tq.architectures
    torch.set_grad_enabled(False)

    # ---- UNet forward: latent 2D, 1D, pixel 2D --------------------------------------------------------
    for name, cfg, xshape, seed in [
        ("unet_latent2d", arch.get_2d_unet_config(CFG, 8, 8), (2, 8, 32, 32), 11),
        ("unet_1d", arch.get_1d_unet_config(CFG, 6, 6), (2, 6, 512), 12),
        ("unet_pixel2d", arch.get_2d_unet_config(CFG, 3, 3), (1, 3, 64, 64), 13),
    ]:
        net = load_seeded(tq.unet.UNetModel(**cfg), seed)
        x = rand(xshape, seed + 100)
        sigma = torch.tensor([2.0, 0.3][: xshape[0]])
        t = 0.25 * sigma.log()
        cond = rand((xshape[0], 5), seed + 200)
        y = net(x, t, cond)
        save(name, x=x, t=t, cond=cond, y=y, seed=seed)

    # ---- autoencoder decoder / encoder ---------------------------------------------------------------------
    enc_cfg, dec_cfg = arch.get_2d_autoencoder_configs(CFG)
    dec = load_seeded(tq.blocks.Decoder(**dec_cfg), 21)
    z = rand((1, 8, 32, 32), 121)
    save("decoder2d", z=z, y=dec(z), seed=21)
    enc = load_seeded(tq.blocks.Encoder(**enc_cfg), 22)
    x = rand((1, 3, 128, 128), 122)
    save("encoder2d", x=x, y=enc(x), seed=22)

    # ---- EDM denoiser + Heun sampler (latent, with autoencoder) -----------------------------------------------
    ae = tq.autoencoder.LightningAutoencoder(enc_cfg, dec_cfg, {})
    edm = tq.edm.LightningEDM(arch.get_2d_unet_config(CFG, 8, 8), {}, num_sampling_steps=4, autoencoder=ae)
    load_seeded(edm, 31)
    x = rand((2, 8, 32, 32), 131, scale=3.0)
    sigma = torch.tensor([5.0, 0.7])
    cond = rand((2, 5), 231)
    save("edm_denoiser", x=x, sigma=sigma, cond=cond, D=edm(x, sigma, None, cond), seed=31)

    sigmas = edm.edm.sampling_sigmas(4)
    eps = rand((2, 8, 32, 32), 132, dtype=torch.float64) * sigmas[0]
    lat = edm.sample_deterministically(eps, sigmas, None, cond)
    dec_out = edm.autoencoder.decode(lat.to(torch.float32))
    save("edm_heun4_latent", eps=eps, sigmas=sigmas, cond=cond, latent=lat, decoded=dec_out, seed=31)

    torch.manual_seed(1234)
    out = edm.sample((2, 3, 128, 128), cond=cond)
    save("edm_sample_seed1234", cond=cond, decoded=out, seed=31, torch_seed=1234)

    # 1D EDM (config 1 family), 3 steps, short signal
    edm1 = tq.edm.LightningEDM(arch.get_1d_unet_config(CFG, 6, 6), {}, num_sampling_steps=3)
    load_seeded(edm1, 41)
    sig1 = edm1.edm.sampling_sigmas(3)
    eps1 = rand((2, 6, 256), 141, dtype=torch.float64) * sig1[0]
    cond1 = rand((2, 5), 241)
    out1 = edm1.sample_deterministically(eps1, sig1, None, cond1)
    import tqdne.representation as rrep

    wave1 = rrep.MovingAverageEnvelope().invert_representation(out1.to(torch.float32))
    save("edm_heun3_1d", eps=eps1, sigmas=sig1, cond=cond1, sample=out1, waveform=wave1, seed=41)

    # ---- representation inverses ---------------------------------------------------------------------------------
    rep = rand((2, 6, 300), 51)
    save("mavg_inverse", rep=rep, wave=rrep.MovingAverageEnvelope().invert_representation(rep))

    ls = object.__new__(rrep.LogSpectrogram)   # the ctor imports librosa; set what it would set
    ls.clip, ls.log_clip, ls.log_max, ls.library = 1e-8, np.log(1e-8), 3, "librosa"
    n_iter = 8
    ls.istft = lambda s: griffinlim_ref.griffinlim(s, hop=32, n_fft=256, n_iter=n_iter, seed=0)
    rep = torch.tanh(rand((1, 3, 128, 128), 52) * 0.5)
    wave = ls.invert_representation(rep)
    save("logspec_inverse_iter8", rep=rep, wave=wave, n_iter=n_iter)


if __name__ == "__main__":
    main()
