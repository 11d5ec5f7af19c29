"""CPU: the oracle restatement against the golden vectors produced by the reference itself."""
import numpy as np
import pytest
import torch

from oracle import griffinlim_ref, torch_ref
from oracle.weights import seeded_state_dict, shapes_of
from tests.conftest import GOLDEN, rel_l2
from tests.helpers import CFG, arch, golden, unet_cfg

TOL = 2e-5  # fp32 CPU restatement vs fp32 CPU reference: same ops, possibly different summation order


def _sd(module, seed):
    return seeded_state_dict(shapes_of(module), seed)


@pytest.mark.parametrize("kind,name", [("latent2d", "unet_latent2d"), ("1d", "unet_1d"), ("pixel2d", "unet_pixel2d"),
                                       ("latent2d_film", "unet_film_2d"), ("1d_film", "unet_film_1d"),
                                       ("1d_causal", "unet_causal_1d"), ("latent2d_causal", "unet_causal_2d"),
                                       ("1d_pool", "unet_poolresample_1d"), ("latent2d_pool", "unet_poolresample_2d"),
                                       ("1d_condembed", "unet_condembed_1d")])
def test_unet_restatement_matches_reference(kind, name):
    import tqdne_b200 as tq

    g = golden(name)
    cfg = unet_cfg(kind)
    sd = _sd(tq.UNetModel(**cfg), g["seed"])
    with torch.no_grad():
        y = torch_ref.unet_forward(sd, cfg, g["x"], g["t"], g["cond"])
    assert rel_l2(y, g["y"]) < TOL


def test_decoder_encoder_restatement():
    import tqdne_b200 as tq

    enc_cfg, dec_cfg = arch().get_2d_autoencoder_configs(CFG)
    g = golden("decoder2d")
    with torch.no_grad():
        y = torch_ref.decoder_forward(_sd(tq.Decoder(**dec_cfg), g["seed"]), dec_cfg, g["z"])
    assert rel_l2(y, g["y"]) < TOL
    g = golden("encoder2d")
    with torch.no_grad():
        y = torch_ref.encoder_forward(_sd(tq.Encoder(**enc_cfg), g["seed"]), enc_cfg, g["x"])
    assert rel_l2(y, g["y"]) < TOL


def _latent_edm():
    import tqdne_b200 as tq

    enc_cfg, dec_cfg = arch().get_2d_autoencoder_configs(CFG)
    ae = tq.LightningAutoencoder(enc_cfg, dec_cfg, {})
    edm = tq.LightningEDM(unet_cfg("latent2d"), {}, num_sampling_steps=4, autoencoder=ae)
    return edm, dec_cfg


def test_sampling_sigmas_match_reference_schedule():
    g = golden("edm_heun4_latent")
    assert torch.equal(torch_ref.sampling_sigmas(4), g["sigmas"])
    import tqdne_b200 as tq

    assert torch.equal(tq.EDM().sampling_sigmas(4), g["sigmas"])
    s25 = tq.EDM().sampling_sigmas(25)
    assert s25.dtype == torch.float32 and len(s25) == 26 and abs(float(s25[0]) - 80.0) < 1e-3 and s25[-1] == 0.0
    assert abs(float(s25[-2]) - 0.002) < 1e-8


def test_denoiser_and_heun_restatement():
    edm, dec_cfg = _latent_edm()
    g = golden("edm_denoiser")
    sd = _sd(edm, g["seed"])
    cfg = unet_cfg("latent2d")
    with torch.no_grad():
        D = torch_ref.denoise(sd, cfg, g["x"], g["sigma"], g["cond"])
    assert rel_l2(D, g["D"]) < TOL
    g = golden("edm_heun4_latent")
    with torch.no_grad():
        lat = torch_ref.heun_sample(sd, cfg, g["eps"], g["sigmas"], g["cond"])
        dec = torch_ref.decoder_forward(sd, dec_cfg, lat.float(), prefix="autoencoder.decoder.")
    assert lat.dtype == torch.float64
    assert rel_l2(lat, g["latent"]) < 1e-4
    assert rel_l2(dec, g["decoded"]) < 1e-4


def test_reference_rng_draw_order_is_known():
    """sample() draws randn_like(mean) (fp32) for the dummy encode, then the fp64 sampler noise (edm.py:155-160)."""
    edm, dec_cfg = _latent_edm()
    g = golden("edm_sample_seed1234")
    sd = _sd(edm, g["seed"])
    torch.manual_seed(int(g["torch_seed"]))
    torch.randn((2, 8, 32, 32), dtype=torch.float32)
    sig = torch_ref.sampling_sigmas(4)
    eps = torch.randn((2, 8, 32, 32), dtype=torch.float64) * sig[0]
    with torch.no_grad():
        lat = torch_ref.heun_sample(sd, unet_cfg("latent2d"), eps, sig, g["cond"])
        dec = torch_ref.decoder_forward(sd, dec_cfg, lat.float(), prefix="autoencoder.decoder.")
    assert rel_l2(dec, g["decoded"]) < 1e-4


def test_1d_heun_and_envelope_inverse():
    import tqdne_b200 as tq

    g = golden("edm_heun3_1d")
    cfg = unet_cfg("1d")
    edm = tq.LightningEDM(cfg, {}, num_sampling_steps=3)
    sd = _sd(edm, g["seed"])
    with torch.no_grad():
        out = torch_ref.heun_sample(sd, cfg, g["eps"], g["sigmas"], g["cond"])
    assert rel_l2(out, g["sample"]) < 1e-4
    wave = torch_ref.mavg_envelope_inverse(g["sample"].float().numpy())
    assert rel_l2(wave, g["waveform"]) < 1e-6
    g = golden("mavg_inverse")
    assert rel_l2(torch_ref.mavg_envelope_inverse(g["rep"].numpy()), g["wave"]) < 1e-7


def test_stft_pair_pinned_against_torch():
    rng = np.random.default_rng(0)
    y = rng.standard_normal(4064)
    S = griffinlim_ref.stft(y)
    win = torch.hann_window(256, periodic=True, dtype=torch.float64)
    St = torch.stft(torch.from_numpy(y), 256, 32, window=win, center=True, pad_mode="constant", return_complex=True)
    assert S.shape == (129, 128)
    assert np.abs(S - St.numpy()).max() < 1e-10
    yi = griffinlim_ref.istft(S)
    yt = torch.istft(St, 256, 32, window=win, center=True, length=4064).numpy()
    assert np.abs(yi - yt).max() < 1e-12 and np.abs(yi - y).max() < 1e-12


def test_logspec_inverse_wrapper_matches_reference_code():
    """Reference un-normalise / exp / Nyquist / reshape code (representation.py:152-175) around the same GL."""
    g = golden("logspec_inverse_iter8")
    w = griffinlim_ref.logspec_inverse(g["rep"].numpy(), n_iter=int(g["n_iter"]), precision="fp64")
    assert w.shape == (1, 3, 4064)
    assert rel_l2(w, g["wave"]) < 1e-9


def test_griffinlim_properties():
    """Unpinned boundary: check what the domain offers -- GL keeps a consistent spectrogram fixed (up to phase)
    and reduces spectral inconsistency; the kernel's FFT decomposition model equals numpy's rfft/irfft."""
    rng = np.random.default_rng(1)
    x = rng.standard_normal(256)
    assert np.abs(griffinlim_ref.kernel_model_rfft256(x) - np.fft.rfft(x)).max() < 1e-12
    X = np.fft.rfft(x) + 0j
    X[0] += 0.3j
    X[128] -= 0.2j
    assert np.abs(griffinlim_ref.kernel_model_irfft256(X) - np.fft.irfft(X, n=256)).max() < 1e-12
    y = rng.standard_normal(4064)
    S = np.abs(griffinlim_ref.stft(y))

    def inconsistency(wave):
        return np.linalg.norm(np.abs(griffinlim_ref.stft(wave)) - S) / np.linalg.norm(S)

    e4 = inconsistency(griffinlim_ref.griffinlim(S, n_iter=4))
    e32 = inconsistency(griffinlim_ref.griffinlim(S, n_iter=32))
    assert e32 < e4 < 1.0


def test_forward_representations_match_reference_golden():
    """oracle restatements of get_representation vs the reference's own methods (oracle/make_golden_forward.py)."""
    g = golden("mavg_forward")
    assert rel_l2(griffinlim_ref.mavg_forward(g["wave"].numpy()), g["rep"]) < 1e-12
    g = golden("logspec_forward")
    w = g["wave"].numpy()
    r64 = griffinlim_ref.logspec_forward(w.astype(np.float64))
    assert r64.shape == (1, 3, 128, 128) and np.abs(r64 - g["rep64"].numpy()).max() < 1e-12
    r32 = griffinlim_ref.logspec_forward(w)
    assert np.abs(r32 - g["rep32"].numpy()).max() < 1e-5


def test_griffinlim_loop_pinned_against_torchaudio():
    """torchaudio.functional.griffinlim is an independent port of librosa's fast Griffin-Lim loop: same
    momentum update, normalisation and stft/istft chain, but reflect padding, all-ones initial phase and eps 1e-16.
    The oracle run in THAT configuration must reproduce it; the default (librosa) configuration differs from it only by
    those three switches."""
    ta = pytest.importorskip("torchaudio")
    rng = np.random.default_rng(3)
    y = rng.standard_normal(4064) * np.exp(-((np.arange(4064) - 1500.0) / 900.0) ** 2)
    S = np.abs(griffinlim_ref.stft(y))                                      # [129, 128] float64, a consistent spectrogram
    win = torch.hann_window(256, periodic=True, dtype=torch.float64)
    for n_iter in (1, 6):
        ref = ta.functional.griffinlim(torch.from_numpy(S), window=win, n_fft=256, hop_length=32, win_length=256, power=1.0,
                                       n_iter=n_iter, momentum=0.99, length=4064, rand_init=False).numpy()
        out = griffinlim_ref.griffinlim(S, n_iter=n_iter, pad_mode="reflect", init="ones", eps=1e-16)
        assert out.shape == ref.shape == (4064,)
        assert np.abs(out - ref).max() < 1e-9 * np.abs(ref).max()


CLASSIFIER_ENCODER = {"in_channels": 3, "model_channels": 64, "channel_mult": (1, 2, 4, 4), "out_channels": 256,
                      "num_res_blocks": 2, "attention_resolutions": (8,), "dims": 2, "conv_kernel_size": 3, "num_heads": 4,
                      "dropout": 0.1, "flash_attention": False}   # experiments/train_classifier.py:70-82


def test_classifier_embedding_and_frechet_distance_restatement():
    """oracle classifier_embed / classifier_forward / frechet_distance against the reference's own
    LithningClassifier and tqdne.metric.frechet_distance (oracle/make_golden_classifier.py)."""
    from tqdne_b200.classifier import LithningClassifier
    from tqdne_b200.metric import frechet_distance

    g = golden("classifier")
    sd = _sd(LithningClassifier(CLASSIFIER_ENCODER, 5), 21)
    with torch.no_grad():
        emb = torch_ref.classifier_embed(sd, CLASSIFIER_ENCODER, g["x"])
        logits = torch_ref.classifier_forward(sd, CLASSIFIER_ENCODER, g["x"])
    assert rel_l2(emb, g["emb"]) < TOL and rel_l2(logits, g["logits"]) < TOL
    a, b = g["fid_a"].numpy(), g["fid_b"].numpy()
    ref = float(g["fid"])
    assert abs(torch_ref.frechet_distance(a, b) - ref) < 1e-9 * abs(ref)
    assert abs(float(frechet_distance(a, b)) - ref) < 1e-9 * abs(ref)   # the product's host-side reduction


def _train_golden():
    z = np.load(GOLDEN / "train_step_1d.npz", allow_pickle=False)
    return z


def test_training_step_oracle_matches_reference_autograd():
    """Autograd through the oracle's denoiser + the EDM loss against the UNMODIFIED reference's LightningEDM.step under
    autograd (oracle/make_golden_train.py): loss, the L2 norm of all 310 parameter gradients, nine whole gradient tensors.
    This pins the checker the GPU training parity tests use."""
    import tqdne_b200 as tq

    z = _train_golden()
    shell = tq.LightningEDM(unet_cfg("1d"), {}, num_sampling_steps=18)
    sd = _sd(shell, int(z["seed"]))
    P = {k: v.clone().requires_grad_(not k.endswith("time_embed.W")) for k, v in sd.items()}
    x, cond, sigma, noise = (torch.from_numpy(z[k]) for k in ("x", "cond", "sigma", "noise"))
    pred = torch_ref.denoise(P, unet_cfg("1d"), x + noise * sigma[:, None, None], sigma, cond)
    loss = ((pred - x) ** 2 * ((sigma**2 + 0.25) / (sigma * 0.5) ** 2)[:, None, None]).mean()
    loss.backward()
    assert abs(float(loss.detach()) - float(z["loss"])) < 1e-5 * float(z["loss"])
    names = [str(n) for n in z["grad_names"]]
    norms = torch.tensor([float(P[n].grad.norm()) for n in names], dtype=torch.float64)
    assert rel_l2(norms, torch.from_numpy(z["grad_norms"])) < 1e-4
    for k in z.files:
        if k.startswith("grad:"):
            assert rel_l2(P[k[5:]].grad, torch.from_numpy(z[k])) < 1e-4, k


# ---- round 2: stochastic sampler and the stated sizes of BASELINE.json configs[0] / configs[3] ---------------------
@pytest.mark.parametrize("name,kind", [("edm_stochastic_latent", "latent2d"), ("edm_stochastic_1d", "1d")])
def test_stochastic_sampler_restatement_matches_reference(name, kind):
    """LightningEDM.sample_stochastically (edm.py:198-230) fed with the recorded th.randn_like draws."""
    import tqdne_b200 as tq

    g = golden(name)
    cfg = unet_cfg(kind)
    sd = _sd(tq.LightningEDM(cfg, {}), g["seed"])
    with torch.no_grad():
        out = torch_ref.heun_sample_stochastic(sd, cfg, g["eps"], g["sigmas"], g["cond"], list(g["noises"]))
    assert out.dtype == torch.float64 and rel_l2(out, g["sample"]) < TOL
    # the churn is real: without it the result differs at the 1e-1 level (the fixture exercises gamma > 0)
    with torch.no_grad():
        det = torch_ref.heun_sample(sd, cfg, g["eps"], g["sigmas"], g["cond"])
    assert rel_l2(det, g["sample"]) > 1e-2


@pytest.mark.parametrize("name,kind", [("unet_1d_4064", "1d"), ("unet_1d_4096", "1d"), ("unet_pixel2d_128", "pixel2d")])
def test_denoiser_restatement_at_stated_sizes(name, kind):
    """One denoiser call of one sample at [1, 6, 4064] / [1, 6, 4096] (508 / 512 attention tokens) and
    [1, 3, 128, 128] (256 tokens): the sizes BASELINE.json configs[0] / configs[3] name."""
    import tqdne_b200 as tq

    g = golden(name)
    cfg = unet_cfg(kind)
    sd = _sd(tq.LightningEDM(cfg, {}), g["seed"])
    with torch.no_grad():
        D = torch_ref.denoise(sd, cfg, g["x"], g["sigma"], g["cond"])
    assert rel_l2(D, g["D"]) < TOL


def test_kernel_fft_network_model_is_exact():
    """The lane / register model of the fused Griffin-Lim kernel's 128-point FFT network (radix 4 x 4 x 4 x 2, two
    shared-memory transposes, one shuffle stage; output order (lane & 15) + 16 p + 64 (lane >> 4)) against np.fft."""
    rng = np.random.default_rng(3)
    x = rng.standard_normal(128) + 1j * rng.standard_normal(128)
    X = np.fft.fft(x)
    L = np.arange(32)
    kk = np.stack([(L & 15) + 16 * p + 64 * (L >> 4) for p in range(4)], 1)
    assert np.abs(griffinlim_ref.kernel_model_fft128_dif(x) - X[kk]).max() < 1e-12


@pytest.mark.parametrize("name,kind", [("edm_condsample_1d", "1d_condsample"), ("edm_condsample_2d", "latent2d_condsample")])
def test_signal_conditioned_sampler_restatement_matches_reference(name, kind):
    """LightningEDM.sample_deterministically with cond_sample (edm.py:109,171-196): the conditioning signal is concatenated
    to the scaled state in front of the UNet at every call."""
    import tqdne_b200 as tq

    g = golden(name)
    cfg = unet_cfg(kind)
    sd = _sd(tq.LightningEDM(cfg, {}), g["seed"])
    with torch.no_grad():
        out = torch_ref.heun_sample(sd, cfg, g["eps"], g["sigmas"], g["cond"], cond_sample=g["cond_sample"])
    assert out.dtype == torch.float64 and rel_l2(out, g["sample"]) < TOL
    with torch.no_grad():   # the conditioning signal matters: zeros in its place give another sample
        other = torch_ref.heun_sample(sd, cfg, g["eps"], g["sigmas"], g["cond"], cond_sample=torch.zeros_like(g["cond_sample"]))
    assert rel_l2(other, g["sample"]) > 1e-3

