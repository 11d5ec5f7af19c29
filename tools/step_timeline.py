"""Where does one bench step (batch 256: 25 Heun steps -> decode -> Griffin-Lim) spend its time?
Times the three stages separately (synchronised wall clock, 3 repetitions after warm-up) and the whole step."""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

import tqdne_b200 as tq  # noqa: E402
from bench import build_state_dict, cond_grid  # noqa: E402
from tqdne_b200.config import LatentSpectrogramConfig  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cfg = LatentSpectrogramConfig()
enc_cfg, dec_cfg = tq.get_2d_autoencoder_configs(cfg)
edm = tq.LightningEDM(tq.get_2d_unet_config(cfg, 8, 8), {}, autoencoder=tq.LightningAutoencoder(enc_cfg, dec_cfg, {}))
edm.load_state_dict(build_state_dict(edm))
edm.eval().cuda().set_engine_precision("bf16")
cond = torch.from_numpy(cond_grid(B)).cuda()
noise = torch.randn(B, 8, 32, 32, device="cuda", dtype=torch.float64)
sig = edm.edm.sampling_sigmas(25, device="cpu")
rep_inv = cfg.representation


def timed(name, fn, reps=3):
    for _ in range(2):
        out = fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    print(f"{name:40s} {(time.perf_counter() - t0) / reps * 1e3:9.3f} ms", flush=True)
    return out


lat = timed("heun (49 denoiser calls + updates)", lambda: edm.sample_deterministically(noise * 80.0, sig, None, cond))


from tqdne_b200 import engine  # noqa: E402


def decode2():
    with edm._engine_stream():
        x = engine.nchw_to_nhwc(lat, torch.float64)
        return edm._decode_latents(x, B, 8, (32, 32))


rep = timed("decode (4 micro-batches of 64)", decode2)
timed("griffin-lim (768 items, 128 iterations)", lambda: rep_inv.invert_representation_device(rep))
timed("whole step through sample()", lambda: rep_inv.invert_representation_device(
    edm.sample((B, 3, 128, 128), cond=cond, noise=noise)))
for mp in [int(a) for a in sys.argv[2:]]:
    edm.max_positions_per_pass = mp * 1024
    timed(f"whole step, Heun micro-batch {mp}", lambda: rep_inv.invert_representation_device(
        edm.sample((B, 3, 128, 128), cond=cond, noise=noise)))
