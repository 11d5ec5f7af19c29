"""A/B inside one process: decoder micro-batch size (edm.decode_micro_batch) at batch 256."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

import tqdne_b200 as tq  # noqa: E402
from bench import build_state_dict  # noqa: E402
from tqdne_b200.config import LatentSpectrogramConfig  # noqa: E402

B = 256
cfg = LatentSpectrogramConfig()
enc_cfg, dec_cfg = tq.get_2d_autoencoder_configs(cfg)
edm = tq.LightningEDM(tq.get_2d_unet_config(cfg, 8, 8), {}, autoencoder=tq.LightningAutoencoder(enc_cfg, dec_cfg, {}))
edm.load_state_dict(build_state_dict(edm))
edm.eval().cuda().set_engine_precision("bf16")
x = torch.randn(B, 1024, 8, device="cuda", dtype=torch.float64)
s = torch.cuda.Stream()
for rep in range(2):
    for mb in (64, 128, 256):
        edm.decode_micro_batch = mb
        with torch.cuda.stream(s):
            for _ in range(3):
                edm._decode_latents(x, B, 8, (32, 32))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            for _ in range(5):
                edm._decode_latents(x, B, 8, (32, 32))
            e1.record(s)
            s.synchronize()
        print(f"decode_micro_batch {mb:3d}: {e0.elapsed_time(e1) / 5:7.2f} ms per 256 samples "
              f"(peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB)", flush=True)
