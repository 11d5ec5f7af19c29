import sys, json
sys.path.insert(0, '.')
import torch, tqdne_b200 as tq
from bench import cond_grid
from tqdne_b200.synthetic_weights import seeded_state_dict, shapes_of
from tqdne_b200.config import SpectrogramConfig
cfg = SpectrogramConfig()
edm = tq.LightningEDM(tq.get_2d_unet_config(cfg, 3, 3), {}, num_sampling_steps=32)
sd = seeded_state_dict(shapes_of(edm), 0)
edm.load_state_dict(sd); edm.eval().cuda().set_engine_precision("bf16")
B = 64
cond = torch.from_numpy(cond_grid(B)).cuda()
noise = torch.randn((B, 3, 128, 128), device="cuda", dtype=torch.float64)
for mb in (16, 32, 64):
    edm.max_positions_per_pass = mb * 128 * 128
    for _ in range(1): edm.sample((B, 3, 128, 128), cond=cond, noise=noise)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); edm.sample((B, 3, 128, 128), cond=cond, noise=noise); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(json.dumps({"micro_batch": mb, "ms": ms, "waveforms_per_s": B / ms * 1e3, "tflops": B / ms * 1e3 * 63 * 271.89e9 / 1e12}), flush=True)
