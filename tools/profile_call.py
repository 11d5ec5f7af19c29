"""One denoiser call (latent UNet, batch 256, bf16) replayed eagerly inside an NVTX range for ncu:

  ncu --nvtx --nvtx-include "denoiser_call/" --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python tools/profile_call.py
"""
import sys
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
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    if what == "unet":
        p = get_unet_plan(edm.unet, B, (32, 32), uniform_t=True)
        p.xin.t.normal_()
        p.set_cond(torch.from_numpy(cond_grid(B)).cuda())
        p.t.fill_(0.3)
        run = p.plan.run
        # the launches this capture belongs to (bench.igemm_fingerprint): travels back with the csv
        Path("gpurun_out").mkdir(exist_ok=True)
        Path(f"gpurun_out/r2_ops_unet_b{B}.names.txt").write_text("\n".join(p.plan.op_names()) + "\n")
    elif what == "decoder":
        p = get_coder_plan(edm.autoencoder.decoder, "decoder", B, (32, 32))
        p.xin.t.normal_()
        run = p.plan.run
    else:
        rep = torch.tanh(torch.randn(B, 3, 128, 128, device="cuda"))
        run = lambda: cfg.representation.invert_representation_device(rep)  # noqa: E731
    for _ in range(3):
        run()
    s.synchronize()
    torch.cuda.nvtx.range_push("denoiser_call")
    run()
    s.synchronize()
    torch.cuda.nvtx.range_pop()
print("ops", getattr(getattr(p, "plan", None), "num_ops", 1) if what != "gl" else 1)
