"""A tiny Lightning-layout checkpoint written with the UNMODIFIED reference classes -- TEST INFRASTRUCTURE.

`hyper_parameters` pickles the reference's own `tqdne.edm.EDM` (module path `tqdne.edm`), exactly what a real
tqdne `.ckpt` holds (reference generate_waveforms.py:170 registers that class as a safe global before loading);
`ema_state` mirrors the EMA callback's on_save_checkpoint (reference ema.py:50-51).  The loader under test must
resolve those pickles without the reference package.

    python -m oracle.make_golden_ckpt      (build container only; writes tests/golden/tiny_edm_reference.ckpt)
"""
from pathlib import Path

import torch

from oracle import reference_loader

OUT = Path(__file__).resolve().parent.parent / "tests" / "golden" / "tiny_edm_reference.ckpt"


def main():
    tq = reference_loader.load()
    torch.manual_seed(71)
    cfg = dict(in_channels=2, out_channels=2, cond_features=5, dims=1, conv_kernel_size=3, model_channels=32,
               channel_mult=(1,), num_res_blocks=1, attention_resolutions=(), num_heads=1, dropout=0.0, flash_attention=False)
    e = tq.edm.EDM()                      # a plain class of class attributes; instance overrides are pickled
    e.sigma_min, e.sigma_max = 0.004, 40.0
    edm = tq.edm.LightningEDM(cfg, {"learning_rate": 1e-4, "max_steps": 10}, num_sampling_steps=7, edm=e)
    sd = {k: v.clone() for k, v in edm.state_dict().items()}
    ema = {k: (v * 0.5).clone() for k, v in edm.named_parameters() if v.requires_grad}
    ckpt = {"state_dict": sd, "ema_state": ema, "pytorch-lightning_version": "2.5.1", "epoch": 3, "global_step": 1234,
            "hyper_parameters": {"unet_config": cfg, "optimizer_params": {"learning_rate": 1e-4, "max_steps": 10},
                                 "num_sampling_steps": 7, "deterministic_sampling": True, "edm": edm.edm}}
    torch.save(ckpt, OUT)
    print(OUT, OUT.stat().st_size, "bytes;", len(sd), "tensors; edm class", type(edm.edm).__module__)


if __name__ == "__main__":
    main()
