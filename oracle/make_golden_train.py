"""Golden vector for the training step -- TEST INFRASTRUCTURE (see oracle/__init__.py).

Runs the UNMODIFIED reference `tqdne.edm.LightningEDM.step` (edm.py:115-134) of the 1D EDM UNet (train_1d_edm config)
under autograd from /root/reference (build container only), dropout switched off (`eval()`; the draw of sigma and noise
inside `step` is reproduced by the seed), and stores the batch, the drawn sigma / noise, the loss, the L2 norm of every
parameter gradient and a few whole gradient tensors.

    python -m oracle.make_golden_train
"""
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch

from oracle import reference_loader
from oracle.weights import seeded_state_dict, shapes_of

OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"
FEATURES = ("hypocentral_distance", "magnitude", "vs30", "hypocentre_depth", "azimuthal_gap")
KEEP = ["unet.out.2.weight", "unet.input_blocks.0.0.weight", "unet.time_mlp.2.weight", "unet.cond_mlp.0.weight",
        "unet.middle_block.1.qkv.bias", "unet.output_blocks.11.0.in_layers.0.weight", "unet.input_blocks.3.0.op.weight",
        "unet.output_blocks.2.2.conv.bias", "unet.input_blocks.1.0.emb_layers.1.weight"]


def main():
    tq = reference_loader.load()
    cfg = SimpleNamespace(features_keys=FEATURES, channels=6)
    ucfg = tq.architectures.get_1d_unet_config(cfg, 6, 6)
    mod = tq.edm.LightningEDM(ucfg, {"learning_rate": 1e-4, "max_steps": 10, "eta_min": 0.0}, num_sampling_steps=18)
    seed_w, seed_b, seed_s = 41, 42, 43
    mod.load_state_dict(seeded_state_dict(shapes_of(mod), seed_w))
    mod.eval()   # dropout off; everything else of step() is unchanged
    g = torch.Generator().manual_seed(seed_b)
    N, L = 2, 512
    x = torch.randn(N, 6, L, generator=g)
    cond = torch.randn(N, 5, generator=g)
    torch.manual_seed(seed_s)
    loss = mod.step({"signal": x, "cond": cond}, 0)
    loss.backward()
    # the same two draws step() made (edm.py:125-127): eps -> sigma, then the noise
    torch.manual_seed(seed_s)
    sigma = mod.edm.sigma(torch.randn(N))
    noise = torch.randn_like(x)
    grads = {n: p.grad for n, p in mod.named_parameters() if p.requires_grad}
    names = sorted(grads)
    out = dict(seed=np.int64(seed_w), x=x.numpy(), cond=cond.numpy(), sigma=sigma.numpy(), noise=noise.numpy(),
               loss=np.float64(loss.item()), grad_names=np.array(names), grad_norms=np.array([float(grads[n].norm()) for n in names]))
    for k in KEEP:
        out["grad:" + k] = grads[k].numpy()
    np.savez_compressed(OUT / "train_step_1d.npz", **out)
    print("loss", loss.item(), "params", len(names), "total grad norm", float(torch.cat([g_.flatten() for g_ in grads.values()]).norm()))


if __name__ == "__main__":
    main()
