"""Shared fixtures: named architectures, seeded weights, golden vectors."""
from types import SimpleNamespace

import numpy as np
import torch

from oracle.weights import seeded_state_dict, shapes_of
from tests.conftest import GOLDEN

FEATURES = ("hypocentral_distance", "magnitude", "vs30", "hypocentre_depth", "azimuthal_gap")
CFG = SimpleNamespace(features_keys=FEATURES, channels=3, latent_channels=8)


def golden(name):
    z = np.load(GOLDEN / f"{name}.npz")
    return {k: (torch.from_numpy(z[k]) if z[k].ndim else z[k].item()) for k in z.files}


def arch():
    import tqdne_b200.architectures as a

    return a


def unet_cfg(kind):
    a = arch()
    return {"latent2d": a.get_2d_unet_config(CFG, 8, 8), "1d": a.get_1d_unet_config(CFG, 6, 6),
            "pixel2d": a.get_2d_unet_config(CFG, 3, 3),
            # signal-conditioned variants: cond_sample supplies the second half of the input channels (edm.py:109)
            "1d_condsample": a.get_1d_unet_config(CFG, 12, 6), "latent2d_condsample": a.get_2d_unet_config(CFG, 16, 8),
            # FiLM ResBlocks (use_scale_shift_norm, unet.py:135-139)
            "latent2d_film": dict(a.get_2d_unet_config(CFG, 8, 8), use_scale_shift_norm=True),
            "1d_film": dict(a.get_1d_unet_config(CFG, 6, 6), use_scale_shift_norm=True),
            # causal attention mask (blocks.py:181-186)
            "1d_causal": dict(a.get_1d_unet_config(CFG, 6, 6), use_causal_mask=True),
            "latent2d_causal": dict(a.get_2d_unet_config(CFG, 8, 8), use_causal_mask=True),
            # conv_resample=False: average-pool Downsample, conv-less nearest Upsample (blocks.py:59-64,104)
            "1d_pool": dict(a.get_1d_unet_config(CFG, 6, 6), conv_resample=False),
            "latent2d_pool": dict(a.get_2d_unet_config(CFG, 8, 8), conv_resample=False),
            # Fourier-embedded conditioning, one feature (unet.py:217-219,386-387)
            "1d_condembed": dict(a.get_1d_unet_config(CFG, 6, 6), cond_features=1, cond_emb_scale=0.5)}[kind]


def seeded(module, seed):
    module.load_state_dict(seeded_state_dict(shapes_of(module), seed))
    return module.eval()
