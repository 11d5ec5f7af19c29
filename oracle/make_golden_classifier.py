"""Golden vectors for the evaluation-time classifier -- TEST INFRASTRUCTURE (see oracle/__init__.py).

Runs the UNMODIFIED reference `tqdne.classifier.LithningClassifier` (embed / forward, classifier.py:51-59) with the
encoder configuration of experiments/train_classifier.py:70-82 and `tqdne.metric.frechet_distance`
(metric.py:13-44) from /root/reference (build container only).  `torchmetrics` is absent: a stand-in module with empty
`Metric` / `MetricCollection` classes is registered first (the metrics are only touched by training / validation steps).

    python -m oracle.make_golden_classifier
"""
import sys
import types
from pathlib import Path

import numpy as np
import torch
from torch import nn

from oracle import reference_loader
from oracle.weights import seeded_state_dict, shapes_of

OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"

ENCODER_CONFIG = {"in_channels": 3, "model_channels": 64, "channel_mult": (1, 2, 4, 4), "out_channels": 256,
                  "num_res_blocks": 2, "attention_resolutions": (8,), "dims": 2, "conv_kernel_size": 3, "num_heads": 4,
                  "dropout": 0.1, "flash_attention": False}
NUM_CLASSES = 5


def main():
    reference_loader.load()
    if "torchmetrics" not in sys.modules:
        tm = types.ModuleType("torchmetrics")

        class Metric(nn.Module):
            pass

        class MetricCollection(nn.Module):
            def __init__(self, metrics=None):
                super().__init__()

        tm.Metric, tm.MetricCollection = Metric, MetricCollection
        sys.modules["torchmetrics"] = tm
    import tqdne.classifier as rcls

    clf = rcls.LithningClassifier(ENCODER_CONFIG, NUM_CLASSES, nn.CrossEntropyLoss(), [], {})
    clf.load_state_dict(seeded_state_dict(shapes_of(clf), 21))
    clf.eval()
    g = torch.Generator().manual_seed(22)
    x = torch.randn(3, 3, 64, 64, generator=g)   # 64 x 64: attention at ds = 8 sees T = 64 tokens
    with torch.no_grad():
        emb = clf.embed(x)
        logits = clf(x)
    out = dict(x=x.numpy(), emb=emb.numpy(), logits=logits.numpy())
    try:
        import scipy.linalg as sla
        import tqdne.metric as rmet

        # the reference calls sqrtm(..., disp=False) -> (sqrtm, error estimate); SciPy >= 1.18 (this image) dropped the
        # keyword.  The reference source stays untouched: the old calling convention is restored around the new function.
        _sqrtm = sla.sqrtm

        def sqrtm_compat(a, disp=True, **kw):
            r = _sqrtm(a, **kw)
            return r if disp else (r, 0.0)

        rmet.linalg.sqrtm = sqrtm_compat

        rng = np.random.RandomState(23)
        a = rng.randn(40, 12) @ rng.randn(12, 12) * 0.3 + 0.5
        b = rng.randn(48, 12) @ rng.randn(12, 12) * 0.3
        out.update(fid_a=a, fid_b=b, fid=np.float64(rmet.frechet_distance(a, b)))
    except Exception as e:  # noqa: BLE001 -- tqdne.metric imports tqdne.representation (librosa at class level on some versions)
        print("frechet_distance golden skipped:", type(e).__name__, e)
    np.savez_compressed(OUT / "classifier.npz", **out)
    print({k: np.shape(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
