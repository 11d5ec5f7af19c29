"""Evaluation loop on the B200 engine -- the batch loop of experiments/evaluate.py:103-147 (`predict`) and of the
validation-end hook tqdne/logging.py:42-90 (`LogCallback`): for every batch of the data set, sample with the batch's
conditioning (`edm.evaluate`), invert the representation, and embed target and prediction with the classifier.

The reference wraps this loop in data-set / HDF5 / wandb plumbing (out of scope); here it takes any
iterable of batches (dicts with "signal", "waveform", "cond" [, "cond_signal"]) and returns the same arrays the reference
writes: target / predicted waveform, target / predicted signal, target / predicted classifier embedding and logits.
Everything between the batch's tensors and the host arrays runs on the device: sampler, decoder, representation
inverse, the forward representation of the classifier's input and the classifier itself.
"""

from __future__ import annotations

import numpy as np
import torch

from .engine import device_guard

KEYS = ("target_waveform", "predicted_waveform", "target_signal", "predicted_signal", "target_classifier_embedding",
        "predicted_classifier_embedding", "target_classifier_pred", "predicted_classifier_pred")


@torch.no_grad()
def predict(batches, edm, classifier, representation, classifier_representation=None, device=None) -> dict:
    """reference: experiments/evaluate.py:20-147.  `representation` is the EDM config's representation (inverse applied to the
    prediction), `classifier_representation` the classifier config's (None or the same type: the classifier consumes the
    EDM's signal directly, evaluate.py:120-122).  Returns {key: float32 array over all rows} for KEYS."""
    device = torch.device(device) if device is not None else next(edm.parameters()).device
    out = {k: [] for k in KEYS}
    same = classifier_representation is None or type(classifier_representation) is type(representation)
    with device_guard(device):
        for batch in batches:
            out["target_waveform"].append(np.asarray(_host(batch["waveform"]), dtype=np.float32))
            out["target_signal"].append(np.asarray(_host(batch["signal"]), dtype=np.float32))
            dev_batch = {k: torch.as_tensor(v).to(device) for k, v in batch.items()}
            pred_signal = edm.evaluate(dev_batch)                                   # evaluate.py:113
            out["predicted_signal"].append(pred_signal.float().cpu().numpy())
            if hasattr(representation, "invert_representation_device"):
                pred_wave_dev = representation.invert_representation_device(pred_signal)
                pred_wave = pred_wave_dev.float().cpu().numpy()
            else:
                pred_wave = np.asarray(representation.invert_representation(pred_signal), dtype=np.float32)
                pred_wave_dev = torch.from_numpy(pred_wave).to(device)
            out["predicted_waveform"].append(pred_wave.astype(np.float32))
            if same:                                                                # evaluate.py:120-133
                tgt_in, pred_in = dev_batch["signal"].float(), pred_signal.float()
            else:
                rep = classifier_representation
                tgt_in = rep.get_representation_device(dev_batch["waveform"].float()).float()
                pred_in = rep.get_representation_device(pred_wave_dev.float()).float()
            for name, x in (("target", tgt_in), ("predicted", pred_in)):
                head = classifier._run(x)       # one encoder pass gives embed(x) and output_layer(embed(x)), evaluate.py:135-146
                out[f"{name}_classifier_embedding"].append(head["emb"].float().cpu().numpy())
                out[f"{name}_classifier_pred"].append(head["logits"].float().cpu().numpy())
    return {k: np.concatenate(v) if v else np.zeros((0,), np.float32) for k, v in out.items()}


def _host(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)
