"""Classifier-embedding metrics of the evaluation callers: drop-in for the neural metrics of tqdne/metric.py.

`FrechetInceptionDistance` / `InceptionScore` (metric.py:139-176) batch the inputs through
`LithningClassifier.embed` / `.forward` -- which run on the B200 engine -- and reduce the [N, 256] embeddings on the
host with NumPy / SciPy exactly like the reference (`frechet_distance`, metric.py:13-44): a 256 x 256 matrix square
root is host work there too.
"""

from __future__ import annotations

import numpy as np
import torch
from scipy import linalg


def frechet_distance(x: np.ndarray, y: np.ndarray, isotropic: bool = False, eps: float = 1e-6):
    """reference: tqdne/metric.py:13-44."""
    mu_x, mu_y = x.mean(0), y.mean(0)
    if isotropic:
        return np.sum((mu_x - mu_y) ** 2) + np.sum((x.std(0) - y.std(0)) ** 2)
    cov_x, cov_y = np.cov(x, rowvar=False), np.cov(y, rowvar=False)
    covmean = linalg.sqrtm(cov_x @ cov_y)
    if not np.isfinite(covmean).all():
        print(f"fid calculation produces singular product; adding {eps} to diagonal of cov estimates")
        offset = np.eye(cov_x.shape[0]) * eps
        covmean = linalg.sqrtm((cov_x + offset) @ (cov_y + offset))
    if np.iscomplexobj(covmean):
        if not np.allclose(np.diagonal(covmean).imag, 0, atol=1e-3):
            raise ValueError(f"Imaginary component {np.max(np.abs(covmean.imag))}")
        covmean = covmean.real
    return np.sum((mu_x - mu_y) ** 2) + np.trace(cov_x) + np.trace(cov_y) - 2 * np.trace(covmean)


class NeuralMetric:
    """reference: tqdne/metric.py:109-136 (waveforms -> representation -> classifier)."""

    def __init__(self, classifier, representation, batch_size: int = 128):
        self.classifier = classifier.eval()
        self.representation = representation
        self.batch_size = batch_size

    @property
    def name(self):
        return self.__class__.__name__

    def _rep(self, wave):
        dev = self.classifier.device
        if hasattr(self.representation, "get_representation_device"):
            w = torch.as_tensor(np.asarray(wave) if not torch.is_tensor(wave) else wave, device=dev, dtype=torch.float32)
            return self.representation.get_representation_device(w).to(torch.float32)
        return torch.as_tensor(self.representation.get_representation(wave), device=dev, dtype=torch.float32)

    def __call__(self, pred, target=None):
        pred = self._rep(pred)
        if target is not None:
            target = self._rep(target)
        return self.compute(pred, target)

    def _batched(self, fn, x):
        return np.concatenate([fn(x[i:i + self.batch_size]).cpu().numpy() for i in range(0, len(x), self.batch_size)])


class FrechetInceptionDistance(NeuralMetric):
    @torch.no_grad()
    def compute(self, pred, target):
        return frechet_distance(self._batched(self.classifier.embed, pred), self._batched(self.classifier.embed, target))


class InceptionScore(NeuralMetric):
    @torch.no_grad()
    def compute(self, pred, target=None):
        prob = self._batched(lambda b: torch.softmax(self.classifier(b), dim=-1), pred)
        marginal = prob.mean(axis=0)
        kl = np.sum(prob * (np.log(prob) - np.log(marginal)), axis=-1)
        return np.exp(kl.mean())
