"""Classifier embedding for the evaluation callers: drop-in for tqdne/classifier.py on the B200 engine.

`embed` (classifier.py:51-55: Encoder -> spatial mean -> SiLU/Linear/SiLU/Linear) feeds the Frechet distance and
`forward` (classifier.py:57-59: + output_layer) the inception score of tqdne/metric.py:139-176.  The Encoder runs its
kernel plan (tcgen05 convolutions, GroupNorm, attention); the pooling and the three small dense layers are one more
plan of C-ABI kernels (`tq_plan_add_spatial_mean`, `tq_plan_add_linear`).  Training the classifier is out of scope.
"""

from __future__ import annotations

import torch
from torch import nn

from .blocks import Encoder
from .engine import device_guard, Plan, nchw_to_nhwc, require_cuda
from .lightning_shim import LightningModule


class LithningClassifier(LightningModule):
    """Same constructor signature, attribute names and state_dict keys as the reference (the class name keeps the
    reference's spelling)."""

    def __init__(self, encoder_config: dict, num_classes: int, loss: nn.Module | None = None, metrics: list | None = None,
                 optimizer_params: dict | None = None):
        super().__init__()
        self.encoder = Encoder(**encoder_config)
        out_channels = encoder_config["out_channels"]
        self.output_MLP = nn.Sequential(nn.SiLU(), nn.Linear(out_channels, out_channels), nn.SiLU(),
                                        nn.Linear(out_channels, out_channels))
        self.output_layer = nn.Linear(out_channels, num_classes)
        self.loss = loss
        self.optimizer_params = optimizer_params
        self.save_hyperparameters(ignore=("loss", "metrics"))

    def _apply(self, fn, *args, **kwargs):
        self.__dict__.pop("_tq_head", None)
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self.__dict__.pop("_tq_head", None)
        return super().load_state_dict(*args, **kwargs)

    def _head(self, N: int, spatial: tuple):
        """Plan of the pooling + dense layers reading the encoder plan's fp32 channels-last output in place."""
        from .lowering import get_coder_plan

        enc = get_coder_plan(self.encoder, "encoder", N, spatial)
        stamp = tuple((p.data_ptr(), p._version) for p in list(self.output_MLP.parameters()) + list(self.output_layer.parameters()))
        cache = self.__dict__.setdefault("_tq_head", {})
        key = (N, tuple(spatial), id(enc))
        hit = cache.get(key)
        if hit is not None and hit["stamp"] == stamp:
            return enc, hit
        dev = enc.out.t.device
        f32 = dict(device=dev, dtype=torch.float32)
        Cc = enc.cout
        plan = Plan(dev, torch.float32)
        pooled = torch.empty(N, Cc, **f32)
        h1 = torch.empty(N, Cc, **f32)
        emb = torch.empty(N, Cc, **f32)
        logits = torch.empty(N, self.output_layer.weight.shape[0], **f32)
        l1, l2 = self.output_MLP[1], self.output_MLP[3]
        w = [t.detach().float().contiguous() for t in (l1.weight, l1.bias, l2.weight, l2.bias, self.output_layer.weight,
                                                       self.output_layer.bias)]
        plan.spatial_mean(enc.out.t, N, enc.out.H * enc.out.W, Cc, enc.out.C, pooled)
        plan.linear(pooled, w[0], w[1], N, act_in=True, y=h1)
        plan.linear(h1, w[2], w[3], N, act_in=True, y=emb)
        plan.linear(emb, w[4], w[5], N, act_in=False, y=logits)
        hit = {"stamp": stamp, "plan": plan, "emb": emb, "logits": logits}
        cache[key] = hit
        return enc, hit

    @torch.no_grad()
    def _run(self, x: torch.Tensor):
        require_cuda(x, "x")
        N, spatial = x.shape[0], tuple(x.shape[2:])
        with device_guard(x.device):
            enc, head = self._head(N, spatial)
            xin = nchw_to_nhwc(x.to(torch.float32), enc.act_dtype, enc.cin_pad)
            enc.xin.t.copy_(xin.view(-1))
            enc.run()
            head["plan"].run()
        return head

    def embed(self, x):
        """reference: classifier.py:51-55."""
        return self._run(x)["emb"].clone()

    def forward(self, x):
        """reference: classifier.py:57-59."""
        return self._run(x)["logits"].clone()

    def training_step(self, batch, batch_idx):  # pragma: no cover
        raise NotImplementedError("tqdne_b200 accelerates sampling and evaluation; train with the reference tqdne package")

    validation_step = training_step
