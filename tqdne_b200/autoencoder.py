"""Latent autoencoder: drop-in for tqdne/autoencoder.py on the B200 engine.

`decode` (autoencoder.py:45-46) is on the sampling hot path; `encode` (autoencoder.py:37-43) is kept with
the reference's reparameterisation semantics.  Both run the Encoder/Decoder kernel plans.
"""

from __future__ import annotations

import torch

from .blocks import Decoder, Encoder
from .lightning_shim import LightningModule


class LightningAutoencoder(LightningModule):
    def __init__(self, encoder_config: dict, decoder_config: dict, optimizer_params: dict, kl_weight: float = 1e-6):
        super().__init__()
        self.encoder = Encoder(**encoder_config)
        self.decoder = Decoder(**decoder_config)
        self.optimizer_params = optimizer_params
        self.kl_weight = kl_weight
        self.config = encoder_config
        self.save_hyperparameters()

    @torch.no_grad()
    def _encode(self, x):
        """mean, log_std = chunk(Encoder(x), 2, dim=1); z = mean + randn * exp(log_std)."""
        mean, log_std = torch.chunk(self.encoder(x), 2, dim=1)
        latent = mean + torch.randn_like(mean) * torch.exp(log_std)
        return latent, mean, log_std

    def encode(self, x):
        return self._encode(x)[0]

    @torch.no_grad()
    def decode(self, x):
        return self.decoder(x)

    def forward(self, x):
        return self.decode(self._encode(x)[0])

    def evaluate(self, batch):
        return self(batch["signal"])

    def step(self, batch, stage="training"):  # pragma: no cover
        raise NotImplementedError("tqdne_b200 accelerates sampling; train with the reference tqdne package")

    def training_step(self, batch, batch_idx):  # pragma: no cover
        return self.step(batch)

    validation_step = training_step
