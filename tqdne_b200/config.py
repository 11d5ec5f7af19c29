"""Shapes of the named experiments (reference: experiments/config.py:7-75, tqdne/generate_waveforms.py:25-41)."""

from __future__ import annotations

from dataclasses import dataclass

from . import representation

FEATURES = ("hypocentral_distance", "magnitude", "vs30", "hypocentre_depth", "azimuthal_gap")


@dataclass
class SpectrogramConfig:
    """128 x 128 log-spectrogram of 4064-sample, 3-component waveforms (experiments/config.py:34-42)."""

    features_keys: tuple = FEATURES
    channels: int = 3
    fs: int = 100
    stft_channels: int = 256
    hop_size: int = 32
    t: int = 4096 - 32

    def __post_init__(self):
        self.representation = representation.LogSpectrogram(stft_channels=self.stft_channels, hop_size=self.hop_size)


@dataclass
class LatentSpectrogramConfig(SpectrogramConfig):
    """HighFEM latent EDM (experiments/config.py:45-50)."""

    latent_channels: int = 8
    kl_weight: float = 1e-6


@dataclass
class MovingAverageEnvelopeConfig:
    """1D EDM on [3 scaled-signal + 3 log-envelope] channels (experiments/config.py:61-67)."""

    features_keys: tuple = FEATURES
    channels: int = 6
    fs: int = 100
    t: int = 4064

    def __post_init__(self):
        self.representation = representation.MovingAverageEnvelope()
