"""Signal <-> representation maps; the inverses run on the GPU.

Drop-in for tqdne/representation.py: `LogSpectrogram(stft_channels, hop_size, clip, log_max, library,
multiprocessing)`, `MovingAverageEnvelope(window_size, log_eps, eps)`, `Identity`, `Normalization` with
`get_representation` / `invert_representation` accepting torch tensors or NumPy arrays and returning NumPy.

`LogSpectrogram.invert_representation` (reference: representation.py:152-175 -> librosa.griffinlim in a pathos
process pool) is ONE CUDA launch for the whole batch (csrc/tq_griffinlim.cu); a CUDA tensor input is consumed in
place, so only the waveforms cross PCIe.  The forward maps (`get_representation`) are device
kernels too (tq_logspec_forward, tq_mavg_envelope_forward): inputs are taken as fp32 (the dataset's storage type),
the arithmetic precision of the STFT follows `precision`.
"""

from __future__ import annotations

import math

import numpy as np
import torch

from . import _lib
from ._lib import TQ_F32, TQ_F64
from .engine import current_stream_ptr, device_guard


def _as_cuda_f32(x, device=None) -> torch.Tensor:
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
    if not isinstance(x, torch.Tensor):
        x = torch.as_tensor(x)
    if not x.is_cuda:
        if not torch.cuda.is_available():
            raise RuntimeError("tqdne_b200: representation inverses run on a CUDA device; none is available "
                               "(there is no CPU fallback)")
        x = x.to(device or "cuda")
    return x.to(torch.float32).contiguous()


def _on_input_device(fn):
    """Run a *_device method with the input's CUDA device current (streams and launches follow the current device)."""
    import functools

    @functools.wraps(fn)
    def wrapper(self, x, *a, **k):
        x = _as_cuda_f32(x)
        with device_guard(x.device):
            return fn(self, x, *a, **k)
    return wrapper


class Representation:
    """Abstract representation (reference: representation.py:9-19)."""

    def get_representation(self, waveform):
        raise NotImplementedError

    def invert_representation(self, representation):
        raise NotImplementedError


class Identity(Representation):
    def get_representation(self, waveform):
        return _np(waveform)

    def invert_representation(self, representation):
        return _np(representation)


def _np(x):
    return x.numpy(force=True) if isinstance(x, torch.Tensor) else np.asarray(x)


class Normalization(Representation):
    def __init__(self, mean, std):
        self.mean, self.std = mean, std

    def get_representation(self, waveform):
        return (_np(waveform) - self.mean) / self.std

    def invert_representation(self, representation):
        return _np(representation) * self.std + self.mean


class MovingAverageEnvelope(Representation):
    """Scaled signal + log moving-average envelope (reference: representation.py:41-60)."""

    def __init__(self, window_size=128, log_eps=1e-6, eps=1e-6):
        self.window_size, self.log_eps, self.eps = window_size, log_eps, eps

    @_on_input_device
    def get_representation_device(self, waveform) -> torch.Tensor:
        """[.., channels, L] -> [.., 2*channels, L] on the device (reference: representation.py:47-55)."""
        w = _as_cuda_f32(waveform)
        lead, (cw, L) = w.shape[:-2], w.shape[-2:]
        n = int(math.prod(lead)) if lead else 1
        out = torch.empty(*lead, 2 * cw, L, device=w.device, dtype=torch.float32)
        _lib.check(_lib.lib().tq_mavg_envelope_forward(w.data_ptr(), out.data_ptr(), n, cw, L, self.window_size,
                                                       self.log_eps, self.eps, current_stream_ptr()), "mavg_envelope_forward")
        return out

    def get_representation(self, waveform):
        return self.get_representation_device(waveform).cpu().numpy()

    @_on_input_device
    def invert_representation_device(self, representation) -> torch.Tensor:
        rep = _as_cuda_f32(representation)
        lead, (c2, L) = rep.shape[:-2], rep.shape[-2:]
        assert c2 % 2 == 0, "expected [.., 2*channels, L]"
        n = int(math.prod(lead)) if lead else 1
        out = torch.empty(*lead, c2 // 2, L, device=rep.device, dtype=torch.float32)
        _lib.check(_lib.lib().tq_mavg_envelope_inverse(rep.data_ptr(), out.data_ptr(), n, c2 // 2, L, self.log_eps,
                                                       self.eps, current_stream_ptr()), "mavg_envelope_inverse")
        return out

    def invert_representation(self, representation):
        return self.invert_representation_device(representation).cpu().numpy()


class LogSpectrogram(Representation):
    """Log-magnitude STFT in [-1, 1]; inverse = exp + fast Griffin-Lim (reference: representation.py:63-175).

    Extra keyword (engine extension): `precision` = "fp64" (default) or "fp32" arithmetic for Griffin-Lim.  The
    reference's locked environment (uv.lock: NumPy 2.2.5) promotes the float32 spectrogram to float64 before librosa, so
    the reference runs complex128: fp64 is the parity mode (kernel = NumPy oracle to 1e-9).  128 momentum iterations
    amplify rounding ~1e3x (NumPy's own complex64 run differs from its complex128 run by 1e-4 .. 1e-3 on decoded
    spectrograms, worst items 1e-1), so "fp32" is an explicit opt-in fast mode, not a parity mode.
    `library` and `multiprocessing` are accepted for signature compatibility and ignored: there is no
    librosa / process pool here.
    """

    n_iter = 128          # librosa.griffinlim(n_iter=128, ...) representation.py:106-108
    momentum = 0.99       # librosa default
    random_state = 0

    def __init__(self, stft_channels=256, hop_size=None, clip=1e-8, log_max=3, library="librosa", multiprocessing=True,
                 precision="fp64"):
        self.clip = clip
        self.log_clip = np.log(clip)
        self.log_max = log_max
        self.library = library
        self.stft_channels = stft_channels
        self.hop_size = stft_channels // 4 if hop_size is None else hop_size
        if precision not in ("fp32", "fp64"):
            raise ValueError("precision must be 'fp32' or 'fp64'")
        self.precision = precision
        self._phase = {}
        self._ws = None
        self.max_items_per_launch = 3072

    def disable_multiprocessing(self):
        """Kept for API compatibility (reference: representation.py:135-138); nothing to close."""

    # ---- forward (the step before the sampling path) ---------------------------------------------------
    @_on_input_device
    def get_representation_device(self, waveform) -> torch.Tensor:
        """waveforms [.., L] -> normalised log-magnitude STFT [.., n_fft/2, 1 + L/hop] in [-1, 1] as a CUDA tensor
        (reference: get_spectrogram + get_representation, representation.py:140-150,163-169; librosa stft semantics)."""
        w = _as_cuda_f32(waveform)
        lead, L = w.shape[:-1], w.shape[-1]
        items = int(math.prod(lead)) if lead else 1
        frames = 1 + L // self.hop_size
        prec = TQ_F64 if self.precision == "fp64" else TQ_F32
        odt = torch.float64 if prec == TQ_F64 else torch.float32
        out = torch.empty(items, self.stft_channels // 2, frames, device=w.device, dtype=odt)
        _lib.check(_lib.lib().tq_logspec_forward(w.reshape(items, L).data_ptr(), out.data_ptr(), prec, items, self.stft_channels,
                                                 self.hop_size, L, frames, float(self.clip), float(self.log_max), prec,
                                                 current_stream_ptr()), "logspec_forward")
        return out.reshape(*lead, self.stft_channels // 2, frames)

    def get_representation(self, waveform):
        return self.get_representation_device(waveform).cpu().numpy()

    def get_spectrogram(self, waveform):
        raise NotImplementedError("tqdne_b200: the complex spectrogram is never materialised -- get_representation() goes "
                                  "from waveforms to the normalised log-magnitudes in one kernel (tq_logspec_forward)")

    # ---- inverse (hot path) ----------------------------------------------------------------------------
    def _phase0(self, frames: int, device) -> torch.Tensor:
        key = (frames, str(device))
        if key not in self._phase:
            rng = np.random.RandomState(seed=self.random_state)
            ph = 2 * np.pi * rng.random(size=(self.stft_channels // 2 + 1, frames))
            # unit phasors exp(i ph) as (cos, sin) pairs in float64, like librosa's util.phasor
            self._phase[key] = torch.from_numpy(np.stack([np.cos(ph), np.sin(ph)], axis=-1)).contiguous().to(device)
        return self._phase[key]

    @_on_input_device
    def invert_representation_device(self, representation) -> torch.Tensor:
        """[.., n_fft/2, frames] in [-1, 1] -> waveforms [.., hop*(frames-1)] as a CUDA tensor."""
        rep = _as_cuda_f32(representation)
        lead, (nb, frames) = rep.shape[:-2], rep.shape[-2:]
        if nb != self.stft_channels // 2:
            raise ValueError(f"expected {self.stft_channels // 2} frequency rows, got {nb}")
        items = int(math.prod(lead)) if lead else 1
        lib = _lib.lib()
        prec = TQ_F64 if self.precision == "fp64" else TQ_F32
        odt = torch.float64 if prec == TQ_F64 else torch.float32
        out_len = self.hop_size * (frames - 1)
        out = torch.empty(items, out_len, device=rep.device, dtype=odt)
        rep2 = rep.reshape(items, nb, frames)
        phase = self._phase0(frames, rep.device)
        step = min(items, self.max_items_per_launch)
        need = lib.tq_griffinlim_ws_bytes(step, self.stft_channels, frames, prec)
        if need < 0:
            raise RuntimeError("tqdne_b200: unsupported STFT size for the Griffin-Lim kernel")
        if self._ws is None or self._ws.numel() < need or self._ws.device != rep.device:
            self._ws = torch.empty(need, device=rep.device, dtype=torch.uint8)
        st = current_stream_ptr()
        for i0 in range(0, items, step):
            n = min(step, items - i0)
            _lib.check(lib.tq_logspec_griffinlim(rep2[i0:i0 + n].data_ptr(), phase.data_ptr(), out[i0:i0 + n].data_ptr(), n,
                                                 self.stft_channels, self.hop_size, frames, self.n_iter,
                                                 float(self.log_clip), float(self.log_max), self.momentum, prec,
                                                 self._ws.data_ptr(), st), "logspec_griffinlim")
        return out.reshape(*lead, out_len)

    def invert_representation(self, representation):
        return self.invert_representation_device(representation).cpu().numpy()

    def invert_spectrogram(self, spec):
        """Magnitudes [.., n_fft/2, frames] -> waveforms (reference: representation.py:152-161)."""
        s = torch.as_tensor(_np(spec)) if not isinstance(spec, torch.Tensor) else spec
        log_spec = torch.log(torch.clamp(s.to(torch.float64), min=1e-300))
        rep = (log_spec - self.log_clip) / (self.log_max - self.log_clip) * 2 - 1
        return self.invert_representation(rep.to(torch.float32))
