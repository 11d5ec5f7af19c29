"""Golden vectors for the forward representations -- TEST INFRASTRUCTURE (see oracle/__init__.py).

Runs the UNMODIFIED reference classes from /root/reference (build container only):
  tqdne.representation.MovingAverageEnvelope.get_representation           (representation.py:47-55)
  tqdne.representation.LogSpectrogram.get_spectrogram / get_representation (representation.py:140-150,163-169)
with oracle.griffinlim_ref.stft standing in for librosa.stft (librosa is absent: PARITY UNPINNED at that boundary,
see oracle/griffinlim_ref.py), and stores inputs / outputs under tests/golden/.

    python -m oracle.make_golden_forward
"""
from pathlib import Path

import numpy as np

from oracle import griffinlim_ref, reference_loader

OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"


def main():
    reference_loader.load()
    import tqdne.representation as rrep

    rng = np.random.RandomState(61)
    # seismogram-like: decaying noise bursts, 3 components, reference length 4064
    t = np.arange(4064)
    env = np.exp(-((t - 900.0) / 600.0) ** 2)[None, None, :] + 1e-3
    wave = (rng.randn(2, 3, 4064) * env).astype(np.float32)
    rep = rrep.MovingAverageEnvelope().get_representation(wave)
    np.savez_compressed(OUT / "mavg_forward.npz", wave=wave, rep=rep.astype(np.float64))

    ls = object.__new__(rrep.LogSpectrogram)   # the ctor imports librosa; set what it would set
    ls.clip, ls.log_clip, ls.log_max, ls.library = 1e-8, np.log(1e-8), 3, "librosa"
    ls.stft = lambda x: griffinlim_ref.stft(x, 256, 32)
    w1 = wave[:1]                                             # one sample keeps the fixture under 1 MB
    rep32 = ls.get_representation(w1)                         # float32 in -> complex64 STFT
    rep64 = ls.get_representation(w1.astype(np.float64))      # float64 in -> complex128 STFT
    np.savez_compressed(OUT / "logspec_forward.npz", wave=w1, rep32=rep32.astype(np.float32), rep64=rep64)
    print("mavg", rep.shape, rep.dtype, "logspec", rep32.shape, rep32.dtype, rep64.dtype)


if __name__ == "__main__":
    main()
