"""NumPy restatement of the log-spectrogram inverse -- TEST INFRASTRUCTURE (see oracle/__init__.py).

PARITY UNPINNED at the librosa boundary (narrowed, see below): the reference calls librosa 0.11.0 (`uv.lock`, pyproject.toml:16)
`griffinlim(S, hop_length=32, n_fft=256, n_iter=128, random_state=0)` (tqdne/representation.py:103-108);
librosa is a third-party dependency that is neither vendored under /root/reference nor installed here, and the
reference has no test pinning its output.  This file restates librosa 0.11's published algorithm:

  stft   : center=True with zero ("constant") padding of n_fft//2, periodic Hann window of n_fft, hop 32,
           rfft of each frame                                   -> [1 + n_fft/2, 1 + len/hop]
  istft  : irfft(frame) * window, overlap-add, divide by sum of squared windows where > tiny, trim n_fft//2
  griffinlim (fast, momentum 0.99, init='random'):
           angles = exp(2 pi i U),  U = RandomState(seed).random(S.shape);  angles *= S
           loop: inverse = istft(angles); rebuilt = stft(inverse); angles = rebuilt - m/(1+m) * tprev;
                 angles /= |angles| + tiny; angles *= S; tprev = rebuilt
           return istft(angles)

The STFT pair is pinned against torch.stft / torch.istft in tests/test_oracle.py, and the Griffin-Lim LOOP
(momentum update, normalisation, stft/istft chaining) is pinned against torchaudio.functional.griffinlim -- an
independent port of librosa's loop that is installed here -- in the configuration both can express (reflect padding,
all-ones initial phase, eps 1e-16).  What remains unpinned is exactly librosa's three deviations from that port:
zero padding, the RandomState(0) unit-phasor initial phase, eps = finfo.tiny.  The surrounding reference
code (un-normalise, exp, append zero Nyquist row; representation.py:152-175) is pinned by executing the
reference's own LogSpectrogram methods with this griffinlim plugged in (oracle/make_golden.py).
"""

from __future__ import annotations

import numpy as np


def hann_periodic(n: int, dtype=np.float64):
    return (0.5 - 0.5 * np.cos(2 * np.pi * np.arange(n) / n)).astype(dtype)


def stft(y: np.ndarray, n_fft: int = 256, hop: int = 32, pad_mode: str = "constant") -> np.ndarray:
    win = hann_periodic(n_fft, y.dtype)
    yp = np.pad(y, n_fft // 2, mode=pad_mode)
    n_frames = 1 + (len(yp) - n_fft) // hop
    idx = np.arange(n_fft)[None, :] + hop * np.arange(n_frames)[:, None]
    frames = yp[idx] * win[None, :]
    return np.fft.rfft(frames, axis=1).T.astype(np.complex128 if y.dtype == np.float64 else np.complex64)


def istft(X: np.ndarray, n_fft: int = 256, hop: int = 32) -> np.ndarray:
    rdt = np.float64 if X.dtype == np.complex128 else np.float32
    win = hann_periodic(n_fft, rdt)
    n_frames = X.shape[1]
    length = n_fft + hop * (n_frames - 1)
    frames = np.fft.irfft(X.T, n=n_fft, axis=1).astype(rdt) * win[None, :]
    y = np.zeros(length, dtype=rdt)
    wss = np.zeros(length, dtype=rdt)
    w2 = win * win
    for t in range(n_frames):
        y[t * hop:t * hop + n_fft] += frames[t]
        wss[t * hop:t * hop + n_fft] += w2
    nz = wss > np.finfo(rdt).tiny
    y[nz] /= wss[nz]
    return y[n_fft // 2: length - n_fft // 2]


def initial_phase(n_bins: int, n_frames: int, seed: int = 0) -> np.ndarray:
    """2*pi*U with U = RandomState(seed).random((n_bins, n_frames)) -- the same for every item."""
    return 2 * np.pi * np.random.RandomState(seed=seed).random(size=(n_bins, n_frames))


def griffinlim(S: np.ndarray, n_iter: int = 128, hop: int = 32, n_fft: int = 256, momentum: float = 0.99,
               seed: int = 0, pad_mode: str = "constant", init: str = "random", eps: float | None = None) -> np.ndarray:
    """S: magnitudes [1 + n_fft/2, frames], float32 or float64 (sets the arithmetic precision).

    The defaults are librosa 0.11's (zero padding, unit phasors from RandomState(seed), eps = tiny).  The three
    keyword deviations exist so that the SAME loop can be pinned against an independent implementation that is
    installed here: torchaudio.functional.griffinlim = (pad_mode="reflect", init="ones", eps=1e-16), see
    tests/test_oracle.py::test_griffinlim_loop_pinned_against_torchaudio."""
    cdt = np.complex128 if S.dtype == np.float64 else np.complex64
    eps = np.finfo(S.dtype).tiny if eps is None else eps
    if init == "random":
        ph = initial_phase(*S.shape, seed=seed)
        angles = (np.cos(ph) + 1j * np.sin(ph)).astype(cdt)
    else:
        angles = np.ones(S.shape, dtype=cdt)
    angles *= S
    tprev = None
    for _ in range(n_iter):
        inverse = istft(angles, n_fft, hop)
        rebuilt = stft(inverse, n_fft, hop, pad_mode)
        angles[:] = rebuilt
        if tprev is not None:
            angles -= (momentum / (1 + momentum)) * tprev
        angles /= np.abs(angles) + eps
        angles *= S
        tprev = rebuilt
    return istft(angles, n_fft, hop)


def logspec_inverse(rep: np.ndarray, clip: float = 1e-8, log_max: float = 3, n_fft: int = 256, hop: int = 32,
                    n_iter: int = 128, precision: str = "fp64") -> np.ndarray:
    """LogSpectrogram.invert_representation (tqdne/representation.py:152-175): rep [.., n_fft/2, frames] float32."""
    rep = np.asarray(rep, dtype=np.float32)
    log_clip = np.log(clip)
    norm = (rep + 1) / 2                                  # float32
    log_spec = norm.astype(np.float64) * (log_max - log_clip) + log_clip
    spec = np.exp(log_spec)
    if precision == "fp32":
        spec = spec.astype(np.float32)
    shape = spec.shape
    flat = spec.reshape(-1, shape[-2], shape[-1])
    flat = np.concatenate([flat, np.zeros_like(flat[:, :1])], axis=1)  # zero Nyquist row
    out = np.array([griffinlim(s, n_iter=n_iter, hop=hop, n_fft=n_fft) for s in flat])
    return out.reshape(shape[:-2] + out.shape[1:])


# ---- forward representations (tqdne/representation.py:47-55 and :140-150,163-169) --------------------------------
def logspec_forward(wave: np.ndarray, clip: float = 1e-8, log_max: float = 3, n_fft: int = 256, hop: int = 32) -> np.ndarray:
    """LogSpectrogram.get_representation with librosa.stft restated by `stft` above: wave [.., L] -> [.., n_fft/2, frames].
    The arithmetic precision follows wave.dtype, like librosa (float32 in -> complex64)."""
    wave = np.asarray(wave)
    shape = wave.shape
    flat = wave.reshape(-1, shape[-1])
    spec = np.array([stft(x, n_fft, hop) for x in flat])[:, :-1]      # drop the Nyquist row
    spec = np.abs(spec.reshape(shape[:-1] + spec.shape[1:]))
    log_clip = np.log(clip)
    log_spec = np.log(np.clip(spec, clip, None))
    return (log_spec - log_clip) / (log_max - log_clip) * 2 - 1


def mavg_forward(wave: np.ndarray, window: int = 128, log_eps: float = 1e-6, eps: float = 1e-6) -> np.ndarray:
    """MovingAverageEnvelope.get_representation: [.., C, L] -> [.., 2C, L]."""
    wave = np.asarray(wave)
    env = np.apply_along_axis(lambda x: np.convolve(x, np.ones(window) / window, mode="same"), axis=-1, arr=np.abs(wave))
    return np.concatenate([wave / (env + eps), np.log(env + log_eps) - np.log(log_eps) / 2], axis=-2)


# ---- a line-by-line NumPy model of the CUDA kernel's FFT decomposition (csrc/tq_griffinlim.cu) -----------
def kernel_model_rfft256(x: np.ndarray) -> np.ndarray:
    """rfft-256 via one 128-point complex FFT + even/odd split, as the kernel computes it."""
    z = x[0::2] + 1j * x[1::2]
    Z = np.fft.fft(z)
    k = np.arange(129)
    zk, zn = Z[k % 128], Z[(128 - k) % 128]
    er, ei = 0.5 * (zk.real + zn.real), 0.5 * (zk.imag - zn.imag)
    dr, di = 0.5 * (zk.real - zn.real), 0.5 * (zk.imag + zn.imag)
    c, s = np.cos(-2 * np.pi * k / 256), np.sin(-2 * np.pi * k / 256)
    return (er + (di * c + dr * s)) + 1j * (ei + (di * s - dr * c))


def kernel_model_irfft256(X: np.ndarray) -> np.ndarray:
    """irfft-256 via pre-twiddle + one 128-point complex inverse FFT, as the kernel computes it."""
    k = np.arange(128)
    xk, xn = X[k].copy(), X[128 - k].copy()
    xk[0] = xk[0].real
    xn[0] = xn[0].real
    er, ei = 0.5 * (xk.real + xn.real), 0.5 * (xk.imag - xn.imag)
    dr, di = 0.5 * (xk.real - xn.real), 0.5 * (xk.imag + xn.imag)
    c, s = np.cos(2 * np.pi * k / 256), np.sin(2 * np.pi * k / 256)
    orr, oi = dr * c - di * s, dr * s + di * c
    Z = (er - oi) + 1j * (ei + orr)
    z = np.fft.ifft(Z)  # includes the 1/128
    x = np.empty(256)
    x[0::2], x[1::2] = z.real, z.imag
    return x


# ---- lane / register model of the fused kernel's 128-point FFT network (csrc/tq_griffinlim.cu, namespace fused) ----
def _radix4(z: np.ndarray, inv: bool) -> np.ndarray:
    t0, t1, t2, t3 = z[:, 0] + z[:, 2], z[:, 0] - z[:, 2], z[:, 1] + z[:, 3], z[:, 1] - z[:, 3]
    it3 = (1j if inv else -1j) * t3
    out = np.empty_like(z)
    out[:, 0], out[:, 2], out[:, 1], out[:, 3] = t0 + t2, t0 - t2, t1 + it3, t1 - it3
    return out


def kernel_model_fft128_dif(x: np.ndarray) -> np.ndarray:
    """fft128<R, false>: x natural -> z[lane, p] = X[(lane & 15) + 16 p + 64 (lane >> 4)] (32 lanes x 4 registers)."""
    L = np.arange(32)
    W = lambda e: np.exp(-2j * np.pi * e / 128)  # noqa: E731
    a, l2, g, l3 = L >> 3, L & 7, L & 15, L >> 4
    z = _radix4(np.stack([x[L + 32 * j] for j in range(4)], 1).astype(complex), False)
    for k in range(1, 4):
        z[:, k] *= W(L * k)
    ex = np.zeros(152, complex)
    for k in range(4):
        ex[k * 40 + L] = z[:, k]
    z = _radix4(np.stack([ex[a * 40 + l2 + 8 * j] for j in range(4)], 1), False)
    for m in range(1, 4):
        z[:, m] *= W(4 * l2 * m)
    ex = np.zeros(152, complex)
    for m in range(4):
        ex[9 * (a + 4 * m) + l2] = z[:, m]
    z = _radix4(np.stack([ex[9 * g + l3 + 2 * j] for j in range(4)], 1), False)
    for q in range(1, 4):
        z[l3 == 1, q] *= W(16 * q)
    other = z[L ^ 16]
    return np.where((l3 == 1)[:, None], other - z, z + other)
