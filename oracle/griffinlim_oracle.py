"""CPU restatement of the reference's Griffin-Lim vocoder path (seq2seq_vc/vocoder/griffin_lim.py:20-106).

TEST INFRASTRUCTURE ONLY.  The reference delegates the arithmetic to **librosa** (``librosa.filters.mel`` for the inverse mel
basis, ``librosa.griffinlim`` -> ``librosa.stft`` / ``librosa.istft`` for the phase reconstruction; setup.cfg:5, unpinned, not
installable here), so this oracle is **PARITY UNPINNED** against librosa itself: it restates librosa's published algorithm for the
reference's call (griffin_lim.py:79-87: ``librosa.griffinlim(S, n_iter, hop_length, win_length, window, center=True)``, i.e. the
fast Griffin-Lim of Perraudin et al. with momentum 0.99, random initial phases, ``pad_mode="constant"`` (librosa >= 0.9)):

    angles = exp(2 pi i U)                                   U ~ uniform[0, 1)   (here: passed in explicitly)
    repeat n_iter times:
        inverse = istft(S * angles)                          window-weighted overlap-add / sum of squared windows, centre trimmed
        rebuilt = stft(inverse)                              zero ("constant") padding of n_fft // 2 on both sides
        angles  = rebuilt - momentum / (1 + momentum) * previous rebuilt
        angles /= |angles| + tiny
    return istft(S * angles)

The STFT / ISTFT pair is cross-checked against scipy.signal (tests/test_griffinlim_oracle.py); the random initial phases of
librosa are replaced by an explicit array so that the CUDA path can be compared element by element.
"""
from __future__ import annotations

import numpy as np

from oracle.logmel_oracle import hann_padded, mel_basis

EPS = 1e-10


def logmel2linear(lmspc: np.ndarray, fs: int, n_fft: int, n_mels: int, fmin=None, fmax=None) -> np.ndarray:
    """griffin_lim.py:20-50: max(EPS, pinv(mel_basis) @ 10 ** lmspc)."""
    assert lmspc.shape[1] == n_mels
    fmin = 0 if fmin is None else fmin
    fmax = fs / 2 if fmax is None else fmax
    mspc = np.power(10.0, lmspc)
    inv = np.linalg.pinv(mel_basis(fs, n_fft, n_mels, fmin, fmax))
    return np.maximum(EPS, np.dot(inv, mspc.T).T)


def stft(y: np.ndarray, n_fft: int, hop: int, win_length=None, pad_mode: str = "constant") -> np.ndarray:
    """librosa.stft(center=True): (1 + len(y) // hop, 1 + n_fft // 2) complex128, frames along axis 0."""
    win = hann_padded(n_fft, win_length)
    pad = n_fft // 2
    yp = np.pad(np.asarray(y, dtype=np.float64), (pad, pad), mode=pad_mode)
    n_frames = 1 + (len(yp) - n_fft) // hop
    idx = np.arange(n_fft)[None, :] + hop * np.arange(n_frames)[:, None]
    return np.fft.rfft(yp[idx] * win[None, :], axis=-1)


def istft(spec: np.ndarray, n_fft: int, hop: int, win_length=None) -> np.ndarray:
    """librosa.istft(center=True, length=None): inverse rFFT of every frame, synthesis window = analysis window, overlap-add,
    division by the sum of squared windows where it exceeds tiny, n_fft // 2 samples trimmed on both sides."""
    win = hann_padded(n_fft, win_length)
    T = spec.shape[0]
    frames = np.fft.irfft(spec, n=n_fft, axis=-1) * win[None, :]
    total = n_fft + hop * (T - 1)
    y = np.zeros(total)
    wss = np.zeros(total)
    for t in range(T):
        y[t * hop:t * hop + n_fft] += frames[t]
        wss[t * hop:t * hop + n_fft] += win ** 2
    nz = wss > np.finfo(np.float32).tiny
    y[nz] /= wss[nz]
    return y[n_fft // 2: total - n_fft // 2]


def griffin_lim(spc: np.ndarray, n_fft: int, n_shift: int, win_length=None, n_iter: int = 32, init_angles: np.ndarray | None = None,
                momentum: float = 0.99, pad_mode: str = "constant", seed: int = 0) -> np.ndarray:
    """spc (T, n_fft // 2 + 1) magnitudes -> waveform (n_shift * (T - 1),); init_angles (T, bins) complex unit phases."""
    S = np.abs(np.asarray(spc, dtype=np.float64))
    assert S.shape[1] == n_fft // 2 + 1
    if init_angles is None:
        init_angles = np.exp(2j * np.pi * np.random.default_rng(seed).random(S.shape))
    angles = np.asarray(init_angles, dtype=np.complex128).copy()
    eps = np.finfo(np.float32).tiny
    tprev = np.zeros_like(angles)
    for _ in range(n_iter):
        inverse = istft(S * angles, n_fft, n_shift, win_length)
        rebuilt = stft(inverse, n_fft, n_shift, win_length, pad_mode)
        angles = rebuilt - (momentum / (1 + momentum)) * tprev
        angles /= np.abs(angles) + eps
        tprev = rebuilt
    return istft(S * angles, n_fft, n_shift, win_length)
