"""CPU restatement of the reference STFT -> log-mel front end.

TEST INFRASTRUCTURE ONLY.  Restates ``logmelfilterbank`` (seq2seq_vc/bin/preprocess.py:30-92).
The arithmetic lives in a third-party dependency that is absent from /root/reference and not
installable here: **librosa** (setup.cfg:5, ``librosa >= 0.8.0``, unpinned).  The reference has no
test or golden vector at that boundary, so this oracle is **PARITY UNPINNED** against librosa
itself; it follows librosa's published algorithm for the call sites preprocess.py:63-70
(``librosa.stft(..., window="hann", pad_mode="reflect")``, center=True default) and :76-82
(``librosa.filters.mel``: Slaney mel scale, Slaney area normalisation, float32).  The risk is bounded by
cross-checks against independent implementations of the same algorithm (tests/test_logmel_oracle.py,
tests/test_oracle_golden.py): scipy.signal.stft for the STFT magnitudes, transformers.audio_utils
(mel_filter_bank / spectrogram, written to reproduce librosa) for the filterbank and the whole log-mel
pipeline (<= 1e-5 in the log10 domain), torchaudio for the filterbank.

librosa semantics restated:
 * reflect-pad n_fft//2 samples on both sides; frames = 1 + len(audio)//hop;
 * window = periodic Hann of win_length (scipy get_window(..., fftbins=True)), zero-padded
   centrally to n_fft; float64 window * float32 frame -> float64 rfft -> cast complex64;
 * magnitude float32; mel = float32 GEMM with the float32 basis; max(eps, .); log10.
"""
from __future__ import annotations

import numpy as np


def hz_to_mel_slaney(f):
    f = np.asanyarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    with np.errstate(divide="ignore"):
        return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-300) / min_log_hz) / logstep, mels)


def mel_to_hz_slaney(m):
    m = np.asanyarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_basis(sr: int, n_fft: int, n_mels: int = 80, fmin: float = 0.0, fmax: float | None = None) -> np.ndarray:
    """librosa.filters.mel(htk=False, norm='slaney', dtype=float32) -> (n_mels, 1 + n_fft//2)."""
    fmax = sr / 2.0 if fmax is None else fmax
    fftfreqs = np.fft.rfftfreq(n_fft, 1.0 / sr)
    mel_f = mel_to_hz_slaney(np.linspace(hz_to_mel_slaney(fmin), hz_to_mel_slaney(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    w = np.zeros((n_mels, fftfreqs.shape[0]), dtype=np.float64)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        w[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2 : n_mels + 2] - mel_f[:n_mels])
    w *= enorm[:, None]
    return w.astype(np.float32)


def hann_padded(n_fft: int, win_length: int | None = None) -> np.ndarray:
    win_length = n_fft if win_length is None else win_length
    n = np.arange(win_length, dtype=np.float64)
    w = 0.5 - 0.5 * np.cos(2.0 * np.pi * n / win_length)  # periodic (fftbins=True)
    lpad = (n_fft - win_length) // 2
    out = np.zeros(n_fft, dtype=np.float64)
    out[lpad : lpad + win_length] = w
    return out


def logmelfilterbank(audio, sampling_rate, fft_size=1024, hop_size=256, win_length=None, window="hann",
                     num_mels=80, fmin=None, fmax=None, eps=1e-10, log_base=10.0) -> np.ndarray:
    """Same signature/returns as the reference (preprocess.py:30-42): (frames, num_mels) float32."""
    if window != "hann":
        raise ValueError("oracle restates the hann window only")
    audio = np.asarray(audio, dtype=np.float32)
    pad = fft_size // 2
    y = np.pad(audio, (pad, pad), mode="reflect")
    n_frames = 1 + (len(y) - fft_size) // hop_size
    idx = np.arange(fft_size)[None, :] + hop_size * np.arange(n_frames)[:, None]
    frames = y[idx]                                   # (frames, n_fft) float32
    win = hann_padded(fft_size, win_length)
    spec = np.fft.rfft(win[None, :] * frames, axis=-1).astype(np.complex64)
    spc = np.abs(spec)                                # float32 (frames, bins)
    fmin = 0 if fmin is None else fmin
    fmax = sampling_rate / 2 if fmax is None else fmax
    basis = mel_basis(sampling_rate, fft_size, num_mels, fmin, fmax)
    mel = np.maximum(np.float32(eps), spc @ basis.T)
    if log_base is None:
        return np.log(mel)
    if log_base == 10.0:
        return np.log10(mel)
    if log_base == 2.0:
        return np.log2(mel)
    raise ValueError(f"{log_base} is not supported.")
