"""The STFT -> log-mel oracle (oracle/logmel_oracle.py) against two INDEPENDENT implementations of the same published
algorithm.  The reference delegates this arithmetic to librosa (bin/preprocess.py:63-82), which is neither vendored in the
reference tree nor installable here, so the oracle stays "parity unpinned" against librosa itself; these checks bound the risk:
  * scipy.signal.stft (boundary='even' == reflect padding, periodic Hann, un-normalised) for the STFT magnitudes;
  * transformers.audio_utils (mel_filter_bank / spectrogram, written to reproduce librosa's slaney filterbank and centred
    STFT) for the filterbank and the whole log-mel pipeline.
Tolerance: 1e-5 in the log10 domain (float32 magnitudes in the oracle vs float64 in the checkers)."""
import numpy as np
import pytest

from oracle import logmel_oracle as lo

CASES = [  # (sr, n_fft, hop, win_length, fmin, fmax)
    (24000, 1024, 256, None, 80, 7600),     # egs/arctic/vc1/conf/vtn.v1.yaml feature settings
    (48000, 2048, 300, None, 0, None),      # BASELINE configs[4] (C5)
    (16000, 1024, 160, 800, 0, 8000),       # win_length < n_fft: centrally zero-padded window
]


def _wave(sr, seed, seconds=0.5):
    rng = np.random.default_rng(seed)
    t = np.arange(int(sr * seconds)) / sr
    x = 0.1 * rng.standard_normal(t.size) + 0.3 * np.sin(2 * np.pi * 440.0 * t) + 0.2 * np.sin(2 * np.pi * 3000.0 * t)
    return np.clip(x, -1, 1).astype(np.float32)


@pytest.mark.parametrize("sr,n_fft,hop,win,fmin,fmax", CASES)
def test_stft_magnitude_vs_scipy(sr, n_fft, hop, win, fmin, fmax):
    from scipy import signal

    x = _wave(sr, 1)
    ours = lo.logmelfilterbank(x, sr, n_fft, hop, win, "hann", 80, fmin, fmax)
    wl = n_fft if win is None else win
    w = np.zeros(n_fft)
    w[(n_fft - wl) // 2:(n_fft - wl) // 2 + wl] = signal.get_window("hann", wl, fftbins=True)
    _, _, Z = signal.stft(x.astype(np.float64), fs=sr, window=w, nperseg=n_fft, noverlap=n_fft - hop, nfft=n_fft, boundary="even",
                          padded=False, return_onesided=True)
    mag = np.abs(Z.T) * w.sum()                                       # scipy divides by sum(window)
    basis = lo.mel_basis(sr, n_fft, 80, fmin, sr / 2 if fmax is None else fmax).astype(np.float64)
    ref = np.log10(np.maximum(1e-10, mag @ basis.T))
    assert ours.shape == (1 + len(x) // hop, 80) and ours.dtype == np.float32
    n = min(len(ref), len(ours))
    assert n >= len(ours) - 1
    assert np.abs(ref[:n] - ours[:n]).max() <= 1e-5


@pytest.mark.parametrize("sr,n_fft,hop,win,fmin,fmax", CASES[:2])
def test_logmel_vs_transformers_audio_utils(sr, n_fft, hop, win, fmin, fmax):
    au = pytest.importorskip("transformers.audio_utils")
    x = _wave(sr, 2)
    fmax_ = sr / 2 if fmax is None else fmax
    mf = au.mel_filter_bank(num_frequency_bins=n_fft // 2 + 1, num_mel_filters=80, min_frequency=fmin, max_frequency=fmax_,
                            sampling_rate=sr, norm="slaney", mel_scale="slaney")
    assert np.abs(mf.T - lo.mel_basis(sr, n_fft, 80, fmin, fmax_)).max() <= 1e-7
    sp = au.spectrogram(x.astype(np.float64), au.window_function(n_fft, "hann", periodic=True), frame_length=n_fft, hop_length=hop,
                        fft_length=n_fft, power=1.0, center=True, pad_mode="reflect", mel_filters=mf, mel_floor=1e-10, log_mel="log10")
    ours = lo.logmelfilterbank(x, sr, n_fft, hop, win, "hann", 80, fmin, fmax)
    n = min(sp.shape[1], len(ours))
    assert n >= len(ours) - 1
    assert np.abs(sp.T[:n] - ours[:n]).max() <= 1e-5


def test_log_bases_and_floor():
    x = np.zeros(4000, dtype=np.float32)                              # silence: every bin sits on the eps floor
    for base, f in ((10.0, np.log10), (2.0, np.log2), (None, np.log)):
        out = lo.logmelfilterbank(x, 16000, 512, 128, None, "hann", 80, 0, 8000, eps=1e-10, log_base=base)
        assert np.allclose(out, f(np.float32(1e-10)))
    with pytest.raises(ValueError):
        lo.logmelfilterbank(x, 16000, 512, 128, log_base=3.0)
