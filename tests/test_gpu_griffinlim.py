"""Griffin-Lim vocoder path on the GPU (seq2seq_vc_b200.griffin_lim / logmel2linear / Spectrogram2Waveform, through the C ABI) vs
the oracle (oracle/griffinlim_oracle.py: librosa's published algorithm, PARITY UNPINNED against librosa itself) on the same
initial phases."""
import numpy as np
import pytest
import torch

from oracle import griffinlim_oracle as glo
from oracle import logmel_oracle as lo

pytestmark = pytest.mark.gpu


def _signal(n, fs, seed):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / fs
    return (0.4 * np.sin(2 * np.pi * 220 * t) + 0.3 * np.sin(2 * np.pi * 1870 * t + 0.7) + 0.05 * rng.standard_normal(n)).astype(np.float32)


@pytest.mark.parametrize("n_fft,hop,win,pad", [(1024, 256, None, "constant"), (2048, 300, 1200, "constant"), (512, 128, 400, "reflect"),
                                               (4096, 1024, None, "constant")])
def test_stft_istft_kernels_vs_oracle(n_fft, hop, win, pad):
    from seq2seq_vc_b200 import api, ops

    T = 33
    x = _signal(hop * (T - 1), 24000, n_fft)
    w = torch.from_numpy(api.hann_window(n_fft, win)).cuda()
    spec = torch.empty(T, n_fft // 2 + 1, 2, device="cuda")
    ops.gl_stft(torch.from_numpy(x).cuda(), w, spec, n_fft, hop, pad == "reflect")
    ref = glo.stft(x, n_fft, hop, win, pad)
    got = torch.view_as_complex(spec).cpu().numpy()
    assert np.abs(got - ref).max() <= 2e-5 * np.abs(ref).max()
    # istft of an arbitrary (non-consistent) spectrum, complex DC / Nyquist bins included
    rng = np.random.default_rng(hop)
    mag = np.abs(rng.standard_normal(ref.shape)).astype(np.float32)
    ang = np.exp(2j * np.pi * rng.random(ref.shape)).astype(np.complex64)
    frames = torch.empty(T, n_fft, device="cuda")
    y = torch.empty(hop * (T - 1), device="cuda")
    ops.gl_istft(torch.from_numpy(mag).cuda(), torch.view_as_real(torch.from_numpy(ang).cuda()).contiguous(), w, frames, y, n_fft, hop)
    yref = glo.istft(mag * ang, n_fft, hop, win)
    assert np.abs(y.cpu().numpy() - yref).max() <= 2e-5 * max(1.0, np.abs(yref).max())


@pytest.mark.parametrize("n_fft,hop,win,n_iter", [(1024, 256, None, 8), (2048, 300, 1200, 32)])
def test_griffin_lim_vs_oracle_on_the_same_initial_phases(n_fft, hop, win, n_iter):
    from seq2seq_vc_b200 import griffin_lim

    T = 48
    x = _signal(hop * (T - 1), 24000, 3)
    S = np.abs(glo.stft(x, n_fft, hop, win)).astype(np.float32)
    ang = np.exp(2j * np.pi * np.random.default_rng(5).random(S.shape))
    got = griffin_lim(S, n_fft, hop, win_length=win, n_iter=n_iter, init_angles=ang)
    ref = glo.griffin_lim(S, n_fft, hop, win, n_iter=n_iter, init_angles=ang)
    assert got.shape == ref.shape == (hop * (T - 1),)
    # fp32 on the device vs float64 in the oracle over n_iter round trips: waveform and the spectral consistency it reaches
    # (32 rounds amplify the rounding differences wherever |rebuilt - c * previous| is tiny, so the long run is held in the L2 sense)
    rel = np.linalg.norm(got - ref) / np.linalg.norm(ref)
    print(f"griffin_lim n_fft {n_fft} n_iter {n_iter}: relative L2 {rel:.2e}, max abs {np.abs(got - ref).max():.2e}")
    assert rel <= (1e-3 if n_iter <= 8 else 3e-2), rel
    e_got = np.linalg.norm(np.abs(glo.stft(got, n_fft, hop, win)) - S) / np.linalg.norm(S)
    e_ref = np.linalg.norm(np.abs(glo.stft(ref, n_fft, hop, win)) - S) / np.linalg.norm(S)
    assert abs(e_got - e_ref) <= 2e-3 and e_got < 0.3


def test_logmel2linear_and_spectrogram2waveform_round_trip():
    """log-mel from the CUDA front end -> pseudo-inverse mel basis -> Griffin-Lim: logmel2linear vs the oracle's formula, decode()
    vs the oracle pipeline on the same phases, and the re-analysed log-mel of the waveform close to the input."""
    from seq2seq_vc_b200 import Spectrogram2Waveform, api, logmel2linear

    fs, n_fft, hop, n_mels = 24000, 1024, 256, 80
    x = _signal(hop * 40, fs, 9)
    lmspc = api.logmelfilterbank(x, fs, fft_size=n_fft, hop_size=hop, num_mels=n_mels, fmin=80, fmax=7600)
    lin = logmel2linear(lmspc, fs, n_fft, n_mels, 80, 7600)
    ref = glo.logmel2linear(lmspc.astype(np.float64), fs, n_fft, n_mels, 80, 7600)
    assert lin.shape == ref.shape and np.abs(lin - ref).max() <= 1e-4 * np.abs(ref).max()
    stats = dict(mean=lmspc.mean(0), scale=lmspc.std(0) + 1e-3)
    s2w = Spectrogram2Waveform(n_fft=n_fft, n_shift=hop, stats=stats, fs=fs, n_mels=n_mels, fmin=80, fmax=7600, griffin_lim_iters=16)
    ang = np.exp(2j * np.pi * np.random.default_rng(1).random(lin.shape))
    norm = torch.from_numpy((lmspc - stats["mean"]) / stats["scale"]).cuda()
    wav = s2w.decode(norm, init_angles=ang)
    assert wav.device.type == "cuda" and wav.shape == (hop * (lmspc.shape[0] - 1),)
    wref = glo.griffin_lim(ref, n_fft, hop, n_iter=16, init_angles=ang)
    assert np.abs(wav.cpu().numpy() - wref).max() <= 1e-2 * np.abs(wref).max()
    back = lo.logmelfilterbank(wav.cpu().numpy(), fs, fft_size=n_fft, hop_size=hop, num_mels=n_mels, fmin=80, fmax=7600)
    m = min(back.shape[0], lmspc.shape[0])
    assert np.abs(back[2:m - 2] - lmspc[2:m - 2]).mean() <= 0.25          # log10 units: Griffin-Lim is approximate, not an inverse
