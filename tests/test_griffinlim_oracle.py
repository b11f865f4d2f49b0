"""The Griffin-Lim oracle (oracle/griffinlim_oracle.py, PARITY UNPINNED against librosa) against independent implementations of
its pieces: scipy.signal for the STFT / ISTFT pair, exact reconstruction properties, and the reference's logmel2linear formula."""
import numpy as np
import pytest
import scipy.signal

from oracle import griffinlim_oracle as gl
from oracle import logmel_oracle as lo


@pytest.mark.parametrize("n_fft,hop,win", [(1024, 256, None), (2048, 300, 1200), (512, 128, 400)])
def test_stft_istft_against_scipy(n_fft, hop, win):
    rng = np.random.default_rng(n_fft)
    T = 40
    x = rng.standard_normal(hop * (T - 1))
    w = lo.hann_padded(n_fft, win)
    ours = gl.stft(x, n_fft, hop, win, "constant")
    assert ours.shape == (T, n_fft // 2 + 1)
    _, _, Z = scipy.signal.stft(np.pad(x, (n_fft // 2, n_fft // 2)), window=w, nperseg=n_fft, noverlap=n_fft - hop, boundary=None,
                                padded=False)
    assert np.abs(ours - (Z * w.sum()).T[:T]).max() <= 1e-9 * np.abs(ours).max()
    back = gl.istft(ours, n_fft, hop, win)
    assert back.shape == x.shape
    assert np.abs(back - x).max() <= 1e-9                        # the weighted overlap-add inverts the STFT exactly
    _, xs = scipy.signal.istft(Z, window=w, nperseg=n_fft, noverlap=n_fft - hop, boundary=False)
    n = min(len(xs) - n_fft // 2, len(back))
    assert np.abs(xs[n_fft // 2:n_fft // 2 + n] - back[:n]).max() <= 1e-8


def test_griffin_lim_reduces_spectral_error_and_is_deterministic_given_phases():
    rng = np.random.default_rng(1)
    n_fft, hop, T = 512, 128, 60
    t = np.arange(hop * (T - 1)) / 16000.0
    x = 0.5 * np.sin(2 * np.pi * 440 * t) + 0.3 * np.sin(2 * np.pi * 1250 * t + 1.0) + 0.01 * rng.standard_normal(t.shape)
    S = np.abs(gl.stft(x, n_fft, hop))
    ang = np.exp(2j * np.pi * rng.random(S.shape))
    err = []
    for n_iter in (0, 4, 32):
        y = gl.griffin_lim(S, n_fft, hop, n_iter=n_iter, init_angles=ang)
        assert y.shape == x.shape
        err.append(np.linalg.norm(np.abs(gl.stft(y, n_fft, hop)) - S) / np.linalg.norm(S))
    assert err[2] < err[1] < err[0] and err[2] < 0.12, err
    assert np.array_equal(gl.griffin_lim(S, n_fft, hop, n_iter=4, init_angles=ang), gl.griffin_lim(S, n_fft, hop, n_iter=4, init_angles=ang))


def test_logmel2linear_inverts_the_mel_projection_in_the_least_squares_sense():
    rng = np.random.default_rng(2)
    fs, n_fft, n_mels = 24000, 1024, 80
    basis = lo.mel_basis(fs, n_fft, n_mels, 80, 7600)
    lin = np.abs(rng.standard_normal((7, n_fft // 2 + 1))) + 0.1
    lmspc = np.log10(np.maximum(1e-10, lin @ basis.T))
    rec = gl.logmel2linear(lmspc, fs, n_fft, n_mels, 80, 7600)
    assert rec.shape == lin.shape and (rec >= 1e-10).all()
    # projecting the reconstruction again reproduces the mel spectrum wherever the clamp at EPS did not bite
    free = np.dot(np.linalg.pinv(basis), (10.0 ** lmspc).T).T
    assert np.abs((free @ basis.T) - 10.0 ** lmspc).max() <= 1e-6 * (10.0 ** lmspc).max()


def test_feature_statistics_host_logic_matches_sklearn(monkeypatch):
    """seq2seq_vc_b200.FeatureStatistics (the device accumulator replaced by its CPU contract) == sklearn StandardScaler.partial_fit
    over the same utterances (bin/compute_statistics.py:128-132), incl. a constant feature (scale 1) and the ragged-batch form."""
    import fake_ops
    import torch
    from sklearn.preprocessing import StandardScaler

    from seq2seq_vc_b200 import FeatureStatistics

    fake_ops.install(monkeypatch)
    rng = np.random.default_rng(5)
    utts = [(rng.standard_normal((t, 12)) * rng.uniform(0.1, 3.0, 12) + rng.uniform(-4, 4, 12)).astype(np.float32) for t in (31, 7, 120, 64)]
    for u in utts:
        u[:, 3] = 1.25                                    # constant feature
    sk = StandardScaler()
    ours = FeatureStatistics(device="cpu")
    for u in utts:
        sk.partial_fit(u)
        ours.partial_fit(u)
    assert ours.n_samples_seen_ == int(sk.n_samples_seen_)
    np.testing.assert_allclose(ours.mean_, sk.mean_, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(ours.var_, sk.var_, rtol=1e-8, atol=1e-12)
    np.testing.assert_allclose(ours.scale_, sk.scale_, rtol=1e-8, atol=1e-12)
    assert ours.scale_[3] == 1.0
    assert ours.stats().dtype == np.float32 and ours.stats().shape == (2, 12)
    batch = np.zeros((4, 120, 12), np.float32)
    for b, u in enumerate(utts):
        batch[b, : len(u)] = u
    both = FeatureStatistics(device="cpu").partial_fit(torch.from_numpy(batch), lens=[len(u) for u in utts])
    np.testing.assert_allclose(both.mean_, sk.mean_, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(both.scale_, sk.scale_, rtol=1e-8, atol=1e-12)
