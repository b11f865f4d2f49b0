"""Monotonic alignment search (bit-exact) and STFT->log-mel on the GPU vs golden vectors and the CPU oracles."""
import os

import numpy as np
import pytest
import torch

from oracle import logmel_oracle, mas_oracle

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def run_mas(lp, tl, fl):
    from seq2seq_vc_b200 import api

    ds, bin_loss, paths = api.viterbi_decode(torch.from_numpy(lp).cuda(), tl, fl, return_paths=True)
    torch.cuda.synchronize()
    return ds.cpu().numpy(), float(bin_loss), paths.cpu().numpy().astype(np.int64)


def test_mas_golden_cases_bit_exact():
    z = np.load(os.path.join(GOLD, "mas.npz"))
    for k in [k for k in z.files if k.startswith("lp.")]:
        lp = z[k]
        ds, _, paths = run_mas(lp[None].copy(), [lp.shape[1]], [lp.shape[0]])
        np.testing.assert_array_equal(paths[0], z["path." + k[3:]], err_msg=k)
        np.testing.assert_array_equal(ds[0], np.bincount(z["path." + k[3:]], minlength=lp.shape[1]).astype(np.float32))
    ds, bin_loss, _ = run_mas(z["vd_lp"], z["vd_tl"].tolist(), z["vd_fl"].tolist())
    np.testing.assert_array_equal(ds, z["vd_ds"])
    assert abs(bin_loss - float(z["vd_bin_loss"])) <= 1e-5


@pytest.mark.parametrize("B,TF,TT,kind", [(8, 300, 75, "rand"), (8, 300, 75, "ties"), (64, 768, 192, "rand"), (4, 1000, 1, "rand"),
                                          (3, 40, 40, "ties"), (2, 2000, 500, "rand")])
def test_mas_fuzz_vs_c_oracle_bit_exact(B, TF, TT, kind):
    rng = np.random.default_rng(B * 1000 + TF)
    lp = torch.log_softmax(torch.from_numpy(rng.standard_normal((B, TF, TT)).astype(np.float32)), -1).numpy()
    if kind == "ties":
        lp = (np.round(lp * 2) / 2).astype(np.float32)
    tl = rng.integers(1, TT + 1, B)
    fl = np.maximum(rng.integers(1, TF + 1, B), tl)          # the reference needs T_feats >= T_text
    tl[0], fl[0] = TT, TF
    for b in range(B):
        lp[b, :, tl[b]:] = -np.inf
    ds, bin_loss, paths = run_mas(lp, tl.tolist(), fl.tolist())
    ds_ref, bl_ref, p_ref = mas_oracle.viterbi_decode_oracle(lp, tl, fl)
    np.testing.assert_array_equal(paths, p_ref)
    np.testing.assert_array_equal(ds, ds_ref)
    assert abs(bin_loss - bl_ref) <= 1e-4 * max(1.0, abs(bl_ref))
    # size-independent properties: monotone path, ends at the last token, durations sum to the length
    for b in range(B):
        p = paths[b, :fl[b]]
        assert p[-1] == tl[b] - 1 and (np.diff(p) >= 0).all() and (np.diff(p) <= 1).all()
        assert ds[b].sum() == fl[b]


def test_mas_bin_loss_gradient():
    from seq2seq_vc_b200 import api

    rng = np.random.default_rng(1)
    lp = torch.log_softmax(torch.from_numpy(rng.standard_normal((3, 30, 9)).astype(np.float32)), -1).cuda().requires_grad_(True)
    ds, bin_loss, paths = api.viterbi_decode(lp, [9, 7, 4], [30, 22, 11], return_paths=True)
    bin_loss.backward()
    g = lp.grad.cpu().numpy()
    for b, fl in enumerate([30, 22, 11]):
        for t in range(fl):
            assert abs(g[b, t, paths[b, t].item()] + 1.0 / (fl * 3)) < 1e-7
    assert abs(np.abs(g).sum() - 1.0) < 1e-5


@pytest.mark.parametrize("sr,n_fft,hop,ns,win", [(24000, 1024, 256, 12000, None), (48000, 2048, 300, 48000, None),
                                                 (16000, 1024, 256, 5000, 800), (48000, 2048, 300, 4801, 1200),
                                                 (16000, 512, 128, 7001, None), (16000, 1024, 256, 16001, None), (8000, 256, 64, 3000, None)])
def test_logmel_vs_oracle(sr, n_fft, hop, ns, win):
    from seq2seq_vc_b200 import api

    rng = np.random.default_rng(ns)
    wav = np.clip(0.1 * rng.standard_normal(ns), -1, 1).astype(np.float32)
    got = api.logmelfilterbank(wav, sr, fft_size=n_fft, hop_size=hop, win_length=win, num_mels=80)
    ref = logmel_oracle.logmelfilterbank(wav, sr, fft_size=n_fft, hop_size=hop, win_length=win, num_mels=80)
    assert got.shape == ref.shape == (1 + ns // hop, 80)
    # 1e-4 absolute in the log10 domain away from the 1e-10 floor
    live = ref > -9.0
    assert np.abs(got - ref)[live].max() <= 1e-4, np.abs(got - ref)[live].max()
    assert np.abs(got - ref).max() <= 5e-2


@pytest.mark.parametrize("num_mels,fmin,fmax", [(80, 80, 7600), (128, 0, None), (160, 0, None), (40, 1000, 20000)])
def test_logmel_2048_filterbank_shapes(num_mels, fmin, fmax):
    """n_fft = 2048 fast path: narrow / full-band / sparse-low banks on the banded tensor-core projection, and a bank with more
    than 128 bands on the kernel's dense projection; two clips whose frame groups straddle the clip boundary, odd sample count."""
    from seq2seq_vc_b200 import api

    rng = np.random.default_rng(num_mels)
    wav = np.clip(0.2 * rng.standard_normal((2, 30001)), -1, 1).astype(np.float32)
    wav[1, 5000:9000] = 0.0
    got = api.logmel_batch(torch.from_numpy(wav).cuda(), 48000, fft_size=2048, hop_size=300, num_mels=num_mels, fmin=fmin,
                           fmax=fmax).cpu().numpy()
    for b in range(2):
        ref = logmel_oracle.logmelfilterbank(wav[b], 48000, fft_size=2048, hop_size=300, num_mels=num_mels, fmin=fmin, fmax=fmax)
        assert got[b].shape == ref.shape
        live = ref > -9.0
        assert np.abs(got[b] - ref)[live].max() <= 1e-4, np.abs(got[b] - ref)[live].max()
        assert np.abs(got[b] - ref).max() <= 5e-2


def test_logmel_2048_unaligned_view():
    """A clip that starts on an odd float offset takes the scalar load path and gives the same frames."""
    from seq2seq_vc_b200 import api

    rng = np.random.default_rng(5)
    buf = torch.from_numpy((0.1 * rng.standard_normal(2 * 20001 + 1)).astype(np.float32)).cuda()
    a = buf[1:].view(2, 20001)
    got = api.logmel_batch(a, 48000, fft_size=2048, hop_size=300, num_mels=80)
    want = api.logmel_batch(a.clone(), 48000, fft_size=2048, hop_size=300, num_mels=80)
    assert torch.equal(got, want)


def test_logmel_batched_and_silence():
    from seq2seq_vc_b200 import api

    rng = np.random.default_rng(0)
    wav = (0.1 * rng.standard_normal((5, 9000))).astype(np.float32)
    wav[3] = 0.0
    mel = api.logmel_batch(torch.from_numpy(wav).cuda(), 24000, fft_size=1024, hop_size=256, num_mels=80).cpu().numpy()
    for b in range(5):
        ref = logmel_oracle.logmelfilterbank(wav[b], 24000, fft_size=1024, hop_size=256, num_mels=80)
        assert np.abs(mel[b] - ref).max() <= 1e-4
    assert np.allclose(mel[3], -10.0)


@pytest.mark.parametrize("n_fft,hop", [(1024, 256), (2048, 300)])
def test_logmel_with_fused_normalisation(n_fft, hop):
    """STFT -> log-mel -> StandardScaler.transform (bin/normalize.py:173-193) in one kernel vs oracle log-mel + sklearn."""
    from sklearn.preprocessing import StandardScaler

    from seq2seq_vc_b200 import api

    rng = np.random.default_rng(3)
    wav = (0.1 * rng.standard_normal((3, 20000))).astype(np.float32)
    refs = [logmel_oracle.logmelfilterbank(w, 24000, fft_size=n_fft, hop_size=hop, num_mels=80) for w in wav]
    scaler = StandardScaler()
    for r in refs:
        scaler.partial_fit(r)                                   # bin/compute_statistics.py:130-132
    got = api.logmel_batch(torch.from_numpy(wav).cuda(), 24000, fft_size=n_fft, hop_size=hop, num_mels=80,
                           mean=scaler.mean_.astype(np.float32), scale=scaler.scale_.astype(np.float32)).cpu().numpy()
    for b in range(3):
        ref = scaler.transform(refs[b])
        assert np.abs(got[b] - ref).max() <= 1e-4 / scaler.scale_.min() + 1e-5
    assert abs(got.mean()) <= 1e-2 and abs(got.std() - 1.0) <= 5e-2


@pytest.mark.parametrize("B,T,D", [(1, 37, 80), (6, 301, 80), (3, 50, 7), (2, 40, 300), (64, 1601, 80), (2, 33, 1032), (2, 33, 258)])
def test_feature_statistics_matches_sklearn(B, T, D):
    """s2s_feat_stats / FeatureStatistics vs sklearn StandardScaler.partial_fit per utterance (bin/compute_statistics.py:128-132):
    float64 sums on the device; mean_ / scale_ agree to 1e-9 relative, the frame count exactly; then log-mel normalised with these
    statistics (the fused store of s2s_logmel_norm) has zero mean and unit variance per bin."""
    from sklearn.preprocessing import StandardScaler

    from seq2seq_vc_b200 import FeatureStatistics

    rng = np.random.default_rng(B * 1000 + T + D)
    lens = rng.integers(1, T + 1, B)
    lens[0] = T
    x = (rng.standard_normal((B, T, D)) * rng.uniform(0.1, 3.0, D) + rng.uniform(-4, 4, D)).astype(np.float32)
    sk = StandardScaler()
    for b in range(B):
        x[b, lens[b]:] = 0
        sk.partial_fit(x[b, : lens[b]])
    ours = FeatureStatistics().partial_fit(torch.from_numpy(x).cuda(), lens=lens)
    assert ours.n_samples_seen_ == int(sk.n_samples_seen_)
    np.testing.assert_allclose(ours.mean_, sk.mean_, rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(ours.scale_, sk.scale_, rtol=1e-9, atol=1e-10)
    one = FeatureStatistics()
    for b in range(B):                                             # the reference's loop: one utterance at a time
        one.partial_fit(x[b, : lens[b]])
    np.testing.assert_allclose(one.mean_, sk.mean_, rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(one.stats(), np.stack([sk.mean_, sk.scale_]).astype(np.float32), rtol=1e-6, atol=1e-7)


def test_statistics_then_fused_normalisation_roundtrip():
    from seq2seq_vc_b200 import FeatureStatistics
    from seq2seq_vc_b200.api import logmel_batch

    g = torch.Generator().manual_seed(3)
    wav = (0.1 * torch.randn(8, 24000, generator=g)).cuda()
    mel = logmel_batch(wav, 24000, 2048, 300, 1200, 80, 80, 7600)
    st = FeatureStatistics().partial_fit(mel)
    stats = st.stats()
    norm = logmel_batch(wav, 24000, 2048, 300, 1200, 80, 80, 7600, mean=torch.from_numpy(stats[0]).cuda(), scale=torch.from_numpy(stats[1]).cuda())
    flat = norm.reshape(-1, 80).double()
    assert flat.mean(0).abs().max().item() <= 1e-4 and (flat.std(0, unbiased=False) - 1).abs().max().item() <= 1e-4
