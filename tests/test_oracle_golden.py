"""The CPU oracle reproduces the committed golden vectors dumped from the live reference."""
import os

import numpy as np
import pytest
import torch

from oracle import logmel_oracle, mas_oracle, vtn_oracle

GOLD = os.path.join(os.path.dirname(__file__), "golden")
TINY_HP = dict(idim=80, odim=80, dprenet_layers=2, dprenet_units=16, adim=32, aheads=2, elayers=1, eunits=48,
               dlayers=2, dunits=48, postnet_layers=3, postnet_filts=5, postnet_chans=16, decoder_reduction_factor=2)


@pytest.mark.parametrize("fixture,r", [("vtn_tiny.npz", 2), ("vtn_r1_tiny.npz", 1), ("vtn_r3_tiny.npz", 3), ("vtn_r4_tiny.npz", 4)])
def test_vtn_oracle_forward_loss_grads(fixture, r):
    """decoder_reduction_factor 2 (class default), 1, 3 and 4 (the recipe's value, egs/arctic/vc1/conf/vtn.v1.yaml:43) with ragged
    target lengths that are not multiples of r (models/vtn.py:227-243,262-274)."""
    z = np.load(os.path.join(GOLD, fixture))
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd.")}
    out, (l1, bce), grads = vtn_oracle.vtn_loss_and_grads(sd, dict(TINY_HP, decoder_reduction_factor=r), torch.from_numpy(z["xs"]), z["ilens"].tolist(),
                                                          torch.from_numpy(z["ys"]), torch.from_numpy(z["labels"]),
                                                          z["olens"].tolist())
    assert np.abs(out["after_outs"].detach().numpy() - z["after_outs"]).max() <= 2e-5
    assert np.abs(out["before_outs"].detach().numpy() - z["before_outs"]).max() <= 2e-5
    assert np.abs(out["logits"].detach().numpy() - z["logits"]).max() <= 2e-5
    assert abs(float(l1) - float(z["l1_loss"])) <= 1e-6 and abs(float(bce) - float(z["bce_loss"])) <= 1e-6
    np.testing.assert_array_equal(out["labels"].numpy(), z["labels_out"])
    assert out["olens"] == z["olens_out"].tolist() and out["ilens_ds_st"] == z["ilens_ds_st"].tolist()
    for i, a in enumerate(out["att_ws"]):
        assert np.abs(a.detach().numpy() - z[f"att_ws.{i}"]).max() <= 1e-6
    for k, g in grads.items():
        ref = z["grad." + k]
        assert np.abs(g.numpy() - ref).max() <= 2e-4 * (np.abs(ref).max() + 1e-5) + 1e-7, k      # 1e-7: scalar sums (alpha) in another order


@pytest.mark.parametrize("layer", ["conv1d", "conv1d_linear"])
def test_vtn_conv_positionwise_oracle_forward_loss_grads(layer):
    """VTN whose Transformer encoder uses MultiLayeredConv1d / Conv1dLinear (kernel size 3) position-wise layers
    (modules/transformer/encoder.py:143-175, multi_layer_conv.py:12-108) against the live-reference dumps."""
    z = np.load(os.path.join(GOLD, f"vtn_{layer}_k3_tiny.npz"))
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd.")}
    out, (l1, bce), grads = vtn_oracle.vtn_loss_and_grads(sd, dict(TINY_HP, elayers=2), torch.from_numpy(z["xs"]), z["ilens"].tolist(),
                                                          torch.from_numpy(z["ys"]), torch.from_numpy(z["labels"]), z["olens"].tolist())
    for k in ("after_outs", "before_outs", "logits"):
        assert np.abs(out[k].detach().numpy() - z[k]).max() <= 2e-5, k
    assert abs(float(l1) - float(z["l1_loss"])) <= 1e-6 and abs(float(bce) - float(z["bce_loss"])) <= 1e-6
    for k, g in grads.items():
        ref = z["grad." + k]
        assert np.abs(g.numpy() - ref).max() <= 2e-4 * (np.abs(ref).max() + 1e-5) + 1e-7, k


@pytest.mark.parametrize("rel", ["legacy", "latest"])
def test_vtn_conformer_oracle_forward_loss_grads(rel):
    """VTN(encoder_type="conformer") (models/vtn.py:83-143): the class-default legacy rel-pos attention (reversed 5000-row table,
    wrap-around rel_shift, attention.py:114-207) and conformer_rel_pos_type="latest", against the live-reference dumps."""
    z = np.load(os.path.join(GOLD, f"vtn_conformer_{rel}_tiny.npz"))
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd.")}
    hp = dict(TINY_HP, elayers=2, encoder_type="conformer", conformer_rel_pos_type=rel)
    args = (torch.from_numpy(z["xs"]), z["ilens"].tolist(), torch.from_numpy(z["ys"]), torch.from_numpy(z["labels"]), z["olens"].tolist())
    bn = {}
    out = vtn_oracle.vtn_forward(sd, hp, *args, training=True, bn_stats=bn)
    for k in ("after_outs", "before_outs", "logits"):
        assert np.abs(out[k].detach().numpy() - z[k]).max() <= 2e-5, k
    for k in [k for k in z.files if k.startswith("attn.encoder.")]:
        assert np.abs(out["attn"]["encoder." + k[len("attn.encoder."):]].detach().numpy() - z[k]).max() <= 1e-6, k
    for i, a in enumerate(out["att_ws"]):
        assert np.abs(a.detach().numpy() - z[f"att_ws.{i}"]).max() <= 1e-6
    for k, v in bn.items():
        assert np.abs(v.numpy() - z["bn_after." + k]).max() <= 1e-5, k
    _, (l1, bce), grads = vtn_oracle.vtn_loss_and_grads(sd, hp, *args)
    assert abs(float(l1) - float(z["l1_loss"])) <= 1e-6 and abs(float(bce) - float(z["bce_loss"])) <= 1e-6
    gmax = max(np.abs(z[k]).max() for k in z.files if k.startswith("grad."))
    for k, g in grads.items():
        ref = z["grad." + k]
        assert np.abs(g.numpy() - ref).max() <= 2e-4 * np.abs(ref).max() + 2e-6 * gmax, k
    oute = vtn_oracle.vtn_forward({**sd, **{k[9:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("bn_after.")}}, hp, *args, training=False)
    assert np.abs(oute["after_outs"].detach().numpy() - z["eval_after_outs"]).max() <= 2e-5


def test_mas_oracle_matches_reference_numba():
    z = np.load(os.path.join(GOLD, "mas.npz"))
    keys = [k for k in z.files if k.startswith("lp.")]
    assert len(keys) >= 20
    for k in keys:
        lp = z[k]
        ref = z["path." + k[3:]]
        np.testing.assert_array_equal(mas_oracle.mas_path_py(lp), ref)
        paths, ds = mas_oracle.mas_batch_c(lp[None], [lp.shape[1]], [lp.shape[0]])
        np.testing.assert_array_equal(paths[0], ref)
    ds, bin_loss, _ = mas_oracle.viterbi_decode_oracle(z["vd_lp"], z["vd_tl"], z["vd_fl"])
    np.testing.assert_array_equal(ds, z["vd_ds"])
    assert abs(bin_loss - float(z["vd_bin_loss"])) <= 1e-6


def test_reference_docstring_kats():
    z = np.load(os.path.join(GOLD, "kats.npz"))
    np.testing.assert_array_equal(~vtn_oracle.non_pad_mask([5, 3, 2], 5).numpy(), z["pad_mask_532"].astype(bool))
    np.testing.assert_array_equal(vtn_oracle.causal_mask(3).numpy(), z["subsequent_mask_3"].astype(bool))
    # guided attention matrix examples of losses/guided_attention_loss.py:72-90
    def ga(olen, ilen, sigma=0.4):
        t = np.arange(olen)[:, None] / olen
        s = np.arange(ilen)[None, :] / ilen
        return 1 - np.exp(-((s - t) ** 2) / (2 * sigma * sigma))
    np.testing.assert_allclose(ga(5, 5), z["ga_5x5"], atol=1e-6)
    np.testing.assert_allclose(ga(6, 3), z["ga_6x3"], atol=1e-6)


def test_mel_filterbank_against_torchaudio():
    ta = pytest.importorskip("torchaudio")
    for sr, n_fft in ((48000, 2048), (16000, 1024), (24000, 2048)):
        ours = logmel_oracle.mel_basis(sr, n_fft, 80, 0.0, sr / 2)
        theirs = ta.functional.melscale_fbanks(n_fft // 2 + 1, 0.0, sr / 2, 80, sr, norm="slaney", mel_scale="slaney").T.numpy()
        np.testing.assert_allclose(ours, theirs, atol=2e-6)


def test_product_filterbank_and_window_match_oracle():
    from seq2seq_vc_b200.api import hann_window, mel_filterbank

    for sr, n_fft, fmin, fmax in ((48000, 2048, 0.0, 24000.0), (16000, 1024, 80.0, 7600.0)):
        np.testing.assert_allclose(mel_filterbank(sr, n_fft, 80, fmin, fmax), logmel_oracle.mel_basis(sr, n_fft, 80, fmin, fmax), atol=1e-7)
    np.testing.assert_allclose(hann_window(2048, None), logmel_oracle.hann_padded(2048, None).astype(np.float32), atol=1e-7)
    np.testing.assert_allclose(hann_window(2048, 1200), logmel_oracle.hann_padded(2048, 1200).astype(np.float32), atol=1e-7)


def test_tts_oracle_forward_and_guided_attention():
    z = np.load(os.path.join(GOLD, "tts_tiny.npz"))
    hp = dict(idim=40, odim=80, dprenet_layers=2, dprenet_units=16, adim=32, aheads=4, elayers=1, eunits=48,
              dlayers=2, dunits=48, postnet_layers=2, postnet_filts=5, postnet_chans=16, decoder_reduction_factor=2)
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd.")}
    out = vtn_oracle.tts_forward(sd, hp, torch.from_numpy(z["tokens"]), z["ilens"].tolist(), torch.from_numpy(z["ys"]),
                                 torch.from_numpy(z["labels"]), z["olens"].tolist(), use_guided_attn_loss=True)
    assert np.abs(out["after_outs"].numpy() - z["after_outs"]).max() <= 2e-5
    assert np.abs(out["att_ws"].numpy() - z["att_ws"]).max() <= 1e-6
    assert out["ilens"] == z["ilens_out"].tolist()
    ga = vtn_oracle.guided_attention_loss(out["att_ws"], out["ilens"], out["olens_in"])
    assert abs(float(ga) - float(z["ga_loss"])) <= 1e-6


AAS_HP = dict(idim=80, odim=80, adim=32, aheads=2, elayers=1, eunits=48, dlayers=1, dunits=48,
              duration_predictor_input_dim=80, duration_predictor_layers=2, duration_predictor_chans=16,
              duration_predictor_kernel_size=3, postnet_layers=2, postnet_filts=5, postnet_chans=16,
              post_encoder_reduction_factor=4, conformer_enc_kernel_size=7, conformer_dec_kernel_size=7)


@pytest.mark.parametrize("fixture", ["aasvc_tiny.npz", "aasvc_conv1d_tiny.npz", "aasvc_conv1d_k3_tiny.npz", "aasvc_conv1d_linear_k3_tiny.npz"])
def test_aasvc_oracle_forward_loss_grads(fixture):
    """AASVC forward, the four losses of AASVCTrainer._train_step and every gradient vs the live-reference dump
    (Linear + Swish position-wise layers of the shipped yaml; MultiLayeredConv1d k = 1 + ReLU of the class default;
    MultiLayeredConv1d and Conv1dLinear with kernel size 3, multi_layer_conv.py:12-108)."""
    from oracle import aasvc_oracle

    z = np.load(os.path.join(GOLD, fixture))
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd.")}
    bn = {}
    out = aasvc_oracle.aasvc_forward(sd, AAS_HP, torch.from_numpy(z["xs"]), z["ilens"].tolist(), torch.from_numpy(z["ys"]),
                                     z["olens"].tolist(), torch.from_numpy(z["dp_inputs"]), training=True, bn_stats=bn)
    for k in ("after_outs", "before_outs", "d_outs"):
        assert np.abs(out[k].detach().numpy() - z[k]).max() <= 2e-5, k
    lp, ref = out["log_p_attn"].detach().numpy(), z["log_p_attn"]
    assert np.array_equal(np.isfinite(lp), np.isfinite(ref)) and np.abs(lp[np.isfinite(ref)] - ref[np.isfinite(ref)]).max() <= 2e-5
    np.testing.assert_array_equal(out["ds"].numpy(), z["ds"])                      # integer alignment path: bit-exact
    assert out["ilens"] == z["ilens_out"].tolist() and out["olens"] == z["olens_out"].tolist()
    for k, v in bn.items():
        assert np.abs(v.numpy() - z["bn_after." + k]).max() <= 1e-5, k
    for k in [k for k in z.files if k.startswith("attn.")]:
        assert np.abs(out["attn"][k[5:]].detach().numpy() - z[k]).max() <= 1e-6, k
    _, parts, grads = aasvc_oracle.aasvc_loss_and_grads(sd, AAS_HP, torch.from_numpy(z["xs"]), z["ilens"].tolist(),
                                                        torch.from_numpy(z["ys"]), z["olens"].tolist(), torch.from_numpy(z["dp_inputs"]))
    for k in ("l1_loss", "forward_sum_loss", "bin_loss", "duration_loss"):
        assert abs(float(parts[k]) - float(z[k])) <= 2e-6 * max(1.0, abs(float(z[k]))), k
    gmax = max(np.abs(z[k]).max() for k in z.files if k.startswith("grad."))
    for k, g in grads.items():
        ref = z["grad." + k]
        # gradients that are zero in exact arithmetic (key bias under softmax, biases ahead of BatchNorm) are round-off noise
        assert np.abs(g.numpy() - ref).max() <= 2e-4 * np.abs(ref).max() + 2e-6 * gmax, k


FS_HP = dict(idim=80, odim=80, adim=32, aheads=2, elayers=2, eunits=48, dlayers=2, dunits=48, duration_predictor_input_dim=80,
             duration_predictor_layers=2, duration_predictor_chans=16, duration_predictor_kernel_size=3, postnet_layers=2, postnet_filts=5,
             postnet_chans=16, conformer_enc_kernel_size=7, conformer_dec_kernel_size=7)


def test_fastspeech_vc_oracle_forward_loss_grads():
    """FastSpeechVC (conformer encoder / decoder, conv2d input layer, LengthRegulator with teacher durations: the configuration of
    egs/arctic/vc2/conf/fs2_vc.melmelmel.v1.yaml) + the two losses of NARVCTrainer._train_step vs the live-reference dump."""
    from oracle import fsvc_oracle

    z = np.load(os.path.join(GOLD, "fsvc_tiny.npz"))
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd.")}
    args = (torch.from_numpy(z["xs"]), z["ilens"].tolist(), torch.from_numpy(z["ys"]), z["olens"].tolist(), torch.from_numpy(z["ds"]),
            torch.from_numpy(z["dp_inputs"]))
    bn = {}
    out = fsvc_oracle.fsvc_forward(sd, FS_HP, *args, training=True, bn_stats=bn)
    for k in ("after_outs", "before_outs", "d_outs"):
        assert np.abs(out[k].detach().numpy() - z[k]).max() <= 2e-5, k
    assert out["ilens"] == z["ilens_out"].tolist() and out["olens"] == z["olens_out"].tolist()
    for k, v in bn.items():
        assert np.abs(v.numpy() - z["bn_after." + k]).max() <= 1e-5, k
    for k in [k for k in z.files if k.startswith("attn.")]:
        assert np.abs(out["attn"][k[5:]].detach().numpy() - z[k]).max() <= 1e-6, k
    _, parts, grads = fsvc_oracle.fsvc_loss_and_grads(sd, FS_HP, *args)
    for k in ("l1_loss", "duration_loss"):
        assert abs(float(parts[k]) - float(z[k])) <= 2e-6 * max(1.0, abs(float(z[k]))), k
    gmax = max(np.abs(z[k]).max() for k in z.files if k.startswith("grad."))
    for k, g in grads.items():
        ref = z["grad." + k]
        assert np.abs(g.numpy() - ref).max() <= 2e-4 * np.abs(ref).max() + 2e-6 * gmax, k
    sd_e = {**sd, **{k[9:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("bn_after.")}}
    oe = fsvc_oracle.fsvc_forward(sd_e, FS_HP, *args, training=False)
    assert np.abs(oe["after_outs"].detach().numpy() - z["eval_after_outs"]).max() <= 2e-5


def test_forward_sum_oracle_matches_reference_ctc():
    """Hand-rolled alpha recursion vs the reference's F.ctc_loss path incl. an infeasible utterance (zero_infinity)."""
    from oracle import aasvc_oracle

    z = np.load(os.path.join(GOLD, "aasvc_tiny.npz"))
    lp = torch.from_numpy(z["fs_lp"]).requires_grad_(True)
    loss = aasvc_oracle.forward_sum_loss(lp, z["fs_tl"].tolist(), z["fs_fl"].tolist())
    assert abs(float(loss) - float(z["fs_loss"])) <= 1e-5
    (g,) = torch.autograd.grad(loss, lp)
    assert np.abs(g.numpy() - z["fs_grad"]).max() <= 1e-5


def test_rel_shift_restatement():
    """bd'[i, j] = bd[i, T-1-i+j] is exactly the reference's pad/view/slice rel_shift (attention.py:237-260)."""
    T = 7
    x = torch.arange(2 * 3 * T * (2 * T - 1), dtype=torch.float32).view(2, 3, T, 2 * T - 1)
    xp = torch.cat([torch.zeros(2, 3, T, 1), x], dim=-1).view(2, 3, 2 * T, T)[:, :, 1:].reshape(2, 3, T, 2 * T - 1)[..., :T]
    idx = T - 1 - torch.arange(T)[:, None] + torch.arange(T)[None, :]
    assert torch.equal(xp, torch.gather(x, 3, idx[None, None].expand(2, 3, T, T)))


def test_inference_oracles_match_reference():
    """Autoregressive VTN.inference and AASVC.inference restatements vs the live-reference dumps."""
    from oracle import aasvc_oracle

    z = np.load(os.path.join(GOLD, "vtn_tiny.npz"))
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd.")}
    sd.update({k[9:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("bn_after.")})
    il = int(z["ilens"][0])
    o, p, a = vtn_oracle.vtn_inference(sd, TINY_HP, torch.from_numpy(z["xs"])[0, :il], 0.9999, 0.0, 1.6)
    assert np.abs(o.numpy() - z["inf_outs"]).max() <= 2e-5 and np.abs(p.numpy() - z["inf_probs"]).max() <= 1e-6
    assert np.abs(a.numpy() - z["inf_att_ws"]).max() <= 1e-6
    z = np.load(os.path.join(GOLD, "aasvc_tiny.npz"))
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd.")}
    sd.update({k[7:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("inf_bn.")})
    sd["duration_predictor.linear.bias"] = torch.from_numpy(z["inf_dp_bias"])
    il = int(z["ilens"][0])
    outs, d = aasvc_oracle.aasvc_inference(sd, AAS_HP, torch.from_numpy(z["xs"])[0, :il], torch.from_numpy(z["dp_inputs"])[0, :il])
    np.testing.assert_array_equal(d.numpy(), z["inf_d_outs"])
    assert np.abs(outs.numpy() - z["inf_outs"]).max() <= 2e-5


# ------------------------------------------------------------------ stochastic duration predictor: oracle groundwork
SDP_HP = dict(channels=16, kernel_size=3, dds_conv_layers=3, flows=4)


def _sdp():
    z = np.load(os.path.join(GOLD, "sdp_tiny.npz"))
    sd = {"duration_predictor." + k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd.")}
    return z, sd


def test_sdp_oracle_state_dict_spec():
    from oracle import sdp_oracle

    z, sd = _sdp()
    spec = dict(sdp_oracle.state_dict_spec(SDP_HP))
    assert set(spec) == set(sd) and all(tuple(sd[k].shape) == spec[k] for k in sd)


def test_sdp_oracle_nll_and_gradients_match_reference():
    """Variational duration NLL as AASVC._forward computes it (models/aas_vc.py:412-419) with the reference's recorded noise:
    value and every parameter gradient of sum(dur_nll) vs the live-reference dump."""
    from oracle import sdp_oracle

    z, sd = _sdp()
    sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    nll = sdp_oracle.aasvc_dur_nll(sd, "duration_predictor", SDP_HP, torch.from_numpy(z["dp_inputs"]), z["text_lens"].tolist(),
                                   torch.from_numpy(z["ds"]), torch.from_numpy(z["e_q"]))
    assert np.abs(nll.detach().numpy() - z["dur_nll"]).max() <= 1e-5 * np.abs(z["dur_nll"]).max()
    nll.sum().backward()
    gmax = max(np.abs(z[k]).max() for k in z.files if k.startswith("grad."))
    n = 0
    for k in z.files:
        if k.startswith("grad."):
            g = sd["duration_predictor." + k[5:]].grad
            assert g is not None, k
            # fp32 on both sides through log / sqrt / softmax chains of 8 splines: the result moves by a few 1e-4 relative
            # with the BLAS thread count alone (seen between an isolated run and the whole suite)
            assert np.abs(g.numpy() - z[k]).max() <= 2e-3 * np.abs(z[k]).max() + 1e-5 * gmax, k
            n += 1
    assert n == sum(1 for k in z.files if k.startswith("sd."))          # every parameter trains (x is detached, not the weights)


def test_sdp_oracle_inverse_durations_match_reference():
    """Inference direction (models/aas_vc.py:385-393): integer durations exact, incl. a wide-noise draw that lands in the
    linear tails of the splines."""
    from oracle import sdp_oracle

    z, sd = _sdp()
    for zk, dk, scale in (("z", "d_outs", 0.8), ("z_wide", "d_outs_wide", 4.0)):
        d = sdp_oracle.aasvc_dur_inference(sd, "duration_predictor", SDP_HP, torch.from_numpy(z["dp_inputs"]), z["text_lens"].tolist(),
                                           torch.from_numpy(z[zk]), noise_scale=scale)
        np.testing.assert_array_equal(d.numpy(), z[dk])
    assert (np.abs(z["z_wide"] * 4.0) > 5.0).any()                       # the tails are exercised


def test_sdp_spline_round_trip_and_tails():
    """Size-independent properties of the spline: inverse(forward(x)) == x, log-determinants cancel, identity outside +-5."""
    from oracle import sdp_oracle

    g = torch.Generator().manual_seed(7)
    x = torch.randn(4, 1, 50, generator=g) * 3.0
    uw, uh, ud = (torch.randn(4, 1, 50, n, generator=g) for n in (10, 10, 9))
    y, lad = sdp_oracle.rq_spline_linear_tails(x, uw, uh, ud, inverse=False)
    xr, ladr = sdp_oracle.rq_spline_linear_tails(y, uw, uh, ud, inverse=True)
    assert (xr - x).abs().max().item() <= 2e-4 and (lad + ladr).abs().max().item() <= 2e-3
    out = x.abs() > 5.0
    assert out.any() and torch.equal(y[out], x[out]) and (lad[out] == 0).all()
    xs = torch.linspace(-6, 6, 400).view(1, 1, 400)                       # one spline, many abscissae: strictly increasing
    ys, _ = sdp_oracle.rq_spline_linear_tails(xs, uw[:1, :, :1].expand(1, 1, 400, 10), uh[:1, :, :1].expand(1, 1, 400, 10),
                                              ud[:1, :, :1].expand(1, 1, 400, 9), inverse=False)
    assert (torch.diff(ys.flatten()) > 0).all()
