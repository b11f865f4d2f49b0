"""Host-side orchestration of the AAS-VC engine checked on CPU against golden vectors from the live reference.

The C-ABI kernels are replaced by their torch-CPU contracts (tests/fake_ops.py): this verifies buffer wiring, strided
GEMM operand views, weight packing and the hand-written backward pass of the Conformer blocks, the rel-pos attention,
the alignment module, Gaussian upsampling, the duration predictor and the loss assembly.  The kernels themselves are
verified on the GPU (tests/test_gpu_aasvc.py).
"""
import os

import numpy as np
import pytest
import torch

import fake_ops
from seq2seq_vc_b200.aasvc_engine import AASVCEngine, beta_binomial_log_prior, nearest_index

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "aasvc_tiny.npz")
AAS_HP = dict(idim=80, odim=80, adim=32, aheads=2, elayers=1, eunits=48, dlayers=1, dunits=48,
              duration_predictor_input_dim=80, duration_predictor_layers=2, duration_predictor_chans=16,
              duration_predictor_kernel_size=3, postnet_layers=2, postnet_filts=5, postnet_chans=16,
              post_encoder_reduction_factor=4, conformer_enc_kernel_size=7, conformer_dec_kernel_size=7)
NO_DROPOUT = dict(transformer_enc_dropout_rate=0.0, transformer_enc_positional_dropout_rate=0.0, transformer_enc_attn_dropout_rate=0.0,
                  transformer_dec_dropout_rate=0.0, transformer_dec_positional_dropout_rate=0.0, transformer_dec_attn_dropout_rate=0.0,
                  duration_predictor_dropout_rate=0.0, postnet_dropout_rate=0.0)


# the shipped yaml's Linear + Swish position-wise layers, and the AASVC class default (MultiLayeredConv1d k = 1 + ReLU)
# plus MultiLayeredConv1d / Conv1dLinear with kernel size 3 (multi_layer_conv.py:12-108)
FIXTURES = {"linear": "aasvc_tiny.npz", "conv1d": "aasvc_conv1d_tiny.npz", "conv1d:3": "aasvc_conv1d_k3_tiny.npz",
            "conv1d-linear:3": "aasvc_conv1d_linear_k3_tiny.npz"}


def _pw(key):
    t, _, k = key.partition(":")
    return dict(positionwise_layer_type=t, positionwise_conv_kernel_size=int(k or 1))


@pytest.fixture(params=sorted(FIXTURES))
def engine(monkeypatch, request):
    fake_ops.install(monkeypatch)
    z = np.load(os.path.join(os.path.dirname(GOLDEN), FIXTURES[request.param]))
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd.")}
    eng = AASVCEngine(dict(AAS_HP, **_pw(request.param), **NO_DROPOUT), device="cpu", bf16=False)
    assert set(eng.state_dict()) == set(sd)
    eng.load_state_dict(sd)
    return eng, z


def run_step(eng, z):
    ilens, olens = z["ilens"].tolist(), z["olens"].tolist()
    xs = torch.from_numpy(z["xs"])[:, :max(ilens)].contiguous()
    ys = torch.from_numpy(z["ys"])[:, :max(olens)].contiguous()
    dpi = torch.from_numpy(z["dp_inputs"])[:, :max(ilens)].contiguous()
    after, before = eng.forward(xs, ys, dpi, ilens, olens)
    losses = eng.loss(ys)
    eng.backward()
    return after, before, losses


def test_forward_and_losses_match_reference(engine):
    eng, z = engine
    after, before, losses = run_step(eng, z)
    assert np.abs(after.numpy() - z["after_outs"]).mean() <= 1e-5
    assert np.abs(before.numpy() - z["before_outs"]).mean() <= 1e-5
    lp, ref = eng.log_p_attn.numpy(), z["log_p_attn"]
    fin = np.isfinite(ref)
    assert np.array_equal(np.isfinite(lp), fin) and np.abs(lp[fin] - ref[fin]).max() <= 2e-5
    np.testing.assert_array_equal(eng.ds.numpy(), z["ds"])
    assert np.abs(eng.d_outs.numpy() - z["d_outs"]).max() <= 1e-5
    for i, k in enumerate(("l1_loss", "forward_sum_loss", "bin_loss", "duration_loss")):
        assert abs(float(losses[i]) - float(z[k])) <= 1e-5 * max(1.0, abs(float(z[k]))), k
    for k in [k for k in z.files if k.startswith("attn.")]:
        assert np.abs(eng.attn[k[5:]].numpy() - z[k]).max() <= 1e-6, k
    assert eng.tlens_host == z["ilens_out"].tolist()


def test_gradients_match_reference(engine):
    eng, z = engine
    run_step(eng, z)
    gmax = max(np.abs(z[k]).max() for k in z.files if k.startswith("grad."))
    worst = 0.0
    for name in eng.store.names():
        ref = z["grad." + name]
        got = eng.store.g(name).numpy()
        err = np.abs(got - ref).max()
        worst = max(worst, err / (np.abs(ref).max() + 1e-12))
        assert err <= 2e-4 * np.abs(ref).max() + 2e-6 * gmax, (name, err, np.abs(ref).max())


def test_bn_running_stats_and_eval(engine):
    eng, z = engine
    run_step(eng, z)
    for k in [k for k in z.files if k.startswith("bn_after.")]:
        assert np.abs(eng.buffers[k[9:]].numpy() - z[k]).max() <= 1e-5, k
    eng.training = False
    ilens, olens = z["ilens"].tolist(), z["olens"].tolist()
    after, _ = eng.forward(torch.from_numpy(z["xs"]), torch.from_numpy(z["ys"]), torch.from_numpy(z["dp_inputs"]), ilens, olens)
    assert np.abs(after.numpy() - z["eval_after_outs"]).mean() <= 1e-5


def test_beta_binomial_prior_matches_scipy():
    from scipy.stats import betabinom

    for N, T in ((12, 40), (5, 17), (1, 3), (30, 31)):
        a = np.arange(1, T + 1, dtype=float)
        b = np.array([T - t + 1 for t in a])
        ref = betabinom.logpmf(np.arange(N)[:, None], N, a, b).T
        got = beta_binomial_log_prior(N, T).numpy()
        assert np.abs(got - ref.astype(np.float32)).max() <= 2e-6 * np.abs(ref).max()


def test_nearest_index_matches_torch_interpolate():
    for tin, tout in ((11, 12), (191, 192), (12, 12), (30, 7), (5, 13)):
        x = torch.arange(tin, dtype=torch.float32)[None, None]
        ref = torch.nn.functional.interpolate(x, size=tout).long().view(-1).tolist()
        assert nearest_index(tin, tout) == ref


def test_inference_matches_reference(engine):
    """AASVC.inference (no ground truth): predicted durations (integers) exact, output mel vs the live-reference dump."""
    eng, z = engine
    sd = eng.state_dict()
    for k in z.files:
        if k.startswith("inf_bn."):
            sd[k[7:]] = torch.from_numpy(z[k])
    sd["duration_predictor.linear.bias"] = torch.from_numpy(z["inf_dp_bias"])
    eng.load_state_dict(sd)
    il = int(z["ilens"][0])
    outs, d_outs = eng.inference(torch.from_numpy(z["xs"])[0, :il], torch.from_numpy(z["dp_inputs"])[0, :il])
    np.testing.assert_array_equal(d_outs.numpy(), z["inf_d_outs"])
    assert outs.shape == z["inf_outs"].shape and np.abs(outs.numpy() - z["inf_outs"]).mean() <= 1e-5
    assert eng.training is True                          # restored


def test_dropin_class_default_positionwise_layer():
    """AASVC() defaults to positionwise_layer_type="conv1d", kernel size 1 (models/aas_vc.py:52-53): the drop-in builds the
    reference's state dict for it ((U, d, 1) Conv1d weights), loads a reference checkpoint, and refuses wider kernels."""
    from seq2seq_vc_b200 import AASVC

    fixed = dict(duration_predictor_type="deterministic", encoder_input_layer="linear", duration_predictor_use_encoder_outputs=False,
                 encoder_normalize_before=True, decoder_normalize_before=True)
    z = np.load(os.path.join(os.path.dirname(GOLDEN), FIXTURES["conv1d"]))
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd.")}
    m = AASVC(**AAS_HP, **fixed)
    msd = m.state_dict()
    assert set(msd) == set(sd) and all(tuple(msd[k].shape) == tuple(sd[k].shape) for k in sd)
    assert tuple(msd["decoder.encoders.0.feed_forward_macaron.w_2.weight"].shape) == (128, 48, 1)
    m.load_state_dict(sd)
    assert torch.equal(m.state_dict()["encoder.encoders.0.feed_forward.w_1.weight"], sd["encoder.encoders.0.feed_forward.w_1.weight"])
    m3 = AASVC(**AAS_HP, **fixed, positionwise_conv_kernel_size=3)          # kernel size 3: same keys and shapes as the reference's
    z3 = np.load(os.path.join(os.path.dirname(GOLDEN), FIXTURES["conv1d:3"]))
    assert {k: tuple(v.shape) for k, v in m3.state_dict().items()} == {k[3:]: z3[k].shape for k in z3.files if k.startswith("sd.")}
    with pytest.raises(NotImplementedError):
        AASVC(**AAS_HP, **fixed, positionwise_conv_kernel_size=4)          # even kernels shorten the sequence in the reference


def test_length_regulator_host_logic(monkeypatch):
    """LengthRegulator drop-in (length_regulator.py:46-97) with the kernels replaced by their CPU contracts vs the reference module."""
    from oracle import ref_shim

    if not ref_shim.available():
        pytest.skip("reference tree not present")
    fake_ops.install(monkeypatch)
    ref_shim.install()
    from seq2seq_vc.modules.length_regulator import LengthRegulator as RefLR
    from seq2seq_vc_b200 import LengthRegulator

    g = torch.Generator().manual_seed(3)
    xs = torch.randn(3, 9, 8, generator=g)
    ds = torch.randint(0, 4, (3, 9), generator=g)
    for alpha, pad in ((1.0, 0.0), (1.7, 2.0)):
        a = xs.clone().requires_grad_(True)
        b = xs.clone().requires_grad_(True)
        got, want = LengthRegulator(pad)(a, ds, alpha), RefLR(pad)(b, ds.clone(), alpha)
        assert torch.equal(got, want)
        w = torch.randn(want.shape, generator=g)
        (got * w).sum().backward()
        (want * w).sum().backward()
        assert torch.allclose(a.grad, b.grad, atol=1e-6)
    assert torch.equal(LengthRegulator()(xs, torch.zeros(3, 9, dtype=torch.long)), RefLR()(xs, torch.zeros(3, 9, dtype=torch.long)))
