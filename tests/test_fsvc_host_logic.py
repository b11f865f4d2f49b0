"""Host-side orchestration of the FastSpeechVC engine on CPU against the live-reference golden vectors (kernels replaced by their
torch-CPU contracts, tests/fake_ops.py); the kernels themselves are verified on the GPU (tests/test_gpu_fsvc.py)."""
import os

import numpy as np
import pytest
import torch

import fake_ops
from seq2seq_vc_b200.fsvc_engine import FastSpeechVCEngine

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "fsvc_tiny.npz")
FS_HP = dict(idim=80, odim=80, adim=32, aheads=2, elayers=2, eunits=48, dlayers=2, dunits=48, duration_predictor_input_dim=80,
             duration_predictor_layers=2, duration_predictor_chans=16, duration_predictor_kernel_size=3, postnet_layers=2, postnet_filts=5,
             postnet_chans=16, conformer_enc_kernel_size=7, conformer_dec_kernel_size=7)
NO_DROPOUT = dict(transformer_enc_dropout_rate=0.0, transformer_enc_positional_dropout_rate=0.0, transformer_enc_attn_dropout_rate=0.0,
                  transformer_dec_dropout_rate=0.0, transformer_dec_positional_dropout_rate=0.0, transformer_dec_attn_dropout_rate=0.0,
                  duration_predictor_dropout_rate=0.0, postnet_dropout_rate=0.0)


def load():
    z = np.load(GOLDEN)
    return z, {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd.")}


def run_step(eng, z):
    ilens, olens = z["ilens"].tolist(), z["olens"].tolist()
    dev = eng.device
    xs = torch.from_numpy(z["xs"])[:, :max(ilens)].contiguous().to(dev)
    ys = torch.from_numpy(z["ys"])[:, :max(olens)].contiguous().to(dev)
    dpi = torch.from_numpy(z["dp_inputs"])[:, :max(ilens)].contiguous().to(dev)
    ds = torch.from_numpy(z["ds"]).to(dev)
    after, before = eng.forward(xs, ys, ds, dpi, ilens, olens)
    losses = eng.loss(ys)
    eng.backward()
    return after, before, losses


def check_against_golden(eng, z, tol_out, tol_grad_mean, tol_grad_max):
    after, before, losses = run_step(eng, z)
    assert np.abs(after.float().cpu().numpy() - z["after_outs"]).mean() <= tol_out
    assert np.abs(before.float().cpu().numpy() - z["before_outs"]).mean() <= tol_out
    assert np.abs(eng.forward_d_outs().cpu().numpy() - z["d_outs"]).max() <= 20 * tol_out
    assert eng.tlens_host == z["ilens_out"].tolist()
    for i, k in enumerate(("l1_loss", "duration_loss")):
        assert abs(losses[i].item() - float(z[k])) <= 10 * tol_out * max(1.0, abs(float(z[k]))), k
    for k in [k for k in z.files if k.startswith("attn.")]:
        assert np.abs(eng.attn[k[5:]].float().cpu().numpy() - z[k]).mean() <= 1e-3, k
    gmax = max(np.abs(z[k]).max() for k in z.files if k.startswith("grad."))
    for name in eng.store.names():
        ref = z["grad." + name]
        got = eng.store.g(name).cpu().numpy()
        assert np.abs(got - ref).mean() <= tol_grad_mean * np.abs(ref).mean() + 1e-6 * gmax, name
        assert np.abs(got - ref).max() <= tol_grad_max * np.abs(ref).max() + 1e-5 * gmax, name
    for k in z.files:
        if k.startswith("bn_after."):
            np.testing.assert_allclose(eng.buffers[k[9:]].cpu().numpy(), z[k], rtol=1e-4, atol=1e-5)
    eng.training = False
    ilens, olens = z["ilens"].tolist(), z["olens"].tolist()
    dev = eng.device
    after_e, _ = eng.forward(torch.from_numpy(z["xs"]).to(dev), torch.from_numpy(z["ys"]).to(dev), torch.from_numpy(z["ds"]).to(dev),
                             torch.from_numpy(z["dp_inputs"]).to(dev), ilens, olens)
    assert np.abs(after_e.float().cpu().numpy() - z["eval_after_outs"]).mean() <= tol_out


def test_forward_losses_gradients_match_reference(monkeypatch):
    fake_ops.install(monkeypatch)
    z, sd = load()
    eng = FastSpeechVCEngine(dict(FS_HP, **NO_DROPOUT), device="cpu", bf16=False)
    assert set(eng.state_dict()) == set(sd)
    eng.load_state_dict(sd)
    check_against_golden(eng, z, 1e-5, 2e-4, 2e-4)


def test_inference_matches_oracle_pipeline(monkeypatch):
    """Predicted durations -> LengthRegulator -> decoder (fastspeech_vc.py:427-470) vs the same pipeline assembled from the oracle."""
    from oracle import aasvc_oracle as ao
    from oracle import fsvc_oracle as fo

    fake_ops.install(monkeypatch)
    z, sd = load()
    sd = {**sd, **{k[9:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("bn_after.")}}
    sd["duration_predictor.linear.bias"] = sd["duration_predictor.linear.bias"] + 1.0      # predicted durations not all zero
    eng = FastSpeechVCEngine(dict(FS_HP, **NO_DROPOUT), device="cpu", bf16=False)
    eng.load_state_dict(sd)
    il = int(z["ilens"][0])
    x = torch.from_numpy(z["xs"])[0, :il]
    outs, d_outs = eng.inference(x, x)
    hp = fo.default_hparams(**FS_HP)
    # oracle: duration predictor inference = clamp(round(exp(pre) - 1), min 0) on the projected side input, then the teacher-forced path
    T2 = ((il - 1) // 2 - 1) // 2
    dpi = ao.dp_projection(sd, "duration_predictor_projection", x[None], T2)
    pre = ao.duration_predictor(sd, "duration_predictor", dict(hp), dpi, [T2], clamp=False)
    ds = torch.clamp(torch.round(torch.exp(pre) - 1.0), min=0).long()
    assert torch.equal(ds[0].float(), d_outs)
    L = int(ds.sum())
    ref = fo.fsvc_forward(sd, FS_HP, x[None], [il], torch.zeros(1, L, 80), [L], ds, x[None], training=False)
    assert outs.shape == (L, 80) and (outs - ref["after_outs"][0]).abs().mean().item() <= 1e-5


def test_dropin_matches_live_reference(monkeypatch):
    """seq2seq_vc_b200.FastSpeechVC registers the reference's parameters in the reference's ORDER with its state-dict keys and shapes,
    loads the reference's state dict, and the NARVCTrainer step written against it (trainers/nar_vc.py:52-96: model -> L1Loss +
    DurationPredictorLoss -> backward) reproduces the reference's outputs, losses and every parameter gradient (CPU contracts)."""
    from oracle import ref_shim

    if not ref_shim.available():
        pytest.skip("reference tree not present")
    fake_ops.install(monkeypatch)
    ref_shim.install()
    from seq2seq_vc.losses import DurationPredictorLoss as RefDurLoss
    from seq2seq_vc.losses import L1Loss as RefL1
    from seq2seq_vc.models.fastspeech_vc import FastSpeechVC as RefFS
    from seq2seq_vc_b200 import DurationPredictorLoss, FastSpeechVC, L1Loss

    kw = dict(FS_HP, encoder_type="conformer", decoder_type="conformer", encoder_input_layer="conv2d", positionwise_layer_type="linear",
              duration_predictor_use_encoder_outputs=False, encoder_normalize_before=True, decoder_normalize_before=True,
              teacher_model_decoder_reduction_factor=1)
    torch.manual_seed(11)
    ref = RefFS(**kw)
    ref_shim.disable_dropout(ref)
    ours = FastSpeechVC(**kw, **NO_DROPOUT)
    assert [n for n, _ in ours.named_parameters()] == [n for n, _ in ref.named_parameters()]
    assert {k: tuple(v.shape) for k, v in ours.state_dict().items()} == {k: tuple(v.shape) for k, v in ref.state_dict().items()}
    ours.load_state_dict(ref.state_dict())
    z, _ = load()
    ilens, olens = torch.from_numpy(z["ilens"]), torch.from_numpy(z["olens"])
    xs, ys, dpi, ds = (torch.from_numpy(z[k]) for k in ("xs", "ys", "dp_inputs", "ds"))
    dlens = torch.from_numpy(z["ilens_out"])
    ref.train()
    ours.train()
    a = ref(xs, ilens, ys, olens, ds, dlens, dpi, dp_lengths=ilens)
    b = ours(xs, ilens, ys, olens, ds, dlens, dpi, dp_lengths=ilens)
    for i in range(3):
        assert (a[i] - b[i]).abs().max().item() <= 3e-5, i
    assert torch.equal(a[3], b[3]) and torch.equal(a[4], b[4]) and torch.equal(a[5], b[5])
    la = RefL1()(a[1], a[0], a[5], a[4]) + RefDurLoss()(a[2], ds, a[3])
    lb = L1Loss()(b[1], b[0], b[5], b[4]) + DurationPredictorLoss()(b[2], ds, b[3])
    assert abs(la.item() - lb.item()) <= 1e-5 * max(1.0, abs(la.item()))
    la.backward()
    lb.backward()
    gref = dict(ref.named_parameters())
    gmax = max(p.grad.abs().max().item() for p in gref.values() if p.grad is not None)
    for n, p in ours.named_parameters():
        r = gref[n].grad
        assert r is not None and p.grad is not None, n
        assert (p.grad - r).abs().max().item() <= 2e-3 * r.abs().max().item() + 1e-5 * gmax, n
    with pytest.raises(NotImplementedError):
        FastSpeechVC(**dict(kw, encoder_type="transformer"))
    with pytest.raises(NotImplementedError):
        FastSpeechVC(**dict(kw, duration_predictor_use_encoder_outputs=True))


def test_dropin_teacher_reduction_factor(monkeypatch):
    """teacher_model_decoder_reduction_factor = 2: the regulated length is 2 * sum(ds) (fastspeech_vc.py:276-279)."""
    from oracle import ref_shim

    if not ref_shim.available():
        pytest.skip("reference tree not present")
    fake_ops.install(monkeypatch)
    ref_shim.install()
    from seq2seq_vc.models.fastspeech_vc import FastSpeechVC as RefFS
    from seq2seq_vc_b200 import FastSpeechVC

    kw = dict(FS_HP, encoder_type="conformer", decoder_type="conformer", encoder_input_layer="conv2d", positionwise_layer_type="linear",
              duration_predictor_use_encoder_outputs=False, encoder_normalize_before=True, decoder_normalize_before=True,
              teacher_model_decoder_reduction_factor=2)
    torch.manual_seed(12)
    ref = RefFS(**kw)
    ref_shim.disable_dropout(ref)
    ours = FastSpeechVC(**kw, **NO_DROPOUT)
    ours.load_state_dict(ref.state_dict())
    z, _ = load()
    ilens = torch.from_numpy(z["ilens"])
    xs, dpi, ds = (torch.from_numpy(z[k]) for k in ("xs", "dp_inputs", "ds"))
    olens = 2 * ds.sum(1)
    ys = torch.randn(xs.shape[0], int(olens.max()), 80, generator=torch.Generator().manual_seed(1))
    dlens = torch.from_numpy(z["ilens_out"])
    ref.train()
    ours.train()
    a = ref(xs, ilens, ys, olens, ds, dlens, dpi, dp_lengths=ilens)
    b = ours(xs, ilens, ys, olens, ds, dlens, dpi, dp_lengths=ilens)
    for i in range(3):
        assert a[i].shape == b[i].shape and (a[i] - b[i]).abs().max().item() <= 3e-5, i
    with pytest.raises(Exception):
        ours(xs, ilens, ys[:, :-1], olens - 1, ds, dlens, dpi, dp_lengths=ilens)          # sum of durations != target length


@pytest.mark.parametrize("pw,k", [("conv1d", 1), ("conv1d", 3), ("conv1d-linear", 3)])
def test_dropin_positionwise_variants_match_live_reference(monkeypatch, pw, k):
    """FastSpeechVC with the MultiLayeredConv1d / Conv1dLinear position-wise layers (multi_layer_conv.py:12-108) in both conformer
    stacks: same registration order, outputs and every parameter gradient as the live reference (CPU contracts)."""
    from oracle import ref_shim

    if not ref_shim.available():
        pytest.skip("reference tree not present")
    fake_ops.install(monkeypatch)
    ref_shim.install()
    from seq2seq_vc.models.fastspeech_vc import FastSpeechVC as RefFS
    from seq2seq_vc_b200 import FastSpeechVC

    kw = dict(FS_HP, encoder_type="conformer", decoder_type="conformer", encoder_input_layer="conv2d", positionwise_layer_type=pw,
              positionwise_conv_kernel_size=k, duration_predictor_use_encoder_outputs=False, encoder_normalize_before=True,
              decoder_normalize_before=True, teacher_model_decoder_reduction_factor=1)
    torch.manual_seed(21 + k)
    ref = RefFS(**kw)
    ref_shim.disable_dropout(ref)
    ours = FastSpeechVC(**kw, **NO_DROPOUT)
    assert [n for n, _ in ours.named_parameters()] == [n for n, _ in ref.named_parameters()]
    ours.load_state_dict(ref.state_dict())
    z, _ = load()
    ilens, olens = torch.from_numpy(z["ilens"]), torch.from_numpy(z["olens"])
    xs, ys, dpi, ds = (torch.from_numpy(z[k_]) for k_ in ("xs", "ys", "dp_inputs", "ds"))
    dlens = torch.from_numpy(z["ilens_out"])
    ref.train()
    ours.train()
    a = ref(xs, ilens, ys, olens, ds, dlens, dpi, dp_lengths=ilens)
    b = ours(xs, ilens, ys, olens, ds, dlens, dpi, dp_lengths=ilens)
    for i in range(3):
        assert (a[i] - b[i]).abs().max().item() <= 5e-5, i
    r = torch.randn(a[1].shape, generator=torch.Generator().manual_seed(2))
    ((a[1] * r).sum() + a[2].sum()).backward()
    ((b[1] * r).sum() + b[2].sum()).backward()
    gref = dict(ref.named_parameters())
    gmax = max(p.grad.abs().max().item() for p in gref.values() if p.grad is not None)
    for n, p in ours.named_parameters():
        rg = gref[n].grad
        assert rg is not None and p.grad is not None, n
        assert (p.grad - rg).abs().max().item() <= 2e-3 * rg.abs().max().item() + 1e-5 * gmax, n
