"""Host-side orchestration of the VTN engine checked on CPU against the reference's golden vectors.

The C-ABI kernels are replaced by their torch-CPU contracts (tests/fake_ops.py) -- this verifies
buffer wiring, strided GEMM operand views, weight packing and the hand-written backward pass.
The kernels themselves are verified on the GPU (tests/test_gpu_*.py).
"""
import os

import numpy as np
import pytest
import torch

import fake_ops
from seq2seq_vc_b200.vtn_engine import VTNEngine

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "vtn_tiny.npz")
TINY_HP = dict(idim=80, odim=80, dprenet_layers=2, dprenet_units=16, adim=32, aheads=2, elayers=1, eunits=48,
               dlayers=2, dunits=48, postnet_layers=3, postnet_filts=5, postnet_chans=16, decoder_reduction_factor=2)
NO_DROPOUT = dict(dprenet_dropout_rate=0.0, transformer_enc_dropout_rate=0.0, enc_positional_dropout_rate=0.0,
                  dec_dropout_rate=0.0, dec_positional_dropout_rate=0.0, postnet_dropout_rate=0.0)


def load_golden():
    z = np.load(GOLDEN)
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd.")}
    return z, sd


@pytest.fixture()
def engine(monkeypatch):
    fake_ops.install(monkeypatch)
    z, sd = load_golden()
    eng = VTNEngine(dict(TINY_HP, **NO_DROPOUT), device="cpu", bf16=False)
    eng.load_state_dict(sd)
    return eng, z


def run_step(eng, z):
    ilens, olens = z["ilens"].tolist(), z["olens"].tolist()
    dev = eng.device
    xs = torch.from_numpy(z["xs"])[:, :max(ilens)].contiguous().to(dev)
    ys = torch.from_numpy(z["ys"])[:, :max(olens)].contiguous().to(dev)
    labels = torch.from_numpy(z["labels"])[:, :max(olens)].contiguous().to(dev)
    after, before, logits = eng.forward(xs, ys, ilens, olens)
    losses = eng.loss(ys, labels)
    eng.backward(eng.d_after, eng.d_before, eng.d_logits)
    return after, before, logits, losses


def test_forward_matches_reference(engine):
    eng, z = engine
    after, before, logits, losses = run_step(eng, z)
    assert np.abs(after.numpy() - z["after_outs"]).mean() <= 1e-5
    assert np.abs(before.numpy() - z["before_outs"]).mean() <= 1e-5
    assert np.abs(logits.numpy() - z["logits"]).mean() <= 1e-5
    assert abs(float(losses[0]) - float(z["l1_loss"])) <= 1e-5
    assert abs(float(losses[1]) - float(z["bce_loss"])) <= 1e-5
    np.testing.assert_array_equal(eng.labels_fix.numpy(), z["labels_out"])
    assert eng.olens_fix_host == z["olens_out"].tolist()
    assert eng.ilens_ds_st == z["ilens_ds_st"].tolist()
    assert eng.olens_in_host == z["olens_in"].tolist()
    nl = eng.hp["dlayers"]
    for i in range(nl):   # reference returns src-attention maps last layer first (vtn.py:280-287)
        got = eng.attn[f"decoder.decoders.{nl - 1 - i}.src_attn"].numpy()
        assert np.abs(got - z[f"att_ws.{i}"]).mean() <= 1e-6


def test_gradients_match_reference(engine):
    eng, z = engine
    run_step(eng, z)
    worst = ("", 0.0)
    for name in eng.store.names():
        ref = z["grad." + name]
        got = eng.store.g(name).numpy()
        # linear_k.bias has a mathematically zero gradient (softmax shift invariance): absolute floor
        scale = np.abs(ref).max() + 1e-5
        err = np.abs(got - ref).max() / scale
        if err > worst[1]:
            worst = (name, err)
        assert err <= 2e-4, (name, err)
    print("worst relative grad error", worst)


@pytest.mark.parametrize("r", [1, 3, 4])
def test_reduction_factors_forward_and_gradients(monkeypatch, r):
    """decoder_reduction_factor 1, 3 and 4 (recipe: egs/arctic/vc1/conf/vtn.v1.yaml:43), ragged olens not divisible by r."""
    fake_ops.install(monkeypatch)
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", f"vtn_r{r}_tiny.npz"))
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd.")}
    eng = VTNEngine(dict(TINY_HP, **NO_DROPOUT, decoder_reduction_factor=r), device="cpu", bf16=False)
    eng.load_state_dict(sd)
    check_reduction_factor(eng, z, 1e-5, 2e-4)


@pytest.mark.parametrize("layer", ["conv1d", "conv1d-linear"])
def test_conv_positionwise_encoder_forward_and_gradients(monkeypatch, layer):
    """Transformer-encoder VTN with MultiLayeredConv1d / Conv1dLinear (kernel size 3) position-wise layers vs the live-reference dump."""
    fake_ops.install(monkeypatch)
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", f"vtn_{layer.replace('-', '_')}_k3_tiny.npz"))
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd.")}
    eng = VTNEngine(dict(TINY_HP, **NO_DROPOUT, elayers=2, positionwise_layer_type=layer, positionwise_conv_kernel_size=3), device="cpu", bf16=False)
    assert set(eng.state_dict()) == set(sd)
    eng.load_state_dict(sd)
    check_reduction_factor(eng, z, 1e-5, 2e-4)


@pytest.mark.parametrize("rel", ["legacy", "latest"])
def test_conformer_encoder_forward_and_gradients(monkeypatch, rel):
    """VTN(encoder_type="conformer") (models/vtn.py:83-143) with the class-default legacy rel-pos attention and with
    conformer_rel_pos_type="latest": outputs, losses, encoder attention maps, BatchNorm running statistics and every gradient vs
    the live-reference dump; eval mode on the updated statistics."""
    fake_ops.install(monkeypatch)
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", f"vtn_conformer_{rel}_tiny.npz"))
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd.")}
    eng = VTNEngine(dict(TINY_HP, **NO_DROPOUT, elayers=2, encoder_type="conformer", conformer_rel_pos_type=rel, enc_attn_dropout_rate=0.0),
                    device="cpu", bf16=False)
    assert set(eng.state_dict()) == set(sd)
    eng.load_state_dict(sd)
    check_conformer(eng, z, 1e-5, 2e-4)


def check_conformer(eng, z, tol_out, tol_grad):
    check_reduction_factor(eng, z, tol_out, tol_grad)
    for k in [k for k in z.files if k.startswith("attn.encoder.")]:
        assert np.abs(eng.attn[k[5:]].float().cpu().numpy() - z[k]).mean() <= 1e-3 if tol_out > 1e-5 else 1e-6, k
    for k in [k for k in z.files if k.startswith("bn_after.")]:
        np.testing.assert_allclose(eng.buffers[k[9:]].float().cpu().numpy(), z[k], rtol=1e-4, atol=1e-5)
    eng.training = False
    ilens, olens = z["ilens"].tolist(), z["olens"].tolist()
    dev = eng.device
    after_e, _, _ = eng.forward(torch.from_numpy(z["xs"])[:, :max(ilens)].contiguous().to(dev),
                                torch.from_numpy(z["ys"])[:, :max(olens)].contiguous().to(dev), ilens, olens)
    assert np.abs(after_e.float().cpu().numpy() - z["eval_after_outs"]).mean() <= tol_out


def check_reduction_factor(eng, z, tol_out, tol_grad):
    after, before, logits, losses = run_step(eng, z)
    assert after.shape == z["after_outs"].shape
    assert np.abs(after.float().cpu().numpy() - z["after_outs"]).mean() <= tol_out
    assert np.abs(before.float().cpu().numpy() - z["before_outs"]).mean() <= tol_out
    assert np.abs(logits.float().cpu().numpy() - z["logits"]).mean() <= tol_out
    assert abs(float(losses[0]) - float(z["l1_loss"])) <= 10 * tol_out and abs(float(losses[1]) - float(z["bce_loss"])) <= 10 * tol_out
    np.testing.assert_array_equal(eng.labels_fix.cpu().numpy(), z["labels_out"])
    assert eng.olens_fix_host == z["olens_out"].tolist() and eng.olens_in_host == z["olens_in"].tolist()
    nl = eng.hp["dlayers"]
    for i in range(nl):
        got = eng.attn[f"decoder.decoders.{nl - 1 - i}.src_attn"].float().cpu().numpy()
        assert np.abs(got - z[f"att_ws.{i}"]).mean() <= 1e-3
    for name in eng.store.names():
        ref = z["grad." + name]
        got = eng.store.g(name).cpu().numpy()
        assert np.abs(got - ref).max() <= tol_grad * (np.abs(ref).max() + 1e-5) + 1e-7, name


def test_bn_running_stats(engine):
    eng, z = engine
    run_step(eng, z)
    for k in z.files:
        if k.startswith("bn_after."):
            np.testing.assert_allclose(eng.buffers[k[9:]].numpy(), z[k], rtol=1e-4, atol=1e-6)


def test_eval_mode_forward(engine):
    eng, z = engine
    run_step(eng, z)            # training step updates the running stats first (as in gen_golden)
    eng.training = False
    ilens, olens = z["ilens"].tolist(), z["olens"].tolist()
    xs = torch.from_numpy(z["xs"])[:, :max(ilens)].contiguous()
    ys = torch.from_numpy(z["ys"])[:, :max(olens)].contiguous()
    after, _, _ = eng.forward(xs, ys, ilens, olens)
    assert np.abs(after.numpy() - z["eval_after_outs"]).mean() <= 1e-5


def test_adam_step_matches_torch(engine):
    eng, z = engine
    run_step(eng, z)
    p0 = eng.store.P.clone()
    g = eng.store.G.clone()
    ref = torch.nn.Parameter(p0.clone())
    ref.grad = g.clone()
    opt = torch.optim.Adam([ref], lr=8e-5)
    torch.nn.utils.clip_grad_norm_([ref], 1.0)
    opt.step()
    eng.lr_dev.fill_(8e-5)
    eng.optimizer_step(1.0)
    assert (eng.store.P - ref.detach()).abs().max() <= 1e-7


# --------------------------------------------------------------------------------------------------
# TransformerTTS (token-embedding encoder) + guided attention loss
# --------------------------------------------------------------------------------------------------
TTS_HP = dict(idim=40, odim=80, dprenet_layers=2, dprenet_units=16, adim=32, aheads=4, elayers=1, eunits=48,
              dlayers=2, dunits=48, postnet_layers=2, postnet_filts=5, postnet_chans=16, decoder_reduction_factor=2,
              encoder_input="embed")


def run_tts_step(eng, z):
    ilens, olens = z["ilens"].tolist(), z["olens"].tolist()
    dev = eng.device
    tokens = torch.from_numpy(z["tokens"])[:, :max(ilens)].contiguous().to(dev)
    ys = torch.from_numpy(z["ys"])[:, :max(olens)].contiguous().to(dev)
    labels = torch.from_numpy(z["labels"])[:, :max(olens)].contiguous().to(dev)
    after, before, logits = eng.forward(tokens, ys, ilens, olens)
    losses = eng.loss(ys, labels)
    ga, d_att = eng.guided_attention(sigma=0.4, alpha=1.0, n_layers=2, n_heads=2)
    eng.backward(eng.d_after, eng.d_before, eng.d_logits, d_att=d_att)
    return after, before, logits, losses, ga


def check_tts(eng, z, tol_out=1e-5, tol_grad=2e-4):
    after, before, logits, losses, ga = run_tts_step(eng, z)
    assert np.abs(after.cpu().numpy() - z["after_outs"]).mean() <= tol_out
    assert np.abs(logits.cpu().numpy() - z["logits"]).mean() <= tol_out
    assert abs(float(losses[0]) - float(z["l1_loss"])) <= 10 * tol_out and abs(float(losses[1]) - float(z["bce_loss"])) <= 10 * tol_out
    assert abs(float(ga) - float(z["ga_loss"])) <= 10 * tol_out
    assert eng.ilens_ds_st == z["ilens_out"].tolist() and eng.olens_in_host == z["olens_in"].tolist()
    nl = eng.hp["dlayers"]
    att = torch.cat([eng.attn[f"decoder.decoders.{l}.src_attn"][:, :2] for l in reversed(range(nl))][:2], dim=1)
    assert np.abs(att.float().cpu().numpy() - z["att_ws"]).mean() <= 1e-5
    for name in eng.store.names():
        ref = z["grad." + name]
        got = eng.store.g(name).cpu().numpy()
        assert np.abs(got - ref).max() <= tol_grad * (np.abs(ref).max() + 1e-5) + 1e-7, name     # 1e-7: gradients that are exactly 0 in
        # exact arithmetic (key biases: softmax is shift-invariant) are pure rounding noise on either side


def test_tts_forward_guided_attention_and_gradients(monkeypatch):
    fake_ops.install(monkeypatch)
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "tts_tiny.npz"))
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd.")}
    eng = VTNEngine(dict(TTS_HP, **NO_DROPOUT), device="cpu", bf16=False)
    eng.load_state_dict(sd)
    check_tts(eng, z)


def test_autoregressive_inference_matches_reference(engine):
    """VTN.inference (models/vtn.py:302-394): outputs, stop probabilities and source-attention maps vs the live-reference dump."""
    eng, z = engine
    sd = eng.state_dict()
    for k in z.files:
        if k.startswith("bn_after."):
            sd[k[9:]] = torch.from_numpy(z[k])
    eng.load_state_dict(sd)
    il = int(z["ilens"][0])
    outs, probs, att = eng.inference(torch.from_numpy(z["xs"])[0, :il], threshold=0.9999, minlenratio=0.0, maxlenratio=1.6)
    assert outs.shape == z["inf_outs"].shape and probs.shape == z["inf_probs"].shape and att.shape == z["inf_att_ws"].shape
    assert np.abs(outs.numpy() - z["inf_outs"]).mean() <= 1e-5
    assert np.abs(probs.numpy() - z["inf_probs"]).max() <= 1e-5
    assert np.abs(att.numpy() - z["inf_att_ws"]).max() <= 1e-5
    assert eng.training is True
    # the cache-free recomputation path gives the same answer
    o2, p2, a2 = eng.inference_recompute(torch.from_numpy(z["xs"])[0, :il], threshold=0.9999, minlenratio=0.0, maxlenratio=1.6)
    assert np.abs(o2.numpy() - z["inf_outs"]).mean() <= 1e-5 and np.abs(a2.numpy() - z["inf_att_ws"]).max() <= 1e-5


def test_tts_inference_matches_reference(monkeypatch):
    """TransformerTTS.inference (<eos> append + token embedding + KV-cache decode) vs the live-reference dump."""
    fake_ops.install(monkeypatch)
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "tts_tiny.npz"))
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd.")}
    sd.update({k[9:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("bn_after.")})
    eng = VTNEngine(dict(TTS_HP, **NO_DROPOUT), device="cpu", bf16=False)
    eng.load_state_dict(sd)
    il = int(z["ilens"][0])
    for fn in (eng.inference, eng.inference_recompute):
        outs, probs, att = fn(torch.from_numpy(z["tokens"])[0, :il], threshold=0.9999, minlenratio=0.0, maxlenratio=1.5)
        assert outs.shape == z["inf_outs"].shape and att.shape == z["inf_att_ws"].shape
        assert np.abs(outs.numpy() - z["inf_outs"]).mean() <= 1e-5 and np.abs(probs.numpy() - z["inf_probs"]).max() <= 1e-5
        assert np.abs(att.numpy() - z["inf_att_ws"]).max() <= 1e-5


def test_conformer_dropin_matches_reference_registration(monkeypatch):
    """seq2seq_vc_b200.VTN(encoder_type="conformer") registers the reference's parameters in the reference's ORDER (optimizer
    state dicts are keyed by order) with its state-dict keys and shapes, loads the reference's state dict and reproduces its
    forward through the drop-in path (CPU contracts)."""
    from oracle import ref_shim

    if not ref_shim.available():
        pytest.skip("reference tree not present")
    fake_ops.install(monkeypatch)
    ref_shim.install()
    from seq2seq_vc.models import VTN as RefVTN
    from seq2seq_vc_b200 import VTN

    kw = dict(TINY_HP, elayers=2, encoder_type="conformer", conformer_enc_kernel_size=7, dprenet_dropout_rate=0.0)
    torch.manual_seed(3)
    ref = RefVTN(**kw)
    ref_shim.disable_dropout(ref)
    ours = VTN(**kw, transformer_enc_dropout_rate=0.0, transformer_enc_positional_dropout_rate=0.0, transformer_enc_attn_dropout_rate=0.0)
    assert [n for n, _ in ours.named_parameters()] == [n for n, _ in ref.named_parameters()]
    assert {k: tuple(v.shape) for k, v in ours.state_dict().items()} == {k: tuple(v.shape) for k, v in ref.state_dict().items()}
    ours.load_state_dict(ref.state_dict())
    ours.engine.hp.update({k: 0.0 for k in ours.engine.hp if "dropout" in k})
    from oracle import vtn_oracle

    xs, ilens, ys, labels, olens = vtn_oracle.synthetic_batch(2, 44, 30, ilens=[44, 37], olens=[30, 23], seed=5)
    ref.train()
    ours.train()
    a = ref(xs, torch.tensor(ilens), ys, labels, torch.tensor(olens))
    b = ours(xs, torch.tensor(ilens), ys, labels, torch.tensor(olens))
    for i in range(3):
        assert (a[i] - b[i]).abs().max().item() <= 2e-5
    assert torch.equal(a[4], b[4]) and torch.equal(a[5], b[5])
    with pytest.raises(NotImplementedError):
        VTN(**dict(kw, conformer_self_attn_layer_type="selfattn"))


@pytest.mark.parametrize("rel", ["legacy", "latest"])
def test_conformer_encoder_inference_matches_oracle(monkeypatch, rel):
    """Autoregressive inference (models/vtn.py:302-394) with the conformer encoder: the oracle vs the live-reference dump, then the
    KV-cache decode and the prefix-recomputing cross-check vs the oracle."""
    from oracle import vtn_oracle

    fake_ops.install(monkeypatch)
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", f"vtn_conformer_{rel}_tiny.npz"))
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd.")}
    hp = dict(TINY_HP, elayers=2, encoder_type="conformer", conformer_rel_pos_type=rel)
    eng = VTNEngine(dict(hp, **NO_DROPOUT, enc_attn_dropout_rate=0.0), device="cpu", bf16=False)
    eng.load_state_dict(sd)
    il = int(z["ilens"][0])
    x = torch.from_numpy(z["xs"])[0, :il]
    sd_eval = {**sd, **{k[9:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("bn_after.")}}     # the dump ran after a training step
    eng.load_state_dict(sd_eval)
    for k in z.files:
        if k.startswith("bn_after."):
            eng.buffers[k[9:]].copy_(torch.from_numpy(z[k]))
    ref_o, ref_p, ref_a = vtn_oracle.vtn_inference(sd_eval, vtn_oracle.default_hparams(**hp), x, threshold=0.9999, minlenratio=0.0, maxlenratio=1.3)
    assert ref_o.shape == z["inf_outs"].shape                                 # oracle == live reference
    assert np.abs(ref_o.numpy() - z["inf_outs"]).mean() <= 1e-5 and np.abs(ref_p.numpy() - z["inf_probs"]).mean() <= 1e-5
    assert np.abs(ref_a.numpy() - z["inf_att_ws"]).mean() <= 1e-6
    for fn in (eng.inference, eng.inference_recompute):
        outs, probs, att = fn(x, threshold=0.9999, minlenratio=0.0, maxlenratio=1.3)
        assert outs.shape == ref_o.shape
        assert (outs - ref_o).abs().mean().item() <= 1e-5 and (probs - ref_p).abs().mean().item() <= 1e-5
        assert (att - ref_a).abs().mean().item() <= 1e-6


def test_train_step_and_engine_form_no_reference_cycle(monkeypatch):
    """A train step registers its shape-eviction callback weakly, so dropping the step and its engine frees both by reference
    counting alone: a dead engine <-> step cycle would keep CUDA graphs and device memory alive until some later cyclic collection,
    and a collection that triggers inside another stream capture invalidates it.  Captures themselves run with the collector off
    (_lib.no_gc) -- torch.cuda.graph() does not collect any more."""
    import gc
    import weakref

    from seq2seq_vc_b200 import VTNTrainStep, _lib

    fake_ops.install(monkeypatch)
    gc.collect()
    was = gc.isenabled()
    gc.disable()
    try:
        eng = VTNEngine(dict(TINY_HP), device="cpu", bf16=False)
        step = VTNTrainStep(eng, lr=1e-3, warmup_steps=1)
        seen = []
        eng._evict_listeners.append(lambda sig: seen.append(sig))
        eng._evict_sig((1, 2, 3, True))                 # weak listeners are resolved and called
        assert seen == [(1, 2, 3, True)]
        w_eng, w_step = weakref.ref(eng), weakref.ref(step)
        del step
        assert w_step() is None                         # the engine does not keep the step alive
        eng._evict_sig((1, 2, 3, True))                 # ... and drops the dead listener
        assert not any(isinstance(cb, weakref.WeakMethod) for cb in eng._evict_listeners)
        del eng
        assert w_eng() is None
        with _lib.no_gc():
            assert not gc.isenabled()
        assert not gc.isenabled()                       # restored to what it was (off, here)
    finally:
        if was:
            gc.enable()
    with _lib.no_gc():
        assert not gc.isenabled()
    assert gc.isenabled() == was
