"""FastSpeechVC hot path on the GPU through the C ABI: the whole engine vs golden vectors dumped from the live reference
(tests/golden/fsvc_tiny.npz, oracle/gen_golden.py fsvc_tiny) and vs the CPU oracle.  Tolerances: mel L1 <= 1e-4, attention-weight
L1 <= 1e-3 (fp32 path); the duration / length-regulator path is integer work and is bit-exact."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "fsvc_tiny.npz")
FS_HP = dict(idim=80, odim=80, adim=32, aheads=2, elayers=2, eunits=48, dlayers=2, dunits=48, duration_predictor_input_dim=80,
             duration_predictor_layers=2, duration_predictor_chans=16, duration_predictor_kernel_size=3, postnet_layers=2, postnet_filts=5,
             postnet_chans=16, conformer_enc_kernel_size=7, conformer_dec_kernel_size=7)
NO_DROPOUT = dict(transformer_enc_dropout_rate=0.0, transformer_enc_positional_dropout_rate=0.0, transformer_enc_attn_dropout_rate=0.0,
                  transformer_dec_dropout_rate=0.0, transformer_dec_positional_dropout_rate=0.0, transformer_dec_attn_dropout_rate=0.0,
                  duration_predictor_dropout_rate=0.0, postnet_dropout_rate=0.0)
FIXED = dict(encoder_type="conformer", decoder_type="conformer", encoder_input_layer="conv2d", positionwise_layer_type="linear",
             duration_predictor_use_encoder_outputs=False, encoder_normalize_before=True, decoder_normalize_before=True,
             teacher_model_decoder_reduction_factor=1)


def _golden():
    z = np.load(GOLDEN)
    return z, {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd.")}


def _step(eng, z):
    ilens, olens = z["ilens"].tolist(), z["olens"].tolist()
    xs = torch.from_numpy(z["xs"])[:, :max(ilens)].contiguous().cuda()
    ys = torch.from_numpy(z["ys"])[:, :max(olens)].contiguous().cuda()
    dpi = torch.from_numpy(z["dp_inputs"])[:, :max(ilens)].contiguous().cuda()
    ds = torch.from_numpy(z["ds"]).cuda()
    after, before = eng.forward(xs, ys, ds, dpi, ilens, olens)
    losses = eng.loss(ys)
    eng.backward()
    torch.cuda.synchronize()
    return after, before, losses


@pytest.mark.parametrize("fp32_gemm", ["simt", "tc"])
def test_golden_tiny_fp32_forward_losses_grads(fp32_gemm):
    """Live-reference dump of one NARVCTrainer training step (trainers/nar_vc.py:52-96) in the fs2_vc yaml's configuration family:
    float32 engine on the CUDA-core GEMM ("simt") and on the fp32-accurate tcgen05 GEMM ("tc", the default float32 mode)."""
    from seq2seq_vc_b200.fsvc_engine import FastSpeechVCEngine

    z, sd = _golden()
    eng = FastSpeechVCEngine(dict(FS_HP, **NO_DROPOUT), device="cuda:0", bf16=False, fp32_gemm=fp32_gemm)
    eng.load_state_dict(sd)
    after, before, losses = _step(eng, z)
    assert np.abs(after.cpu().numpy() - z["after_outs"]).mean() <= 1e-4
    assert np.abs(before.cpu().numpy() - z["before_outs"]).mean() <= 1e-4
    assert np.abs(eng.forward_d_outs().cpu().numpy() - z["d_outs"]).max() <= 1e-4
    assert eng.tlens_host == z["ilens_out"].tolist()
    for i, k in enumerate(("l1_loss", "duration_loss")):
        assert abs(losses[i].item() - float(z[k])) <= 1e-4 * max(1.0, abs(float(z[k]))), k
    for k in [k for k in z.files if k.startswith("attn.")]:
        assert np.abs(eng.attn[k[5:]].cpu().numpy() - z[k]).mean() <= 1e-3, k
    gmax = max(np.abs(z[k]).max() for k in z.files if k.startswith("grad."))
    for name in eng.store.names():
        ref = z["grad." + name]
        got = eng.store.g(name).cpu().numpy()
        if fp32_gemm == "simt":
            assert np.abs(got - ref).max() <= 1e-3 * np.abs(ref).max() + 1e-5 * gmax, name
        else:      # tensor-core accumulation order: a ReLU input within ~1e-6 of zero can flip (see tests/test_gpu_aasvc.py)
            assert np.abs(got - ref).mean() <= 2e-3 * np.abs(ref).mean() + 1e-6 * gmax, name
            assert np.abs(got - ref).max() <= 0.1 * np.abs(ref).max() + 1e-5 * gmax, name
    for k in z.files:
        if k.startswith("bn_after."):
            np.testing.assert_allclose(eng.buffers[k[9:]].cpu().numpy(), z[k], rtol=1e-4, atol=1e-6)
    eng.training = False
    ilens, olens = z["ilens"].tolist(), z["olens"].tolist()
    after_e, _ = eng.forward(torch.from_numpy(z["xs"]).cuda(), torch.from_numpy(z["ys"]).cuda(), torch.from_numpy(z["ds"]).cuda(),
                             torch.from_numpy(z["dp_inputs"]).cuda(), ilens, olens)
    assert np.abs(after_e.cpu().numpy() - z["eval_after_outs"]).mean() <= 1e-4


def test_mid_size_fp32_vs_oracle_and_bf16_drift():
    """A ragged batch at the shipped yaml's widths (adim 384, 2 heads, kernel 15 would take the CPU oracle minutes: adim 128 here)
    with teacher_model_decoder_reduction_factor = 2: float32 engine vs the CPU oracle (forward, losses, every gradient), the bf16
    tensor-core engine within the drift the other model families show."""
    from oracle import fsvc_oracle as fo
    from seq2seq_vc_b200.fsvc_engine import FastSpeechVCEngine

    hp = dict(idim=80, odim=80, adim=128, aheads=2, elayers=2, eunits=256, dlayers=2, dunits=256, duration_predictor_input_dim=80,
              duration_predictor_layers=2, duration_predictor_chans=64, duration_predictor_kernel_size=3, postnet_layers=3, postnet_filts=5,
              postnet_chans=64, conformer_enc_kernel_size=15, conformer_dec_kernel_size=15, teacher_model_decoder_reduction_factor=2,
              **NO_DROPOUT)
    g = torch.Generator().manual_seed(7)
    B, T = 4, 150
    ilens = [150, 133, 101, 64]
    tl = [((i - 2 + 1) // 2 - 2 + 1) // 2 for i in ilens]
    xs = torch.randn(B, T, 80, generator=g)
    ds = torch.randint(0, 4, (B, max(tl)), generator=g)
    for b in range(B):
        xs[b, ilens[b]:] = 0
        ds[b, tl[b]:] = 0
        ds[b, 0] = max(int(ds[b, 0]), 1)
    olens = (2 * ds.sum(1)).tolist()
    ys = torch.randn(B, max(olens), 80, generator=g)
    for b in range(B):
        ys[b, olens[b]:] = 0
    eng = FastSpeechVCEngine(hp, device="cuda:0", bf16=False, seed=5)
    sd = {k: v.detach().cpu().clone() for k, v in eng.state_dict().items()}
    out, parts, grads = fo.fsvc_loss_and_grads(sd, hp, xs, ilens, ys, olens, ds, xs)
    ref = dict(after_outs=out["after_outs"].detach(), before_outs=out["before_outs"].detach(), grads=grads,
               l1_loss=parts["l1_loss"].item(), duration_loss=parts["duration_loss"].item())
    after, before = eng.forward(xs.cuda(), ys.cuda(), ds.cuda(), xs.cuda(), ilens, olens)
    losses = eng.loss(ys.cuda())
    eng.backward()
    torch.cuda.synchronize()
    assert (after.cpu() - ref["after_outs"]).abs().mean().item() <= 1e-4
    assert (before.cpu() - ref["before_outs"]).abs().mean().item() <= 1e-4
    for i, k in enumerate(("l1_loss", "duration_loss")):
        assert abs(losses[i].item() - float(ref[k])) <= 1e-4 * max(1.0, abs(float(ref[k]))), k
    gmax = max(v.abs().max().item() for v in ref["grads"].values())
    for name in eng.store.names():
        r, got = ref["grads"][name], eng.store.g(name).cpu()
        assert r is not None, name
        assert (got - r).abs().mean().item() <= 2e-3 * r.abs().mean().item() + 1e-6 * gmax, name
    G32 = eng.store.G.clone().cpu()
    e16 = FastSpeechVCEngine(hp, device="cuda:0", bf16=True, seed=5)
    e16.load_state_dict(sd)
    a16, _ = e16.forward(xs.cuda(), ys.cuda(), ds.cuda(), xs.cuda(), ilens, olens)
    l16 = e16.loss(ys.cuda())
    e16.backward()
    torch.cuda.synchronize()
    assert torch.isfinite(l16).all()
    assert (a16.float().cpu() - ref["after_outs"]).abs().mean().item() <= 0.05
    assert abs(l16[0].item() - float(ref["l1_loss"])) <= 0.02 * float(ref["l1_loss"])
    assert torch.nn.functional.cosine_similarity(G32, e16.store.G.cpu(), dim=0).item() >= 0.98


def test_dropout_training_is_reproducible_and_trains():
    """Default dropout rates: same-seed steps agree up to float-atomic summation order, another seed differs; 30 fused engine steps
    (forward + loss + backward + clip + Adam) on one batch lower the L1 loss in both precisions."""
    from seq2seq_vc_b200.fsvc_engine import FastSpeechVCEngine

    z, sd = _golden()
    eng = FastSpeechVCEngine(dict(FS_HP), device="cuda:0", bf16=False, seed=3)
    eng.load_state_dict(sd)
    a1, _, l1 = _step(eng, z)
    a1, g1 = a1.clone(), eng.store.G.clone()
    a2, _, _ = _step(eng, z)
    assert (a1 - a2).abs().max().item() <= 1e-4
    assert (g1 - eng.store.G).abs().max().item() <= 1e-3 * g1.abs().max().item()
    assert torch.isfinite(g1).all() and torch.isfinite(l1).all()
    eng.seed_dev += 1
    a3, _, _ = _step(eng, z)
    assert (a1 - a3).abs().mean().item() >= 1e-2
    for bf16 in (False, True):
        e = FastSpeechVCEngine(dict(FS_HP), device="cuda:0", bf16=bf16, seed=1)
        e.load_state_dict(sd)
        e.lr_dev.fill_(2e-3)
        first = last = None
        for _ in range(30):
            _, _, l = _step(e, z)
            e.optimizer_step()
            last = l.clone()
            first = last if first is None else first
        assert torch.isfinite(last).all() and last[0].item() < 0.9 * first[0].item(), (bf16, first, last)


def test_dropin_module_losses_and_autograd():
    """seq2seq_vc_b200.FastSpeechVC + L1Loss + DurationPredictorLoss used exactly as NARVCTrainer._train_step uses the reference
    classes (trainers/nar_vc.py:52-96): same kwargs, 6-tuple, state-dict names; gradients through torch autograd match the
    live-reference dump; the stock torch optimizer moves the engine's parameters."""
    from seq2seq_vc_b200 import DurationPredictorLoss, FastSpeechVC, L1Loss

    z, sd = _golden()
    model = FastSpeechVC(**FS_HP, **NO_DROPOUT, **FIXED, init_type="xavier_uniform", use_masking=True).to("cuda:0")
    assert set(model.state_dict().keys()) == set(sd.keys())
    model.load_state_dict(sd)
    model.train()
    ilens, olens, dlens = torch.from_numpy(z["ilens"]), torch.from_numpy(z["olens"]), torch.from_numpy(z["ilens_out"])
    xs, ys, dpi, ds = (torch.from_numpy(z[k]).cuda() for k in ("xs", "ys", "dp_inputs", "ds"))
    before, after, d_outs, ilens_, olens_, ys_ = model(xs, ilens, ys, olens, ds, dlens, dpi, dp_lengths=ilens)
    assert np.abs(after.detach().cpu().numpy() - z["after_outs"]).mean() <= 1e-4
    assert np.abs(before.detach().cpu().numpy() - z["before_outs"]).mean() <= 1e-4
    assert np.abs(d_outs.detach().cpu().numpy() - z["d_outs"]).max() <= 1e-4
    assert ilens_.tolist() == z["ilens_out"].tolist() and olens_.tolist() == z["olens_out"].tolist()
    l1 = L1Loss()(after, before, ys_, olens_)
    dur = DurationPredictorLoss()(d_outs, ds, ilens_)
    for got, k in ((l1, "l1_loss"), (dur, "duration_loss")):
        assert abs(got.item() - float(z[k])) <= 1e-4 * max(1.0, abs(float(z[k]))), k
    (l1 + dur).backward()
    gmax = max(np.abs(z[k]).max() for k in z.files if k.startswith("grad."))
    for name, p in model.named_parameters():
        ref = z["grad." + name]
        assert p.grad is not None, name
        assert np.abs(p.grad.cpu().numpy() - ref).mean() <= 2e-3 * np.abs(ref).mean() + 1e-6 * gmax, name
        assert np.abs(p.grad.cpu().numpy() - ref).max() <= 0.1 * np.abs(ref).max() + 1e-5 * gmax, name
    att = model.decoder.encoders[1].self_attn.attn
    assert att is not None and tuple(att.shape) == tuple(z["attn.decoder.encoders.1.self_attn"].shape)
    p0 = model.engine.store.P.clone()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
    opt.step()
    assert not torch.equal(p0, model.engine.store.P)


def test_inference_matches_oracle_pipeline():
    """FastSpeechVC.inference (fastspeech_vc.py:427-470): predicted durations (integer path: bit-exact) -> LengthRegulator ->
    decoder, vs the same pipeline assembled from the CPU oracle; speed control alpha changes the output length like the reference's
    round(d * alpha)."""
    from oracle import aasvc_oracle as ao
    from oracle import fsvc_oracle as fo
    from seq2seq_vc_b200 import FastSpeechVC

    z, sd = _golden()
    sd = {**sd, **{k[9:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("bn_after.")}}
    sd["duration_predictor.linear.bias"] = sd["duration_predictor.linear.bias"] + 1.0      # predicted durations not all zero
    model = FastSpeechVC(**FS_HP, **NO_DROPOUT, **FIXED).to("cuda:0")
    model.load_state_dict(sd)
    model.eval()
    il = int(z["ilens"][1])
    x = torch.from_numpy(z["xs"])[1, :il]
    outs, d_outs = model.inference(x.cuda(), dp_input=x.cuda())
    hp = fo.default_hparams(**FS_HP)
    T2 = ((il - 1) // 2 - 1) // 2
    dpi = ao.dp_projection(sd, "duration_predictor_projection", x[None], T2)
    pre = ao.duration_predictor(sd, "duration_predictor", dict(hp), dpi, [T2], clamp=False)
    ds = torch.clamp(torch.round(torch.exp(pre) - 1.0), min=0).long()
    assert torch.equal(ds[0].float(), d_outs.cpu())
    L = int(ds.sum())
    ref = fo.fsvc_forward(sd, FS_HP, x[None], [il], torch.zeros(1, L, 80), [L], ds, x[None], training=False)
    assert tuple(outs.shape) == (L, 80) and (outs.cpu() - ref["after_outs"][0]).abs().mean().item() <= 1e-4
    outs2, _ = model.inference(x.cuda(), dp_input=x.cuda(), alpha=1.5)
    assert outs2.shape[0] == int(torch.round(ds[0].float() * 1.5).long().sum())


def test_fused_train_step_graph_matches_eager_and_prefetch():
    """NARVCTrainStep: CUDA-graph replay == eager launches (same seeds, dropout on) from pinned host batches staged with prefetch();
    losses finite, parameters move."""
    from seq2seq_vc_b200 import FastSpeechVC, NARVCTrainStep

    z, sd = _golden()
    ilens, olens = z["ilens"].tolist(), z["olens"].tolist()
    host = [torch.from_numpy(z[k]).pin_memory() for k in ("xs", "ys", "ds", "dp_inputs")]
    outs = []
    for use_graph in (False, True):
        model = FastSpeechVC(**FS_HP, **FIXED, seed=11).to("cuda:0")
        model.load_state_dict(sd)
        st = NARVCTrainStep(model, lr=1e-3, warmup_steps=10, use_graph=use_graph)
        ls = []
        for i in range(4):
            if i % 2 == 1:
                st.prefetch(*host)
            ls.append(st(host[0], ilens, host[1], olens, host[2], host[3]).clone())
        torch.cuda.synchronize()
        outs.append((torch.stack(ls).cpu(), model.engine.store.P.clone().cpu()))
        assert st.steps == 4 and (not use_graph or st.replayed_launches > 0)
    assert torch.isfinite(outs[0][0]).all() and torch.isfinite(outs[1][0]).all()
    assert (outs[0][0] - outs[1][0]).abs().max().item() <= 5e-3 * outs[0][0].abs().max().item()
    assert (outs[0][1] - outs[1][1]).abs().max().item() <= 1e-4
    assert (model.engine.state_dict()["feat_out.weight"].cpu() - sd["feat_out.weight"]).abs().max().item() > 0


@pytest.mark.parametrize("pw,k", [("conv1d", 1), ("conv1d", 3), ("conv1d-linear", 3)])
def test_positionwise_variants_fp32_vs_oracle(pw, k):
    """MultiLayeredConv1d / Conv1dLinear position-wise layers in both conformer stacks of FastSpeechVC: float32 engine vs the CPU
    oracle (forward, losses, every gradient) on the golden batch with freshly initialised parameters."""
    from oracle import fsvc_oracle as fo
    from seq2seq_vc_b200.fsvc_engine import FastSpeechVCEngine

    z, _ = _golden()
    hp = dict(FS_HP, positionwise_layer_type=pw, positionwise_conv_kernel_size=k, **NO_DROPOUT)
    eng = FastSpeechVCEngine(hp, device="cuda:0", bf16=False, seed=9)
    sd = {n: v.detach().cpu().clone() for n, v in eng.state_dict().items()}
    ilens, olens = z["ilens"].tolist(), z["olens"].tolist()
    xs, ys, dpi, ds = (torch.from_numpy(z[n]) for n in ("xs", "ys", "dp_inputs", "ds"))
    out, parts, grads = fo.fsvc_loss_and_grads(sd, hp, xs, ilens, ys, olens, ds, dpi)
    after, before, losses = _step(eng, z)
    assert (after.cpu() - out["after_outs"].detach()).abs().mean().item() <= 1e-4
    for i, n in enumerate(("l1_loss", "duration_loss")):
        assert abs(losses[i].item() - parts[n].item()) <= 1e-4 * max(1.0, abs(parts[n].item())), n
    gmax = max(g.abs().max().item() for g in grads.values())
    for n in eng.store.names():
        r, got = grads[n], eng.store.g(n).cpu()
        assert (got - r).abs().mean().item() <= 2e-3 * r.abs().mean().item() + 1e-6 * gmax, n
