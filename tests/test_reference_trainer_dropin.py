"""The drop-in claim, end to end: the REFERENCE's own trainers (seq2seq_vc/trainers/ar_vc.py:59-112 ARVCTrainer._train_step,
trainers/aas_vc.py:56-162 AASVCTrainer._train_step) drive this package's VTN / AASVC modules unchanged -- reference batch dict,
reference criterions, torch.optim.Adam, reference WarmupLR, clip_grad_norm_ -- and land on the same parameters and logged
losses as when they drive the reference models.

Build container only (needs /root/reference); CPU: the C-ABI kernels are replaced by their contracts (tests/fake_ops.py),
so this pins the Python boundary (forward signature, returned structures, autograd wiring into the hand-written backward,
nn.Parameter views the optimizer updates in place).  The same modules run the real kernels in tests/test_gpu_*.py."""
import sys
import types

import numpy as np
import pytest
import torch

import fake_ops
from oracle import aasvc_oracle, ref_shim, vtn_oracle

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present")


class _Tqdm:
    def update(self, n):
        pass


@pytest.fixture()
def trainers(monkeypatch, tmp_path):
    ref_shim.install()
    for name in ("tensorboardX", "soundfile", "matplotlib", "matplotlib.pyplot", "h5py"):      # trainer-only imports
        if name not in sys.modules:
            mod = types.ModuleType(name)
            mod.SummaryWriter = lambda *a, **k: types.SimpleNamespace(add_scalar=lambda *a, **k: None)
            mod.use = lambda *a, **k: None
            monkeypatch.setitem(sys.modules, name, mod)
    import seq2seq_vc.trainers.aas_vc as t_aas
    import seq2seq_vc.trainers.ar_vc as t_ar

    fake_ops.install(monkeypatch)
    return t_ar.ARVCTrainer, t_aas.AASVCTrainer, str(tmp_path)


def _run(trainer_cls, model, criterion, config, batch, steps):
    from seq2seq_vc.schedulers.warmup_lr import WarmupLR

    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    sched = WarmupLR(opt, warmup_steps=3)
    tr = trainer_cls(0, 0, {"train": None, "dev": None}, {"train": None, "dev": None}, model, None, criterion, opt, sched, config,
                     torch.device("cpu"))
    tr.tqdm = _Tqdm()
    tr.backward_steps, tr.all_loss = 0, 0.0
    tr.grads = []           # the (clipped) gradients every optimizer step consumed, by parameter name
    names = [n for n, _ in model.named_parameters()]
    opt.register_step_pre_hook(lambda o, a, k: tr.grads.append(
        {n: (p.grad.detach().clone() if p.grad is not None else None) for n, p in zip(names, model.parameters())}))
    for _ in range(steps):
        tr._train_step(batch)
    return tr


def _same_parameter_order(ref, ours):
    """optimizer.state_dict() is keyed by parameter ORDER (trainers/base.py:93,120): resuming a reference checkpoint in the
    drop-in only works if named_parameters() enumerates the same names and shapes in the same order."""
    r = [(n, tuple(p.shape)) for n, p in ref.named_parameters()]
    o = [(n, tuple(p.shape)) for n, p in ours.named_parameters()]
    assert r == o, [(a, b) for a, b in zip(r, o) if a != b][:4]


def _close(sd_ref, sd_our, tol=1.5e-4):
    """Parameters after a few Adam steps.  Adam divides by sqrt(v) ~ |g|, so an element whose gradient is at round-off level
    moves by up to lr per step in a direction round-off decides: single elements are bounded by `tol` (a fraction of lr),
    while the MEAN difference per tensor must stay at 1e-5 -- systematic errors (a wrong Adam clock, moments landing on the
    wrong parameter, a missing 1/k) showed up as mean differences of 1e-4 and more.  Parameters whose gradient is zero in exact
    arithmetic (key biases: softmax shift invariance; the depthwise-conv bias in front of BatchNorm) only ever see round-off
    and are bounded by steps x lr."""
    for k, v in sd_ref.items():
        if v.dtype.is_floating_point:
            d = (v - sd_our[k]).abs()
            if k.endswith("linear_k.bias") or k.endswith("depthwise_conv.bias"):
                assert d.max().item() <= 3.5e-3, k
            else:
                assert d.max().item() <= tol and d.mean().item() <= 1e-5, (k, d.max().item(), d.mean().item())
        else:
            assert torch.equal(v, sd_our[k].to(v.device)), k          # BatchNorm num_batches_tracked counters


def _compare_grads(t_ref, t_our):
    """Adam divides by sqrt(v) ~ |g| in its first steps, so parameters whose gradient is (near) zero move by lr-sized steps
    decided by round-off: the tight comparison is on the gradients each optimizer step consumed; the parameters themselves
    are then only bounded by a fraction of the learning rate."""
    assert len(t_ref.grads) == len(t_our.grads) > 0
    for step, (gr, go) in enumerate(zip(t_ref.grads, t_our.grads)):
        gmax = max(float(g.abs().max()) for g in gr.values() if g is not None)
        for n, g in gr.items():
            if g is None:
                assert go[n] is None or float(go[n].abs().max()) == 0.0, n
                continue
            assert go[n] is not None, n
            rel = 2e-4 if step == 0 else 2e-3          # later steps start from parameters that already differ by round-off
            assert (g - go[n]).abs().max().item() <= rel * float(g.abs().max()) + rel * 1e-2 * gmax, (step, n)


VTN_HP = dict(idim=80, odim=80, dprenet_layers=2, dprenet_units=16, adim=32, aheads=2, elayers=1, eunits=48, dlayers=1, dunits=48,
              postnet_layers=2, postnet_filts=5, postnet_chans=16, decoder_reduction_factor=2, dprenet_dropout_rate=0.0)


@pytest.mark.parametrize("criterions", ["reference", "dropin"])
def test_reference_arvc_trainer_drives_dropin_vtn(trainers, criterions):
    """criterions = "dropin": the criterion factory (bin/vc_train.py:397-403) resolves the loss in this package too."""
    ARVCTrainer, _, outdir = trainers
    import seq2seq_vc_b200
    from seq2seq_vc.losses import Seq2SeqLoss
    from seq2seq_vc.models import VTN as RefVTN

    torch.manual_seed(3)
    ref = RefVTN(**VTN_HP)
    ref_shim.disable_dropout(ref)
    ref.train()
    ours = seq2seq_vc_b200.VTN(**VTN_HP, transformer_enc_dropout_rate=0.0)
    for k in ("enc_positional_dropout_rate", "dec_dropout_rate", "dec_positional_dropout_rate", "postnet_dropout_rate"):
        ours.engine.hp[k] = 0.0       # rates the reference hard-codes too: the harness zeroes them on both sides
    ours.load_state_dict(ref.state_dict())
    ours.train()
    _same_parameter_order(ref, ours)
    # the batch comes out of the reference's own collater (collaters/ar_vc.py:64-73): float32 zero-padded features, int64
    # CPU lengths, stop labels, spembs None
    from seq2seq_vc.collaters.ar_vc import ARVCCollater

    rng = np.random.default_rng(11)
    items = [dict(src_feat=rng.standard_normal((t, 80)).astype(np.float32), trg_feat=rng.standard_normal((l, 80)).astype(np.float32))
             for t, l in ((40, 24), (33, 17))]
    batch = ARVCCollater()(items)
    assert batch["ilens"].dtype == torch.int64 and batch["labels"][1, 16:].eq(1).all()
    config = dict(outdir=outdir, grad_norm=1.0, train_max_steps=10 ** 9, distributed=False, save_interval_steps=10 ** 9,
                  eval_interval_steps=10 ** 9, log_interval_steps=10 ** 9)
    t_ref = _run(ARVCTrainer, ref, {"Seq2SeqLoss": Seq2SeqLoss(bce_pos_weight=10.0)}, config, batch, 3)
    our_crit = Seq2SeqLoss(bce_pos_weight=10.0) if criterions == "reference" else seq2seq_vc_b200.Seq2SeqLoss(bce_pos_weight=10.0)
    t_our = _run(ARVCTrainer, ours, {"Seq2SeqLoss": our_crit}, config, batch, 3)
    assert t_ref.steps == t_our.steps == 3
    for k in ("train/l1_loss", "train/bce_loss", "train/loss"):
        assert abs(t_ref.total_train_loss[k] - t_our.total_train_loss[k]) <= 1e-4 * max(1.0, abs(t_ref.total_train_loss[k])), k
    _compare_grads(t_ref, t_our)
    sd_ref, sd_our = ref.state_dict(), ours.state_dict()
    assert set(sd_ref) == set(sd_our)
    _close(sd_ref, sd_our)
    assert any((sd_our[k] - v0).abs().max().item() > 1e-4 for k, v0 in _fresh_vtn_state(3).items() if v0.dtype.is_floating_point and v0.dim() > 1)


def _fresh_vtn_state(seed):
    from seq2seq_vc.models import VTN as RefVTN

    torch.manual_seed(seed)
    return RefVTN(**VTN_HP).state_dict()


AAS_HP = dict(idim=80, odim=80, adim=32, aheads=2, elayers=1, eunits=48, dlayers=1, dunits=48, duration_predictor_input_dim=80,
              duration_predictor_layers=2, duration_predictor_chans=16, duration_predictor_kernel_size=3, postnet_layers=2,
              postnet_filts=5, postnet_chans=16, post_encoder_reduction_factor=4, conformer_enc_kernel_size=7,
              conformer_dec_kernel_size=7)
AAS_FIXED = dict(positionwise_layer_type="linear", positionwise_conv_kernel_size=1, duration_predictor_use_encoder_outputs=False,
                 encoder_normalize_before=True, decoder_normalize_before=True, duration_predictor_type="deterministic",
                 encoder_input_layer="linear")
AAS_NO_DROPOUT = dict(transformer_enc_dropout_rate=0.0, transformer_enc_positional_dropout_rate=0.0, transformer_enc_attn_dropout_rate=0.0,
                      transformer_dec_dropout_rate=0.0, transformer_dec_positional_dropout_rate=0.0, transformer_dec_attn_dropout_rate=0.0,
                      duration_predictor_dropout_rate=0.0, postnet_dropout_rate=0.0)


@pytest.mark.parametrize("accum,criterions", [(1, "reference"), (2, "reference"), (1, "dropin")])
def test_reference_aasvc_trainer_drives_dropin_aasvc(trainers, accum, criterions):
    """AASVCTrainer._train_step with the reference's L1Loss / ForwardSumLoss / DurationPredictorLoss and lambda_align, incl. its
    gradient_accumulate_steps path (trainers/aas_vc.py:141-149)."""
    _, AASVCTrainer, outdir = trainers
    import seq2seq_vc_b200
    from seq2seq_vc.losses import DurationPredictorLoss, ForwardSumLoss, L1Loss
    from seq2seq_vc.models import AASVC as RefAASVC

    torch.manual_seed(5)
    ref = RefAASVC(**AAS_HP, **AAS_FIXED, **AAS_NO_DROPOUT)
    ref_shim.disable_dropout(ref)
    ref.train()
    ours = seq2seq_vc_b200.AASVC(**AAS_HP, **AAS_FIXED, **AAS_NO_DROPOUT)
    ours.load_state_dict(ref.state_dict())
    ours.train()
    _same_parameter_order(ref, ours)
    from seq2seq_vc.collaters.nar_vc import NARVCCollater            # collaters/nar_vc.py:71-91

    rng = np.random.default_rng(21)
    items = []
    for t, l in ((44, 36), (37, 29)):
        src = rng.standard_normal((t, 80)).astype(np.float32)
        items.append(dict(src_feat=src, trg_feat=rng.standard_normal((l, 80)).astype(np.float32), dp_input=src.copy()))
    batch = NARVCCollater()(items)
    config = dict(outdir=outdir, grad_norm=1.0, train_max_steps=10 ** 9, distributed=False, save_interval_steps=10 ** 9,
                  eval_interval_steps=10 ** 9, log_interval_steps=10 ** 9, lambda_align=2.0, dp_train_start_steps=0,
                  criterions=["L1Loss", "ForwardSumLoss", "DurationPredictorLoss"], gradient_accumulate_steps=accum)
    crit = lambda: {"L1Loss": L1Loss(), "ForwardSumLoss": ForwardSumLoss(), "DurationPredictorLoss": DurationPredictorLoss()}
    t_ref = _run(AASVCTrainer, ref, crit(), config, batch, 2 * accum)
    our_crit = crit() if criterions == "reference" else {"L1Loss": seq2seq_vc_b200.L1Loss(), "ForwardSumLoss": seq2seq_vc_b200.ForwardSumLoss(),
                                                         "DurationPredictorLoss": seq2seq_vc_b200.DurationPredictorLoss()}
    t_our = _run(AASVCTrainer, ours, our_crit, config, batch, 2 * accum)
    assert t_ref.steps == t_our.steps == 2
    for k, v in t_ref.total_train_loss.items():
        assert abs(v - t_our.total_train_loss[k]) <= 2e-4 * max(1.0, abs(v)), k
    _compare_grads(t_ref, t_our)
    sd_ref, sd_our = ref.state_dict(), ours.state_dict()
    assert set(sd_ref) == set(sd_our)
    _close(sd_ref, sd_our)
    fresh = _fresh_aasvc_state()
    assert sum(int((v - fresh[k]).abs().max().item() > 1e-4) for k, v in sd_our.items() if v.dtype.is_floating_point and v.dim() > 1) > 10


_AAS0 = {}


def _fresh_aasvc_state():
    if not _AAS0:
        from seq2seq_vc.models import AASVC as RefAASVC

        torch.manual_seed(5)
        _AAS0.update(RefAASVC(**AAS_HP, **AAS_FIXED, **AAS_NO_DROPOUT).state_dict())
    return _AAS0


TTS_HP = dict(idim=40, odim=80, dprenet_layers=2, dprenet_units=16, adim=32, aheads=2, elayers=2, eunits=48, dlayers=2, dunits=48,
              postnet_layers=2, postnet_filts=5, postnet_chans=16, decoder_reduction_factor=2, dprenet_dropout_rate=0.0,
              use_guided_attn_loss=True, num_heads_applied_guided_attn=2, num_layers_applied_guided_attn=2)


def test_reference_artts_trainer_drives_dropin_transformer_tts(trainers, monkeypatch):
    """ARTTSTrainer._train_step (trainers/ar_tts.py:25-72): token inputs in the collater's 6-tuple, the reference's
    GuidedMultiHeadAttentionLoss applied to the att_ws the drop-in returns (its gradient flows back into our backward)."""
    _, _, outdir = trainers
    import seq2seq_vc.trainers.ar_tts as t_tts
    import seq2seq_vc_b200
    from seq2seq_vc.losses import GuidedMultiHeadAttentionLoss, Seq2SeqLoss
    from seq2seq_vc.models.transformer_tts import TransformerTTS as RefTTS

    torch.manual_seed(9)
    ref = RefTTS(**TTS_HP)
    ref_shim.disable_dropout(ref)
    ref.train()
    ours = seq2seq_vc_b200.TransformerTTS(**TTS_HP)
    for k in ("transformer_enc_dropout_rate", "enc_positional_dropout_rate", "dec_dropout_rate", "dec_positional_dropout_rate",
              "postnet_dropout_rate"):
        ours.engine.hp[k] = 0.0
    ours.load_state_dict(ref.state_dict())
    ours.train()
    _same_parameter_order(ref, ours)
    from seq2seq_vc.collaters.ar_tts import ARTTSCollater             # collaters/ar_tts.py:64: the 6-tuple batch

    rng = np.random.default_rng(33)
    batch = ARTTSCollater()([(rng.integers(1, TTS_HP["idim"] - 1, size=t), rng.standard_normal((l, 80)).astype(np.float32))
                             for t, l in ((11, 26), (7, 18))])
    assert batch[0].dtype == torch.int64 and batch[5] is None
    config = dict(outdir=outdir, grad_norm=1.0, train_max_steps=10 ** 9, distributed=False, save_interval_steps=10 ** 9,
                  eval_interval_steps=10 ** 9, log_interval_steps=10 ** 9, use_guided_attn_loss=True)
    crit = lambda: {"Seq2SeqLoss": Seq2SeqLoss(), "guided_attn": GuidedMultiHeadAttentionLoss(sigma=0.4, alpha=1.0)}
    t_ref = _run(t_tts.ARTTSTrainer, ref, crit(), config, batch, 3)
    t_our = _run(t_tts.ARTTSTrainer, ours, crit(), config, batch, 3)
    assert t_ref.steps == t_our.steps == 3
    for k, v in t_ref.total_train_loss.items():
        assert abs(v - t_our.total_train_loss[k]) <= 1e-4 * max(1.0, abs(v)), k
    assert t_ref.total_train_loss["train/guided_attn_loss"] > 0
    _compare_grads(t_ref, t_our)
    sd_ref, sd_our = ref.state_dict(), ours.state_dict()
    assert set(sd_ref) == set(sd_our)
    _close(sd_ref, sd_our)


def test_reference_trainer_checkpoint_and_freeze_with_dropin(trainers, tmp_path):
    """trainers/base.py:85-121 save_checkpoint / load_checkpoint and :226 freeze_modules (yaml `freeze-mods`, matched on
    state-dict key prefixes, utils/model_io.py:95-111) on the drop-in VTN: a reference checkpoint loads into it and vice
    versa, frozen prefixes stay put (no .grad, untouched by Adam and by the clip norm) while the rest trains as in the reference."""
    ARVCTrainer, _, outdir = trainers
    import seq2seq_vc_b200
    from seq2seq_vc.losses import Seq2SeqLoss
    from seq2seq_vc.models import VTN as RefVTN

    torch.manual_seed(13)
    ref = RefVTN(**VTN_HP)
    ref_shim.disable_dropout(ref)
    ref.train()
    ours = seq2seq_vc_b200.VTN(**VTN_HP, transformer_enc_dropout_rate=0.0)
    for k in ("enc_positional_dropout_rate", "dec_dropout_rate", "dec_positional_dropout_rate", "postnet_dropout_rate"):
        ours.engine.hp[k] = 0.0
    ours.train()
    xs, ilens, ys, labels, olens = vtn_oracle.synthetic_batch(2, 40, 24, ilens=[40, 33], olens=[24, 17], seed=12)
    batch = dict(xs=xs, ys=ys, labels=labels, ilens=torch.tensor(ilens), olens=torch.tensor(olens))
    config = dict(outdir=outdir, grad_norm=1.0, train_max_steps=10 ** 9, distributed=False, save_interval_steps=10 ** 9,
                  eval_interval_steps=10 ** 9, log_interval_steps=10 ** 9)
    # a checkpoint written by the reference trainer around the reference model loads into the drop-in through the trainer
    t_ref = _run(ARVCTrainer, ref, {"Seq2SeqLoss": Seq2SeqLoss()}, config, batch, 1)
    ck = str(tmp_path / "ck" / "checkpoint-1steps.pkl")
    t_ref.save_checkpoint(ck)
    t_our = _run(ARVCTrainer, ours, {"Seq2SeqLoss": Seq2SeqLoss()}, config, batch, 0)
    t_our.load_checkpoint(ck)             # parameters, Adam moments (keyed by parameter order), scheduler, step counter
    assert t_our.steps == 1
    for k, v in ref.state_dict().items():
        assert torch.equal(v, ours.state_dict()[k]), k
    # ... and back: the drop-in's checkpoint is a plain reference checkpoint
    ck2 = str(tmp_path / "ck" / "ours.pkl")
    t_our.save_checkpoint(ck2)
    blob = torch.load(ck2, map_location="cpu")
    assert set(blob) == {"optimizer", "scheduler", "steps", "epochs", "model"}
    RefVTN(**VTN_HP).load_state_dict(blob["model"])
    # init-mods (yaml `init-mods`, trainers/ar_vc.py:30-57 load_trained_modules -> utils/model_io.py:12-57): partial transfer
    # by state-dict key prefix into a differently initialised drop-in
    torch.manual_seed(99)
    other = seq2seq_vc_b200.VTN(**VTN_HP, transformer_enc_dropout_rate=0.0)
    other.load_state_dict(RefVTN(**VTN_HP).state_dict())
    before = {k: v.clone() for k, v in other.state_dict().items()}
    t_other = _run(ARVCTrainer, other, {"Seq2SeqLoss": Seq2SeqLoss()}, config, batch, 0)
    t_other.load_trained_modules(ck, ["encoder", "postnet"])
    after = other.state_dict()
    for k, v in ref.state_dict().items():
        if not v.dtype.is_floating_point:
            continue
        if k.startswith(("encoder.", "postnet.")):
            assert torch.equal(after[k], blob["model"][k]), k
        else:
            assert torch.equal(after[k], before[k]), k
    assert not torch.equal(after["encoder.embed.conv.0.weight"], before["encoder.embed.conv.0.weight"])
    # freeze-mods: encoder prefix frozen on both sides, two more steps
    frozen0 = {k: v.clone() for k, v in ours.state_dict().items() if k.startswith("encoder.")}
    for t in (t_ref, t_our):
        t.freeze_modules(["encoder"])
        for _ in range(2):
            t._train_step(batch)
    assert all(p.grad is None for n, p in ours.named_parameters() if n.startswith("encoder."))
    for k, v in frozen0.items():
        if v.dtype.is_floating_point:
            assert torch.equal(v, ours.state_dict()[k]), k
    _close(ref.state_dict(), ours.state_dict())
    assert any((v - ours.state_dict()[k]).abs().max().item() > 0 for k, v in blob["model"].items() if k.startswith("decoder.") and v.dim() > 1)


# ------------------------------------------------------------------ --distributed (bin/vc_train.py:423-431), world size 2, gloo
def _ddp_worker(rank, world, port, out_dir, family="vtn"):
    import os

    here = os.path.dirname(os.path.abspath(__file__))
    for p in (here, os.path.dirname(here)):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist

    import fake_ops as fo
    import seq2seq_vc_b200
    import seq2seq_vc_b200.api as api
    import seq2seq_vc_b200.ops as ops
    from oracle import ref_shim as rs, vtn_oracle as vo

    rs.install()
    for name in ("tensorboardX", "soundfile", "matplotlib", "matplotlib.pyplot", "h5py"):
        if name not in sys.modules:
            mod = types.ModuleType(name)
            mod.SummaryWriter = lambda *a, **k: types.SimpleNamespace(add_scalar=lambda *a, **k: None)
            mod.use = lambda *a, **k: None
            sys.modules[name] = mod
    for n in fo.ALL:
        if hasattr(ops, n) and n != "logmel":
            setattr(ops, n, getattr(fo, n))
    api._require_cuda = lambda t, who: None
    from seq2seq_vc.losses import Seq2SeqLoss
    from seq2seq_vc.models import VTN as RefVTN
    from seq2seq_vc.trainers.ar_vc import ARVCTrainer

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    if family == "aasvc":
        # AASVCTrainer with gradient_accumulate_steps = 2: every micro-step's backward averages over the ranks
        from oracle import aasvc_oracle as ao
        from seq2seq_vc.losses import DurationPredictorLoss, ForwardSumLoss, L1Loss
        from seq2seq_vc.models import AASVC as RefAASVC
        from seq2seq_vc.trainers.aas_vc import AASVCTrainer

        torch.manual_seed(3)
        ref = RefAASVC(**AAS_HP, **AAS_FIXED, **AAS_NO_DROPOUT)
        rs.disable_dropout(ref)
        ref.train()
        torch.manual_seed(100 + rank)
        ours = seq2seq_vc_b200.AASVC(**AAS_HP, **AAS_FIXED, **AAS_NO_DROPOUT)
        if rank == 0:
            ours.load_state_dict(ref.state_dict())
        ours.train()
        xs, ilens, ys, olens, dpi = ao.synthetic_batch(2, 44, 36, ilens=[44, 37 - rank], olens=[36, 29 + rank], seed=70 + rank)
        batch = dict(xs=xs, ys=ys, ilens=torch.tensor(ilens), olens=torch.tensor(olens), dp_inputs=dpi, dplens=torch.tensor(ilens))
        config = dict(outdir=out_dir, grad_norm=1.0, train_max_steps=10 ** 9, distributed=True, save_interval_steps=10 ** 9,
                      eval_interval_steps=10 ** 9, log_interval_steps=10 ** 9, lambda_align=2.0, dp_train_start_steps=0,
                      criterions=["L1Loss", "ForwardSumLoss", "DurationPredictorLoss"], gradient_accumulate_steps=2)
        crit = lambda: {"L1Loss": L1Loss(), "ForwardSumLoss": ForwardSumLoss(), "DurationPredictorLoss": DurationPredictorLoss()}
        # find_unused_parameters: the duration predictor takes no part in the first window (no duration loss at step 0)
        t_ref = _run(AASVCTrainer, torch.nn.parallel.DistributedDataParallel(ref, find_unused_parameters=True), crit(), config, batch, 4)
        t_our = _run(AASVCTrainer, seq2seq_vc_b200.DistributedDataParallel(ours), crit(), config, batch, 4)
        assert t_ref.steps == t_our.steps == 2
        t_our.save_checkpoint(os.path.join(out_dir, f"ours{rank}.pkl"))
        torch.save({k: v.clone() for k, v in ref.state_dict().items()}, os.path.join(out_dir, f"ref{rank}.pt"))
        dist.destroy_process_group()
        return
    torch.manual_seed(3)
    ref = RefVTN(**VTN_HP)
    rs.disable_dropout(ref)
    ref.train()
    torch.manual_seed(100 + rank)                   # rank-dependent init: the wrapper must broadcast rank 0's parameters
    ours = seq2seq_vc_b200.VTN(**VTN_HP, transformer_enc_dropout_rate=0.0)
    for k in ("enc_positional_dropout_rate", "dec_dropout_rate", "dec_positional_dropout_rate", "postnet_dropout_rate"):
        ours.engine.hp[k] = 0.0
    if rank == 0:
        ours.load_state_dict(ref.state_dict())
    ours.train()
    xs, ilens, ys, labels, olens = vo.synthetic_batch(2, 40, 24, ilens=[40, 33 - rank], olens=[24, 17 + rank], seed=50 + rank)
    batch = dict(xs=xs, ys=ys, labels=labels, ilens=torch.tensor(ilens), olens=torch.tensor(olens))
    config = dict(outdir=out_dir, grad_norm=1.0, train_max_steps=10 ** 9, distributed=True, save_interval_steps=10 ** 9,
                  eval_interval_steps=10 ** 9, log_interval_steps=10 ** 9)
    t_ref = _run(ARVCTrainer, torch.nn.parallel.DistributedDataParallel(ref), {"Seq2SeqLoss": Seq2SeqLoss()}, config, batch, 2)
    t_our = _run(ARVCTrainer, seq2seq_vc_b200.DistributedDataParallel(ours), {"Seq2SeqLoss": Seq2SeqLoss()}, config, batch, 2)
    t_our.save_checkpoint(os.path.join(out_dir, f"ours{rank}.pkl"))          # goes through self.model.module
    torch.save({k: v.clone() for k, v in ref.state_dict().items()}, os.path.join(out_dir, f"ref{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("family", ["vtn", "aasvc"])
def test_reference_trainer_distributed_with_dropin_wrapper(trainers, tmp_path, family):
    """Two ranks, rank-specific batches, ARVCTrainer / AASVCTrainer (gradient_accumulate_steps = 2) with config["distributed"]:
    the reference model under torch DDP (apex is not installed here) vs the drop-in under seq2seq_vc_b200.DistributedDataParallel
    land on the same parameters; the replicas stay identical; the wrapper broadcast rank 0's initial parameters."""
    import socket

    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_ddp_worker, args=(2, port, str(tmp_path), family), nprocs=2, join=True)
    ours0 = torch.load(tmp_path / "ours0.pkl", map_location="cpu")["model"]
    ours1 = torch.load(tmp_path / "ours1.pkl", map_location="cpu")["model"]
    ref0 = torch.load(tmp_path / "ref0.pt")
    for k, v in ref0.items():
        if v.dtype.is_floating_point and "running_" not in k:
            assert torch.equal(ours0[k], ours1[k]), k                      # replicas in lock-step
    _close({k: v for k, v in ref0.items() if "running_" not in k}, ours0)


# ------------------------------------------------------------------ fused steps vs the reference trainers
def test_fused_vtn_train_step_matches_reference_trainer(trainers):
    """seq2seq_vc_b200.VTNTrainStep (own clip + Adam + WarmupLR on flat buffers) == ARVCTrainer._train_step around the reference
    model with torch.optim.Adam / clip_grad_norm_ / WarmupLR: same losses, same parameters after 3 steps."""
    ARVCTrainer, _, outdir = trainers
    import seq2seq_vc_b200
    from seq2seq_vc.losses import Seq2SeqLoss
    from seq2seq_vc.models import VTN as RefVTN

    torch.manual_seed(17)
    ref = RefVTN(**VTN_HP)
    ref_shim.disable_dropout(ref)
    ref.train()
    ours = seq2seq_vc_b200.VTN(**VTN_HP, transformer_enc_dropout_rate=0.0)
    for k in ("enc_positional_dropout_rate", "dec_dropout_rate", "dec_positional_dropout_rate", "postnet_dropout_rate"):
        ours.engine.hp[k] = 0.0
    ours.load_state_dict(ref.state_dict())
    xs, ilens, ys, labels, olens = vtn_oracle.synthetic_batch(2, 40, 24, ilens=[40, 33], olens=[24, 17], seed=19)
    batch = dict(xs=xs, ys=ys, labels=labels, ilens=torch.tensor(ilens), olens=torch.tensor(olens))
    config = dict(outdir=outdir, grad_norm=1.0, train_max_steps=10 ** 9, distributed=False, save_interval_steps=10 ** 9,
                  eval_interval_steps=10 ** 9, log_interval_steps=10 ** 9)
    t_ref = _run(ARVCTrainer, ref, {"Seq2SeqLoss": Seq2SeqLoss()}, config, batch, 3)       # Adam lr 1e-3, WarmupLR(3), clip 1.0
    step = seq2seq_vc_b200.VTNTrainStep(ours, lr=1e-3, warmup_steps=3, grad_norm=1.0, bce_pos_weight=10.0)
    tot = torch.zeros(2)
    for _ in range(3):
        tot += step(xs, ilens, ys, labels, olens).float().cpu()
    assert abs(float(tot[0]) - t_ref.total_train_loss["train/l1_loss"]) <= 1e-4 * max(1.0, t_ref.total_train_loss["train/l1_loss"])
    assert abs(float(tot[1]) - t_ref.total_train_loss["train/bce_loss"]) <= 1e-4 * max(1.0, t_ref.total_train_loss["train/bce_loss"])
    sd_our = ours.engine.state_dict()
    _close(ref.state_dict(), sd_our)


@pytest.mark.parametrize("accum", [1, 2])
def test_fused_aasvc_train_step_matches_reference_trainer(trainers, accum):
    """seq2seq_vc_b200.AASVCTrainStep == AASVCTrainer._train_step around the reference model: lambda_align = 2, no duration loss
    before dp_train_start_steps, gradient accumulation, clip, Adam, WarmupLR."""
    _, AASVCTrainer, outdir = trainers
    import seq2seq_vc_b200
    from seq2seq_vc.losses import DurationPredictorLoss, ForwardSumLoss, L1Loss
    from seq2seq_vc.models import AASVC as RefAASVC

    torch.manual_seed(23)
    ref = RefAASVC(**AAS_HP, **AAS_FIXED, **AAS_NO_DROPOUT)
    ref_shim.disable_dropout(ref)
    ref.train()
    ours = seq2seq_vc_b200.AASVC(**AAS_HP, **AAS_FIXED, **AAS_NO_DROPOUT)
    ours.load_state_dict(ref.state_dict())
    xs, ilens, ys, olens, dpi = aasvc_oracle.synthetic_batch(2, 44, 36, ilens=[44, 37], olens=[36, 29], seed=29)
    batch = dict(xs=xs, ys=ys, ilens=torch.tensor(ilens), olens=torch.tensor(olens), dp_inputs=dpi, dplens=torch.tensor(ilens))
    config = dict(outdir=outdir, grad_norm=1.0, train_max_steps=10 ** 9, distributed=False, save_interval_steps=10 ** 9,
                  eval_interval_steps=10 ** 9, log_interval_steps=10 ** 9, lambda_align=2.0, dp_train_start_steps=0,
                  criterions=["L1Loss", "ForwardSumLoss", "DurationPredictorLoss"], gradient_accumulate_steps=accum)
    crit = {"L1Loss": L1Loss(), "ForwardSumLoss": ForwardSumLoss(), "DurationPredictorLoss": DurationPredictorLoss()}
    t_ref = _run(AASVCTrainer, ref, crit, config, batch, 3 * accum)
    step = seq2seq_vc_b200.AASVCTrainStep(ours, lr=1e-3, warmup_steps=3, grad_norm=1.0, dp_train_start_steps=0,
                                          gradient_accumulate_steps=accum)
    tot = torch.zeros(4)
    for _ in range(3 * accum):
        tot += step(xs, ilens, ys, olens, dpi).float().cpu() / accum
    assert step.steps == t_ref.steps == 3
    for i, k in enumerate(("train/l1_loss", "train/forward_sum_loss", "train/binary_loss", "train/duration_loss")):
        assert abs(float(tot[i]) - t_ref.total_train_loss[k]) <= 2e-4 * max(1.0, abs(t_ref.total_train_loss[k])), k
    sd_our = ours.engine.state_dict()
    _close(ref.state_dict(), sd_our)


def test_fused_tts_train_step_with_guided_attention_matches_reference_trainer(trainers):
    """VTNTrainStep(guided_attn=...) around TransformerTTS == ARTTSTrainer._train_step with use_guided_attn_loss
    (trainers/ar_tts.py:49-53, losses/guided_attention_loss.py:109-165)."""
    _, _, outdir = trainers
    import seq2seq_vc.trainers.ar_tts as t_tts
    import seq2seq_vc_b200
    from seq2seq_vc.losses import GuidedMultiHeadAttentionLoss, Seq2SeqLoss
    from seq2seq_vc.models.transformer_tts import TransformerTTS as RefTTS

    torch.manual_seed(31)
    ref = RefTTS(**TTS_HP)
    ref_shim.disable_dropout(ref)
    ref.train()
    ours = seq2seq_vc_b200.TransformerTTS(**TTS_HP)
    for k in ("transformer_enc_dropout_rate", "enc_positional_dropout_rate", "dec_dropout_rate", "dec_positional_dropout_rate",
              "postnet_dropout_rate"):
        ours.engine.hp[k] = 0.0
    ours.load_state_dict(ref.state_dict())
    g = torch.Generator().manual_seed(37)
    ilens, olens = [11, 7], [26, 18]
    tokens = torch.randint(1, TTS_HP["idim"] - 1, (2, 11), generator=g)
    ys = torch.randn(2, 26, 80, generator=g)
    labels = torch.zeros(2, 26)
    for b in range(2):
        tokens[b, ilens[b]:] = 0
        ys[b, olens[b]:] = 0
        labels[b, olens[b] - 1:] = 1.0
    config = dict(outdir=outdir, grad_norm=1.0, train_max_steps=10 ** 9, distributed=False, save_interval_steps=10 ** 9,
                  eval_interval_steps=10 ** 9, log_interval_steps=10 ** 9, use_guided_attn_loss=True)
    crit = {"Seq2SeqLoss": Seq2SeqLoss(), "guided_attn": GuidedMultiHeadAttentionLoss(sigma=0.4, alpha=1.0)}
    t_ref = _run(t_tts.ARTTSTrainer, ref, crit, config, (tokens, torch.tensor(ilens), ys, labels, torch.tensor(olens), None), 3)
    step = seq2seq_vc_b200.VTNTrainStep(ours, lr=1e-3, warmup_steps=3, grad_norm=1.0,
                                        guided_attn=dict(sigma=0.4, alpha=1.0, n_layers=2, n_heads=2))
    tot, ga = torch.zeros(2), 0.0
    for _ in range(3):
        tot += step(tokens, ilens, ys, labels, olens).float().cpu()
        ga += float(step.ga_loss)
    assert abs(float(tot[0]) - t_ref.total_train_loss["train/l1_loss"]) <= 1e-4 * max(1.0, t_ref.total_train_loss["train/l1_loss"])
    assert abs(float(tot[1]) - t_ref.total_train_loss["train/bce_loss"]) <= 1e-4 * max(1.0, t_ref.total_train_loss["train/bce_loss"])
    assert abs(ga - t_ref.total_train_loss["train/guided_attn_loss"]) <= 1e-4 * max(1.0, t_ref.total_train_loss["train/guided_attn_loss"])
    sd_our = ours.engine.state_dict()
    _close(ref.state_dict(), sd_our)


# ------------------------------------------------------------------ checkpoints move between the reference trainer and the fused steps
@pytest.mark.parametrize("family", ["vtn", "aasvc"])
def test_checkpoints_move_between_reference_trainer_and_fused_step(trainers, tmp_path, family):
    """trainers/base.py:85-121: a checkpoint the reference trainer wrote resumes in the fused step (parameters, Adam moments and
    per-parameter step counts, scheduler position) and continues exactly like the reference would; and the fused step's
    state_dict() is a checkpoint the reference trainer resumes from.  For AAS-VC the duration predictor's Adam clock is one
    step behind everybody else's at that point (no duration loss at step 0)."""
    ARVCTrainer, AASVCTrainer, outdir = trainers
    import seq2seq_vc_b200
    from seq2seq_vc.losses import DurationPredictorLoss, ForwardSumLoss, L1Loss, Seq2SeqLoss
    from seq2seq_vc.models import AASVC as RefAASVC, VTN as RefVTN

    base = dict(outdir=outdir, grad_norm=1.0, train_max_steps=10 ** 9, distributed=False, save_interval_steps=10 ** 9,
                eval_interval_steps=10 ** 9, log_interval_steps=10 ** 9)
    if family == "vtn":
        def make_ref():
            torch.manual_seed(41)
            m = RefVTN(**VTN_HP)
            ref_shim.disable_dropout(m)
            return m.train()

        def make_ours():
            m = seq2seq_vc_b200.VTN(**VTN_HP, transformer_enc_dropout_rate=0.0)
            for k in ("enc_positional_dropout_rate", "dec_dropout_rate", "dec_positional_dropout_rate", "postnet_dropout_rate"):
                m.engine.hp[k] = 0.0
            return m

        xs, ilens, ys, labels, olens = vtn_oracle.synthetic_batch(2, 40, 24, ilens=[40, 33], olens=[24, 17], seed=43)
        batch = dict(xs=xs, ys=ys, labels=labels, ilens=torch.tensor(ilens), olens=torch.tensor(olens))
        trainer_cls, config = ARVCTrainer, base
        crit = lambda: {"Seq2SeqLoss": Seq2SeqLoss()}
        fused = lambda m: seq2seq_vc_b200.VTNTrainStep(m, lr=1e-3, warmup_steps=3, grad_norm=1.0)
        call = lambda st: st(xs, ilens, ys, labels, olens)
    else:
        def make_ref():
            torch.manual_seed(41)
            m = RefAASVC(**AAS_HP, **AAS_FIXED, **AAS_NO_DROPOUT)
            ref_shim.disable_dropout(m)
            return m.train()

        make_ours = lambda: seq2seq_vc_b200.AASVC(**AAS_HP, **AAS_FIXED, **AAS_NO_DROPOUT)
        xs, ilens, ys, olens, dpi = aasvc_oracle.synthetic_batch(2, 44, 36, ilens=[44, 37], olens=[36, 29], seed=43)
        batch = dict(xs=xs, ys=ys, ilens=torch.tensor(ilens), olens=torch.tensor(olens), dp_inputs=dpi, dplens=torch.tensor(ilens))
        trainer_cls = AASVCTrainer
        config = dict(base, lambda_align=2.0, dp_train_start_steps=0, criterions=["L1Loss", "ForwardSumLoss", "DurationPredictorLoss"])
        crit = lambda: {"L1Loss": L1Loss(), "ForwardSumLoss": ForwardSumLoss(), "DurationPredictorLoss": DurationPredictorLoss()}
        fused = lambda m: seq2seq_vc_b200.AASVCTrainStep(m, lr=1e-3, warmup_steps=3, grad_norm=1.0)
        call = lambda st: st(xs, ilens, ys, olens, dpi)

    # reference trainer: 2 steps -> checkpoint -> (a) the reference continues, (b) the fused step resumes from the file
    ref = make_ref()
    t_ref = _run(trainer_cls, ref, crit(), config, batch, 2)
    ck = str(tmp_path / "ck" / "checkpoint-2steps.pkl")
    t_ref.save_checkpoint(ck)
    t_ref._train_step(batch)
    ours = make_ours()
    st = fused(ours)
    st.load_state_dict(torch.load(ck, map_location="cpu"))
    assert st.steps == 2
    call(st)
    _close(ref.state_dict(), ours.engine.state_dict())

    # fused step: 2 steps -> state_dict() -> (a) the fused step continues, (b) the reference trainer resumes from the file
    ours2 = make_ours()
    ours2.load_state_dict(make_ref().state_dict())
    st2 = fused(ours2)
    call(st2)
    call(st2)
    ck2 = str(tmp_path / "ck" / "fused-2steps.pkl")
    torch.save(st2.state_dict(), ck2)
    call(st2)
    ref2 = make_ref()
    with torch.no_grad():
        for p_ in ref2.parameters():
            p_.add_(1.0)                                  # whatever it holds is replaced by the checkpoint
    t_ref2 = _run(trainer_cls, ref2, crit(), config, batch, 0)
    t_ref2.load_checkpoint(ck2)
    assert t_ref2.steps == 2
    t_ref2._train_step(batch)
    _close(ref2.state_dict(), ours2.engine.state_dict())
    # and both continuations agree with each other (same initial state, same batch)
    _close(ref.state_dict(), ours2.engine.state_dict())


# ------------------------------------------------------------------ the trainer's evaluation hook (trainers/base.py:165-184)
class _RecordingPlt:
    """matplotlib.pyplot stand-in that records every array handed to plot / imshow."""

    def __init__(self):
        self.arrays = []

    def plot(self, a, *args, **kw):
        self.arrays.append(np.array(a))

    def imshow(self, a, *args, **kw):
        self.arrays.append(np.array(a))

    def __getattr__(self, name):
        return lambda *a, **k: None


def test_reference_trainer_eval_hook_with_dropin_inference(trainers, monkeypatch, tmp_path):
    """ARVCTrainer._genearete_and_save_intermediate_result (trainers/ar_vc.py:114-225, called from _eval_epoch every
    eval_interval_steps): feeds padded rows of the dev batch to model.inference(x, config["inference"], spemb=None) and plots
    outs / probs / att_ws.  Everything it draws with the drop-in equals what it draws with the reference model."""
    ARVCTrainer, _, outdir = trainers
    import seq2seq_vc.trainers.ar_vc as t_ar
    import seq2seq_vc_b200
    from seq2seq_vc.losses import Seq2SeqLoss
    from seq2seq_vc.models import VTN as RefVTN

    torch.manual_seed(51)
    ref = RefVTN(**VTN_HP)
    ours = seq2seq_vc_b200.VTN(**VTN_HP)
    ours.load_state_dict(ref.state_dict())
    ref.eval()
    ours.eval()
    xs, ilens, ys, labels, olens = vtn_oracle.synthetic_batch(2, 40, 24, ilens=[40, 33], olens=[24, 17], seed=53)
    batch = dict(xs=xs, ys=ys, labels=labels, ilens=torch.tensor(ilens), olens=torch.tensor(olens))
    drawn = []
    for i, model in enumerate((ref, ours)):
        config = dict(outdir=str(tmp_path / f"run{i}"), grad_norm=1.0, train_max_steps=10 ** 9, distributed=False,
                      inference=dict(threshold=0.5, minlenratio=0.0, maxlenratio=0.6), num_save_intermediate_results=4)
        plt = _RecordingPlt()
        monkeypatch.setattr(t_ar, "plt", plt)
        tr = _run(ARVCTrainer, model, {"Seq2SeqLoss": Seq2SeqLoss()}, config, batch, 0)
        tr._genearete_and_save_intermediate_result(batch)
        drawn.append(plt.arrays)
    assert len(drawn[0]) == len(drawn[1]) > 0
    for a, b in zip(*drawn):
        assert a.shape == b.shape
        assert np.abs(a - b).max() <= 2e-4, np.abs(a - b).max()
    assert any(a.ndim == 2 and a.shape[0] > 1 for a in drawn[1])


def test_reference_aasvc_trainer_eval_hook_with_dropin_inference(trainers, monkeypatch, tmp_path):
    """AASVCTrainer._genearete_and_save_intermediate_result (trainers/aas_vc.py:205-290) calls
    model.inference(x[:ilen], y[:olen], spembs=None, dp_input=dp_input) -- WITH the ground-truth target -- and unpacks five
    values (outs, d_outs, ds, log_p_attn, ilens_): the drop-in serves that form; durations from the alignment search are
    exact, everything drawn (outs, target, log_p_attn) matches the reference model's."""
    _, AASVCTrainer, outdir = trainers
    import seq2seq_vc.trainers.aas_vc as t_aas
    import seq2seq_vc_b200
    from seq2seq_vc.losses import L1Loss
    from seq2seq_vc.models import AASVC as RefAASVC

    torch.manual_seed(61)
    ref = RefAASVC(**AAS_HP, **AAS_FIXED, **AAS_NO_DROPOUT)
    with torch.no_grad():
        ref.duration_predictor.linear.bias.add_(1.0)          # so that the predicted durations are not all zero
    ours = seq2seq_vc_b200.AASVC(**AAS_HP, **AAS_FIXED, **AAS_NO_DROPOUT)
    ours.load_state_dict(ref.state_dict())
    ref.eval()
    ours.eval()
    xs, ilens, ys, olens, dpi = aasvc_oracle.synthetic_batch(2, 44, 36, ilens=[44, 37], olens=[36, 29], seed=63)
    batch = dict(xs=xs, ys=ys, ilens=torch.tensor(ilens), olens=torch.tensor(olens), dp_inputs=dpi, dplens=torch.tensor(ilens))
    # the 5-tuple itself
    r = ref.inference(xs[0, :44], ys[0, :36], spembs=None, dp_input=dpi[0])
    o = ours.inference(xs[0, :44], ys[0, :36], spembs=None, dp_input=dpi[0])
    assert len(r) == len(o) == 5
    assert (r[0] - o[0]).abs().max().item() <= 2e-4 and torch.equal(r[1].long(), o[1].long())
    assert torch.equal(r[2], o[2]) and int(r[4]) == int(o[4])                     # MAS durations: exact
    fin = torch.isfinite(r[3])
    assert torch.equal(fin, torch.isfinite(o[3])) and (r[3][fin] - o[3][fin]).abs().max().item() <= 2e-4
    drawn = []
    for i, model in enumerate((ref, ours)):
        config = dict(outdir=str(tmp_path / f"run{i}"), grad_norm=1.0, train_max_steps=10 ** 9, distributed=False,
                      num_save_intermediate_results=4, criterions=["L1Loss"])
        plt = _RecordingPlt()
        monkeypatch.setattr(t_aas, "plt", plt)
        tr = _run(AASVCTrainer, model, {"L1Loss": L1Loss()}, config, batch, 0)
        tr._genearete_and_save_intermediate_result(batch)
        drawn.append(plt.arrays)
    assert len(drawn[0]) == len(drawn[1]) > 0
    for a, b in zip(*drawn):
        assert a.shape == b.shape
        m = np.isfinite(a)
        assert np.array_equal(m, np.isfinite(b)) and np.abs(a[m] - b[m]).max() <= 2e-4


# ------------------------------------------------------------------ FastSpeechVC (trainers/nar_vc.py)
FS_HP = dict(idim=80, odim=80, adim=32, aheads=2, elayers=2, eunits=48, dlayers=2, dunits=48, duration_predictor_input_dim=80,
             duration_predictor_layers=2, duration_predictor_chans=16, duration_predictor_kernel_size=3, postnet_layers=2, postnet_filts=5,
             postnet_chans=16, conformer_enc_kernel_size=7, conformer_dec_kernel_size=7)
FS_FIXED = dict(encoder_type="conformer", decoder_type="conformer", encoder_input_layer="conv2d", positionwise_layer_type="linear",
                duration_predictor_use_encoder_outputs=False, encoder_normalize_before=True, decoder_normalize_before=True,
                teacher_model_decoder_reduction_factor=1)


def _fs_batch():
    """Out of the reference's own collater (collaters/nar_vc.py:49-91, the branch with teacher durations)."""
    from seq2seq_vc.collaters.nar_vc import NARVCCollater

    rng = np.random.default_rng(31)
    items = []
    for t in (62, 51):
        tl = ((t - 1) // 2 - 1) // 2
        d = rng.integers(0, 4, tl).astype(np.int64)
        d[0] = max(d[0], 1)
        src = rng.standard_normal((t, 80)).astype(np.float32)
        items.append(dict(src_feat=src, trg_feat=rng.standard_normal((int(d.sum()), 80)).astype(np.float32), dp_input=src.copy(), duration=d))
    return NARVCCollater()(items)


def _nar_trainer(monkeypatch):
    import seq2seq_vc.trainers.nar_vc as t_nar

    return t_nar.NARVCTrainer


FS_CONFIG = dict(grad_norm=1.0, train_max_steps=10 ** 9, distributed=False, save_interval_steps=10 ** 9, eval_interval_steps=10 ** 9,
                 log_interval_steps=10 ** 9)


@pytest.mark.parametrize("criterions", ["reference", "dropin"])
def test_reference_narvc_trainer_drives_dropin_fastspeech_vc(trainers, monkeypatch, criterions):
    """NARVCTrainer._train_step (trainers/nar_vc.py:52-103) with the reference's batch dict, L1Loss / DurationPredictorLoss,
    torch.optim.Adam, WarmupLR and clip_grad_norm_ around seq2seq_vc_b200.FastSpeechVC vs around the reference model."""
    _, _, outdir = trainers
    import seq2seq_vc_b200
    from seq2seq_vc.losses import DurationPredictorLoss, L1Loss
    from seq2seq_vc.models import FastSpeechVC as RefFS

    NARVCTrainer = _nar_trainer(monkeypatch)
    torch.manual_seed(37)
    ref = RefFS(**FS_HP, **FS_FIXED)
    ref_shim.disable_dropout(ref)
    ref.train()
    ours = seq2seq_vc_b200.FastSpeechVC(**FS_HP, **FS_FIXED, **AAS_NO_DROPOUT)
    ours.load_state_dict(ref.state_dict())
    ours.train()
    _same_parameter_order(ref, ours)
    batch = _fs_batch()
    config = dict(FS_CONFIG, outdir=outdir)
    crit = lambda: {"L1Loss": L1Loss(), "DurationPredictorLoss": DurationPredictorLoss()}
    t_ref = _run(NARVCTrainer, ref, crit(), config, batch, 3)
    our_crit = crit() if criterions == "reference" else {"L1Loss": seq2seq_vc_b200.L1Loss(),
                                                         "DurationPredictorLoss": seq2seq_vc_b200.DurationPredictorLoss()}
    t_our = _run(NARVCTrainer, ours, our_crit, config, batch, 3)
    assert t_ref.steps == t_our.steps == 3
    for k in ("train/l1_loss", "train/duration_loss", "train/loss"):
        assert abs(t_ref.total_train_loss[k] - t_our.total_train_loss[k]) <= 1e-4 * max(1.0, abs(t_ref.total_train_loss[k])), k
    _compare_grads(t_ref, t_our)
    sd_ref, sd_our = ref.state_dict(), ours.state_dict()
    assert set(sd_ref) == set(sd_our)
    _close(sd_ref, sd_our)


def test_fused_narvc_train_step_matches_reference_trainer(trainers, monkeypatch):
    """seq2seq_vc_b200.NARVCTrainStep (own clip + Adam + WarmupLR on flat buffers) == NARVCTrainer._train_step around the reference
    FastSpeechVC: same logged losses, same parameters after 3 steps."""
    _, _, outdir = trainers
    import seq2seq_vc_b200
    from seq2seq_vc.losses import DurationPredictorLoss, L1Loss
    from seq2seq_vc.models import FastSpeechVC as RefFS

    NARVCTrainer = _nar_trainer(monkeypatch)
    torch.manual_seed(41)
    ref = RefFS(**FS_HP, **FS_FIXED)
    ref_shim.disable_dropout(ref)
    ref.train()
    ours = seq2seq_vc_b200.FastSpeechVC(**FS_HP, **FS_FIXED, **AAS_NO_DROPOUT)
    ours.load_state_dict(ref.state_dict())
    batch = _fs_batch()
    t_ref = _run(NARVCTrainer, ref, {"L1Loss": L1Loss(), "DurationPredictorLoss": DurationPredictorLoss()}, dict(FS_CONFIG, outdir=outdir), batch, 3)
    step = seq2seq_vc_b200.NARVCTrainStep(ours, lr=1e-3, warmup_steps=3, grad_norm=1.0)
    tot = torch.zeros(2)
    for _ in range(3):
        tot += step(batch["xs"], batch["ilens"].tolist(), batch["ys"], batch["olens"].tolist(), batch["durations"], batch["dp_inputs"]).float().cpu()
    assert step.steps == t_ref.steps == 3
    for i, k in enumerate(("train/l1_loss", "train/duration_loss")):
        assert abs(float(tot[i]) - t_ref.total_train_loss[k]) <= 2e-4 * max(1.0, abs(t_ref.total_train_loss[k])), k
    _close(ref.state_dict(), ours.engine.state_dict())
    ck = step.state_dict()                           # reference checkpoint layout (trainers/base.py:86-104)
    assert ck["steps"] == 3 and set(ck["model"]) == set(ref.state_dict())


def test_reference_narvc_trainer_eval_hook_with_dropin_inference(trainers, monkeypatch, tmp_path):
    """NARVCTrainer._genearete_and_save_intermediate_result (trainers/nar_vc.py:103-203) calls
    model.inference(x, spembs=None, dp_input=dp_input) on every (padded) row of the batch and unpacks (outs, d_outs): the drop-in's
    predicted durations are exact and everything drawn matches the reference model's."""
    _, _, outdir = trainers
    import seq2seq_vc.trainers.nar_vc as t_nar
    import seq2seq_vc_b200
    from seq2seq_vc.losses import L1Loss
    from seq2seq_vc.models import FastSpeechVC as RefFS

    torch.manual_seed(67)
    ref = RefFS(**FS_HP, **FS_FIXED)
    with torch.no_grad():
        ref.duration_predictor.linear.bias.add_(1.0)          # so that the predicted durations are not all zero
    ours = seq2seq_vc_b200.FastSpeechVC(**FS_HP, **FS_FIXED, **AAS_NO_DROPOUT)
    ours.load_state_dict(ref.state_dict())
    ref.eval()
    ours.eval()
    batch = _fs_batch()
    r = ref.inference(batch["xs"][0], spembs=None, dp_input=batch["dp_inputs"][0])
    o = ours.inference(batch["xs"][0], spembs=None, dp_input=batch["dp_inputs"][0])
    assert len(r) == len(o) == 2
    assert torch.equal(r[1].long(), o[1].long()) and r[0].shape == o[0].shape and (r[0] - o[0]).abs().max().item() <= 2e-4
    r2 = ref.inference(batch["xs"][1], spembs=None, dp_input=batch["dp_inputs"][1], alpha=1.3)
    o2 = ours.inference(batch["xs"][1], spembs=None, dp_input=batch["dp_inputs"][1], alpha=1.3)
    assert r2[0].shape == o2[0].shape and (r2[0] - o2[0]).abs().max().item() <= 2e-4
    drawn = []
    for i, model in enumerate((ref, ours)):
        config = dict(FS_CONFIG, outdir=str(tmp_path / f"run{i}"), num_save_intermediate_results=4)
        plt = _RecordingPlt()
        monkeypatch.setattr(t_nar, "plt", plt)
        tr = _run(t_nar.NARVCTrainer, model, {"L1Loss": L1Loss()}, config, batch, 0)
        tr.vocoder = None
        tr._genearete_and_save_intermediate_result(batch)
        drawn.append(plt.arrays)
    assert len(drawn[0]) == len(drawn[1]) > 0
    for a, b in zip(*drawn):
        assert a.shape == b.shape and np.abs(a - b).max() <= 2e-4


def _recipes():
    import glob

    return sorted(p.split("/egs/")[1] for p in glob.glob("/root/reference/egs/*/*/conf/*.yaml"))


@pytest.mark.parametrize("recipe", _recipes())
def test_every_shipped_recipe_constructs_the_dropin_with_the_reference_state_dict(recipe):
    """`getattr(seq2seq_vc_b200, config["model_type"])(**config["model_params"])` on every recipe under egs/ that defines a model
    (bin/vc_train.py:346-350, bin/tts_train.py:219): the drop-in builds -- none is refused -- and its parameters / buffers carry the
    reference model's names, shapes and registration order.  (Fine-tuning recipes without `model_params` take them from the
    pre-trained checkpoint's config and are skipped, as are the f0 range files.)"""
    import os

    import yaml

    ref_shim.install()
    import seq2seq_vc.models as ref_models
    import seq2seq_vc_b200

    with open(os.path.join("/root/reference/egs", recipe)) as f:
        config = yaml.load(f, Loader=yaml.Loader)
    if not isinstance(config, dict) or "model_params" not in config:
        pytest.skip("no model_params in this file")
    cls = config.get("model_type", "VTN")
    params = dict(config["model_params"])
    if cls == "TransformerTTS":
        params["idim"] = 78                                  # tts_train.py:219 writes the vocabulary size
    params.setdefault("idim", config.get("num_mels", 80))    # vc recipes carry idim / odim themselves; defaults for safety
    params.setdefault("odim", config.get("num_mels", 80))
    torch.manual_seed(0)
    ref = getattr(ref_models, cls)(**params)
    ours = getattr(seq2seq_vc_b200, cls)(**params)
    assert [(n, tuple(p.shape)) for n, p in ours.named_parameters()] == [(n, tuple(p.shape)) for n, p in ref.named_parameters()]
    assert {k: tuple(v.shape) for k, v in ours.state_dict().items()} == {k: tuple(v.shape) for k, v in ref.state_dict().items()}
