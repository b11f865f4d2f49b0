"""Data-parallel step (world size 2, gloo, CPU): the flat-gradient all-reduce + 1/world scaling inside the
Adam kernel equals one process stepping on the mean of the two ranks' gradients.  Kernels are replaced
by their CPU contracts (tests/fake_ops.py): this covers the N>1 host logic, NCCL runs on the GPU box."""
import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
HP = dict(idim=80, odim=80, dprenet_layers=2, dprenet_units=16, adim=32, aheads=2, elayers=1, eunits=48, dlayers=1,
          dunits=48, postnet_layers=2, postnet_filts=5, postnet_chans=16, decoder_reduction_factor=2,
          dprenet_dropout_rate=0.0, transformer_enc_dropout_rate=0.0, enc_positional_dropout_rate=0.0,
          dec_dropout_rate=0.0, dec_positional_dropout_rate=0.0, postnet_dropout_rate=0.0)


def _install_fakes():
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import fake_ops
    import seq2seq_vc_b200.ops as ops

    for n in fake_ops.ALL:
        if hasattr(ops, n) and n not in ("mas", "logmel"):
            setattr(ops, n, getattr(fake_ops, n))


def _batch(rank):
    from oracle import vtn_oracle

    return vtn_oracle.synthetic_batch(2, 40, 24, ilens=[40, 33], olens=[24, 17], seed=100 + rank)


def _worker(rank, world, port, out_dir):
    _install_fakes()
    import torch.distributed as dist

    from seq2seq_vc_b200 import VTNEngine, VTNTrainStep

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    eng = VTNEngine(HP, device="cpu", bf16=False, seed=5)
    step = VTNTrainStep(eng, lr=1e-3, warmup_steps=1, use_graph=False)
    xs, ilens, ys, labels, olens = _batch(rank)
    for _ in range(2):
        step(xs, ilens, ys, labels, olens)
    torch.save(eng.store.P.clone(), os.path.join(out_dir, f"p{rank}.pt"))
    dist.destroy_process_group()


def test_two_rank_step_matches_mean_gradient_step(tmp_path, monkeypatch):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    p0, p1 = torch.load(tmp_path / "p0.pt"), torch.load(tmp_path / "p1.pt")
    assert torch.equal(p0, p1), "replicas diverged"

    # single process: average the two ranks' gradients by hand, then the same optimizer tail
    import fake_ops
    from seq2seq_vc_b200 import VTNEngine, VTNTrainStep

    fake_ops.install(monkeypatch)
    engs = [VTNEngine(HP, device="cpu", bf16=False, seed=5) for _ in range(2)]
    ref = VTNEngine(HP, device="cpu", bf16=False, seed=5)
    stepper = VTNTrainStep(ref, lr=1e-3, warmup_steps=1)
    for it in range(2):
        g = torch.zeros_like(ref.store.G)
        for r, e in enumerate(engs):
            e.store.P.copy_(ref.store.P)
            for k in ref.buffers:
                pass
            xs, ilens, ys, labels, olens = _batch(r)
            e.forward(xs, ys, ilens, olens)
            e.loss(ys, labels)
            e.backward(e.d_after, e.d_before, e.d_logits)
            g += e.store.G
        ref.store.G.copy_(g / 2)
        stepper.steps += 1
        ref.lr_dev.fill_(stepper.lr_at(stepper.steps))
        ref.optimizer_step(1.0)
    assert (ref.store.P - p0).abs().max().item() <= 1e-6


# ---------------------------------------------------------------------------------------------- AAS-VC (same sharding rule)
AAS_HP = dict(idim=80, odim=80, adim=32, aheads=2, elayers=1, eunits=48, dlayers=1, dunits=48, duration_predictor_input_dim=80,
              duration_predictor_layers=2, duration_predictor_chans=16, duration_predictor_kernel_size=3, postnet_layers=2,
              postnet_filts=5, postnet_chans=16, post_encoder_reduction_factor=4, conformer_enc_kernel_size=7,
              conformer_dec_kernel_size=7, transformer_enc_dropout_rate=0.0, transformer_enc_positional_dropout_rate=0.0,
              transformer_enc_attn_dropout_rate=0.0, transformer_dec_dropout_rate=0.0, transformer_dec_positional_dropout_rate=0.0,
              transformer_dec_attn_dropout_rate=0.0, duration_predictor_dropout_rate=0.0, postnet_dropout_rate=0.0)


def _install_fakes_all():
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import fake_ops
    import seq2seq_vc_b200.ops as ops

    for n in fake_ops.ALL:
        if hasattr(ops, n) and n != "logmel":
            setattr(ops, n, getattr(fake_ops, n))


def _aas_batch(rank):
    from oracle import aasvc_oracle

    return aasvc_oracle.synthetic_batch(2, 44, 36, ilens=[44, 37], olens=[36, 29], seed=200 + rank)


def _aas_worker(rank, world, port, out_dir):
    _install_fakes_all()
    import torch.distributed as dist

    from seq2seq_vc_b200 import AASVCEngine, AASVCTrainStep

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    eng = AASVCEngine(AAS_HP, device="cpu", bf16=False, seed=5)
    step = AASVCTrainStep(eng, lr=1e-3, warmup_steps=1, use_graph=False)
    xs, ilens, ys, olens, dpi = _aas_batch(rank)
    for _ in range(2):
        step(xs, ilens, ys, olens, dpi)
    torch.save(eng.store.P.clone(), os.path.join(out_dir, f"a{rank}.pt"))
    dist.destroy_process_group()


def test_aasvc_two_rank_step_matches_mean_gradient_step(tmp_path, monkeypatch):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_aas_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    p0, p1 = torch.load(tmp_path / "a0.pt"), torch.load(tmp_path / "a1.pt")
    assert torch.equal(p0, p1), "replicas diverged"

    import fake_ops
    from seq2seq_vc_b200 import AASVCEngine, AASVCTrainStep

    fake_ops.install(monkeypatch)
    engs = [AASVCEngine(AAS_HP, device="cpu", bf16=False, seed=5) for _ in range(2)]
    ref = AASVCEngine(AAS_HP, device="cpu", bf16=False, seed=5)
    stepper = AASVCTrainStep(ref, lr=1e-3, warmup_steps=1)
    for it in range(2):
        g = torch.zeros_like(ref.store.G)
        for r, e in enumerate(engs):
            e.store.P.copy_(ref.store.P)
            xs, ilens, ys, olens, dpi = _aas_batch(r)
            e.forward(xs, ys, dpi, ilens, olens)
            e.loss(ys, duration_loss=it > 0)          # trainers/aas_vc.py:113: no duration loss at step 0
            e.backward()
            g += e.store.G
        ref.store.G.copy_(g / 2)
        stepper.steps += 1
        ref.lr_dev.fill_(stepper.lr_at(stepper.steps))
        ref.optimizer_step(1.0, duration_predictor_active=it > 0)      # the predictor's Adam clock starts with its loss
    assert (ref.store.P - p0).abs().max().item() <= 1e-6


# ------------------------------------------------------------- gradient accumulation (trainers/aas_vc.py:141-149)
def _aas_micro_batch(rank, micro):
    from oracle import aasvc_oracle

    return aasvc_oracle.synthetic_batch(2, 44, 36, ilens=[44, 37], olens=[36, 29], seed=300 + 10 * micro + rank)


def _aas_accum_worker(rank, world, port, out_dir):
    _install_fakes_all()
    import torch.distributed as dist

    from seq2seq_vc_b200 import AASVCEngine, AASVCTrainStep

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    calls = []
    real = dist.all_reduce
    dist.all_reduce = lambda *a, **k: (calls.append(1), real(*a, **k))[1]
    eng = AASVCEngine(AAS_HP, device="cpu", bf16=False, seed=5)
    step = AASVCTrainStep(eng, lr=1e-3, warmup_steps=1, use_graph=False, gradient_accumulate_steps=2)
    for it in range(2):
        for micro in range(2):
            before = eng.store.P.clone()
            step(*_aas_micro_batch(rank, micro))
            if micro == 0:      # non-boundary micro-step: no collective, no optimizer work
                assert torch.equal(before, eng.store.P) and len(calls) == it and step.steps == it
    assert len(calls) == 2 and step.steps == 2 and step.backward_steps == 4
    torch.save(eng.store.P.clone(), os.path.join(out_dir, f"acc{rank}.pt"))
    dist.destroy_process_group()


def test_aasvc_gradient_accumulation_two_ranks(tmp_path, monkeypatch):
    """2 ranks x 2 micro-steps == one process stepping on the mean of the four gradients; the all-reduce runs once per
    optimizer step only."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_aas_accum_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    p0, p1 = torch.load(tmp_path / "acc0.pt"), torch.load(tmp_path / "acc1.pt")
    assert torch.equal(p0, p1), "replicas diverged"

    import fake_ops
    from seq2seq_vc_b200 import AASVCEngine, AASVCTrainStep

    fake_ops.install(monkeypatch)
    e = AASVCEngine(AAS_HP, device="cpu", bf16=False, seed=5)
    ref = AASVCEngine(AAS_HP, device="cpu", bf16=False, seed=5)
    stepper = AASVCTrainStep(ref, lr=1e-3, warmup_steps=1)
    for it in range(2):
        g = torch.zeros_like(ref.store.G)
        e.store.P.copy_(ref.store.P)
        for r in range(2):
            for micro in range(2):
                xs, ilens, ys, olens, dpi = _aas_micro_batch(r, micro)
                e.forward(xs, ys, dpi, ilens, olens)
                e.loss(ys, duration_loss=it > 0)
                e.backward()
                g += e.store.G
        ref.store.G.copy_(g / 4)
        stepper.steps += 1
        ref.lr_dev.fill_(stepper.lr_at(stepper.steps))
        ref.optimizer_step(1.0, duration_predictor_active=it > 0)
    assert (ref.store.P - p0).abs().max().item() <= 1e-6


def test_aasvc_gradient_accumulation_rejects_zero():
    from seq2seq_vc_b200 import AASVCTrainStep

    with pytest.raises(ValueError):
        AASVCTrainStep(object(), gradient_accumulate_steps=0)


# ------------------------------------------------------------- FastSpeechVC (NARVCTrainStep), world size 2
FS_HP = dict(idim=80, odim=80, adim=32, aheads=2, elayers=1, eunits=48, dlayers=1, dunits=48, duration_predictor_input_dim=80,
             duration_predictor_layers=2, duration_predictor_chans=16, duration_predictor_kernel_size=3, postnet_layers=2, postnet_filts=5,
             postnet_chans=16, conformer_enc_kernel_size=7, conformer_dec_kernel_size=7, transformer_enc_dropout_rate=0.0,
             transformer_enc_positional_dropout_rate=0.0, transformer_enc_attn_dropout_rate=0.0, transformer_dec_dropout_rate=0.0,
             transformer_dec_positional_dropout_rate=0.0, transformer_dec_attn_dropout_rate=0.0, duration_predictor_dropout_rate=0.0,
             postnet_dropout_rate=0.0)


def _fs_batch(rank):
    g = torch.Generator().manual_seed(400 + rank)
    B, T = 2, 46
    ilens = [46, 39 - rank]
    tl = [((i - 2 + 1) // 2 - 2 + 1) // 2 for i in ilens]
    xs = torch.randn(B, T, 80, generator=g)
    ds = torch.randint(0, 4, (B, max(tl)), generator=g)
    for b in range(B):
        xs[b, ilens[b]:] = 0
        ds[b, tl[b]:] = 0
        ds[b, 0] = max(int(ds[b, 0]), 1)
    olens = ds.sum(1).tolist()
    ys = torch.randn(B, max(olens), 80, generator=g)
    for b in range(B):
        ys[b, olens[b]:] = 0
    return xs, ilens, ys, olens, ds


def _fs_worker(rank, world, port, out_dir):
    _install_fakes_all()
    import torch.distributed as dist

    from seq2seq_vc_b200 import NARVCTrainStep
    from seq2seq_vc_b200.fsvc_engine import FastSpeechVCEngine

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    eng = FastSpeechVCEngine(FS_HP, device="cpu", bf16=False, seed=5)
    step = NARVCTrainStep(eng, lr=1e-3, warmup_steps=1, use_graph=False)
    xs, ilens, ys, olens, ds = _fs_batch(rank)
    for _ in range(2):
        step(xs, ilens, ys, olens, ds, xs)
    torch.save(eng.store.P.clone(), os.path.join(out_dir, f"f{rank}.pt"))
    dist.destroy_process_group()


def test_fastspeech_vc_two_rank_step_matches_mean_gradient_step(tmp_path, monkeypatch):
    """NARVCTrainStep under a world-size-2 gloo group (rank-specific ragged batches and durations): the replicas stay identical and
    equal one process stepping on the mean of the two gradients -- the path shards by utterance batch, one all-reduce per step."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_fs_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    p0, p1 = torch.load(tmp_path / "f0.pt"), torch.load(tmp_path / "f1.pt")
    assert torch.equal(p0, p1), "replicas diverged"

    import fake_ops
    from seq2seq_vc_b200 import NARVCTrainStep
    from seq2seq_vc_b200.fsvc_engine import FastSpeechVCEngine

    fake_ops.install(monkeypatch)
    engs = [FastSpeechVCEngine(FS_HP, device="cpu", bf16=False, seed=5) for _ in range(2)]
    ref = FastSpeechVCEngine(FS_HP, device="cpu", bf16=False, seed=5)
    stepper = NARVCTrainStep(ref, lr=1e-3, warmup_steps=1)
    for it in range(2):
        g = torch.zeros_like(ref.store.G)
        for r, e in enumerate(engs):
            e.store.P.copy_(ref.store.P)
            e.seed_dev.copy_(ref.seed_dev)
            xs, ilens, ys, olens, ds = _fs_batch(r)
            e.forward(xs, ys, ds, xs, ilens, olens)
            e.loss(ys)
            e.backward()
            g += e.store.G
        ref.store.G.copy_(g / 2)
        stepper.steps += 1
        ref.lr_dev.fill_(stepper.lr_at(stepper.steps))
        ref.optimizer_step(1.0)
    assert (ref.store.P - p0).abs().max().item() <= 1e-6
