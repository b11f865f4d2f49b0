#!/usr/bin/env python
"""Benchmark of the seq2seq-vc training hot path on B200 (contract: see the task statement / DESIGN.md).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference|torch_gpu] [--workload c2|c1|c2b64|c3|c3b16|c4|c5]

  --impl reference : the UNMODIFIED reference modules (baseline/_ref, installed by __graft_entry__.build()) on the host cores
  --impl torch_gpu : the same reference modules on cuda (fp32 and bf16 autocast): the PyTorch-eager incumbent on this box
  --workload c5    : STFT -> log-mel feature extraction (BASELINE.json configs[4]), HBM roofline

One "step" = one full VTN training step (forward + Seq2SeqLoss + backward + gradient all-reduce +
clip_grad_norm + Adam) over one synthetic padded mel batch.  Default workload = BASELINE.json
configs[1]: VTN-base (6+6 layers, d=384, 8 heads, r=2), bf16 compute, batch 32 x (512 -> 1024
frames, 80 mel) per GPU.  Metric: target mel-frames / second, whole job (sum over GPUs).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOADS = {
    # name: (model hparams, per-GPU batch, T, L, bf16)
    "c2": (dict(idim=80, odim=80, adim=384, aheads=8, elayers=6, dlayers=6, eunits=1536, dunits=1536, decoder_reduction_factor=2),
           32, 512, 1024, True, "VTN-base 6+6 d384 h8 r2, B32 x (512->1024, 80-mel), bf16"),
    "c2b64": (dict(idim=80, odim=80, adim=384, aheads=8, elayers=6, dlayers=6, eunits=1536, dunits=1536, decoder_reduction_factor=2),
              64, 512, 1024, True, "VTN-base 6+6 d384 h8 r2, B64 x (512->1024, 80-mel), bf16"),
    "c4": (dict(idim=80, odim=80, adim=384, aheads=4, elayers=6, dlayers=6, eunits=1536, dunits=1536, decoder_reduction_factor=2),
           64, 160, 1000, True, "TransformerTTS 6+6 d384 h4 r2, B64 x (160 tokens -> 1000 frames, 80-mel), bf16"),
    # AAS-VC (egs/arctic/vc2/conf/aas_vc.melmelmel.v1.yaml, deterministic duration predictor): BASELINE.json configs[2]
    "c3": (dict(idim=80, odim=80, adim=384, aheads=2, elayers=4, eunits=1536, dlayers=4, dunits=1536, duration_predictor_input_dim=80,
                duration_predictor_layers=2, duration_predictor_chans=256, duration_predictor_kernel_size=3, postnet_layers=5,
                postnet_filts=5, postnet_chans=256, post_encoder_reduction_factor=4, conformer_enc_kernel_size=15,
                conformer_dec_kernel_size=15),
           64, 768, 768, True, "AAS-VC Conformer 4+4 (enc d384, dec d1536, h2, k15), B64 x (768->768, 80-mel), bf16, MAS + forward-sum on device"),
    "c3b16": (dict(idim=80, odim=80, adim=384, aheads=2, elayers=4, eunits=1536, dlayers=4, dunits=1536, duration_predictor_input_dim=80,
                   duration_predictor_layers=2, duration_predictor_chans=256, duration_predictor_kernel_size=3, postnet_layers=5,
                   postnet_filts=5, postnet_chans=256, post_encoder_reduction_factor=4, conformer_enc_kernel_size=15,
                   conformer_dec_kernel_size=15),
              16, 768, 768, True, "AAS-VC Conformer 4+4 (enc d384, dec d1536, h2, k15), B16 x (768->768, 80-mel), bf16 (the recipe's batch size)"),
    # the shipped yaml UNMODIFIED: stochastic duration predictor (VITS flows, egs/arctic/vc2/conf/aas_vc.melmelmel.v1.yaml:57)
    "c3s": (dict(idim=80, odim=80, adim=384, aheads=2, elayers=4, eunits=1536, dlayers=4, dunits=1536, duration_predictor_input_dim=80,
                 duration_predictor_layers=2, duration_predictor_chans=256, duration_predictor_kernel_size=3, postnet_layers=5,
                 postnet_filts=5, postnet_chans=256, post_encoder_reduction_factor=4, conformer_enc_kernel_size=15,
                 conformer_dec_kernel_size=15, duration_predictor_type="stochastic"),
            64, 768, 768, True, "AAS-VC Conformer 4+4 (enc d384, dec d1536, h2, k15) with the stochastic duration predictor, B64 x (768->768, 80-mel), bf16"),
    # FastSpeechVC (egs/arctic/vc2/conf/fs2_vc.melmelmel.v1.yaml): conformer 4+4 behind the conv2d input layer, teacher durations;
    # a widening workload, not a BASELINE.json config
    "c6": (dict(idim=80, odim=80, adim=384, aheads=2, elayers=4, eunits=1536, dlayers=4, dunits=1536, duration_predictor_input_dim=80,
                duration_predictor_layers=2, duration_predictor_chans=256, duration_predictor_kernel_size=3, postnet_layers=5,
                postnet_filts=5, postnet_chans=256, conformer_enc_kernel_size=15, conformer_dec_kernel_size=15,
                teacher_model_decoder_reduction_factor=1),
           64, 768, 768, True, "FastSpeechVC Conformer 4+4 (d384, h2, k15, conv2d input layer, teacher durations), B64 x (768->768, 80-mel), bf16"),
    "c1": (dict(idim=80, odim=80, adim=256, aheads=4, elayers=2, dlayers=2, eunits=1024, dunits=1024, decoder_reduction_factor=2),
           4, 200, 400, False, "VTN-small 2+2 d256 h4 r2, B4 x (200->400, 80-mel), fp32"),
    # STFT -> log-mel (BASELINE.json configs[4]): 256 clips x 10 s @ 48 kHz, n_fft 2048, hop 300, 80 mels
    "c5": (dict(sr=48000, fft_size=2048, hop_size=300, num_mels=80), 256, 480000, 1601, False,
           "STFT->log-mel, 256 x 10 s clips @ 48 kHz, n_fft 2048, hop 300, 80 mels, fp32"),
}
METRIC = "target mel-frames/sec, VTN-base enc-dec training step (80-mel, src512/tgt1024)"
METRIC_C5 = "mel-frames/sec, STFT->log-mel feature extraction (48 kHz, n_fft 2048, hop 300, 80-mel)"
METRIC_FS = "target mel-frames/sec, FastSpeechVC Conformer non-AR training step (80-mel, 768 frames)"
METRIC_AAS = "target mel-frames/sec, AAS-VC Conformer non-AR training step (80-mel, 768 frames)"


def is_aas(workload):
    return workload.startswith("c3")


def aasvc_fwd_flops(hp, T, L):
    """Forward FLOPs per utterance of AASVC (linear_pos is batch-independent and left out)."""
    d, pr = hp["adim"], hp["post_encoder_reduction_factor"]
    C, Tt = d * pr, T // pr

    def conformer(n, dm, U, K, Tn):
        per = 8 * Tn * dm * U + 8 * Tn * dm * dm + 4 * Tn * Tn * dm + 2 * Tn * (2 * Tn - 1) * dm + 6 * Tn * dm * dm + 2 * Tn * dm * K
        return n * per

    T1, F1 = (T - 1) // 2, 39
    T2, F2 = (T1 - 1) // 2, 19
    dp_proj = 2 * 9 * d * T1 * F1 + 2 * 9 * d * d * T2 * F2 + 2 * T2 * (d * F2) * d
    enc = 2 * T * 80 * d + conformer(hp["elayers"], d, hp["eunits"], hp["conformer_enc_kernel_size"], T)
    align = 2 * Tt * 4 * C * C + 2 * L * (3 * 80 * C + 4 * C * C) + 3 * L * Tt * C
    up = 2 * L * Tt * C
    dec = conformer(hp["dlayers"], C, hp["dunits"], hp["conformer_dec_kernel_size"], L) + 2 * L * C * 80
    post = 2 * 5 * L * (80 * 256 + 3 * 256 * 256 + 256 * 80)
    return dp_proj + enc + align + up + dec + post

UNIT = "frames/s"


def vtn_fwd_flops(hp, T, L, tts=False):
    """Forward FLOPs per utterance (SURVEY.md section 8d formulas)."""
    d, r = hp["adim"], hp["decoder_reduction_factor"]
    Ue, Ud = hp["eunits"], hp["dunits"]
    T1, F1 = (T - 1) // 2, 39
    T2, F2 = (T1 - 1) // 2, 19
    Lr = L // r
    conv = 2 * 9 * d * T1 * F1 + 2 * 9 * d * d * T2 * F2 + 2 * T2 * (d * F2) * d
    if tts:
        T2, conv = T + 1, 0
    enc = hp["elayers"] * (8 * T2 * d * d + 4 * T2 * T2 * d + 4 * T2 * d * Ue)
    pre = 2 * Lr * (80 * 256 + 256 * 256 + 256 * d)
    dec = hp["dlayers"] * (8 * Lr * d * d + 4 * Lr * Lr * d + 4 * Lr * d * d + 4 * T2 * d * d + 4 * Lr * T2 * d + 4 * Lr * d * Ud)
    heads = 2 * Lr * d * (80 * r + r)
    post = 2 * 5 * (Lr * r) * (80 * 256 + 3 * 256 * 256 + 256 * 80)
    return conv + enc + pre + dec + heads + post


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons for this rank's GPU during the timed region."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self._stop_evt = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.15)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        reasons = []
        for i, name in enumerate(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")):
            if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def synthetic_batch(B, T, L, seed, tts=False):
    g = torch.Generator().manual_seed(seed)
    xs = torch.randint(1, 79, (B, T), generator=g) if tts else torch.randn(B, T, 80, generator=g)
    ys = torch.randn(B, L, 80, generator=g)
    labels = torch.zeros(B, L)
    labels[:, L - 1:] = 1.0
    return xs, [T] * B, ys, labels, [L] * B


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


# --------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path on the host cores
# --------------------------------------------------------------------------------------------------
def cpu_port_steps(hp, B, T, L, steps, warmup, tts=False):
    """fwd + Seq2SeqLoss + bwd + clip + Adam of the oracle (plain torch fp32 CPU) on B utterances."""
    from oracle import vtn_oracle

    torch.set_num_threads(os.cpu_count() or 1)
    ohp = vtn_oracle.default_hparams(**hp)
    sd = vtn_oracle.init_state_dict(ohp, seed=0)
    if tts:
        sd = {k: v for k, v in sd.items() if not k.startswith("encoder.embed.")}
        emb = torch.randn(hp["idim"], hp["adim"])
        emb[0] = 0
        sd["encoder.embed.0.weight"] = emb
        sd["encoder.embed.1.alpha"] = torch.tensor(1.0)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.dtype.is_floating_point and "running_" not in k}
    full = dict(sd)
    full.update(params)
    opt = torch.optim.Adam(list(params.values()), lr=8e-5)
    xs, ilens, ys, labels, olens = synthetic_batch(B, T, L, 1234, tts)
    times = []
    fwd = vtn_oracle.tts_forward if tts else vtn_oracle.vtn_forward
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        out = fwd(full, ohp, xs, ilens, ys, labels, olens, training=True)
        l1, bce = vtn_oracle.seq2seq_loss(out["after_outs"], out["before_outs"], out["logits"], out["ys"], out["labels"], out["olens"])
        opt.zero_grad()
        (l1 + bce).backward()
        torch.nn.utils.clip_grad_norm_(list(params.values()), 1.0)
        opt.step()
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return sum(times) / len(times)


def cpu_port_steps_aas(hp, B, T, L, steps, warmup):
    """fwd + the four AAS-VC losses + bwd + clip + Adam of the oracle (plain torch fp32 CPU) on B utterances."""
    from oracle import aasvc_oracle as ao

    torch.set_num_threads(os.cpu_count() or 1)
    sd = ao.init_state_dict(hp, seed=0)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.dtype.is_floating_point and "running_" not in k}
    full = dict(sd)
    full.update(params)
    opt = torch.optim.Adam(list(params.values()), lr=8e-5)
    xs, ilens, ys, olens, dpi = ao.synthetic_batch(B, T, L, seed=1234)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        out = ao.aasvc_forward(full, hp, xs, ilens, ys, olens, dpi, training=True)
        total, _ = ao.aasvc_losses(out)
        opt.zero_grad()
        total.backward()
        torch.nn.utils.clip_grad_norm_(list(params.values()), 1.0)
        opt.step()
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return sum(times) / len(times)


def fs_batch(B, T, L, seed):
    """Synthetic FastSpeechVC batch: mels + ragged teacher durations, every utterance summing to L (the data loader trims the target
    to the duration sum, fs2_vc.melmelmel.v1.yaml: teacher_duration_reduction_factor)."""
    g = torch.Generator().manual_seed(seed)
    xs, ys = torch.randn(B, T, 80, generator=g), torch.randn(B, L, 80, generator=g)
    Tt = ((T - 1) // 2 - 1) // 2
    ds = torch.zeros(B, Tt, dtype=torch.int64)
    for b in range(B):
        cuts = torch.sort(torch.randint(0, L + 1, (Tt - 1,), generator=g)).values
        ds[b] = torch.diff(torch.cat([torch.zeros(1, dtype=torch.int64), cuts, torch.full((1,), L, dtype=torch.int64)]))
    assert int(ds.sum(1).min()) == L == int(ds.sum(1).max())
    return xs, ys, ds


def cpu_port_steps_fs(hp, B, T, L, steps, warmup):
    """fwd + L1 / duration losses + bwd + clip + Adam of the FastSpeechVC oracle (plain torch fp32 CPU) on B utterances."""
    from oracle import fsvc_oracle as fo
    from seq2seq_vc_b200.fsvc_engine import param_groups, buffer_specs, default_hparams

    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(0)
    full_hp = default_hparams(**hp)
    sd = {}
    for grp in param_groups(full_hp):
        for name, shape in grp:
            sd[name] = (torch.randn(*shape, generator=g) * (0.05 if len(shape) > 1 else 0.0) + (1.0 if name.endswith(("norm.weight", "norm1.weight", "norm2.weight")) else 0.0))
    for name, shape, dtype in buffer_specs(full_hp):
        sd[name] = torch.ones(shape, dtype=dtype) if name.endswith("running_var") else torch.zeros(shape, dtype=dtype)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.dtype.is_floating_point and "running_" not in k}
    full = dict(sd)
    full.update(params)
    opt = torch.optim.Adam(list(params.values()), lr=8e-5)
    xs, ys, ds = fs_batch(B, T, L, 1234)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        out = fo.fsvc_forward(full, hp, xs, [T] * B, ys, [L] * B, ds, xs, training=True)
        total, _ = fo.fsvc_losses(out, ds)
        opt.zero_grad()
        total.backward()
        torch.nn.utils.clip_grad_norm_(list(params.values()), 1.0)
        opt.step()
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return sum(times) / len(times)


def reference_modules():
    """The unmodified reference package from baseline/_ref (pip-installed + completed by __graft_entry__.build()), imported
    through oracle/ref_shim.py (lazy numba.jit, the missing diffsinger module).  None when it is not there."""
    root = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(root, "seq2seq_vc", "modules")):
        return None
    os.environ["S2SVC_REFERENCE_ROOT"] = root
    try:
        from oracle import ref_shim

        ref_shim.install()
        import seq2seq_vc.losses as losses
        import seq2seq_vc.models as models
        from seq2seq_vc.schedulers.warmup_lr import WarmupLR
    except Exception as e:          # noqa: BLE001 -- any import problem means "fall back to the port", reported in the line
        sys.stderr.write(f"reference import failed: {e!r}\n")
        return None
    return models, losses, WarmupLR


def reference_step_fn(mods, workload, B, device, autocast=False, seed=1234):
    """ARVCTrainer._train_step / ARTTSTrainer._train_step (trainers/ar_vc.py:59-107) around the reference's own VTN /
    TransformerTTS, Seq2SeqLoss, torch.optim.Adam, WarmupLR and clip_grad_norm_, on `device`.  Returns (step(), frames/step)."""
    models, losses, WarmupLR = mods
    hp, _, T, L, _, _ = WORKLOADS[workload]
    tts = workload == "c4"
    torch.manual_seed(0)
    model = (models.TransformerTTS if tts else models.VTN)(**hp).to(device)
    model.train()
    crit = losses.Seq2SeqLoss().to(device)
    opt = torch.optim.Adam(model.parameters(), lr=8e-5)
    sched = WarmupLR(opt, warmup_steps=4000)
    xs, ilens, ys, labels, olens = synthetic_batch(B, T, L, seed, tts)
    xs, ys, labels = xs.to(device), ys.to(device), labels.to(device)
    ilens_t, olens_t = torch.tensor(ilens, device=device), torch.tensor(olens, device=device)

    def step():
        with torch.autocast(device_type="cuda", dtype=torch.bfloat16, enabled=autocast):
            after, before, logits, ys_, labels_, olens_, _ = model(xs, ilens_t, ys, labels, olens_t)
            l1, bce = crit(after, before, logits, ys_, labels_, olens_)
        loss = l1 + bce
        opt.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        sched.step()
        return loss

    return step, B * L


def cpu_reference_steps(mods, workload, B, steps, warmup):
    torch.set_num_threads(os.cpu_count() or 1)
    step, _ = reference_step_fn(mods, workload, B, torch.device("cpu"))
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        step()
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return sum(times) / len(times)


def logmel_cpu_frames_per_s(clips, n_samples, params):
    """Oracle restatement of bin/preprocess.py:30-92 (librosa semantics) on `clips` clips, all host cores (joblib)."""
    import numpy as np
    from joblib import Parallel, delayed

    from oracle import logmel_oracle

    rng = np.random.default_rng(1234)
    wavs = np.clip(0.1 * rng.standard_normal((clips, n_samples)), -1, 1).astype(np.float32)
    cores = os.cpu_count() or 1
    f = lambda w: logmel_oracle.logmelfilterbank(w, params["sr"], fft_size=params["fft_size"], hop_size=params["hop_size"],
                                                 num_mels=params["num_mels"]).shape[0]
    f(wavs[0])
    t0 = time.perf_counter()
    frames = sum(Parallel(n_jobs=cores)(delayed(f)(w) for w in wavs))
    return frames / (time.perf_counter() - t0), cores


def run_reference(args, rank):
    if rank != 0:
        return
    hp, B, T, L, bf16, desc = WORKLOADS[args.workload]
    if args.workload == "c5":
        clips = max(2, min(2 * (os.cpu_count() or 1), 64))
        val, cores = logmel_cpu_frames_per_s(clips, T, hp)
        line = {"impl": "reference", "metric": METRIC_C5, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": clips * L / val * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": {"workload": desc, "sample": f"{clips} clips (of {B}) over {cores} processes"},
                "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                                 "sample": f"numpy/scipy restatement of preprocess.py:30-92 with librosa's conventions (librosa itself is not "
                                           f"installable offline), {clips} x 10 s clips, joblib over {cores} cores"},
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return
    mods = reference_modules() if not (is_aas(args.workload) or args.workload == "c6") else None
    if mods is not None:
        # the reference's own modules; the full batch when a step stays within ~a minute, else a bounded sample (stated)
        import psutil

        roomy = psutil.virtual_memory().available >= (96 << 30) and (os.cpu_count() or 1) >= 16
        Bs = B if (args.workload == "c1" or (roomy and B <= 32)) else min(B, 4)      # ~0.3 GB of fp32 activations per C2 utterance
        timed = max(1, min(args.steps, 5 if Bs <= 4 else 2))
        sec = cpu_reference_steps(mods, args.workload, Bs, timed, 1)
        val = Bs * L / sec
        cores = os.cpu_count() or 1
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": desc, "same_config": Bs == B,
                           "sample": f"{Bs} utterances per step (of {B}); frames/s scales linearly in B", "timed_steps": timed},
                "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "reference",
                                 "sample": f"unmodified reference modules (baseline/_ref: VTN/TransformerTTS + Seq2SeqLoss + Adam + WarmupLR + "
                                           f"clip_grad_norm_, trainers/ar_vc.py:59-107), fp32, {Bs} x ({T}->{L}) per step, "
                                           f"{torch.get_num_threads()} threads, {timed} steps after 1 warm-up"},
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return
    aas = is_aas(args.workload) or args.workload == "c6"
    Bs = 1 if aas else min(B, 4)
    timed = max(1, min(args.steps, 2 if aas else 6))      # bounded sample: the CPU arm must end within minutes whatever K is
    if args.workload == "c6":
        sec = cpu_port_steps_fs(hp, Bs, T, L, timed, 1)
    elif aas:
        sec = cpu_port_steps_aas(hp, Bs, T, L, timed, 1)
    else:
        sec = cpu_port_steps(hp, Bs, T, L, timed, max(1, min(args.warmup, 1)), args.workload == "c4")
    val = Bs * L / sec
    cores = os.cpu_count() or 1
    line = {"impl": "reference", "metric": (METRIC_FS if args.workload == "c6" else METRIC_AAS) if aas else METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "sample": f"{Bs} utterances per step (of {B}); frames/s scales linearly in B",
                       "timed_steps": timed},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"oracle port of the reference PyTorch-CPU path, {Bs} x ({T}->{L}) per step, {torch.get_num_threads()} threads"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def gemm_roofline(stepper, batch, dev, table_path=None, run=None):
    """Per-launch CUDA-event timing of every tcgen05 GEMM launch of one (eager) training step."""
    from seq2seq_vc_b200 import ops

    rec = []
    orig = ops.gemm

    def timed(a, b, c, **kw):
        if kw.get("mode", 0) != 1:
            return orig(a, b, c, **kw)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = orig(a, b, c, **kw)
        e1.record()
        M = kw.get("M") or c.shape[-2]
        nb = 1
        for s in c.shape[:-2]:
            nb *= s
        key = (nb, M, c.shape[-1], a.shape[-1], kw.get("taps", 1), "Amn" if a.stride(-1) != 1 else "Ak",
               "Bmn" if b.stride(-1) != 1 else "Bk", str(c.dtype).replace("torch.", ""), bool(kw.get("accumulate")))
        rec.append((2.0 * nb * M * c.shape[-1] * a.shape[-1] * kw.get("taps", 1), e0, e1, key))
        return out

    orig_grouped = ops.gemm_grouped

    def timed_grouped(problems, mode=0):
        if mode != 1 or len(problems) < 2:      # single problems go through ops.gemm (timed above)
            return orig_grouped(problems, mode)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        orig_grouped(problems, mode)          # ops.gemm_grouped resolves the library itself: not re-entered through `timed`
        e1.record()
        fl = sum(2.0 * c.shape[-2] * c.shape[-1] * a.shape[-1] for a, b, c, kw in problems)
        rec.append((fl, e0, e1, ("grouped dW", len(problems))))

    ops.gemm = timed
    ops.gemm_grouped = timed_grouped
    try:
        eng = stepper.engine
        if run is None:
            xs, ilens, ys, labels, olens = batch
            eng.prepare(xs.shape[0], xs.shape[1], ys.shape[1], ilens, olens)
            run = lambda: stepper._fwd_bwd(xs, ys, labels)
        for _ in range(2):
            rec.clear()
            # park the GPU behind a ~40 ms spin so that the eager launches below queue up back to back: the
            # event pairs then bracket pure kernel execution instead of host launch latency
            torch.cuda._sleep(80_000_000)
            run()
            torch.cuda.synchronize()
    finally:
        ops.gemm = orig
        ops.gemm_grouped = orig_grouped
    flops = sum(r[0] for r in rec)
    ms = sum(r[1].elapsed_time(r[2]) for r in rec)
    if table_path:
        agg = {}
        for f, a, b, key in rec:
            e = agg.setdefault(key, [0, 0.0, 0.0])
            e[0] += 1
            e[1] += a.elapsed_time(b)
            e[2] += f
        with open(table_path, "w") as fh:
            fh.write("# (batch, M, N, K, taps, A-major, B-major, C dtype, accumulate): launches, total ms, TFLOP/s\n")
            for key, (n, t, f) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                fh.write(f"{key}: n={n} ms={t:.3f} tflops={f / (t * 1e-3) / 1e12:.1f}\n")
    return flops, ms, len(rec)


def incumbent_torch_gpu(workload, steps=5, warmup=3):
    """PyTorch-eager incumbent on this box: the UNMODIFIED reference modules on cuda, same step
    (trainers/ar_vc.py:59-107: forward, Seq2SeqLoss, backward, clip_grad_norm_, Adam, WarmupLR), fp32 and bf16 autocast."""
    mods = reference_modules()
    if mods is None:
        return {"unavailable": "baseline/_ref not importable"}
    hp, B, T, L, _, desc = WORKLOADS[workload]
    dev = torch.device("cuda", torch.cuda.current_device())
    out = {"what": "unmodified reference modules on cuda (cuBLAS / cuDNN / ATen), same batch and step", "batch": B}
    for name, ac in (("fp32", False), ("bf16_autocast", True)):
        try:
            torch.backends.cuda.matmul.allow_tf32 = False
            torch.backends.cudnn.allow_tf32 = False
            step, frames = reference_step_fn(mods, workload, B, dev, autocast=ac)
            for _ in range(warmup):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                loss = step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[name] = {"ms_per_step": ms, "frames_per_s": frames / (ms * 1e-3), "loss": float(loss)}
            del step
            torch.cuda.empty_cache()
        except Exception as e:      # noqa: BLE001
            out[name] = {"error": repr(e)[:200]}
    return out


def dropin_boundary(workload, steps=10, warmup=4, use_graph=False):
    """The reference trainer's step (trainers/ar_vc.py:59-107) around the DROP-IN module -- seq2seq_vc_b200.VTN -> Seq2SeqLoss ->
    backward -> clip_grad_norm_ -> torch.optim.Adam (+ WarmupLR) -- on the same batch as the fused step: what a user who only
    swaps the model / criterion classes gets, timed with CUDA events incl. the host-side glue."""
    from seq2seq_vc_b200 import VTN, Seq2SeqLoss, TransformerTTS

    hp, B, T, L, bf16, desc = WORKLOADS[workload]
    tts = workload == "c4"
    dev = torch.device("cuda", torch.cuda.current_device())
    model = (TransformerTTS if tts else VTN)(**hp, compute_dtype="bf16" if bf16 else "float32", device=dev, seed=0, use_graph=use_graph)
    model.train()
    crit = Seq2SeqLoss()
    opt = torch.optim.Adam(model.parameters(), lr=8e-5)
    xs, ilens, ys, labels, olens = synthetic_batch(B, T, L, 1234, tts)
    xs, ys, labels = xs.to(dev), ys.to(dev), labels.to(dev)
    ilens_t, olens_t = torch.tensor(ilens), torch.tensor(olens)          # the collater hands CPU int64 length tensors

    def step():
        after, before, logits, ys_, labels_, olens_, _ = model(xs, ilens_t, ys, labels, olens_t)
        l1, bce = crit(after, before, logits, ys_, labels_, olens_)
        loss = l1 + bce
        opt.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        return loss

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    out = {"ms_per_step": ms, "frames_per_s": B * L / (ms * 1e-3), "loss": float(loss),
           "what": "reference-style step around the drop-in module (model -> Seq2SeqLoss -> backward -> clip_grad_norm_ -> torch.optim.Adam), "
                   + ("forward / backward of the module replayed from CUDA graphs (use_graph=True)" if use_graph else "eager")}
    del model, opt
    torch.cuda.empty_cache()
    return out


def run_torch_gpu(args, rank):
    if rank != 0:
        return
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    hp, B, T, L, bf16, desc = WORKLOADS[args.workload]
    res = incumbent_torch_gpu(args.workload, steps=max(1, min(args.steps, 10)), warmup=max(3, min(args.warmup, 5)))
    best = max((v.get("frames_per_s", 0.0) for v in res.values() if isinstance(v, dict)), default=0.0)
    print(json.dumps({"impl": "torch_gpu", "metric": METRIC, "value": best, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
                      "warmup": args.warmup, "higher_is_better": True, "data": "synthetic", "config": {"workload": desc}, "detail": res}), flush=True)


def run_c6(args, rank, world):
    """FastSpeechVC training step (NARVCTrainer._train_step, trainers/nar_vc.py:52-103) through NARVCTrainStep; shards by utterance
    batch like the other models (gradient all-reduce only)."""
    from seq2seq_vc_b200 import FastSpeechVC, NARVCTrainStep, _lib

    hp, B, T, L, bf16, desc = WORKLOADS["c6"]
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _lib.device_check()
    xs, ys, ds = fs_batch(B, T, L, 1234 + rank)
    ilens, olens = [T] * B, [L] * B
    yaml_fixed = dict(positionwise_layer_type="linear", duration_predictor_use_encoder_outputs=False, encoder_normalize_before=True,
                      decoder_normalize_before=True, encoder_type="conformer", decoder_type="conformer", encoder_input_layer="conv2d",
                      transformer_enc_dropout_rate=0.2, transformer_enc_positional_dropout_rate=0.2, transformer_enc_attn_dropout_rate=0.2,
                      transformer_dec_dropout_rate=0.2, transformer_dec_positional_dropout_rate=0.2, transformer_dec_attn_dropout_rate=0.2)
    model = FastSpeechVC(**hp, **yaml_fixed, compute_dtype="bf16", device=dev, seed=0)
    step = NARVCTrainStep(model, lr=8e-5, warmup_steps=4000, use_graph=not args.no_graph)
    dev_in = [t.to(dev) for t in (xs, ys, ds)]
    pin_in = [t.pin_memory() for t in (xs, ys, ds)]
    call = lambda x, y, d: step(x, ilens, y, olens, d, x)          # duration_predictor_feat: mel -> dp_inputs are the source mels

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        call(*dev_in)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = _lib.launch_count() + step.replayed_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        call(*dev_in)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count() + step.replayed_launches - l0
    for _ in range(2):
        call(*pin_in).cpu()
    barrier()
    step.prefetch(pin_in[0], pin_in[1], pin_in[2], pin_in[0])
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for _ in range(args.steps):
        out = call(*pin_in)
        step.prefetch(pin_in[0], pin_in[1], pin_in[2], pin_in[0])
        host_losses = out.cpu()
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3)
    clocks = sampler.stop()
    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()
    if rank != 0:
        return
    assert all(v == v for v in host_losses.tolist()), "NaN loss"
    frames = B * L * world * args.steps
    pk, pk_src = peaks()
    line = {"metric": METRIC_FS, "value": frames / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": desc, "global_batch": B * world, "per_gpu_batch": B, "parallelism": f"dp{world}", "cuda_graph": not args.no_graph,
                       "l2": "per-step working set exceeds the 126 MB L2; no flush needed", "losses_last_step": host_losses.tolist(),
                       "tc_fallbacks": int(_lib.load().s2s_tc_fallback_count())},
            "clocks": clocks,
            "e2e": {"value": frames / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": 2 * xs.numel() * 4 + ys.numel() * 4 + ds.numel() * 8 + 3 * B * 4, "d2h_bytes_per_step": 8,
                    "h2d_overlap": "next step's inputs prefetched on a side stream (step.prefetch)"},
            "gpu_launches": int(launches)}
    if world == 1:
        step.engine.prepare(B, T, L, ilens, olens)
        run = lambda: step._fwd_bwd(dev_in[0], dev_in[1], dev_in[2], dev_in[0])
        gf, gms, n = gemm_roofline(step, None, dev, args.gemm_table, run)
        peak = pk.get("bf16_tflops_sustained", pk.get("bf16_tflops"))
        ach = gf / (gms * 1e-3) / 1e12
        line["roofline"] = {"bound": "tensor", "kernel": "gemm_tc_kernel (tcgen05.mma bf16, all shapes of one step)", "achieved": ach, "peak": peak,
                            "unit": "TFLOP/s", "frac": ach / peak, "traffic": None, "peak_source": pk_src + " (sustained)", "launches_per_step": n,
                            "algorithmic_gflop_per_launch_avg": gf / n / 1e9, "avg_launch_us": gms * 1e3 / n,
                            "gemm_share_of_step": gms / (ms / args.steps),
                            "how": "CUDA events around every mode-1 s2s_gemm launch of one fwd+bwd queued behind a GPU spin, right after the timed region"}
        if not args.no_cpu_baseline:
            sec = cpu_port_steps_fs(hp, 1, T, L, 1, 1)
            line["cpu_baseline"] = {"value": L / sec, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                                    "sample": f"oracle port of the reference PyTorch-CPU path (fp32), 1 x ({T}->{L}) per step (of {B}), 1 step after "
                                              f"1 warm-up, {torch.get_num_threads()} threads"}
    print(json.dumps(line), flush=True)


def run_c5(args, rank, world):
    """STFT -> log-mel (BASELINE.json configs[4]): every rank extracts its own 256 clips (the path shards by clip)."""
    from seq2seq_vc_b200 import _lib, api

    hp, B, ns, nf, _, desc = WORKLOADS["c5"]
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _lib.device_check()
    g = torch.Generator().manual_seed(1234 + rank)
    wav_h = (0.1 * torch.randn(B, ns, generator=g)).clamp_(-1, 1).pin_memory()
    wav = wav_h.to(dev)
    mel = torch.empty(B, nf, hp["num_mels"], device=dev)
    mel_h = torch.empty(B, nf, hp["num_mels"]).pin_memory()
    wav_stage = torch.empty_like(wav)
    kw = dict(fft_size=hp["fft_size"], hop_size=hp["hop_size"], num_mels=hp["num_mels"])

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        api.logmel_batch(wav, hp["sr"], out=mel, **kw)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = _lib.launch_count()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for a, b in evs:
        a.record()
        api.logmel_batch(wav, hp["sr"], out=mel, **kw)
        b.record()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    kernel_ms = sum(a.elapsed_time(b) for a, b in evs) / args.steps
    launches = _lib.launch_count() - l0
    for _ in range(2):
        wav_stage.copy_(wav_h, non_blocking=True)
        api.logmel_batch(wav_stage, hp["sr"], out=mel, **kw)
        mel_h.copy_(mel, non_blocking=True)
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for _ in range(args.steps):
        wav_stage.copy_(wav_h, non_blocking=True)           # H2D of the step's clips from pinned host memory
        api.logmel_batch(wav_stage, hp["sr"], out=mel, **kw)
        mel_h.copy_(mel, non_blocking=True)                 # D2H of the features (the product of this path)
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3)
    clocks = sampler.stop()
    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()
    if rank != 0:
        return
    assert torch.isfinite(mel_h).all()
    frames = B * nf * world * args.steps
    alg_bytes = wav.numel() * 4 + mel.numel() * 4
    pk, pk_src = peaks()
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r02_logmel_traffic.json")))["dram_bytes_per_launch"]
    except Exception:
        pass
    ach = alg_bytes / (kernel_ms * 1e-3) / 1e9
    line = {"metric": METRIC_C5, "value": frames / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "clips_per_gpu": B, "parallelism": f"dp{world} (clips sharded, no collective)",
                       "l2": "491 MB of input per step exceeds the 126 MB L2; no flush needed"},
            "clocks": clocks,
            "e2e": {"value": frames / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e / args.steps, "h2d_bytes_per_step": wav.numel() * 4,
                    "d2h_bytes_per_step": mel.numel() * 4},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "logmel kernels (STFT + mel + log of one batch)", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s",
                         "frac": ach / pk["hbm_gbs"], "traffic": traffic, "peak_source": pk_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_us": kernel_ms * 1e3,
                         "note": "SURVEY section 8(d): read 256 x 480000 x 4 B, write 256 x 1601 x 80 x 4 B = 622.7 MB per launch; DRAM traffic equals "
                                 "it (ncu: 609 MB), but every sample is used by 6.8 overlapping 2048-point FFTs, so the kernel is bound by instruction "
                                 "issue / shared-memory bandwidth of the FFT (1.83 k warp instructions and 465 shared-memory wavefronts per frame), "
                                 "not by HBM: see DESIGN.md section 4"}}
    if world == 1 and not args.no_cpu_baseline:
        clips = max(2, min(os.cpu_count() or 1, 32))
        val, cores = logmel_cpu_frames_per_s(clips, ns, hp)
        line["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"numpy/scipy restatement of preprocess.py:30-92 (librosa conventions), {clips} clips over {cores} processes"}
    print(json.dumps(line), flush=True)


def run_ours(args, rank, world):
    from seq2seq_vc_b200 import AASVC, AASVCTrainStep, VTN, TransformerTTS, VTNTrainStep, _lib

    if args.workload == "c5":
        return run_c5(args, rank, world)
    if args.workload == "c6":
        return run_c6(args, rank, world)
    hp, B, T, L, bf16, desc = WORKLOADS[args.workload]
    tts = args.workload == "c4"
    aas = is_aas(args.workload)
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _lib.device_check()
    xs, ilens, ys, labels, olens = synthetic_batch(B, T, L, 1234 + rank, tts)
    dxs, dys, dlabels = xs.to(dev), ys.to(dev), labels.to(dev)
    pxs, pys, plabels = xs.pin_memory(), ys.pin_memory(), labels.pin_memory()
    if aas:
        yaml_fixed = dict(positionwise_layer_type="linear", duration_predictor_use_encoder_outputs=False, encoder_normalize_before=True,
                          decoder_normalize_before=True, encoder_input_layer="linear",
                          **({} if "duration_predictor_type" in hp else {"duration_predictor_type": "deterministic"}),
                          transformer_enc_dropout_rate=0.2, transformer_enc_positional_dropout_rate=0.2,
                          transformer_enc_attn_dropout_rate=0.2, transformer_dec_dropout_rate=0.2,
                          transformer_dec_positional_dropout_rate=0.2, transformer_dec_attn_dropout_rate=0.2)
        model = AASVC(**hp, **yaml_fixed, compute_dtype="bf16" if bf16 else "float32", device=dev, seed=0)
        inner = AASVCTrainStep(model, lr=8e-5, warmup_steps=4000, use_graph=not args.no_graph)
        inner.steps = 1                                   # past dp_train_start_steps: the duration loss is part of every timed step

        class _Adapter:                                   # same call shape as VTNTrainStep; dp_inputs = the source mels (duration_predictor_feat: mel)
            engine = inner.engine

            def __call__(self, x, il, y, lab, ol):
                return inner(x, il, y, ol, x)

            def prefetch(self, x, y, lab):
                inner.prefetch(x, y, x)

            @property
            def replayed_launches(self):
                return inner.replayed_launches

        stepper = _Adapter()
    else:
        model = (TransformerTTS if tts else VTN)(**hp, compute_dtype="bf16" if bf16 else "float32", device=dev, seed=0)
        stepper = VTNTrainStep(model, lr=8e-5, warmup_steps=4000, use_graph=not args.no_graph)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # ---- device-resident inputs: `value`
    for _ in range(max(args.warmup, 3)):
        stepper(dxs, ilens, dys, dlabels, olens)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = _lib.launch_count() + stepper.replayed_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        losses = stepper(dxs, ilens, dys, dlabels, olens)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count() + stepper.replayed_launches - l0
    # ---- end to end through the public step API with pinned HOST buffers + loss read-back: `e2e`
    # Every step copies its inputs from pinned host memory and reads its losses back; the copy of step n + 1 is issued (prefetch, a
    # side stream) right after step n is launched, so it runs under step n's kernels -- what a pin_memory DataLoader gives the
    # reference trainer.  S2S_BENCH_NO_PREFETCH=1 puts the copy back on the compute stream.
    use_pf = os.environ.get("S2S_BENCH_NO_PREFETCH", "0") != "1"
    for _ in range(2):
        stepper(pxs, ilens, pys, plabels, olens).cpu()
    barrier()
    if use_pf:
        stepper.prefetch(pxs, pys, plabels)
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for _ in range(args.steps):
        out = stepper(pxs, ilens, pys, plabels, olens)
        if use_pf:
            stepper.prefetch(pxs, pys, plabels)                             # H2D copy of the next step's inputs
        host_losses = out.cpu()                                             # D2H read of this step's losses
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3)
    clocks = sampler.stop()
    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()
    if rank != 0:
        return
    assert all(map(lambda v: v == v, host_losses.tolist())), "NaN loss"
    frames = B * L * world * args.steps
    value = frames / (ms * 1e-3)
    e2e = frames / (ms_e2e * 1e-3)
    pk, pk_src = peaks()
    flops_step = 3.0 * (aasvc_fwd_flops(hp, T, L) if aas else vtn_fwd_flops(hp, T, L, tts)) * B
    line = {"metric": METRIC_AAS if aas else METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if bf16 else "f32", "data": "synthetic",
            "config": {"workload": desc, "global_batch": B * world, "per_gpu_batch": B, "parallelism": f"dp{world}",
                       "frames_per_sec_per_gpu": value / world, "cuda_graph": not args.no_graph,
                       "l2": "per-step working set (activations + 0.5 GB of parameter/optimizer state) exceeds the 126 MB L2; no flush needed",
                       "losses_last_step": host_losses.tolist(), "model_tflops_per_step": flops_step / 1e12,
                       "model_tflops_per_sec": flops_step / (ms / args.steps * 1e-3) / 1e12,
                       "tc_fallbacks": int(_lib.load().s2s_tc_fallback_count()),
                       "gemm_pdl": int(os.environ.get("S2S_GEMM_PDL", "0") or 0)},
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": (2 * xs.numel() * 4 + ys.numel() * 4 + 3 * B * 4) if aas else
                    (xs.numel() * xs.element_size() + (ys.numel() + labels.numel()) * 4 + 5 * B * 4),
                    "d2h_bytes_per_step": 16 if aas else 8,
                    "h2d_overlap": "next step's inputs prefetched on a side stream (step.prefetch)" if use_pf else "copies on the compute stream"},
            "gpu_launches": int(launches)}
    if world == 1:
        if bf16:
            run = None
            if aas:
                inner.engine.prepare(B, T, L, ilens, olens)
                run = lambda: inner._fwd_bwd(dxs, dys, dxs, True)
            gf, gms, n = gemm_roofline(stepper, (dxs, ilens, dys, dlabels, olens), dev, args.gemm_table, run)
            peak = pk.get("bf16_tflops_sustained", pk.get("bf16_tflops"))
            ach = gf / (gms * 1e-3) / 1e12
            traffic = None
            for fn in ("r01_gemm_tc_traffic.json", "r01_gemm_tc_traffic_c3.json", "r02_gemm_tc_traffic.json", "r02_gemm_tc_traffic_c3.json"):   # later files win     # ncu dram bytes per launch, per workload
                try:
                    tj = json.load(open(os.path.join(ROOT, "profiles", fn)))
                    if tj.get("workload") == args.workload:
                        traffic = tj["dram_bytes_per_launch"]
                except Exception:
                    pass
            line["roofline"] = {"bound": "tensor", "kernel": "gemm_tc_kernel (tcgen05.mma bf16, all shapes of one step)", "achieved": ach,
                                "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": traffic, "peak_source": pk_src + " (sustained)",
                                "launches_per_step": n, "algorithmic_gflop_per_launch_avg": gf / n / 1e9,
                                "avg_launch_us": gms * 1e3 / n, "gemm_share_of_step": gms / (ms / args.steps),
                                "how": "CUDA events around every mode-1 s2s_gemm launch of one fwd+bwd queued behind a GPU spin (no host gaps), right after the timed region"}
        if not args.no_cpu_baseline:
            mods = None if aas else reference_modules()
            if mods is not None:
                Bs = min(B, 4)
                sec = cpu_reference_steps(mods, args.workload, Bs, 2, 1)
                line["cpu_baseline"] = {"value": Bs * L / sec, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "reference",
                                        "sample": f"unmodified reference modules (baseline/_ref), fp32, {Bs} x ({T}->{L}) per step (of {B}; frames/s "
                                                  f"scales linearly in B), 2 steps after 1 warm-up, {torch.get_num_threads()} threads"}
            else:
                Bs = 1 if aas else min(B, 4)
                sec = cpu_port_steps_aas(hp, Bs, T, L, 1, 1) if aas else cpu_port_steps(hp, Bs, T, L, 2, 1, tts)
                line["cpu_baseline"] = {"value": Bs * L / sec, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                                        "sample": f"oracle port of the reference PyTorch-CPU path (fp32), {Bs} x ({T}->{L}) per step, 2 steps after 1 warm-up"}
        if not aas and not args.no_incumbent:
            # the practical incumbent on this box, timed after our own arm on the same GPU
            del stepper, model
            torch.cuda.empty_cache()
            line["dropin_boundary"] = dropin_boundary(args.workload)
            line["dropin_boundary_graph"] = dropin_boundary(args.workload, use_graph=True)
            line["torch_gpu_incumbent"] = incumbent_torch_gpu(args.workload)
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "torch_gpu"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-incumbent", action="store_true", help="skip the PyTorch-eager-on-this-GPU timing of the reference modules")
    ap.add_argument("--gemm-table", default=None, help="write a per-shape timing table of the tensor-core GEMM launches")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.impl == "torch_gpu":
        run_torch_gpu(args, rank)
        return
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"          # keep stdout to the single JSON line
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
        torch.distributed.init_process_group("nccl")
    try:
        run_ours(args, rank, world)
    finally:
        if world > 1:
            torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
