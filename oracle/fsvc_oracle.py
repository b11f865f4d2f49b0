"""CPU restatement of FastSpeechVC's teacher-forced training path (seq2seq_vc/models/fastspeech_vc.py:200-425, the conformer
configuration of egs/arctic/vc2/conf/fs2_vc.melmelmel.v1.yaml) and the loss assembly of NARVCTrainer._train_step
(trainers/nar_vc.py:53-99: L1Loss + DurationPredictorLoss).

TEST INFRASTRUCTURE ONLY.  Pinned against the live reference through tests/golden/fsvc_tiny.npz (oracle/gen_golden.py) and
tests/test_oracle_vs_reference.py.  Built from the AAS-VC oracle's pieces (the two models share every module but the length
regulation): Conv2dSubsampling + RelPositionalEncoding input layer, conformer encoder, Conv2dSubsampling projection + nearest
interpolation + DurationPredictor on the side input, LengthRegulator with the TEACHER's durations (length_regulator.py:69-97),
conformer decoder, feat_out, Postnet.
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

from oracle import aasvc_oracle as ao
from oracle.vtn_oracle import linear, non_pad_mask, postnet


def default_hparams(**over):
    hp = dict(idim=80, odim=80, adim=384, aheads=2, elayers=4, eunits=1536, dlayers=4, dunits=1536, duration_predictor_input_dim=80,
              duration_predictor_layers=2, duration_predictor_chans=256, duration_predictor_kernel_size=3, postnet_layers=5, postnet_filts=5,
              postnet_chans=256, conformer_enc_kernel_size=15, conformer_dec_kernel_size=15, teacher_model_decoder_reduction_factor=1)
    hp.update(over)
    return hp


def length_regulate(hs: torch.Tensor, ds: torch.Tensor, pad_value: float = 0.0) -> torch.Tensor:
    """LengthRegulator.forward (length_regulator.py:69-97), alpha = 1: row i of utterance b repeated ds[b, i] times, ragged rows padded."""
    rows = [torch.repeat_interleave(h, d.long(), dim=0) for h, d in zip(hs, ds)]
    out = hs.new_full((len(rows), max(r.shape[0] for r in rows), hs.shape[2]), pad_value)
    for b, r in enumerate(rows):
        out[b, : r.shape[0]] = r
    return out


def fsvc_forward(sd, hp, xs, ilens, ys, olens, ds, dp_inputs, training: bool = True, bn_stats=None):
    """FastSpeechVC.forward -> _forward (fastspeech_vc.py:200-305,382-425).  ds (B, T') int64 teacher durations, T' the encoder
    length after the conv2d input layer; ilens / olens python ints."""
    hp = default_hparams(**hp)
    ilens = [int(v) for v in ilens]
    olens = [int(v) for v in olens]
    H = hp["aheads"]
    xs = xs[:, : max(ilens)]
    ys = ys[:, : max(olens)]
    attn: Dict[str, torch.Tensor] = {}
    B = xs.shape[0]
    # encoder: Conv2dSubsampling + RelPositionalEncoding (conformer/encoder.py:124-131) + conformer blocks; the key mask is the
    # subsampled source mask (subsampling.py:92-94)
    x_mask = non_pad_mask(ilens, xs.shape[1]).unsqueeze(-2)
    x = xs.unsqueeze(1)
    x = torch.relu(F.conv2d(x, sd["encoder.embed.conv.0.weight"], sd["encoder.embed.conv.0.bias"], stride=2))
    x = torch.relu(F.conv2d(x, sd["encoder.embed.conv.2.weight"], sd["encoder.embed.conv.2.bias"], stride=2))
    b, c, t, f = x.shape
    x = linear(x.transpose(1, 2).reshape(b, t, c * f), sd, "encoder.embed.out.0")
    mask = x_mask[:, :, :-2:2][:, :, :-2:2]
    hs = ao.conformer_layers(sd, "encoder", hp["elayers"], H, x, mask, training, bn_stats, attn)
    tlens = [((i - 2 + 1) // 2 - 2 + 1) // 2 for i in ilens]                   # fastspeech_vc.py:236-238
    # duration predictor on the projected side input (fastspeech_vc.py:244-275)
    dpi = ao.dp_projection(sd, "duration_predictor_projection", dp_inputs, hs.shape[1])
    d_outs = ao.duration_predictor(sd, "duration_predictor", hp, dpi, tlens, clamp=False)
    # length regulator with the teacher's durations (:276-279), conformer decoder on the result (:281-305)
    up = length_regulate(hs, ds * hp["teacher_model_decoder_reduction_factor"])
    L = up.shape[1]
    h_mask = non_pad_mask(olens, L).unsqueeze(-2)
    zs = ao.conformer_layers(sd, "decoder", hp["dlayers"], H, up, h_mask, training, bn_stats, attn)
    before = linear(zs, sd, "feat_out").view(B, -1, hp["odim"])
    after = before + postnet(sd, hp, before.transpose(1, 2), training, bn_stats).transpose(1, 2)
    return dict(before_outs=before, after_outs=after, d_outs=d_outs, ilens=tlens, olens=olens, ys=ys, attn=attn)


def fsvc_losses(out, ds):
    """trainers/nar_vc.py:73-82: L1Loss(after, before, ys, olens) + DurationPredictorLoss(d_outs, durations, ilens)."""
    l1 = ao.l1_loss(out["after_outs"], out["before_outs"], out["ys"], out["olens"])
    dur = ao.duration_loss(out["d_outs"], ds.to(torch.float32), out["ilens"])
    return l1 + dur, dict(l1_loss=l1, duration_loss=dur)


def fsvc_loss_and_grads(sd, hp, xs, ilens, ys, olens, ds, dp_inputs, training=True):
    params = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items() if v.dtype.is_floating_point and "running_" not in k}
    full = dict(sd)
    full.update(params)
    out = fsvc_forward(full, hp, xs, ilens, ys, olens, ds, dp_inputs, training=training)
    total, parts = fsvc_losses(out, ds)
    grads = torch.autograd.grad(total, list(params.values()), allow_unused=True)
    return out, parts, {k: g for k, g in zip(params.keys(), grads)}
