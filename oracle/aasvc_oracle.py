"""CPU oracle for the AAS-VC (Conformer, non-autoregressive) training path -- TEST INFRASTRUCTURE ONLY.

Functional fp32 restatement (plain PyTorch on CPU, autograd for gradients) of
  seq2seq_vc/models/aas_vc.py:279-471            AASVC._forward (teacher-forced branch)
  seq2seq_vc/modules/conformer/encoder.py:249-293, encoder_layer.py:79-179, convolution.py:56-79
  seq2seq_vc/modules/transformer/attention.py:209-305   RelPositionMultiHeadedAttention (+ rel_shift)
  seq2seq_vc/layers/positional_encoding.py:238-309      RelPositionalEncoding
  seq2seq_vc/modules/alignments.py:12-60,281-310        AlignmentModule, viterbi_decode
  seq2seq_vc/modules/length_regulator.py:100-154        GaussianUpsampling
  seq2seq_vc/modules/duration_predictor.py:27-128       DurationPredictor (deterministic)
  seq2seq_vc/losses/{l1_loss,forward_sum_loss,duration_predictor_loss}.py
  seq2seq_vc/trainers/aas_vc.py:56-134                  loss assembly (lambda_align)
for the configuration family of egs/arctic/vc2/conf/aas_vc.melmelmel.v1.yaml with the README-sanctioned
deterministic duration predictor: encoder/decoder reduction factor 1, `linear` input layer, macaron +
CNN conformer blocks, rel_pos / rel_selfattn, pre-LN, Conv2dSubsampling projection of the duration
predictor input.  Pinned against the live reference (tests/test_oracle_vs_reference.py) and against
golden vectors dumped from it (tests/golden/aasvc_tiny.npz, oracle/gen_golden.py).

Third-party arithmetic on this path: scipy.stats.betabinom.logpmf (forward_sum_loss.py:9,107; scipy is
in the image, used here exactly as the reference's call site does) and torch's F.ctc_loss, which this
file restates as an explicit log-domain alpha recursion (forward_sum_loss.py:58-76).
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence

import numpy as np
import torch
import torch.nn.functional as F

from oracle import mas_oracle
from oracle.vtn_oracle import BN_EPS, BN_MOMENTUM, LN_EPS, linear, non_pad_mask, postnet

EMBED_LN_EPS = 1e-5     # torch.nn.LayerNorm default in the `linear` input layer (conformer/encoder.py:119)
MAX_DP_OUTPUT = 10      # models/aas_vc.py:35
LOG_BLANK = -1.0        # log(e**-1), forward_sum_loss.py:31,56
GAUSS_DELTA = 0.1       # length_regulator.py:107


def default_hparams(**over):
    """model_params of egs/arctic/vc2/conf/aas_vc.melmelmel.v1.yaml (deterministic duration predictor)."""
    hp = dict(idim=80, odim=80, adim=384, aheads=2, elayers=4, eunits=1536, dlayers=4, dunits=1536,
              duration_predictor_input_dim=80, duration_predictor_layers=2, duration_predictor_chans=256,
              duration_predictor_kernel_size=3, postnet_layers=5, postnet_filts=5, postnet_chans=256,
              post_encoder_reduction_factor=4, conformer_enc_kernel_size=15, conformer_dec_kernel_size=15)
    hp.update(over)
    return hp


def layer_norm(x, sd, prefix, eps=LN_EPS):
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + ".weight"], sd[prefix + ".bias"], eps)


def rel_pos_table(T: int, d: int) -> torch.Tensor:
    """pos_emb (2T-1, d) of RelPositionalEncoding.forward (positional_encoding.py:263-309): row k is the
    sinusoid at relative position T-1-k (positive positions first, reversed; then the negative ones)."""
    pos = torch.arange(T - 1, -T, -1, dtype=torch.float32).unsqueeze(1)
    div = torch.exp(torch.arange(0, d, 2, dtype=torch.float32) * -(math.log(10000.0) / d))
    pe = torch.zeros(2 * T - 1, d)
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe


def swish(x):
    return x * torch.sigmoid(x)


def rel_attention(sd, prefix, x, pos_emb, mask, n_head, store=None):
    """RelPositionMultiHeadedAttention.forward (attention.py:262-305); rel_shift (:237-260) is restated as the
    gather bd'[i, j] = bd[i, T-1-i+j]."""
    B, T, d = x.shape
    dk = d // n_head
    q = linear(x, sd, prefix + ".linear_q").view(B, T, n_head, dk)
    k = linear(x, sd, prefix + ".linear_k").view(B, T, n_head, dk).transpose(1, 2)
    v = linear(x, sd, prefix + ".linear_v").view(B, T, n_head, dk).transpose(1, 2)
    p = F.linear(pos_emb, sd[prefix + ".linear_pos.weight"]).view(2 * T - 1, n_head, dk).transpose(0, 1)   # (H, 2T-1, dk)
    qu = (q + sd[prefix + ".pos_bias_u"]).transpose(1, 2)
    qv = (q + sd[prefix + ".pos_bias_v"]).transpose(1, 2)
    ac = torch.matmul(qu, k.transpose(-2, -1))
    bd = torch.matmul(qv, p.transpose(-2, -1).unsqueeze(0))                      # (B, H, T, 2T-1)
    idx = (T - 1 - torch.arange(T)[:, None] + torch.arange(T)[None, :])          # (T, T)
    bd = torch.gather(bd, 3, idx[None, None].expand(B, n_head, T, T))
    scores = (ac + bd) / math.sqrt(dk)
    dead = ~mask.unsqueeze(1)
    scores = scores.masked_fill(dead, torch.finfo(scores.dtype).min)
    pr = torch.softmax(scores, dim=-1).masked_fill(dead, 0.0)
    if store is not None:
        store[prefix] = pr
    ctx = torch.matmul(pr, v).transpose(1, 2).reshape(B, T, d)
    return linear(ctx, sd, prefix + ".linear_out")


def ffn_swish(sd, prefix, x):
    """Position-wise layer of a conformer block (conformer/encoder.py:181-198), told apart by the weight rank:
    "linear"  -> PositionwiseFeedForward with Swish (positionwise_feed_forward.py:12-32), weights (U, d);
    "conv1d"  -> MultiLayeredConv1d (multi_layer_conv.py:13-62): Conv1d(k) -> ReLU -> dropout -> Conv1d(k) over time with
                 padding (k-1)//2, weights (U, d, k); the activation is ReLU whatever the encoder's activation_type."""
    w1 = sd[prefix + ".w_1.weight"]
    if w1.dim() == 3:
        k = w1.shape[2]
        h = F.relu(F.conv1d(x.transpose(1, 2), w1, sd[prefix + ".w_1.bias"], padding=(k - 1) // 2))
        if sd[prefix + ".w_2.weight"].dim() == 2:       # "conv1d-linear": Conv1dLinear (multi_layer_conv.py:66-108), w_2 is a Linear
            return linear(h.transpose(1, 2), sd, prefix + ".w_2")
        return F.conv1d(h, sd[prefix + ".w_2.weight"], sd[prefix + ".w_2.bias"], padding=(k - 1) // 2).transpose(1, 2)
    return linear(swish(linear(x, sd, prefix + ".w_1")), sd, prefix + ".w_2")


def conv_module(sd, prefix, x, training, bn_stats=None):
    """ConvolutionModule.forward (conformer/convolution.py:56-79); BatchNorm sees padded frames (no mask)."""
    d = x.shape[-1]
    x = x.transpose(1, 2)
    x = F.conv1d(x, sd[prefix + ".pointwise_conv1.weight"], sd[prefix + ".pointwise_conv1.bias"])
    x = F.glu(x, dim=1)
    w = sd[prefix + ".depthwise_conv.weight"]
    x = F.conv1d(x, w, sd[prefix + ".depthwise_conv.bias"], padding=(w.shape[-1] - 1) // 2, groups=d)
    if training:
        mean = x.mean(dim=(0, 2))
        var_b = x.var(dim=(0, 2), unbiased=False)
        if bn_stats is not None:
            n = x.shape[0] * x.shape[2]
            bn_stats[prefix + ".norm.running_mean"] = (1 - BN_MOMENTUM) * sd[prefix + ".norm.running_mean"] + BN_MOMENTUM * mean.detach()
            bn_stats[prefix + ".norm.running_var"] = ((1 - BN_MOMENTUM) * sd[prefix + ".norm.running_var"]
                                                      + BN_MOMENTUM * (var_b.detach() * n / (n - 1)))
    else:
        mean, var_b = sd[prefix + ".norm.running_mean"], sd[prefix + ".norm.running_var"]
    x = (x - mean[None, :, None]) * torch.rsqrt(var_b[None, :, None] + BN_EPS)
    x = x * sd[prefix + ".norm.weight"][None, :, None] + sd[prefix + ".norm.bias"][None, :, None]
    x = swish(x)
    x = F.conv1d(x, sd[prefix + ".pointwise_conv2.weight"], sd[prefix + ".pointwise_conv2.bias"])
    return x.transpose(1, 2)


def conformer_layers(sd, prefix, n_layers, n_head, x, mask, training, bn_stats=None, attn_store=None):
    """RelPositionalEncoding + n x conformer EncoderLayer (pre-LN, macaron, CNN) + after_norm
    (conformer/encoder.py:249-293, encoder_layer.py:79-179)."""
    T, d = x.shape[1], x.shape[2]
    x = x * math.sqrt(d)
    pos_emb = rel_pos_table(T, d)
    for l in range(n_layers):
        p = f"{prefix}.encoders.{l}"
        x = x + 0.5 * ffn_swish(sd, p + ".feed_forward_macaron", layer_norm(x, sd, p + ".norm_ff_macaron"))
        x = x + rel_attention(sd, p + ".self_attn", layer_norm(x, sd, p + ".norm_mha"), pos_emb, mask, n_head, attn_store)
        x = x + conv_module(sd, p + ".conv_module", layer_norm(x, sd, p + ".norm_conv"), training, bn_stats)
        x = x + 0.5 * ffn_swish(sd, p + ".feed_forward", layer_norm(x, sd, p + ".norm_ff"))
        x = layer_norm(x, sd, p + ".norm_final")
    return layer_norm(x, sd, prefix + ".after_norm")


def dp_projection(sd, prefix, xs, T_out):
    """Conv2dSubsampling(use_pos_enc=False) + per-utterance nearest F.interpolate to the padded encoder
    length (subsampling.py:74-94, aas_vc.py:335-351).  Nearest: src = floor(dst * T_in / T_out)."""
    x = xs.unsqueeze(1)
    x = torch.relu(F.conv2d(x, sd[prefix + ".conv.0.weight"], sd[prefix + ".conv.0.bias"], stride=2))
    x = torch.relu(F.conv2d(x, sd[prefix + ".conv.2.weight"], sd[prefix + ".conv.2.bias"], stride=2))
    b, c, t, f = x.shape
    x = linear(x.transpose(1, 2).reshape(b, t, c * f), sd, prefix + ".out")
    idx = interp_index(t, T_out)
    return x[:, idx]


def interp_index(T_in: int, T_out: int) -> torch.Tensor:
    """Source row of every output row for F.interpolate(mode="nearest"): floor(dst * (T_in / T_out)) with the
    scale held in float32, as torch's nearest kernel computes it."""
    scale = np.float32(T_in) / np.float32(T_out)
    idx = np.floor(np.arange(T_out, dtype=np.float32) * scale).astype(np.int64)
    return torch.from_numpy(np.minimum(idx, T_in - 1))


def alignment_log_p(sd, prefix, text, feats, text_lens):
    """AlignmentModule.forward (alignments.py:28-60): -L2 distance, text padding -> -inf, log-softmax."""
    t = text.transpose(1, 2)
    t = torch.relu(F.conv1d(t, sd[prefix + ".t_conv1.weight"], sd[prefix + ".t_conv1.bias"], padding=1))
    t = F.conv1d(t, sd[prefix + ".t_conv2.weight"], sd[prefix + ".t_conv2.bias"]).transpose(1, 2)
    f = feats.transpose(1, 2)
    f = torch.relu(F.conv1d(f, sd[prefix + ".f_conv1.weight"], sd[prefix + ".f_conv1.bias"], padding=1))
    f = torch.relu(F.conv1d(f, sd[prefix + ".f_conv2.weight"], sd[prefix + ".f_conv2.bias"], padding=1))
    f = F.conv1d(f, sd[prefix + ".f_conv3.weight"], sd[prefix + ".f_conv3.bias"]).transpose(1, 2)
    dist = torch.norm(f.unsqueeze(2) - t.unsqueeze(1), p=2, dim=3)
    score = -dist
    pad = ~non_pad_mask(text_lens, text.shape[1])
    score = score.masked_fill(pad.unsqueeze(-2), -np.inf)
    return F.log_softmax(score, dim=-1)


def gaussian_upsampling(hs, ds, feats_lens, text_lens, T_feats):
    """GaussianUpsampling.forward (length_regulator.py:111-154).  Padded output frames use t = 0."""
    t = torch.arange(T_feats, dtype=torch.float32)[None].repeat(ds.shape[0], 1)
    t = t * non_pad_mask(feats_lens, T_feats).float()
    c = ds.cumsum(dim=-1) - ds / 2
    energy = -GAUSS_DELTA * (t.unsqueeze(-1) - c.unsqueeze(1)) ** 2
    dmask = non_pad_mask(text_lens, ds.shape[1])
    energy = energy.masked_fill(~dmask.unsqueeze(1), -float("inf"))
    return torch.matmul(torch.softmax(energy, dim=2), hs)


def duration_predictor(sd, prefix, hp, xs, text_lens, clamp: bool = True):
    """DurationPredictor.forward (duration_predictor.py:83-114) + clamp(max=10) (aas_vc.py:408-410; AAS-VC only: FastSpeechVC uses
    the predictor's output as it is, fastspeech_vc.py:270-275)."""
    k = hp["duration_predictor_kernel_size"]
    x = xs.transpose(1, 2)
    for i in range(hp["duration_predictor_layers"]):
        p = f"{prefix}.conv.{i}"
        x = torch.relu(F.conv1d(x, sd[p + ".0.weight"], sd[p + ".0.bias"], padding=(k - 1) // 2))
        x = layer_norm(x.transpose(1, 2), sd, p + ".2").transpose(1, 2)
    out = linear(x.transpose(1, 2), sd, prefix + ".linear").squeeze(-1)
    out = out * non_pad_mask(text_lens, xs.shape[1])
    return torch.clamp(out, max=MAX_DP_OUTPUT) if clamp else out


def beta_binomial_prior(N: int, T: int) -> np.ndarray:
    """(T, N) float64 log-prior, forward_sum_loss.py:100-114 (scipy betabinom at the reference's call site)."""
    from scipy.stats import betabinom

    alpha = np.arange(1, T + 1, dtype=float)
    beta = np.array([T - t + 1 for t in alpha])
    k = np.arange(N)[..., None]
    return betabinom.logpmf(k, N, alpha, beta).T


def ctc_forward_sum_utt(lp: np.ndarray):
    """One utterance of the reference's F.ctc_loss call (forward_sum_loss.py:58-74), float64.

    lp (T, N): label log-probabilities (log_p_attn + prior); states [blank, 1, blank, ..., N, blank] with a
    constant blank log-probability LOG_BLANK.  Returns (nll, grad (T, N)) where grad is what torch's
    ctc_loss backward hands to log_probs for a unit upstream gradient:
        grad[t, k] = exp(lp[t, k]) - exp(alpha_t(2k+1) + beta_t(2k+1) - lp[t, k] + nll)
    (aten/src/ATen/native/LossCTC.cpp, "eq. (16)": the derivative w.r.t. *logits* under an implicit
    log-softmax).  The reference feeds un-normalised rows (prior added, constant blank), so the exp(lp)
    term does not cancel downstream: it is part of the reference's training signal and is reproduced
    here on purpose.  An infeasible utterance (nll = inf) gives (0, zeros): zero_infinity=True.
    """
    T, N = lp.shape
    S = 2 * N + 1
    NEG = -np.inf
    emit = np.full((T, S), LOG_BLANK, dtype=np.float64)
    emit[:, 1::2] = lp
    skip = np.zeros(S, dtype=bool)
    skip[3::2] = True                                  # label states may skip the blank between distinct labels

    def lse3(a, b, c):
        m = np.maximum(np.maximum(a, b), c)
        ms = np.where(np.isfinite(m), m, 0.0)
        with np.errstate(divide="ignore"):
            return np.log(np.exp(a - ms) + np.exp(b - ms) + np.exp(c - ms)) + ms

    alpha = np.full((T, S), NEG)
    alpha[0, :2] = emit[0, :2]
    for t in range(1, T):
        a0 = alpha[t - 1]
        a1 = np.concatenate([[NEG], a0[:-1]])
        a2 = np.where(skip, np.concatenate([[NEG, NEG], a0[:-2]]), NEG)
        alpha[t] = lse3(a0, a1, a2) + emit[t]
    ll = lse3(alpha[T - 1, S - 1], alpha[T - 1, S - 2] if S > 1 else NEG, NEG)
    if not np.isfinite(ll):
        return 0.0, np.zeros((T, N))
    beta = np.full((T, S), NEG)
    beta[T - 1, S - 1] = emit[T - 1, S - 1]
    if S > 1:
        beta[T - 1, S - 2] = emit[T - 1, S - 2]
    skip_from = np.zeros(S, dtype=bool)
    skip_from[1:S - 2:2] = True                        # from label state s to label state s + 2
    for t in range(T - 2, -1, -1):
        b0 = beta[t + 1]
        b1 = np.concatenate([b0[1:], [NEG]])
        b2 = np.where(skip_from, np.concatenate([b0[2:], [NEG, NEG]]), NEG)
        beta[t] = lse3(b0, b1, b2) + emit[t]
    nll = -ll
    with np.errstate(over="ignore"):
        grad = np.exp(lp) - np.exp(alpha[:, 1::2] + beta[:, 1::2] - lp + nll)
    grad = np.where(np.isfinite(lp), grad, 0.0)
    return float(nll), grad


class _ForwardSum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, log_p_attn, text_lens, feats_lens):
        B = log_p_attn.shape[0]
        grad = np.zeros(tuple(log_p_attn.shape), dtype=np.float64)
        total = 0.0
        lpn = log_p_attn.detach().double().numpy()
        for b in range(B):
            N, T = int(text_lens[b]), int(feats_lens[b])
            prior = beta_binomial_prior(N, T).astype(np.float32).astype(np.float64)   # cast to fp32 before the add (:48-49)
            nll, g = ctc_forward_sum_utt(lpn[b, :T, :N] + prior)
            total += nll / N                                                          # reduction="mean": / target length
            grad[b, :T, :N] = g / N
        ctx.save_for_backward(torch.from_numpy(grad / B).to(log_p_attn.dtype))
        return log_p_attn.new_tensor(total / B)

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return g * grad, None, None


def forward_sum_loss(log_p_attn, text_lens, feats_lens):
    """ForwardSumLoss.forward (forward_sum_loss.py:26-76): mean_b [CTC-NLL_b / N_b]; the gradient is the one
    torch's ctc_loss backward produces for the reference (see ctc_forward_sum_utt)."""
    return _ForwardSum.apply(log_p_attn, list(text_lens), list(feats_lens))


def duration_loss(d_outs, ds, text_lens):
    """DurationPredictorLoss.forward (duration_predictor_loss.py:29-50)."""
    m = non_pad_mask(text_lens, ds.shape[1])
    return F.mse_loss(d_outs.masked_select(m), torch.log(ds.masked_select(m).float() + 1.0))


def l1_loss(after, before, ys, olens):
    """L1Loss.forward (l1_loss.py:24-49)."""
    m = non_pad_mask(olens, ys.shape[1]).unsqueeze(-1)
    y = ys.masked_select(m)
    return (before.masked_select(m) - y).abs().mean() + (after.masked_select(m) - y).abs().mean()


def aasvc_forward(sd, hp, xs, ilens, ys, olens, dp_inputs, training: bool = True, bn_stats=None):
    """AASVC.forward (aas_vc.py:473-529 -> _forward :279-471), teacher-forced.  ilens / olens: python ints."""
    hp = default_hparams(**hp)
    ilens = [int(v) for v in ilens]
    olens = [int(v) for v in olens]
    pr, H = hp["post_encoder_reduction_factor"], hp["aheads"]
    xs = xs[:, : max(ilens)]
    ys = ys[:, : max(olens)]
    attn: Dict[str, torch.Tensor] = {}
    B, T = xs.shape[0], xs.shape[1]

    # encoder (aas_vc.py:307-308; conformer/encoder.py:117-123 `linear` input layer)
    x_mask = non_pad_mask(ilens, T).unsqueeze(-2)
    x = layer_norm(linear(xs, sd, "encoder.embed.0"), sd, "encoder.embed.1", EMBED_LN_EPS)
    hs = conformer_layers(sd, "encoder", hp["elayers"], H, x, x_mask, training, bn_stats, attn)

    # post-encoder reduction (aas_vc.py:319-332)
    Tt = T // pr
    hs = hs[:, : Tt * pr].reshape(B, Tt, hs.shape[2] * pr)
    tlens = [i // pr for i in ilens]

    # duration-predictor input (aas_vc.py:335-351)
    dpi = dp_projection(sd, "duration_predictor_projection", dp_inputs, Tt)

    # alignment + MAS (aas_vc.py:401-404; alignments.py:281-310)
    log_p = alignment_log_p(sd, "alignment_module", hs, ys, tlens)
    ds_np, _, paths = mas_oracle.viterbi_decode_oracle(log_p.detach().numpy(), tlens, olens)
    ds = torch.from_numpy(ds_np)
    bin_loss = log_p.new_zeros(())
    for b in range(B):
        t_idx = torch.arange(olens[b])
        bin_loss = bin_loss - log_p[b, t_idx, torch.from_numpy(paths[b, : olens[b]].astype(np.int64))].mean()
    bin_loss = bin_loss / B

    d_outs = duration_predictor(sd, "duration_predictor", hp, dpi, tlens)         # aas_vc.py:407-411
    L = ys.shape[1]
    up = gaussian_upsampling(hs, ds, olens, tlens, L)                             # aas_vc.py:422-427

    h_mask = non_pad_mask(olens, L).unsqueeze(-2)
    zs = conformer_layers(sd, "decoder", hp["dlayers"], H, up, h_mask, training, bn_stats, attn)
    before = linear(zs, sd, "feat_out").view(B, -1, hp["odim"])
    after = before + postnet(sd, hp, before.transpose(1, 2), training, bn_stats).transpose(1, 2)
    return dict(before_outs=before, after_outs=after, ds=ds, ilens=tlens, olens=olens, olens_reduced=olens, ys=ys,
                bin_loss=bin_loss, log_p_attn=log_p, d_outs=d_outs, attn=attn, paths=paths)


def aasvc_losses(out, lambda_align: float = 2.0):
    """Loss assembly of AASVCTrainer._train_step (trainers/aas_vc.py:73-134), duration loss enabled."""
    l1 = l1_loss(out["after_outs"], out["before_outs"], out["ys"], out["olens"])
    fs = forward_sum_loss(out["log_p_attn"], out["ilens"], out["olens_reduced"])
    dur = duration_loss(out["d_outs"], out["ds"], out["ilens"])
    total = l1 + lambda_align * (fs + out["bin_loss"]) + dur
    return total, dict(l1_loss=l1, forward_sum_loss=fs, bin_loss=out["bin_loss"], duration_loss=dur)


def aasvc_loss_and_grads(sd, hp, xs, ilens, ys, olens, dp_inputs, training=True, lambda_align: float = 2.0):
    params = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()
              if v.dtype.is_floating_point and "running_" not in k}
    full = dict(sd)
    full.update(params)
    out = aasvc_forward(full, hp, xs, ilens, ys, olens, dp_inputs, training=training)
    total, parts = aasvc_losses(out, lambda_align)
    grads = torch.autograd.grad(total, list(params.values()), allow_unused=True)
    return out, parts, {k: g for k, g in zip(params.keys(), grads)}


def state_dict_spec(hp) -> List:
    """(name, shape) of every AASVC state-dict entry for this configuration family."""
    hp = default_hparams(**hp)
    d, H, pr = hp["adim"], hp["aheads"], hp["post_encoder_reduction_factor"]
    idim, odim = hp["idim"], hp["odim"]
    spec = []

    def lin(n, o, i, bias=True):
        spec.append((n + ".weight", (o, i)))
        if bias:
            spec.append((n + ".bias", (o,)))

    def ln(n, c):
        spec.extend([(n + ".weight", (c,)), (n + ".bias", (c,))])

    def bn(n, c):
        ln(n, c)
        spec.extend([(n + ".running_mean", (c,)), (n + ".running_var", (c,)), (n + ".num_batches_tracked", ())])

    def conformer(prefix, n_layers, dm, units, k):
        for l in range(n_layers):
            p = f"{prefix}.encoders.{l}"
            spec.extend([(p + ".self_attn.pos_bias_u", (H, dm // H)), (p + ".self_attn.pos_bias_v", (H, dm // H))])
            for s in ("linear_q", "linear_k", "linear_v", "linear_out"):
                lin(f"{p}.self_attn.{s}", dm, dm)
            lin(p + ".self_attn.linear_pos", dm, dm, bias=False)
            for ff in ("feed_forward", "feed_forward_macaron"):
                if hp.get("positionwise_layer_type", "linear") in ("conv1d", "conv1d-linear"):
                    pk_ = hp.get("positionwise_conv_kernel_size", 1)
                    w2 = (dm, units, pk_) if hp["positionwise_layer_type"] == "conv1d" else (dm, units)
                    spec.extend([(f"{p}.{ff}.w_1.weight", (units, dm, pk_)), (f"{p}.{ff}.w_1.bias", (units,)),
                                 (f"{p}.{ff}.w_2.weight", w2), (f"{p}.{ff}.w_2.bias", (dm,))])
                else:
                    lin(f"{p}.{ff}.w_1", units, dm)
                    lin(f"{p}.{ff}.w_2", dm, units)
            spec.extend([(p + ".conv_module.pointwise_conv1.weight", (2 * dm, dm, 1)), (p + ".conv_module.pointwise_conv1.bias", (2 * dm,)),
                         (p + ".conv_module.depthwise_conv.weight", (dm, 1, k)), (p + ".conv_module.depthwise_conv.bias", (dm,))])
            bn(p + ".conv_module.norm", dm)
            spec.extend([(p + ".conv_module.pointwise_conv2.weight", (dm, dm, 1)), (p + ".conv_module.pointwise_conv2.bias", (dm,))])
            for n in ("norm_ff", "norm_mha", "norm_ff_macaron", "norm_conv", "norm_final"):
                ln(f"{p}.{n}", dm)
        ln(prefix + ".after_norm", dm)

    lin("encoder.embed.0", d, idim)
    ln("encoder.embed.1", d)
    conformer("encoder", hp["elayers"], d, hp["eunits"], hp["conformer_enc_kernel_size"])
    ch, k = hp["duration_predictor_chans"], hp["duration_predictor_kernel_size"]
    for i in range(hp["duration_predictor_layers"]):
        spec.extend([(f"duration_predictor.conv.{i}.0.weight", (ch, d if i == 0 else ch, k)), (f"duration_predictor.conv.{i}.0.bias", (ch,))])
        ln(f"duration_predictor.conv.{i}.2", ch)
    lin("duration_predictor.linear", 1, ch)
    f2 = ((hp["duration_predictor_input_dim"] - 1) // 2 - 1) // 2
    spec.extend([("duration_predictor_projection.conv.0.weight", (d, 1, 3, 3)), ("duration_predictor_projection.conv.0.bias", (d,)),
                 ("duration_predictor_projection.conv.2.weight", (d, d, 3, 3)), ("duration_predictor_projection.conv.2.bias", (d,))])
    lin("duration_predictor_projection.out", d, d * f2)
    C = d * pr
    for n, ic, kk in (("t_conv1", C, 3), ("t_conv2", C, 1), ("f_conv1", odim, 3), ("f_conv2", C, 3), ("f_conv3", C, 1)):
        spec.extend([(f"alignment_module.{n}.weight", (C, ic, kk)), (f"alignment_module.{n}.bias", (C,))])
    conformer("decoder", hp["dlayers"], C, hp["dunits"], hp["conformer_dec_kernel_size"])
    lin("feat_out", odim, C)
    pc, pk = hp["postnet_chans"], hp["postnet_filts"]
    for i in range(hp["postnet_layers"]):
        ic = odim if i == 0 else pc
        oc = odim if i == hp["postnet_layers"] - 1 else pc
        spec.append((f"postnet.postnet.{i}.0.weight", (oc, ic, pk)))
        bn(f"postnet.postnet.{i}.1", oc)
    return spec


def init_state_dict(hp, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Random AASVC state dict with the reference's key names / shapes (uniform +-1/sqrt(fan_in); affine 1/0)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in state_dict_spec(hp):
        if name.endswith("num_batches_tracked"):
            sd[name] = torch.zeros((), dtype=torch.int64)
        elif name.endswith("running_mean"):
            sd[name] = torch.zeros(shape)
        elif name.endswith("running_var"):
            sd[name] = torch.ones(shape)
        elif "norm" in name or name.startswith("encoder.embed.1") or ".conv." in name and name.split(".")[-2] == "2" \
                or (name.startswith("postnet") and ".1." in name):
            sd[name] = torch.ones(shape) if name.endswith("weight") else torch.zeros(shape)
        else:
            fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else int(shape[0])
            if name.endswith(".bias"):
                w = sd.get(name[:-5] + ".weight")
                fan_in = int(np.prod(w.shape[1:])) if w is not None else fan_in
            sd[name] = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(max(fan_in, 1))
    return sd


def synthetic_batch(B, T, L, idim=80, odim=80, ilens=None, olens=None, seed=1234):
    """Seeded N(0,1) batch: xs (B,T,idim), ys (B,L,odim), dp_inputs = xs; padding zeroed."""
    g = torch.Generator().manual_seed(seed)
    ilens = list(ilens) if ilens is not None else [T] * B
    olens = list(olens) if olens is not None else [L] * B
    xs = torch.randn(B, T, idim, generator=g)
    ys = torch.randn(B, L, odim, generator=g)
    for b in range(B):
        xs[b, ilens[b]:] = 0
        ys[b, olens[b]:] = 0
    return xs, ilens, ys, olens, xs.clone()


def aasvc_inference(sd, hp, x, dp_input):
    """AASVC.inference without ground truth (aas_vc.py:531-603 -> _forward(is_inference=True) :371-398): eval-mode encoder,
    durations = clamp(round(exp(d) - 1), 0, 10) (duration_predictor.py:92-96, aas_vc.py:389), Gaussian upsampling over
    T_feats = sum(ds) frames without masks, decoder without attention mask, postnet with running BatchNorm statistics.
    x (T, idim), dp_input (T_dp, dp_idim) -> (outs (L, odim), d_outs (T_text,) int64)."""
    hp = default_hparams(**hp)
    pr, H = hp["post_encoder_reduction_factor"], hp["aheads"]
    xs, dpi = x.unsqueeze(0), dp_input.unsqueeze(0)
    T = xs.shape[1]
    x_mask = non_pad_mask([T], T).unsqueeze(-2)
    e = layer_norm(linear(xs, sd, "encoder.embed.0"), sd, "encoder.embed.1", EMBED_LN_EPS)
    hs = conformer_layers(sd, "encoder", hp["elayers"], H, e, x_mask, False)
    Tt = T // pr
    hs = hs[:, : Tt * pr].reshape(1, Tt, hs.shape[2] * pr)
    dp_in = dp_projection(sd, "duration_predictor_projection", dpi, Tt)
    k = hp["duration_predictor_kernel_size"]
    z = dp_in.transpose(1, 2)
    for i in range(hp["duration_predictor_layers"]):
        p = f"duration_predictor.conv.{i}"
        z = torch.relu(F.conv1d(z, sd[p + ".0.weight"], sd[p + ".0.bias"], padding=(k - 1) // 2))
        z = layer_norm(z.transpose(1, 2), sd, p + ".2").transpose(1, 2)
    pre = linear(z.transpose(1, 2), sd, "duration_predictor.linear").squeeze(-1)
    d_outs = torch.clamp(torch.clamp(torch.round(pre.exp() - 1.0), min=0).long(), max=MAX_DP_OUTPUT)
    ds = d_outs.clone()
    if ds.sum() == 0:
        ds[ds.sum(dim=1).eq(0)] = 1                                              # length_regulator.py:127-135
    L = int(ds.sum())
    t = torch.arange(L, dtype=torch.float32)[None]
    c = ds.cumsum(dim=-1) - ds / 2
    energy = -GAUSS_DELTA * (t.unsqueeze(-1) - c.unsqueeze(1)) ** 2
    up = torch.matmul(torch.softmax(energy, dim=2), hs)
    full = torch.ones(1, 1, L, dtype=torch.bool)
    zs = conformer_layers(sd, "decoder", hp["dlayers"], H, up, full, False)
    before = linear(zs, sd, "feat_out").view(1, -1, hp["odim"])
    after = before + postnet(sd, hp, before.transpose(1, 2), False).transpose(1, 2)
    return after[0], d_outs[0]
