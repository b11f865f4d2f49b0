"""Functional fp32 CPU restatement of the reference VTN training forward + Seq2SeqLoss.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Parity pinned: checked against the live
reference (imported through oracle/ref_shim.py in the build container) by
tests/test_oracle_vs_reference.py and against committed golden vectors tests/golden/vtn_tiny.npz.

The model is expressed as pure functions over a flat ``state_dict`` (name -> tensor) so that it
shares no code structure with the reference's nn.Module tree; every function cites the
reference file:line whose arithmetic it restates (paths relative to /root/reference).
All dropout is treated as identity (p = 0): parity is defined on the deterministic path
(SURVEY.md §8c), and the reference's always-on Prenet dropout must be disabled
(dprenet_dropout_rate=0) on the reference side when comparing.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence

import torch
import torch.nn.functional as F

LN_EPS = 1e-12  # seq2seq_vc/modules/transformer/layer_norm.py:23
BN_EPS = 1e-5   # torch.nn.BatchNorm1d default, seq2seq_vc/modules/pre_postnets.py:122
BN_MOMENTUM = 0.1


def default_hparams(**over):
    """Constructor defaults of the reference VTN (seq2seq_vc/models/vtn.py:15-62)."""
    hp = dict(idim=80, odim=80, dprenet_layers=2, dprenet_units=256, adim=384, aheads=4,
              elayers=6, eunits=1536, dlayers=6, dunits=1536, postnet_layers=5, postnet_filts=5,
              postnet_chans=256, decoder_reduction_factor=2)
    hp.update(over)
    return hp


def sinusoid_table(length: int, d_model: int) -> torch.Tensor:
    """PE table, seq2seq_vc/layers/positional_encoding.py:36-57 (non-reversed branch)."""
    pos = torch.arange(0, length, dtype=torch.float32).unsqueeze(1)
    div = torch.exp(torch.arange(0, d_model, 2, dtype=torch.float32) * -(math.log(10000.0) / d_model))
    pe = torch.zeros(length, d_model)
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe


def non_pad_mask(lengths: Sequence[int], maxlen: int) -> torch.Tensor:
    """(B, maxlen) bool, True on valid frames. seq2seq_vc/layers/utils.py:90-121,124-210."""
    lens = torch.as_tensor(list(lengths), dtype=torch.int64)
    return torch.arange(maxlen, dtype=torch.int64)[None, :] < lens[:, None]


def causal_mask(n: int) -> torch.Tensor:
    """Lower-triangular (n, n) bool. seq2seq_vc/modules/transformer/mask.py:9-22."""
    return torch.ones(n, n, dtype=torch.bool).tril()


def layer_norm(x, sd, prefix):
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + ".weight"], sd[prefix + ".bias"], LN_EPS)


def linear(x, sd, prefix):
    return F.linear(x, sd[prefix + ".weight"], sd[prefix + ".bias"])


def attention(sd, prefix, q_in, kv_in, mask, n_head, store: Dict[str, torch.Tensor] | None = None):
    """MultiHeadedAttention.forward, seq2seq_vc/modules/transformer/attention.py:40-111.

    mask: (B, 1, T2) or (B, T1, T2) bool (True = attend).  Masked scores are filled with
    finfo.min before the softmax and the probabilities are zeroed after it (:76-83).
    """
    B, T1, d = q_in.shape
    T2 = kv_in.shape[1]
    dk = d // n_head
    q = linear(q_in, sd, prefix + ".linear_q").view(B, T1, n_head, dk).transpose(1, 2)
    k = linear(kv_in, sd, prefix + ".linear_k").view(B, T2, n_head, dk).transpose(1, 2)
    v = linear(kv_in, sd, prefix + ".linear_v").view(B, T2, n_head, dk).transpose(1, 2)
    scores = torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(dk)
    dead = ~mask.unsqueeze(1)
    scores = scores.masked_fill(dead, torch.finfo(scores.dtype).min)
    p = torch.softmax(scores, dim=-1).masked_fill(dead, 0.0)
    if store is not None:
        store[prefix] = p
    ctx = torch.matmul(p, v).transpose(1, 2).reshape(B, T1, d)
    return linear(ctx, sd, prefix + ".linear_out")


def feed_forward(sd, prefix, x):
    """PositionwiseFeedForward (ReLU), modules/transformer/positionwise_feed_forward.py:12-32; told apart by the weight rank:
    MultiLayeredConv1d (Conv1d(k) -> ReLU -> Conv1d(k)) and Conv1dLinear (Conv1d(k) -> ReLU -> Linear) over time with zero padding
    (k-1)//2 (modules/transformer/multi_layer_conv.py:12-108)."""
    w1 = sd[prefix + ".w_1.weight"]
    if w1.dim() == 3:
        k = w1.shape[2]
        h = torch.relu(F.conv1d(x.transpose(1, 2), w1, sd[prefix + ".w_1.bias"], padding=(k - 1) // 2))
        if sd[prefix + ".w_2.weight"].dim() == 2:
            return linear(h.transpose(1, 2), sd, prefix + ".w_2")
        return F.conv1d(h, sd[prefix + ".w_2.weight"], sd[prefix + ".w_2.bias"], padding=(k - 1) // 2).transpose(1, 2)
    return linear(torch.relu(linear(x, sd, prefix + ".w_1")), sd, prefix + ".w_2")


def conv2d_subsample(sd, prefix, xs, x_mask):
    """Conv2dSubsampling + ScaledPositionalEncoding, modules/transformer/subsampling.py:74-94."""
    x = xs.unsqueeze(1)
    x = torch.relu(F.conv2d(x, sd[prefix + ".conv.0.weight"], sd[prefix + ".conv.0.bias"], stride=2))
    x = torch.relu(F.conv2d(x, sd[prefix + ".conv.2.weight"], sd[prefix + ".conv.2.bias"], stride=2))
    b, c, t, f = x.shape
    x = linear(x.transpose(1, 2).reshape(b, t, c * f), sd, prefix + ".out.0")
    x = x + sd[prefix + ".out.1.alpha"] * sinusoid_table(t, x.shape[-1])[None]
    return x, x_mask[:, :, :-2:2][:, :, :-2:2]


LEGACY_PE_MAX_LEN = 5000


def legacy_rel_pos_table(T: int, d: int) -> torch.Tensor:
    """pos_emb of LegacyRelPositionalEncoding.forward (layers/positional_encoding.py:192-235): the table is built ONCE at
    construction, reversed, for max_len = 5000 (positional_encoding.py:44,52-60) and then only sliced, so row k holds the
    sinusoid of position 4999 - k whatever the utterance length."""
    pos = torch.arange(LEGACY_PE_MAX_LEN - 1, LEGACY_PE_MAX_LEN - 1 - T, -1, dtype=torch.float32).unsqueeze(1)
    div = torch.exp(torch.arange(0, d, 2, dtype=torch.float32) * -(math.log(10000.0) / d))
    pe = torch.zeros(T, d)
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe


def legacy_rel_shift(bd: torch.Tensor) -> torch.Tensor:
    """LegacyRelPositionMultiHeadedAttention.rel_shift (attention.py:138-157) restated as an index map: with one zero column
    padded on the left, out[i, j] is flat element T + i T + j of the (T, T + 1) block, i.e. zero when that index is a multiple
    of T + 1 and bd[f // (T + 1), f % (T + 1) - 1] otherwise (rows wrap around)."""
    T = bd.shape[-1]
    f = T + torch.arange(T)[:, None] * T + torch.arange(T)[None, :]
    r, c = f // (T + 1), f % (T + 1)
    src = (r * T + (c - 1).clamp(min=0)).reshape(-1)
    out = bd.reshape(*bd.shape[:-2], T * T)[..., src].reshape(bd.shape)
    return out * (c > 0).to(bd.dtype)


def legacy_rel_attention(sd, prefix, x, pos_emb, mask, n_head, store=None):
    """LegacyRelPositionMultiHeadedAttention.forward (attention.py:159-207)."""
    B, T, d = x.shape
    dk = d // n_head
    q = linear(x, sd, prefix + ".linear_q").view(B, T, n_head, dk)
    k = linear(x, sd, prefix + ".linear_k").view(B, T, n_head, dk).transpose(1, 2)
    v = linear(x, sd, prefix + ".linear_v").view(B, T, n_head, dk).transpose(1, 2)
    p = F.linear(pos_emb, sd[prefix + ".linear_pos.weight"]).view(T, n_head, dk).transpose(0, 1)      # (H, T, dk)
    qu = (q + sd[prefix + ".pos_bias_u"]).transpose(1, 2)
    qv = (q + sd[prefix + ".pos_bias_v"]).transpose(1, 2)
    ac = torch.matmul(qu, k.transpose(-2, -1))
    bd = legacy_rel_shift(torch.matmul(qv, p.transpose(-2, -1).unsqueeze(0)))
    scores = (ac + bd) / math.sqrt(dk)
    dead = ~mask.unsqueeze(1)
    scores = scores.masked_fill(dead, torch.finfo(scores.dtype).min)
    pr = torch.softmax(scores, dim=-1).masked_fill(dead, 0.0)
    if store is not None:
        store[prefix] = pr
    ctx = torch.matmul(pr, v).transpose(1, 2).reshape(B, T, d)
    return linear(ctx, sd, prefix + ".linear_out")


def conformer_encoder(sd, hp, xs, x_mask, training=True, bn_stats=None, attn_store=None):
    """ConformerEncoder as VTN builds it (models/vtn.py:122-143): Conv2dSubsampling whose positional layer is the (legacy) rel-pos
    encoding (x * sqrt(d), pos_emb on the side), macaron conformer blocks with the convolution module, after_norm
    (conformer/encoder.py:249-293, encoder_layer.py:79-179)."""
    from oracle import aasvc_oracle as ao           # conformer block pieces shared with AAS-VC (imported lazily: it imports this module)

    prefix = "encoder.embed"
    x = xs.unsqueeze(1)
    x = torch.relu(F.conv2d(x, sd[prefix + ".conv.0.weight"], sd[prefix + ".conv.0.bias"], stride=2))
    x = torch.relu(F.conv2d(x, sd[prefix + ".conv.2.weight"], sd[prefix + ".conv.2.bias"], stride=2))
    b, c, t, f = x.shape
    x = linear(x.transpose(1, 2).reshape(b, t, c * f), sd, prefix + ".out.0")
    mask = x_mask[:, :, :-2:2][:, :, :-2:2]
    d = x.shape[-1]
    legacy = hp.get("conformer_rel_pos_type", "legacy") == "legacy"
    x = x * math.sqrt(d)
    pos_emb = legacy_rel_pos_table(t, d) if legacy else ao.rel_pos_table(t, d)
    att = legacy_rel_attention if legacy else ao.rel_attention
    for l in range(hp["elayers"]):
        p = f"encoder.encoders.{l}"
        x = x + 0.5 * ao.ffn_swish(sd, p + ".feed_forward_macaron", ao.layer_norm(x, sd, p + ".norm_ff_macaron"))
        x = x + att(sd, p + ".self_attn", ao.layer_norm(x, sd, p + ".norm_mha"), pos_emb, mask, hp["aheads"], attn_store)
        x = x + ao.conv_module(sd, p + ".conv_module", ao.layer_norm(x, sd, p + ".norm_conv"), training, bn_stats)
        x = x + 0.5 * ao.ffn_swish(sd, p + ".feed_forward", ao.layer_norm(x, sd, p + ".norm_ff"))
        x = ao.layer_norm(x, sd, p + ".norm_final")
    return ao.layer_norm(x, sd, "encoder.after_norm"), mask


def encoder(sd, hp, xs, x_mask, attn_store=None, training=True, bn_stats=None):
    """Transformer Encoder (pre-LN), modules/transformer/encoder.py:283-329 + encoder_layer.py:61-119; the conformer encoder of
    VTN(encoder_type="conformer") when the hyper-parameters say so."""
    if hp.get("encoder_type", "transformer") == "conformer":
        return conformer_encoder(sd, hp, xs, x_mask, training, bn_stats, attn_store)
    x, mask = conv2d_subsample(sd, "encoder.embed", xs, x_mask)
    for l in range(hp["elayers"]):
        p = f"encoder.encoders.{l}"
        n = layer_norm(x, sd, p + ".norm1")
        x = x + attention(sd, p + ".self_attn", n, n, mask, hp["aheads"], attn_store)
        n = layer_norm(x, sd, p + ".norm2")
        x = x + feed_forward(sd, p + ".feed_forward", n)
    return layer_norm(x, sd, "encoder.after_norm"), mask


def encoder_tts(sd, hp, tokens, x_mask, attn_store=None):
    """TransformerTTS encoder: Embedding(padding_idx=0) + ScaledPE input layer, then the same pre-LN
    layers (models/transformer_tts.py:63-77, modules/transformer/encoder.py:131-135,283-329)."""
    x = F.embedding(tokens, sd["encoder.embed.0.weight"], padding_idx=0)
    x = x + sd["encoder.embed.1.alpha"] * sinusoid_table(x.shape[1], x.shape[-1])[None]
    for l in range(hp["elayers"]):
        p = f"encoder.encoders.{l}"
        n = layer_norm(x, sd, p + ".norm1")
        x = x + attention(sd, p + ".self_attn", n, n, x_mask, hp["aheads"], attn_store)
        n = layer_norm(x, sd, p + ".norm2")
        x = x + feed_forward(sd, p + ".feed_forward", n)
    return layer_norm(x, sd, "encoder.after_norm"), x_mask


def guided_attention_loss(att_ws, ilens, olens, sigma=0.4, alpha=1.0):
    """GuidedMultiHeadAttentionLoss.forward, losses/guided_attention_loss.py:57-96,142-165."""
    B, H, T_out, T_in = att_ws.shape
    w = torch.zeros(B, T_out, T_in)
    m = torch.zeros(B, T_out, T_in, dtype=torch.bool)
    for b, (il, ol) in enumerate(zip(ilens, olens)):
        gy, gx = torch.meshgrid(torch.arange(ol).float(), torch.arange(il).float(), indexing="ij")
        w[b, :ol, :il] = 1.0 - torch.exp(-((gx / il - gy / ol) ** 2) / (2 * sigma ** 2))
        m[b, :ol, :il] = True
    losses = w.unsqueeze(1) * att_ws
    return alpha * losses.masked_select(m.unsqueeze(1).expand_as(losses)).mean()


def tts_forward(sd, hp, tokens, ilens, ys, labels, olens, training: bool = True, use_guided_attn_loss: bool = False,
                num_heads_applied_guided_attn: int = 2, num_layers_applied_guided_attn: int = 2):
    """TransformerTTS.forward, seq2seq_vc/models/transformer_tts.py:129-229."""
    ilens = [int(v) for v in ilens]
    olens = [int(v) for v in olens]
    r, odim = hp["decoder_reduction_factor"], hp["odim"]
    tokens = tokens[:, : max(ilens)]
    ys = ys[:, : max(olens)]
    labels = labels[:, : max(olens)]
    tokens = F.pad(tokens, [0, 1], "constant", 0).clone()                         # :139-142
    for b, l in enumerate(ilens):
        tokens[b, l] = hp["idim"] - 1
    ilens = [i + 1 for i in ilens]
    attn: Dict[str, torch.Tensor] = {}
    x_mask = non_pad_mask(ilens, tokens.shape[1]).unsqueeze(-2)
    hs, h_mask = encoder_tts(sd, hp, tokens, x_mask, attn)
    ys_in = ys[:, r - 1 :: r] if r > 1 else ys
    olens_in = [o // r for o in olens]
    ys_in = torch.cat([ys_in.new_zeros(ys_in.shape[0], 1, odim), ys_in[:, :-1]], dim=1)
    y_mask = non_pad_mask(olens_in, ys_in.shape[1]).unsqueeze(-2) & causal_mask(ys_in.shape[1])[None]
    zs = decoder(sd, hp, ys_in, y_mask, hs, h_mask, attn)
    B = zs.shape[0]
    before = linear(zs, sd, "feat_out").view(B, -1, odim)
    logits = linear(zs, sd, "prob_out").view(B, -1)
    after = before + postnet(sd, hp, before.transpose(1, 2), training).transpose(1, 2)
    if r > 1:
        olens = [o - o % r for o in olens]
        ys = ys[:, : max(olens)]
        labels = labels[:, : max(olens)].clone()
        for b, o in enumerate(olens):
            labels[b, o - 1] = 1.0
    att_ws = []
    if use_guided_attn_loss:                                                       # :205-219
        for idx, l in enumerate(reversed(range(hp["dlayers"]))):
            att_ws.append(attn[f"decoder.decoders.{l}.src_attn"][:, :num_heads_applied_guided_attn])
            if idx + 1 == num_layers_applied_guided_attn:
                break
        att_ws = torch.cat(att_ws, dim=1)
    return dict(after_outs=after, before_outs=before, logits=logits, ys=ys, labels=labels, olens=olens, att_ws=att_ws,
                ilens=ilens, olens_in=olens_in, attn=attn)


def decoder(sd, hp, ys_in, y_mask, memory, mem_mask, attn_store=None):
    """Transformer Decoder (post-LN), modules/transformer/decoder.py:207-237 + decoder_layer.py:63-134.

    Input layer: Prenet (Linear+ReLU)x n, Linear(units->adim), ScaledPE
    (modules/pre_postnets.py:14-66; models/vtn.py:155-163).
    """
    x = ys_in
    for i in range(hp["dprenet_layers"]):
        x = torch.relu(linear(x, sd, f"decoder.embed.0.0.prenet.{i}.0"))
    x = linear(x, sd, "decoder.embed.0.1")
    x = x + sd["decoder.embed.1.alpha"] * sinusoid_table(x.shape[1], x.shape[-1])[None]
    for l in range(hp["dlayers"]):
        p = f"decoder.decoders.{l}"
        x = layer_norm(x + attention(sd, p + ".self_attn", x, x, y_mask, hp["aheads"], attn_store), sd, p + ".norm1")
        x = layer_norm(x + attention(sd, p + ".src_attn", x, memory, mem_mask, hp["aheads"], attn_store), sd, p + ".norm2")
        x = layer_norm(x + feed_forward(sd, p + ".feed_forward", x), sd, p + ".norm3")
    return x


def postnet(sd, hp, x_bct, training: bool, bn_stats: Dict[str, torch.Tensor] | None = None):
    """Postnet: n x [Conv1d(k, no bias) -> BatchNorm1d -> tanh] (last without tanh).

    modules/pre_postnets.py:69-185.  training=True uses batch statistics over (B, T) including
    padded frames (no mask is applied anywhere, SURVEY.md §7 hard part 5/6) and, when ``bn_stats``
    is given, records the running-stat update (momentum 0.1, unbiased variance).
    """
    n_layers = hp["postnet_layers"]
    pad = (hp["postnet_filts"] - 1) // 2
    for i in range(n_layers):
        p = f"postnet.postnet.{i}"
        x_bct = F.conv1d(x_bct, sd[p + ".0.weight"], None, stride=1, padding=pad)
        if training:
            mean = x_bct.mean(dim=(0, 2))
            var_b = x_bct.var(dim=(0, 2), unbiased=False)
            if bn_stats is not None:
                n = x_bct.shape[0] * x_bct.shape[2]
                bn_stats[p + ".1.running_mean"] = (1 - BN_MOMENTUM) * sd[p + ".1.running_mean"] + BN_MOMENTUM * mean.detach()
                bn_stats[p + ".1.running_var"] = (1 - BN_MOMENTUM) * sd[p + ".1.running_var"] + BN_MOMENTUM * (var_b.detach() * n / (n - 1))
        else:
            mean, var_b = sd[p + ".1.running_mean"], sd[p + ".1.running_var"]
        x_bct = (x_bct - mean[None, :, None]) * torch.rsqrt(var_b[None, :, None] + BN_EPS)
        x_bct = x_bct * sd[p + ".1.weight"][None, :, None] + sd[p + ".1.bias"][None, :, None]
        if i != n_layers - 1:
            x_bct = torch.tanh(x_bct)
    return x_bct


def vtn_forward(sd, hp, xs, ilens, ys, labels, olens, training: bool = True, bn_stats=None):
    """VTN.forward, seq2seq_vc/models/vtn.py:207-300.  ilens / olens: python int lists.

    Returns a dict with the reference's 7-tuple fields plus every attention map under "attn".
    """
    ilens = [int(v) for v in ilens]
    olens = [int(v) for v in olens]
    r = hp["decoder_reduction_factor"]
    odim = hp["odim"]
    xs = xs[:, : max(ilens)]
    ys = ys[:, : max(olens)]
    labels = labels[:, : max(olens)]
    attn: Dict[str, torch.Tensor] = {}

    x_mask = non_pad_mask(ilens, xs.shape[1]).unsqueeze(-2)                        # vtn.py:217,553-572
    hs, h_mask = encoder(sd, hp, xs, x_mask, attn, training, bn_stats)

    ys_in = ys[:, r - 1 :: r] if r > 1 else ys                                   # vtn.py:227-240
    olens_in = [o // r for o in olens]
    ys_in = torch.cat([ys_in.new_zeros(ys_in.shape[0], 1, odim), ys_in[:, :-1]], dim=1)  # vtn.py:523-527
    y_mask = non_pad_mask(olens_in, ys_in.shape[1]).unsqueeze(-2) & causal_mask(ys_in.shape[1])[None]  # vtn.py:574-602
    zs = decoder(sd, hp, ys_in, y_mask, hs, h_mask, attn)

    B = zs.shape[0]
    before = linear(zs, sd, "feat_out").view(B, -1, odim)                         # vtn.py:249
    logits = linear(zs, sd, "prob_out").view(B, -1)                               # vtn.py:251
    after = before + postnet(sd, hp, before.transpose(1, 2), training, bn_stats).transpose(1, 2)

    if r > 1:                                                                     # vtn.py:262-274
        assert all(o >= r for o in olens)
        olens = [o - o % r for o in olens]
        ys = ys[:, : max(olens)]
        labels = labels[:, : max(olens)].clone()
        for b, o in enumerate(olens):
            labels[b, o - 1] = 1.0
    ilens_ds = [((i - 2 + 1) // 2 - 2 + 1) // 2 for i in ilens]                   # vtn.py:279
    att_ws = [attn[f"decoder.decoders.{l}.src_attn"] for l in reversed(range(hp["dlayers"]))]  # vtn.py:280-287
    return dict(after_outs=after, before_outs=before, logits=logits, ys=ys, labels=labels,
                olens=olens, att_ws=att_ws, ilens_ds_st=ilens_ds, olens_in=olens_in, attn=attn)


def seq2seq_loss(after, before, logits, ys, labels, olens, bce_pos_weight: float = 10.0):
    """Seq2SeqLoss.forward, seq2seq_vc/losses/seq2seq_loss.py:30-59 -> (l1_loss, bce_loss)."""
    m = non_pad_mask(olens, ys.shape[1])
    m3 = m.unsqueeze(-1)
    y = ys.masked_select(m3)
    l1 = (after.masked_select(m3) - y).abs().mean() + (before.masked_select(m3) - y).abs().mean()
    bce = F.binary_cross_entropy_with_logits(logits.masked_select(m), labels.masked_select(m),
                                             pos_weight=torch.tensor(bce_pos_weight))
    return l1, bce


def vtn_loss_and_grads(sd, hp, xs, ilens, ys, labels, olens, training=True):
    """Forward + Seq2SeqLoss + autograd grads for every floating parameter of ``sd``."""
    params = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items() if v.dtype.is_floating_point and "running_" not in k}
    full = dict(sd)
    full.update(params)
    out = vtn_forward(full, hp, xs, ilens, ys, labels, olens, training=training)
    l1, bce = seq2seq_loss(out["after_outs"], out["before_outs"], out["logits"], out["ys"], out["labels"], out["olens"])
    loss = l1 + bce
    grads = torch.autograd.grad(loss, list(params.values()), allow_unused=True)
    return out, (l1, bce), {k: g for k, g in zip(params.keys(), grads)}


def init_state_dict(hp, seed: int = 0, scale: float = 1.0) -> Dict[str, torch.Tensor]:
    """Random VTN state dict with the reference's key names and shapes (SURVEY.md §8b 'state dict').

    The initial values follow torch's default Linear/Conv init *distributions* (uniform
    +-1/sqrt(fan_in)); they are not meant to reproduce torch's RNG stream.
    """
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def uni(shape, fan_in):
        b = scale / math.sqrt(fan_in)
        return (torch.rand(shape, generator=g) * 2 - 1) * b

    def lin(name, o, i):
        sd[name + ".weight"] = uni((o, i), i)
        sd[name + ".bias"] = uni((o,), i)

    def ln(name, n):
        sd[name + ".weight"] = 1 + 0.1 * (torch.rand(n, generator=g) - 0.5)
        sd[name + ".bias"] = 0.1 * (torch.rand(n, generator=g) - 0.5)

    def mha(name, d):
        for s in ("linear_q", "linear_k", "linear_v", "linear_out"):
            lin(f"{name}.{s}", d, d)

    d, idim, odim = hp["adim"], hp["idim"], hp["odim"]
    f2 = ((idim - 1) // 2 - 1) // 2
    sd["encoder.embed.conv.0.weight"] = uni((d, 1, 3, 3), 9)
    sd["encoder.embed.conv.0.bias"] = uni((d,), 9)
    sd["encoder.embed.conv.2.weight"] = uni((d, d, 3, 3), 9 * d)
    sd["encoder.embed.conv.2.bias"] = uni((d,), 9 * d)
    lin("encoder.embed.out.0", d, d * f2)
    sd["encoder.embed.out.1.alpha"] = torch.tensor(1.0)
    for l in range(hp["elayers"]):
        p = f"encoder.encoders.{l}"
        mha(p + ".self_attn", d)
        lin(p + ".feed_forward.w_1", hp["eunits"], d)
        lin(p + ".feed_forward.w_2", d, hp["eunits"])
        ln(p + ".norm1", d)
        ln(p + ".norm2", d)
    ln("encoder.after_norm", d)
    u = hp["dprenet_units"]
    for i in range(hp["dprenet_layers"]):
        lin(f"decoder.embed.0.0.prenet.{i}.0", u, odim if i == 0 else u)
    lin("decoder.embed.0.1", d, u)
    sd["decoder.embed.1.alpha"] = torch.tensor(1.0)
    for l in range(hp["dlayers"]):
        p = f"decoder.decoders.{l}"
        mha(p + ".self_attn", d)
        mha(p + ".src_attn", d)
        lin(p + ".feed_forward.w_1", hp["dunits"], d)
        lin(p + ".feed_forward.w_2", d, hp["dunits"])
        for n in ("norm1", "norm2", "norm3"):
            ln(f"{p}.{n}", d)
    r = hp["decoder_reduction_factor"]
    lin("feat_out", odim * r, d)
    lin("prob_out", r, d)
    ch, k = hp["postnet_chans"], hp["postnet_filts"]
    for i in range(hp["postnet_layers"]):
        ic = odim if i == 0 else ch
        oc = odim if i == hp["postnet_layers"] - 1 else ch
        p = f"postnet.postnet.{i}"
        sd[p + ".0.weight"] = uni((oc, ic, k), ic * k)
        ln(p + ".1", oc)
        sd[p + ".1.running_mean"] = torch.zeros(oc)
        sd[p + ".1.running_var"] = torch.ones(oc)
        sd[p + ".1.num_batches_tracked"] = torch.tensor(0, dtype=torch.int64)
    return sd


def synthetic_batch(B, T, L, idim=80, odim=80, ilens=None, olens=None, seed=1234):
    """Seeded synthetic padded mel batch: N(0,1) features, zero padding,
    stop labels 1 from the last valid frame on (seq2seq_vc/collaters/ar_vc.py:48-62)."""
    g = torch.Generator().manual_seed(seed)
    ilens = list(ilens) if ilens is not None else [T] * B
    olens = list(olens) if olens is not None else [L] * B
    xs = torch.randn(B, T, idim, generator=g)
    ys = torch.randn(B, L, odim, generator=g)
    labels = torch.zeros(B, L)
    for b in range(B):
        xs[b, ilens[b]:] = 0
        ys[b, olens[b]:] = 0
        labels[b, olens[b] - 1:] = 1.0
    return xs, ilens, ys, labels, olens


def vtn_inference(sd, hp, x, threshold=0.5, minlenratio=0.0, maxlenratio=10.0, tts: bool = False):
    """VTN.inference (models/vtn.py:302-394) restated by full-prefix recomputation (equal to forward_one_step by
    causality); eval-mode BatchNorm; Prenet dropout must be 0 for a deterministic comparison.
    x (T, idim) -> (outs (L, odim), probs (L,), att_ws (#dlayers, H, L/r, T'))."""
    hp = default_hparams(**hp)
    r, odim = hp["decoder_reduction_factor"], hp["odim"]
    if tts:     # TransformerTTS.inference (models/transformer_tts.py:231-330): <eos> = idim - 1 appended, token-embedding encoder
        xs = F.pad(x, [0, 1], "constant", hp["idim"] - 1).unsqueeze(0)
        hs, _ = encoder_tts(sd, hp, xs, torch.ones(1, 1, xs.shape[1], dtype=torch.bool))
    else:
        xs = x.unsqueeze(0)
        T = xs.shape[1]
        hs, _ = encoder(sd, hp, xs, torch.ones(1, 1, T, dtype=torch.bool), training=False)      # model.eval(): running BatchNorm statistics
    T2 = hs.shape[1]
    mem_mask = torch.ones(1, 1, T2, dtype=torch.bool)
    maxlen, minlen = int(T2 * maxlenratio / r), int(T2 * minlenratio / r)
    ys = hs.new_zeros(1, 1, odim)
    outs, probs = [], []
    idx = 0
    while True:
        idx += 1
        attn: Dict[str, torch.Tensor] = {}
        zs = decoder(sd, hp, ys, causal_mask(idx)[None], hs, mem_mask, attn)
        z = zs[:, -1]
        outs.append(linear(z, sd, "feat_out").view(r, odim))
        probs.append(torch.sigmoid(linear(z, sd, "prob_out"))[0])
        ys = torch.cat([ys, outs[-1][-1].view(1, 1, odim)], dim=1)
        if int(sum(probs[-1] >= threshold)) > 0 or idx >= maxlen:
            if idx < minlen:
                continue
            o = torch.cat(outs, dim=0).unsqueeze(0).transpose(1, 2)
            o = o + postnet(sd, hp, o, False)
            att = torch.stack([attn[f"decoder.decoders.{l}.src_attn"][0] for l in range(hp["dlayers"])], dim=0)
            return o.transpose(2, 1).squeeze(0), torch.cat(probs, dim=0), att
