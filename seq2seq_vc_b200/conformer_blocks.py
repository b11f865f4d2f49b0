"""Conformer encoder blocks shared by the AAS-VC engine and the conformer-encoder VTN: explicit forward / backward over the
C-ABI kernels (reference: modules/conformer/{encoder,encoder_layer,convolution}.py, modules/transformer/attention.py:114-305,
layers/positional_encoding.py:192-309, modules/transformer/multi_layer_conv.py).

A mixin over EngineBase (buffers, scratch, LayerNorm / Linear helpers, dropout sites).  The host class provides `hp` with
`positionwise_layer_type` / `positionwise_conv_kernel_size` and, optionally, `conformer_rel_pos_type`:
  "latest"  RelPositionMultiHeadedAttention + RelPositionalEncoding (2T-1 relative positions, attention.py:209-305) -- AAS-VC;
  "legacy"  LegacyRelPositionMultiHeadedAttention + LegacyRelPositionalEncoding (T positions taken from a reversed absolute
            table, row k = PE(4999 - k), and the wrap-around rel_shift, attention.py:114-207) -- the default of
            VTN(encoder_type="conformer") (models/vtn.py:83-99).
"""
from __future__ import annotations

import math
from typing import Dict

import torch

from . import ops
from .engine_base import _r8

_f32 = torch.float32
LEGACY_PE_MAX_LEN = 5000        # LegacyRelPositionalEncoding builds its reversed table once for max_len = 5000 (positional_encoding.py:44,200)


def rel_pos_table(T: int, d: int) -> torch.Tensor:
    """pos_emb (2T-1, d) of RelPositionalEncoding (layers/positional_encoding.py:263-309): row k = PE(T-1-k)."""
    pos = torch.arange(T - 1, -T, -1, dtype=_f32).unsqueeze(1)
    div = torch.exp(torch.arange(0, d, 2, dtype=_f32) * -(math.log(10000.0) / d))
    pe = torch.zeros(2 * T - 1, d)
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe


def legacy_rel_pos_table(T: int, d: int) -> torch.Tensor:
    """pos_emb (T, d) of LegacyRelPositionalEncoding (layers/positional_encoding.py:192-235): the first T rows of an absolute
    table built once, REVERSED, for max_len = 5000 -- row k = PE(4999 - k), whatever the utterance length."""
    assert T <= LEGACY_PE_MAX_LEN, "sequences longer than the reference's positional table (5000 frames)"
    pos = torch.arange(LEGACY_PE_MAX_LEN - 1, LEGACY_PE_MAX_LEN - 1 - T, -1, dtype=_f32).unsqueeze(1)
    div = torch.exp(torch.arange(0, d, 2, dtype=_f32) * -(math.log(10000.0) / d))
    pe = torch.zeros(T, d)
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe


def conformer_param_groups(g, hp: dict, prefix: str, n_layers: int, dm: int, units: int, k: int, H: int) -> None:
    """Appends the reference state-dict names / shapes of `n_layers` conformer blocks + after_norm under `prefix` to the group
    list `g` (Q / K / V projections adjacent: one fused GEMM)."""
    def lin(name, o, i, bias=True):
        g.append([(name + ".weight", (o, i))])
        if bias:
            g.append([(name + ".bias", (o,))])

    def ln(name, n):
        g.append([(name + ".weight", (n,))])
        g.append([(name + ".bias", (n,))])

    for l in range(n_layers):
        p = f"{prefix}.encoders.{l}"
        g.append([(p + ".self_attn.pos_bias_u", (H, dm // H))])
        g.append([(p + ".self_attn.pos_bias_v", (H, dm // H))])
        g.append([(f"{p}.self_attn.{s}.weight", (dm, dm)) for s in ("linear_q", "linear_k", "linear_v")])
        g.append([(f"{p}.self_attn.{s}.bias", (dm,)) for s in ("linear_q", "linear_k", "linear_v")])
        lin(p + ".self_attn.linear_out", dm, dm)
        lin(p + ".self_attn.linear_pos", dm, dm, bias=False)
        for ff in ("feed_forward", "feed_forward_macaron"):
            if hp.get("positionwise_layer_type", "linear") != "linear":     # Conv1d weights keep their (out, in, k) shape
                pk = hp.get("positionwise_conv_kernel_size", 1)
                g.append([(f"{p}.{ff}.w_1.weight", (units, dm, pk))])
                g.append([(f"{p}.{ff}.w_1.bias", (units,))])
                g.append([(f"{p}.{ff}.w_2.weight", (dm, units, pk) if hp["positionwise_layer_type"] == "conv1d" else (dm, units))])
                g.append([(f"{p}.{ff}.w_2.bias", (dm,))])
            else:
                lin(f"{p}.{ff}.w_1", units, dm)
                lin(f"{p}.{ff}.w_2", dm, units)
        g.append([(p + ".conv_module.pointwise_conv1.weight", (2 * dm, dm, 1))])
        g.append([(p + ".conv_module.pointwise_conv1.bias", (2 * dm,))])
        g.append([(p + ".conv_module.depthwise_conv.weight", (dm, 1, k))])
        g.append([(p + ".conv_module.depthwise_conv.bias", (dm,))])
        ln(p + ".conv_module.norm", dm)
        g.append([(p + ".conv_module.pointwise_conv2.weight", (dm, dm, 1))])
        g.append([(p + ".conv_module.pointwise_conv2.bias", (dm,))])
        for n in ("norm_ff", "norm_mha", "norm_ff_macaron", "norm_conv", "norm_final"):
            ln(f"{p}.{n}", dm)
    ln(prefix + ".after_norm", dm)


def conformer_buffer_specs(prefix: str, n_layers: int, dm: int):
    """BatchNorm running statistics of the convolution modules."""
    out = []
    for l in range(n_layers):
        p = f"{prefix}.encoders.{l}.conv_module.norm"
        out.extend([(p + ".running_mean", (dm,), _f32), (p + ".running_var", (dm,), _f32), (p + ".num_batches_tracked", (), torch.int64)])
    return out


class ConformerBlocks:
    """Mixin: conformer blocks (macaron FFN, rel-pos self-attention, convolution module, FFN, final LayerNorm)."""

    def _legacy_rel(self) -> bool:
        return self.hp.get("conformer_rel_pos_type", "latest") == "legacy"

    def _n_pos(self, T: int) -> int:
        return T if self._legacy_rel() else 2 * T - 1

    def _rel_table(self, T: int, d: int) -> torch.Tensor:
        cache: Dict = self.__dict__.setdefault("_relpe", {})
        key = (T, d, self._legacy_rel())
        t = cache.get(key)
        if t is None:
            t = (legacy_rel_pos_table if self._legacy_rel() else rel_pos_table)(T, d).to(self.device)
            cache[key] = t
        return t

    def _ffn_fwd(self, x, p, ff, tag, U, rate, out):
        """out = x + 0.5 * dropout(w_2(dropout(act(w_1 LN(x)))))   (encoder_layer.py:115-123,157-163); act = Swish for the
        "linear" position-wise layer, ReLU (in the GEMM epilogue, with its dropout) for "conv1d" k = 1 (multi_layer_conv.py:13-62)."""
        st = self.store
        B, T, dm = x.shape
        norm = "norm_ff_macaron" if ff == "feed_forward_macaron" else "norm_ff"
        n = self._ln_fwd(x, f"{p}.{norm}", f"{tag}.ln")
        if self._ffn_is_conv():
            return self._ffn_conv_fwd(x, n, p, ff, tag, U, rate, out)
        h = self.buf(tag + ".h", (B * T, U))
        w1, w2 = self.W(f"{p}.{ff}.w_1.weight").view(U, dm), self.W(f"{p}.{ff}.w_2.weight").view(dm, U)
        if self.hp["positionwise_layer_type"] == "conv1d":
            self._lin_fwd(n.view(B * T, dm), w1, st.p(f"{p}.{ff}.w_1.bias"), h, relu=True, drop=self.named_drop(tag + ".d1", rate))
        else:
            hpre = self.buf(tag + ".hpre", (B * T, U))
            self._lin_fwd(n.view(B * T, dm), w1, st.p(f"{p}.{ff}.w_1.bias"), hpre)
            ops.swish_fwd(hpre, h, self.named_drop(tag + ".d1", rate))
        bh = self.buf(tag + ".bhalf", (dm,), _f32)
        ops.scale_dropout(st.p(f"{p}.{ff}.w_2.bias"), bh, 0.5)
        ops.gemm(h, w2, out.view(B * T, dm), bias=bh, alpha=0.5, drop=self.named_drop(tag + ".d2", rate),
                 residual=x.view(B * T, dm), mode=self.mode)
        return out

    def _ffn_bwd(self, g, x, p, ff, tag, U, rate, gout):
        """g = d(out) (B,T,dm) -> gout = d(x) = g + LN'(...) ; accumulates the FFN parameter gradients."""
        st = self.store
        B, T, dm = x.shape
        norm = "norm_ff_macaron" if ff == "feed_forward_macaron" else "norm_ff"
        n = self.buf(f"{tag}.ln.y", (B, T, dm))
        if self._ffn_is_conv():
            return self._ffn_conv_bwd(g, x, n, p, ff, tag, U, rate, gout)
        h = self.buf(tag + ".h", (B * T, U))
        dy = self._scratch("cf.dy", (B * T, dm))
        ops.scale_dropout(g.view(B * T, dm), dy, 0.5, self.named_drop(tag + ".d2", rate))
        dh = self._scratch("cf.dh", (B * T, U))
        w1, w2 = self.W(f"{p}.{ff}.w_1.weight").view(U, dm), self.W(f"{p}.{ff}.w_2.weight").view(dm, U)
        self._lin_bwd(dy, h, w2, st.g(f"{p}.{ff}.w_2.weight").view(dm, U), st.g(f"{p}.{ff}.w_2.bias"), dx=dh)
        if self.hp["positionwise_layer_type"] == "conv1d":
            ops.relu_bwd(dh, h, dh, self.named_drop(tag + ".d1", rate).scale)    # h = dropout(relu(.)): zero where cut or dropped
        else:
            ops.swish_bwd(dh, self.buf(tag + ".hpre", (B * T, U)), dh, self.named_drop(tag + ".d1", rate))
        dn = self._scratch("cf.dn", (B, T, dm))
        self._lin_bwd(dh, n.view(B * T, dm), w1, st.g(f"{p}.{ff}.w_1.weight").view(U, dm), st.g(f"{p}.{ff}.w_1.bias"),
                      dx=dn.view(B * T, dm))
        self._ln_bwd(dn, x, f"{p}.{norm}", f"{tag}.ln", gout, dres=g)
        return gout

    # ---- MultiLayeredConv1d (kernel size > 1) / Conv1dLinear position-wise layers (multi_layer_conv.py:12-108)
    def _ffn_is_conv(self) -> bool:
        t = self.hp["positionwise_layer_type"]
        return t == "conv1d-linear" or (t == "conv1d" and self.hp.get("positionwise_conv_kernel_size", 1) > 1)

    def _ffn_conv_fwd(self, x, n, p, ff, tag, U, rate, out, scale: float = 0.5):
        """out = x + scale * dropout(w_2(dropout(relu(w_1 n)))) (scale 0.5: conformer block, 1: Transformer encoder layer) with w_1 a Conv1d(k) over time (zero padding (k-1)/2 per utterance
        row, padded frames take part like any other frame: the reference does not mask inside the block) and w_2 a Conv1d(k)
        ("conv1d") or a Linear ("conv1d-linear").  The convolutions are taps-GEMMs over zero-haloed channels-last rows."""
        st = self.store
        B, T, dm = x.shape
        halo = (self.hp["positionwise_conv_kernel_size"] - 1) // 2
        Lp = T + 2 * halo
        npad = self.buf(tag + ".npad", (B, Lp, dm))
        ops.pad_rows(n, npad, halo)
        h = self._conv1d_fwd(npad, f"{p}.{ff}.w_1", T, True, tag + ".c1")             # (B, Lp, U), zero halos
        d1 = self.named_drop(tag + ".d1", rate)
        if d1.p > 0.0:
            h = ops.scale_dropout(h, self.buf(tag + ".hd", (B, Lp, U)), 1.0, d1)
        y = self._scratch("cf.y", (B, T, dm))
        if self.hp["positionwise_layer_type"] == "conv1d":
            z2 = self._conv1d_fwd(h, f"{p}.{ff}.w_2", T, False, tag + ".c2")
            ops.unpad_rows(z2, y, halo)
        else:
            hu = self.buf(tag + ".hu", (B, T, U))
            ops.unpad_rows(h, hu, halo)
            self._lin_fwd(hu.view(B * T, U), self.W(f"{p}.{ff}.w_2.weight"), st.p(f"{p}.{ff}.w_2.bias"), y.view(B * T, dm))
        ops.scale_dropout(y, y, scale, self.named_drop(tag + ".d2", rate))
        ops.add(x, y, out)
        return out

    def _ffn_conv_bwd(self, g, x, n, p, ff, tag, U, rate, gout):
        dn = self._ffn_conv_bwd_core(g, n, p, ff, tag, U, rate)
        norm = "norm_ff_macaron" if ff == "feed_forward_macaron" else "norm_ff"
        self._ln_bwd(dn, x, f"{p}.{norm}", f"{tag}.ln", gout, dres=g)
        return gout

    def _ffn_conv_bwd_core(self, g, n, p, ff, tag, U, rate, scale: float = 0.5, dn=None):
        """g = d(out) -> d(n) (gradient at the LayerNorm output that feeds w_1); accumulates the parameter gradients."""
        st = self.store
        B, T, dm = n.shape
        halo = (self.hp["positionwise_conv_kernel_size"] - 1) // 2
        Lp = T + 2 * halo
        d1 = self.named_drop(tag + ".d1", rate)
        h = self.buf(tag + (".hd" if d1.p > 0.0 else ".c1.z"), (B, Lp, U))            # dropout(relu(.)): zero where cut or dropped
        dy = self._scratch("cf.dy", (B, T, dm))
        ops.scale_dropout(g, dy, scale, self.named_drop(tag + ".d2", rate))
        dh = self._scratch("cf.dhp", (B, Lp, U))
        if self.hp["positionwise_layer_type"] == "conv1d":
            dz2 = self._scratch("cf.dz2", (B, Lp, dm))
            ops.pad_rows(dy, dz2, halo)
            self._conv1d_bwd(dz2, h, f"{p}.{ff}.w_2", T, tag + ".c2", dh)
        else:
            hu = self.buf(tag + ".hu", (B, T, U))
            dhu = self._scratch("cf.dhu", (B, T, U))
            self._lin_bwd(dy.view(B * T, dm), hu.view(B * T, U), self.W(f"{p}.{ff}.w_2.weight"), st.g(f"{p}.{ff}.w_2.weight"),
                          st.g(f"{p}.{ff}.w_2.bias"), dx=dhu.view(B * T, U))
            ops.pad_rows(dhu, dh, halo)
        ops.relu_bwd(dh, h, dh, d1.scale)
        npad = self.buf(tag + ".npad", (B, Lp, dm))
        dnp = self._scratch("cf.dnp", (B, Lp, dm))
        self._conv1d_bwd(dh, npad, f"{p}.{ff}.w_1", T, tag + ".c1", dnp)
        if dn is None:
            dn = self._scratch("cf.dn", (B, T, dm))
        ops.unpad_rows(dnp, dn, halo)
        return dn

    def _relattn_fwd(self, x, p, tag, H, klens, pos_emb, rate, attn_rate, out):
        """out = x + dropout(RelPositionMultiHeadedAttention(LN(x)))   (encoder_layer.py:125-150, attention.py:262-305)."""
        st = self.store
        B, T, dm = x.shape
        dk = dm // H
        NP = self._n_pos(T)
        ld, ldb = _r8(T), _r8(NP)
        sc = 1.0 / math.sqrt(dk)
        n = self._ln_fwd(x, p + ".norm_mha", tag + ".ln")
        qkv = self.buf(tag + ".qkv", (B, T, 3, H, dk))
        self._lin_fwd(n.view(B * T, dm), self.Wspan([p + ".self_attn.linear_q.weight"], (3 * dm, dm)),
                      st.span(st.P, [p + ".self_attn.linear_q.bias"], (3 * dm,)), qkv.view(B * T, 3 * dm))
        qu = self.buf(tag + ".qu", (B, T, H, dk))
        qv = self.buf(tag + ".qv", (B, T, H, dk))
        ops.bias_add2(qkv.view(B * T, 3 * dm)[:, :dm], st.p(p + ".self_attn.pos_bias_u"), st.p(p + ".self_attn.pos_bias_v"), qu, qv)
        pp = self.buf(tag + ".pp", (NP, H, dk))
        self._lin_fwd(pos_emb, self.W(p + ".self_attn.linear_pos.weight"), None, pp.view(NP, dm))
        k, v = qkv[:, :, 1], qkv[:, :, 2]
        P = self.buf(tag + ".P", (B, H, T, ld))
        ops.gemm(qu.permute(0, 2, 1, 3), k.permute(0, 2, 1, 3), P[..., :T], alpha=sc, mode=self.mode)        # matrix_ac
        BD = self._scratch("cf.bd", (H, B * T, ldb))
        ops.gemm(qv.view(B * T, H, dk).permute(1, 0, 2), pp.permute(1, 0, 2), BD[..., :NP], alpha=sc, mode=self.mode)   # matrix_bd
        (ops.relshift_legacy_add if self._legacy_rel() else ops.relshift_add)(P, BD.view(H, B, T, ldb), T)
        drop = self.named_drop(tag + ".attn", attn_rate)
        Pd = self.buf(tag + ".Pd", (B, H, T, ld)) if drop.p > 0 else None
        ops.softmax_fwd(P, klens, False, T, Pd, drop)
        self.attn[p + ".self_attn"] = P[..., :T]
        Pv = Pd if Pd is not None else P
        ctx = self.buf(tag + ".ctx", (B, T, dm))
        ops.gemm(Pv[..., :T], v.permute(0, 2, 3, 1), ctx.view(B, T, H, dk).permute(0, 2, 1, 3), mode=self.mode)
        self._lin_fwd(ctx.view(B * T, dm), self.W(p + ".self_attn.linear_out.weight"), st.p(p + ".self_attn.linear_out.bias"),
                      out.view(B * T, dm), drop=self.named_drop(tag + ".out", rate), residual=x.view(B * T, dm))
        return out

    def _relattn_bwd(self, g, x, p, tag, H, pos_emb, rate, attn_rate, gout):
        st = self.store
        B, T, dm = x.shape
        dk = dm // H
        NP = self._n_pos(T)
        ld, ldb = _r8(T), _r8(NP)
        sc = 1.0 / math.sqrt(dk)
        mode = self.mode
        n = self.buf(tag + ".ln.y", (B, T, dm))
        qkv = self.buf(tag + ".qkv", (B, T, 3, H, dk))
        qu = self.buf(tag + ".qu", (B, T, H, dk))
        qv = self.buf(tag + ".qv", (B, T, H, dk))
        pp = self.buf(tag + ".pp", (NP, H, dk))
        P = self.buf(tag + ".P", (B, H, T, ld))
        drop = self.named_drop(tag + ".attn", attn_rate)
        Pv = self.buf(tag + ".Pd", (B, H, T, ld)) if drop.p > 0 else P
        ctx = self.buf(tag + ".ctx", (B, T, dm))
        k, v = qkv[:, :, 1], qkv[:, :, 2]
        do = self._drop_bwd(g.view(B * T, dm), self.named_drop(tag + ".out", rate), self._scratch("cf.dy", (B * T, dm)))
        dctx = self._scratch("cf.dctx", (B, T, dm))
        self._lin_bwd(do, ctx.view(B * T, dm), self.W(p + ".self_attn.linear_out.weight"), st.g(p + ".self_attn.linear_out.weight"),
                      st.g(p + ".self_attn.linear_out.bias"), dx=dctx.view(B * T, dm))
        dqkv = self._scratch("cf.dqkv", (B, T, 3, H, dk))
        dk_, dv = dqkv[:, :, 1], dqkv[:, :, 2]
        dP = self._scratch("cf.dP", (B, H, T, ld))
        dctx4 = dctx.view(B, T, H, dk).permute(0, 2, 1, 3)
        ops.gemm(dctx4, v.permute(0, 2, 1, 3), dP[..., :T], mode=mode)
        ops.gemm(Pv[..., :T].transpose(-1, -2), dctx4.transpose(-1, -2), dv.permute(0, 2, 1, 3), mode=mode)
        ops.softmax_bwd(P, dP, T, sc, drop)
        dS = dP                                                                  # gradient w.r.t. the un-scaled ac and bd'
        dBD = self._scratch("cf.bd", (H, B * T, ldb))
        (ops.relshift_legacy_bwd if self._legacy_rel() else ops.relshift_bwd)(dS, dBD.view(H, B, T, ldb), T)
        dqu = self._scratch("cf.dqu", (B, T, H, dk))
        dqv = self._scratch("cf.dqv", (B, T, H, dk))
        ops.gemm(dS[..., :T], k.permute(0, 2, 3, 1), dqu.permute(0, 2, 1, 3), mode=mode)
        ops.gemm(dS[..., :T].transpose(-1, -2), qu.permute(0, 2, 3, 1), dk_.permute(0, 2, 1, 3), mode=mode)
        ops.gemm(dBD[..., :NP], pp.permute(1, 2, 0), dqv.view(B * T, H, dk).permute(1, 0, 2), mode=mode)
        dpp = self._scratch("cf.dpp", (NP, H, dk))
        ops.gemm(dBD[..., :NP].transpose(-1, -2), qv.view(B * T, H, dk).permute(1, 2, 0), dpp.permute(1, 0, 2), mode=mode)
        self._lin_bwd(dpp.view(NP, dm), pos_emb, self.W(p + ".self_attn.linear_pos.weight"), st.g(p + ".self_attn.linear_pos.weight"),
                      None, dx=None)
        ops.colsum(dqu.view(B * T, dm), st.g(p + ".self_attn.pos_bias_u").view(dm))
        ops.colsum(dqv.view(B * T, dm), st.g(p + ".self_attn.pos_bias_v").view(dm))
        ops.add_strided(dqu.view(B * T, dm), dqv.view(B * T, dm), dqkv.view(B * T, 3 * dm)[:, :dm])
        dn = self._scratch("cf.dn", (B, T, dm))
        self._lin_bwd(dqkv.view(B * T, 3 * dm), n.view(B * T, dm), self.Wspan([p + ".self_attn.linear_q.weight"], (3 * dm, dm)),
                      st.span(st.G, [p + ".self_attn.linear_q.weight"], (3 * dm, dm)), st.span(st.G, [p + ".self_attn.linear_q.bias"], (3 * dm,)),
                      dx=dn.view(B * T, dm))
        self._ln_bwd(dn, x, p + ".norm_mha", tag + ".ln", gout, dres=g)
        return gout

    def _convmod_fwd(self, x, p, tag, K, rate, out):
        """out = x + dropout(ConvolutionModule(LN(x)))   (encoder_layer.py:152-158, convolution.py:56-79)."""
        st = self.store
        B, T, dm = x.shape
        cm = p + ".conv_module"
        n = self._ln_fwd(x, p + ".norm_conv", tag + ".ln")
        pw1 = self.buf(tag + ".pw1", (B * T, 2 * dm))
        self._lin_fwd(n.view(B * T, dm), self.W(cm + ".pointwise_conv1.weight").view(2 * dm, dm), st.p(cm + ".pointwise_conv1.bias"), pw1)
        glu = self.buf(tag + ".glu", (B, T, dm))
        ops.glu_fwd(pw1, glu)
        z = self.buf(tag + ".z", (B, T, dm))
        ops.dwconv_fwd(glu, st.p(cm + ".depthwise_conv.weight").view(dm, K), st.p(cm + ".depthwise_conv.bias"), z)
        mean = self.buf(tag + ".mean", (dm,), _f32)
        invstd = self.buf(tag + ".invstd", (dm,), _f32)
        if self.training:
            sums = self.buf(tag + ".sums", (2 * dm,), _f32)
            sums.zero_()
            ops.bn_stats(z, sums, T, 0)
            ops.bn_finalize(sums, mean, invstd, self.buffers[cm + ".norm.running_mean"], self.buffers[cm + ".norm.running_var"], B * T)
            self.buffers[cm + ".norm.num_batches_tracked"] += 1
        else:
            ops.bn_eval_stats(self.buffers[cm + ".norm.running_mean"], self.buffers[cm + ".norm.running_var"], mean, invstd)
        y = self.buf(tag + ".y", (B, T, dm))
        ops.bn_apply(z, mean, invstd, st.p(cm + ".norm.weight"), st.p(cm + ".norm.bias"), y, T, 0, 2)
        self._lin_fwd(y.view(B * T, dm), self.W(cm + ".pointwise_conv2.weight").view(dm, dm), st.p(cm + ".pointwise_conv2.bias"),
                      out.view(B * T, dm), drop=self.named_drop(tag + ".out", rate), residual=x.view(B * T, dm))
        return out

    def _convmod_bwd(self, g, x, p, tag, K, rate, gout):
        st = self.store
        B, T, dm = x.shape
        cm = p + ".conv_module"
        n = self.buf(tag + ".ln.y", (B, T, dm))
        pw1 = self.buf(tag + ".pw1", (B * T, 2 * dm))
        glu = self.buf(tag + ".glu", (B, T, dm))
        z = self.buf(tag + ".z", (B, T, dm))
        y = self.buf(tag + ".y", (B, T, dm))
        mean = self.buf(tag + ".mean", (dm,), _f32)
        invstd = self.buf(tag + ".invstd", (dm,), _f32)
        do = self._drop_bwd(g.view(B * T, dm), self.named_drop(tag + ".out", rate), self._scratch("cf.dy", (B * T, dm)))
        dy = self._scratch("cf.dctx", (B, T, dm))
        self._lin_bwd(do, y.view(B * T, dm), self.W(cm + ".pointwise_conv2.weight").view(dm, dm),
                      st.g(cm + ".pointwise_conv2.weight").view(dm, dm), st.g(cm + ".pointwise_conv2.bias"), dx=dy.view(B * T, dm))
        gam, bet = st.p(cm + ".norm.weight"), st.p(cm + ".norm.bias")
        sums = self._scratch("cf.bsums", (2 * dm,), _f32)
        sums.zero_()
        dz = self._scratch("cf.dz", (B, T, dm))
        ops.bn_bwd_reduce(dy, y, z, mean, invstd, gam, bet, sums, T, 0, 2)
        if self.training:
            ops.bn_bwd_apply(dy, y, z, mean, invstd, gam, bet, sums, dz, st.g(cm + ".norm.weight"), st.g(cm + ".norm.bias"), T, 0, 2)
        else:
            ops.bn_bwd_apply(dy, y, z, mean, invstd, gam, bet, None, dz, None, None, T, 0, 2)
            ops.add(st.g(cm + ".norm.bias"), sums[:dm], st.g(cm + ".norm.bias"))
            ops.add(st.g(cm + ".norm.weight"), sums[dm:], st.g(cm + ".norm.weight"))
        dglu = self._scratch("cf.dglu", (B, T, dm))
        ops.dwconv_bwd(dz, glu, st.p(cm + ".depthwise_conv.weight").view(dm, K), dglu, st.g(cm + ".depthwise_conv.weight").view(dm, K),
                       st.g(cm + ".depthwise_conv.bias"))
        dpw1 = self._scratch("cf.dpw1", (B * T, 2 * dm))
        ops.glu_bwd(dglu.view(B * T, dm), pw1, dpw1)
        dn = self._scratch("cf.dn", (B, T, dm))
        self._lin_bwd(dpw1, n.view(B * T, dm), self.W(cm + ".pointwise_conv1.weight").view(2 * dm, dm),
                      st.g(cm + ".pointwise_conv1.weight").view(2 * dm, dm), st.g(cm + ".pointwise_conv1.bias"), dx=dn.view(B * T, dm))
        self._ln_bwd(dn, x, p + ".norm_conv", tag + ".ln", gout, dres=g)
        return gout

    def _conformer_fwd(self, x, prefix, n_layers, H, U, K, klens, rate, pos_rate, attn_rate):
        """RelPositionalEncoding dropout of pos_emb + n conformer blocks + after_norm (conformer/encoder.py:249-293)."""
        B, T, dm = x.shape
        NP = self._n_pos(T)
        zero = self._scratch("cf.zero_pe", (1, NP, dm))
        zero.zero_()
        pos_emb = self.buf(prefix + ".pos_emb", (NP, dm))
        one = self.__dict__.get("_one")
        if one is None:
            one = self._one = torch.ones(1, dtype=_f32, device=self.device)
        ops.scaled_pe_fwd(zero, self._rel_table(T, dm), one, pos_emb.view(1, NP, dm), self.named_drop(prefix + ".posemb", pos_rate))
        for l in range(n_layers):
            p = f"{prefix}.encoders.{l}"
            x1 = self._ffn_fwd(x, p, "feed_forward_macaron", p + ".mac", U, rate, self.buf(p + ".x1", (B, T, dm)))
            x2 = self._relattn_fwd(x1, p, p + ".sa", H, klens, pos_emb, rate, attn_rate, self.buf(p + ".x2", (B, T, dm)))
            x3 = self._convmod_fwd(x2, p, p + ".cv", K, rate, self.buf(p + ".x3", (B, T, dm)))
            x4 = self._ffn_fwd(x3, p, "feed_forward", p + ".ff", U, rate, self.buf(p + ".x4", (B, T, dm)))
            x = self._ln_fwd(x4, p + ".norm_final", p + ".lnz")
        self.__dict__.setdefault("_last", {})[prefix] = x
        return self._ln_fwd(x, prefix + ".after_norm", prefix + ".after")

    def _conformer_bwd(self, g, x0, prefix, n_layers, H, U, K, rate, pos_rate, attn_rate):
        """g = d(after_norm output) -> returns d(x0) (input of the first block, after the positional scaling)."""
        B, T, dm = x0.shape
        pos_emb = self.buf(prefix + ".pos_emb", (self._n_pos(T), dm))
        ga = self._scratch(prefix + ".ga", (B, T, dm))
        gb = self._scratch(prefix + ".gb", (B, T, dm))
        self._ln_bwd(g, self._last[prefix], prefix + ".after_norm", prefix + ".after", ga)
        cur, other = ga, gb
        for l in reversed(range(n_layers)):
            p = f"{prefix}.encoders.{l}"
            xin = self.buf(f"{prefix}.encoders.{l - 1}.lnz.y", (B, T, dm)) if l > 0 else x0
            x1, x2, x3, x4 = (self.buf(p + f".x{i}", (B, T, dm)) for i in (1, 2, 3, 4))
            self._ln_bwd(cur, x4, p + ".norm_final", p + ".lnz", other)
            cur, other = other, cur
            self._ffn_bwd(cur, x3, p, "feed_forward", p + ".ff", U, rate, other)
            cur, other = other, cur
            self._convmod_bwd(cur, x2, p, p + ".cv", K, rate, other)
            cur, other = other, cur
            self._relattn_bwd(cur, x1, p, p + ".sa", H, pos_emb, rate, attn_rate, other)
            cur, other = other, cur
            self._ffn_bwd(cur, xin, p, "feed_forward_macaron", p + ".mac", U, rate, other)
            cur, other = other, cur
        return cur
