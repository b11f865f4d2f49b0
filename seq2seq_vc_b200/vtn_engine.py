"""VTN training step on B200: explicit forward / backward over the C-ABI kernels.

This is the hot path behind ``seq2seq_vc_b200.VTN`` (reference: seq2seq_vc/models/vtn.py:207-300,
modules/transformer/{encoder,decoder,encoder_layer,decoder_layer,attention,subsampling}.py,
modules/pre_postnets.py, losses/seq2seq_loss.py:30-59, trainers/ar_vc.py:99-107).  There is no
autograd tape: every activation the backward pass needs lives in a named, shape-cached HBM buffer,
the backward is written out op by op, and dropout masks are regenerated from a counter-based RNG,
so one step is a fixed sequence of kernel launches that can be captured in a CUDA graph.

PyTorch is used only to own device memory; all arithmetic happens in libs2svc_b200.so.

Data layout in HBM
  activations   (B, T, d) row-major ("tokens x channels"), dtype f32 (parity mode) or bf16
  attention     scores / probabilities (B, H, T1, ld) with ld = T2 rounded up to 8
  conv buffers  channels-last; postnet buffers carry a zero halo of (k-1)/2 frames per utterance
                so that Conv1d is a `taps`-GEMM over overlapping row windows (no im2col)
  parameters    one flat float32 buffer (ParamStore.P) + flat grads / Adam moments (+ bf16 shadow)
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib, ops
from ._lib import NO_DROP, Drop
from .engine_base import EngineBase, _r8, sinusoid_table
from .conformer_blocks import ConformerBlocks, conformer_buffer_specs, conformer_param_groups
from .params import ParamStore

_f32 = torch.float32
_i32 = torch.int32


def default_hparams(**over) -> dict:
    """Constructor defaults of the reference VTN (seq2seq_vc/models/vtn.py:15-62)."""
    hp = dict(idim=80, odim=80, dprenet_layers=2, dprenet_units=256, adim=384, aheads=4, elayers=6, eunits=1536,
              dlayers=6, dunits=1536, postnet_layers=5, postnet_filts=5, postnet_chans=256,
              dprenet_dropout_rate=0.5, transformer_enc_dropout_rate=0.1,
              # fixed by the reference's Encoder / Decoder / Postnet defaults
              enc_positional_dropout_rate=0.1, dec_dropout_rate=0.1, dec_positional_dropout_rate=0.1,
              postnet_dropout_rate=0.5, decoder_reduction_factor=2,
              initial_encoder_alpha=1.0, initial_decoder_alpha=1.0,
              # "conv2d": VTN (Conv2dSubsampling + ScaledPE); "embed": TransformerTTS (token embedding + <eos> + ScaledPE)
              encoder_input="conv2d",
              # "transformer" | "conformer" (models/vtn.py:83-140: Conv2dSubsampling + rel-pos encoding, macaron conformer blocks with
              # the convolution module; the decoder stays a Transformer decoder).  conformer_rel_pos_type "legacy" is the class default
              # (LegacyRelPositionalEncoding + LegacyRelPositionMultiHeadedAttention), "latest" the 2T-1 form AAS-VC uses
              encoder_type="transformer", conformer_rel_pos_type="legacy", conformer_enc_kernel_size=7,
              enc_attn_dropout_rate=0.1, positionwise_layer_type="linear", positionwise_conv_kernel_size=1)
    hp.update(over)
    if hp["encoder_type"] not in ("transformer", "conformer"):
        raise NotImplementedError(f"encoder_type {hp['encoder_type']!r}")
    if hp["encoder_type"] == "conformer" and hp["encoder_input"] != "conv2d":
        raise NotImplementedError("the conformer encoder takes the Conv2dSubsampling input layer (models/vtn.py:129)")
    if hp["positionwise_layer_type"] not in ("linear", "conv1d", "conv1d-linear"):
        raise NotImplementedError("Support only linear or conv1d.")
    if hp["positionwise_layer_type"] != "linear" and (hp["positionwise_conv_kernel_size"] % 2 != 1 or hp["encoder_input"] == "embed"):
        raise NotImplementedError("conv position-wise layers: odd kernel sizes, VTN encoders only (TransformerTTS builds none)")
    return hp


def param_groups(hp: dict) -> List[List[Tuple[str, Tuple[int, ...]]]]:
    """Reference state-dict names / shapes, grouped so that fused operands are contiguous."""
    d, idim, odim = hp["adim"], hp["idim"], hp["odim"]
    f2 = ((idim - 1) // 2 - 1) // 2
    g: List[List[Tuple[str, Tuple[int, ...]]]] = []

    def lin(name, o, i):
        g.append([(name + ".weight", (o, i))])
        g.append([(name + ".bias", (o,))])

    def ln(name, n):
        g.append([(name + ".weight", (n,))])
        g.append([(name + ".bias", (n,))])

    def mha(name, fuse):
        g.append([(f"{name}.{s}.weight", (d, d)) for s in fuse])
        g.append([(f"{name}.{s}.bias", (d,)) for s in fuse])
        for s in ("linear_q", "linear_k", "linear_v", "linear_out"):
            if s not in fuse:
                lin(f"{name}.{s}", d, d)

    if hp.get("encoder_input", "conv2d") == "embed":      # models/transformer_tts.py:63-77
        g.append([("encoder.embed.0.weight", (idim, d))])
        g.append([("encoder.embed.1.alpha", ())])
    else:
        g.append([("encoder.embed.conv.0.weight", (d, 1, 3, 3))])
        g.append([("encoder.embed.conv.0.bias", (d,))])
        g.append([("encoder.embed.conv.2.weight", (d, d, 3, 3))])
        g.append([("encoder.embed.conv.2.bias", (d,))])
        lin("encoder.embed.out.0", d, d * f2)
        if hp.get("encoder_type", "transformer") != "conformer":      # the conformer's rel-pos encoding has no alpha
            g.append([("encoder.embed.out.1.alpha", ())])
    if hp.get("encoder_type", "transformer") == "conformer":
        conformer_param_groups(g, hp, "encoder", hp["elayers"], d, hp["eunits"], hp["conformer_enc_kernel_size"], hp["aheads"])
    else:
        for l in range(hp["elayers"]):
            p = f"encoder.encoders.{l}"
            mha(p + ".self_attn", ("linear_q", "linear_k", "linear_v"))
            if hp.get("positionwise_layer_type", "linear") != "linear":     # MultiLayeredConv1d / Conv1dLinear (multi_layer_conv.py:12-108)
                pk = hp.get("positionwise_conv_kernel_size", 1)
                g.append([(p + ".feed_forward.w_1.weight", (hp["eunits"], d, pk))])
                g.append([(p + ".feed_forward.w_1.bias", (hp["eunits"],))])
                g.append([(p + ".feed_forward.w_2.weight", (d, hp["eunits"], pk) if hp["positionwise_layer_type"] == "conv1d" else (d, hp["eunits"]))])
                g.append([(p + ".feed_forward.w_2.bias", (d,))])
            else:
                lin(p + ".feed_forward.w_1", hp["eunits"], d)
                lin(p + ".feed_forward.w_2", d, hp["eunits"])
            ln(p + ".norm1", d)
            ln(p + ".norm2", d)
        ln("encoder.after_norm", d)
    u = hp["dprenet_units"]
    for i in range(hp["dprenet_layers"]):
        lin(f"decoder.embed.0.0.prenet.{i}.0", u, odim if i == 0 else u)
    lin("decoder.embed.0.1", d, u)
    g.append([("decoder.embed.1.alpha", ())])
    for l in range(hp["dlayers"]):
        p = f"decoder.decoders.{l}"
        mha(p + ".self_attn", ("linear_q", "linear_k", "linear_v"))
        mha(p + ".src_attn", ("linear_k", "linear_v"))
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
        g.append([(f"postnet.postnet.{i}.0.weight", (oc, ic, k))])
        ln(f"postnet.postnet.{i}.1", oc)
    return g


def buffer_specs(hp: dict) -> List[Tuple[str, Tuple[int, ...], torch.dtype]]:
    """Non-trainable state-dict entries (BatchNorm running statistics)."""
    out = []
    if hp.get("encoder_type", "transformer") == "conformer":
        out += conformer_buffer_specs("encoder", hp["elayers"], hp["adim"])
    ch, odim = hp["postnet_chans"], hp["odim"]
    for i in range(hp["postnet_layers"]):
        oc = odim if i == hp["postnet_layers"] - 1 else ch
        p = f"postnet.postnet.{i}.1"
        out += [(p + ".running_mean", (oc,), _f32), (p + ".running_var", (oc,), _f32),
                (p + ".num_batches_tracked", (), torch.int64)]
    return out


class VTNEngine(ConformerBlocks, EngineBase):
    """Owns parameters, activation buffers and the explicit forward/backward of one VTN step."""

    def __init__(self, hp: dict, device="cuda:0", bf16: bool = False, seed: int = 0, fp32_gemm: str = "tc"):
        """fp32_gemm (float32 engines only): "tc" = fp32-accurate tcgen05 GEMM (bf16-split operands), "simt" = CUDA-core GEMM."""
        self.fp32_gemm = fp32_gemm
        self.hp = default_hparams(**hp)
        hp = self.hp
        assert hp["adim"] % hp["aheads"] == 0
        self._setup(param_groups(hp), buffer_specs(hp), device, bf16, seed)
        self.losses = torch.zeros(2, dtype=_f32, device=self.device)
        self._loss_ws = torch.zeros(4, dtype=_f32, device=self.device)
        self.init_parameters(seed)

    # ------------------------------------------------------------------ parameters
    def init_parameters(self, seed: int = 0) -> None:
        """torch-default Linear / Conv init distributions (uniform +-1/sqrt(fan_in)); LN/BN affine = 1/0."""
        g = torch.Generator().manual_seed(seed)
        hp = self.hp
        for name, (off, shape) in self.store.offsets.items():
            n = 1
            for s in shape:
                n *= s
            if name.endswith("alpha"):
                v = torch.full((1,), hp["initial_encoder_alpha"] if name.startswith("encoder") else hp["initial_decoder_alpha"])
            elif name == "encoder.embed.0.weight" and hp["encoder_input"] == "embed":
                v = torch.randn(shape, generator=g)
                v[0] = 0.0                                  # nn.Embedding(padding_idx=0)
                v = v.reshape(-1)
            elif "norm" in name or (name.startswith("postnet") and ".1." in name):
                v = torch.ones(n) if name.endswith("weight") else torch.zeros(n)
            else:
                wname = name[:-5] + ".weight" if name.endswith(".bias") else name
                wshape = self.store.offsets[wname][1]
                fan_in = 1
                for s in wshape[1:]:
                    fan_in *= s
                v = (torch.rand(n, generator=g) * 2 - 1) / math.sqrt(fan_in)
            self.store.P[off:off + n].copy_(v.to(self.device))
        self.p16_dirty = True


    # ------------------------------------------------------------------ buffers


    # ------------------------------------------------------------------ building blocks


    # ------------------------------------------------------------------ encoder / decoder stacks
    def _encode(self, xs: torch.Tensor) -> torch.Tensor:
        """Encoder front end + encoder layers + after_norm -> memory (B, T2, d).  Uses self.shapes / self.klens_enc."""
        hp, st = self.hp, self.store
        embed = hp["encoder_input"] == "embed"
        sh = self.shapes
        B, T, T1, F1, T2, F2 = sh["B"], sh["T"], sh["T1"], sh["F1"], sh["T2"], sh["F2"]
        idim, d, H = hp["idim"], hp["adim"], hp["aheads"]
        dk = d // H
        self.xs = xs
        x = self.buf("enc.x0", (B, T2, d))
        if embed:
            # ---- token embedding + <eos> + ScaledPE in one kernel (transformer_tts.py:63-77,139-142)
            ops.embed_pe_fwd(xs, self.ilens_dev, st.p("encoder.embed.0.weight"), self.pe(d, T2), st.p("encoder.embed.1.alpha"), x,
                             idim - 1, 0, self.drop(hp["enc_positional_dropout_rate"]))
        else:
            # ---- packed conv weights (activation dtype)
            w2p = self.buf("w.conv2p", (d, 9, d))          # [oc][tap][ic]
            ops.transpose_last2(st.p("encoder.embed.conv.2.weight"), w2p, d, d, 9)
            woutp = self.buf("w.outp", (d, F2, d))         # [n][f][c]
            ops.transpose_last2(st.p("encoder.embed.out.0.weight"), woutp, d, d, F2)
            # ---- encoder front end (subsampling.py:74-94)
            y1 = self.buf("enc.y1", (B, T1, F1, d))
            self._conv1_fwd(xs, "encoder.embed", y1, "enc")
            col = self._scratch("col", (B * T2 * F2, 9 * d))
            ops.im2col_s2(y1, col)
            self._col_of = self._sig                       # the patch matrix stays valid until backward() turns it into dcol
            y2 = self.buf("enc.y2", (B * T2 * F2, d))
            ops.gemm(col, w2p.view(d, 9 * d), y2, bias=st.p("encoder.embed.conv.2.bias"), relu=True, mode=self.mode)
            elin = self.buf("enc.elin", (B * T2, d))
            ops.gemm(y2.view(B * T2, F2 * d), woutp.view(d, F2 * d), elin, bias=st.p("encoder.embed.out.0.bias"), mode=self.mode)
            if hp["encoder_type"] == "conformer":
                # rel-pos encoding of the input layer: x * sqrt(d) then the positional dropout (positional_encoding.py:192-235 /
                # :263-309); conformer blocks + after_norm (conformer/encoder.py:249-293)
                ops.scale_dropout(elin.view(B, T2, d), x, math.sqrt(d), NO_DROP, self.named_drop("enc.posx", hp["enc_positional_dropout_rate"]))
                return self._conformer_fwd(x, "encoder", hp["elayers"], H, hp["eunits"], hp["conformer_enc_kernel_size"], self.klens_enc,
                                           hp["transformer_enc_dropout_rate"], hp["enc_positional_dropout_rate"], hp["enc_attn_dropout_rate"])
            ops.scaled_pe_fwd(elin.view(B, T2, d), self.pe(d, T2), st.p("encoder.embed.out.1.alpha"), x,
                              self.drop(hp["enc_positional_dropout_rate"]))

        # ---- encoder layers (pre-LN; encoder_layer.py:61-119)
        pe_ = hp["transformer_enc_dropout_rate"]
        for l in range(hp["elayers"]):
            p = f"encoder.encoders.{l}"
            n1 = self._ln_fwd(x, p + ".norm1", p + ".ln1")
            qkv = self.buf(p + ".qkv", (B, T2, 3, H, dk))
            self._lin_fwd(n1.view(B * T2, d), self.Wspan([p + ".self_attn.linear_q.weight"], (3 * d, d)),
                          st.span(st.P, [p + ".self_attn.linear_q.bias"], (3 * d,)), qkv.view(B * T2, 3 * d))
            ctx = self._attn_core_fwd(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], self.klens_enc, False, p + ".sa", p + ".self_attn")
            xm = self.buf(p + ".xmid", (B, T2, d))
            self._lin_fwd(ctx.view(B * T2, d), self.W(p + ".self_attn.linear_out.weight"), st.p(p + ".self_attn.linear_out.bias"),
                          xm.view(B * T2, d), drop=self.drop(pe_), residual=x.view(B * T2, d))
            n2 = self._ln_fwd(xm, p + ".norm2", p + ".ln2")
            xn = self.buf(p + ".xout", (B, T2, d))
            if hp["positionwise_layer_type"] != "linear":       # MultiLayeredConv1d / Conv1dLinear: name-keyed dropout sites
                self._ffn_conv_fwd(xm, n2, p, "feed_forward", p + ".ffc", hp["eunits"], pe_, xn, scale=1.0)
            else:
                h = self.buf(p + ".ffh", (B * T2, hp["eunits"]))
                self._lin_fwd(n2.view(B * T2, d), self.W(p + ".feed_forward.w_1.weight"), st.p(p + ".feed_forward.w_1.bias"), h,
                              relu=True, drop=self.drop(pe_))
                self._lin_fwd(h, self.W(p + ".feed_forward.w_2.weight"), st.p(p + ".feed_forward.w_2.bias"), xn.view(B * T2, d),
                              drop=self.drop(pe_), residual=xm.view(B * T2, d))
            x = xn
        self.enc_last = x
        mem = self._ln_fwd(x, "encoder.after_norm", "enc.after")

        return mem

    def _decode(self, ys_in: torch.Tensor, mem: torch.Tensor) -> torch.Tensor:
        """Prenet + ScaledPE + decoder layers over the (already shifted / thinned) decoder inputs ys_in (B, Lr, odim)."""
        hp, st = self.hp, self.store
        B, Lr, odim = ys_in.shape
        T2 = mem.shape[1]
        d, H = hp["adim"], hp["aheads"]
        dk = d // H
        u = hp["dprenet_units"]
        hcur = ys_in.view(B * Lr, odim)
        for i in range(hp["dprenet_layers"]):
            nm = f"decoder.embed.0.0.prenet.{i}.0"
            out = self.buf(f"dec.prenet{i}", (B * Lr, u))
            # Prenet dropout is always on in the reference (pre_postnets.py:65), also in eval()
            self._site += 1
            pd = hp["dprenet_dropout_rate"]
            dr = Drop(pd, self.base_seed, self._site, self.seed_dev) if pd > 0 else NO_DROP
            self._lin_fwd(hcur, self.W(nm + ".weight"), st.p(nm + ".bias"), out, relu=True, drop=dr)
            hcur = out
        dlin = self.buf("dec.elin", (B * Lr, d))
        self._lin_fwd(hcur, self.W("decoder.embed.0.1.weight"), st.p("decoder.embed.0.1.bias"), dlin)
        x = self.buf("dec.x0", (B, Lr, d))
        ops.scaled_pe_fwd(dlin.view(B, Lr, d), self.pe(d, Lr), st.p("decoder.embed.1.alpha"), x,
                          self.drop(hp["dec_positional_dropout_rate"]))

        # ---- decoder layers (post-LN; decoder_layer.py:63-134)
        pdrop = hp["dec_dropout_rate"]
        for l in range(hp["dlayers"]):
            p = f"decoder.decoders.{l}"
            qkv = self.buf(p + ".qkv", (B, Lr, 3, H, dk))
            self._lin_fwd(x.view(B * Lr, d), self.Wspan([p + ".self_attn.linear_q.weight"], (3 * d, d)),
                          st.span(st.P, [p + ".self_attn.linear_q.bias"], (3 * d,)), qkv.view(B * Lr, 3 * d))
            ctx = self._attn_core_fwd(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], self.olens_in, True, p + ".sa", p + ".self_attn")
            t1 = self.buf(p + ".t1", (B, Lr, d))
            self._lin_fwd(ctx.view(B * Lr, d), self.W(p + ".self_attn.linear_out.weight"), st.p(p + ".self_attn.linear_out.bias"),
                          t1.view(B * Lr, d), drop=self.drop(pdrop), residual=x.view(B * Lr, d))
            x1 = self._ln_fwd(t1, p + ".norm1", p + ".ln1")
            q = self.buf(p + ".q", (B, Lr, H, dk))
            self._lin_fwd(x1.view(B * Lr, d), self.W(p + ".src_attn.linear_q.weight"), st.p(p + ".src_attn.linear_q.bias"),
                          q.view(B * Lr, d))
            kv = self.buf(p + ".kv", (B, T2, 2, H, dk))
            self._lin_fwd(mem.view(B * T2, d), self.Wspan([p + ".src_attn.linear_k.weight"], (2 * d, d)),
                          st.span(st.P, [p + ".src_attn.linear_k.bias"], (2 * d,)), kv.view(B * T2, 2 * d))
            ctx2 = self._attn_core_fwd(q, kv[:, :, 0], kv[:, :, 1], self.klens_enc, False, p + ".ca", p + ".src_attn")
            t2 = self.buf(p + ".t2", (B, Lr, d))
            self._lin_fwd(ctx2.view(B * Lr, d), self.W(p + ".src_attn.linear_out.weight"), st.p(p + ".src_attn.linear_out.bias"),
                          t2.view(B * Lr, d), drop=self.drop(pdrop), residual=x1.view(B * Lr, d))
            x2 = self._ln_fwd(t2, p + ".norm2", p + ".ln2")
            h = self.buf(p + ".ffh", (B * Lr, hp["dunits"]))
            self._lin_fwd(x2.view(B * Lr, d), self.W(p + ".feed_forward.w_1.weight"), st.p(p + ".feed_forward.w_1.bias"), h,
                          relu=True, drop=self.drop(pdrop))
            t3 = self.buf(p + ".t3", (B, Lr, d))
            self._lin_fwd(h, self.W(p + ".feed_forward.w_2.weight"), st.p(p + ".feed_forward.w_2.bias"), t3.view(B * Lr, d),
                          drop=self.drop(pdrop), residual=x2.view(B * Lr, d))
            x = self._ln_fwd(t3, p + ".norm3", p + ".ln3")
        zs = x

        return zs

    # ------------------------------------------------------------------ forward
    def prepare(self, B: int, T: int, L: int, ilens: Sequence[int], olens: Sequence[int]) -> None:
        """Derive the per-utterance length vectors on the host (they arrive as CPU ints from the
        collater) and ship them in ONE small H2D copy; never synchronises.  Kept out of forward()
        so that a captured CUDA graph of the step only ever sees the device-resident copy."""
        hp = self.hp
        r = hp["decoder_reduction_factor"]
        embed = hp["encoder_input"] == "embed"
        T2 = T + 1 if embed else (((T - 1) // 2) - 1) // 2
        self._use_sig((B, T, L, self.training))
        ilens = [int(v) for v in ilens]
        olens = [int(v) for v in olens]
        assert len(ilens) == B and len(olens) == B
        if embed:
            klens_enc = [i + 1 for i in ilens]                       # <eos> appended (transformer_tts.py:139-142)
        else:
            klens_enc = [min(T2, (i + 3) // 4) for i in ilens]       # mask[:, :, :-2:2][:, :, :-2:2] (subsampling.py:92-94)
        olens_in = [o // r for o in olens]
        olens_fix = [o - o % r for o in olens]
        if embed:
            self.ilens_ds_st = klens_enc                                        # transformer_tts.py:222 returns ilens + 1
        else:
            self.ilens_ds_st = [((i - 2 + 1) // 2 - 2 + 1) // 2 for i in ilens]   # vtn.py:279
        self._ship_lens([klens_enc, olens_in, olens_fix, ilens, self.ilens_ds_st])
        self.olens_in_host, self.olens_fix_host = olens_in, olens_fix
        self._prepared = (B, T, L)

    def forward(self, xs: torch.Tensor, ys: torch.Tensor, ilens: Optional[Sequence[int]] = None,
                olens: Optional[Sequence[int]] = None):
        """xs (B,T,idim) / ys (B,L,odim) float32 device tensors already trimmed to max length.
        ilens / olens: host ints; omit them when prepare() was already called for this batch.

        Returns (after (B,L',odim), before, logits (B,L')) in activation dtype; attention maps in
        self.attn; L' = (L // r) * r.
        """
        hp, st = self.hp, self.store
        embed = hp["encoder_input"] == "embed"
        B, T = xs.shape[0], xs.shape[1]
        idim = hp["idim"]
        L, odim = ys.shape[1], ys.shape[2]
        r, d, H = hp["decoder_reduction_factor"], hp["adim"], hp["aheads"]
        dk = d // H
        assert ys.dtype == _f32 and xs.is_contiguous() and ys.is_contiguous()
        assert xs.dtype == (torch.int64 if embed else _f32), "xs: int64 token ids (TransformerTTS) or float32 features (VTN)"
        T1, F1 = (T - 1) // 2, (idim - 1) // 2
        T2, F2 = (T1 - 1) // 2, (F1 - 1) // 2
        if embed:
            T1 = F1 = F2 = 0
            T2 = T + 1
        Lr = L // r
        self._sig = (B, T, L, self.training)
        self._site = 0
        self.attn = {}
        self.sync_shadow()
        self.shapes = dict(B=B, T=T, L=L, T1=T1, F1=F1, T2=T2, F2=F2, Lr=Lr)

        if ilens is not None:
            self.prepare(B, T, L, ilens, olens)
        assert self._prepared == (B, T, L), "prepare(B, T, L, ilens, olens) must precede forward() for this batch shape"
        lens = self.buf("lens", (5, B), _i32)
        self.klens_enc, self.olens_in, self.olens_fix, self.ilens_dev, self.ilens_ds_dev = lens[0], lens[1], lens[2], lens[3], lens[4]

        mem = self._encode(xs)

        # ---- decoder input (vtn.py:227-243,523-527; pre_postnets.py:60-66)
        ys_in = self.buf("dec.ys_in", (B, Lr, odim))
        ops.shift_thin(ys, ys_in, r)
        zs = self._decode(ys_in, mem)

        # ---- output heads (vtn.py:249-251)
        Lo = Lr * r
        before = self.buf("out.before", (B, Lo, odim))
        self._lin_fwd(zs.view(B * Lr, d), self.W("feat_out.weight"), st.p("feat_out.bias"), before.view(B * Lr, odim * r))
        logits = self.buf("out.logits", (B, Lo))
        self._lin_fwd(zs.view(B * Lr, d), self.W("prob_out.weight"), st.p("prob_out.bias"), logits.view(B * Lr, r))
        self.zs = zs

        # ---- postnet (pre_postnets.py:105-185): Conv1d(k) as taps-GEMM over haloed channels-last rows
        after = self._postnet_fwd(before, lambda i: self.drop(hp["postnet_dropout_rate"]))
        self.before, self.after, self.logits = before, after, logits
        return after, before, logits

    # ------------------------------------------------------------------ autoregressive inference
    @torch.no_grad()
    def inference(self, x: torch.Tensor, threshold: float = 0.5, minlenratio: float = 0.0, maxlenratio: float = 10.0):
        """VTN.inference (models/vtn.py:302-394): x (T, idim) float32 -> (outs (L, odim), probs (L,), att_ws (#dlayers, H, L/r, T')).

        KV-cache decode: the encoder and the source-attention K/V projections run once; every step is ~70 single-row
        kernels (s2s_gemv, s2s_decode_attn, LayerNorm, s2s_decode_pe, s2s_decode_advance) whose only step-dependent input
        is a DEVICE position counter, so ONE captured CUDA graph is replayed for all steps.  The stop test reads the
        step's logits back to the host each step, exactly as the reference does (`int(sum(probs[-1] >= threshold))`)."""
        hp, st = self.hp, self.store
        embed = hp["encoder_input"] == "embed"          # TransformerTTS: token ids, <eos> appended by the embedding kernel (transformer_tts.py:254)
        r, d, H, odim, idim = hp["decoder_reduction_factor"], hp["adim"], hp["aheads"], hp["odim"], hp["idim"]
        dk = d // H
        T = x.shape[0]
        T1, F1 = (T - 1) // 2, (idim - 1) // 2
        T2, F2 = (T1 - 1) // 2, (F1 - 1) // 2
        if embed:
            T1 = F1 = F2 = 0
            T2 = T + 1
        assert T2 >= 1, "input too short for Conv2dSubsampling"
        nl, u, npre = hp["dlayers"], hp["dprenet_units"], hp["dprenet_layers"]
        was_training = self.training
        self.training = False
        try:
            for key in [k for k in self._bufs if isinstance(k[0], tuple) and len(k[0]) == 4 and k[0][3] is False]:
                del self._bufs[key]
            mk = lambda v: torch.tensor([v], dtype=_i32).to(self.device)
            self._sig = (1, T, 0, False)
            self._site = 0
            self.attn = {}
            self.sync_shadow()
            self.shapes = dict(B=1, T=T, L=0, T1=T1, F1=F1, T2=T2, F2=F2, Lr=0)
            self.klens_enc, self.ilens_dev = mk(T2), mk(T)            # encoder(x, None): no padding mask
            mem = self._encode((x.to(torch.int64) if embed else x.to(_f32)).contiguous().unsqueeze(0))
            site0 = self._site
            maxlen = max(int(T2 * maxlenratio / r), 1)
            minlen = int(T2 * minlenratio / r)
            cap = maxlen + 1
            ldp = _r8(T2)
            f32b = lambda name, shape: self.buf(name, shape, _f32)
            # source-attention K / V of every layer, projected once; self-attention caches
            srckv, cache = [], []
            for l in range(nl):
                p = f"decoder.decoders.{l}"
                kv = self.buf(f"inf.srckv{l}", (T2, 2, H, dk))
                self._lin_fwd(mem.view(T2, d), self.Wspan([p + ".src_attn.linear_k.weight"], (2 * d, d)),
                              st.span(st.P, [p + ".src_attn.linear_k.bias"], (2 * d,)), kv.view(T2, 2 * d))
                srckv.append(kv)
                cache.append(self.buf(f"inf.cache{l}", (cap, 2, H, dk)))
            pos = self.buf("inf.pos", (1,), _i32)
            pos.zero_()
            nxt = self.buf("inf.next", (odim,))
            nxt.zero_()
            frames, logits = f32b("inf.frames", (cap, r * odim)), f32b("inf.logits", (cap, r))
            att = f32b("inf.att", (cap, nl, H, ldp))
            vec = lambda name, n: self.buf("inf.v." + name, (n,))
            mean1, rstd1 = f32b("inf.mean", (1,)), f32b("inf.rstd", (1,))
            pd = hp["dprenet_dropout_rate"]
            sc = 1.0 / math.sqrt(dk)

            def ln(xv, name, out):
                ops.layernorm_fwd(xv.view(1, 1, d), st.p(name + ".weight"), st.p(name + ".bias"), out.view(1, 1, d), mean1, rstd1, 1e-12)
                return out

            def step():
                h = nxt
                for i in range(npre):                                   # Prenet dropout is always on (pre_postnets.py:65)
                    nm = f"decoder.embed.0.0.prenet.{i}.0"
                    dr = Drop(pd, self.base_seed, site0 + 1 + i, self.seed_dev) if pd > 0 else NO_DROP
                    h = ops.gemv(self.W(nm + ".weight"), st.p(nm + ".bias"), h, vec(f"pre{i}", u), relu=True, drop=dr, pos_dev=pos)
                e = ops.gemv(self.W("decoder.embed.0.1.weight"), st.p("decoder.embed.0.1.bias"), h, vec("emb", d))
                xv = ops.decode_pe(e, self.pe(d, cap), st.p("decoder.embed.1.alpha"), pos, vec("x0", d))
                for l in range(nl):
                    p = f"decoder.decoders.{l}"
                    qkv = ops.gemv(self.Wspan([p + ".self_attn.linear_q.weight"], (3 * d, d)), st.span(st.P, [p + ".self_attn.linear_q.bias"], (3 * d,)),
                                   xv, vec(f"qkv{l}", 3 * d))
                    ctx = ops.decode_attn(qkv[:d], qkv[d:2 * d], qkv[2 * d:], cache[l][:, 0], cache[l][:, 1], H, dk, -1, cap, pos, sc,
                                          vec(f"ctx{l}", d))
                    t1 = ops.gemv(self.W(p + ".self_attn.linear_out.weight"), st.p(p + ".self_attn.linear_out.bias"), ctx, vec(f"t1{l}", d), residual=xv)
                    x1 = ln(t1, p + ".norm1", vec(f"x1{l}", d))
                    q2 = ops.gemv(self.W(p + ".src_attn.linear_q.weight"), st.p(p + ".src_attn.linear_q.bias"), x1, vec(f"q2{l}", d))
                    ctx2 = ops.decode_attn(q2, None, None, srckv[l][:, 0], srckv[l][:, 1], H, dk, T2, T2, pos, sc, vec(f"ctx2{l}", d),
                                           probs=att[:, l], ldp=ldp, probs_step_stride=nl * H * ldp)
                    t2 = ops.gemv(self.W(p + ".src_attn.linear_out.weight"), st.p(p + ".src_attn.linear_out.bias"), ctx2, vec(f"t2{l}", d), residual=x1)
                    x2 = ln(t2, p + ".norm2", vec(f"x2{l}", d))
                    hh = ops.gemv(self.W(p + ".feed_forward.w_1.weight"), st.p(p + ".feed_forward.w_1.bias"), x2, vec(f"ffh{l}", hp["dunits"]), relu=True)
                    t3 = ops.gemv(self.W(p + ".feed_forward.w_2.weight"), st.p(p + ".feed_forward.w_2.bias"), hh, vec(f"t3{l}", d), residual=x2)
                    xv = ln(t3, p + ".norm3", vec(f"x3{l}", d))
                feat = ops.gemv(self.W("feat_out.weight"), st.p("feat_out.bias"), xv, vec("feat", odim * r))
                logit = ops.gemv(self.W("prob_out.weight"), st.p("prob_out.bias"), xv, vec("logit", r))
                ops.decode_advance(feat, logit, nxt, frames, logits, pos, odim, r)

            graph = None
            idx = 0
            while True:
                idx += 1
                if idx == 1 or self.device.type != "cuda":
                    step()                                              # eager (allocates the step's buffers on the first call)
                    if idx == 1 and self.device.type == "cuda":
                        torch.cuda.synchronize()
                        graph = torch.cuda.CUDAGraph()
                        with _lib.graph_capture(graph):
                            step()
                else:
                    graph.replay()
                p_step = torch.sigmoid(logits[idx - 1])
                stop = bool((p_step >= threshold).any().item()) or idx >= maxlen        # host read-back (vtn.py:369)
                if stop and idx >= minlen:
                    break
            L = idx * r
            before = frames[:idx].reshape(1, L, odim).to(self.adt).contiguous()
            self._sig = (1, T, -L, False)
            after = self._postnet_fwd(before, lambda i: NO_DROP)
            att_ws = att[:idx, :, :, :T2].permute(1, 2, 0, 3).contiguous()
            return after[0].float().clone(), torch.sigmoid(logits[:idx]).reshape(-1).clone(), att_ws
        finally:
            self.training = was_training

    @torch.no_grad()
    def inference_recompute(self, x: torch.Tensor, threshold: float = 0.5, minlenratio: float = 0.0, maxlenratio: float = 10.0):
        """Same contract as inference(), computed without a KV cache (kept as the cross-check of the decode kernels).

        The reference decodes with Decoder.forward_one_step, whose "cache" holds previous layer *outputs* and re-projects K/V
        of the whole prefix every step; by causality that equals running the decoder over the prefix and reading its last
        row, which is what this does (prefix lengths bucketed to multiples of 64 so that buffers / shapes are reused).
        One host read-back of the stop probabilities per step, as in the reference (`int(sum(probs[-1] >= threshold))`)."""
        hp, st = self.hp, self.store
        embed = hp["encoder_input"] == "embed"
        r, d, H, odim, idim = hp["decoder_reduction_factor"], hp["adim"], hp["aheads"], hp["odim"], hp["idim"]
        T = x.shape[0]
        T1, F1 = (T - 1) // 2, (idim - 1) // 2
        T2, F2 = (T1 - 1) // 2, (F1 - 1) // 2
        if embed:
            T1 = F1 = F2 = 0
            T2 = T + 1
        assert T2 >= 1, "input too short for Conv2dSubsampling"
        was_training = self.training
        self.training = False
        try:
            for key in [k for k in self._bufs if isinstance(k[0], tuple) and len(k[0]) == 4 and k[0][3] is False]:
                del self._bufs[key]
            mk = lambda v: torch.tensor([v], dtype=_i32).to(self.device)
            self._sig = (1, T, 0, False)
            self._site = 0
            self.attn = {}
            self.sync_shadow()
            self.shapes = dict(B=1, T=T, L=0, T1=T1, F1=F1, T2=T2, F2=F2, Lr=0)
            self.klens_enc, self.ilens_dev = mk(T2), mk(T)            # encoder(x, None): no padding mask
            xs = (x.to(torch.int64) if embed else x.to(_f32)).contiguous().unsqueeze(0)
            mem = self._encode(xs)
            site_after_encoder = self._site
            maxlen = int(T2 * maxlenratio / r)
            minlen = int(T2 * minlenratio / r)
            cap = max(maxlen, 1) + 1
            hist = torch.zeros(1, (cap + 63) // 64 * 64, odim, dtype=self.adt, device=self.device)   # decoder inputs: row 0 = zero frame
            outs, probs = [], []
            graphs: Dict[int, object] = {}
            use_graph = self.device.type == "cuda"
            idx = 0

            def step_body(Lq):
                """Decoder over the first Lq rows of `hist` + both output heads on every row (row idx-1 is read afterwards):
                a fixed launch sequence per bucket, captured once into a CUDA graph and replayed for the bucket's steps."""
                self._sig = (1, T, Lq, False)
                self._site = site_after_encoder
                ys_in = self.buf("dec.ys_in", (1, Lq, odim))
                ys_in.copy_(hist[:, :Lq])
                zs = self._decode(ys_in, mem)                          # (1, Lq, d); rows >= idx are don't-care (causal)
                feat_all = self.buf("inf.feat", (Lq, odim * r))
                logit_all = self.buf("inf.logit", (Lq, r))
                self._lin_fwd(zs.view(Lq, d), self.W("feat_out.weight"), st.p("feat_out.bias"), feat_all)
                self._lin_fwd(zs.view(Lq, d), self.W("prob_out.weight"), st.p("prob_out.bias"), logit_all)
                return feat_all, logit_all

            while True:
                idx += 1
                Lq = (idx + 63) // 64 * 64
                if Lq not in graphs:
                    self.olens_in = mk(Lq)
                    feat_all, logit_all = step_body(Lq)                # eager: allocates this bucket's buffers
                    graphs[Lq] = None
                    if use_graph:
                        torch.cuda.synchronize()
                        g = torch.cuda.CUDAGraph()
                        with _lib.graph_capture(g):
                            step_body(Lq)
                        graphs[Lq] = g
                elif graphs[Lq] is not None:
                    graphs[Lq].replay()
                else:
                    feat_all, logit_all = step_body(Lq)
                frame = feat_all[idx - 1].view(r, odim)
                outs.append(frame.clone())
                p_step = torch.sigmoid(logit_all[idx - 1].float()).view(r)
                probs.append(p_step)
                if idx < hist.shape[1]:
                    hist[0, idx].copy_(frame[-1])
                stop = bool((p_step >= threshold).any().item()) or idx >= maxlen     # host read-back (vtn.py:369)
                if stop and idx >= minlen:
                    break
            L = idx * r
            before = torch.cat(outs, dim=0).view(1, L, odim).contiguous()
            self._sig = (1, T, -L, False)
            after = self._postnet_fwd(before, lambda i: NO_DROP)
            att = torch.stack([self.attn[f"decoder.decoders.{l}.src_attn"][0, :, :idx].float() for l in range(hp["dlayers"])], dim=0)
            return after[0].float().clone(), torch.cat(probs, dim=0), att.clone()
        finally:
            self.training = was_training

    # ------------------------------------------------------------------ loss
    def loss(self, ys: torch.Tensor, labels: torch.Tensor, pos_weight: float = 10.0):
        """Seq2SeqLoss value + gradient w.r.t. the three outputs (losses/seq2seq_loss.py:30-59)."""
        B, Lo, odim = self.after.shape
        r = self.hp["decoder_reduction_factor"]
        labels_fix = self.buf("loss.labels", (B, Lo), _f32)
        olens_out = self.buf("loss.olens", (B,), _i32)
        if r > 1:
            ops.fix_targets(labels, self.olens_fix, labels_fix, olens_out, r)      # vtn.py:262-274
        else:
            labels_fix.copy_(labels[:, :Lo])
        self.labels_fix = labels_fix
        self.d_after = self.buf("loss.d_after", self.after.shape)
        self.d_before = self.buf("loss.d_before", self.after.shape)
        self.d_logits = self.buf("loss.d_logits", self.logits.shape)
        ops.seq2seq_loss(self.after, self.before, self.logits, ys, labels_fix, self.olens_fix, pos_weight, self.losses,
                         self.d_after, self.d_before, self.d_logits, self._loss_ws)
        return self.losses

    def guided_attention(self, sigma: float = 0.4, alpha: float = 1.0, n_layers: int = 2, n_heads: int = 2):
        """GuidedMultiHeadAttentionLoss over the source-attention maps of the last `n_layers` decoder layers x first
        `n_heads` heads (models/transformer_tts.py:205-219, losses/guided_attention_loss.py:142-165).
        Returns (loss (1,) device tensor, {attention name: d loss / d P in the (B, H, T_out, ld) layout})."""
        s, hp = self.shapes, self.hp
        B, Lr, T2, H = s["B"], s["Lr"], s["T2"], hp["aheads"]
        ld = _r8(T2)
        nh = min(n_heads, H)
        names = [f"decoder.decoders.{l}" for l in reversed(range(hp["dlayers"]))][:n_layers]
        total = self.buf("ga.loss", (1,), _f32)
        total.zero_()
        part = self.buf("ga.part", (1,), _f32)
        ws = self.buf("ga.ws", (2,), _f32)
        d_att = {}
        for p in names:
            P = self.buf(p + ".ca.P", (B, H, Lr, ld))
            sel = self._scratch("ga.sel", (B, nh, Lr, ld))
            sel.copy_(P[:, :nh])
            dsel = self._scratch("ga.dsel", (B, nh, Lr, ld))
            ops.guided_attn_loss(sel, self.ilens_ds_dev, self.olens_in, T2, sigma, alpha / len(names), part, dsel, ws)
            ops.add(total, part, total)
            dfull = self.buf(p + ".ca.d_att", (B, H, Lr, ld), zero=True)      # heads >= nh stay zero
            dfull[:, :nh].copy_(dsel)
            d_att[p + ".src_attn"] = dfull
        return total, d_att

    # ------------------------------------------------------------------ backward
    def backward(self, d_after: torch.Tensor, d_before: torch.Tensor, d_logits: torch.Tensor,
                 d_att: Optional[Dict[str, torch.Tensor]] = None, zero_grad: bool = True, on_decoder_done=None) -> None:
        """Accumulates parameter gradients into ParamStore.G (d_* are overwritten as scratch).
        on_decoder_done(): called once every gradient of the postnet, the output heads and the decoder is final and only the
        encoder's backward remains -- the data-parallel step starts all-reducing that part of the flat buffer there."""
        hp, st = self.hp, self.store
        s = self.shapes
        B, T, L, T1, F1, T2, F2, Lr = s["B"], s["T"], s["L"], s["T1"], s["F1"], s["T2"], s["F2"], s["Lr"]
        r, d, H, odim = hp["decoder_reduction_factor"], hp["adim"], hp["aheads"], hp["odim"]
        dk = d // H
        Lo = Lr * r
        d_att = d_att or {}
        if zero_grad:
            st.G.zero_()
        # dropout sites are re-enumerated in forward order; collect them first
        sites = self._site_table()

        # ---- postnet
        dbefore_tot = self._postnet_bwd(d_after, d_before, lambda i: sites[f"post{i}"])

        # ---- output heads
        zs = self.zs
        g = self._scratch("g.dec_a", (B, Lr, d))
        self._lin_bwd(dbefore_tot.view(B * Lr, odim * r), zs.view(B * Lr, d), self.W("feat_out.weight"), st.g("feat_out.weight"),
                      st.g("feat_out.bias"), dx=g.view(B * Lr, d))
        self._lin_bwd(d_logits.view(B * Lr, r), zs.view(B * Lr, d), self.W("prob_out.weight"), st.g("prob_out.weight"),
                      st.g("prob_out.bias"), dx=g.view(B * Lr, d), dx_accumulate=True)

        # ---- decoder layers (reverse)
        mem = self.buf("encoder.after.y" if hp["encoder_type"] == "conformer" else "enc.after.y", (B, T2, d))
        dmem = self._scratch("g.mem", (B, T2, d))
        dmem.zero_()
        for l in reversed(range(hp["dlayers"])):
            p = f"decoder.decoders.{l}"
            # weight / bias gradients of the layer are recorded and issued together at its end (EngineBase._flush_defer): the three
            # LayerNorm-input gradients and their dropout' copies are dW operands, so they get a buffer each instead of sharing one
            deferring = self._begin_defer()
            gts = [self._scratch(f"g.dec_b{j if deferring else 0}", (B, Lr, d)) for j in range(3)]
            gds = [self._scratch(f"g.dec_c{j if deferring else 0}", (B * Lr, d)) for j in range(3)]
            gt, gd = gts[0], gds[0]
            xin = self.buf(f"decoder.decoders.{l - 1}.ln3.y", (B, Lr, d)) if l > 0 else self.buf("dec.x0", (B, Lr, d))
            t1, t2, t3 = (self.buf(p + f".t{i}", (B, Lr, d)) for i in (1, 2, 3))
            x1 = self.buf(p + ".ln1.y", (B, Lr, d))
            x2 = self.buf(p + ".ln2.y", (B, Lr, d))
            h = self.buf(p + ".ffh", (B * Lr, hp["dunits"]))
            # x3 = LN3(t3), t3 = drop(h W2) + x2
            dt3 = self._ln_bwd(g, t3, p + ".norm3", p + ".ln3", gt, dx_drop=gd.view(B, Lr, d), drop=sites[p + ".ff2"])
            df = gd if sites[p + ".ff2"].p > 0.0 else dt3.view(B * Lr, d)
            dh = self._scratch("g.ffh", (B * Lr, hp["dunits"]))
            self._lin_bwd(df, h, self.W(p + ".feed_forward.w_2.weight"), st.g(p + ".feed_forward.w_2.weight"),
                          st.g(p + ".feed_forward.w_2.bias"), dx=dh, dx_gate=h, dx_gate_scale=sites[p + ".ff1"].scale)
            dx2 = g
            self._lin_bwd(dh, x2.view(B * Lr, d), self.W(p + ".feed_forward.w_1.weight"), st.g(p + ".feed_forward.w_1.weight"),
                          st.g(p + ".feed_forward.w_1.bias"), dx=dx2.view(B * Lr, d), dx_residual=dt3.view(B * Lr, d))
            # x2 = LN2(t2), t2 = drop(ctx2 Wo) + x1
            gt, gd = gts[1], gds[1]
            dt2 = self._ln_bwd(dx2, t2, p + ".norm2", p + ".ln2", gt, dx_drop=gd.view(B, Lr, d), drop=sites[p + ".ca_out"])
            do = gd if sites[p + ".ca_out"].p > 0.0 else dt2.view(B * Lr, d)
            ctx2 = self.buf(p + ".ca.ctx", (B, Lr, d))
            dctx = self._scratch("g.ctx", (B, Lr, d))
            self._lin_bwd(do, ctx2.view(B * Lr, d), self.W(p + ".src_attn.linear_out.weight"), st.g(p + ".src_attn.linear_out.weight"),
                          st.g(p + ".src_attn.linear_out.bias"), dx=dctx.view(B * Lr, d))
            q = self.buf(p + ".q", (B, Lr, H, dk))
            kv = self.buf(p + ".kv", (B, T2, 2, H, dk))
            dq = self._scratch("g.q", (B, Lr, H, dk))
            dkv = self._scratch("g.kv", (B, T2, 2, H, dk))
            self._attn_core_bwd(dctx, q, kv[:, :, 0], kv[:, :, 1], dq, dkv[:, :, 0], dkv[:, :, 1], p + ".ca", d_att.get(p + ".src_attn"),
                                klens=self.klens_enc, causal=False)
            self._lin_bwd(dkv.view(B * T2, 2 * d), mem.view(B * T2, d), self.Wspan([p + ".src_attn.linear_k.weight"], (2 * d, d)),
                          st.span(st.G, [p + ".src_attn.linear_k.weight"], (2 * d, d)), st.span(st.G, [p + ".src_attn.linear_k.bias"], (2 * d,)),
                          dx=dmem.view(B * T2, d), dx_accumulate=True)
            dx1 = g
            self._lin_bwd(dq.view(B * Lr, d), x1.view(B * Lr, d), self.W(p + ".src_attn.linear_q.weight"), st.g(p + ".src_attn.linear_q.weight"),
                          st.g(p + ".src_attn.linear_q.bias"), dx=dx1.view(B * Lr, d), dx_residual=dt2.view(B * Lr, d))
            # x1 = LN1(t1), t1 = drop(ctx Wo) + xin
            gt, gd = gts[2], gds[2]
            dt1 = self._ln_bwd(dx1, t1, p + ".norm1", p + ".ln1", gt, dx_drop=gd.view(B, Lr, d), drop=sites[p + ".sa_out"])
            do = gd if sites[p + ".sa_out"].p > 0.0 else dt1.view(B * Lr, d)
            ctx = self.buf(p + ".sa.ctx", (B, Lr, d))
            self._lin_bwd(do, ctx.view(B * Lr, d), self.W(p + ".self_attn.linear_out.weight"), st.g(p + ".self_attn.linear_out.weight"),
                          st.g(p + ".self_attn.linear_out.bias"), dx=dctx.view(B * Lr, d))
            qkv = self.buf(p + ".qkv", (B, Lr, 3, H, dk))
            dqkv = self._scratch("g.qkv", (B, Lr, 3, H, dk))
            self._attn_core_bwd(dctx, qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], dqkv[:, :, 0], dqkv[:, :, 1], dqkv[:, :, 2], p + ".sa",
                                d_att.get(p + ".self_attn"), klens=self.olens_in, causal=True)
            self._lin_bwd(dqkv.view(B * Lr, 3 * d), xin.view(B * Lr, d), self.Wspan([p + ".self_attn.linear_q.weight"], (3 * d, d)),
                          st.span(st.G, [p + ".self_attn.linear_q.weight"], (3 * d, d)), st.span(st.G, [p + ".self_attn.linear_q.bias"], (3 * d,)),
                          dx=g.view(B * Lr, d), dx_residual=dt1.view(B * Lr, d))
            self._flush_defer()

        # ---- decoder input layer
        de = self._scratch("g.dec_b0", (B, Lr, d))
        ops.scaled_pe_bwd(g, self.pe(d, Lr), de, st.g("decoder.embed.1.alpha"), sites["dec.pe"])
        u = hp["dprenet_units"]
        npre = hp["dprenet_layers"]
        hin = self.buf(f"dec.prenet{npre - 1}", (B * Lr, u)) if npre > 0 else self.buf("dec.ys_in", (B, Lr, odim)).view(B * Lr, odim)
        pbufs = ("g.prenet_a", "g.prenet_b")
        dhp = self._scratch(pbufs[0], (B * Lr, u)) if npre > 0 else None
        # relu' (+ the always-on prenet dropout scale, pre_postnets.py:60-66) of prenet layer i rides in the dX GEMM of the layer above
        self._lin_bwd(de.view(B * Lr, d), hin, self.W("decoder.embed.0.1.weight"), st.g("decoder.embed.0.1.weight"),
                      st.g("decoder.embed.0.1.bias"), dx=dhp, dx_gate=hin if npre > 0 else None,
                      dx_gate_scale=sites[f"prenet{npre - 1}"].scale if npre > 0 else 1.0)
        for i in reversed(range(npre)):
            nm = f"decoder.embed.0.0.prenet.{i}.0"
            hin = self.buf(f"dec.prenet{i - 1}", (B * Lr, u)) if i > 0 else self.buf("dec.ys_in", (B, Lr, odim)).view(B * Lr, odim)
            dnext = self._scratch(pbufs[(npre - i) % 2], (B * Lr, u)) if i > 0 else None
            self._lin_bwd(dhp, hin, self.W(nm + ".weight"), st.g(nm + ".weight"), st.g(nm + ".bias"), dx=dnext,
                          dx_gate=hin if i > 0 else None, dx_gate_scale=sites[f"prenet{i - 1}"].scale if i > 0 else 1.0)
            dhp = dnext

        if on_decoder_done is not None:
            on_decoder_done()

        # ---- encoder (reverse)
        ge = self._scratch("g.enc_a", (B, T2, d))
        gte = self._scratch("g.enc_b", (B, T2, d))
        gde = self._scratch("g.enc_c", (B * T2, d))
        gdf = self._scratch("g.enc_e", (B, T2, d))        # dropout'(g) for the next (lower) layer's ff2 branch
        nle = hp["elayers"]
        conformer = hp["encoder_type"] == "conformer"
        if conformer:
            gx0 = self._conformer_bwd(dmem, self.buf("enc.x0", (B, T2, d)), "encoder", nle, H, hp["eunits"], hp["conformer_enc_kernel_size"],
                                      hp["transformer_enc_dropout_rate"], hp["enc_positional_dropout_rate"], hp["enc_attn_dropout_rate"])
            ops.scale_dropout(gx0, gte, math.sqrt(d), NO_DROP, self.named_drop("enc.posx", hp["enc_positional_dropout_rate"]))
            nle = 0
        else:
            self._ln_bwd(dmem, self.enc_last, "encoder.after_norm", "enc.after", ge, dx_drop=gdf if nle > 0 else None,
                         drop=sites[f"encoder.encoders.{nle - 1}.ff2"] if nle > 0 else NO_DROP)
        g = ge
        for l in reversed(range(nle)):
            p = f"encoder.encoders.{l}"
            xin = self.buf(f"encoder.encoders.{l - 1}.xout", (B, T2, d)) if l > 0 else self.buf("enc.x0", (B, T2, d))
            xm = self.buf(p + ".xmid", (B, T2, d))
            n1 = self.buf(p + ".ln1.y", (B, T2, d))
            n2 = self.buf(p + ".ln2.y", (B, T2, d))
            self._begin_defer()
            dn2 = gte
            if hp["positionwise_layer_type"] != "linear":
                self._ffn_conv_bwd_core(g, n2, p, "feed_forward", p + ".ffc", hp["eunits"], hp["transformer_enc_dropout_rate"], scale=1.0, dn=dn2)
            else:
                h = self.buf(p + ".ffh", (B * T2, hp["eunits"]))
                df = gdf.view(B * T2, d) if sites[p + ".ff2"].p > 0.0 else g.view(B * T2, d)
                dh = self._scratch("g.effh", (B * T2, hp["eunits"]))
                self._lin_bwd(df, h, self.W(p + ".feed_forward.w_2.weight"), st.g(p + ".feed_forward.w_2.weight"),
                              st.g(p + ".feed_forward.w_2.bias"), dx=dh, dx_gate=h, dx_gate_scale=sites[p + ".ff1"].scale)
                self._lin_bwd(dh, n2.view(B * T2, d), self.W(p + ".feed_forward.w_1.weight"), st.g(p + ".feed_forward.w_1.weight"),
                              st.g(p + ".feed_forward.w_1.bias"), dx=dn2.view(B * T2, d))
            gm = self._scratch("g.enc_d", (B, T2, d))
            self._ln_bwd(dn2, xm, p + ".norm2", p + ".ln2", gm, dres=g, dx_drop=gde.view(B, T2, d), drop=sites[p + ".sa_out"])   # g_mid = g + LN2'(dn2)
            do = gde if sites[p + ".sa_out"].p > 0.0 else gm.view(B * T2, d)
            ctx = self.buf(p + ".sa.ctx", (B, T2, d))
            dctx = self._scratch("g.ectx", (B, T2, d))
            self._lin_bwd(do, ctx.view(B * T2, d), self.W(p + ".self_attn.linear_out.weight"), st.g(p + ".self_attn.linear_out.weight"),
                          st.g(p + ".self_attn.linear_out.bias"), dx=dctx.view(B * T2, d))
            qkv = self.buf(p + ".qkv", (B, T2, 3, H, dk))
            dqkv = self._scratch("g.eqkv", (B, T2, 3, H, dk))
            self._attn_core_bwd(dctx, qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], dqkv[:, :, 0], dqkv[:, :, 1], dqkv[:, :, 2], p + ".sa",
                                d_att.get(p + ".self_attn"), klens=self.klens_enc, causal=False)
            dn1 = gte
            self._lin_bwd(dqkv.view(B * T2, 3 * d), n1.view(B * T2, d), self.Wspan([p + ".self_attn.linear_q.weight"], (3 * d, d)),
                          st.span(st.G, [p + ".self_attn.linear_q.weight"], (3 * d, d)), st.span(st.G, [p + ".self_attn.linear_q.bias"], (3 * d,)),
                          dx=dn1.view(B * T2, d))
            self._flush_defer()       # before g / gdf (operands of this layer's w_2 gradient) are overwritten
            self._ln_bwd(dn1, xin, p + ".norm1", p + ".ln1", g, dres=gm, dx_drop=gdf if l > 0 else None,
                         drop=sites[f"encoder.encoders.{l - 1}.ff2"] if l > 0 else NO_DROP)          # g_prev = g_mid + LN1'(dn1)

        # ---- encoder front end
        if hp["encoder_input"] == "embed":
            ops.embed_pe_bwd(g, self.xs, self.ilens_dev, self.pe(d, T2), st.g("encoder.embed.0.weight"), st.g("encoder.embed.1.alpha"),
                             hp["idim"] - 1, 0, sites["enc.pe"])
            return
        delin = gte
        if not conformer:
            ops.scaled_pe_bwd(g, self.pe(d, T2), delin, st.g("encoder.embed.out.1.alpha"), sites["enc.pe"])
        y2 = self.buf("enc.y2", (B * T2 * F2, d))
        woutp = self.buf("w.outp", (d, F2, d))
        gwoutp = self._scratch("g.woutp", (d, F2 * d), _f32)
        dy2 = self._scratch("g.y2", (B * T2 * F2, d))
        mode = self.mode
        ops.gemm(delin.view(B * T2, d).t(), y2.view(B * T2, F2 * d).t(), gwoutp, mode=mode)
        ops.transpose_last2(gwoutp, st.g("encoder.embed.out.0.weight"), d, F2, d, accumulate=True)
        ops.colsum(delin.view(B * T2, d), st.g("encoder.embed.out.0.bias"))
        ops.gemm(delin.view(B * T2, d), woutp.view(d, F2 * d).t(), dy2.view(B * T2, F2 * d), mode=mode, gate=y2.view(B * T2, F2 * d))
        w2p = self.buf("w.conv2p", (d, 9, d))
        col = self._scratch("col", (B * T2 * F2, 9 * d))
        y1 = self.buf("enc.y1", (B, T1, F1, d))
        if getattr(self, "_col_of", None) != self._sig:
            ops.im2col_s2(y1, col)                  # normally still there from this step's forward (0.5 GB at C2 is cheap on 180 GB)
        self._col_of = None
        gw2p = self._scratch("g.w2p", (d, 9 * d), _f32)
        ops.gemm(dy2.t(), col.t(), gw2p, mode=mode)
        ops.transpose_last2(gw2p, st.g("encoder.embed.conv.2.weight"), d, 9, d, accumulate=True)
        ops.colsum(dy2, st.g("encoder.embed.conv.2.bias"))
        dcol = col
        ops.gemm(dy2, w2p.view(d, 9 * d).t(), dcol, mode=mode)
        dy1 = self._scratch("g.y1", (B, T1, F1, d))
        ops.col2im_s2_relu(dcol, y1, dy1)          # scatter-add + conv.0's ReLU' in one pass
        self._conv1_bwd(self.xs, dy1, "encoder.embed", "enc")


    def encoder_span(self) -> int:
        """Number of leading elements of the flat parameter / gradient buffers that belong to the encoder (its parameters are
        registered first): [0, n) is final only at the very end of backward(), [n, numel) already at on_decoder_done."""
        names = self.store.names()
        first_dec = next(i for i, n in enumerate(names) if not n.startswith("encoder."))
        return self.store.offsets[names[first_dec]][0]

    def _site_table(self) -> Dict[str, Drop]:
        """Re-enumerate the forward's dropout sites (same order as forward())."""
        hp = self.hp
        saved = self._site
        self._site = 0
        t: Dict[str, Drop] = {}
        if hp["encoder_type"] != "conformer":           # the conformer blocks use name-keyed dropout sites (named_drop)
            t["enc.pe"] = self.drop(hp["enc_positional_dropout_rate"])
            for l in range(hp["elayers"]):
                p = f"encoder.encoders.{l}"
                t[p + ".sa_out"] = self.drop(hp["transformer_enc_dropout_rate"])
                conv_ffn = hp["positionwise_layer_type"] != "linear"          # conv position-wise layers use name-keyed sites
                t[p + ".ff1"] = NO_DROP if conv_ffn else self.drop(hp["transformer_enc_dropout_rate"])
                t[p + ".ff2"] = NO_DROP if conv_ffn else self.drop(hp["transformer_enc_dropout_rate"])
        for i in range(hp["dprenet_layers"]):
            self._site += 1
            pd = hp["dprenet_dropout_rate"]
            t[f"prenet{i}"] = Drop(pd, self.base_seed, self._site, self.seed_dev) if pd > 0 else NO_DROP
        t["dec.pe"] = self.drop(hp["dec_positional_dropout_rate"])
        for l in range(hp["dlayers"]):
            p = f"decoder.decoders.{l}"
            t[p + ".sa_out"] = self.drop(hp["dec_dropout_rate"])
            t[p + ".ca_out"] = self.drop(hp["dec_dropout_rate"])
            t[p + ".ff1"] = self.drop(hp["dec_dropout_rate"])
            t[p + ".ff2"] = self.drop(hp["dec_dropout_rate"])
        for i in range(hp["postnet_layers"]):
            t[f"post{i}"] = self.drop(hp["postnet_dropout_rate"])
        assert saved == 0 or saved == self._site, (saved, self._site)
        return t

    # ------------------------------------------------------------------ optimizer tail
