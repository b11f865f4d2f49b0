"""FastSpeechVC (conformer, non-autoregressive, teacher durations) training step on B200: explicit forward / backward over the C-ABI
kernels.

Hot path behind ``seq2seq_vc_b200.FastSpeechVC`` (reference: seq2seq_vc/models/fastspeech_vc.py:200-425 teacher-forced branch in the
configuration of egs/arctic/vc2/conf/fs2_vc.melmelmel.v1.yaml, the loss assembly of trainers/nar_vc.py:53-99).  The model shares every
block with AAS-VC -- conformer encoder / decoder (conformer_blocks.py), the Conv2dSubsampling projection + DurationPredictor on the
side input, feat_out, Postnet -- and differs in three places: the encoder's input layer is Conv2dSubsampling + RelPositionalEncoding
(conformer/encoder.py:124-131: the encoder runs at T' = ((T - 1) // 2 - 1) // 2 frames), there is no alignment module / MAS, and the
encoder output is expanded to frame level by the LengthRegulator with the TEACHER's integer durations (length_regulator.py:69-97:
`s2s_lr_cumsum` + `s2s_lr_fwd`, the adjoint `s2s_lr_bwd` in backward) instead of Gaussian upsampling.  Losses: L1 (after + before) and
the duration predictor's log-domain MSE against the teacher durations.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import ops
from ._lib import NO_DROP
from .aasvc_engine import AASVCEngine
from .conformer_blocks import conformer_buffer_specs, conformer_param_groups

_f32 = torch.float32
_i32 = torch.int32
NO_CLAMP = 3.0e38


def default_hparams(**over) -> dict:
    """model_params of egs/arctic/vc2/conf/fs2_vc.melmelmel.v1.yaml (+ the constructor defaults of FastSpeechVC)."""
    hp = dict(idim=80, odim=80, adim=384, aheads=2, elayers=4, eunits=1536, dlayers=4, dunits=1536, duration_predictor_input_dim=80,
              duration_predictor_layers=2, duration_predictor_chans=256, duration_predictor_kernel_size=3, postnet_layers=5, postnet_filts=5,
              postnet_chans=256, conformer_enc_kernel_size=15, conformer_dec_kernel_size=15, transformer_enc_dropout_rate=0.2,
              transformer_enc_positional_dropout_rate=0.2, transformer_enc_attn_dropout_rate=0.2, transformer_dec_dropout_rate=0.2,
              transformer_dec_positional_dropout_rate=0.2, transformer_dec_attn_dropout_rate=0.2, duration_predictor_dropout_rate=0.1,
              postnet_dropout_rate=0.5, positionwise_layer_type="linear", positionwise_conv_kernel_size=1,
              teacher_model_decoder_reduction_factor=1,
              # fixed for this model: no post-encoder reduction, deterministic predictor, no alignment loss
              post_encoder_reduction_factor=1, duration_predictor_type="deterministic", lambda_align=0.0)
    hp.update(over)
    if hp["positionwise_layer_type"] not in ("linear", "conv1d", "conv1d-linear"):
        raise NotImplementedError("Support only linear or conv1d.")
    return hp


def param_groups(hp: dict) -> List[List[Tuple[str, Tuple[int, ...]]]]:
    """Reference state-dict names / shapes in the reference's registration order (encoder, duration predictor, its projection,
    decoder, feat_out, postnet); Q / K / V projections adjacent."""
    d, H, idim, odim = hp["adim"], hp["aheads"], hp["idim"], hp["odim"]
    g: List[List[Tuple[str, Tuple[int, ...]]]] = []

    def lin(name, o, i):
        g.append([(name + ".weight", (o, i))])
        g.append([(name + ".bias", (o,))])

    def ln(name, n):
        g.append([(name + ".weight", (n,))])
        g.append([(name + ".bias", (n,))])

    def conv2d_sub(prefix, out_name, in_dim):
        f2 = ((in_dim - 1) // 2 - 1) // 2
        g.append([(prefix + ".conv.0.weight", (d, 1, 3, 3))])
        g.append([(prefix + ".conv.0.bias", (d,))])
        g.append([(prefix + ".conv.2.weight", (d, d, 3, 3))])
        g.append([(prefix + ".conv.2.bias", (d,))])
        lin(out_name, d, d * f2)

    conv2d_sub("encoder.embed", "encoder.embed.out.0", idim)
    conformer_param_groups(g, hp, "encoder", hp["elayers"], d, hp["eunits"], hp["conformer_enc_kernel_size"], H)
    ch, k = hp["duration_predictor_chans"], hp["duration_predictor_kernel_size"]
    for i in range(hp["duration_predictor_layers"]):
        g.append([(f"duration_predictor.conv.{i}.0.weight", (ch, d if i == 0 else ch, k))])
        g.append([(f"duration_predictor.conv.{i}.0.bias", (ch,))])
        ln(f"duration_predictor.conv.{i}.2", ch)
    lin("duration_predictor.linear", 1, ch)
    conv2d_sub("duration_predictor_projection", "duration_predictor_projection.out", hp["duration_predictor_input_dim"])
    conformer_param_groups(g, hp, "decoder", hp["dlayers"], d, hp["dunits"], hp["conformer_dec_kernel_size"], H)
    lin("feat_out", odim, d)
    pc, pk = hp["postnet_chans"], hp["postnet_filts"]
    for i in range(hp["postnet_layers"]):
        ic = odim if i == 0 else pc
        oc = odim if i == hp["postnet_layers"] - 1 else pc
        g.append([(f"postnet.postnet.{i}.0.weight", (oc, ic, pk))])
        ln(f"postnet.postnet.{i}.1", oc)
    return g


def buffer_specs(hp: dict):
    out = conformer_buffer_specs("encoder", hp["elayers"], hp["adim"]) + conformer_buffer_specs("decoder", hp["dlayers"], hp["adim"])
    for i in range(hp["postnet_layers"]):
        c = hp["odim"] if i == hp["postnet_layers"] - 1 else hp["postnet_chans"]
        p = f"postnet.postnet.{i}.1"
        out += [(p + ".running_mean", (c,), _f32), (p + ".running_var", (c,), _f32), (p + ".num_batches_tracked", (), torch.int64)]
    return out


class FastSpeechVCEngine(AASVCEngine):
    """Owns parameters, activation buffers and the explicit forward / loss / backward of one FastSpeechVC step."""

    LOSS_NAMES = ("l1_loss", "duration_loss")

    def __init__(self, hp: dict, device="cuda:0", bf16: bool = False, seed: int = 0, fp32_gemm: str = "tc"):
        self.fp32_gemm = fp32_gemm
        self.hp = default_hparams(**hp)
        hp = self.hp
        assert hp["adim"] % hp["aheads"] == 0
        self.stochastic = False
        self._setup(param_groups(hp), buffer_specs(hp), device, bf16, seed)
        self.losses = torch.zeros(2, dtype=_f32, device=self.device)
        self._l1_pair = torch.zeros(2, dtype=_f32, device=self.device)
        self._loss_ws = torch.zeros(4, dtype=_f32, device=self.device)
        self._one = torch.ones(1, dtype=_f32, device=self.device)
        self._relpe: Dict = {}
        self._interp: Dict = {}
        self._prior_key: Dict = {}
        self._last: Dict[str, torch.Tensor] = {}
        self.init_parameters(seed)

    # ------------------------------------------------------------------ lengths
    def prepare(self, B: int, T: int, L: int, ilens: Sequence[int], olens: Sequence[int]) -> None:
        """Host-side length vectors in ONE small H2D copy: encoder key lengths after the conv2d input layer (the subsampled mask,
        subsampling.py:92-94), the lengths the model reports / the duration loss masks with (fastspeech_vc.py:236-238), target lengths."""
        T2 = (((T - 1) // 2) - 1) // 2
        self._use_sig((B, T, L, self.training))
        ilens = [int(v) for v in ilens]
        olens = [int(v) for v in olens]
        assert len(ilens) == B and len(olens) == B
        klens = [min(T2, (i + 3) // 4) for i in ilens]
        tlens = [((i - 2 + 1) // 2 - 2 + 1) // 2 for i in ilens]
        self._ship_lens([klens, tlens, olens])
        self.tlens_host, self.olens_host = tlens, olens
        self._prepared = (B, T, L)

    # ------------------------------------------------------------------ forward
    def forward(self, xs: torch.Tensor, ys: torch.Tensor, ds: torch.Tensor, dp_inputs: torch.Tensor, ilens: Optional[Sequence[int]] = None,
                olens: Optional[Sequence[int]] = None):
        """xs (B,T,idim), ys (B,L,odim), dp_inputs (B,T_dp,dp_idim) float32; ds (B,T') int64 teacher durations (device), T' the
        encoder length; L = max_b sum(ds[b]) * teacher factor (the caller -- collater or drop-in module -- knows it on the host).
        Returns (after (B,L,odim), before) in activation dtype; sets self.d_outs after loss()."""
        hp, st = self.hp, self.store
        B, T, idim = xs.shape
        L, odim = ys.shape[1], ys.shape[2]
        d, H = hp["adim"], hp["aheads"]
        T2 = (((T - 1) // 2) - 1) // 2
        assert xs.dtype == _f32 and ys.dtype == _f32 and dp_inputs.dtype == _f32 and ds.dtype == torch.int64
        assert xs.is_contiguous() and ys.is_contiguous() and dp_inputs.is_contiguous() and ds.shape == (B, T2)
        self._sig = (B, T, L, self.training)
        self.attn = {}
        self._last = {}
        self.sync_shadow()
        if ilens is not None:
            self.prepare(B, T, L, ilens, olens)
        assert self._prepared == (B, T, L), "prepare(B, T, L, ilens, olens) must precede forward() for this batch shape"
        lens = self.buf("lens", (3, B), _i32)
        self.klens_dev, self.tlens_dev, self.olens_dev = lens[0], lens[1], lens[2]
        self.shapes = dict(B=B, T=T, L=L, Tt=T2, Tdp=dp_inputs.shape[1])
        self.xs, self.dp_inputs = xs, dp_inputs
        (er, epr, ear), (dr_, dpr, dar) = self._rates()
        # ---- encoder: Conv2dSubsampling + RelPositionalEncoding (x * sqrt(d), positional dropout) + conformer blocks
        elin = self._conv2d_sub_fwd(xs, "encoder.embed", "encoder.embed.out.0", "enc")
        x0 = self.buf("enc.x0", (B, T2, d))
        ops.scale_dropout(elin.view(B, T2, d), x0, math.sqrt(d), NO_DROP, self.named_drop("enc.pos", epr))
        hs = self._conformer_fwd(x0, "encoder", hp["elayers"], H, hp["eunits"], hp["conformer_enc_kernel_size"], self.klens_dev, er, epr, ear)
        self.hs = hs
        # ---- duration predictor on the projected side input (fastspeech_vc.py:244-275)
        self._dp_input_fwd(dp_inputs, T2)
        self._dp_forward(T2)
        # ---- length regulator with the teacher's durations (fastspeech_vc.py:276-279)
        self.ds = ds.contiguous()
        cum = self.buf("lr.cum", (B, T2 + 1), _i32)
        ops.lr_cumsum(self.ds, cum, float(hp["teacher_model_decoder_reduction_factor"]))
        up = self.buf("lr.out", (B, L, d))
        ops.lr_fwd(hs, cum, up, 0.0)
        # ---- decoder: RelPositionalEncoding + conformer blocks, feat_out, postnet (fastspeech_vc.py:281-305)
        xd0 = self.buf("dec.x0", (B, L, d))
        ops.scale_dropout(up, xd0, math.sqrt(d), self.named_drop("dec.pos", dpr))
        zs = self._conformer_fwd(xd0, "decoder", hp["dlayers"], H, hp["dunits"], hp["conformer_dec_kernel_size"], self.olens_dev, dr_, dpr, dar)
        self.zs = zs
        before = self.buf("out.before", (B, L, odim))
        self._lin_fwd(zs.view(B * L, d), self.W("feat_out.weight"), st.p("feat_out.bias"), before.view(B * L, odim))
        after = self._postnet_fwd(before, lambda i: self.named_drop(f"post{i}", hp["postnet_dropout_rate"]))
        self.before, self.after = before, after
        return after, before

    def forward_d_outs(self) -> torch.Tensor:
        """d_outs (B, T') = masked log-domain predictor output, as the reference returns it (duration_predictor.py:98-101)."""
        B, Tt = self.shapes["B"], self.shapes["Tt"]
        m = (torch.arange(Tt, device=self.device)[None, :] < self.tlens_dev[:, None])
        return self.dp_pre.view(B, Tt).float() * m

    # ------------------------------------------------------------------ losses (trainers/nar_vc.py:73-82)
    def loss(self, ys: torch.Tensor):
        B, L, odim = self.after.shape
        Tt = self.shapes["Tt"]
        self.d_after = self.buf("loss.d_after", self.after.shape)
        self.d_before = self.buf("loss.d_before", self.after.shape)
        zl = self.buf("loss.zero_logits", (B, L), zero=True)
        zlab = self.buf("loss.zero_labels", (B, L), _f32, zero=True)
        dzl = self.buf("loss.d_logits", (B, L))
        ops.seq2seq_loss(self.after, self.before, zl, ys, zlab, self.olens_dev, 1.0, self._l1_pair, self.d_after, self.d_before, dzl, self._loss_ws)
        self.losses[0:1].copy_(self._l1_pair[0:1])
        self.d_outs = self.buf("dp.d_outs", (B, Tt), _f32)
        self.d_dp_pre = self.buf("dp.d_pre", (B * Tt, 1))
        self.ds_f = self.buf("lr.ds_f", (B, Tt), _f32)
        self.ds_f.copy_(self.ds)
        # (no clamp of the predictor's output in this model: aas_vc.py:408-410 has one, fastspeech_vc.py:270-275 does not)
        ops.duration_loss(self.dp_pre, self.ds_f, self.tlens_dev, self.d_outs, self.losses[1:2], self.d_dp_pre, 1.0, clamp_max=NO_CLAMP)
        return self.losses

    def total_loss(self) -> torch.Tensor:
        return self.losses[0] + self.losses[1]

    # ------------------------------------------------------------------ backward
    def backward(self, d_after=None, d_before=None, d_dp_pre=None, zero_grad: bool = True) -> None:
        hp, st = self.hp, self.store
        d_after = self.d_after if d_after is None else d_after
        d_before = self.d_before if d_before is None else d_before
        d_dp_pre = self.d_dp_pre if d_dp_pre is None else d_dp_pre
        s = self.shapes
        B, T, L, Tt = s["B"], s["T"], s["L"], s["Tt"]
        d, H, odim = hp["adim"], hp["aheads"], hp["odim"]
        (er, epr, ear), (dr_, dpr, dar) = self._rates()
        if zero_grad:
            st.G.zero_()
        dbefore = self._postnet_bwd(d_after, d_before, lambda i: self.named_drop(f"post{i}", hp["postnet_dropout_rate"]))
        gz = self._scratch("g.zs", (B, L, d))
        self._lin_bwd(dbefore.view(B * L, odim), self.zs.view(B * L, d), self.W("feat_out.weight"), st.g("feat_out.weight"),
                      st.g("feat_out.bias"), dx=gz.view(B * L, d))
        xd0 = self.buf("dec.x0", (B, L, d))
        gx = self._conformer_bwd(gz, xd0, "decoder", hp["dlayers"], H, hp["dunits"], hp["conformer_dec_kernel_size"], dr_, dpr, dar)
        gup = self._scratch("g.up", (B, L, d))
        ops.scale_dropout(gx, gup, math.sqrt(d), self.named_drop("dec.pos", dpr))
        # ---- length regulator adjoint: d_hs[b, i] = sum of d_up over row i's run of frames
        dhs = self._scratch("g.hs", (B, Tt, d))
        ops.lr_bwd(gup, self.buf("lr.cum", (B, Tt + 1), _i32), dhs)
        # ---- duration predictor + its projection
        self._dp_backward(d_dp_pre, Tt)
        # ---- encoder
        x0 = self.buf("enc.x0", (B, Tt, d))
        gx0 = self._conformer_bwd(dhs, x0, "encoder", hp["elayers"], H, hp["eunits"], hp["conformer_enc_kernel_size"], er, epr, ear)
        gelin = self._scratch("g.elin", (B, Tt, d))
        ops.scale_dropout(gx0, gelin, math.sqrt(d), NO_DROP, self.named_drop("enc.pos", epr))
        self._conv2d_sub_bwd(gelin.view(B * Tt, d), self.xs, "encoder.embed", "encoder.embed.out.0", "enc")

    def optimizer_step(self, max_norm: float = 1.0, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                       grad_scale: float = 1.0) -> None:
        """clip_grad_norm_ + Adam over the flat buffers (trainers/nar_vc.py:88-96)."""
        st = self.store
        ops.step_advance(self.step_dev, self.seed_dev)
        self._sqn.zero_()
        ops.sqnorm(st.G, self._sqn)
        ops.adam_step(st.P, st.G, st.M, st.V, st.P16, self.lr_dev, betas[0], betas[1], eps, weight_decay, self.step_dev, self._sqn, max_norm,
                      grad_scale)
        self.p16_dirty = False

    # ------------------------------------------------------------------ inference (fastspeech_vc.py:427-470)
    def inference(self, x: torch.Tensor, dp_input: torch.Tensor, alpha: float = 1.0):
        """x (T, idim), dp_input (T_dp, dp_idim): predicted durations -> LengthRegulator -> decoder.  Returns (outs (L, odim), d_outs (T',))."""
        hp, st = self.hp, self.store
        was = self.training
        self.training = False
        try:
            T = x.shape[0]
            xs = x.to(_f32).contiguous().unsqueeze(0)
            dpi = dp_input.to(_f32).contiguous().unsqueeze(0)
            d, H, odim = hp["adim"], hp["aheads"], hp["odim"]
            T2 = (((T - 1) // 2) - 1) // 2
            self._use_sig((1, T, -1, False))
            self._sig = (1, T, -1, False)
            self.attn, self._last = {}, {}
            self.sync_shadow()
            klens = torch.full((1,), T2, dtype=_i32, device=self.device)
            self.shapes = dict(B=1, T=T, L=0, Tt=T2, Tdp=dpi.shape[1])
            elin = self._conv2d_sub_fwd(xs, "encoder.embed", "encoder.embed.out.0", "enc")
            x0 = self.buf("enc.x0", (1, T2, d))
            ops.scale_dropout(elin.view(1, T2, d), x0, math.sqrt(d))
            hs = self._conformer_fwd(x0, "encoder", hp["elayers"], H, hp["eunits"], hp["conformer_enc_kernel_size"], klens, 0.0, 0.0, 0.0)
            self._dp_input_fwd(dpi, T2)
            self._dp_forward(T2)
            d_outs = torch.empty(T2, dtype=_f32, device=self.device)
            ops.duration_infer(self.dp_pre, d_outs, clamp_max=NO_CLAMP)          # clamp(round(exp(.) - 1), min 0) (duration_predictor.py:116-128)
            ds = (d_outs * hp["teacher_model_decoder_reduction_factor"]).to(torch.int64).view(1, T2).contiguous()
            cum = torch.empty(1, T2 + 1, dtype=_i32, device=self.device)
            ops.lr_cumsum(ds, cum, alpha)
            L = int(cum[0, T2].item())
            if L == 0:                                                           # all predicted durations 0: every duration becomes 1
                ops.lr_cumsum(ds, cum, alpha, all_ones=True)
                L = T2
            up = torch.empty(1, L, d, dtype=self.adt, device=self.device)
            ops.lr_fwd(hs, cum, up, 0.0)
            self._sig = (1, T, L, False)
            xd0 = self.buf("dec.x0", (1, L, d))
            ops.scale_dropout(up, xd0, math.sqrt(d))
            olens = torch.full((1,), L, dtype=_i32, device=self.device)
            zs = self._conformer_fwd(xd0, "decoder", hp["dlayers"], H, hp["dunits"], hp["conformer_dec_kernel_size"], olens, 0.0, 0.0, 0.0)
            before = self.buf("out.before", (1, L, odim))
            self._lin_fwd(zs.view(L, d), self.W("feat_out.weight"), st.p("feat_out.bias"), before.view(L, odim))
            after = self._postnet_fwd(before, lambda i: NO_DROP)
            return after[0].float().clone(), d_outs
        finally:
            self.training = was
