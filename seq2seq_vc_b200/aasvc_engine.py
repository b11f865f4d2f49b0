"""AAS-VC (Conformer, non-autoregressive) training step on B200: explicit forward / backward over the C-ABI kernels.

Hot path behind ``seq2seq_vc_b200.AASVC`` (reference: seq2seq_vc/models/aas_vc.py:279-471 teacher-forced branch,
modules/conformer/{encoder,encoder_layer,convolution}.py, modules/transformer/attention.py:209-305,
layers/positional_encoding.py:238-309, modules/alignments.py:12-60,281-310, modules/length_regulator.py:100-154,
modules/duration_predictor.py:27-128, losses/{l1_loss,forward_sum_loss,duration_predictor_loss}.py and the loss
assembly of trainers/aas_vc.py:56-134).  Configuration family: egs/arctic/vc2/conf/aas_vc.melmelmel.v1.yaml with the
deterministic duration predictor (encoder / decoder reduction factor 1, `linear` input layer, macaron + CNN conformer
blocks, rel_pos / rel_selfattn, pre-LN, Conv2dSubsampling projection of the duration-predictor input).

As in VTNEngine there is no autograd tape: activations live in named HBM buffers, the backward is written out op by
op, dropout masks are regenerated from a counter-based RNG, and nothing synchronises with the host -- the monotonic
alignment search, the forward-sum recursion and the beta-binomial prior lookup all stay on the device / in a cached
device tensor, so one step is a fixed launch sequence that can be captured in a CUDA graph.

Data layout in HBM
  activations    (B, T, d) row-major, f32 (parity mode) or bf16; conv-module tensors are channels-last too
  attention      ac / probabilities (B, H, T, ld) with ld = T rounded up to 8; the un-shifted rel-pos term bd is
                 (H, B, T, ldb >= 2T-1) so that both of its GEMMs see one (B*T)-row operand per head
  alignment      log_p_attn (B, T_feats, T_text) float32 (the reference's 4-D (B,T_feats,T_text,C) difference tensor is
                 never formed); prior / alpha workspace share that shape
"""
from __future__ import annotations

import math
import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import ops
from ._lib import NO_DROP, Drop
from .conformer_blocks import ConformerBlocks, conformer_param_groups, rel_pos_table  # noqa: F401
from .engine_base import EngineBase, _r8

_f32 = torch.float32
_i32 = torch.int32

EMBED_LN_EPS = 1e-5     # torch.nn.LayerNorm default in the `linear` input layer (conformer/encoder.py:119)


def default_hparams(**over) -> dict:
    """model_params of egs/arctic/vc2/conf/aas_vc.melmelmel.v1.yaml (+ the constructor defaults of AASVC)."""
    hp = dict(idim=80, odim=80, adim=384, aheads=2, elayers=4, eunits=1536, dlayers=4, dunits=1536,
              duration_predictor_input_dim=80, duration_predictor_layers=2, duration_predictor_chans=256,
              duration_predictor_kernel_size=3, postnet_layers=5, postnet_filts=5, postnet_chans=256,
              post_encoder_reduction_factor=4, conformer_enc_kernel_size=15, conformer_dec_kernel_size=15,
              transformer_enc_dropout_rate=0.2, transformer_enc_positional_dropout_rate=0.2,
              transformer_enc_attn_dropout_rate=0.2, transformer_dec_dropout_rate=0.2,
              transformer_dec_positional_dropout_rate=0.2, transformer_dec_attn_dropout_rate=0.2,
              duration_predictor_dropout_rate=0.1, postnet_dropout_rate=0.5, lambda_align=2.0,
              # "linear": PositionwiseFeedForward + Swish (the shipped yaml); "conv1d": MultiLayeredConv1d + ReLU (kernel size 1 is
              # the AASVC class default, models/aas_vc.py:52-53; odd kernel sizes > 1 run as taps-GEMMs over haloed rows);
              # "conv1d-linear": Conv1dLinear (multi_layer_conv.py:66-108)
              positionwise_layer_type="linear", positionwise_conv_kernel_size=1,
              # "deterministic": DurationPredictor + DurationPredictorLoss; "stochastic": StochasticDurationPredictor (VITS flows),
              # the shipped yaml's default (aas_vc.melmelmel.v1.yaml:57; constructor defaults models/aas_vc.py:105-110)
              duration_predictor_type="deterministic", stochastic_duration_predictor_kernel_size=3,
              stochastic_duration_predictor_dropout_rate=0.5, stochastic_duration_predictor_flows=4,
              stochastic_duration_predictor_dds_conv_layers=3, stochastic_duration_predictor_noise_scale=0.8)
    hp.update(over)
    if hp["duration_predictor_type"] not in ("deterministic", "stochastic"):
        raise ValueError(f"Duration predictor type: {hp['duration_predictor_type']} is not supported.")
    if hp["positionwise_layer_type"] not in ("linear", "conv1d", "conv1d-linear"):
        raise NotImplementedError("Support only linear or conv1d.")          # conformer/encoder.py:205
    if hp["positionwise_layer_type"] != "linear" and hp["positionwise_conv_kernel_size"] % 2 != 1:
        raise NotImplementedError("even positionwise_conv_kernel_size (the reference's padding (k - 1) // 2 then shortens the sequence)")
    return hp


def sdp_hparams(hp: dict) -> dict:
    """StochasticDurationPredictor(channels=adim, ...) as AASVC builds it (models/aas_vc.py:178-186)."""
    return dict(channels=hp["adim"], kernel_size=hp["stochastic_duration_predictor_kernel_size"],
                dds_conv_layers=hp["stochastic_duration_predictor_dds_conv_layers"], flows=hp["stochastic_duration_predictor_flows"])


def beta_binomial_log_prior(N: int, T: int) -> torch.Tensor:
    """(T, N) float32 log-pmf of BetaBinomial(k; n=N, a=t, b=T-t+1), t = 1..T, k = 0..N-1, exactly the table
    ForwardSumLoss._generate_prior builds with scipy on the host (losses/forward_sum_loss.py:100-114); float64 lgamma
    on the host, cast to float32 before the add as the reference does (:48-49)."""
    k = torch.arange(N, dtype=torch.float64)[None, :]
    a = torch.arange(1, T + 1, dtype=torch.float64)[:, None]
    b = T - a + 1
    n = float(N)
    lg = torch.lgamma
    log_binom = lg(torch.tensor(n + 1, dtype=torch.float64)) - lg(k + 1) - lg(n - k + 1)
    betaln = lambda x, y: lg(x) + lg(y) - lg(x + y)
    return (log_binom + betaln(k + a, n - k + b) - betaln(a, b)).to(_f32)


def nearest_index(T_in: int, T_out: int) -> List[int]:
    """Source row of every output row for F.interpolate(mode="nearest") (aas_vc.py:345-348): floor(dst * scale) with
    scale = T_in / T_out in float32, clamped to T_in - 1."""
    scale = torch.tensor(T_in, dtype=_f32) / torch.tensor(T_out, dtype=_f32)
    idx = torch.floor(torch.arange(T_out, dtype=_f32) * scale).to(torch.int64).clamp_(max=T_in - 1)
    return idx.tolist()


def param_groups(hp: dict) -> List[List[Tuple[str, Tuple[int, ...]]]]:
    """Reference state-dict names / shapes; Q/K/V projections of every attention are adjacent (one fused GEMM)."""
    d, H, pr = hp["adim"], hp["aheads"], hp["post_encoder_reduction_factor"]
    idim, odim = hp["idim"], hp["odim"]
    g: List[List[Tuple[str, Tuple[int, ...]]]] = []

    def lin(name, o, i, bias=True):
        g.append([(name + ".weight", (o, i))])
        if bias:
            g.append([(name + ".bias", (o,))])

    def ln(name, n):
        g.append([(name + ".weight", (n,))])
        g.append([(name + ".bias", (n,))])

    def conformer(prefix, n_layers, dm, units, k):
        conformer_param_groups(g, hp, prefix, n_layers, dm, units, k, H)

    lin("encoder.embed.0", d, idim)
    ln("encoder.embed.1", d)
    conformer("encoder", hp["elayers"], d, hp["eunits"], hp["conformer_enc_kernel_size"])
    if hp["duration_predictor_type"] == "stochastic":
        from . import sdp
        for name, shape in sdp.param_spec(sdp_hparams(hp), "duration_predictor"):
            g.append([(name, shape)])
    else:
        ch, k = hp["duration_predictor_chans"], hp["duration_predictor_kernel_size"]
        for i in range(hp["duration_predictor_layers"]):
            g.append([(f"duration_predictor.conv.{i}.0.weight", (ch, d if i == 0 else ch, k))])
            g.append([(f"duration_predictor.conv.{i}.0.bias", (ch,))])
            ln(f"duration_predictor.conv.{i}.2", ch)
        lin("duration_predictor.linear", 1, ch)
    f2 = ((hp["duration_predictor_input_dim"] - 1) // 2 - 1) // 2
    g.append([("duration_predictor_projection.conv.0.weight", (d, 1, 3, 3))])
    g.append([("duration_predictor_projection.conv.0.bias", (d,))])
    g.append([("duration_predictor_projection.conv.2.weight", (d, d, 3, 3))])
    g.append([("duration_predictor_projection.conv.2.bias", (d,))])
    lin("duration_predictor_projection.out", d, d * f2)
    C = d * pr
    for n, ic, kk in (("t_conv1", C, 3), ("t_conv2", C, 1), ("f_conv1", odim, 3), ("f_conv2", C, 3), ("f_conv3", C, 1)):
        g.append([(f"alignment_module.{n}.weight", (C, ic, kk))])
        g.append([(f"alignment_module.{n}.bias", (C,))])
    conformer("decoder", hp["dlayers"], C, hp["dunits"], hp["conformer_dec_kernel_size"])
    lin("feat_out", odim, C)
    pc, pk = hp["postnet_chans"], hp["postnet_filts"]
    for i in range(hp["postnet_layers"]):
        ic = odim if i == 0 else pc
        oc = odim if i == hp["postnet_layers"] - 1 else pc
        g.append([(f"postnet.postnet.{i}.0.weight", (oc, ic, pk))])
        ln(f"postnet.postnet.{i}.1", oc)
    return g


def buffer_specs(hp: dict) -> List[Tuple[str, Tuple[int, ...], torch.dtype]]:
    """BatchNorm running statistics (conformer conv modules and postnet)."""
    out = []

    def bn(p, c):
        out.extend([(p + ".running_mean", (c,), _f32), (p + ".running_var", (c,), _f32), (p + ".num_batches_tracked", (), torch.int64)])

    d, C = hp["adim"], hp["adim"] * hp["post_encoder_reduction_factor"]
    for l in range(hp["elayers"]):
        bn(f"encoder.encoders.{l}.conv_module.norm", d)
    for l in range(hp["dlayers"]):
        bn(f"decoder.encoders.{l}.conv_module.norm", C)
    for i in range(hp["postnet_layers"]):
        bn(f"postnet.postnet.{i}.1", hp["odim"] if i == hp["postnet_layers"] - 1 else hp["postnet_chans"])
    return out


class AASVCEngine(ConformerBlocks, EngineBase):
    """Owns parameters, activation buffers and the explicit forward / loss / backward of one AAS-VC step."""

    align_dot = os.environ.get("S2S_ALIGN_DOT", "1") != "0"      # A/B switch: tensor-core pairwise distances in the bf16 engine

    LOSS_NAMES = ("l1_loss", "forward_sum_loss", "bin_loss", "duration_loss")

    def __init__(self, hp: dict, device="cuda:0", bf16: bool = False, seed: int = 0, fp32_gemm: str = "tc"):
        """fp32_gemm (float32 engines only): "tc" = fp32-accurate tcgen05 GEMM (bf16-split operands), "simt" = CUDA-core GEMM."""
        self.fp32_gemm = fp32_gemm
        self.hp = default_hparams(**hp)
        hp = self.hp
        assert hp["adim"] % hp["aheads"] == 0
        self._setup(param_groups(hp), buffer_specs(hp), device, bf16, seed)
        self.losses = torch.zeros(4, dtype=_f32, device=self.device)       # l1, forward-sum, bin, duration
        self._l1_pair = torch.zeros(2, dtype=_f32, device=self.device)
        self._loss_ws = torch.zeros(4, dtype=_f32, device=self.device)
        self._one = torch.ones(1, dtype=_f32, device=self.device)
        # flat-buffer span [a, b) of the duration predictor (+ its input projection) and its own Adam clock: see optimizer_step
        names = self.store.names()
        dp = [i for i, n in enumerate(names) if n.startswith(("duration_predictor.", "duration_predictor_projection."))]
        assert dp and dp == list(range(dp[0], dp[-1] + 1)), "duration-predictor parameters must be contiguous in the flat buffer"
        self._dp_span = (self.store.offsets[names[dp[0]]][0],
                         self.store.offsets[names[dp[-1] + 1]][0] if dp[-1] + 1 < len(names) else self.store.numel)
        self.dp_step_dev = torch.zeros(1, dtype=_f32, device=self.device)
        self._prior_cache: Dict[Tuple, torch.Tensor] = {}
        self._prior_key: Dict[Tuple, Tuple] = {}
        self._prior_tables: Dict[Tuple[int, int], torch.Tensor] = {}
        self._relpe: Dict[Tuple[int, int], torch.Tensor] = {}
        self._interp: Dict[Tuple[int, int], Tuple[torch.Tensor, ...]] = {}
        self.stochastic = hp["duration_predictor_type"] == "stochastic"
        self.sdp = None
        if self.stochastic:
            from . import sdp

            # autograd leaves over the flat store: detached views whose .grad IS the matching slice of the flat gradient buffer
            self._sdp_leaves: Dict[str, torch.Tensor] = {}
            for name, _ in sdp.param_spec(sdp_hparams(hp), "duration_predictor"):
                leaf = self.store.p(name).detach().requires_grad_(True)
                leaf.grad = self.store.g(name)
                self._sdp_leaves[name] = leaf
            self.sdp = sdp.StochasticDurationPredictor(
                sdp_hparams(hp), "duration_predictor", self._sdp_leaves.__getitem__, gemm_mode=0 if fp32_gemm == "simt" else 2,
                dropout_rate=hp["stochastic_duration_predictor_dropout_rate"],
                drop_of=lambda name, p: self.named_drop("sdp." + name, p))
        self.init_parameters(seed)

    # ------------------------------------------------------------------ parameters
    def init_parameters(self, seed: int = 0) -> None:
        """torch-default Linear / Conv init distributions (uniform +-1/sqrt(fan_in)); LN/BN affine = 1/0; xavier-uniform
        pos_bias_u/v (attention.py:233-234)."""
        g = torch.Generator().manual_seed(seed)
        sdp_init = {}
        if self.hp["duration_predictor_type"] == "stochastic":
            from . import sdp
            sdp_init = sdp.init_params(sdp_hparams(self.hp), "duration_predictor", seed + 1)
        for name, (off, shape) in self.store.offsets.items():
            n = 1
            for s in shape:
                n *= s
            if name in sdp_init:
                self.store.P[off:off + n].copy_(sdp_init[name].reshape(-1).to(self.device))
                continue
            is_affine = (("norm" in name) or name.startswith("encoder.embed.1") or (name.startswith("postnet") and ".1." in name)
                         or (name.startswith("duration_predictor.conv") and name.split(".")[-2] == "2"))
            if is_affine:
                v = torch.ones(n) if name.endswith("weight") else torch.zeros(n)
            elif "pos_bias" in name:
                bound = math.sqrt(6.0 / (shape[0] + shape[1]))
                v = (torch.rand(n, generator=g) * 2 - 1) * bound
            else:
                wname = name[:-5] + ".weight" if name.endswith(".bias") else name
                wshape = self.store.offsets[wname][1]
                fan_in = 1
                for s in wshape[1:]:
                    fan_in *= s
                v = (torch.rand(n, generator=g) * 2 - 1) / math.sqrt(fan_in)
            self.store.P[off:off + n].copy_(v.to(self.device))
        self.p16_dirty = True

    # ------------------------------------------------------------------ host-side preparation (no device sync)
    def prepare(self, B: int, T: int, L: int, ilens: Sequence[int], olens: Sequence[int]) -> None:
        """Length vectors, the beta-binomial prior and the interpolation index for this batch: built on the host from
        the CPU ints the collater provides, cached by value, shipped with non-blocking copies."""
        hp = self.hp
        pr = hp["post_encoder_reduction_factor"]
        ilens = [int(v) for v in ilens]
        olens = [int(v) for v in olens]
        assert len(ilens) == B and len(olens) == B and max(ilens) <= T and max(olens) <= L
        Tt = T // pr
        assert Tt >= 1, "source too short for the post-encoder reduction factor"
        tlens = [i // pr for i in ilens]
        assert all(t >= 1 for t in tlens) and all(o >= 1 for o in olens)
        self._use_sig((B, T, L, self.training))
        self._ship_lens([ilens, tlens, olens])
        self.ilens_host, self.tlens_host, self.olens_host = ilens, tlens, olens
        # beta-binomial prior (B, L, Tt) in a shape-stable device buffer (a captured CUDA graph keeps reading it);
        # re-filled only when the length pattern changes.  -inf never enters: the kernel reads t < olen, k < tlen only
        key = (L, Tt, tuple(tlens), tuple(olens))
        pbuf = self.buf("prior", (B, L, Tt), _f32)
        if self._prior_key.get(self._sig) != key:
            prior = self._prior_cache.get(key)
            if prior is None:
                ph = torch.zeros(B, L, Tt, dtype=_f32)
                for b in range(B):
                    tab = self._prior_tables.get((tlens[b], olens[b]))
                    if tab is None:
                        tab = beta_binomial_log_prior(tlens[b], olens[b])
                        self._prior_tables[(tlens[b], olens[b])] = tab
                    ph[b, :olens[b], :tlens[b]] = tab
                prior = ph.pin_memory() if self.device.type == "cuda" else ph
                if len(self._prior_cache) > 8:
                    self._prior_cache.clear()
                self._prior_cache[key] = prior
            pbuf.copy_(prior, non_blocking=True)
            self._prior_key[self._sig] = key
        self.prior = pbuf
        self._prepared = (B, T, L)

    def _evict_sig(self, sig) -> None:
        super()._evict_sig(sig)
        self._prior_key.pop(sig, None)

    def _interp_tables(self, T_in: int, T_out: int):
        """(start, ones) for the forward gather and (run start, run length) per source row for its adjoint."""
        t = self._interp.get((T_in, T_out))
        if t is None:
            idx = nearest_index(T_in, T_out)
            first = [0] * T_in
            cnt = [0] * T_in
            for j, i in enumerate(idx):
                if cnt[i] == 0:
                    first[i] = j
                cnt[i] += 1
            mk = lambda v: torch.tensor(v, dtype=_i32).to(self.device)
            t = (mk(idx), mk([1] * T_out), mk(first), mk(cnt))
            self._interp[(T_in, T_out)] = t
        return t

    # ------------------------------------------------------------------ conformer block
    # ------------------------------------------------------------------ forward
    def _rates(self):
        hp = self.hp
        return ((hp["transformer_enc_dropout_rate"], hp["transformer_enc_positional_dropout_rate"], hp["transformer_enc_attn_dropout_rate"]),
                (hp["transformer_dec_dropout_rate"], hp["transformer_dec_positional_dropout_rate"], hp["transformer_dec_attn_dropout_rate"]))

    def _encoder_side(self, xs: torch.Tensor, dp_inputs: torch.Tensor) -> None:
        """Encoder input layer + conformer encoder + post-encoder reduction -> self.hs (B,Tt,C); duration-predictor input
        projection + duration predictor -> self.dp_pre (B*Tt, 1) (log-domain pre-activation) and self.dp_last."""
        hp, st = self.hp, self.store
        B, T, idim = xs.shape
        d, H, pr = hp["adim"], hp["aheads"], hp["post_encoder_reduction_factor"]
        C, Tt = d * pr, T // pr
        (er, epr, ear), _ = self._rates()
        # ---- encoder input layer: Linear -> LayerNorm(1e-5) -> dropout -> x * sqrt(d) -> dropout (conformer/encoder.py:117-123)
        xa = xs
        if self.bf16:
            xa = ops.cast(xs, self.buf("enc.xs16", (B, T, idim)))
        e0 = self.buf("enc.e0", (B, T, d))
        self._lin_fwd(xa.view(B * T, idim), self.W("encoder.embed.0.weight"), st.p("encoder.embed.0.bias"), e0.view(B * T, d))
        ln0 = self._ln_fwd(e0, "encoder.embed.1", "enc.ln0", EMBED_LN_EPS)
        x0 = self.buf("enc.x0", (B, T, d))
        ops.scale_dropout(ln0, x0, math.sqrt(d), self.named_drop("enc.embed.drop", er), self.named_drop("enc.pos", epr))
        henc = self._conformer_fwd(x0, "encoder", hp["elayers"], H, hp["eunits"], hp["conformer_enc_kernel_size"], self.ilens_dev,
                                   er, epr, ear)
        # ---- post-encoder reduction (aas_vc.py:319-332): (B,T,d) -> (B,Tt,C)
        if T % pr == 0:
            hs = henc.view(B, Tt, C)
        else:
            hs = self.buf("hs.red", (B, Tt, C))
            hs.view(B, Tt * pr, d).copy_(henc[:, :Tt * pr])
        self.hs = hs
        self._dp_input_fwd(dp_inputs, Tt)
        if self.stochastic:
            return          # the stochastic predictor needs the MAS durations: it runs after the alignment search (_sdp_forward)
        self._dp_forward(Tt)

    def _dp_input_fwd(self, dp_inputs: torch.Tensor, Tt: int) -> torch.Tensor:
        """Duration-predictor input: Conv2dSubsampling projection + nearest interpolation to the encoder length (aas_vc.py:335-351,
        fastspeech_vc.py:247-260) -> self.dp_in (B, Tt, d)."""
        d = self.hp["adim"]
        B, Tdp = dp_inputs.shape[0], dp_inputs.shape[1]
        Tp = (((Tdp - 1) // 2) - 1) // 2
        proj = self._conv2d_sub_fwd(dp_inputs, "duration_predictor_projection", "duration_predictor_projection.out", "dpp")
        idx, ones, _, _ = self._interp_tables(Tp, Tt)
        dpi = self.buf("dp.in", (B, Tt, d))
        ops.gather_rows(proj.view(B, Tp, d), idx, ones, dpi)
        self.dp_in = dpi
        return dpi

    def _dp_forward(self, Tt: int) -> None:
        """DurationPredictor (duration_predictor.py:83-101) on self.dp_in -> self.dp_pre (B*Tt, 1) (log-domain pre-activation), self.dp_last."""
        hp, st = self.hp, self.store
        dpi = self.dp_in
        B = dpi.shape[0]
        k = hp["duration_predictor_kernel_size"]
        halo = (k - 1) // 2
        ch = hp["duration_predictor_chans"]
        cur = dpi
        for i in range(hp["duration_predictor_layers"]):
            ic = cur.shape[2]
            xp = ops.pad_rows(cur, self.buf(f"dp.pad{i}", (B, Tt + 2 * halo, ic)), halo)
            z = self._conv1d_fwd(xp, f"duration_predictor.conv.{i}.0", Tt, True, f"dp.c{i}")
            zu = ops.unpad_rows(z, self.buf(f"dp.zu{i}", (B, Tt, ch)), halo)
            nl = self._ln_fwd(zu, f"duration_predictor.conv.{i}.2", f"dp.ln{i}")
            drop = self.named_drop(f"dp.drop{i}", hp["duration_predictor_dropout_rate"])
            if drop.p > 0:
                nl = ops.scale_dropout(nl, self.buf(f"dp.do{i}", (B, Tt, ch)), 1.0, drop)
            cur = nl
        self.dp_last = cur
        self.dp_pre = self.buf("dp.pre", (B * Tt, 1))
        self._lin_fwd(cur.view(B * Tt, ch), self.W("duration_predictor.linear.weight"), st.p("duration_predictor.linear.bias"), self.dp_pre)

    def _decoder_side(self, ds: torch.Tensor, L: int):
        """Gaussian upsampling of self.hs with durations ds (B,Tt) to L frames, conformer decoder, feat_out, postnet."""
        hp, st = self.hp, self.store
        hs = self.hs
        B, Tt, C = hs.shape
        H, odim = hp["aheads"], hp["odim"]
        _, (dr_, dpr, dar) = self._rates()
        # ---- Gaussian upsampling (length_regulator.py:111-154)
        ldp = _r8(Tt)
        Pg = self.buf("up.P", (B, L, ldp))
        ops.gauss_weights(ds, self.olens_dev, self.tlens_dev, Pg)
        up = self.buf("up.out", (B, L, C))
        ops.gemm(Pg[..., :Tt], hs.transpose(1, 2), up, mode=self.mode)
        # ---- decoder: RelPositionalEncoding (x * sqrt(C), dropout) + conformer blocks (aas_vc.py:449-452)
        xd0 = self.buf("dec.x0", (B, L, C))
        ops.scale_dropout(up, xd0, math.sqrt(C), self.named_drop("dec.pos", dpr))
        zs = self._conformer_fwd(xd0, "decoder", hp["dlayers"], H, hp["dunits"], hp["conformer_dec_kernel_size"], self.olens_dev,
                                 dr_, dpr, dar)
        self.zs = zs
        before = self.buf("out.before", (B, L, odim))
        self._lin_fwd(zs.view(B * L, C), self.W("feat_out.weight"), st.p("feat_out.bias"), before.view(B * L, odim))
        after = self._postnet_fwd(before, lambda i: self.named_drop(f"post{i}", hp["postnet_dropout_rate"]))
        self.before, self.after = before, after
        return after, before

    def forward(self, xs: torch.Tensor, ys: torch.Tensor, dp_inputs: torch.Tensor, ilens: Optional[Sequence[int]] = None,
                olens: Optional[Sequence[int]] = None):
        """xs (B,T,idim), ys (B,L,odim), dp_inputs (B,T_dp,dp_idim) float32 device tensors already trimmed to the max
        lengths.  Returns (after (B,L,odim), before) in activation dtype; also sets self.log_p_attn (B,L,T_text) f32,
        self.ds (B,T_text) f32, self.d_outs (B,T_text) f32, self.paths, self.attn."""
        hp, st = self.hp, self.store
        B, T, idim = xs.shape
        L, odim = ys.shape[1], ys.shape[2]
        d, pr = hp["adim"], hp["post_encoder_reduction_factor"]
        C, Tt = d * pr, T // pr
        assert xs.dtype == _f32 and ys.dtype == _f32 and dp_inputs.dtype == _f32
        assert xs.is_contiguous() and ys.is_contiguous() and dp_inputs.is_contiguous()
        self._sig = (B, T, L, self.training)
        self.attn = {}
        self._last: Dict[str, torch.Tensor] = {}
        self.sync_shadow()
        if ilens is not None:
            self.prepare(B, T, L, ilens, olens)
        assert self._prepared == (B, T, L), "prepare(B, T, L, ilens, olens) must precede forward() for this batch shape"
        lens = self.buf("lens", (3, B), _i32)
        self.ilens_dev, self.tlens_dev, self.olens_dev = lens[0], lens[1], lens[2]
        self.shapes = dict(B=B, T=T, L=L, Tt=Tt, Tdp=dp_inputs.shape[1])
        self.xs, self.dp_inputs = xs, dp_inputs

        self._encoder_side(xs, dp_inputs)
        hs = self.hs

        self._alignment_and_mas(ys)
        if self.stochastic:
            self._sdp_forward()
        return self._decoder_side(self.ds, L)

    def _text_maskf(self, B: int, Tt: int) -> torch.Tensor:
        """(B*Tt,) float 0/1 text mask on the device (index bookkeeping on a tiny tensor, no model arithmetic)."""
        return (torch.arange(Tt, device=self.device)[None, :] < self.tlens_dev[:, None]).to(_f32).reshape(-1)

    def _sdp_forward(self) -> None:
        """dur_nll (B,) = StochasticDurationPredictor(dp_in^T, mask, w = ds) / sum(mask)  (models/aas_vc.py:412-419): the
        predictor's graph (kernels of this library sequenced by torch's tape) is kept for backward()."""
        from . import sdp

        s = self.shapes
        B, Tt = s["B"], s["Tt"]
        x32 = self.dp_in if self.dp_in.dtype == _f32 else ops.cast(self.dp_in, self.buf("sdp.x32", self.dp_in.shape, _f32))
        self._sdp_mask = self._text_maskf(B, Tt)
        e_q = sdp.randn((B, 2, Tt), self.device, self.base_seed, self.seed_dev, 7001)
        self._sdp_eq = e_q
        self.sdp.dropout_rate = self.hp["stochastic_duration_predictor_dropout_rate"] if self.training else 0.0
        with torch.enable_grad():
            self.sdp_nll = self.sdp.nll(x32, self.tlens_dev, self._sdp_mask, self.ds, e_q)
        self._sdp_norm = 1.0 / float(sum(self.tlens_host)) if getattr(self, "tlens_host", None) else None
        self.dur_nll = self.buf("sdp.dur_nll", (B,), _f32)
        self.dur_nll.zero_()
        ops.axpy(self.sdp_nll.detach(), self.dur_nll, self._sdp_norm)

    def _alignment_and_mas(self, ys: torch.Tensor) -> None:
        """Alignment module + monotonic alignment search on self.hs (B, T_text, C) and ys (B, L, odim); needs self.shapes,
        self.tlens_dev, self.olens_dev.  Fills self.log_p_attn, self.paths, self.ds, self.d_logp_mas, self.losses[2]."""
        hp, st = self.hp, self.store
        s = self.shapes
        B, L, Tt = s["B"], s["L"], s["Tt"]
        odim = hp["odim"]
        C = hp["adim"] * hp["post_encoder_reduction_factor"]
        hs = self.hs
        # ---- alignment module (alignments.py:28-60)
        ya = ys
        if self.bf16:
            ya = ops.cast(ys, self.buf("al.ys16", (B, L, odim)))
        tpad = ops.pad_rows(hs, self.buf("al.tpad", (B, Tt + 2, C)), 1)
        t1 = self._conv1d_fwd(tpad, "alignment_module.t_conv1", Tt, True, "al.t1")
        t1u = ops.unpad_rows(t1, self.buf("al.t1u", (B, Tt, C)), 1)
        text = self.buf("al.text", (B, Tt, C))
        self._lin_fwd(t1u.view(B * Tt, C), self.W("alignment_module.t_conv2.weight").view(C, C), st.p("alignment_module.t_conv2.bias"),
                      text.view(B * Tt, C))
        fpad = ops.pad_rows(ya, self.buf("al.fpad", (B, L + 2, odim)), 1)
        f1 = self._conv1d_fwd(fpad, "alignment_module.f_conv1", L, True, "al.f1")
        f2 = self._conv1d_fwd(f1, "alignment_module.f_conv2", L, True, "al.f2")
        f2u = ops.unpad_rows(f2, self.buf("al.f2u", (B, L, C)), 1)
        feats = self.buf("al.feats", (B, L, C))
        self._lin_fwd(f2u.view(B * L, C), self.W("alignment_module.f_conv3.weight").view(C, C), st.p("alignment_module.f_conv3.bias"),
                      feats.view(B * L, C))
        logp = self.buf("al.logp", (B, L, Tt), _f32)
        lse = self.buf("al.lse", (B, L), _f32)
        if self.mode == 1 and self.align_dot:
            # bf16 engine: the pairwise distances through the tensor cores, ||f - x||^2 = |f|^2 + |x|^2 - 2 f.x  (the direct-difference
            # kernel took 1.42 ms of the C3 step on the CUDA cores; this form ~0.1 ms).  The float32 parity path keeps the direct kernel.
            nf = ops.row_sqnorm(feats.view(B * L, C), self.buf("al.nf", (B * L,), _f32))
            nt = ops.row_sqnorm(text.view(B * Tt, C), self.buf("al.nt", (B * Tt,), _f32))
            ops.gemm(feats, text, logp, mode=1)
            ops.align_logp_from_dot(logp, nf, nt, self.tlens_dev, lse)
        else:
            ops.align_logp_fwd(feats, text, self.tlens_dev, logp, lse)
        self.log_p_attn = logp

        # ---- monotonic alignment search (alignments.py:281-310): durations + bin loss (+ its gradient), on the device
        self.paths = self.buf("mas.paths", (B, L), _i32)
        self.ds = self.buf("mas.ds", (B, Tt), _f32)
        self.d_logp_mas = self.buf("mas.dlogp", (B, L, Tt), _f32)
        self.d_logp_mas.zero_()
        ops.mas_into(logp, self.tlens_dev, self.olens_dev, self.paths, self.ds, self.losses[2:3], self.d_logp_mas,
                     self._mas_ws(B, L, Tt))

    @torch.no_grad()
    def inference(self, x: torch.Tensor, dp_input: torch.Tensor, y: Optional[torch.Tensor] = None, noise: Optional[torch.Tensor] = None):
        """AASVC.inference (aas_vc.py:531-603, _forward(is_inference=True) :371-398) for one utterance:
        x (T, idim), dp_input (T_dp, dp_idim) float32 device tensors -> (outs (L, odim) float32, d_outs (T_text,) int64).
        Eval-mode BatchNorm, no dropout; durations = clamp(round(exp(d) - 1), 0, 10); T_feats = sum(durations) is read back
        to the host once (the reference does the same in GaussianUpsampling: `ds.sum().int()`).
        With a ground-truth target y (L_y, odim) -- the "debug usage" the reference trainer's evaluation hook relies on
        (trainers/aas_vc.py:243-245) -- the alignment module and the alignment search also run on (encoder output, y) and the
        call returns (outs, d_outs, ds (T_text,) float32, log_p_attn (L_y, T_text) float32); outs still follow d_outs."""
        hp = self.hp
        T = x.shape[0]
        pr = hp["post_encoder_reduction_factor"]
        Tt = T // pr
        assert Tt >= 1, "source too short for the post-encoder reduction factor"
        was_training = self.training
        self.training = False
        try:
            # eval buffers are keyed by utterance length: drop the ones of earlier utterances
            for key in [k for k in self._bufs if isinstance(k[0], tuple) and len(k[0]) == 4 and k[0][3] is False]:
                del self._bufs[key]
            self._sig = (1, T, -1, False)
            self.attn = {}
            self._last = {}
            self.sync_shadow()
            mk = lambda v: torch.tensor([v], dtype=_i32).to(self.device)
            self.ilens_dev, self.tlens_dev = mk(T), mk(Tt)
            xs = x.to(_f32).contiguous().unsqueeze(0)
            dpi = dp_input.to(_f32).contiguous().unsqueeze(0)
            self._encoder_side(xs, dpi)
            gt = None
            if y is not None:
                Ly = y.shape[0]
                assert Ly >= Tt, "MAS needs at least one target frame per encoder position"
                self._sig = (1, T, -(Ly + 2), False)          # buffers of this side computation, dropped with the other eval buffers
                self.olens_dev = mk(Ly)
                self.shapes = dict(B=1, T=T, L=Ly, Tt=Tt, Tdp=dpi.shape[1])
                self._alignment_and_mas(y.to(_f32).contiguous().unsqueeze(0))
                gt = (self.ds[0].clone(), self.log_p_attn[0].clone())
                self._sig = (1, T, -1, False)
            ds = self.buf("inf.ds", (1, Tt), _f32)
            if self.stochastic:
                # d_outs = clamp(sdp(dp_in^T, mask, inverse=True, noise_scale), max=10)   (models/aas_vc.py:385-393)
                from . import sdp

                x32 = self.dp_in if self.dp_in.dtype == _f32 else self.dp_in.float()
                zn = sdp.randn((1, 2, Tt), self.device, self.base_seed, self.seed_dev, 7002) if noise is None else noise.to(self.device, _f32).reshape(1, 2, Tt)
                self.sdp.dropout_rate = 0.0
                ds.copy_(self.sdp.inverse(x32, self.tlens_dev, self._text_maskf(1, Tt), zn, hp["stochastic_duration_predictor_noise_scale"]))
            else:
                ops.duration_infer(self.dp_pre, ds)
            L = int(ds.sum().item())                          # the one host read-back of this path
            if L == 0:                                        # length_regulator.py:127-135 (all-zero prediction): every token gets one frame
                ds.fill_(1.0)
                L = Tt
            self._sig = (1, T, L, False)
            self.olens_dev = mk(L)
            self.shapes = dict(B=1, T=T, L=L, Tt=Tt, Tdp=dpi.shape[1])
            after, _ = self._decoder_side(ds, L)
            if gt is not None:
                return after[0].float().clone(), ds[0].to(torch.int64), gt[0], gt[1]
            return after[0].float().clone(), ds[0].to(torch.int64)
        finally:
            self.training = was_training

    def _mas_ws(self, B, L, Tt):
        n = ops.mas_workspace_bytes(B, L, Tt)
        return self.buf("mas.ws", (max(n, 8),), torch.uint8)

    # ------------------------------------------------------------------ losses
    def loss(self, ys: torch.Tensor, duration_loss: bool = True):
        """L1Loss + ForwardSumLoss + bin loss + DurationPredictorLoss as AASVCTrainer._train_step assembles them
        (trainers/aas_vc.py:73-134): total = l1 + lambda_align * (forward_sum + bin) + duration.  Fills self.losses
        (l1, forward_sum, bin, duration) and the gradients w.r.t. after / before / log_p_attn / duration pre-activation."""
        hp = self.hp
        B, L, odim = self.after.shape
        Tt = self.shapes["Tt"]
        lam = float(hp["lambda_align"])
        self.d_after = self.buf("loss.d_after", self.after.shape)
        self.d_before = self.buf("loss.d_before", self.after.shape)
        zl = self.buf("loss.zero_logits", (B, L), zero=True)
        zlab = self.buf("loss.zero_labels", (B, L), _f32, zero=True)
        dzl = self.buf("loss.d_logits", (B, L))
        ops.seq2seq_loss(self.after, self.before, zl, ys, zlab, self.olens_dev, 1.0, self._l1_pair, self.d_after, self.d_before, dzl,
                         self._loss_ws)
        self.losses[0:1].copy_(self._l1_pair[0:1])
        # forward-sum: writes lambda * d(fs)/d(logp) into d_logp, then += lambda * d(bin)/d(logp) from the MAS kernel
        self.d_logp = self.buf("loss.d_logp", (B, L, Tt), _f32)
        alpha_ws = self.buf("loss.alpha", (B, L, Tt), _f32)
        ops.forward_sum(self.log_p_attn, self.prior, self.tlens_dev, self.olens_dev, alpha_ws, self.losses[1:2], self.d_logp, lam)
        ops.axpy(self.d_logp_mas, self.d_logp, lam)
        if self.stochastic:
            # StochasticDurationPredictorLoss: duration_loss = sum(dur_nll) (trainers/aas_vc.py:125-127)
            self.losses[3:4].zero_()
            self._sdp_weight = 1.0 if duration_loss else 0.0
            if duration_loss:
                from . import _lib
                _lib.check(_lib.load().s2s_rowsum_acc(_lib.ptr(self.dur_nll), _lib.ptr(self.losses[3:4]), 1.0, 1, B, _lib.stream()), "rowsum_acc")
            return self.losses
        self.d_outs = self.buf("dp.d_outs", (B, Tt), _f32)
        self.d_dp_pre = self.buf("dp.d_pre", (B * Tt, 1))
        ops.duration_loss(self.dp_pre, self.ds, self.tlens_dev, self.d_outs, self.losses[3:4], self.d_dp_pre,
                          1.0 if duration_loss else 0.0)
        if not duration_loss:
            self.losses[3:4].zero_()       # trainers/aas_vc.py:131-133: before dp_train_start_steps the loss is 0.0, also in the logs
        return self.losses

    def optimizer_step(self, max_norm: float = 1.0, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                       grad_scale: float = 1.0, duration_predictor_active: bool = True) -> None:
        """clip_grad_norm_ + Adam over the flat buffers (trainers/aas_vc.py:151-158) with torch.optim.Adam's PER-PARAMETER
        clock: while the duration loss is off (steps <= dp_train_start_steps, aas_vc.py:119) the reference's duration
        predictor has no .grad, Adam skips it and its bias-correction step count starts later than everybody else's.  The
        predictor's span of the flat buffers therefore has its own device step counter and is left untouched while inactive."""
        st = self.store
        ops.step_advance(self.step_dev, self.seed_dev)
        if duration_predictor_active:
            ops.step_advance(self.dp_step_dev, None)
        self._sqn.zero_()
        ops.sqnorm(st.G, self._sqn)
        a, b = self._dp_span
        for lo, hi, step in ((0, a, self.step_dev), (a, b, self.dp_step_dev), (b, st.numel, self.step_dev)):
            if hi <= lo or (step is self.dp_step_dev and not duration_predictor_active):
                continue
            ops.adam_step(st.P[lo:hi], st.G[lo:hi], st.M[lo:hi], st.V[lo:hi], st.P16[lo:hi] if st.P16 is not None else None,
                          self.lr_dev, betas[0], betas[1], eps, weight_decay, step, self._sqn, max_norm, grad_scale)
        self.p16_dirty = False  # adam_step refreshed the bf16 shadow

    def total_loss(self) -> torch.Tensor:
        lam = float(self.hp["lambda_align"])
        return self.losses[0] + lam * (self.losses[1] + self.losses[2]) + self.losses[3]

    # ------------------------------------------------------------------ backward
    def forward_d_outs(self) -> torch.Tensor:
        """d_outs = min(pre * mask, 10) without any loss (the drop-in module's forward; aas_vc.py:408-411)."""
        B, Tt = self.shapes["B"], self.shapes["Tt"]
        self.d_outs = self.buf("dp.d_outs", (B, Tt), _f32)
        ops.duration_loss(self.dp_pre, self.ds, self.tlens_dev, self.d_outs, None, None)
        return self.d_outs

    def backward(self, d_after=None, d_before=None, d_logp=None, d_dp_pre=None, zero_grad: bool = True) -> None:
        """Accumulates parameter gradients into ParamStore.G.  Defaults: the gradients loss() left behind; the drop-in
        module passes the ones autograd hands it (d_logp already includes the bin-loss share)."""
        hp, st = self.hp, self.store
        d_after = self.d_after if d_after is None else d_after
        d_before = self.d_before if d_before is None else d_before
        d_logp = self.d_logp if d_logp is None else d_logp
        if not self.stochastic:
            d_dp_pre = self.d_dp_pre if d_dp_pre is None else d_dp_pre
        s = self.shapes
        B, T, L, Tt = s["B"], s["T"], s["L"], s["Tt"]
        d, H, pr, odim = hp["adim"], hp["aheads"], hp["post_encoder_reduction_factor"], hp["odim"]
        C = d * pr
        mode = self.mode
        if zero_grad:
            st.G.zero_()
        er, epr, ear = hp["transformer_enc_dropout_rate"], hp["transformer_enc_positional_dropout_rate"], hp["transformer_enc_attn_dropout_rate"]
        dr_, dpr, dar = hp["transformer_dec_dropout_rate"], hp["transformer_dec_positional_dropout_rate"], hp["transformer_dec_attn_dropout_rate"]

        # ---- postnet + feat_out
        dbefore = self._postnet_bwd(d_after, d_before, lambda i: self.named_drop(f"post{i}", hp["postnet_dropout_rate"]))
        gz = self._scratch("g.zs", (B, L, C))
        self._lin_bwd(dbefore.view(B * L, odim), self.zs.view(B * L, C), self.W("feat_out.weight"), st.g("feat_out.weight"),
                      st.g("feat_out.bias"), dx=gz.view(B * L, C))
        # ---- decoder
        xd0 = self.buf("dec.x0", (B, L, C))
        gx = self._conformer_bwd(gz, xd0, "decoder", hp["dlayers"], H, hp["dunits"], hp["conformer_dec_kernel_size"], dr_, dpr, dar)
        gup = self._scratch("g.up", (B, L, C))
        ops.scale_dropout(gx, gup, math.sqrt(C), self.named_drop("dec.pos", dpr))
        # ---- Gaussian upsampling: d_hs = P^T d_up   (ds carries no gradient: it is the integer MAS output)
        ldp = _r8(Tt)
        Pg = self.buf("up.P", (B, L, ldp))
        dhs = self._scratch("g.hs", (B, Tt, C))
        ops.gemm(Pg[..., :Tt].transpose(1, 2), gup.transpose(1, 2), dhs, mode=mode)

        # ---- duration predictor
        if self.stochastic:
            # the predictor sees a detached input (duration_predictor.py:236): its loss reaches its own parameters only, and the
            # input projection receives no gradient at all.  d_dp_pre: None = the fused step (weight from loss()), or the
            # (B,) gradient of dur_nll handed in by the drop-in module's autograd node
            self.sdp_backward(d_dp_pre)
            return self._backward_alignment_and_encoder(dhs, d_logp)
        self._dp_backward(d_dp_pre, Tt)

        return self._backward_alignment_and_encoder(dhs, d_logp)

    def _dp_backward(self, d_dp_pre: torch.Tensor, Tt: int) -> None:
        """Backward of the duration predictor and of its input projection (the side input itself needs no gradient)."""
        hp, st = self.hp, self.store
        s = self.shapes
        B, d = s["B"], hp["adim"]
        ch = hp["duration_predictor_chans"]
        k = hp["duration_predictor_kernel_size"]
        halo = (k - 1) // 2
        gcur = self._scratch("g.dp_a", (B, Tt, ch))
        self._lin_bwd(d_dp_pre, self.dp_last.view(B * Tt, ch), self.W("duration_predictor.linear.weight"),
                      st.g("duration_predictor.linear.weight"), st.g("duration_predictor.linear.bias"), dx=gcur.view(B * Tt, ch))
        for i in reversed(range(hp["duration_predictor_layers"])):
            ic = d if i == 0 else ch
            drop = self.named_drop(f"dp.drop{i}", hp["duration_predictor_dropout_rate"])
            if drop.p > 0:
                gcur = ops.scale_dropout(gcur, self._scratch("g.dp_b", (B, Tt, ch)), 1.0, drop)
            zu = self.buf(f"dp.zu{i}", (B, Tt, ch))
            gzu = self._scratch("g.dp_c", (B, Tt, ch))
            self._ln_bwd(gcur, zu, f"duration_predictor.conv.{i}.2", f"dp.ln{i}", gzu)
            ops.relu_bwd(gzu, zu, gzu, 1.0)
            gzp = ops.pad_rows(gzu, self._scratch("g.dp_zp", (B, Tt + 2 * halo, ch)), halo)
            xp = self.buf(f"dp.pad{i}", (B, Tt + 2 * halo, ic))
            gxp = self._scratch(f"g.dp_xp{i % 2}", (B, Tt + 2 * halo, ic))
            self._conv1d_bwd(gzp, xp, f"duration_predictor.conv.{i}.0", Tt, f"dp.c{i}", gxp)
            gnext = self._scratch(f"g.dp_n{i % 2}", (B, Tt, ic))
            ops.unpad_rows(gxp, gnext, halo)
            gcur = gnext
        gdpi = gcur                                                                                   # (B, Tt, d)
        Tdp = s["Tdp"]
        Tp = (((Tdp - 1) // 2) - 1) // 2
        _, _, first, cnt = self._interp_tables(Tp, Tt)
        gproj = self._scratch("g.dpproj", (B, Tp, d))
        ops.gather_rows(gdpi, first, cnt, gproj)
        self._conv2d_sub_bwd(gproj.view(B * Tp, d), self.dp_inputs, "duration_predictor_projection", "duration_predictor_projection.out", "dpp")

    def sdp_backward(self, g_dur_nll: Optional[torch.Tensor] = None) -> None:
        """Gradient of the stochastic predictor's loss into the flat gradient buffer (its leaves' .grad are views of it).
        g_dur_nll (B,): d loss / d dur_nll from the drop-in module's tape; None: the fused step, sum(dur_nll) * weight of loss()."""
        B = self.shapes["B"]
        if g_dur_nll is None:
            w = getattr(self, "_sdp_weight", 1.0)
            if w == 0.0:
                self.sdp_nll = None
                return
            g = torch.full((B,), w * self._sdp_norm, dtype=_f32, device=self.device)
        else:
            g = ops.axpy(g_dur_nll.to(_f32).contiguous(), torch.zeros(B, dtype=_f32, device=self.device), self._sdp_norm)
        for name, leaf in self._sdp_leaves.items():
            if leaf.grad is None or leaf.grad.data_ptr() != self.store.g(name).data_ptr():
                leaf.grad = self.store.g(name)
        self.sdp_nll.backward(gradient=g)
        self.sdp_nll = None

    def _backward_alignment_and_encoder(self, dhs: torch.Tensor, d_logp: torch.Tensor) -> None:
        """Tail of backward(): alignment module (d log_p_attn -> d feats, d text), post-encoder reduction, conformer encoder and
        input layer; dhs (B, Tt, C) is the gradient that reached the encoder output through the Gaussian upsampling."""
        hp, st = self.hp, self.store
        s = self.shapes
        B, T, L, Tt = s["B"], s["T"], s["L"], s["Tt"]
        d, H, pr, odim = hp["adim"], hp["aheads"], hp["post_encoder_reduction_factor"], hp["odim"]
        C = d * pr
        mode = self.mode
        er, epr, ear = hp["transformer_enc_dropout_rate"], hp["transformer_enc_positional_dropout_rate"], hp["transformer_enc_attn_dropout_rate"]
        # ---- alignment module: d(log_p_attn) -> d(feats), d(text)
        ldw = _r8(Tt)
        Wm = self._scratch("g.alW", (B, L, ldw))
        rowsum = self._scratch("g.alrow", (B, L), _f32)
        colsum = self._scratch("g.alcol", (B, Tt), _f32)
        ops.align_logp_bwd(d_logp, self.log_p_attn, self.buf("al.lse", (B, L), _f32), self.tlens_dev, Wm, rowsum, colsum)
        feats = self.buf("al.feats", (B, L, C))
        text = self.buf("al.text", (B, Tt, C))
        dfeats = self._scratch("g.alfeats", (B, L, C))
        dtext = self._scratch("g.altext", (B, Tt, C))
        ops.rowscale(feats, rowsum, dfeats)
        ops.gemm(Wm[..., :Tt], text.transpose(1, 2), dfeats, alpha=-1.0, residual=dfeats, mode=mode)
        ops.rowscale(text, colsum, dtext)
        ops.gemm(Wm[..., :Tt].transpose(1, 2), feats.transpose(1, 2), dtext, alpha=-1.0, residual=dtext, mode=mode)
        # feats path: f_conv3 (k1) <- relu f_conv2 (k3) <- relu f_conv1 (k3) <- ys
        f2u = self.buf("al.f2u", (B, L, C))
        gf2u = self._scratch("g.al_a", (B, L, C))
        self._lin_bwd(dfeats.view(B * L, C), f2u.view(B * L, C), self.W("alignment_module.f_conv3.weight").view(C, C),
                      st.g("alignment_module.f_conv3.weight").view(C, C), st.g("alignment_module.f_conv3.bias"), dx=gf2u.view(B * L, C),
                      dx_gate=f2u.view(B * L, C))                    # ReLU' of f_conv2 rides in the GEMM epilogue
        gf2p = ops.pad_rows(gf2u, self._scratch("g.al_p", (B, L + 2, C)), 1)
        f1 = self.buf("al.f1.z", (B, L + 2, C))
        gf1 = self._scratch("g.al_p2", (B, L + 2, C))
        self._conv1d_bwd(gf2p, f1, "alignment_module.f_conv2", L, "al.f2", gf1)
        ops.relu_bwd(gf1, f1, gf1, 1.0)
        self._conv1d_bwd(gf1, self.buf("al.fpad", (B, L + 2, odim)), "alignment_module.f_conv1", L, "al.f1", None)
        # text path: t_conv2 (k1) <- relu t_conv1 (k3) <- hs
        t1u = self.buf("al.t1u", (B, Tt, C))
        gt1u = self._scratch("g.al_a", (B, Tt, C))
        self._lin_bwd(dtext.view(B * Tt, C), t1u.view(B * Tt, C), self.W("alignment_module.t_conv2.weight").view(C, C),
                      st.g("alignment_module.t_conv2.weight").view(C, C), st.g("alignment_module.t_conv2.bias"), dx=gt1u.view(B * Tt, C),
                      dx_gate=t1u.view(B * Tt, C))
        gt1p = ops.pad_rows(gt1u, self._scratch("g.al_p", (B, Tt + 2, C)), 1)
        ghp = self._scratch("g.al_p2", (B, Tt + 2, C))
        self._conv1d_bwd(gt1p, self.buf("al.tpad", (B, Tt + 2, C)), "alignment_module.t_conv1", Tt, "al.t1", ghp)
        ghs_al = self._scratch("g.al_hs", (B, Tt, C))
        ops.unpad_rows(ghp, ghs_al, 1)
        ops.add(dhs, ghs_al, dhs)

        # ---- back through the post-encoder reduction into the encoder
        if T % pr == 0:
            genc = dhs.view(B, T, d)
        else:
            genc = self._scratch("g.enc_full", (B, T, d))
            genc.zero_()
            genc[:, :Tt * pr].copy_(dhs.view(B, Tt * pr, d))
        x0 = self.buf("enc.x0", (B, T, d))
        gx0 = self._conformer_bwd(genc, x0, "encoder", hp["elayers"], H, hp["eunits"], hp["conformer_enc_kernel_size"], er, epr, ear)
        gln0 = self._scratch("g.ln0", (B, T, d))
        ops.scale_dropout(gx0, gln0, math.sqrt(d), self.named_drop("enc.embed.drop", er), self.named_drop("enc.pos", epr))
        e0 = self.buf("enc.e0", (B, T, d))
        ge0 = self._scratch("g.e0", (B, T, d))
        self._ln_bwd(gln0, e0, "encoder.embed.1", "enc.ln0", ge0)
        xa = self.buf("enc.xs16", (B, T, hp["idim"])) if self.bf16 else self.xs
        self._lin_bwd(ge0.view(B * T, d), xa.view(B * T, hp["idim"]), self.W("encoder.embed.0.weight"), st.g("encoder.embed.0.weight"),
                      st.g("encoder.embed.0.bias"), dx=None)
