"""Stochastic duration predictor of AAS-VC on the C-ABI kernels (csrc/ops_sdp.cu).

Reference: StochasticDurationPredictor (seq2seq_vc/modules/duration_predictor.py:131-304) with DilatedDepthSeparableConv /
ConvFlow / ElementwiseAffineFlow / LogFlow / FlipFlow (modules/vits/flow.py:19-310) and the rational-quadratic spline
(modules/vits/transform.py:12-216), called as AASVC._forward does (models/aas_vc.py:385-393 inference, :412-419 training).
It is the default of the shipped recipe (egs/arctic/vc2/conf/aas_vc.melmelmel.v1.yaml:57).

The predictor sees its conditioning input DETACHED (duration_predictor.py:236) and the MAS durations as data, so it is an
isolated sub-network: its loss only reaches its own parameters.  Every layer op below is a kernel of libs2svc_b200.so wrapped
in a torch.autograd.Function (forward kernel / backward kernel pair); torch's tape only ORDERS those launches and sums the
fan-in of the few tensors that are used twice.  All float32 (logs, square roots and 1e-3 floors do not survive bf16); the
tensors are (B, T_text, C) small.  Layout: channels-last activations (B, T, C), flow state z (B, 2, T) with the FlipFlow folded
into which row the coupling writes.  Noise is an explicit input (the reference draws it inside forward); `randn` is the
library's own counter-based generator.
"""
from __future__ import annotations

import ctypes
import math
from typing import Callable, Dict, List, Tuple

import torch

from . import _lib, ops
from ._lib import NO_DROP, Drop, check, ptr, stream

_f32 = torch.float32
_i32 = torch.int32
BINS = 10
NPAR = 3 * BINS - 1
LN_EPS = 1e-5                                    # flow.py:139


def _L():
    return _lib.load()


def param_spec(hp: dict, prefix: str = "duration_predictor") -> List[Tuple[str, Tuple[int, ...]]]:
    """Reference state-dict names / shapes (duration_predictor.py:143-206), in registration order."""
    C, k, nl, n = hp["channels"], hp["kernel_size"], hp["dds_conv_layers"], hp["flows"]
    spec: List[Tuple[str, Tuple[int, ...]]] = []

    def conv(name, o, i, kk=1):
        spec.extend([(name + ".weight", (o, i, kk)), (name + ".bias", (o,))])

    def dds(name):
        for i in range(nl):
            conv(f"{name}.convs.{i}.0", C, 1, k)
            spec.extend([(f"{name}.convs.{i}.2.weight", (C,)), (f"{name}.convs.{i}.2.bias", (C,))])
            conv(f"{name}.convs.{i}.5", C, C)
            spec.extend([(f"{name}.convs.{i}.7.weight", (C,)), (f"{name}.convs.{i}.7.bias", (C,))])

    def flows(name):
        spec.extend([(f"{name}.0.m", (2, 1)), (f"{name}.0.logs", (2, 1))])
        for i in range(n):
            p = f"{name}.{1 + 2 * i}"
            conv(p + ".input_conv", C, 1)
            dds(p + ".dds_conv")
            conv(p + ".proj", NPAR, C)

    conv(prefix + ".pre", C, C)
    dds(prefix + ".dds")
    conv(prefix + ".proj", C, C)
    flows(prefix + ".flows")
    conv(prefix + ".post_pre", C, 1)
    dds(prefix + ".post_dds")
    conv(prefix + ".post_proj", C, C)
    flows(prefix + ".post_flows")
    return spec


def init_params(hp: dict, prefix: str, seed: int) -> Dict[str, torch.Tensor]:
    """torch-default Conv1d / LayerNorm initialisation; ConvFlow.proj starts at zero (flow.py:256-258: identity splines)."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    spec = param_spec(hp, prefix)
    shapes = dict(spec)
    for name, shape in spec:
        n = 1
        for s in shape:
            n *= s
        if name.endswith(".m") or name.endswith(".logs"):
            v = torch.zeros(shape)
        elif ".proj." in name and "flows" in name:
            v = torch.zeros(shape)
        elif name.endswith((".2.weight", ".7.weight")):
            v = torch.ones(shape)
        elif name.endswith((".2.bias", ".7.bias")):
            v = torch.zeros(shape)
        else:
            wshape = shapes[name[:-5] + ".weight"] if name.endswith(".bias") else shape
            fan_in = wshape[1] * wshape[2]
            v = ((torch.rand(n, generator=g) * 2 - 1) / math.sqrt(fan_in)).reshape(shape)
        out[name] = v
    return out


def randn(shape, device, seed: int, seed_dev: torch.Tensor, site: int) -> torch.Tensor:
    """Standard-normal draws from the library's counter-based generator (s2s_randn)."""
    out = torch.empty(shape, dtype=_f32, device=device)
    check(_L().s2s_randn(ptr(out), out.numel(), seed & 0xFFFFFFFFFFFFFFFF, ptr(seed_dev), site, stream()), "randn")
    return out


# =================================================================================================
# kernel pairs as autograd nodes
# =================================================================================================
class _Linear(torch.autograd.Function):
    """y (N, O) = x (N, K) W^T + b: s2s_gemm; backward = dX GEMM, dW GEMM, bias column sum."""

    @staticmethod
    def forward(ctx, x, w, b, mode):
        x = x.contiguous()
        w2 = w.reshape(w.shape[0], -1)
        y = torch.empty(x.shape[0], w2.shape[0], dtype=_f32, device=x.device)
        m = mode if (mode != 2 or w2.shape[0] >= 8) else 0        # the split-operand tcgen05 GEMM pads K itself; N >= 8 for TMA
        ops.gemm(x, w2, y, bias=b, mode=m)
        ctx.save_for_backward(x, w2)
        ctx.mode, ctx.wshape = m, w.shape
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w2 = ctx.saved_tensors
        dy = dy.contiguous()
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            ops.gemm(dy, w2.t(), dx, mode=ctx.mode)
        dw = torch.zeros_like(w2)
        ops.gemm(dy.t(), x.t(), dw, accumulate=True, mode=ctx.mode)
        db = torch.zeros(w2.shape[0], dtype=_f32, device=x.device)
        ops.colsum(dy, db)
        return dx, dw.reshape(ctx.wshape), db, None


class _Outer(torch.autograd.Function):
    """Conv1d(1 -> C, k = 1): y (N, C) = x (N) w^T + b   (s2s_outer_fwd / bwd)"""

    @staticmethod
    def forward(ctx, x, w, b):
        x = x.contiguous().view(-1)
        C = w.shape[0]
        w1 = w.reshape(C).contiguous()
        y = torch.empty(x.numel(), C, dtype=_f32, device=x.device)
        check(_L().s2s_outer_fwd(ptr(x), ptr(w1), ptr(b), ptr(y), x.numel(), C, stream()), "outer_fwd")
        ctx.save_for_backward(x, w1)
        ctx.wshape = w.shape
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w1 = ctx.saved_tensors
        dy = dy.contiguous()
        C = w1.numel()
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dw, db = torch.zeros_like(w1), torch.zeros_like(w1)
        check(_L().s2s_outer_bwd(ptr(dy), ptr(x), ptr(w1), ptr(dx), ptr(dw), ptr(db), x.numel(), C, stream()), "outer_bwd")
        return dx, dw.reshape(ctx.wshape), db


class _DWConv(torch.autograd.Function):
    """Dilated depthwise Conv1d on the masked input (s2s_dwconv_dilated_fwd / bwd); x (B, T, C), w (C, 1, K)."""

    @staticmethod
    def forward(ctx, x, w, b, tlens, dil):
        x = x.contiguous()
        B, T, C = x.shape
        K = w.shape[-1]
        w2 = w.reshape(C, K).contiguous()
        y = torch.empty_like(x)
        check(_L().s2s_dwconv_dilated_fwd(ptr(x), ptr(tlens), ptr(w2), ptr(b), ptr(y), B, T, C, K, dil, stream()), "dwconv_dilated_fwd")
        ctx.save_for_backward(x, w2, tlens)
        ctx.dil, ctx.wshape = dil, w.shape
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w2, tlens = ctx.saved_tensors
        dy = dy.contiguous()
        B, T, C = x.shape
        K = w2.shape[1]
        dx = torch.empty_like(x)
        dw, db = torch.zeros_like(w2), torch.zeros(C, dtype=_f32, device=x.device)
        check(_L().s2s_dwconv_dilated_bwd(ptr(dy), ptr(x), ptr(tlens), ptr(w2), ptr(dx), ptr(dw), ptr(db), B, T, C, K, ctx.dil, stream()),
              "dwconv_dilated_bwd")
        return dx, dw.reshape(ctx.wshape), db, None, None


class _LayerNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta):
        x = x.contiguous()
        rows = x.numel() // x.shape[-1]
        y = torch.empty_like(x)
        mean, rstd = torch.empty(rows, dtype=_f32, device=x.device), torch.empty(rows, dtype=_f32, device=x.device)
        ops.layernorm_fwd(x, gamma, beta, y, mean, rstd, LN_EPS)
        ctx.save_for_backward(x, gamma, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma, mean, rstd = ctx.saved_tensors
        dx = torch.empty_like(x)
        dg, db = torch.zeros_like(gamma), torch.zeros_like(gamma)
        ops.layernorm_bwd(dy.contiguous(), x, gamma, mean, rstd, dx, dg, db)
        return dx, dg, db


class _Gelu(torch.autograd.Function):
    """exact GELU followed by (optional) dropout"""

    @staticmethod
    def forward(ctx, x, drop):
        x = x.contiguous()
        y = torch.empty_like(x)
        check(_L().s2s_gelu_fwd(ptr(x), ptr(y), x.numel(), stream()), "gelu_fwd")
        if drop.p > 0.0:
            ops.dropout_bwd(y.view(-1, y.shape[-1]), y.view(-1, y.shape[-1]), drop)       # y *= mask / (1 - p): the same kernel both ways
        ctx.save_for_backward(x)
        ctx.drop = drop
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        dy = dy.contiguous()
        if ctx.drop.p > 0.0:
            g = torch.empty_like(dy)
            ops.dropout_bwd(dy.view(-1, dy.shape[-1]), g.view(-1, g.shape[-1]), ctx.drop)
            dy = g
        dx = torch.empty_like(x)
        check(_L().s2s_gelu_bwd(ptr(dy), ptr(x), ptr(dx), x.numel(), stream()), "gelu_bwd")
        return dx, None


class _Add(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        return ops.add(a.contiguous(), b.contiguous(), torch.empty_like(a))

    @staticmethod
    def backward(ctx, g):
        return g, g


class _RowMask(torch.autograd.Function):
    """x (B, T, C) * mask (B, T): s2s_rowscale both ways"""

    @staticmethod
    def forward(ctx, x, maskf):
        ctx.save_for_backward(maskf)
        return ops.rowscale(x.contiguous(), maskf, torch.empty_like(x))

    @staticmethod
    def backward(ctx, g):
        (maskf,) = ctx.saved_tensors
        return ops.rowscale(g.contiguous(), maskf, torch.empty_like(g)), None


class _Affine(torch.autograd.Function):
    """ElementwiseAffineFlow forward direction; threads the per-utterance accumulator nll through"""

    @staticmethod
    def forward(ctx, z, m, logs, tlens, nll, sign):
        z = z.contiguous()
        B, _, T = z.shape
        y, out = torch.empty_like(z), nll.clone()
        check(_L().s2s_sdp_affine_fwd(ptr(z), ptr(m), ptr(logs), ptr(tlens), ptr(y), ptr(out), sign, B, T, 0, stream()), "sdp_affine_fwd")
        ctx.save_for_backward(z, logs, tlens)
        ctx.sign = sign
        return y, out

    @staticmethod
    def backward(ctx, gy, gnll):
        z, logs, tlens = ctx.saved_tensors
        B, _, T = z.shape
        dz = torch.empty_like(z)
        dm, dl = torch.zeros(2, 1, dtype=_f32, device=z.device), torch.zeros(2, 1, dtype=_f32, device=z.device)
        check(_L().s2s_sdp_affine_bwd(ptr(z), ptr(logs), ptr(tlens), ptr(gy.contiguous()), ptr(gnll.contiguous()), ctx.sign, ptr(dz), ptr(dm), ptr(dl),
                                      B, T, stream()), "sdp_affine_bwd")
        return dz, dm, dl, None, gnll, None


class _Coupling(torch.autograd.Function):
    """ConvFlow's spline coupling + FlipFlow (flow.py:263-310,79-93): z = (xa, xb), h = spline parameters predicted from xa;
    z' = (spline(xb; h), xa) * mask, nll += sign * sum_t log|det|."""

    @staticmethod
    def forward(ctx, z, h, tlens, nll, sign, hidden):
        z, h = z.contiguous(), h.contiguous()
        B, _, T = z.shape
        out, lad, acc = torch.empty_like(z), torch.empty(B, T, dtype=_f32, device=z.device), nll.clone()
        L = _L()
        check(L.s2s_rq_spline_fwd(ptr(z) + 4 * T, 2 * T, ptr(h), ptr(tlens), ptr(out), 2 * T, ptr(lad), B, T, hidden, 0, stream()), "rq_spline_fwd")
        out[:, 1].copy_(z[:, 0])                  # xa passes through (already masked); the flip puts it second
        check(L.s2s_rowsum_acc(ptr(lad), ptr(acc), sign, B, T, stream()), "rowsum_acc")
        ctx.save_for_backward(z, h, tlens)
        ctx.sign, ctx.hidden = sign, hidden
        return out, acc

    @staticmethod
    def backward(ctx, gout, gnll):
        z, h, tlens = ctx.saved_tensors
        B, _, T = z.shape
        gout, gnll = gout.contiguous(), gnll.contiguous()
        glad = torch.empty(B, T, dtype=_f32, device=z.device)
        L = _L()
        check(L.s2s_rowbcast(ptr(gnll), ptr(glad), ctx.sign, B, T, stream()), "rowbcast")
        dz, dh = torch.empty_like(z), torch.empty_like(h)
        check(L.s2s_rq_spline_bwd(ptr(z) + 4 * T, 2 * T, ptr(h), ptr(tlens), ptr(gout), 2 * T, ptr(glad), ptr(dz) + 4 * T, 2 * T, ptr(dh), B, T,
                                  ctx.hidden, stream()), "rq_spline_bwd")
        dz[:, 0].copy_(gout[:, 1])                # the pass-through row (the tape adds the gradient that arrives through h)
        return dz, dh, None, gnll, None, None


class _Head(torch.autograd.Function):
    @staticmethod
    def forward(ctx, zq, w, tlens, nll):
        zq = zq.contiguous()
        B, _, T = zq.shape
        out, acc = torch.empty_like(zq), nll.clone()
        check(_L().s2s_sdp_head_fwd(ptr(zq), ptr(w), ptr(tlens), ptr(out), ptr(acc), B, T, stream()), "sdp_head_fwd")
        ctx.save_for_backward(zq, w, tlens)
        return out, acc

    @staticmethod
    def backward(ctx, gout, gnll):
        zq, w, tlens = ctx.saved_tensors
        B, _, T = zq.shape
        dz = torch.empty_like(zq)
        check(_L().s2s_sdp_head_bwd(ptr(zq), ptr(w), ptr(tlens), ptr(gout.contiguous()), ptr(gnll.contiguous()), ptr(dz), B, T, stream()), "sdp_head_bwd")
        return dz, None, None, gnll


class _Gauss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, tlens, nll, sign):
        z = z.contiguous()
        B, _, T = z.shape
        acc = nll.clone()
        check(_L().s2s_sdp_gauss_fwd(ptr(z), ptr(tlens), ptr(acc), sign, B, T, stream()), "sdp_gauss_fwd")
        ctx.save_for_backward(z, tlens)
        ctx.sign = sign
        return acc

    @staticmethod
    def backward(ctx, gnll):
        z, tlens = ctx.saved_tensors
        B, _, T = z.shape
        dz = torch.empty_like(z)
        check(_L().s2s_sdp_gauss_bwd(ptr(z), ptr(tlens), ptr(gnll.contiguous()), ctx.sign, ptr(dz), 0, B, T, stream()), "sdp_gauss_bwd")
        return dz, None, gnll, None


# =================================================================================================
# the predictor
# =================================================================================================
class StochasticDurationPredictor:
    """Functional over a parameter getter `P(name) -> float32 tensor` (views into the engine's flat store or nn.Parameters)."""

    def __init__(self, hp: dict, prefix: str, P: Callable[[str], torch.Tensor], gemm_mode: int = 2, dropout_rate: float = 0.5,
                 drop_of: Callable[[str, float], Drop] = None):
        self.hp, self.prefix, self.P, self.mode = hp, prefix, P, gemm_mode
        self.dropout_rate = dropout_rate
        self.drop_of = drop_of or (lambda name, p: NO_DROP)

    # ---- building blocks
    def _lin(self, name, x2d):
        return _Linear.apply(x2d, self.P(name + ".weight"), self.P(name + ".bias"), self.mode)

    def _dds(self, name, x, tlens, maskf, g=None, dropout: float = 0.0):
        """flow.py:192-211: x (+ g); per layer x += drop(GELU(LN(pw(GELU(LN(dw_dilated(x mask))))))); x mask at the end."""
        B, T, C = x.shape
        k = self.hp["kernel_size"]
        if g is not None:
            x = _Add.apply(x, g)
        for i in range(self.hp["dds_conv_layers"]):
            p = f"{name}.convs.{i}"
            y = _DWConv.apply(x, self.P(p + ".0.weight"), self.P(p + ".0.bias"), tlens, k ** i)
            y = _Gelu.apply(_LayerNorm.apply(y, self.P(p + ".2.weight"), self.P(p + ".2.bias")), NO_DROP)
            y = self._lin(p + ".5", y.view(B * T, C)).view(B, T, C)
            y = _Gelu.apply(_LayerNorm.apply(y, self.P(p + ".7.weight"), self.P(p + ".7.bias")), self.drop_of(p, dropout))
            x = _Add.apply(x, y)
        return _RowMask.apply(x, maskf)

    def _condition(self, x, tlens, maskf):
        B, T, C = x.shape
        h = self._lin(self.prefix + ".pre", x.detach().reshape(B * T, C)).view(B, T, C)        # stop gradient (duration_predictor.py:236)
        h = self._dds(self.prefix + ".dds", h, tlens, maskf, dropout=self.dropout_rate)
        h = self._lin(self.prefix + ".proj", h.view(B * T, C)).view(B, T, C)
        return _RowMask.apply(h, maskf)

    def _conv_flow(self, name, z, g, tlens, maskf, nll, sign):
        B, _, T = z.shape
        C = self.hp["channels"]
        h = _Outer.apply(z[:, 0].reshape(B * T), self.P(name + ".input_conv.weight"), self.P(name + ".input_conv.bias")).view(B, T, C)
        h = self._dds(name + ".dds_conv", h, tlens, maskf, g=g)
        h = self._lin(name + ".proj", h.view(B * T, C)).view(B, T, NPAR)
        return _Coupling.apply(z, h, tlens, nll, sign, float(C))

    def _flow_stack(self, name, z, g, tlens, maskf, nll):
        """[ElementwiseAffine, (ConvFlow, Flip) x n] forward; log-determinants enter nll with sign -1."""
        z, nll = _Affine.apply(z, self.P(name + ".0.m"), self.P(name + ".0.logs"), tlens, nll, -1.0)
        for i in range(self.hp["flows"]):
            z, nll = self._conv_flow(f"{name}.{1 + 2 * i}", z, g, tlens, maskf, nll, -1.0)
        return z, nll

    # ---- the two directions
    def nll(self, x, tlens, maskf, w, e_q):
        """Negative variational lower bound per utterance (duration_predictor.py:243-289).
        x (B, T, C) conditioning, tlens (B) int32, maskf (B*T) float 0/1, w (B, T) durations, e_q (B, 2, T) standard-normal noise."""
        B, T, C = x.shape
        p = self.prefix
        xc = self._condition(x, tlens, maskf)
        hw = _Outer.apply(w.reshape(B * T), self.P(p + ".post_pre.weight"), self.P(p + ".post_pre.bias")).view(B, T, C)
        hw = self._dds(p + ".post_dds", hw, tlens, maskf, dropout=self.dropout_rate)
        hw = _RowMask.apply(self._lin(p + ".post_proj", hw.view(B * T, C)).view(B, T, C), maskf)
        nll = torch.zeros(B, dtype=_f32, device=x.device)
        nll = _Gauss.apply(e_q, tlens, nll, -1.0)                                   # logq's Gaussian term
        z, nll = self._flow_stack(p + ".post_flows", e_q, _Add.apply(xc, hw), tlens, maskf, nll)
        z, nll = _Head.apply(z, w.contiguous(), tlens, nll)
        z, nll = self._flow_stack(p + ".flows", z, xc, tlens, maskf, nll)
        return _Gauss.apply(z, tlens, nll, 1.0)

    @torch.no_grad()
    def inverse(self, x, tlens, maskf, z, noise_scale: float = 0.8, clamp_max: float = 10.0):
        """Durations from noise (duration_predictor.py:290-304, models/aas_vc.py:385-393): flows reversed, the first ConvFlow
        skipped, dur = min(ceil(exp(z0) mask), clamp_max).  z (B, 2, T) standard-normal noise."""
        B, T, C = x.shape
        p, L = self.prefix, _L()
        xc = self._condition(x, tlens, maskf)
        n = self.hp["flows"]
        z = ops.axpy(z.contiguous(), torch.zeros_like(z), noise_scale)
        for i in reversed(range(1, n)):
            name = f"{p}.flows.{1 + 2 * i}"
            # Flip first: (a, b) -> (b, a); then the coupling inverts its second row from its first
            h = _Outer.apply(z[:, 1].reshape(B * T), self.P(name + ".input_conv.weight"), self.P(name + ".input_conv.bias")).view(B, T, C)
            h = self._dds(name + ".dds_conv", h, tlens, maskf, g=xc)
            h = self._lin(name + ".proj", h.view(B * T, C)).view(B, T, NPAR).contiguous()
            out = torch.empty_like(z)
            check(L.s2s_rq_spline_fwd(ptr(z), 2 * T, ptr(h), ptr(tlens), ptr(out) + 4 * T, 2 * T, None, B, T, float(C), 1, stream()), "rq_spline_fwd")
            out[:, 0].copy_(ops.rowscale(z[:, 1].contiguous().view(B * T, 1), maskf, torch.empty(B * T, 1, dtype=_f32, device=z.device)).view(B, T))
            z = out
        zf = torch.empty_like(z)
        zf[:, 0].copy_(z[:, 1])
        zf[:, 1].copy_(z[:, 0])
        y = torch.empty_like(zf)
        check(L.s2s_sdp_affine_fwd(ptr(zf), ptr(self.P(p + ".flows.0.m")), ptr(self.P(p + ".flows.0.logs")), ptr(tlens), ptr(y), None, 0.0, B, T, 1,
                                   stream()), "sdp_affine_fwd")
        dur = torch.empty(B, T, dtype=_f32, device=z.device)
        check(L.s2s_sdp_durations(ptr(y), ptr(tlens), ptr(dur), clamp_max, B, T, stream()), "sdp_durations")
        return dur
