"""CPU restatement of the stochastic duration predictor of AAS-VC.

TEST INFRASTRUCTURE ONLY (tests/ may import it; the product must not).  Functional, state-dict driven restatement of
  * StochasticDurationPredictor.forward, both directions   (seq2seq_vc/modules/duration_predictor.py:131-304)
  * DilatedDepthSeparableConv, ConvFlow, ElementwiseAffineFlow, LogFlow, FlipFlow   (modules/vits/flow.py:19-310)
  * the piecewise rational-quadratic spline with linear tails   (modules/vits/transform.py:12-216)
and of the call sites in AASVC._forward (models/aas_vc.py:385-393 inference, :412-419 training: nll / sum(mask)).

Differences in FORM (not in arithmetic) from the reference, chosen so that a CUDA kernel can follow the same recipe:
  * the Gaussian noise is an explicit argument (`e_q` for the posterior flows, `z` for the inverse pass) -- the reference
    draws it with torch.randn inside forward; the fixtures record the reference's draw (oracle/gen_golden.py);
  * the spline is evaluated densely with clamped inputs and a select, instead of boolean-mask gather / scatter of the
    in-range elements (transform.py:61-93); out-of-range elements are the identity with log|det| = 0 in both.
    One reference quirk is NOT reproduced: when no element at all lies inside [-5, 5] the reference's masked gather is empty
    and torch.min() raises (transform.py:116); the dense form returns the identity for every element instead.
Dropout is not restated (the oracle is the dropout-free function, as for the other oracles).

Pinned against the live reference: tests/golden/sdp_tiny.npz (nll, every parameter gradient, inverse durations),
tests/test_oracle_golden.py::test_sdp_oracle_*.
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import numpy as np
import torch
import torch.nn.functional as F

MIN_BIN_WIDTH = MIN_BIN_HEIGHT = MIN_DERIVATIVE = 1e-3      # transform.py:9-11
BINS, TAIL_BOUND = 10, 5.0                                  # flow.py:223-224
LN_EPS = 1e-5                                               # flow.py:139
LOG_EPS = 1e-5                                              # flow.py:60


def conv1x1(sd, name, x):
    return F.conv1d(x, sd[name + ".weight"], sd[name + ".bias"])


def dds_conv(sd: Dict[str, torch.Tensor], prefix: str, x, x_mask, kernel_size: int, layers: int, g=None):
    """flow.py:192-211: x += GELU(LN(pw(GELU(LN(dw_dilated(x * mask)))))) per layer, masked at the end."""
    if g is not None:
        x = x + g
    C = x.shape[1]
    for i in range(layers):
        dil = kernel_size ** i
        pad = (kernel_size * dil - dil) // 2
        p = f"{prefix}.convs.{i}"
        y = F.conv1d(x * x_mask, sd[p + ".0.weight"], sd[p + ".0.bias"], padding=pad, dilation=dil, groups=C)
        y = F.gelu(F.layer_norm(y.transpose(1, 2), (C,), sd[p + ".2.weight"], sd[p + ".2.bias"], LN_EPS).transpose(1, 2))
        y = conv1x1(sd, p + ".5", y)
        y = F.gelu(F.layer_norm(y.transpose(1, 2), (C,), sd[p + ".7.weight"], sd[p + ".7.bias"], LN_EPS).transpose(1, 2))
        x = x + y
    return x * x_mask


def _knots(unnorm, lo: float, hi: float, min_size: float):
    """softmax -> floor at min_size -> cumulative knots on [lo, hi] with the ends pinned (transform.py:117-125,128-135)."""
    nb = unnorm.shape[-1]
    w = min_size + (1 - min_size * nb) * F.softmax(unnorm, dim=-1)
    cum = F.pad(torch.cumsum(w, dim=-1), (1, 0), value=0.0)
    cum = (hi - lo) * cum + lo
    cum = torch.cat([torch.full_like(cum[..., :1], lo), cum[..., 1:-1], torch.full_like(cum[..., :1], hi)], dim=-1)
    return cum, cum[..., 1:] - cum[..., :-1]


def rq_spline_linear_tails(x, uw, uh, ud, inverse: bool, tail_bound: float = TAIL_BOUND):
    """transform.py:44-93 (linear tails) around :96-209 (the spline).  x (...), uw / uh (..., bins), ud (..., bins - 1).
    Returns (y, log|det|); for the inverse direction log|det| is already negated as in the reference."""
    inside = (x >= -tail_bound) & (x <= tail_bound)
    const = float(np.log(np.exp(1 - MIN_DERIVATIVE) - 1))           # boundary derivative = 1 (transform.py:66-68)
    ud = torch.cat([torch.full_like(ud[..., :1], const), ud, torch.full_like(ud[..., :1], const)], dim=-1)
    xc = x.clamp(-tail_bound, tail_bound)
    cumw, widths = _knots(uw, -tail_bound, tail_bound, MIN_BIN_WIDTH)
    cumh, heights = _knots(uh, -tail_bound, tail_bound, MIN_BIN_HEIGHT)
    deriv = MIN_DERIVATIVE + F.softplus(ud)
    locs = (cumh if inverse else cumw).clone()
    locs[..., -1] += 1e-6                                           # _searchsorted (transform.py:212-216)
    idx = (torch.sum(xc[..., None] >= locs, dim=-1) - 1)[..., None]
    pick = lambda t: t.gather(-1, idx)[..., 0]
    in_cumw, in_w, in_cumh, in_h = pick(cumw), pick(widths), pick(cumh), pick(heights)
    delta = heights / widths
    in_delta, d0, d1 = pick(delta), pick(deriv), pick(deriv[..., 1:])
    if inverse:
        t = (xc - in_cumh) * (d0 + d1 - 2 * in_delta)
        a = t + in_h * (in_delta - d0)
        b = in_h * d0 - t
        c = -in_delta * (xc - in_cumh)
        root = (2 * c) / (-b - torch.sqrt(b.pow(2) - 4 * a * c))
        y = root * in_w + in_cumw
        th = root
    else:
        th = (xc - in_cumw) / in_w
    tt = th * (1 - th)
    denom = in_delta + (d0 + d1 - 2 * in_delta) * tt
    if not inverse:
        y = in_cumh + in_h * (in_delta * th.pow(2) + d0 * tt) / denom
    num = in_delta.pow(2) * (d1 * th.pow(2) + 2 * in_delta * tt + d0 * (1 - th).pow(2))
    lad = torch.log(num) - 2 * torch.log(denom)
    if inverse:
        lad = -lad
    return torch.where(inside, y, x), torch.where(inside, lad, torch.zeros_like(lad))


def conv_flow(sd, prefix, x, x_mask, g, hidden: int, kernel_size: int, layers: int, inverse: bool = False):
    """flow.py:263-310: the second channel goes through a spline whose parameters are predicted from the first."""
    xa, xb = x.split(x.size(1) // 2, 1)
    h = conv1x1(sd, prefix + ".input_conv", xa)
    h = dds_conv(sd, prefix + ".dds_conv", h, x_mask, kernel_size, layers, g=g)
    h = conv1x1(sd, prefix + ".proj", h) * x_mask
    b, c, t = xa.shape
    h = h.reshape(b, c, -1, t).permute(0, 1, 3, 2)
    den = math.sqrt(hidden)
    xb, lad = rq_spline_linear_tails(xb, h[..., :BINS] / den, h[..., BINS:2 * BINS] / den, h[..., 2 * BINS:], inverse)
    y = torch.cat([xa, xb], 1) * x_mask
    return y, torch.sum(lad * x_mask, [1, 2])


def affine_flow(sd, prefix, x, x_mask, inverse: bool = False):
    """flow.py:96-112."""
    m, logs = sd[prefix + ".m"], sd[prefix + ".logs"]
    if inverse:
        return (x - m) * torch.exp(-logs) * x_mask, None
    return (m + torch.exp(logs) * x) * x_mask, torch.sum(logs * x_mask, [1, 2])


def _flow_stack(sd, prefix, z, x_mask, g, hp, n_flows):
    """[ElementwiseAffine, (ConvFlow, Flip) x n] forward with summed log-determinants (duration_predictor.py:171-181,256-258)."""
    z, total = affine_flow(sd, f"{prefix}.0", z, x_mask)
    for i in range(n_flows):
        z, ld = conv_flow(sd, f"{prefix}.{1 + 2 * i}", z, x_mask, g, hp["channels"], hp["kernel_size"], hp["dds_conv_layers"])
        total = total + ld
        z = torch.flip(z, [1])
    return z, total


def _condition(sd, prefix, hp, x, x_mask):
    x = conv1x1(sd, prefix + ".pre", x.detach())                                  # stop gradient (duration_predictor.py:236)
    x = dds_conv(sd, prefix + ".dds", x, x_mask, hp["kernel_size"], hp["dds_conv_layers"])
    return conv1x1(sd, prefix + ".proj", x) * x_mask


def sdp_nll(sd, prefix: str, hp: dict, x, x_mask, w, e_q):
    """Negative variational lower bound of the durations, per utterance (duration_predictor.py:243-289).
    x (B, C, T) conditioning, x_mask (B, 1, T) float, w (B, 1, T) durations, e_q (B, 2, T) standard-normal noise."""
    x = _condition(sd, prefix, hp, x, x_mask)
    h_w = conv1x1(sd, prefix + ".post_pre", w)
    h_w = dds_conv(sd, prefix + ".post_dds", h_w, x_mask, hp["kernel_size"], hp["dds_conv_layers"])
    h_w = conv1x1(sd, prefix + ".post_proj", h_w) * x_mask
    e_q = e_q * x_mask
    z_q, logdet_q = _flow_stack(sd, prefix + ".post_flows", e_q, x_mask, x + h_w, hp, hp["flows"])
    z_u, z1 = torch.split(z_q, [1, 1], 1)
    u = torch.sigmoid(z_u) * x_mask
    z0 = (w - u) * x_mask
    logdet_q = logdet_q + torch.sum((F.logsigmoid(z_u) + F.logsigmoid(-z_u)) * x_mask, [1, 2])
    logq = torch.sum(-0.5 * (math.log(2 * math.pi) + e_q ** 2) * x_mask, [1, 2]) - logdet_q
    y0 = torch.log(torch.clamp_min(z0, LOG_EPS)) * x_mask                         # LogFlow (flow.py:62-65)
    logdet = torch.sum(-y0, [1, 2])
    z, ld = _flow_stack(sd, prefix + ".flows", torch.cat([y0, z1], 1), x_mask, x, hp, hp["flows"])
    logdet = logdet + ld
    nll = torch.sum(0.5 * (math.log(2 * math.pi) + z ** 2) * x_mask, [1, 2]) - logdet
    return nll + logq


def sdp_inverse(sd, prefix: str, hp: dict, x, x_mask, z, noise_scale: float = 0.8):
    """Durations from noise (duration_predictor.py:290-304): the flows reversed, the first ConvFlow ("useless vflow")
    skipped, dur = ceil(exp(z0) * mask).  z (B, 2, T) standard-normal noise (scaled here by noise_scale)."""
    x = _condition(sd, prefix, hp, x, x_mask)
    n = hp["flows"]
    z = z * noise_scale
    # reversed(flows) = [Flip, Conv_{n-1}, ..., Flip, Conv_0, Affine]; "flows[:-2] + [flows[-1]]" drops Conv_0
    for i in reversed(range(1, n)):
        z = torch.flip(z, [1])
        z, _ = conv_flow(sd, f"{prefix}.flows.{1 + 2 * i}", z, x_mask, x, hp["channels"], hp["kernel_size"], hp["dds_conv_layers"], inverse=True)
    z = torch.flip(z, [1])
    z, _ = affine_flow(sd, prefix + ".flows.0", z, x_mask, inverse=True)
    logw = z[:, :1]
    return torch.ceil(torch.exp(logw) * x_mask)


def aasvc_dur_nll(sd, prefix, hp, dp_inputs, text_lens: List[int], ds, e_q):
    """models/aas_vc.py:412-419: dur_nll (B,) = sdp(dp_inputs^T, mask, w = ds) / sum(mask); the trainer adds its sum to the
    loss (trainers/aas_vc.py:125-127).  dp_inputs (B, T_text, C), ds (B, T_text) integer durations from MAS."""
    T = dp_inputs.shape[1]
    mask = (torch.arange(T)[None, :] < torch.tensor(text_lens)[:, None]).to(dp_inputs.dtype)[:, None, :]
    nll = sdp_nll(sd, prefix, hp, dp_inputs.transpose(1, 2), mask, ds.to(dp_inputs.dtype)[:, None, :], e_q)
    return nll / mask.sum()


def aasvc_dur_inference(sd, prefix, hp, dp_inputs, text_lens: List[int], z, noise_scale: float = 0.8, max_dur: float = 10.0):
    """models/aas_vc.py:385-393: d_outs = clamp(sdp(..., inverse=True, noise_scale), max=MAX_DP_OUTPUT)."""
    T = dp_inputs.shape[1]
    mask = (torch.arange(T)[None, :] < torch.tensor(text_lens)[:, None]).to(dp_inputs.dtype)[:, None, :]
    return torch.clamp(sdp_inverse(sd, prefix, hp, dp_inputs.transpose(1, 2), mask, z, noise_scale).squeeze(1), max=max_dur)


def state_dict_spec(hp: dict, prefix: str = "duration_predictor") -> List[Tuple[str, Tuple[int, ...]]]:
    C, k, L, n = hp["channels"], hp["kernel_size"], hp["dds_conv_layers"], hp["flows"]
    spec: List[Tuple[str, Tuple[int, ...]]] = []

    def conv(name, o, i, kk=1):
        spec.extend([(name + ".weight", (o, i, kk)), (name + ".bias", (o,))])

    def dds(name):
        for i in range(L):
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
            conv(p + ".proj", 3 * BINS - 1, C)

    conv(prefix + ".pre", C, C)
    dds(prefix + ".dds")
    conv(prefix + ".proj", C, C)
    flows(prefix + ".flows")
    conv(prefix + ".post_pre", C, 1)
    dds(prefix + ".post_dds")
    conv(prefix + ".post_proj", C, C)
    flows(prefix + ".post_flows")
    return spec
