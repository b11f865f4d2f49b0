"""Tensor-level wrappers over the C ABI (include/s2svc_b200.h).

Each function takes torch CUDA tensors as *memory handles* (pointer + shape + strides), checks
layouts, and issues exactly the C call; no arithmetic happens in PyTorch.  GEMM_MODE selects the
s2s_gemm path: 0 = fp32 CUDA-core (parity), 1 = bf16 tcgen05.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import NO_DROP, ColsumDesc, Drop, GemmDesc, check, dt, ptr, stream

_i32 = torch.int32


def _L():
    return _lib.load()


def _batch_strides(t: torch.Tensor, nb: int) -> Tuple[int, int, int, int]:
    """(size1, size2, stride1, stride2) of the (up to two) leading batch dims of t."""
    if nb == 0:
        return 1, 1, 0, 0
    if nb == 1:
        return 1, t.shape[0], 0, t.stride(0)
    return t.shape[0], t.shape[1], t.stride(0), t.stride(1)


def _lead_strides(t: torch.Tensor, core: int) -> Tuple[int, int]:
    """Strides of the leading batch dims of an operand (missing dims broadcast with stride 0)."""
    n = t.dim() - core
    if n == 0:
        return 0, 0
    if n == 1:
        return 0, t.stride(0)
    assert n == 2
    return t.stride(0), t.stride(1)


# mode 2: 3 = two-way bf16 split (a0 b0 + a1 b0 + a0 b1, ~2^-16 per product), 6 = three-way split.  Measured on C1
# (tools/parity_modes.py, profiles/r02_parity_modes_c1.json): both give mel L1 1.2e-5 / 1.5e-5 against the CPU reference restatement (CUDA-core
# fp32: 1.4e-6, bf16: 1.2e-2) -- the tensor core's fp32 accumulation, not the split, sets the floor -- so the cheaper one is used.
SPLIT_TERMS = 3
_split_ws: dict = {}     # per-device workspace of the fp32-accurate mode; outgrown buffers stay alive for captured graphs
_split_ws_retired: list = []


def _split_workspace(g: GemmDesc, device) -> torch.Tensor:
    need = int(_L().s2s_gemm_workspace_bytes(ctypes.byref(g)))
    ws = _split_ws.get(device)
    if ws is None or ws.numel() < need:
        if ws is not None:
            _split_ws_retired.append(ws)
        ws = torch.empty(max(need, 1 << 22), dtype=torch.uint8, device=device)
        _split_ws[device] = ws
    return ws


def gemm(a: torch.Tensor, b: torch.Tensor, c: torch.Tensor, *, mode: int = 0, **kw) -> torch.Tensor:
    """One s2s_gemm call: see _gemm_desc for the operand conventions and epilogue options."""
    g = _gemm_desc(a, b, c, **kw)
    if mode == 2:
        g.split_terms = SPLIT_TERMS
        ws = _split_workspace(g, c.device)
        g.ws, g.ws_bytes = ptr(ws), ws.numel()
    check(_L().s2s_gemm(ctypes.byref(g), mode, stream()), "s2s_gemm")
    return c


def gemm_grouped(problems, mode: int = 0) -> None:
    """Several independent GEMMs [(a, b, c, kwargs), ...] issued together (s2s_gemm_grouped): in mode 1 the weight-gradient
    products of one layer (bf16 operands contiguous along M / N, float32 c accumulated in place) share one persistent launch."""
    if not problems:
        return
    if mode == 2 or len(problems) == 1:
        for a, b, c, kw in problems:
            gemm(a, b, c, mode=mode, **kw)
        return
    arr = (GemmDesc * len(problems))()
    for i, (a, b, c, kw) in enumerate(problems):
        arr[i] = _gemm_desc(a, b, c, **kw)
    check(_L().s2s_gemm_grouped(arr, len(problems), mode, stream()), "s2s_gemm_grouped")


def colsum_multi(items) -> None:
    """[(x2d, out), ...]: out[c] += sum_r x2d[r, c] for every pair, in one launch (s2s_colsum_multi); one dtype per call."""
    if not items:
        return
    arr = (ColsumDesc * len(items))()
    for i, (x2d, out) in enumerate(items):
        assert x2d.dim() == 2 and x2d.stride(1) == 1 and x2d.dtype == items[0][0].dtype
        arr[i] = ColsumDesc(ptr(x2d), x2d.shape[0], x2d.shape[1], x2d.stride(0), ptr(out))
    check(_L().s2s_colsum_multi(arr, len(items), dt(items[0][0]), stream()), "s2s_colsum_multi")


def _gemm_desc(a: torch.Tensor, b: torch.Tensor, c: torch.Tensor, *, bias: Optional[torch.Tensor] = None,
               residual: Optional[torch.Tensor] = None, alpha: float = 1.0, relu: bool = False,
               accumulate: bool = False, drop: Drop = NO_DROP, taps: int = 1,
               row_mask: Optional[Tuple[int, int, int, int]] = None, M: Optional[int] = None,
               gate: Optional[torch.Tensor] = None, gate_scale: float = 1.0) -> GemmDesc:
    """c[..., m, n] = epilogue(alpha * sum_t sum_k a[..., m + t, k] * b[..., n, (t,) k]).

    a: (..., rows, K); b: (..., N, K) or (..., N, taps, K) when taps > 1; c: (..., M, N).
    Leading batch dims (0-2) come from c; a / b with fewer dims broadcast (stride 0).
    gate (layout of c): c = gate > 0 ? c * gate_scale : 0 -- relu' (and the dropout scale) of the layer whose output `gate` is.
    mode: 0 fp32 CUDA cores, 1 bf16 tcgen05, 2 float32 operands on tcgen05 through a bf16 split (SPLIT_TERMS partial products).
    """
    nb = c.dim() - 2
    assert 0 <= nb <= 2
    Mc, N = c.shape[-2], c.shape[-1]
    M = Mc if M is None else M
    K = a.shape[-1]
    assert c.stride(-1) == 1 or N == 1, "C must be contiguous along n"
    g = GemmDesc()
    g.M, g.N, g.K, g.taps = M, N, K, taps
    g.A, g.a_dtype = ptr(a), dt(a)
    g.a_rs, g.a_cs = a.stride(-2), a.stride(-1)
    if a.shape[-1] == 1:
        g.a_cs = 1
    g.a_bs1, g.a_bs2 = _lead_strides(a, 2)
    g.B, g.b_dtype = ptr(b), dt(b)
    if taps > 1:
        assert b.shape[-2] == taps and b.shape[-3] == N and b.shape[-1] == K
        g.b_rs, g.b_ts, g.b_cs = b.stride(-3), b.stride(-2), b.stride(-1)
        nbb_core = 3
    else:
        assert b.shape[-2] == N and b.shape[-1] == K, (b.shape, N, K)
        g.b_rs, g.b_ts, g.b_cs = b.stride(-2), 0, b.stride(-1)
        nbb_core = 2
    if K == 1:
        g.b_cs = 1
    g.b_bs1, g.b_bs2 = _lead_strides(b, nbb_core)
    g.C, g.c_dtype = ptr(c), dt(c)
    g.c_rs = c.stride(-2)
    g.batch1, g.batch2, g.c_bs1, g.c_bs2 = _batch_strides(c, nb)
    g.bias = ptr(bias)
    if gate is not None:
        assert residual is None, "one R operand: residual or gate"
        residual, g.r_mode, g.r_scale = gate, 1, float(gate_scale)
    if residual is not None:
        assert residual.dtype == c.dtype and residual.stride() == c.stride(), "residual / gate must share C's layout"
    g.R = ptr(residual)
    g.alpha, g.relu, g.accumulate = float(alpha), int(relu), int(accumulate)
    g.drop = drop.c()
    if row_mask is not None:
        g.mask_period, g.mask_offset, g.mask_lo, g.mask_hi = row_mask
    return g


def layernorm_fwd(x, gamma, beta, y, mean, rstd, eps=1e-12):
    d = x.shape[-1]
    rows = x.numel() // d
    assert x.is_contiguous() and y.is_contiguous()
    check(_L().s2s_layernorm_fwd(ptr(x), ptr(gamma), ptr(beta), ptr(y), ptr(mean), ptr(rstd), rows, d, eps, dt(x), stream()),
          "layernorm_fwd")
    return y


def layernorm_bwd(dy, x, gamma, mean, rstd, dx, dgamma, dbeta, dres=None, dx_drop=None, drop: Drop = NO_DROP):
    """dx_drop (optional second output) = dropout'(dx) with the mask of `drop`."""
    d = x.shape[-1]
    rows = x.numel() // d
    assert dy.is_contiguous() and x.is_contiguous() and (dres is None or dres.is_contiguous())
    if dx_drop is not None:
        assert dx_drop.is_contiguous() and dx_drop.dtype == dx.dtype and dx_drop.numel() == dx.numel()
        check(_L().s2s_layernorm_bwd_drop(ptr(dy), ptr(x), ptr(gamma), ptr(mean), ptr(rstd), ptr(dres), ptr(dx), ptr(dx_drop),
                                          ctypes.byref(drop.c()), ptr(dgamma), ptr(dbeta), rows, d, dt(x), stream()), "layernorm_bwd_drop")
        return dx
    check(_L().s2s_layernorm_bwd(ptr(dy), ptr(x), ptr(gamma), ptr(mean), ptr(rstd), ptr(dres), ptr(dx), ptr(dgamma), ptr(dbeta), rows, d,
                                 dt(x), stream()), "layernorm_bwd")
    return dx


def skinny_linear_fwd(x2d, w, bias, y2d):
    """y2d[r, j] = x2d[r] . w[j] + bias[j] for N = w.shape[0] <= 4."""
    rows, K = x2d.shape
    N = w.shape[0]
    assert x2d.is_contiguous() and w.is_contiguous() and y2d.is_contiguous() and x2d.dtype == w.dtype == y2d.dtype
    check(_L().s2s_skinny_linear_fwd(ptr(x2d), ptr(w), ptr(bias), ptr(y2d), rows, K, N, dt(x2d), stream()), "skinny_linear_fwd")
    return y2d


def skinny_linear_bwd(dy2d, x2d, w, dw, dbias, dx, dx_accumulate=False):
    rows, K = x2d.shape
    N = w.shape[0]
    assert dy2d.is_contiguous() and x2d.is_contiguous() and w.is_contiguous() and (dx is None or dx.is_contiguous())
    check(_L().s2s_skinny_linear_bwd(ptr(dy2d), ptr(x2d), ptr(w), ptr(dw), ptr(dbias), ptr(dx), int(dx_accumulate), rows, K, N,
                                     dt(x2d), stream()), "skinny_linear_bwd")


def colsum(x2d, out):
    """out[c] += sum_r x2d[r, c]"""
    assert x2d.dim() == 2 and x2d.stride(1) == 1
    check(_L().s2s_colsum(ptr(x2d), x2d.shape[0], x2d.shape[1], x2d.stride(0), ptr(out), dt(x2d), stream()), "colsum")


def relu_bwd(dy, y, dx, scale=1.0):
    assert dy.is_contiguous() and y.is_contiguous() and dx.is_contiguous()
    check(_L().s2s_relu_bwd(ptr(dy), ptr(y), ptr(dx), dy.numel(), scale, dt(dy), stream()), "relu_bwd")
    return dx


def dropout_bwd(dy, dx, drop: Drop):
    assert dy.is_contiguous() and dx.is_contiguous()
    cols = dy.shape[-1]
    check(_L().s2s_dropout_bwd(ptr(dy), ptr(dx), dy.numel() // cols, cols, ctypes.byref(drop.c()), dt(dy), stream()), "dropout_bwd")
    return dx


def add(a, b, out):
    assert a.is_contiguous() and b.is_contiguous() and out.is_contiguous() and a.dtype == b.dtype == out.dtype
    check(_L().s2s_add(ptr(a), ptr(b), ptr(out), a.numel(), dt(a), stream()), "add")
    return out


def softmax_fwd(S, klens, causal: bool, T2: int, Pd=None, drop: Drop = NO_DROP):
    """In place over S (B, H, T1, ld); columns >= T2 are written as zeros."""
    B, H, T1, ld = S.shape
    assert S.is_contiguous()
    check(_L().s2s_softmax_fwd(ptr(S), ptr(S), ptr(Pd), ptr(klens), B, H, T1, T2, ld, int(causal), ctypes.byref(drop.c()),
                               dt(S), stream()), "softmax_fwd")
    return S


def softmax_bwd(P, dP, T2: int, scale: float, drop: Drop = NO_DROP):
    B, H, T1, ld = P.shape
    assert P.is_contiguous() and dP.is_contiguous()
    check(_L().s2s_softmax_bwd(ptr(P), ptr(dP), B, H, T1, T2, ld, scale, ctypes.byref(drop.c()), dt(P), stream()), "softmax_bwd")
    return dP


FUSED_ATTN_DK = (16, 32, 48, 64, 96, 128)


def attn_probs_fwd(q, k, P, klens, causal: bool, T2: int, scale: float):
    """P (B,H,T1,ld) = masked softmax(scale * q k^T); q (B,T1,H,dk) / k (B,T2,H,dk) are strided bf16 views."""
    B, T1, H, dk = q.shape
    assert q.dtype == torch.bfloat16 and k.dtype == torch.bfloat16 and P.dtype == torch.bfloat16 and P.is_contiguous()
    assert q.stride(3) == 1 and k.stride(3) == 1 and k.shape[1] == T2 and P.shape[:3] == (B, H, T1)
    check(_L().s2s_attn_probs_fwd(ptr(q), q.stride(0), q.stride(1), q.stride(2), ptr(k), k.stride(0), k.stride(1), k.stride(2), ptr(P),
                                  ptr(klens), B, H, T1, T2, dk, P.shape[3], float(scale), int(causal), stream()), "attn_probs_fwd")
    return P


def attn_probs_bwd(dctx, v, P, d_att, dS, T2: int, scale: float):
    """dS (B,H,T1,ld) = scale * P * (dP - sum P dP), dP = dctx v^T (+ d_att); dctx (B,T1,H,dk), v (B,T2,H,dk) strided bf16 views."""
    B, T1, H, dk = dctx.shape
    assert dctx.dtype == torch.bfloat16 and v.dtype == torch.bfloat16 and P.is_contiguous() and dS.is_contiguous() and P.shape == dS.shape
    assert dctx.stride(3) == 1 and v.stride(3) == 1 and v.shape[1] == T2 and (d_att is None or (d_att.is_contiguous() and d_att.shape == P.shape))
    check(_L().s2s_attn_probs_bwd(ptr(dctx), dctx.stride(0), dctx.stride(1), dctx.stride(2), ptr(v), v.stride(0), v.stride(1), v.stride(2),
                                  ptr(P), ptr(d_att), ptr(dS), B, H, T1, T2, dk, P.shape[3], float(scale), stream()), "attn_probs_bwd")
    return dS


def _bthd(t):
    assert t.dim() == 4 and t.stride(3) == 1 and t.dtype == torch.bfloat16, "(B,T,H,d_k) bf16 view with contiguous d_k"
    return ptr(t), t.stride(0), t.stride(1), t.stride(2)


def attn_lse_shape(B: int, H: int, T1: int):
    """Shape of the row-statistics / D workspaces of attn_fwd_tc / attn_bwd_tc (row pitch = T1 rounded up to 64)."""
    return (B, H, (T1 + 63) // 64 * 64)


def attn_fwd_tc(q, k, v, ctx, lse, klens, causal: bool, scale: float, P=None):
    """Fused attention forward on tcgen05: ctx (B,T1,H,dk) = softmax(scale q k^T, masked) v; lse (B,H,T1p) float32 row
    statistics for the backward; P (B,H,T1,ld) bf16 receives the probabilities when given."""
    B, T1, H, dk = q.shape
    T2 = k.shape[1]
    assert k.shape == v.shape and k.stride() == v.stride() and ctx.shape == q.shape
    assert lse.dtype == torch.float32 and lse.is_contiguous() and tuple(lse.shape) == attn_lse_shape(B, H, T1)
    ld = 0
    if P is not None:
        assert P.dtype == torch.bfloat16 and P.is_contiguous() and tuple(P.shape[:3]) == (B, H, T1)
        ld = P.shape[3]
    qp, kp, vp, cp = _bthd(q), _bthd(k), _bthd(v), _bthd(ctx)
    check(_L().s2s_attn_fwd_tc(*qp, kp[0], vp[0], *kp[1:], *cp, ptr(lse), ptr(P), ld, ptr(klens), B, H, T1, T2, dk, float(scale),
                               int(causal), stream()), "attn_fwd_tc")
    return ctx


def attn_bwd_tc(q, k, v, ctx, dctx, lse, dvec, dq, dk_, dv, klens, causal: bool, scale: float):
    """Fused attention backward on tcgen05 (P recomputed from q, k, lse): fills dq (B,T1,H,dk), dk_ / dv (B,T2,H,dk)."""
    B, T1, H, dk = q.shape
    T2 = k.shape[1]
    assert k.stride() == v.stride() and ctx.stride() == dctx.stride() and dk_.stride() == dv.stride()
    assert lse.is_contiguous() and dvec.is_contiguous() and tuple(dvec.shape) == attn_lse_shape(B, H, T1) and dvec.dtype == torch.float32
    qp, kp, vp, cp, gp, dqp, dkp, dvp = _bthd(q), _bthd(k), _bthd(v), _bthd(ctx), _bthd(dctx), _bthd(dq), _bthd(dk_), _bthd(dv)
    check(_L().s2s_attn_bwd_tc(*qp, kp[0], vp[0], *kp[1:], cp[0], gp[0], *cp[1:], ptr(lse), ptr(dvec), *dqp, dkp[0], dvp[0], *dkp[1:],
                               ptr(klens), B, H, T1, T2, dk, float(scale), int(causal), stream()), "attn_bwd_tc")


def scaled_pe_fwd(x, pe, alpha, y, drop: Drop = NO_DROP):
    B, T, d = x.shape
    assert x.is_contiguous() and y.is_contiguous() and pe.shape[0] >= T and pe.shape[1] == d and pe.is_contiguous()
    check(_L().s2s_scaled_pe_fwd(ptr(x), ptr(pe), ptr(alpha), ptr(y), B, T, d, ctypes.byref(drop.c()), dt(x), stream()), "scaled_pe_fwd")
    return y


def scaled_pe_bwd(dy, pe, dx, dalpha, drop: Drop = NO_DROP):
    B, T, d = dy.shape
    assert dy.is_contiguous()
    check(_L().s2s_scaled_pe_bwd(ptr(dy), ptr(pe), ptr(dx), ptr(dalpha), B, T, d, ctypes.byref(drop.c()), dt(dy), stream()), "scaled_pe_bwd")
    return dx


def embed_pe_fwd(tokens, ilens, weight, pe, alpha, y, eos, padding_idx=0, drop: Drop = NO_DROP):
    """y (B, T_in + 1, d) = dropout(weight[tok] + alpha * pe) with the <eos> append folded in."""
    B, T_in = tokens.shape
    T_out, d = y.shape[1], y.shape[2]
    assert tokens.dtype == torch.int64 and tokens.is_contiguous() and ilens.dtype == _i32 and y.is_contiguous()
    check(_L().s2s_embed_pe_fwd(ptr(tokens), ptr(ilens), ptr(weight), ptr(pe), ptr(alpha), ptr(y), B, T_in, T_out, d, eos, padding_idx,
                                ctypes.byref(drop.c()), dt(y), stream()), "embed_pe_fwd")
    return y


def embed_pe_bwd(dy, tokens, ilens, pe, dweight, dalpha, eos, padding_idx=0, drop: Drop = NO_DROP):
    B, T_in = tokens.shape
    T_out, d = dy.shape[1], dy.shape[2]
    assert dy.is_contiguous()
    check(_L().s2s_embed_pe_bwd(ptr(dy), ptr(tokens), ptr(ilens), ptr(pe), ptr(dweight), ptr(dalpha), B, T_in, T_out, d, eos,
                                padding_idx, ctypes.byref(drop.c()), dt(dy), stream()), "embed_pe_bwd")


def conv1_fwd(x, w, bias, y1):
    B, T, F = x.shape
    C = w.shape[0]
    assert x.dtype == torch.float32 and x.is_contiguous() and w.is_contiguous() and y1.is_contiguous()
    check(_L().s2s_conv1_fwd(ptr(x), ptr(w), ptr(bias), ptr(y1), B, T, F, C, dt(y1), stream()), "conv1_fwd")
    return y1


def conv1_bwd(x, dy1, dw, dbias):
    B, T, F = x.shape
    C = dy1.shape[-1]
    check(_L().s2s_conv1_bwd(ptr(x), ptr(dy1), ptr(dw), ptr(dbias), B, T, F, C, dt(dy1), stream()), "conv1_bwd")


def im2col_s2(y1, col):
    B, T1, F1, C = y1.shape
    check(_L().s2s_im2col_s2(ptr(y1), ptr(col), B, T1, F1, C, dt(y1), stream()), "im2col_s2")
    return col


def col2im_s2(dcol, dy1):
    B, T1, F1, C = dy1.shape
    check(_L().s2s_col2im_s2(ptr(dcol), ptr(dy1), B, T1, F1, C, dt(dy1), stream()), "col2im_s2")
    return dy1


def col2im_s2_relu(dcol, y1, dy1):
    """dy1 = col2im(dcol) * (y1 > 0): the scatter-add with conv.0's ReLU' fused into its store."""
    B, T1, F1, C = dy1.shape
    assert y1.shape == dy1.shape and y1.dtype == dy1.dtype and y1.is_contiguous()
    check(_L().s2s_col2im_s2_relu(ptr(dcol), ptr(y1), ptr(dy1), B, T1, F1, C, dt(dy1), stream()), "col2im_s2_relu")
    return dy1


def shift_thin(ys, out, r):
    B, L, odim = ys.shape
    Lr = out.shape[1]
    assert ys.dtype == torch.float32 and ys.is_contiguous() and out.is_contiguous()
    check(_L().s2s_shift_thin(ptr(ys), ptr(out), B, L, Lr, odim, r, dt(out), stream()), "shift_thin")
    return out


def fix_targets(labels, olens, labels_out, olens_out, r):
    B, Lin = labels.shape
    Lout = labels_out.shape[1]
    assert labels.is_contiguous() and labels_out.is_contiguous() and olens.dtype == _i32
    check(_L().s2s_fix_targets(ptr(labels), ptr(olens), ptr(labels_out), ptr(olens_out), B, Lin, Lout, r, stream()), "fix_targets")


def bn_stats(x, sums, L, halo):
    B, Lp, C = x.shape
    check(_L().s2s_bn_stats(ptr(x), ptr(sums), B, L, halo, C, dt(x), stream()), "bn_stats")


def bn_finalize(sums, mean, invstd, running_mean, running_var, count, eps=1e-5, momentum=0.1):
    C = mean.numel()
    check(_L().s2s_bn_finalize(ptr(sums), ptr(mean), ptr(invstd), ptr(running_mean), ptr(running_var), count, C, eps, momentum,
                               stream()), "bn_finalize")


def bn_apply(x, mean, invstd, gamma, beta, y, L, halo, use_tanh, drop: Drop = NO_DROP):
    """use_tanh: activation code -- 0/False identity, 1/True tanh, 2 Swish."""
    B, Lp, C = x.shape
    check(_L().s2s_bn_apply(ptr(x), ptr(mean), ptr(invstd), ptr(gamma), ptr(beta), ptr(y), B, L, halo, C, int(use_tanh),
                            ctypes.byref(drop.c()), dt(x), stream()), "bn_apply")
    return y


def bn_bwd_reduce(dy, y, x, mean, invstd, gamma, beta, sums, L, halo, use_tanh, drop: Drop = NO_DROP):
    B, Lp, C = x.shape
    check(_L().s2s_bn_bwd_reduce(ptr(dy), ptr(y), ptr(x), ptr(mean), ptr(invstd), ptr(gamma), ptr(beta), ptr(sums), B, L, halo, C,
                                 int(use_tanh), ctypes.byref(drop.c()), dt(x), stream()), "bn_bwd_reduce")


def bn_bwd_apply(dy, y, x, mean, invstd, gamma, beta, sums, dx, dgamma, dbeta, L, halo, use_tanh, drop: Drop = NO_DROP):
    B, Lp, C = x.shape
    check(_L().s2s_bn_bwd_apply(ptr(dy), ptr(y), ptr(x), ptr(mean), ptr(invstd), ptr(gamma), ptr(beta), ptr(sums), ptr(dx),
                                ptr(dgamma), ptr(dbeta), B, L, halo, C, int(use_tanh), ctypes.byref(drop.c()), dt(x), stream()),
          "bn_bwd_apply")
    return dx


def bn_eval_stats(running_mean, running_var, mean, invstd, eps=1e-5):
    check(_L().s2s_bn_eval_stats(ptr(running_mean), ptr(running_var), ptr(mean), ptr(invstd), mean.numel(), eps, stream()),
          "bn_eval_stats")


def pack_conv1d_w(w, wp, wpt):
    OC, IC, K = w.shape
    assert w.dtype == torch.float32 and w.is_contiguous()
    od = dt(wp if wp is not None else wpt)
    check(_L().s2s_pack_conv1d_w(ptr(w), ptr(wp), ptr(wpt), OC, IC, K, od, stream()), "pack_conv1d_w")


def pad_rows(x, y, halo):
    B, L, C = x.shape
    check(_L().s2s_pad_rows(ptr(x), ptr(y), B, L, halo, C, dt(x), stream()), "pad_rows")
    return y


def unpad_rows(x, y, halo):
    B, L, C = y.shape
    check(_L().s2s_unpad_rows(ptr(x), ptr(y), B, L, halo, C, dt(x), stream()), "unpad_rows")
    return y


def seq2seq_loss(after, before, logits, ys, labels, olens, pos_weight, losses, d_after, d_before, d_logits, ws):
    B, L, odim = after.shape
    assert after.dtype == before.dtype == logits.dtype and ys.dtype == torch.float32 and labels.dtype == torch.float32
    assert all(t.is_contiguous() for t in (after, before, logits, ys, labels))
    check(_L().s2s_seq2seq_loss(ptr(after), ptr(before), ptr(logits), ptr(ys), ptr(labels), ptr(olens), B, L, ys.shape[1],
                                labels.shape[1], odim, pos_weight,
                                ptr(losses), ptr(d_after), ptr(d_before), ptr(d_logits), ptr(ws), dt(after), stream()),
          "seq2seq_loss")


def guided_attn_loss(att, ilens, olens, T_in, sigma, alpha, loss, d_att, ws):
    B, H, T_out, ld = att.shape
    assert att.is_contiguous()
    check(_L().s2s_guided_attn_loss(ptr(att), ptr(ilens), ptr(olens), B, H, T_out, T_in, ld, sigma, alpha, ptr(loss), ptr(d_att),
                                    ptr(ws), dt(att), stream()), "guided_attn_loss")


def sqnorm(g, out):
    check(_L().s2s_sqnorm(ptr(g), g.numel(), ptr(out), stream()), "sqnorm")


def adam_step(p, g, m, v, p16, lr_dev, beta1, beta2, eps, wd, step_dev, sqn, max_norm, grad_scale=1.0):
    check(_L().s2s_adam_step(ptr(p), ptr(g), ptr(m), ptr(v), ptr(p16), p.numel(), ptr(lr_dev), beta1, beta2, eps, wd,
                             ptr(step_dev), ptr(sqn), max_norm, grad_scale, stream()), "adam_step")


def step_advance(step_dev, seed_dev):
    check(_L().s2s_step_advance(ptr(step_dev), ptr(seed_dev), stream()), "step_advance")


def cast(src, dst):
    assert src.is_contiguous() and dst.is_contiguous() and src.numel() == dst.numel()
    check(_L().s2s_cast(ptr(src), ptr(dst), src.numel(), dt(src), dt(dst), stream()), "cast")
    return dst


def transpose_last2(src, dst, N, A, Bd, accumulate=False):
    """dst[n][b][a] (+)= src[n][a][b]"""
    assert src.is_contiguous() and dst.is_contiguous() and src.numel() == N * A * Bd == dst.numel()
    check(_L().s2s_transpose_last2(ptr(src), ptr(dst), N, A, Bd, dt(src), dt(dst), int(accumulate), stream()), "transpose_last2")
    return dst


def mas(log_p, text_lens, feats_lens, want_grad: bool = False):
    """Monotonic alignment search.  Returns (paths int32 (B,T_feats), ds f32 (B,T_text), bin_loss f32 (1,), d_log_p|None)."""
    B, TF, TT = log_p.shape
    assert log_p.dtype == torch.float32 and log_p.is_contiguous()
    dev = log_p.device
    paths = torch.empty(B, TF, dtype=_i32, device=dev)
    ds = torch.empty(B, TT, dtype=torch.float32, device=dev)
    bin_loss = torch.empty(1, dtype=torch.float32, device=dev)
    d_log_p = torch.zeros_like(log_p) if want_grad else None
    nbytes = int(_L().s2s_mas_workspace_bytes(B, TF, TT))
    ws = torch.empty(max(nbytes, 8), dtype=torch.uint8, device=dev)
    check(_L().s2s_mas(ptr(log_p), ptr(text_lens), ptr(feats_lens), B, TF, TT, ptr(paths), ptr(ds), ptr(bin_loss), ptr(d_log_p),
                       ptr(ws), nbytes, stream()), "mas")
    return paths, ds, bin_loss, d_log_p


def mas_workspace_bytes(B: int, TF: int, TT: int) -> int:
    return int(_L().s2s_mas_workspace_bytes(B, TF, TT))


def mas_into(log_p, text_lens, feats_lens, paths, ds, bin_loss, d_log_p, ws):
    """s2s_mas into caller-owned buffers (no allocation: CUDA-graph friendly).  d_log_p must be pre-zeroed."""
    B, TF, TT = log_p.shape
    assert log_p.dtype == torch.float32 and log_p.is_contiguous() and paths.dtype == _i32 and ds.dtype == torch.float32
    check(_L().s2s_mas(ptr(log_p), ptr(text_lens), ptr(feats_lens), B, TF, TT, ptr(paths), ptr(ds), ptr(bin_loss), ptr(d_log_p),
                       ptr(ws), ws.numel(), stream()), "mas")


def logmel(wav, window, basis, mel, n_fft, hop, eps, log_base, mean=None, scale=None):
    """STFT -> log-mel; with mean / scale (n_mels) the StandardScaler normalisation is fused into the store."""
    B, ns = wav.shape
    n_mels = basis.shape[0]
    assert wav.dtype == torch.float32 and wav.is_contiguous() and mel.is_contiguous()
    lb = 0.0 if log_base is None else float(log_base)
    if mean is not None:
        assert scale is not None and mean.dtype == torch.float32 and scale.dtype == torch.float32 and mean.numel() == n_mels == scale.numel()
        check(_L().s2s_logmel_norm(ptr(wav), ptr(window), ptr(basis), ptr(mean), ptr(scale), ptr(mel), B, ns, n_fft, hop, n_mels, eps, lb,
                                   stream()), "logmel_norm")
    else:
        check(_L().s2s_logmel(ptr(wav), ptr(window), ptr(basis), ptr(mel), B, ns, n_fft, hop, n_mels, eps, lb, stream()), "logmel")
    return mel


# ----------------------------------------------------------------------------------------------
# Conformer block
# ----------------------------------------------------------------------------------------------
def gl_istft(mag, angles, window, frames, y, n_fft, hop):
    """y = istft(mag * angles): mag (T, bins) float32, angles (T, bins, 2) float32 (re, im); frames (T, n_fft) workspace."""
    T = mag.shape[0]
    assert mag.dtype == angles.dtype == torch.float32 and mag.is_contiguous() and angles.is_contiguous() and y.numel() == hop * (T - 1)
    check(_L().s2s_gl_istft(ptr(mag), ptr(angles), ptr(window), ptr(frames), ptr(y), T, n_fft, hop, stream()), "gl_istft")
    return y


def gl_stft(y, window, spec, n_fft, hop, pad_reflect=False):
    T = spec.shape[0]
    assert y.dtype == spec.dtype == torch.float32 and y.is_contiguous() and spec.is_contiguous()
    check(_L().s2s_gl_stft(ptr(y), ptr(window), ptr(spec), T, y.numel(), n_fft, hop, int(bool(pad_reflect)), stream()), "gl_stft")
    return spec


def gl_update(rebuilt, tprev, angles, c):
    check(_L().s2s_gl_update(ptr(rebuilt), ptr(tprev), ptr(angles), rebuilt.numel() // 2, float(c), stream()), "gl_update")
    return angles


def bias_add2(q, u, v, qu, qv):
    """qu = q + u, qv = q + v for a (rows, d) view q with arbitrary row stride (slice of the fused QKV buffer)."""
    d = q.shape[-1]
    rows = q.numel() // d
    assert q.stride(-1) == 1 and qu.is_contiguous() and qv.is_contiguous()
    q2 = q.reshape(rows, d) if q.is_contiguous() else q
    ldq = q.stride(-2) if q.dim() >= 2 else d
    if q.dim() > 2:     # (B, T, d) slice of (B, T, 3, d): uniform row stride required
        assert all(q.stride(i) == q.stride(i + 1) * q.shape[i + 1] for i in range(q.dim() - 2)), "non-uniform row stride"
    check(_L().s2s_bias_add2(ptr(q2), ldq, ptr(u), ptr(v), ptr(qu), ptr(qv), rows, d, dt(q), stream()), "bias_add2")


def add_strided(a, b, out):
    """out (rows, d; uniform row stride) = a + b with a, b contiguous."""
    d = out.shape[-1]
    rows = out.numel() // d
    assert a.is_contiguous() and b.is_contiguous() and out.stride(-1) == 1
    if out.dim() > 2:
        assert all(out.stride(i) == out.stride(i + 1) * out.shape[i + 1] for i in range(out.dim() - 2)), "non-uniform row stride"
    check(_L().s2s_add_strided(ptr(a), ptr(b), ptr(out), out.stride(-2), rows, d, dt(out), stream()), "add_strided")
    return out


def relshift_add(S, BD, T):
    """S (B,H,T,ldS) += rel_shift(BD) with BD stored (H,B,T,ldB)."""
    B, H, T1, ldS = S.shape
    assert S.is_contiguous() and BD.is_contiguous() and BD.shape[:3] == (H, B, T1) and T1 == T
    check(_L().s2s_relshift_add(ptr(S), ptr(BD), B, H, T, ldS, BD.shape[3], dt(S), stream()), "relshift_add")
    return S


def relshift_bwd(dS, dBD, T):
    B, H, T1, ldS = dS.shape
    assert dS.is_contiguous() and dBD.is_contiguous() and dBD.shape[:3] == (H, B, T1) and T1 == T
    check(_L().s2s_relshift_bwd(ptr(dS), ptr(dBD), B, H, T, ldS, dBD.shape[3], dt(dS), stream()), "relshift_bwd")
    return dBD


def relshift_legacy_add(S, BD, T):
    """S (B,H,T,ldS) += legacy rel_shift(BD) (attention.py:138-157) with BD (H,B,T,ldB >= T) the T x T bd term."""
    B, H, T1, ldS = S.shape
    assert S.is_contiguous() and BD.is_contiguous() and BD.shape[:3] == (H, B, T1) and T1 == T
    check(_L().s2s_relshift_legacy_add(ptr(S), ptr(BD), B, H, T, ldS, BD.shape[3], dt(S), stream()), "relshift_legacy_add")
    return S


def relshift_legacy_bwd(dS, dBD, T):
    B, H, T1, ldS = dS.shape
    assert dS.is_contiguous() and dBD.is_contiguous() and dBD.shape[:3] == (H, B, T1) and T1 == T
    check(_L().s2s_relshift_legacy_bwd(ptr(dS), ptr(dBD), B, H, T, ldS, dBD.shape[3], dt(dS), stream()), "relshift_legacy_bwd")
    return dBD


def glu_fwd(x, y):
    C = y.shape[-1]
    assert x.is_contiguous() and y.is_contiguous() and x.shape[-1] == 2 * C
    check(_L().s2s_glu_fwd(ptr(x), ptr(y), y.numel() // C, C, dt(x), stream()), "glu_fwd")
    return y


def glu_bwd(dy, x, dx):
    C = dy.shape[-1]
    assert dy.is_contiguous() and x.is_contiguous() and dx.is_contiguous()
    check(_L().s2s_glu_bwd(ptr(dy), ptr(x), ptr(dx), dy.numel() // C, C, dt(x), stream()), "glu_bwd")
    return dx


def dwconv_fwd(x, w, bias, y):
    B, T, C = x.shape
    K = w.shape[-1]
    assert x.is_contiguous() and y.is_contiguous() and w.is_contiguous() and w.dtype == torch.float32 and w.numel() == C * K
    check(_L().s2s_dwconv_fwd(ptr(x), ptr(w), ptr(bias), ptr(y), B, T, C, K, dt(x), stream()), "dwconv_fwd")
    return y


def dwconv_bwd(dy, x, w, dx, dw, dbias=None):
    B, T, C = x.shape
    K = w.shape[-1]
    assert dy.is_contiguous() and x.is_contiguous() and (dx is None or dx.is_contiguous())
    check(_L().s2s_dwconv_bwd(ptr(dy), ptr(x), ptr(w), ptr(dx), ptr(dw), ptr(dbias), B, T, C, K, dt(x), stream()), "dwconv_bwd")


def swish_fwd(x, y, drop: Drop = NO_DROP):
    assert x.is_contiguous() and y.is_contiguous()
    check(_L().s2s_swish_fwd(ptr(x), ptr(y), x.numel(), ctypes.byref(drop.c()), dt(x), stream()), "swish_fwd")
    return y


def swish_bwd(dy, x, dx, drop: Drop = NO_DROP):
    assert dy.is_contiguous() and x.is_contiguous() and dx.is_contiguous()
    check(_L().s2s_swish_bwd(ptr(dy), ptr(x), ptr(dx), x.numel(), ctypes.byref(drop.c()), dt(x), stream()), "swish_bwd")
    return dx


def scale_dropout(x, y, scale, drop1: Drop = NO_DROP, drop2: Drop = NO_DROP):
    assert x.is_contiguous() and y.is_contiguous()
    check(_L().s2s_scale_dropout(ptr(x), ptr(y), x.numel(), float(scale), ctypes.byref(drop1.c()), ctypes.byref(drop2.c()),
                                 dt(x), stream()), "scale_dropout")
    return y


def axpy(x, y, alpha):
    assert x.is_contiguous() and y.is_contiguous() and x.dtype == y.dtype and x.numel() == y.numel()
    check(_L().s2s_axpy(ptr(x), ptr(y), x.numel(), float(alpha), dt(x), stream()), "axpy")
    return y


def rowscale(x, s, out):
    C = x.shape[-1]
    assert x.is_contiguous() and out.is_contiguous() and s.dtype == torch.float32 and s.numel() == x.numel() // C
    check(_L().s2s_rowscale(ptr(x), ptr(s), ptr(out), x.numel() // C, C, dt(x), stream()), "rowscale")
    return out


def gather_rows(x, start, count, y):
    B, Tin, C = x.shape
    Tout = y.shape[1]
    assert x.is_contiguous() and y.is_contiguous() and start.dtype == _i32 and count.dtype == _i32 and start.numel() == Tout
    check(_L().s2s_gather_rows(ptr(x), ptr(start), ptr(count), ptr(y), B, Tin, Tout, C, dt(x), stream()), "gather_rows")
    return y


def lr_cumsum(ds, cum, alpha=1.0, all_ones=False):
    """cum (B, T+1) int32 = exclusive prefix sums of round(ds * alpha); cum[:, T] = output lengths (length_regulator.py:81-96)."""
    B, T = ds.shape
    assert ds.dtype == torch.int64 and ds.is_contiguous() and cum.dtype == _i32 and cum.shape == (B, T + 1)
    check(_L().s2s_lr_cumsum(ptr(ds), ptr(cum), B, T, float(alpha), int(bool(all_ones)), stream()), "lr_cumsum")
    return cum


def lr_fwd(x, cum, y, pad_value=0.0):
    B, T, D = x.shape
    assert x.is_contiguous() and y.is_contiguous() and y.shape[0] == B and y.shape[2] == D and x.dtype == y.dtype
    check(_L().s2s_lr_fwd(ptr(x), ptr(cum), ptr(y), B, T, y.shape[1], D, float(pad_value), dt(x), stream()), "lr_fwd")
    return y


def lr_bwd(dy, cum, dx):
    B, T, D = dx.shape
    assert dy.is_contiguous() and dx.is_contiguous() and dy.dtype == dx.dtype
    check(_L().s2s_lr_bwd(ptr(dy), ptr(cum), ptr(dx), B, T, dy.shape[1], D, dt(dx), stream()), "lr_bwd")
    return dx


# ----------------------------------------------------------------------------------------------
# AAS-VC alignment block
# ----------------------------------------------------------------------------------------------
def align_logp_fwd(feats, text, text_lens, logp, lse):
    B, TF, C = feats.shape
    TT = text.shape[1]
    assert feats.is_contiguous() and text.is_contiguous() and logp.dtype == torch.float32 and logp.is_contiguous()
    check(_L().s2s_align_logp_fwd(ptr(feats), ptr(text), ptr(text_lens), ptr(logp), ptr(lse), B, TF, TT, C, dt(feats), stream()),
          "align_logp_fwd")
    return logp


def align_logp_bwd(dlogp, logp, lse, text_lens, W, rowsum, colsum):
    B, TF, TT = logp.shape
    assert dlogp.is_contiguous() and logp.is_contiguous() and W.is_contiguous() and dlogp.dtype == torch.float32
    check(_L().s2s_align_logp_bwd(ptr(dlogp), ptr(logp), ptr(lse), ptr(text_lens), ptr(W), ptr(rowsum), ptr(colsum), B, TF, TT,
                                  W.shape[-1], dt(W), stream()), "align_logp_bwd")


def forward_sum(logp, prior, text_lens, feats_lens, alpha_ws, loss, dlogp, grad_scale=1.0, blank_logp=-1.0):
    B, TF, TT = logp.shape
    assert logp.is_contiguous() and prior.is_contiguous() and prior.shape == logp.shape and alpha_ws.numel() >= logp.numel()
    check(_L().s2s_forward_sum(ptr(logp), ptr(prior), ptr(text_lens), ptr(feats_lens), B, TF, TT, float(blank_logp), ptr(alpha_ws),
                               ptr(loss), ptr(dlogp), float(grad_scale), stream()), "forward_sum")


def gauss_weights(ds, feats_lens, text_lens, P, delta=0.1):
    B, TF, ld = P.shape
    TT = ds.shape[1]
    assert ds.dtype == torch.float32 and ds.is_contiguous() and P.is_contiguous()
    check(_L().s2s_gauss_weights(ptr(ds), ptr(feats_lens), ptr(text_lens), ptr(P), B, TF, TT, ld, float(delta), dt(P), stream()),
          "gauss_weights")
    return P


def duration_loss(pre, ds, text_lens, d_outs, loss, d_pre, grad_scale=1.0, offset=1.0, clamp_max=10.0, g_douts=None):
    B, TT = ds.shape
    assert pre.is_contiguous() and ds.is_contiguous() and pre.numel() == B * TT
    assert g_douts is None or (g_douts.dtype == torch.float32 and g_douts.is_contiguous() and g_douts.numel() == B * TT)
    check(_L().s2s_duration_loss(ptr(pre), ptr(ds), ptr(text_lens), B, TT, float(offset), float(clamp_max), float(grad_scale),
                                 ptr(g_douts), ptr(d_outs), ptr(loss), ptr(d_pre), dt(pre), stream()), "duration_loss")


def duration_infer(pre, d, offset=1.0, clamp_max=10.0):
    assert pre.is_contiguous() and d.is_contiguous() and d.dtype == torch.float32 and pre.numel() == d.numel()
    check(_L().s2s_duration_infer(ptr(pre), ptr(d), d.numel(), float(offset), float(clamp_max), dt(pre), stream()), "duration_infer")
    return d


# ----------------------------------------------------------------------------------------------
# single-position decode (KV cache)
# ----------------------------------------------------------------------------------------------
def gemv(W, bias, x, y, residual=None, relu=False, drop: Drop = NO_DROP, pos_dev=None):
    """y (N,) = act(W (N,K) @ x (K,) + bias) * dropout + residual."""
    N, K = W.shape
    assert W.is_contiguous() and x.numel() == K and y.numel() == N and W.dtype == x.dtype == y.dtype
    check(_L().s2s_gemv(ptr(W), ptr(bias), ptr(x), ptr(residual), ptr(y), N, K, int(relu), ctypes.byref(drop.c()), ptr(pos_dev), dt(W),
                        stream()), "gemv")
    return y


def decode_attn(q, knew, vnew, kcache, vcache, H, dk, fixed_S, S_cap, pos_dev, scale, ctx, probs=None, ldp=0, probs_step_stride=0):
    """One decode step of multi-head attention over cached keys / values (see include/s2svc_b200.h: s2s_decode_attn).
    kcache / vcache: views whose element (s, h, j) sits at base + s * row_stride + h * dk + j with row_stride = kcache.stride(0)."""
    assert kcache.stride(-1) == 1 and vcache.stride() == kcache.stride() and kcache.dtype == q.dtype
    check(_L().s2s_decode_attn(ptr(q), ptr(knew), ptr(vnew), ptr(kcache), ptr(vcache), kcache.stride(0), H, dk, int(fixed_S), int(S_cap),
                               ptr(pos_dev), float(scale), ptr(ctx), ptr(probs), int(ldp), int(probs_step_stride), dt(q), stream()),
          "decode_attn")
    return ctx


def decode_pe(x, pe, alpha, pos_dev, y):
    check(_L().s2s_decode_pe(ptr(x), ptr(pe), ptr(alpha), ptr(pos_dev), ptr(y), x.numel(), dt(x), stream()), "decode_pe")
    return y


def decode_advance(feat, logit, next_in, frames, logits, pos_dev, odim, r):
    assert frames.dtype == torch.float32 and logits.dtype == torch.float32 and pos_dev.dtype == _i32
    check(_L().s2s_decode_advance(ptr(feat), ptr(logit), ptr(next_in), ptr(frames), ptr(logits), ptr(pos_dev), odim, r, dt(feat), stream()),
          "decode_advance")


def feat_stats(feats, lens, acc):
    """acc (2 D + 1) float64 += per-feature sum, sum of squares and the number of valid frames of feats (B, T, D) float32
    (bin/compute_statistics.py:128-132: what StandardScaler.partial_fit accumulates)."""
    B, T, D = feats.shape
    assert feats.dtype == torch.float32 and feats.is_contiguous() and acc.dtype == torch.float64 and acc.numel() == 2 * D + 1
    assert lens is None or (lens.dtype == _i32 and lens.numel() == B)
    check(_L().s2s_feat_stats(ptr(feats), ptr(lens) if lens is not None else None, ptr(acc), B, T, D, stream()), "feat_stats")
    return acc


def im2col2d(y, col, k, s):
    """col ((B T2 F2), k*k*C) patches of the channels-last map y (B, T1, F1, C) for a k x k convolution of stride s."""
    B, T1, F1, C = y.shape
    T2, F2 = (T1 - k) // s + 1, (F1 - k) // s + 1
    assert y.is_contiguous() and col.is_contiguous() and col.numel() == B * T2 * F2 * k * k * C and col.dtype == y.dtype
    check(_L().s2s_im2col2d(ptr(y), ptr(col), B, T1, F1, C, k, s, dt(y), stream()), "im2col2d")
    return col


def col2im2d(dcol, gate, dy, k, s):
    """dy (B, T1, F1, C) = adjoint of im2col2d applied to dcol, times ReLU'(gate) when gate (the forward map) is given."""
    B, T1, F1, C = dy.shape
    assert dcol.is_contiguous() and dy.is_contiguous() and dcol.dtype == dy.dtype and (gate is None or (gate.shape == dy.shape and gate.is_contiguous()))
    check(_L().s2s_col2im2d(ptr(dcol), ptr(gate) if gate is not None else None, ptr(dy), B, T1, F1, C, k, s, dt(dy), stream()), "col2im2d")
    return dy


def row_sqnorm(x2d, out):
    """out[r] = sum_c x2d[r, c]^2 (float32)."""
    rows, C = x2d.shape
    assert x2d.is_contiguous() and out.dtype == torch.float32 and out.numel() == rows
    check(_L().s2s_row_sqnorm(ptr(x2d), ptr(out), rows, C, dt(x2d), stream()), "row_sqnorm")
    return out


def align_logp_from_dot(logp, nf, nt, text_lens, lse):
    """logp (B, TF, TT) float32 holds feats . text on entry and log_softmax(-||feats - text||) on return (align_logp_fwd's result)."""
    B, TF, TT = logp.shape
    assert logp.is_contiguous() and logp.dtype == torch.float32 and nf.numel() == B * TF and nt.numel() == B * TT
    check(_L().s2s_align_logp_from_dot(ptr(logp), ptr(nf), ptr(nt), ptr(text_lens), ptr(lse), B, TF, TT, stream()), "align_logp_from_dot")
    return logp


def conv1_xcol(x, xcol):
    """xcol (B T1 F1, 16): the nine taps of Conv2d(1, C, 3, 2) around every output position of x (B, T, F) float32, a column of
    ones, six zero columns."""
    B, T, F = x.shape
    assert x.dtype == torch.float32 and x.is_contiguous() and xcol.is_contiguous() and xcol.shape == (B * ((T - 1) // 2) * ((F - 1) // 2), 16)
    check(_L().s2s_conv1_xcol(ptr(x), ptr(xcol), B, T, F, dt(xcol), stream()), "conv1_xcol")
    return xcol


def conv1_fwd_tc(x, w, bias, y1, xcol, w16, mode=1):
    """conv1_fwd through the tensor cores: y1 (B, T1, F1, C) = relu(patches(x) w16^T), w16 = [w | bias | 0]; leaves xcol for backward."""
    C = w.shape[0]
    conv1_xcol(x, xcol)
    assert w16.shape == (C, 16) and w16.dtype == xcol.dtype == y1.dtype and y1.is_contiguous()
    check(_L().s2s_conv1_pack_w(ptr(w), ptr(bias), ptr(w16), C, dt(w16), stream()), "conv1_pack_w")
    gemm(xcol, w16, y1.view(-1, C), relu=True, mode=mode)
    return y1


def conv1_bwd_tc(x, dy1, dw, dbias, xcol, g16, mode=1, xcol_ready=False):
    """conv1_bwd through the tensor cores: dw (C,1,3,3) += dy1^T patches(x), dbias += colsum(dy1).  xcol (B T1 F1, 16) in dy1's dtype
    (rebuilt here unless the forward left it: xcol_ready) and g16 (C, 16) float32 are caller-provided scratch."""
    C = dy1.shape[-1]
    P = dy1.numel() // C
    assert dy1.is_contiguous() and xcol.shape == (P, 16) and xcol.dtype == dy1.dtype
    assert g16.shape == (C, 16) and g16.dtype == torch.float32 and dw.is_contiguous() and dw.numel() == C * 9
    if not xcol_ready:
        conv1_xcol(x, xcol)
    g16.zero_()
    gemm(dy1.view(P, C).t(), xcol.t(), g16, accumulate=True, mode=mode)      # accumulate-in-place float32 output: split-K over the positions
    check(_L().s2s_conv1_dw_scatter(ptr(g16), ptr(dw), ptr(dbias), C, stream()), "conv1_dw_scatter")
