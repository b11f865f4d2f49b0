"""Per-kernel parity: every C-ABI entry point vs its float64 torch-CPU contract (tests/fake_ops.py)."""
import math

import numpy as np
import pytest
import torch

import fake_ops as F

pytestmark = pytest.mark.gpu

DT = [torch.float32, torch.bfloat16]


def tol(dt, scale=1.0):
    return (2e-5 if dt == torch.float32 else 2e-2) * scale


def dev(*ts):
    return [None if t is None else t.cuda() for t in ts]


def close(a, b, atol, what=""):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    err = (a - b).abs().max().item()
    assert err <= atol, f"{what}: max abs err {err} > {atol}"


def rnd(*shape, dt=torch.float32, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed + sum(shape))
    return (torch.randn(*shape, generator=g) * scale).to(dt)


@pytest.fixture(scope="module")
def ops():
    from seq2seq_vc_b200 import _lib, ops

    _lib.device_check()
    return ops


@pytest.mark.parametrize("mode,dt", [(0, torch.float32), (0, torch.bfloat16), (1, torch.bfloat16)])
@pytest.mark.parametrize("M,N,K", [(64, 48, 32), (257, 130, 72), (1000, 384, 384), (127, 8, 19 * 16)])
def test_gemm_plain_and_epilogue(ops, mode, dt, M, N, K):
    a, b = rnd(M, K, dt=dt, seed=1), rnd(N, K, dt=dt, seed=2, scale=1 / math.sqrt(K))
    bias, res = rnd(N, seed=3), rnd(M, N, dt=dt, seed=4)
    for kw in (dict(), dict(bias=bias, relu=True), dict(bias=bias, residual=res, alpha=0.5), dict(accumulate=True),
               dict(gate=res, gate_scale=1.25), dict(bias=bias, gate=res, gate_scale=0.5)) + (
                   (dict(bias=bias, gate=res, accumulate=True),) if mode != 1 else ()):
        c0 = rnd(M, N, dt=dt, seed=5)
        ref = F.gemm(a, b, c0.clone(), **kw)
        da, db, dc = dev(a, b, c0.clone())
        dkw = {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in kw.items()}
        ops.gemm(da, db, dc, mode=mode, **dkw)
        close(dc, ref, tol(dt, 2 * (8 if mode == 2 else 1)), f"gemm {kw.keys()}")


@pytest.mark.parametrize("mode,dt", [(0, torch.float32), (1, torch.bfloat16), (2, torch.float32)])
def test_gemm_transposed_operands_and_f32_accumulate(ops, mode, dt):
    M, N, K = 96, 200, 333          # dW = dy^T x : both operands contiguous along the non-reduced dim
    dy, x = rnd(K, M, dt=dt, seed=1), rnd(K, N, dt=dt, seed=2)
    g0 = rnd(M, N, seed=3)
    ref = F.gemm(dy.t(), x.t(), g0.clone(), accumulate=True)
    ddy, dx, dg = dev(dy, x, g0.clone())
    ops.gemm(ddy.t(), dx.t(), dg, accumulate=True, mode=mode)
    close(dg, ref, tol(dt, 20 * (8 if mode == 2 else 1)), "dW gemm")
    w = rnd(N, M, dt=dt, seed=4)    # dx = dy W : B operand n-major
    c = torch.zeros(K, M, dtype=dt)
    ref = F.gemm(x, w.t(), c.clone())
    dxx, dw, dc = dev(x, w, c)
    ops.gemm(dxx, dw.t(), dc, mode=mode)
    close(dc, ref, tol(dt, 20 * (8 if mode == 2 else 1)), "dx gemm")


@pytest.mark.parametrize("mode,dt", [(0, torch.float32), (1, torch.bfloat16), (2, torch.float32)])
@pytest.mark.parametrize("T1,T2,dk", [(37, 29, 16), (128, 127, 48), (64, 200, 64)])
def test_gemm_attention_views(ops, mode, dt, T1, T2, dk):
    B, H = 3, 4
    d = H * dk
    qkv = rnd(B, T1, 3, H, dk, dt=dt, seed=1)
    kv = rnd(B, T2, 2, H, dk, dt=dt, seed=2)
    ld = (T2 + 7) // 8 * 8
    q, k, v = qkv[:, :, 0], kv[:, :, 0], kv[:, :, 1]
    P = torch.zeros(B, H, T1, ld, dtype=dt)
    ref = F.gemm(q.permute(0, 2, 1, 3), k.permute(0, 2, 1, 3), P.clone()[..., :T2], alpha=1 / math.sqrt(dk))
    dq, dkv, dP = dev(qkv, kv, P)
    ops.gemm(dq[:, :, 0].permute(0, 2, 1, 3), dkv[:, :, 0].permute(0, 2, 1, 3), dP[..., :T2], alpha=1 / math.sqrt(dk), mode=mode)
    close(dP[..., :T2], ref, tol(dt, 4 * (8 if mode == 2 else 1)), "QK^T")
    Pm = torch.softmax(rnd(B, H, T1, ld, seed=3), -1).to(dt)
    ctx = torch.zeros(B, T1, d, dtype=dt)
    ref = F.gemm(Pm[..., :T2], v.permute(0, 2, 3, 1), ctx.clone().view(B, T1, H, dk).permute(0, 2, 1, 3))
    dPm, dctx = dev(Pm, ctx)
    ops.gemm(dPm[..., :T2], dkv[:, :, 1].permute(0, 2, 3, 1), dctx.view(B, T1, H, dk).permute(0, 2, 1, 3), mode=mode)
    close(dctx.view(B, T1, H, dk).permute(0, 2, 1, 3), ref, tol(dt, 2 * (8 if mode == 2 else 1)), "PV")
    # dv[b,s,h,:] = sum_t P[t,s] dctx[t,:]  (both operands contiguous along the output dims)
    g = rnd(B, T1, d, dt=dt, seed=4)
    dv = torch.zeros(B, T2, 2, H, dk, dtype=dt)
    g4 = g.view(B, T1, H, dk).permute(0, 2, 1, 3)
    ref = F.gemm(Pm[..., :T2].transpose(-1, -2), g4.transpose(-1, -2), dv.clone()[:, :, 1].permute(0, 2, 1, 3))
    dg, ddv = dev(g, dv)
    ops.gemm(dPm[..., :T2].transpose(-1, -2), dg.view(B, T1, H, dk).permute(0, 2, 1, 3).transpose(-1, -2),
             ddv[:, :, 1].permute(0, 2, 1, 3), mode=mode)
    close(ddv[:, :, 1].permute(0, 2, 1, 3), ref, tol(dt, 4 * (8 if mode == 2 else 1)), "dV")


@pytest.mark.parametrize("mode,dt", [(0, torch.float32), (1, torch.bfloat16), (2, torch.float32)])
def test_gemm_conv1d_taps(ops, mode, dt):
    B, L, ic, oc, k = 3, 50, 80, 64, 5
    halo, Lp = 2, 54
    x = rnd(B, Lp, ic, dt=dt, seed=1)
    x[:, :halo] = 0
    x[:, halo + L:] = 0
    w = rnd(oc, k, ic, dt=dt, seed=2, scale=0.1)
    z = torch.zeros(B, Lp, oc, dtype=dt)
    M = B * Lp - 2 * halo
    kw = dict(taps=k, row_mask=(Lp, halo, halo, halo + L), M=M)
    ref = z.clone()
    F.gemm(x.view(B * Lp, ic), w, ref.view(B * Lp, oc)[halo:], **kw)
    dx, dw, dz = dev(x, w, z)
    ops.gemm(dx.view(B * Lp, ic), dw, dz.view(B * Lp, oc)[halo:], mode=mode, **kw)
    close(dz, ref, tol(dt, 4 * (8 if mode == 2 else 1)), "taps gemm")
    # against torch conv1d directly
    y = torch.nn.functional.conv1d(x[:, halo:halo + L].float().transpose(1, 2), w.float().permute(0, 2, 1), padding=halo)
    close(dz[:, halo:halo + L], y.transpose(1, 2), tol(dt, 8 * (8 if mode == 2 else 1)), "conv1d")


def test_gemm_grouped_weight_gradients_and_multi_colsum(ops):
    """s2s_gemm_grouped: the dW = dy^T x products of one layer in ONE tcgen05 launch == the same products one by one (bit-level
    differences only from the red.add order), incl. ragged M / N / K and a problem the grouped kernel must hand back
    (K-major operand); s2s_colsum_multi == s2s_colsum per tensor."""
    bf = torch.bfloat16
    shapes = [(1152, 384, 16384), (384, 384, 16384), (768, 384, 4064), (1536, 384, 16384), (384, 1536, 16384), (80, 256, 3001), (160, 384, 777)]
    probs, refs = [], []
    for i, (M, N, K) in enumerate(shapes):
        dy, x = rnd(K, M, dt=bf, seed=10 + i).cuda(), rnd(K, N, dt=bf, seed=20 + i).cuda()
        g0 = rnd(M, N, seed=30 + i).cuda()
        ref = g0.clone()
        ops.gemm(dy.t(), x.t(), ref, accumulate=True, mode=1)
        probs.append((dy.t(), x.t(), g0, dict(accumulate=True)))
        refs.append(ref)
    n0 = ops._lib.launch_count()
    ops.gemm_grouped(probs, mode=1)
    assert ops._lib.launch_count() - n0 == 1, "seven weight gradients, one launch"
    for (M, N, K), (_, _, got, _), ref in zip(shapes, probs, refs):
        close(got, ref, 2e-3 * math.sqrt(K), f"grouped dW {M}x{N}x{K}")
        exact = (rnd(K, M, dt=bf, seed=0).float().t() @ rnd(K, N, dt=bf, seed=0).float()) if False else None
    # against the float64 contract as well
    M, N, K = shapes[0]
    dy, x = rnd(K, M, dt=bf, seed=10), rnd(K, N, dt=bf, seed=20)
    want = F.gemm(dy.t(), x.t(), rnd(M, N, seed=30), accumulate=True)
    close(probs[0][2], want, tol(bf, 40), "grouped dW vs contract")
    # a set with a member outside the grouped form: still every product is computed
    a, b = rnd(300, 200, dt=bf, seed=1).cuda(), rnd(96, 200, dt=bf, seed=2).cuda()
    c = torch.zeros(300, 96, dtype=bf, device="cuda")
    g1 = torch.zeros(384, 384, device="cuda")
    dy, x = rnd(5000, 384, dt=bf, seed=3).cuda(), rnd(5000, 384, dt=bf, seed=4).cuda()
    ops.gemm_grouped([(a, b, c, {}), (dy.t(), x.t(), g1, dict(accumulate=True))], mode=1)
    close(c, F.gemm(a.cpu(), b.cpu(), torch.zeros(300, 96, dtype=bf)), tol(bf, 4), "mixed set: plain member")
    close(g1, F.gemm(dy.cpu().t(), x.cpu().t(), torch.zeros(384, 384), accumulate=True), tol(bf, 40), "mixed set: dW member")
    # multi-colsum
    items, want = [], []
    for i, (rows, cols) in enumerate([(16384, 384), (4064, 768), (16384, 1536), (777, 80), (5, 8)]):
        xx = rnd(rows, cols, dt=bf, seed=40 + i).cuda()
        out = rnd(cols, seed=50 + i).cuda()
        w = out.clone()
        ops.colsum(xx, w)
        items.append((xx, out))
        want.append(w)
    n0 = ops._lib.launch_count()
    ops.colsum_multi(items)
    assert ops._lib.launch_count() - n0 == 1
    for (xx, out), w in zip(items, want):
        close(out, w, 2e-3 * math.sqrt(xx.shape[0]), "multi colsum")


def test_gemm_cta_pairs(ops):
    """The CTA-pair variant of the tcgen05 GEMM (cluster of 2, one 256 x BN cta_group::2 MMA per k-step, each CTA staging half
    of B), forced through the test hook, on every operand layout / epilogue the engines use; then the cost model's own
    choice on a shape where it picks pairs."""
    from seq2seq_vc_b200 import _lib
    from seq2seq_vc_b200._lib import Drop

    lib = _lib.load()
    bf = torch.bfloat16
    try:
        lib.s2s_debug_gemm_tile(2)
        for M, N, K in [(257, 130, 72), (1000, 384, 384), (640, 256, 200)]:
            test_gemm_plain_and_epilogue(ops, 1, bf, M, N, K)
        test_gemm_attention_views(ops, 1, bf, 300, 200, 64)
        test_gemm_conv1d_taps(ops, 1, bf)
        # weight gradient (both operands MN-major, fp32 accumulate in place, split-K) and data gradient (B MN-major)
        M, N, K = 300, 520, 1333
        dy, x = rnd(K, M, dt=bf, seed=1), rnd(K, N, dt=bf, seed=2)
        g0 = rnd(M, N, seed=3)
        ref = F.gemm(dy.t(), x.t(), g0.clone(), accumulate=True)
        ddy, dx, dg = dev(dy, x, g0.clone())
        ops.gemm(ddy.t(), dx.t(), dg, accumulate=True, mode=1)
        close(dg, ref, tol(bf, 40), "pair dW gemm")
        w = rnd(N, M, dt=bf, seed=4, scale=1 / math.sqrt(N))
        c = torch.zeros(K, M, dtype=bf)
        ref = F.gemm(x, w.t(), c.clone())
        dxx, dw, dc = dev(x, w, c)
        ops.gemm(dxx, dw.t(), dc, mode=1)
        close(dc, ref, tol(bf, 4), "pair dx gemm")
        # dropout + residual epilogue: same mask as the 128-row kernel (the mask is a function of the element index only)
        a, b = rnd(900, 256, dt=bf, seed=5), rnd(384, 256, dt=bf, seed=6, scale=1 / 16)
        res = rnd(900, 384, dt=bf, seed=7)
        da, db, dres = dev(a, b, res)
        outs = []
        for cg in (2, 1):
            lib.s2s_debug_gemm_tile(cg)
            o = torch.empty(900, 384, dtype=bf, device="cuda")
            ops.gemm(da, db, o, residual=dres, drop=Drop(0.3, seed=11, site=5), mode=1)
            outs.append(o)
        assert torch.equal(outs[0], outs[1])
    finally:
        lib.s2s_debug_gemm_tile(0)
    test_gemm_plain_and_epilogue(ops, 1, bf, 40000, 512, 256)


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("rows,d", [(7, 32), (1000, 384), (33, 50), (3001, 1536), (130, 256), (65, 1032), (40, 2048), (9, 2056), (5000, 80), (20011, 384),
                                     (4064, 512)])
def test_layernorm(ops, dt, rows, d):
    x, dy, res = rnd(rows, 1, d, dt=dt, seed=1), rnd(rows, 1, d, dt=dt, seed=2), rnd(rows, 1, d, dt=dt, seed=3)
    gam, bet = 1 + 0.1 * rnd(d, seed=4), 0.1 * rnd(d, seed=5)
    y, mean, rstd = torch.empty_like(x), torch.empty(rows), torch.empty(rows)
    F.layernorm_fwd(x, gam, bet, y, mean, rstd)
    dx, dg, db = torch.empty_like(x), torch.zeros(d), torch.zeros(d)
    F.layernorm_bwd(dy, x, gam, mean, rstd, dx, dg, db, dres=res)
    cx, cdy, cres, cg, cb = dev(x, dy, res, gam, bet)
    cy, cm, cr = torch.empty_like(cx), torch.empty(rows, device="cuda"), torch.empty(rows, device="cuda")
    ops.layernorm_fwd(cx, cg, cb, cy, cm, cr)
    close(cy, y, tol(dt, 4), "ln fwd")
    close(cm, mean, 1e-5, "mean")
    cdx, cdg, cdb = torch.empty_like(cx), torch.zeros(d, device="cuda"), torch.zeros(d, device="cuda")
    ops.layernorm_bwd(cdy, cx, cg, cm, cr, cdx, cdg, cdb, dres=cres)
    close(cdx, dx, tol(dt, 8), "ln dx")
    close(cdg, dg, tol(dt, 4) * math.sqrt(rows) + 1e-3, "ln dgamma")
    close(cdb, db, tol(dt, 4) * math.sqrt(rows) + 1e-3, "ln dbeta")


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("T1,T2,ld", [(19, 21, 24), (19, 21, 21), (70, 127, 128), (33, 768, 768), (12, 1000, 1000), (5, 1030, 1032)])
def test_softmax_fwd_bwd(ops, dt, causal, T1, T2, ld):
    B, H = 3, 2
    S = rnd(B, H, T1, ld, dt=dt, seed=1, scale=2.0)
    klens = torch.tensor([T2, max(1, T2 // 2 + 3), 0], dtype=torch.int32)
    P = F.softmax_fwd(S.clone(), klens, causal, T2)
    cS, ck = dev(S, klens)
    ops.softmax_fwd(cS, ck, causal, T2)
    close(cS, P, tol(dt), "softmax")
    assert (cS[2] == 0).all(), "rows without a visible key must be exactly zero"
    dP = rnd(B, H, T1, ld, dt=dt, seed=2)
    ref = F.softmax_bwd(P, dP.clone(), T2, 0.25)
    cdP = dP.cuda()
    ops.softmax_bwd(cS, cdP, T2, 0.25)
    close(cdP, ref, tol(dt, 2), "softmax bwd")


@pytest.mark.parametrize("dt", DT)
def test_scaled_pe_and_elementwise(ops, dt):
    B, T, d = 3, 17, 32
    x, pe, alpha = rnd(B, T, d, dt=dt, seed=1), rnd(40, d, seed=2), torch.tensor(0.7)
    y = F.scaled_pe_fwd(x, pe, alpha, torch.empty_like(x))
    cx, cpe, ca = dev(x, pe, alpha)
    cy = ops.scaled_pe_fwd(cx, cpe, ca, torch.empty_like(cx))
    close(cy, y, tol(dt), "pe fwd")
    dal, cdal = torch.zeros(()), torch.zeros((), device="cuda")
    F.scaled_pe_bwd(x, pe, torch.empty_like(x), dal)
    ops.scaled_pe_bwd(cx, cpe, torch.empty_like(cx), cdal)
    close(cdal, dal, tol(dt, 50), "dalpha")
    a, b = rnd(1001, dt=dt, seed=3), rnd(1001, dt=dt, seed=4)
    close(ops.add(a.cuda(), b.cuda(), torch.empty(1001, dtype=dt, device="cuda")), a + b, tol(dt), "add")
    close(ops.relu_bwd(a.cuda(), b.cuda(), torch.empty(1001, dtype=dt, device="cuda"), 2.0), F.relu_bwd(a, b, torch.empty_like(a), 2.0), tol(dt), "relu_bwd")
    out, cout = torch.zeros(48), torch.zeros(48, device="cuda")
    m = rnd(300, 48, dt=dt, seed=5)
    F.colsum(m, out)
    ops.colsum(m.cuda(), cout)
    close(cout, out, tol(dt, 20), "colsum")
    src = rnd(5, 7, 9, seed=6)
    dst = ops.transpose_last2(src.cuda(), torch.empty(5, 9, 7, dtype=dt, device="cuda"), 5, 7, 9)
    close(dst, src.transpose(1, 2), tol(dt), "transpose")
    w = rnd(6, 4, 5, seed=7)
    wp, wpt = torch.empty(6, 5, 4, dtype=dt, device="cuda"), torch.empty(4, 5, 6, dtype=dt, device="cuda")
    ops.pack_conv1d_w(w.cuda(), wp, wpt)
    rp, rpt = torch.empty(6, 5, 4), torch.empty(4, 5, 6)
    F.pack_conv1d_w(w, rp, rpt)
    close(wp, rp, tol(dt), "wp")
    close(wpt, rpt, tol(dt), "wpt")


@pytest.mark.parametrize("dt", DT)
def test_conv2d_subsampling_pieces(ops, dt):
    B, T, Fq, C = 2, 23, 80, 16
    x, w, bias = rnd(B, T, Fq, seed=1), rnd(C, 1, 3, 3, seed=2, scale=0.3), rnd(C, seed=3, scale=0.1)
    T1, F1 = (T - 1) // 2, (Fq - 1) // 2
    T2, F2 = (T1 - 1) // 2, (F1 - 1) // 2
    y1 = F.conv1_fwd(x, w, bias, torch.empty(B, T1, F1, C, dtype=dt))
    cy1 = ops.conv1_fwd(x.cuda(), w.cuda(), bias.cuda(), torch.empty(B, T1, F1, C, dtype=dt, device="cuda"))
    close(cy1, y1, tol(dt), "conv1")
    col = F.im2col_s2(y1, torch.empty(B * T2 * F2, 9 * C, dtype=dt))
    ccol = ops.im2col_s2(cy1, torch.empty(B * T2 * F2, 9 * C, dtype=dt, device="cuda"))
    close(ccol, col, tol(dt), "im2col")
    dcol = rnd(B * T2 * F2, 9 * C, dt=dt, seed=4)
    dy1 = F.col2im_s2(dcol, torch.empty(B, T1, F1, C, dtype=dt))
    cdy1 = ops.col2im_s2(dcol.cuda(), torch.empty(B, T1, F1, C, dtype=dt, device="cuda"))
    close(cdy1, dy1, tol(dt, 4), "col2im")
    y1g = torch.randn(B, T1, F1, C).to(dt)
    close(ops.col2im_s2_relu(dcol.cuda(), y1g.cuda(), torch.empty(B, T1, F1, C, dtype=dt, device="cuda")),
          F.col2im_s2_relu(dcol, y1g, torch.empty(B, T1, F1, C, dtype=dt)), tol(dt, 4), "col2im_relu")
    dw, db = torch.zeros(C, 1, 3, 3), torch.zeros(C)
    F.conv1_bwd(x, dy1, dw, db)
    cdw, cdb = torch.zeros(C, 1, 3, 3, device="cuda"), torch.zeros(C, device="cuda")
    ops.conv1_bwd(x.cuda(), cdy1, cdw, cdb)
    close(cdw, dw, tol(dt, 100), "conv1 dw")
    close(cdb, db, tol(dt, 100), "conv1 db")


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("use_tanh", [True, False])
def test_batchnorm_pipeline(ops, dt, use_tanh):
    B, L, halo, C = 3, 21, 2, 16
    Lp = L + 2 * halo
    x, dy = rnd(B, Lp, C, dt=dt, seed=1), rnd(B, Lp, C, dt=dt, seed=2)
    gam, bet = 1 + 0.1 * rnd(C, seed=3), 0.1 * rnd(C, seed=4)
    rm, rv = torch.zeros(C), torch.ones(C)

    def run(o, to):
        sums = to(torch.zeros(2 * C))
        X, DY, G, Bt, RM, RV = to(x), to(dy), to(gam), to(bet), to(rm.clone()), to(rv.clone())
        mean, invstd = to(torch.empty(C)), to(torch.empty(C))
        o.bn_stats(X, sums, L, halo)
        o.bn_finalize(sums, mean, invstd, RM, RV, B * L)
        y = o.bn_apply(X, mean, invstd, G, Bt, torch.empty_like(X), L, halo, use_tanh)
        bs = to(torch.zeros(2 * C))
        o.bn_bwd_reduce(DY, y, X, mean, invstd, G, Bt, bs, L, halo, use_tanh)
        dg, db = to(torch.zeros(C)), to(torch.zeros(C))
        dx = o.bn_bwd_apply(DY, y, X, mean, invstd, G, Bt, bs, torch.empty_like(X), dg, db, L, halo, use_tanh)
        return y, dx, dg, db, RM, RV

    ref = run(F, lambda t: t)
    got = run(ops, lambda t: t.cuda())
    for a, b, name, s in zip(got, ref, ("y", "dx", "dgamma", "dbeta", "running_mean", "running_var"), (2, 8, 40, 40, 2, 2)):
        close(a, b, tol(dt, s), name)


@pytest.mark.parametrize("dt", DT)
def test_losses(ops, dt):
    B, L, odim = 3, 22, 80
    after, before, logits = rnd(B, L, odim, dt=dt, seed=1), rnd(B, L, odim, dt=dt, seed=2), rnd(B, L, dt=dt, seed=3)
    ys, labels = rnd(B, L + 3, odim, seed=4), (rnd(B, L + 1, seed=5) > 0.5).float()
    olens = torch.tensor([22, 10, 1], dtype=torch.int32)
    lo, da, db, dl = torch.zeros(2), torch.empty_like(after), torch.empty_like(after), torch.empty_like(logits)
    F.seq2seq_loss(after, before, logits, ys, labels, olens, 10.0, lo, da, db, dl, None)
    clo = torch.zeros(2, device="cuda")
    cda, cdb, cdl = torch.empty_like(after.cuda()), torch.empty_like(after.cuda()), torch.empty_like(logits.cuda())
    ops.seq2seq_loss(after.cuda(), before.cuda(), logits.cuda(), ys.cuda(), labels.cuda(), olens.cuda(), 10.0, clo, cda, cdb, cdl,
                     torch.zeros(4, device="cuda"))
    close(clo, lo, 1e-4 if dt == torch.float32 else 1e-3, "losses")
    close(cda, da, 1e-6, "d_after")
    close(cdl, dl, 1e-4, "d_logits")
    z = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "kats.npz"))
    att = torch.from_numpy(z["gmh_att"]).to(dt)
    loss = torch.zeros(1, device="cuda")
    ops.guided_attn_loss(att.cuda(), torch.from_numpy(z["gmh_ilens"]).int().cuda(), torch.from_numpy(z["gmh_olens"]).int().cuda(),
                         att.shape[-1], 0.4, 1.0, loss, torch.empty_like(att.cuda()), torch.zeros(2, device="cuda"))
    assert abs(loss.item() - float(z["gmh_loss"])) <= (1e-5 if dt == torch.float32 else 2e-3)


def test_glue_and_optimizer(ops):
    ys = rnd(2, 9, 5, seed=1)
    out = ops.shift_thin(ys.cuda(), torch.empty(2, 4, 5, device="cuda"), 2)
    close(out, F.shift_thin(ys, torch.empty(2, 4, 5), 2), 0, "shift_thin")
    labels, olens = torch.zeros(2, 9), torch.tensor([9, 5], dtype=torch.int32)
    lo, oo = torch.empty(2, 8), torch.empty(2, dtype=torch.int32)
    F.fix_targets(labels, olens, lo, oo, 2)
    clo, coo = torch.empty(2, 8, device="cuda"), torch.empty(2, dtype=torch.int32, device="cuda")
    ops.fix_targets(labels.cuda(), olens.cuda(), clo, coo, 2)
    close(clo, lo, 0, "labels")
    assert coo.tolist() == oo.tolist()
    n = 100003
    p, g = rnd(n, seed=2), rnd(n, seed=3, scale=0.01)
    m, v = torch.zeros(n), torch.zeros(n)
    cp, cg, cm, cv = dev(p.clone(), g, m.clone(), v.clone())
    step, lr, sq = torch.tensor([1.0]), torch.tensor([1e-3]), torch.zeros(1)
    F.sqnorm(g, sq)
    F.adam_step(p, g, m, v, None, lr, 0.9, 0.999, 1e-8, 0.0, step, sq, 1.0)
    csq = torch.zeros(1, device="cuda")
    ops.sqnorm(cg, csq)
    close(csq, sq, 1e-4 * float(sq), "sqnorm")
    p16 = torch.empty(n, dtype=torch.bfloat16, device="cuda")
    ops.adam_step(cp, cg, cm, cv, p16, lr.cuda(), 0.9, 0.999, 1e-8, 0.0, step.cuda(), csq, 1.0)
    close(cp, p, 1e-6, "adam p")
    close(p16, p, 2e-2, "bf16 shadow")


@pytest.mark.parametrize("dt", DT)
def test_dropout_is_consistent_between_forward_and_backward(ops, dt):
    from seq2seq_vc_b200._lib import Drop

    M, N, K = 300, 64, 32
    a, b = rnd(M, K, dt=dt, seed=1), rnd(N, K, dt=dt, seed=2)
    seed_dev = torch.tensor([5], dtype=torch.int64, device="cuda")
    drop = Drop(0.3, seed=11, site=4, seed_dev=seed_dev)
    for mode in ((0, 1) if dt == torch.bfloat16 else (0,)):
        plain = ops.gemm(a.cuda(), b.cuda(), torch.empty(M, N, dtype=dt, device="cuda"), mode=mode).float()
        dropped = ops.gemm(a.cuda(), b.cuda(), torch.empty(M, N, dtype=dt, device="cuda"), drop=drop, mode=mode).float()
        mask = ops.dropout_bwd(torch.ones(M, N, dtype=dt, device="cuda"), torch.empty(M, N, dtype=dt, device="cuda"), drop).float()
        frac = (mask == 0).float().mean().item()
        assert abs(frac - 0.3) < 0.02, frac
        err = ((dropped - plain * mask).abs() / (1 + plain.abs())).max().item()
        assert err <= tol(dt, 2), f"dropout mask mismatch {err}"
    seed_dev += 1
    mask2 = ops.dropout_bwd(torch.ones(M, N, dtype=dt, device="cuda"), torch.empty(M, N, dtype=dt, device="cuda"), drop).float()
    assert (mask2 != mask).float().mean().item() > 0.2, "device-side seed must change the mask"


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("N", [1, 2, 4])
def test_skinny_linear(ops, dt, N):
    rows, K = 777, 96
    x, w, bias, dy = rnd(rows, K, dt=dt, seed=1), rnd(N, K, dt=dt, seed=2, scale=0.1), rnd(N, seed=3), rnd(rows, N, dt=dt, seed=4)
    y = F.skinny_linear_fwd(x, w, bias, torch.empty(rows, N, dtype=dt))
    cy = ops.skinny_linear_fwd(x.cuda(), w.cuda(), bias.cuda(), torch.empty(rows, N, dtype=dt, device="cuda"))
    close(cy, y, tol(dt, 4), "skinny fwd")
    dw, db, dx = torch.zeros(N, K), torch.zeros(N), rnd(rows, K, dt=dt, seed=5)
    cdw, cdb, cdx = dw.cuda(), db.cuda(), dx.clone().cuda()
    F.skinny_linear_bwd(dy, x, w, dw, db, dx, True)
    ops.skinny_linear_bwd(dy.cuda(), x.cuda(), w.cuda(), cdw, cdb, cdx, True)
    close(cdw, dw, tol(dt, 60), "skinny dw")
    close(cdb, db, tol(dt, 60), "skinny dbias")
    close(cdx, dx, tol(dt, 4), "skinny dx")


@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("B,H,T1,T2,dk", [(2, 3, 70, 70, 48), (3, 2, 130, 50, 64), (2, 4, 64, 127, 48), (2, 2, 200, 600, 96), (1, 8, 512, 512, 48),
                                           (3, 1, 17, 9, 16)])
def test_fused_attention_probs_fwd_bwd(ops, causal, B, H, T1, T2, dk):
    """scores + softmax (and dP + softmax') fused kernels vs GEMM + softmax contracts; q/k/v are strided slices of a fused buffer."""
    if causal and T1 != T2:
        pytest.skip("causal masks are only used for self-attention")
    dt = torch.bfloat16
    ld = (T2 + 7) // 8 * 8
    qkv = rnd(B, max(T1, T2), 3, H, dk, dt=dt, seed=1)
    q, k, v = qkv[:, :T1, 0], qkv[:, :T2, 1], qkv[:, :T2, 2]
    klens = torch.tensor([T2, max(1, T2 // 2 + 1), 0][:B], dtype=torch.int32)
    scale = 1.0 / math.sqrt(dk)
    P = F.attn_probs_fwd(q, k, torch.empty(B, H, T1, ld, dtype=dt), klens, causal, T2, scale)
    dqkv = qkv.cuda()
    gP = torch.full((B, H, T1, ld), 3.0, dtype=dt, device="cuda")
    ops.attn_probs_fwd(dqkv[:, :T1, 0], dqkv[:, :T2, 1], gP, klens.cuda(), causal, T2, scale)
    close(gP, P, 1.5e-2, "fused probs")
    assert (gP.float().sum(-1).cpu() - P.float().sum(-1)).abs().max().item() <= 3e-2
    if B == 3:
        assert (gP[2] == 0).all(), "rows without a visible key must be exactly zero"
    dctx = rnd(B, T1, H * dk, dt=dt, seed=2)
    for with_att in (False, True):
        d_att = rnd(B, H, T1, ld, dt=dt, seed=3, scale=0.5) if with_att else None
        ref = F.attn_probs_bwd(dctx.view(B, T1, H, dk), v, P, d_att, torch.empty(B, H, T1, ld, dtype=dt), T2, scale)
        gdS = torch.full((B, H, T1, ld), 5.0, dtype=dt, device="cuda")
        ops.attn_probs_bwd(dctx.cuda().view(B, T1, H, dk), dqkv[:, :T2, 2], P.cuda(), None if d_att is None else d_att.cuda(), gdS, T2, scale)
        tol_abs = 2e-2 * max(1.0, ref.float().abs().max().item())
        close(gdS, ref, tol_abs, f"fused dS (d_att={with_att})")


@pytest.mark.parametrize("dt", DT)
def test_layernorm_bwd_with_dropout_output(ops, dt):
    """s2s_layernorm_bwd_drop: dx equals the plain backward, dx_drop equals s2s_dropout_bwd(dx) with the same descriptor."""
    from seq2seq_vc_b200._lib import Drop

    rows, d = 777, 384
    x, dy, dres = rnd(rows, d, dt=dt, seed=1).cuda(), rnd(rows, d, dt=dt, seed=2).cuda(), rnd(rows, d, dt=dt, seed=3).cuda()
    gamma, beta = rnd(d, seed=4).cuda(), rnd(d, seed=5).cuda()
    y, mean, rstd = torch.empty_like(x), torch.empty(rows, device="cuda"), torch.empty(rows, device="cuda")
    ops.layernorm_fwd(x, gamma, beta, y, mean, rstd)
    seed_dev = torch.full((1,), 5, dtype=torch.int64, device="cuda")
    drop = Drop(0.25, seed=9, site=4, seed_dev=seed_dev)
    dx0, dg0, db0 = torch.empty_like(x), torch.zeros(d, device="cuda"), torch.zeros(d, device="cuda")
    ops.layernorm_bwd(dy, x, gamma, mean, rstd, dx0, dg0, db0, dres=dres)
    dx1, dxd, dg1, db1 = torch.empty_like(x), torch.empty_like(x), torch.zeros(d, device="cuda"), torch.zeros(d, device="cuda")
    ops.layernorm_bwd(dy, x, gamma, mean, rstd, dx1, dg1, db1, dres=dres, dx_drop=dxd, drop=drop)
    assert torch.equal(dx0, dx1)
    close(dg1, dg0, 1e-3 * max(1.0, dg0.abs().max().item()), "dgamma")
    want = ops.dropout_bwd(dx0, torch.empty_like(dx0), drop)
    assert torch.equal(dxd == 0, want == 0), "same mask"
    close(dxd, want, 1e-2 * max(1.0, want.float().abs().max().item()) if dt == torch.bfloat16 else 1e-5, "dx_drop")   # bf16: the fused
    # output scales the fp32 value before its single rounding, the two-kernel form rounds twice
    frac = (dxd == 0).float().mean().item()
    assert abs(frac - 0.25) < 0.02, frac


@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("emit", [False, True])
@pytest.mark.parametrize("B,H,T1,T2,dk", [(2, 3, 70, 70, 48), (3, 2, 130, 50, 64), (2, 4, 64, 127, 48), (2, 2, 200, 600, 96), (1, 8, 512, 512, 48),
                                           (3, 1, 17, 9, 16), (2, 2, 300, 300, 128), (2, 2, 129, 257, 32)])
def test_flash_attention_tc_fwd_bwd(ops, causal, emit, B, H, T1, T2, dk):
    """tcgen05 flash attention (S / P in tensor + shared memory) vs the float64 contract: context, row statistics, optional
    probability output, and dq / dk / dv of the recomputing backward.  q/k/v are strided slices of one fused buffer."""
    if causal and T1 != T2:
        pytest.skip("causal masks are only used for self-attention")
    dt = torch.bfloat16
    ld = (T2 + 7) // 8 * 8
    qkv = rnd(B, max(T1, T2), 3, H, dk, dt=dt, seed=1)
    q, k, v = qkv[:, :T1, 0], qkv[:, :T2, 1], qkv[:, :T2, 2]
    klens = torch.tensor([T2, max(1, T2 // 2 + 1), 0][:B], dtype=torch.int32)
    scale = 1.0 / math.sqrt(dk)
    shp = F.attn_lse_shape(B, H, T1)
    ctx, lse = torch.empty(B, T1, H, dk, dtype=dt), torch.empty(shp)
    P = torch.empty(B, H, T1, ld, dtype=dt) if emit else None
    F.attn_fwd_tc(q, k, v, ctx, lse, klens, causal, scale, P)
    dqkv = qkv.cuda()
    gq, gk, gv = dqkv[:, :T1, 0], dqkv[:, :T2, 1], dqkv[:, :T2, 2]
    gctx = torch.full((B, T1, H * dk), 7.0, dtype=dt, device="cuda").view(B, T1, H, dk)
    glse = torch.full(shp, 3.0, device="cuda")
    gP = torch.full((B, H, T1, ld), 3.0, dtype=dt, device="cuda") if emit else None
    ops.attn_fwd_tc(gq, gk, gv, gctx, glse, klens.cuda(), causal, scale, gP)
    torch.cuda.synchronize()
    close(gctx, ctx, 2e-2, "flash ctx")
    fin = torch.isfinite(lse[..., :T1])
    assert torch.equal(torch.isfinite(glse[..., :T1]).cpu(), fin), "rows without a visible key carry lse = +inf"
    assert (glse[..., :T1].cpu()[fin] - lse[..., :T1][fin]).abs().max().item() <= 2e-3
    if emit:
        close(gP, P, 1.5e-2, "flash probabilities")
        assert (gP.float().sum(-1).cpu() - P.float().sum(-1)).abs().max().item() <= 3e-2
    if B == 3:
        assert (gctx[2] == 0).all(), "an utterance without visible keys gives a zero context"
        if emit:
            assert (gP[2] == 0).all()
    # backward (recompute): contract fed with the DEVICE forward's ctx / lse, as the engine does
    dctx = rnd(B, T1, H, dk, dt=dt, seed=2)
    dq, dk_, dv = torch.empty(B, T1, H, dk, dtype=dt), torch.empty(B, T2, H, dk, dtype=dt), torch.empty(B, T2, H, dk, dtype=dt)
    F.attn_bwd_tc(q, k, v, gctx.cpu(), dctx, lse, torch.empty(shp), dq, dk_, dv, klens, causal, scale)
    gd = torch.full((B, max(T1, T2), 3, H, dk), 9.0, dtype=dt, device="cuda")
    gdvec = torch.empty(shp, device="cuda")
    ops.attn_bwd_tc(gq, gk, gv, gctx, dctx.cuda(), glse, gdvec, gd[:, :T1, 0], gd[:, :T2, 1], gd[:, :T2, 2], klens.cuda(), causal, scale)
    torch.cuda.synchronize()
    for name, got, ref in (("dq", gd[:, :T1, 0], dq), ("dk", gd[:, :T2, 1], dk_), ("dv", gd[:, :T2, 2], dv)):
        close(got, ref, 3e-2 * max(1.0, ref.float().abs().max().item()), "flash " + name)
    assert (gd[:, T1:, 0] == 9.0).all() and (gd[:, T2:, 1:] == 9.0).all(), "rows outside the views must stay untouched"


@pytest.mark.parametrize("dt", DT)
def test_decode_kernels(ops, dt):
    """Single-position decode kernels (KV cache) vs their contracts."""
    N, K, H, dk, cap, T2 = 136, 384, 8, 48, 40, 27
    d = H * dk
    W, b, x, res = rnd(N, K, dt=dt, seed=1, scale=0.05), rnd(N, seed=2), rnd(K, dt=dt, seed=3), rnd(N, dt=dt, seed=4)
    for kw in (dict(), dict(relu=True), dict(residual=res)):
        ref = F.gemv(W, b, x, torch.empty(N, dtype=dt), **kw)
        dkw = {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in kw.items()}
        got = ops.gemv(W.cuda(), b.cuda(), x.cuda(), torch.empty(N, dtype=dt, device="cuda"), **dkw)
        close(got, ref, tol(dt, 4), f"gemv {list(kw)}")
    Wo = rnd(5, 19, dt=dt, seed=5)                                     # odd K: scalar path
    close(ops.gemv(Wo.cuda(), None, x[:19].cuda().contiguous(), torch.empty(5, dtype=dt, device="cuda")),
          F.gemv(Wo, None, x[:19], torch.empty(5, dtype=dt)), tol(dt, 4), "gemv odd K")
    # self attention with cache append at pos = 11, then source attention over T2 keys with probability read-out
    cache = rnd(cap, 2, H, dk, dt=dt, seed=6)
    qkv = rnd(3 * d, dt=dt, seed=7)
    pos = torch.tensor([11], dtype=torch.int32)
    c_ref = cache.clone()
    ctx_ref = F.decode_attn(qkv[:d], qkv[d:2 * d], qkv[2 * d:], c_ref[:, 0], c_ref[:, 1], H, dk, -1, cap, pos, 0.2, torch.empty(d, dtype=dt))
    c_gpu, dq = cache.cuda(), qkv.cuda()
    ctx = ops.decode_attn(dq[:d], dq[d:2 * d], dq[2 * d:], c_gpu[:, 0], c_gpu[:, 1], H, dk, -1, cap, pos.cuda(), 0.2,
                          torch.empty(d, dtype=dt, device="cuda"))
    close(ctx, ctx_ref, tol(dt, 4), "decode self-attention")
    assert torch.equal(c_gpu.cpu(), c_ref), "cache row `pos` must hold the new key / value, other rows untouched"
    kv = rnd(T2, 2, H, dk, dt=dt, seed=8)
    nl, ldp = 3, 32
    att_ref, att = torch.zeros(cap, nl, H, ldp), torch.full((cap, nl, H, ldp), 9.0, device="cuda")
    F.decode_attn(qkv[:d], None, None, kv[:, 0], kv[:, 1], H, dk, T2, T2, pos, 0.2, torch.empty(d, dtype=dt), probs=att_ref[:, 1], ldp=ldp,
                  probs_step_stride=nl * H * ldp)
    dkv = kv.cuda()
    ops.decode_attn(dq[:d], None, None, dkv[:, 0], dkv[:, 1], H, dk, T2, T2, pos.cuda(), 0.2, torch.empty(d, dtype=dt, device="cuda"),
                    probs=att[:, 1], ldp=ldp, probs_step_stride=nl * H * ldp)
    close(att[11, 1], att_ref[11, 1], 1e-3 if dt == torch.bfloat16 else 1e-6, "source-attention probabilities of the step")
    assert (att[10] == 9.0).all() and (att[11, 0] == 9.0).all()
    # positional encoding + advance
    pe, alpha = rnd(cap, d, seed=9), torch.tensor([0.7])
    e = rnd(d, dt=dt, seed=10)
    close(ops.decode_pe(e.cuda(), pe.cuda(), alpha.cuda(), pos.cuda(), torch.empty(d, dtype=dt, device="cuda")),
          F.decode_pe(e, pe, alpha, pos, torch.empty(d, dtype=dt)), tol(dt), "decode pe")
    r, odim = 2, 80
    feat, logit = rnd(r * odim, dt=dt, seed=11), rnd(r, dt=dt, seed=12)
    nxt, frames, logits, p2 = (torch.zeros(odim, dtype=dt, device="cuda"), torch.zeros(cap, r * odim, device="cuda"),
                               torch.zeros(cap, r, device="cuda"), pos.cuda())
    ops.decode_advance(feat.cuda(), logit.cuda(), nxt, frames, logits, p2, odim, r)
    assert p2.item() == 12 and torch.equal(nxt.cpu(), feat[odim:]) and torch.equal(frames[11].cpu(), feat.float())
    assert torch.equal(logits[11].cpu(), logit.float()) and frames[12].abs().sum().item() == 0


@pytest.mark.parametrize("B,T,Fq,C", [(2, 23, 80, 16), (3, 64, 80, 384), (32, 512, 80, 384)])
def test_conv1_weight_gradient_through_gemm(ops, B, T, Fq, C):
    """conv.0's weight / bias gradient as dy1^T x (16-column patch matrix of the input) on the tcgen05 GEMM (bf16 engine) vs the
    CUDA-core kernel on the same bf16 dy1: the only difference is the bf16 rounding of the input patches (2^-9 relative per term)."""
    bf = torch.bfloat16
    T1, F1 = (T - 1) // 2, (Fq - 1) // 2
    g = torch.Generator().manual_seed(B + T)
    x = torch.randn(B, T, Fq, generator=g).cuda()
    dy1 = (torch.randn(B, T1, F1, C, generator=g) * 0.1).to(bf).cuda()
    dw0, db0 = torch.zeros(C, 1, 3, 3, device="cuda"), torch.zeros(C, device="cuda")
    ops.conv1_bwd(x, dy1, dw0, db0)
    dw1, db1 = torch.ones(C, 1, 3, 3, device="cuda"), torch.ones(C, device="cuda")          # accumulates into what is there
    xcol = torch.empty(B * T1 * F1, 16, dtype=bf, device="cuda")
    g16 = torch.empty(C, 16, device="cuda")
    ops.conv1_bwd_tc(x, dy1, dw1, db1, xcol, g16)
    assert (xcol[:, 9] == 1).all() and (xcol[:, 10:] == 0).all()
    ref_col = torch.nn.functional.unfold(x.unsqueeze(1), 3, stride=2).transpose(1, 2).reshape(-1, 9)      # (B T1 F1, 9), taps kt*3+kf
    assert torch.equal(xcol[:, :9].float(), ref_col.to(bf).float())
    scale = dw0.abs().max().item()
    assert ((dw1 - 1) - dw0).abs().max().item() <= 1e-2 * scale + 1e-3
    assert ((db1 - 1) - db0).abs().max().item() <= 2e-3 * db0.abs().max().item() + 1e-3
    cos = torch.nn.functional.cosine_similarity((dw1 - 1).flatten(), dw0.flatten(), dim=0).item()
    assert cos >= 0.9999, cos
    # forward on the same path: relu(patches x [w | bias | 0]^T) vs the direct kernel (fp32 input and weights, bf16 output)
    w, bias = (torch.randn(C, 1, 3, 3, generator=g) * 0.3).cuda(), (torch.randn(C, generator=g) * 0.1).cuda()
    y_ref = ops.conv1_fwd(x, w, bias, torch.empty(B, T1, F1, C, dtype=bf, device="cuda"))
    y_tc = ops.conv1_fwd_tc(x, w, bias, torch.empty(B, T1, F1, C, dtype=bf, device="cuda"), xcol, torch.empty(C, 16, dtype=bf, device="cuda"))
    err = (y_tc.float() - y_ref.float()).abs()
    assert err.max().item() <= 3e-2 * max(1.0, y_ref.float().abs().max().item()) and err.mean().item() <= 3e-3
    assert torch.equal(y_tc == 0, y_ref == 0) or ((y_tc == 0) != (y_ref == 0)).float().mean().item() < 2e-3      # ReLU cut at the same places
    dw2, db2 = torch.zeros(C, 1, 3, 3, device="cuda"), torch.zeros(C, device="cuda")
    ops.conv1_bwd_tc(x, dy1, dw2, db2, xcol, g16, xcol_ready=True)                                               # the forward left xcol
    assert (dw2 - (dw1 - 1)).abs().max().item() <= 1e-3 * scale + 1e-5
