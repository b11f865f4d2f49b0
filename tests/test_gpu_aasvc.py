"""AAS-VC hot path on the GPU through the C ABI: every new kernel vs its float64 contract (tests/fake_ops.py), the
forward-sum kernel vs the reference's F.ctc_loss golden vectors, and the whole engine vs golden vectors dumped from
the live reference and vs the CPU oracle.  Tolerances: mel L1 <= 1e-4, attention-weight L1 <= 1e-3 (fp32 path),
durations (integer alignment path) bit-exact."""
import math
import os

import numpy as np
import pytest
import torch

import fake_ops as F

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "aasvc_tiny.npz")
AAS_HP = dict(idim=80, odim=80, adim=32, aheads=2, elayers=1, eunits=48, dlayers=1, dunits=48,
              duration_predictor_input_dim=80, duration_predictor_layers=2, duration_predictor_chans=16,
              duration_predictor_kernel_size=3, postnet_layers=2, postnet_filts=5, postnet_chans=16,
              post_encoder_reduction_factor=4, conformer_enc_kernel_size=7, conformer_dec_kernel_size=7)
NO_DROPOUT = dict(transformer_enc_dropout_rate=0.0, transformer_enc_positional_dropout_rate=0.0, transformer_enc_attn_dropout_rate=0.0,
                  transformer_dec_dropout_rate=0.0, transformer_dec_positional_dropout_rate=0.0, transformer_dec_attn_dropout_rate=0.0,
                  duration_predictor_dropout_rate=0.0, postnet_dropout_rate=0.0)
DT = [torch.float32, torch.bfloat16]


def tol(dt, scale=1.0):
    return (2e-5 if dt == torch.float32 else 2e-2) * scale


def close(a, b, atol, what=""):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    assert a.shape == b.shape, (what, a.shape, b.shape)
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


# ---------------------------------------------------------------------------------------------- conformer kernels
@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("rows,d", [(37, 24), (300, 384)])
def test_bias_add2_and_add_strided(ops, dt, rows, d):
    qkv = rnd(rows, 3 * d, dt=dt, seed=1)
    u, v = rnd(d, seed=2), rnd(d, seed=3)
    qu, qv = torch.empty(rows, d, dtype=dt), torch.empty(rows, d, dtype=dt)
    F.bias_add2(qkv[:, :d], u, v, qu, qv)
    dq = qkv.cuda()
    gqu, gqv = torch.empty(rows, d, dtype=dt, device="cuda"), torch.empty(rows, d, dtype=dt, device="cuda")
    ops.bias_add2(dq[:, :d], u.cuda(), v.cuda(), gqu, gqv)
    close(gqu, qu, tol(dt), "qu")
    close(gqv, qv, tol(dt), "qv")
    out = torch.zeros(rows, 3 * d, dtype=dt, device="cuda")
    ops.add_strided(gqu, gqv, out[:, :d])
    close(out[:, :d], (qu.double() + qv.double()).to(dt), tol(dt, 2), "add_strided")
    assert out[:, d:].abs().max().item() == 0.0


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("B,H,T", [(2, 2, 7), (3, 2, 50), (2, 1, 129)])
def test_relshift_add_and_bwd(ops, dt, B, H, T):
    ld, ldb = (T + 7) // 8 * 8, (2 * T - 1 + 7) // 8 * 8
    S, BD = rnd(B, H, T, ld, dt=dt, seed=1), rnd(H, B, T, ldb, dt=dt, seed=2)
    ref = F.relshift_add(S.clone(), BD, T)
    g = ops.relshift_add(S.cuda(), BD.cuda(), T)
    close(g[..., :T], ref[..., :T], tol(dt, 2), "relshift_add")
    dS = rnd(B, H, T, ld, dt=dt, seed=3)
    dref = F.relshift_bwd(dS, torch.empty(H, B, T, ldb, dtype=dt), T)
    dg = ops.relshift_bwd(dS.cuda(), torch.full((H, B, T, ldb), 7.0, dtype=dt, device="cuda"), T)
    close(dg, dref, 0.0, "relshift_bwd")
    # adjointness: <shift(BD), dS> == <BD, shift^T(dS)>
    sh = F.relshift_add(torch.zeros(B, H, T, ld, dtype=torch.float64), BD.double(), T)
    lhs = (sh[..., :T] * dS.double()[..., :T]).sum()
    rhs = (BD.double() * dref.double()).sum()
    assert abs(float(lhs - rhs)) <= 1e-6 * max(1.0, abs(float(lhs)))


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,H,T", [(2, 2, 11), (1, 3, 64), (3, 2, 37)])
def test_relshift_legacy_add_and_bwd(ops, dt, B, H, T):
    """Legacy rel_shift (attention.py:138-157: pad, re-view, drop the first row -- rows wrap around) and its adjoint vs the
    reference's construction itself (tests/fake_ops._legacy_shift) and vs each other (<shift(x), y> == <x, shift^T(y)>)."""
    g = torch.Generator().manual_seed(T)
    ld, ldb = (T + 7) // 8 * 8, (T + 7) // 8 * 8
    S = torch.randn(B, H, T, ld, generator=g).to(dt)
    BD = torch.randn(H, B, T, ldb, generator=g).to(dt)
    ref = F.relshift_legacy_add(S.clone().float(), BD.float(), T)
    got = ops.relshift_legacy_add(S.cuda(), BD.cuda(), T)
    close(got[..., :T], ref[..., :T], tol(dt, 2), "relshift_legacy_add")
    dS = torch.randn(B, H, T, ld, generator=g).to(dt)
    dref = F.relshift_legacy_bwd(dS, torch.empty(H, B, T, ldb, dtype=dt), T)
    dg = ops.relshift_legacy_bwd(dS.cuda(), torch.full((H, B, T, ldb), 7.0, dtype=dt, device="cuda"), T)
    close(dg, dref, 0.0, "relshift_legacy_bwd")
    sh = F.relshift_legacy_add(torch.zeros(B, H, T, ld, dtype=torch.float64), BD.double(), T)
    lhs = (sh[..., :T] * dS[..., :T].double()).sum()
    rhs = (BD.double()[..., :T] * dg.cpu().double()[..., :T]).sum()
    assert abs(lhs - rhs) <= 1e-6 * max(1.0, abs(lhs))



@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("rows,C", [(50, 12), (333, 384)])
def test_glu_swish_scale(ops, dt, rows, C):
    x = rnd(rows, 2 * C, dt=dt, seed=1)
    y = F.glu_fwd(x, torch.empty(rows, C, dtype=dt))
    gy = ops.glu_fwd(x.cuda(), torch.empty(rows, C, dtype=dt, device="cuda"))
    close(gy, y, tol(dt), "glu_fwd")
    dy = rnd(rows, C, dt=dt, seed=2)
    dx = F.glu_bwd(dy, x, torch.empty(rows, 2 * C, dtype=dt))
    gdx = ops.glu_bwd(dy.cuda(), x.cuda(), torch.empty(rows, 2 * C, dtype=dt, device="cuda"))
    close(gdx, dx, tol(dt, 2), "glu_bwd")
    h = rnd(rows, C, dt=dt, seed=3, scale=2.0)
    close(ops.swish_fwd(h.cuda(), torch.empty_like(h, device="cuda")), F.swish_fwd(h, torch.empty_like(h)), tol(dt, 2), "swish_fwd")
    close(ops.swish_bwd(dy.cuda(), h.cuda(), torch.empty_like(h, device="cuda")), F.swish_bwd(dy, h, torch.empty_like(h)), tol(dt, 2),
          "swish_bwd")
    close(ops.scale_dropout(h.cuda(), torch.empty_like(h, device="cuda"), 3.5), F.scale_dropout(h, torch.empty_like(h), 3.5), tol(dt, 8),
          "scale")
    acc = rnd(rows, C, dt=dt, seed=4)
    close(ops.axpy(h.cuda(), acc.clone().cuda(), -0.5), F.axpy(h, acc.clone(), -0.5), tol(dt, 4), "axpy")
    s = rnd(rows, seed=5)
    close(ops.rowscale(h.cuda(), s.cuda(), torch.empty_like(h, device="cuda")), F.rowscale(h, s, torch.empty_like(h)), tol(dt, 8), "rowscale")


def test_swish_dropout_fwd_bwd_consistency(ops):
    """The backward regenerates the forward mask: d/dx sum(y * w) through the dropped swish."""
    from seq2seq_vc_b200._lib import Drop

    n = 4096
    x = rnd(n, seed=1, scale=1.5).cuda()
    drop = Drop(0.3, seed=5, site=9)
    y = ops.swish_fwd(x, torch.empty_like(x), drop)
    kept = (y != 0) | (x == 0)
    frac = 1.0 - kept.float().mean().item()
    assert 0.25 <= frac <= 0.35
    ref = x * torch.sigmoid(x) / 0.7
    close(y[kept], ref[kept], 1e-5, "kept values scaled by 1/(1-p)")
    dy = torch.ones_like(x)
    dx = ops.swish_bwd(dy, x, torch.empty_like(x), drop)
    assert (dx[~kept] == 0).all()
    sg = torch.sigmoid(x)
    close(dx[kept], ((sg + x * sg * (1 - sg)) / 0.7)[kept], 1e-5, "swish' under the same mask")
    y2 = ops.scale_dropout(x, torch.empty_like(x), 2.0, drop, Drop(0.2, seed=5, site=10))
    z = (y2 == 0).float().mean().item()
    assert 0.38 <= z <= 0.50                       # 1 - 0.7 * 0.8 = 0.44


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("B,T,C,K", [(2, 19, 12, 7), (3, 50, 32, 15), (2, 40, 384, 31), (2, 9, 16, 15), (2, 300, 72, 15), (1, 131, 64, 5),
                                      (3, 257, 136, 7)])
def test_dwconv_fwd_bwd(ops, dt, B, T, C, K):
    x, w, bias = rnd(B, T, C, dt=dt, seed=1), rnd(C, K, seed=2, scale=0.3), rnd(C, seed=3)
    y = F.dwconv_fwd(x, w, bias, torch.empty(B, T, C, dtype=dt))
    gy = ops.dwconv_fwd(x.cuda(), w.cuda(), bias.cuda(), torch.empty(B, T, C, dtype=dt, device="cuda"))
    close(gy, y, tol(dt, 4), "dwconv_fwd")
    dy = rnd(B, T, C, dt=dt, seed=4)
    dx, dw, db = torch.empty(B, T, C, dtype=dt), torch.zeros(C, K), torch.zeros(C)
    F.dwconv_bwd(dy, x, w, dx, dw, db)
    gdx, gdw, gdb = torch.empty(B, T, C, dtype=dt, device="cuda"), torch.zeros(C, K, device="cuda"), torch.zeros(C, device="cuda")
    ops.dwconv_bwd(dy.cuda(), x.cuda(), w.cuda(), gdx, gdw, gdb)
    close(gdx, dx, tol(dt, 4), "dwconv dx")
    close(gdw, dw, tol(torch.float32, 20) * math.sqrt(B * T), "dwconv dw")
    close(gdb, db, tol(torch.float32, 20) * math.sqrt(B * T), "dwconv dbias")


@pytest.mark.parametrize("dt", DT)
def test_bn_swish(ops, dt):
    B, L, C = 3, 21, 32
    x = rnd(B, L, C, dt=dt, seed=1)
    mean, invstd = rnd(C, seed=2, scale=0.1), rnd(C, seed=3).abs() + 0.5
    gam, bet = rnd(C, seed=4), rnd(C, seed=5)
    y = F.bn_apply(x, mean, invstd, gam, bet, torch.empty_like(x), L, 0, 2)
    gy = ops.bn_apply(x.cuda(), mean.cuda(), invstd.cuda(), gam.cuda(), bet.cuda(), torch.empty_like(x, device="cuda"), L, 0, 2)
    close(gy, y, tol(dt, 4), "bn swish fwd")
    dy = rnd(B, L, C, dt=dt, seed=6)
    sums = torch.zeros(2 * C)
    F.bn_bwd_reduce(dy, y, x, mean, invstd, gam, bet, sums, L, 0, 2)
    gs = torch.zeros(2 * C, device="cuda")
    ops.bn_bwd_reduce(dy.cuda(), gy, x.cuda(), mean.cuda(), invstd.cuda(), gam.cuda(), bet.cuda(), gs, L, 0, 2)
    close(gs, sums, tol(dt, 40), "bn swish sums")
    dx, dg, db = torch.empty_like(x), torch.zeros(C), torch.zeros(C)
    F.bn_bwd_apply(dy, y, x, mean, invstd, gam, bet, sums, dx, dg, db, L, 0, 2)
    gdx, gdg, gdb = torch.empty_like(x, device="cuda"), torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    ops.bn_bwd_apply(dy.cuda(), gy, x.cuda(), mean.cuda(), invstd.cuda(), gam.cuda(), bet.cuda(), sums.cuda(), gdx, gdg, gdb, L, 0, 2)
    close(gdx, dx, tol(dt, 8), "bn swish dx")


@pytest.mark.parametrize("dt", DT)
def test_gather_rows(ops, dt):
    from seq2seq_vc_b200.aasvc_engine import nearest_index

    B, Tin, Tout, C = 3, 11, 12, 24
    x = rnd(B, Tin, C, dt=dt, seed=1)
    idx = nearest_index(Tin, Tout)
    ref = x[:, idx]
    st, ct = torch.tensor(idx, dtype=torch.int32), torch.ones(Tout, dtype=torch.int32)
    got = ops.gather_rows(x.cuda(), st.cuda(), ct.cuda(), torch.empty(B, Tout, C, dtype=dt, device="cuda"))
    close(got, ref, 0.0, "nearest gather")
    dy = rnd(B, Tout, C, dt=dt, seed=2)
    first, cnt = [0] * Tin, [0] * Tin
    for j, i in enumerate(idx):
        if cnt[i] == 0:
            first[i] = j
        cnt[i] += 1
    dref = torch.zeros(B, Tin, C, dtype=torch.float64)
    dref.index_add_(1, torch.tensor(idx), dy.double())
    got = ops.gather_rows(dy.cuda(), torch.tensor(first, dtype=torch.int32).cuda(), torch.tensor(cnt, dtype=torch.int32).cuda(),
                          torch.empty(B, Tin, C, dtype=dt, device="cuda"))
    close(got, dref.to(dt), tol(dt, 4), "gather adjoint")


# ---------------------------------------------------------------------------------------------- alignment block
@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("B,TF,TT,C", [(2, 13, 5, 24), (3, 70, 33, 128), (2, 130, 65, 40)])
def test_align_logp_fwd_bwd(ops, dt, B, TF, TT, C):
    feats, text = rnd(B, TF, C, dt=dt, seed=1), rnd(B, TT, C, dt=dt, seed=2)
    tl = torch.tensor([TT, max(1, TT // 2), max(1, TT - 3)][:B], dtype=torch.int32)
    logp, lse = torch.empty(B, TF, TT), torch.empty(B, TF)
    F.align_logp_fwd(feats, text, tl, logp, lse)
    glogp, glse = torch.empty(B, TF, TT, device="cuda"), torch.empty(B, TF, device="cuda")
    ops.align_logp_fwd(feats.cuda(), text.cuda(), tl.cuda(), glogp, glse)
    fin = torch.isfinite(logp)
    assert torch.equal(torch.isfinite(glogp).cpu(), fin)
    close(glogp.cpu()[fin], logp[fin], 5e-5 * math.sqrt(C), "log_p_attn")     # inputs are identical (already rounded): fp32 math
    close(glse, lse, 5e-5 * math.sqrt(C), "lse")
    dlogp = rnd(B, TF, TT, seed=3)
    ld = (TT + 7) // 8 * 8
    W, rs, cs = torch.empty(B, TF, ld, dtype=dt), torch.empty(B, TF), torch.empty(B, TT)
    F.align_logp_bwd(dlogp, logp, lse, tl, W, rs, cs)
    gW = torch.full((B, TF, ld), 3.0, dtype=dt, device="cuda")
    grs, gcs = torch.empty(B, TF, device="cuda"), torch.empty(B, TT, device="cuda")
    ops.align_logp_bwd(dlogp.cuda(), logp.cuda(), lse.cuda(), tl.cuda(), gW, grs, gcs)
    scale = W.float().abs().max().item()
    close(gW, W, tol(dt, 2) * scale, "W")
    close(grs, rs, tol(dt, 4) * scale * math.sqrt(TT), "rowsum")
    close(gcs, cs, tol(dt, 4) * scale * math.sqrt(TF), "colsum")


@pytest.mark.parametrize("B,TF,TT,C", [(2, 13, 5, 24), (3, 70, 33, 128), (2, 130, 65, 40), (4, 768, 192, 384)])
def test_align_logp_through_tensor_core_product(ops, B, TF, TT, C):
    """The bf16 engine's form of the same block: |f|^2 + |x|^2 - 2 f.x with the product on the tcgen05 GEMM (bf16 operands, fp32
    accumulation) == the direct-difference kernel on the same bf16 inputs, up to the expansion's fp32 cancellation."""
    bf = torch.bfloat16
    feats, text = rnd(B, TF, C, dt=bf, seed=1).cuda(), rnd(B, TT, C, dt=bf, seed=2).cuda()
    tl = torch.tensor([TT, max(1, TT // 2), max(1, TT - 3), TT][:B], dtype=torch.int32).cuda()
    ref, rlse = torch.empty(B, TF, TT, device="cuda"), torch.empty(B, TF, device="cuda")
    ops.align_logp_fwd(feats, text, tl, ref, rlse)
    nf = ops.row_sqnorm(feats.view(B * TF, C), torch.empty(B * TF, device="cuda"))
    nt = ops.row_sqnorm(text.view(B * TT, C), torch.empty(B * TT, device="cuda"))
    close(nf, (feats.float() ** 2).sum(-1).view(-1), 1e-4 * C, "row_sqnorm")
    got, glse = torch.full((B, TF, TT), 9.0, device="cuda"), torch.empty(B, TF, device="cuda")
    ops.gemm(feats, text, got, mode=1)
    ops.align_logp_from_dot(got, nf, nt, tl, glse)
    fin = torch.isfinite(ref)
    assert torch.equal(torch.isfinite(got), fin)
    close(got[fin], ref[fin], 2e-4, "log_p_attn via f.x")
    close(glse, rlse, 2e-4, "lse")
    cpu, clse = torch.empty(B, TF, TT), torch.empty(B, TF)
    dot = torch.einsum("btc,bsc->bts", feats.float().cpu(), text.float().cpu())
    F.align_logp_from_dot(cpu.copy_(dot), nf.cpu(), nt.cpu(), tl.cpu(), clse)
    close(got.cpu()[fin.cpu()], cpu[fin.cpu()], 2e-4, "contract")


def test_forward_sum_matches_reference_ctc_golden(ops):
    """Loss and gradient of the reference's ForwardSumLoss (F.ctc_loss path, incl. an infeasible utterance)."""
    from seq2seq_vc_b200.aasvc_engine import beta_binomial_log_prior

    z = np.load(GOLDEN)
    lp = torch.from_numpy(z["fs_lp"])
    tl, fl = z["fs_tl"].tolist(), z["fs_fl"].tolist()
    B, TF, TT = lp.shape
    prior = torch.zeros(B, TF, TT)
    for b in range(B):
        prior[b, :fl[b], :tl[b]] = beta_binomial_log_prior(tl[b], fl[b])
    loss, g = torch.zeros(1, device="cuda"), torch.full((B, TF, TT), 5.0, device="cuda")
    ops.forward_sum(lp.cuda(), prior.cuda(), torch.tensor(tl, dtype=torch.int32).cuda(), torch.tensor(fl, dtype=torch.int32).cuda(),
                    torch.empty(B, TF, TT, device="cuda"), loss, g)
    assert abs(loss.item() - float(z["fs_loss"])) <= 2e-5 * max(1.0, abs(float(z["fs_loss"])))
    close(g, torch.from_numpy(z["fs_grad"]), 2e-5, "forward-sum gradient (torch ctc_loss backward semantics)")


@pytest.mark.parametrize("B,TF,TT", [(3, 50, 12), (4, 200, 48), (2, 768, 192)])
def test_forward_sum_vs_oracle(ops, B, TF, TT):
    lp = torch.log_softmax(rnd(B, TF, TT, seed=1), -1)
    tl = torch.tensor([TT, max(1, TT - 5), max(1, TT // 2), 1][:B], dtype=torch.int32)
    fl = torch.tensor([TF, max(1, TF - 9), max(1, TF // 2), TF][:B], dtype=torch.int32)
    prior = rnd(B, TF, TT, seed=2, scale=0.5) - 2.0
    loss, g = torch.zeros(1), torch.zeros(B, TF, TT)
    F.forward_sum(lp, prior, tl, fl, None, loss, g, grad_scale=2.0)
    gl, gg = torch.zeros(1, device="cuda"), torch.empty(B, TF, TT, device="cuda")
    ops.forward_sum(lp.cuda(), prior.cuda(), tl.cuda(), fl.cuda(), torch.empty(B, TF, TT, device="cuda"), gl, gg, 2.0)
    assert abs(gl.item() - loss.item()) <= 1e-4 * max(1.0, abs(loss.item()))
    close(gg, g, 1e-4, "forward-sum gradient")


@pytest.mark.parametrize("dt", DT)
def test_gauss_weights_and_duration_loss(ops, dt):
    B, TF, TT = 3, 44, 12
    ds = torch.tensor([[1.] * 11 + [33.], [30.] + [1.] * 10 + [0.], [22.] + [1.] * 8 + [0.] * 3])
    fl, tl = torch.tensor([44, 40, 30], dtype=torch.int32), torch.tensor([12, 11, 9], dtype=torch.int32)
    ld = 16
    P = F.gauss_weights(ds, fl, tl, torch.empty(B, TF, ld, dtype=dt))
    gP = ops.gauss_weights(ds.cuda(), fl.cuda(), tl.cuda(), torch.full((B, TF, ld), 2.0, dtype=dt, device="cuda"))
    close(gP, P, tol(dt), "gaussian upsampling weights")
    assert torch.equal(gP[1, 41].cpu(), gP[1, 0].cpu())          # padded output frames replicate frame 0 (reference quirk)
    pre = rnd(B * TT, 1, dt=dt, seed=1, scale=4.0)
    pre[5] = 11.0
    d_outs, loss, d_pre = torch.empty(B, TT), torch.zeros(1), torch.empty(B * TT, 1, dtype=dt)
    F.duration_loss(pre, ds, tl, d_outs, loss, d_pre)
    gd, gl, gp = torch.empty(B, TT, device="cuda"), torch.zeros(1, device="cuda"), torch.empty(B * TT, 1, dtype=dt, device="cuda")
    ops.duration_loss(pre.cuda(), ds.cuda(), tl.cuda(), gd, gl, gp)
    close(gd, d_outs, 1e-6, "d_outs")
    assert abs(gl.item() - loss.item()) <= 1e-5 * max(1.0, loss.item())
    close(gp, d_pre, tol(dt), "d_pre")
    assert gd[0, 5].item() == 10.0 and gp[5].item() == 0.0       # clamp(max=10) blocks the gradient


# ---------------------------------------------------------------------------------------------- whole engine
def _golden(fixture=None):
    z = np.load(GOLDEN if fixture is None else os.path.join(os.path.dirname(GOLDEN), fixture))
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd.")}
    return z, sd


def _step(eng, z):
    ilens, olens = z["ilens"].tolist(), z["olens"].tolist()
    xs = torch.from_numpy(z["xs"])[:, :max(ilens)].contiguous().cuda()
    ys = torch.from_numpy(z["ys"])[:, :max(olens)].contiguous().cuda()
    dpi = torch.from_numpy(z["dp_inputs"])[:, :max(ilens)].contiguous().cuda()
    after, before = eng.forward(xs, ys, dpi, ilens, olens)
    losses = eng.loss(ys)
    eng.backward()
    torch.cuda.synchronize()
    return after, before, losses


@pytest.mark.parametrize("fp32_gemm", ["simt", "tc"])
@pytest.mark.parametrize("pw,fixture", [("linear", "aasvc_tiny.npz"), ("conv1d", "aasvc_conv1d_tiny.npz"),
                                        ("conv1d:3", "aasvc_conv1d_k3_tiny.npz"), ("conv1d-linear:3", "aasvc_conv1d_linear_k3_tiny.npz")])
def test_golden_tiny_fp32_forward_losses_grads(pw, fixture, fp32_gemm):
    """Live-reference dumps of one training step: the shipped yaml's Linear + Swish position-wise layers and the AASVC class
    default (MultiLayeredConv1d k = 1 + ReLU, models/aas_vc.py:52-53); float32 engine on the CUDA-core GEMM ("simt") and on
    the fp32-accurate tcgen05 GEMM ("tc", the default float32 mode)."""
    from seq2seq_vc_b200.aasvc_engine import AASVCEngine

    z, sd = _golden(fixture)
    pw, _, pk = pw.partition(":")      # also MultiLayeredConv1d / Conv1dLinear with kernel size 3 (multi_layer_conv.py:12-108)
    eng = AASVCEngine(dict(AAS_HP, positionwise_layer_type=pw, positionwise_conv_kernel_size=int(pk or 1), **NO_DROPOUT), device="cuda:0",
                      bf16=False, fp32_gemm=fp32_gemm)
    eng.load_state_dict(sd)
    after, before, losses = _step(eng, z)
    assert np.abs(after.cpu().numpy() - z["after_outs"]).mean() <= 1e-4
    assert np.abs(before.cpu().numpy() - z["before_outs"]).mean() <= 1e-4
    lp, ref = eng.log_p_attn.cpu().numpy(), z["log_p_attn"]
    fin = np.isfinite(ref)
    assert np.array_equal(np.isfinite(lp), fin) and np.abs(lp[fin] - ref[fin]).max() <= 1e-4
    np.testing.assert_array_equal(eng.ds.cpu().numpy(), z["ds"])                  # integer alignment path: bit-exact
    assert np.abs(eng.d_outs.cpu().numpy() - z["d_outs"]).max() <= 1e-4
    for i, k in enumerate(("l1_loss", "forward_sum_loss", "bin_loss", "duration_loss")):
        assert abs(losses[i].item() - float(z[k])) <= 1e-4 * max(1.0, abs(float(z[k]))), k
    for k in [k for k in z.files if k.startswith("attn.")]:
        assert np.abs(eng.attn[k[5:]].cpu().numpy() - z[k]).mean() <= 1e-3, k
    gmax = max(np.abs(z[k]).max() for k in z.files if k.startswith("grad."))
    for name in eng.store.names():
        ref = z["grad." + name]
        got = eng.store.g(name).cpu().numpy()
        if fp32_gemm == "simt":
            assert np.abs(got - ref).max() <= 1e-3 * np.abs(ref).max() + 1e-5 * gmax, name
        else:
            # the tensor-core accumulator rounds differently from a sequential fp32 FMA chain: a ReLU input within ~1e-6 of zero
            # can land on the other side (seen: ONE element of alignment_module.f_conv1.weight off by 6 % of the tensor's
            # maximum); bound the mean error tightly and the worst element loosely
            assert np.abs(got - ref).mean() <= 2e-3 * np.abs(ref).mean() + 1e-6 * gmax, name
            assert np.abs(got - ref).max() <= 0.1 * np.abs(ref).max() + 1e-5 * gmax, name
    for k in z.files:
        if k.startswith("bn_after."):
            np.testing.assert_allclose(eng.buffers[k[9:]].cpu().numpy(), z[k], rtol=1e-4, atol=1e-6)
    eng.training = False
    ilens, olens = z["ilens"].tolist(), z["olens"].tolist()
    after_e, _ = eng.forward(torch.from_numpy(z["xs"]).cuda(), torch.from_numpy(z["ys"]).cuda(), torch.from_numpy(z["dp_inputs"]).cuda(),
                             ilens, olens)
    assert np.abs(after_e.cpu().numpy() - z["eval_after_outs"]).mean() <= 1e-4


def test_mid_size_fp32_vs_oracle_and_bf16_drift():
    """A mid-size ragged batch (d=64, 2+2 layers, k=15, T=120 -> L=96): fp32 CUDA path vs the CPU oracle on the same
    weights; then the bf16 tensor-core path with its drift printed and bounded."""
    from oracle import aasvc_oracle as ao
    from seq2seq_vc_b200.aasvc_engine import AASVCEngine

    hp = dict(idim=80, odim=80, adim=64, aheads=2, elayers=2, eunits=128, dlayers=2, dunits=128, duration_predictor_input_dim=80,
              duration_predictor_layers=2, duration_predictor_chans=32, duration_predictor_kernel_size=3, postnet_layers=3,
              postnet_filts=5, postnet_chans=32, post_encoder_reduction_factor=4, conformer_enc_kernel_size=15,
              conformer_dec_kernel_size=15)
    sd = ao.init_state_dict(hp, seed=4)
    ilens, olens = [120, 101, 77], [96, 90, 55]
    xs, ilens, ys, olens, dpi = ao.synthetic_batch(3, 120, 96, ilens=ilens, olens=olens, seed=8)
    out, parts, grads = ao.aasvc_loss_and_grads(sd, hp, xs, ilens, ys, olens, dpi)
    eng = AASVCEngine(dict(hp, **NO_DROPOUT), device="cuda:0", bf16=False)
    eng.load_state_dict(sd)
    after, before = eng.forward(xs.cuda(), ys.cuda(), dpi.cuda(), ilens, olens)
    losses = eng.loss(ys.cuda())
    eng.backward()
    torch.cuda.synchronize()
    assert (after.cpu() - out["after_outs"].detach()).abs().mean().item() <= 1e-4
    assert torch.equal(eng.ds.cpu(), out["ds"])
    for i, k in enumerate(eng.LOSS_NAMES):
        assert abs(losses[i].item() - float(parts[k])) <= 1e-4 * max(1.0, abs(float(parts[k]))), k
    gmax = max(float(g.abs().max()) for g in grads.values() if g is not None)
    for name in eng.store.names():
        ref = grads[name]
        got = eng.store.g(name).cpu()
        assert (got - ref).abs().max().item() <= 2e-3 * float(ref.abs().max()) + 2e-5 * gmax, name
    # bf16 tensor-core path: same step, drift reported (durations may legitimately differ when bf16 flips a MAS decision)
    eng16 = AASVCEngine(dict(hp, **NO_DROPOUT), device="cuda:0", bf16=True)
    eng16.load_state_dict(sd)
    a16, _ = eng16.forward(xs.cuda(), ys.cuda(), dpi.cuda(), ilens, olens)
    l16 = eng16.loss(ys.cuda())
    eng16.backward()
    torch.cuda.synchronize()
    drift = (a16.float().cpu() - out["after_outs"].detach()).abs().mean().item()
    same_ds = (eng16.ds.cpu() == out["ds"]).float().mean().item()
    print(f"bf16 drift: after L1 {drift:.3e}, durations equal {same_ds:.3f}, losses {l16.tolist()} vs {[float(parts[k]) for k in eng.LOSS_NAMES]}")
    assert drift <= 0.15 and torch.isfinite(l16).all()
    assert eng16.ds.sum(1).cpu().tolist() == [float(o) for o in olens]            # durations always sum to the target length
    g16 = eng16.store.G
    assert torch.isfinite(g16).all()
    cos = torch.nn.functional.cosine_similarity(g16.cpu(), eng.store.G.cpu(), dim=0).item()
    print(f"bf16 gradient cosine vs fp32 path: {cos:.4f}")
    assert cos >= 0.9


def test_c3_widths_bf16_drift_vs_fp32_oracle():
    """The widths bench.py times for AAS-VC (BASELINE configs[2]: encoder d384, x4 post-encoder reduction -> decoder d1536, 2 heads
    of d_k 768, k15, 4+4 layers) at a length the CPU oracle finishes in seconds (B 2 x 192 -> 176 frames): bf16 path vs the float32
    oracle on the same weights; drift printed, recorded and bounded; the float32 tensor-core mode held to 1e-4 / exact durations."""
    import json

    from oracle import aasvc_oracle as ao
    from seq2seq_vc_b200.aasvc_engine import AASVCEngine

    hp = dict(idim=80, odim=80, adim=384, aheads=2, elayers=4, eunits=1536, dlayers=4, dunits=1536, duration_predictor_input_dim=80,
              duration_predictor_layers=2, duration_predictor_chans=256, duration_predictor_kernel_size=3, postnet_layers=5,
              postnet_filts=5, postnet_chans=256, post_encoder_reduction_factor=4, conformer_enc_kernel_size=15,
              conformer_dec_kernel_size=15)
    sd = ao.init_state_dict(hp, seed=4)
    ilens, olens = [192, 164], [176, 151]
    xs, ilens, ys, olens, dpi = ao.synthetic_batch(2, 192, 176, ilens=ilens, olens=olens, seed=8)
    out, parts, grads = ao.aasvc_loss_and_grads(sd, hp, xs, ilens, ys, olens, dpi)
    rec = {}
    for mode, kw in (("fp32_tc", dict(bf16=False)), ("bf16", dict(bf16=True))):
        eng = AASVCEngine(dict(hp, **NO_DROPOUT), device="cuda:0", **kw)
        eng.load_state_dict(sd)
        after, before = eng.forward(xs.cuda(), ys.cuda(), dpi.cuda(), ilens, olens)
        losses = eng.loss(ys.cuda())
        eng.backward()
        torch.cuda.synchronize()
        cos = {}
        gmax = max(float(g.abs().max()) for g in grads.values() if g is not None)
        for name in eng.store.names():
            ref = grads[name]
            # gradients that are zero in exact arithmetic (a bias in front of BatchNorm) are rounding noise on both sides: skipped
            if ref is not None and ref.numel() >= 256 and float(ref.abs().max()) >= 1e-5 * gmax:
                cos[name] = torch.nn.functional.cosine_similarity(eng.store.g(name).cpu().flatten(), ref.flatten(), dim=0).item()
        worst = min(cos, key=cos.get)
        rec[mode] = dict(mel_L1=(after.float().cpu() - out["after_outs"].detach()).abs().mean().item(),
                         durations_equal=(eng.ds.cpu() == out["ds"]).float().mean().item(),
                         loss_rel_err={k: abs(losses[i].item() - float(parts[k])) / max(1.0, abs(float(parts[k]))) for i, k in enumerate(eng.LOSS_NAMES)},
                         grad_cosine_min=cos[worst], grad_cosine_min_tensor=worst, grad_cosine_median=float(np.median(list(cos.values()))))
        print(mode, rec[mode])
        assert eng.ds.sum(1).cpu().tolist() == [float(o) for o in olens]
        del eng
        torch.cuda.empty_cache()
    path = os.path.join(os.path.dirname(os.path.dirname(__file__)), "gpurun_out", "r02_bf16_drift.json")
    try:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        data = json.load(open(path)) if os.path.exists(path) else {}
        data["c3_widths_B2_192to176_aasvc"] = rec
        json.dump(data, open(path, "w"), indent=1)
    except OSError:
        pass
    assert rec["fp32_tc"]["mel_L1"] <= 1e-4 and rec["fp32_tc"]["durations_equal"] == 1.0 and rec["fp32_tc"]["grad_cosine_min"] >= 0.999
    assert max(rec["fp32_tc"]["loss_rel_err"].values()) <= 1e-4
    assert rec["bf16"]["mel_L1"] <= 0.15 and rec["bf16"]["grad_cosine_median"] >= 0.95


def test_dropout_training_step_runs_and_is_reproducible():
    """Default dropout rates: two forwards with the same seed state are identical, and backward is finite."""
    from seq2seq_vc_b200.aasvc_engine import AASVCEngine

    z, sd = _golden()
    eng = AASVCEngine(dict(AAS_HP), device="cuda:0", bf16=False, seed=3)
    eng.load_state_dict(sd)
    a1, _, l1 = _step(eng, z)
    a1 = a1.clone()
    g1 = eng.store.G.clone()
    a2, _, l2 = _step(eng, z)
    # same masks; BatchNorm statistics and gradients are accumulated with float atomics, so two runs agree up to
    # summation order, while a different seed changes the output by orders of magnitude more
    assert (a1 - a2).abs().max().item() <= 1e-4
    assert (g1 - eng.store.G).abs().max().item() <= 1e-3 * g1.abs().max().item()
    assert torch.isfinite(g1).all() and torch.isfinite(l1).all()
    eng.seed_dev += 1
    a3, _, _ = _step(eng, z)
    assert (a1 - a3).abs().mean().item() >= 1e-2


def test_conv1d_positionwise_dropout_and_bf16():
    """MultiLayeredConv1d (k = 1) + ReLU position-wise layers with dropout on: same-seed steps agree, the bf16 tensor-core path
    stays close to the fp32 one (no dropout), everything finite."""
    from seq2seq_vc_b200.aasvc_engine import AASVCEngine

    z, sd = _golden("aasvc_conv1d_tiny.npz")
    eng = AASVCEngine(dict(AAS_HP, positionwise_layer_type="conv1d"), device="cuda:0", bf16=False, seed=3)
    eng.load_state_dict(sd)
    a1, _, l1 = _step(eng, z)
    a1, g1 = a1.clone(), eng.store.G.clone()
    a2, _, _ = _step(eng, z)
    assert (a1 - a2).abs().max().item() <= 1e-4
    assert (g1 - eng.store.G).abs().max().item() <= 1e-3 * g1.abs().max().item()
    assert torch.isfinite(g1).all() and torch.isfinite(l1).all()
    assert eng.store.g("encoder.encoders.0.feed_forward.w_1.weight").abs().max().item() > 0
    engs = [AASVCEngine(dict(AAS_HP, positionwise_layer_type="conv1d", **NO_DROPOUT), device="cuda:0", bf16=b) for b in (False, True)]
    outs = []
    for e in engs:
        e.load_state_dict(sd)
        a, _, l = _step(e, z)
        torch.cuda.synchronize()
        outs.append((a.float().cpu(), e.store.G.clone().cpu()))
        assert torch.isfinite(l).all()
    assert (outs[0][0] - outs[1][0]).abs().mean().item() <= 0.15
    assert torch.nn.functional.cosine_similarity(outs[0][1], outs[1][1], dim=0).item() >= 0.9


def test_dropin_module_losses_and_autograd():
    """seq2seq_vc_b200.AASVC + L1Loss + ForwardSumLoss + DurationPredictorLoss used exactly as AASVCTrainer._train_step
    uses the reference classes (trainers/aas_vc.py:56-134): same kwargs, dict keys, state-dict names; gradients through
    torch autograd match the live-reference dump."""
    from seq2seq_vc_b200 import AASVC, DurationPredictorLoss, ForwardSumLoss, L1Loss

    z, sd = _golden()
    model = AASVC(**AAS_HP, **NO_DROPOUT, positionwise_layer_type="linear", positionwise_conv_kernel_size=1,
                  duration_predictor_use_encoder_outputs=False, encoder_normalize_before=True, decoder_normalize_before=True,
                  duration_predictor_type="deterministic", encoder_input_layer="linear", init_type="xavier_uniform", use_masking=True,
                  encoder_input_conv_kernel_size=3, prodiff_denoiser_layers=20).to("cuda:0")
    assert set(model.state_dict().keys()) == set(sd.keys())
    model.load_state_dict(sd)
    model.train()
    ilens, olens = torch.from_numpy(z["ilens"]), torch.from_numpy(z["olens"])
    xs, ys, dpi = (torch.from_numpy(z[k]).cuda() for k in ("xs", "ys", "dp_inputs"))
    ret = model(xs, ilens, ys, olens, dpi, dp_lengths=ilens)
    assert set(ret) >= {"before_outs", "after_outs", "ds", "ilens", "olens", "olens_reduced", "ys", "bin_loss", "log_p_attn", "d_outs"}
    assert np.abs(ret["after_outs"].detach().cpu().numpy() - z["after_outs"]).mean() <= 1e-4
    np.testing.assert_array_equal(ret["ds"].cpu().numpy(), z["ds"])
    assert ret["ilens"].tolist() == z["ilens_out"].tolist() and ret["olens"].tolist() == z["olens_out"].tolist()
    l1 = L1Loss()(ret["after_outs"], ret["before_outs"], ret["ys"], ret["olens"])
    fs = ForwardSumLoss()(ret["log_p_attn"], ret["ilens"], ret["olens_reduced"])
    dur = DurationPredictorLoss()(ret["d_outs"], ret["ds"], ret["ilens"])
    for got, k in ((l1, "l1_loss"), (fs, "forward_sum_loss"), (ret["bin_loss"], "bin_loss"), (dur, "duration_loss")):
        assert abs(got.item() - float(z[k])) <= 1e-4 * max(1.0, abs(float(z[k]))), k
    (l1 + 2.0 * (fs + ret["bin_loss"]) + dur).backward()
    gmax = max(np.abs(z[k]).max() for k in z.files if k.startswith("grad."))
    for name, p in model.named_parameters():
        ref = z["grad." + name]
        assert p.grad is not None, name
        assert np.abs(p.grad.cpu().numpy() - ref).max() <= 1e-3 * np.abs(ref).max() + 1e-5 * gmax, name
    att = model.encoder.encoders[0].self_attn.attn
    assert att is not None and tuple(att.shape) == tuple(z["attn.encoder.encoders.0.self_attn"].shape)
    # an optimizer step through the stock torch optimizer moves the engine's flat parameters (shared storage)
    before = model.engine.store.P.clone()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
    opt.step()
    assert not torch.equal(before, model.engine.store.P)


def test_fused_train_step_graph_matches_eager():
    """AASVCTrainStep: CUDA-graph replay == eager launches (same seeds, dropout on), losses finite, parameters move."""
    from seq2seq_vc_b200 import AASVCEngine, AASVCTrainStep

    z, sd = _golden()
    ilens, olens = z["ilens"].tolist(), z["olens"].tolist()
    xs, ys, dpi = (torch.from_numpy(z[k]).cuda() for k in ("xs", "ys", "dp_inputs"))
    outs = []
    for use_graph in (False, True):
        eng = AASVCEngine(dict(AAS_HP), device="cuda:0", bf16=False, seed=11)
        eng.load_state_dict(sd)
        st = AASVCTrainStep(eng, lr=1e-3, warmup_steps=10, use_graph=use_graph)
        ls = []
        for _ in range(4):
            ls.append(st(xs, ilens, ys, olens, dpi).clone())
        torch.cuda.synchronize()
        outs.append((torch.stack(ls).cpu(), eng.store.P.clone().cpu()))
    assert torch.isfinite(outs[0][0]).all() and torch.isfinite(outs[1][0]).all()
    # step 0 runs without the duration loss in both modes; replays see fresh dropout seeds exactly like eager steps
    assert (outs[0][0] - outs[1][0]).abs().max().item() <= 5e-3 * outs[0][0].abs().max().item()
    assert (outs[0][1] - outs[1][1]).abs().max().item() <= 1e-4
    assert (outs[0][1] - sd_flat_like(outs[0][1], sd)).abs().max().item() > 0


def test_fused_train_step_gradient_accumulation():
    """gradient_accumulate_steps = 2 (trainers/aas_vc.py:141-149): non-boundary micro-steps leave the parameters alone,
    CUDA-graph replay == eager launches, and (dropout off) one accumulated step == one step on the hand-averaged gradient."""
    from seq2seq_vc_b200 import AASVCEngine, AASVCTrainStep

    z, sd = _golden()
    ilens, olens = z["ilens"].tolist(), z["olens"].tolist()
    xs, ys, dpi = (torch.from_numpy(z[k]).cuda() for k in ("xs", "ys", "dp_inputs"))
    xs2, ys2 = xs * 0.9, ys * 1.1
    outs = []
    for use_graph in (False, True):
        eng = AASVCEngine(dict(AAS_HP), device="cuda:0", bf16=False, seed=11)
        eng.load_state_dict(sd)
        st = AASVCTrainStep(eng, lr=1e-3, warmup_steps=10, use_graph=use_graph, gradient_accumulate_steps=2)
        for it in range(4):
            before = eng.store.P.clone()
            st(xs, ilens, ys, olens, dpi)
            assert torch.equal(before, eng.store.P) and st.steps == it
            st(xs2, ilens, ys2, olens, dpi)
            assert st.steps == it + 1 and not torch.equal(before, eng.store.P)
        torch.cuda.synchronize()
        outs.append(eng.store.P.clone().cpu())
    assert torch.isfinite(outs[0]).all()
    assert (outs[0] - outs[1]).abs().max().item() <= 1e-4

    acc, ref = (AASVCEngine(dict(AAS_HP, **NO_DROPOUT), device="cuda:0", bf16=False, seed=11) for _ in range(2))
    acc.load_state_dict(sd)
    ref.load_state_dict(sd)
    st = AASVCTrainStep(acc, lr=1e-3, warmup_steps=10, gradient_accumulate_steps=2)
    st(xs, ilens, ys, olens, dpi)
    st(xs2, ilens, ys2, olens, dpi)
    g = torch.zeros_like(ref.store.G)
    for a, b in ((xs, ys), (xs2, ys2)):
        ref.training = True
        ref.prepare(a.shape[0], a.shape[1], b.shape[1], ilens, olens)
        ref.forward(a, b, dpi)
        ref.loss(b, duration_loss=False)
        ref.backward()
        g += ref.store.G
    ref.store.G.copy_(g / 2)
    ref.lr_dev.fill_(st.lr_at(1))
    ref.optimizer_step(1.0, duration_predictor_active=False)       # first window: no duration loss yet (trainers/aas_vc.py:119)
    torch.cuda.synchronize()
    assert (ref.store.P - acc.store.P).abs().max().item() <= 1e-5


def sd_flat_like(flat, sd):
    """Flat parameter vector of the initial state (for 'did anything move' checks)."""
    from seq2seq_vc_b200 import AASVCEngine

    eng = AASVCEngine(dict(AAS_HP), device="cpu", bf16=False, seed=11)
    eng.load_state_dict(sd)
    return eng.store.P.clone()


def test_inference_golden_and_dropin():
    """AASVC.inference (no ground truth) on the GPU: integer durations exact, mel L1 <= 1e-4 vs the live-reference dump."""
    from seq2seq_vc_b200 import AASVC
    from seq2seq_vc_b200.aasvc_engine import AASVCEngine

    z, sd = _golden()
    sd = dict(sd)
    sd.update({k[7:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("inf_bn.")})
    sd["duration_predictor.linear.bias"] = torch.from_numpy(z["inf_dp_bias"])
    il = int(z["ilens"][0])
    x, dpi = torch.from_numpy(z["xs"])[0, :il].cuda(), torch.from_numpy(z["dp_inputs"])[0, :il].cuda()
    eng = AASVCEngine(dict(AAS_HP, **NO_DROPOUT), device="cuda:0", bf16=False)
    eng.load_state_dict(sd)
    outs, d_outs = eng.inference(x, dpi)
    np.testing.assert_array_equal(d_outs.cpu().numpy(), z["inf_d_outs"])
    assert tuple(outs.shape) == z["inf_outs"].shape and np.abs(outs.cpu().numpy() - z["inf_outs"]).mean() <= 1e-4
    model = AASVC(**AAS_HP, positionwise_layer_type="linear", duration_predictor_use_encoder_outputs=False, encoder_normalize_before=True,
                  decoder_normalize_before=True, duration_predictor_type="deterministic", encoder_input_layer="linear").to("cuda:0")
    model.load_state_dict(sd)
    model.eval()
    o2, d2 = model.inference(x, dp_input=dpi)
    assert torch.equal(d2.cpu(), d_outs.cpu()) and np.abs(o2.cpu().numpy() - z["inf_outs"]).mean() <= 1e-4
    # a second, longer utterance re-uses the engine (eval buffers of the previous length are dropped)
    o3, d3 = eng.inference(torch.randn(93, 80, device="cuda"), torch.randn(93, 80, device="cuda"))
    assert o3.shape == (int(d3.sum().item()) or 23, 80) and torch.isfinite(o3).all()


def test_edge_cases_alignment_block(ops):
    """Single-token text, T_feats == T_text (the only feasible path is the diagonal), zero-length rows, T < K depthwise conv."""
    from seq2seq_vc_b200.aasvc_engine import beta_binomial_log_prior

    # one text token: log-softmax over a single column is 0, MAS assigns every frame to it, forward-sum = -sum(lp)/1
    feats, text = rnd(2, 9, 16, seed=1).cuda(), rnd(2, 3, 16, seed=2).cuda()
    tl, fl = torch.tensor([1, 3], dtype=torch.int32).cuda(), torch.tensor([9, 3], dtype=torch.int32).cuda()
    logp, lse = torch.empty(2, 9, 3, device="cuda"), torch.empty(2, 9, device="cuda")
    ops.align_logp_fwd(feats, text, tl, logp, lse)
    assert (logp[0, :, 0] == 0).all() and torch.isinf(logp[0, :, 1:]).all()
    paths, ds, bl, _ = ops.mas(logp, tl, fl)
    assert ds[0].tolist() == [9.0, 0.0, 0.0] and ds[1].tolist() == [1.0, 1.0, 1.0]      # diagonal when T_feats == T_text
    prior = torch.zeros(2, 9, 3)
    prior[0, :9, :1] = beta_binomial_log_prior(1, 9)
    prior[1, :3, :3] = beta_binomial_log_prior(3, 3)
    loss, g = torch.zeros(1, device="cuda"), torch.empty(2, 9, 3, device="cuda")
    ops.forward_sum(logp, prior.cuda(), tl, fl, torch.empty(2, 9, 3, device="cuda"), loss, g)
    lp_ref, g_ref = torch.zeros(1), torch.zeros(2, 9, 3)
    F.forward_sum(logp.cpu(), prior, tl.cpu(), fl.cpu(), None, lp_ref, g_ref)
    assert abs(loss.item() - lp_ref.item()) <= 1e-4 * max(1.0, abs(lp_ref.item()))
    close(g, g_ref, 1e-5, "forward-sum gradient on degenerate lattices")
    # depthwise conv shorter than its kernel, and a zero-row LayerNorm / colsum call
    x, w = rnd(2, 5, 64, seed=3), rnd(64, 15, seed=4, scale=0.3)
    y = F.dwconv_fwd(x, w, None, torch.empty(2, 5, 64))
    close(ops.dwconv_fwd(x.cuda(), w.cuda(), None, torch.empty(2, 5, 64, device="cuda")), y, 2e-5, "dwconv T < K")
    # Gaussian upsampling with all-zero durations in one utterance: uniform weights over the valid tokens
    ds0 = torch.tensor([[0.0, 0.0, 0.0, 0.0], [2.0, 1.0, 0.0, 0.0]])
    P = ops.gauss_weights(ds0.cuda(), torch.tensor([3, 3], dtype=torch.int32).cuda(), torch.tensor([4, 2], dtype=torch.int32).cuda(),
                          torch.empty(2, 3, 8, device="cuda"))
    ref = F.gauss_weights(ds0, torch.tensor([3, 3]), torch.tensor([4, 2]), torch.empty(2, 3, 8))
    close(P, ref, 1e-6, "gaussian weights with zero durations")
    assert torch.isfinite(P).all()


STOCH_HP = dict(AAS_HP, duration_predictor_type="stochastic")


def test_stochastic_duration_predictor_in_the_engine_matches_the_oracle():
    """AAS-VC with the shipped recipe's stochastic duration predictor (aas_vc.melmelmel.v1.yaml:57): dur_nll (models/aas_vc.py:412-419)
    and every gradient of sum(dur_nll) equal the CPU oracle fed with the engine's own predictor input, MAS durations and noise draw;
    the input projection receives no gradient (the predictor detaches its input, duration_predictor.py:236)."""
    from oracle import sdp_oracle
    from seq2seq_vc_b200.aasvc_engine import AASVCEngine, sdp_hparams

    eng = AASVCEngine(dict(STOCH_HP, **NO_DROPOUT, stochastic_duration_predictor_dropout_rate=0.0), device="cuda:0", bf16=False, seed=3)
    g = torch.Generator().manual_seed(9)
    for name in eng.store.names():          # ConvFlow.proj starts at zero (identity splines): perturb so every branch is exercised
        if name.startswith("duration_predictor.") and ((".proj." in name and "flows" in name) or name.endswith((".m", ".logs"))):
            eng.store.p(name).add_((0.2 * torch.randn(eng.store.p(name).shape, generator=g)).cuda())
    z, _ = _golden()
    ilens, olens = z["ilens"].tolist(), z["olens"].tolist()
    xs, ys, dpi = (torch.from_numpy(z[k]).cuda() for k in ("xs", "ys", "dp_inputs"))
    eng.forward(xs, ys, dpi, ilens, olens)
    losses = eng.loss(ys)
    eng.backward()
    torch.cuda.synchronize()
    sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in eng.state_dict().items() if k.startswith("duration_predictor.")}
    ref = sdp_oracle.aasvc_dur_nll(sd, "duration_predictor", sdp_hparams(eng.hp), eng.dp_in.float().cpu(), eng.tlens_host, eng.ds.cpu(), eng._sdp_eq.cpu())
    assert (eng.dur_nll.cpu() - ref.detach()).abs().max().item() <= 1e-4 * ref.detach().abs().max().item()
    assert abs(losses[3].item() - float(ref.sum())) <= 1e-4 * abs(float(ref.sum()))
    ref.sum().backward()
    gmax = max(float(p.grad.abs().max()) for p in sd.values())
    for k, p in sd.items():
        got = eng.store.g(k).cpu()
        assert (got - p.grad).abs().max().item() <= 6e-3 * float(p.grad.abs().max()) + 2e-5 * gmax, k
    for name in eng.store.names():
        if name.startswith("duration_predictor_projection."):
            assert (eng.store.g(name) == 0).all(), name


def test_stochastic_recipe_fused_step_and_dropin_and_inference():
    """The stochastic recipe end to end: fused AASVCTrainStep under CUDA graphs (bf16) trains with finite, decreasing loss; the
    drop-in module returns `dur_nll` and back-propagates sum(dur_nll) into the predictor only; inference draws durations."""
    from seq2seq_vc_b200 import AASVC, AASVCTrainStep, ForwardSumLoss, L1Loss

    z, _ = _golden()
    ilens, olens = z["ilens"].tolist(), z["olens"].tolist()
    xs, ys, dpi = (torch.from_numpy(z[k]).cuda() for k in ("xs", "ys", "dp_inputs"))
    kw = dict(positionwise_layer_type="linear", duration_predictor_use_encoder_outputs=False, encoder_normalize_before=True,
              decoder_normalize_before=True, encoder_input_layer="linear", duration_predictor_type="stochastic")
    model = AASVC(**AAS_HP, **kw, compute_dtype="bf16", device="cuda:0", seed=4)
    step = AASVCTrainStep(model, lr=2e-3, warmup_steps=1, use_graph=True)
    hist = []
    for it in range(30):
        losses = step(xs, ilens, ys, olens, dpi)
        hist.append(losses.clone())
    hist = torch.stack(hist).cpu()
    assert torch.isfinite(hist).all()
    assert hist[-5:, 0].mean() < hist[:5, 0].mean() and hist[-5:, 3].mean() < hist[1:6, 3].mean(), hist[:, [0, 3]]
    # drop-in module through torch autograd with the reference-style loss assembly (trainers/aas_vc.py:73-134)
    m2 = AASVC(**AAS_HP, **kw, compute_dtype="float32", device="cuda:0", seed=4)
    ret = m2(xs, torch.tensor(ilens), ys, torch.tensor(olens), dpi, dp_lengths=torch.tensor(ilens))
    assert "dur_nll" in ret and "d_outs" not in ret and ret["dur_nll"].shape == (xs.shape[0],)
    l1 = L1Loss()(ret["after_outs"], ret["before_outs"], ret["ys"], ret["olens"])
    fs = ForwardSumLoss()(ret["log_p_attn"], ret["ilens"], ret["olens_reduced"])
    total = l1 + 2.0 * (fs + ret["bin_loss"]) + torch.sum(ret["dur_nll"].float())
    total.backward()
    got = {n: p.grad for n, p in m2.named_parameters()}
    assert all(got[n] is not None and torch.isfinite(got[n]).all() for n in got if n.startswith("duration_predictor."))
    assert any(got[n].abs().max() > 0 for n in got if n.startswith("duration_predictor."))
    assert all(got[n] is None for n in got if n.startswith("duration_predictor_projection."))
    assert got["encoder.embed.0.weight"] is not None
    # inference: stochastic durations, clamped to <= 10 (models/aas_vc.py:385-393)
    m2.eval()
    outs, d = m2.inference(xs[0, :ilens[0]], dp_input=dpi[0])
    assert d.dtype == torch.int64 and int(d.max()) <= 10 and outs.shape[0] == int(d.sum()) and torch.isfinite(outs).all()


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_length_regulator_matches_reference_semantics(dtype):
    """seq2seq_vc_b200.LengthRegulator vs the reference's forward (length_regulator.py:69-97: repeat_interleave + pad_list):
    ragged durations with zeros, alpha scaling (torch.round), the all-zero rescue, pad_value, and the gradient (run sums)."""
    from seq2seq_vc_b200 import LengthRegulator

    g = torch.Generator().manual_seed(7)
    B, T, D = 3, 11, 40
    xs = torch.randn(B, T, D, generator=g).to(dtype)
    ds = torch.randint(0, 5, (B, T), generator=g)
    ds[1, 6:] = 0

    def ref(xs, ds, alpha, pad):
        if alpha != 1.0:
            ds = torch.round(ds.float() * alpha).long()
        if ds.sum() == 0:
            ds = ds.clone()
            ds[ds.sum(dim=1).eq(0)] = 1
        rep = [torch.repeat_interleave(x, d, dim=0) for x, d in zip(xs, ds)]
        out = xs.new_full((len(rep), max(r.shape[0] for r in rep), xs.shape[2]), pad)
        for i, r in enumerate(rep):
            out[i, :r.shape[0]] = r
        return out

    for alpha, pad in ((1.0, 0.0), (1.3, -1.5), (0.5, 0.0)):
        x_dev = xs.cuda().requires_grad_(True)
        got = LengthRegulator(pad_value=pad)(x_dev, ds.cuda(), alpha)
        x_ref = xs.clone().float().requires_grad_(True)
        want = ref(x_ref, ds, alpha, pad)
        assert got.shape == want.shape
        assert torch.equal(got.detach().cpu().float(), want.detach().to(dtype).float())
        w = torch.randn(want.shape, generator=g)
        (got.float() * w.cuda()).sum().backward()
        (want * w).sum().backward()
        tol = 1e-5 if dtype == torch.float32 else 5e-2
        assert torch.allclose(x_dev.grad.cpu().float(), x_ref.grad, atol=tol, rtol=tol)
    z = LengthRegulator()(xs.cuda(), torch.zeros(B, T, dtype=torch.long).cuda())
    assert torch.equal(z.cpu(), xs)                                   # all durations 0 -> every duration becomes 1
