"""Stochastic duration predictor on the GPU (seq2seq_vc_b200/sdp.py over csrc/ops_sdp.cu) vs the live-reference dump
tests/golden/sdp_tiny.npz (noise recorded from the reference's own draw) and vs the CPU oracle at a wider shape."""
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden", "sdp_tiny.npz")
SDP_HP = dict(channels=16, kernel_size=3, dds_conv_layers=3, flows=4)


def _load():
    z = np.load(GOLD)
    sd = {k[3:]: torch.from_numpy(z[k]).cuda().requires_grad_(True) for k in z.files if k.startswith("sd.")}
    return z, sd


def _predictor(sd, hp, mode=2):
    from seq2seq_vc_b200.sdp import StochasticDurationPredictor

    return StochasticDurationPredictor(hp, "duration_predictor", lambda n: sd[n[len("duration_predictor."):]], gemm_mode=mode, dropout_rate=0.0)


def _masks(text_lens, T):
    tl = torch.tensor(text_lens, dtype=torch.int32).cuda()
    maskf = (torch.arange(T)[None, :] < torch.tensor(text_lens)[:, None]).float().reshape(-1).cuda()
    return tl, maskf


@pytest.mark.parametrize("mode", [0, 2])
def test_nll_and_every_gradient_match_the_reference(mode):
    """dur_nll = nll / sum(mask) (models/aas_vc.py:412-419): value <= 1e-5 relative, every parameter gradient of sum(dur_nll)."""
    z, sd = _load()
    sdp = _predictor(sd, SDP_HP, mode)
    B, T, C = z["dp_inputs"].shape
    tl, maskf = _masks(z["text_lens"].tolist(), T)
    nll = sdp.nll(torch.from_numpy(z["dp_inputs"]).cuda(), tl, maskf, torch.from_numpy(z["ds"]).float().cuda(), torch.from_numpy(z["e_q"]).cuda())
    dur_nll = nll / float(z["text_lens"].sum())
    ref = z["dur_nll"]
    assert np.abs(dur_nll.detach().cpu().numpy() - ref).max() <= (1e-5 if mode == 0 else 5e-5) * np.abs(ref).max()
    dur_nll.sum().backward()
    torch.cuda.synchronize()
    gmax = max(np.abs(z[k]).max() for k in z.files if k.startswith("grad."))
    n = 0
    for k in z.files:
        if k.startswith("grad."):
            g = sd[k[5:]].grad
            assert g is not None, k
            # mode 2 (fp32-accurate tcgen05 GEMM): the tensor core's fp32 accumulation moves the log / softmax / spline chains a little more
            assert np.abs(g.cpu().numpy() - z[k]).max() <= (2e-3 if mode == 0 else 6e-3) * np.abs(z[k]).max() + 1e-5 * gmax, k
            n += 1
    assert n == len(sd)


def test_inverse_durations_exact():
    """Inference direction (models/aas_vc.py:385-393): integer durations bit-exact, incl. the wide-noise draw in the spline tails."""
    z, sd = _load()
    sdp = _predictor(sd, SDP_HP, 0)
    B, T, C = z["dp_inputs"].shape
    tl, maskf = _masks(z["text_lens"].tolist(), T)
    x = torch.from_numpy(z["dp_inputs"]).cuda()
    for zk, dk, scale in (("z", "d_outs", 0.8), ("z_wide", "d_outs_wide", 4.0)):
        d = sdp.inverse(x, tl, maskf, torch.from_numpy(z[zk]).cuda(), noise_scale=scale)
        np.testing.assert_array_equal(d.cpu().numpy(), z[dk])


def test_spline_kernel_round_trip_gradient_and_tails():
    """Size-independent properties of the spline kernels: inverse(forward(x)) == x with cancelling log-determinants, identity
    outside +-5, and the dual-number gradient against central differences of the forward kernel."""
    import ctypes

    from seq2seq_vc_b200 import _lib
    from seq2seq_vc_b200._lib import check, ptr, stream

    L = _lib.load()
    g = torch.Generator().manual_seed(7)
    B, T, hidden = 4, 50, 16.0
    x = (torch.randn(B, 2, T, generator=g) * 3.0).cuda()
    h = torch.randn(B, T, 29, generator=g).cuda()
    tl = torch.tensor([50, 41, 17, 50], dtype=torch.int32).cuda()
    y, lad = torch.empty_like(x), torch.empty(B, T, device="cuda")
    check(L.s2s_rq_spline_fwd(ptr(x) + 4 * T, 2 * T, ptr(h), ptr(tl), ptr(y), 2 * T, ptr(lad), B, T, hidden, 0, stream()))
    xr, ladr = torch.empty_like(x), torch.empty(B, T, device="cuda")
    check(L.s2s_rq_spline_fwd(ptr(y), 2 * T, ptr(h), ptr(tl), ptr(xr), 2 * T, ptr(ladr), B, T, hidden, 1, stream()))
    mask = (torch.arange(T)[None, :] < tl.cpu()[:, None]).cuda()
    xb = x[:, 1]
    assert ((xr[:, 0] - xb) * mask).abs().max().item() <= 5e-4 and ((lad + ladr) * mask).abs().max().item() <= 5e-3
    out = (xb.abs() > 5.0) & mask
    assert out.any() and torch.equal(y[:, 0][out], xb[out]) and (lad[out] == 0).all()
    assert (y[:, 0][~mask] == 0).all() and (lad[~mask] == 0).all()
    # gradient: dual numbers vs central differences in float64-ish steps
    gy, gl = torch.randn(B, T, generator=g).cuda(), torch.randn(B, T, generator=g).cuda()
    gyz = torch.zeros_like(x)
    gyz[:, 0] = gy
    dx, dh = torch.zeros_like(x), torch.empty_like(h)
    check(L.s2s_rq_spline_bwd(ptr(x) + 4 * T, 2 * T, ptr(h), ptr(tl), ptr(gyz), 2 * T, ptr(gl), ptr(dx) + 4 * T, 2 * T, ptr(dh), B, T, hidden, stream()))

    def f(hh):
        yy, ll = torch.empty_like(x), torch.empty(B, T, device="cuda")
        check(L.s2s_rq_spline_fwd(ptr(x) + 4 * T, 2 * T, ptr(hh), ptr(tl), ptr(yy), 2 * T, ptr(ll), B, T, hidden, 0, stream()))
        return (gy * yy[:, 0] + gl * ll).double()

    eps = 1e-2
    for j in (0, 7, 13, 19, 22, 28):
        e = torch.zeros_like(h)
        e[..., j] = eps
        num = ((f(h + e) - f(h - e)) / (2 * eps)).float()
        inside = mask & (xb.abs() < 4.5)
        err = ((num - dh[..., j]) * inside).abs().max().item()
        assert err <= 5e-2 * max(1.0, dh[..., j].abs().max().item()), (j, err)       # fp32 central differences across moving knots


def test_wide_shape_vs_cpu_oracle_and_dropout_runs():
    """The recipe's width (C = 384, T_text = 48, B = 4) against the CPU oracle on the same weights / noise; then a step with the
    reference's dropout rate 0.5 in the two conditioning stacks (finite, different from the dropout-free value)."""
    from oracle import sdp_oracle
    from seq2seq_vc_b200 import sdp as S
    from seq2seq_vc_b200._lib import Drop

    hp = dict(channels=384, kernel_size=3, dds_conv_layers=3, flows=4)
    sd = S.init_params(hp, "duration_predictor", seed=3)
    g = torch.Generator().manual_seed(5)
    for k, v in sd.items():                        # de-trivialise the zero-initialised spline heads and affine flows
        if (".proj." in k and "flows" in k) or k.endswith((".m", ".logs")):
            v.add_(0.05 * torch.randn(v.shape, generator=g))
    B, T, C = 4, 48, 384
    tlens = [48, 40, 31, 7]
    x = torch.randn(B, T, C, generator=g)
    ds = (torch.randint(0, 9, (B, T), generator=g) * (torch.arange(T)[None, :] < torch.tensor(tlens)[:, None])).float()
    e_q = torch.randn(B, 2, T, generator=g)
    cpu = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    ref = sdp_oracle.aasvc_dur_nll(cpu, "duration_predictor", hp, x, tlens, ds, e_q)
    ref.sum().backward()
    dev = {k: v.clone().cuda().requires_grad_(True) for k, v in sd.items()}
    tl, maskf = _masks(tlens, T)
    pred = S.StochasticDurationPredictor(hp, "duration_predictor", lambda n: dev[n], gemm_mode=2, dropout_rate=0.0)
    nll = pred.nll(x.cuda(), tl, maskf, ds.cuda(), e_q.cuda()) / float(sum(tlens))
    assert (nll.detach().cpu() - ref.detach()).abs().max().item() <= 1e-4 * ref.detach().abs().max().item()
    nll.sum().backward()
    gmax = max(float(p.grad.abs().max()) for p in cpu.values())
    for k in cpu:
        a, b = dev[k].grad.cpu(), cpu[k].grad
        assert (a - b).abs().max().item() <= 5e-3 * float(b.abs().max()) + 2e-5 * gmax, k
    seed_dev = torch.zeros(1, dtype=torch.int64, device="cuda")
    predd = S.StochasticDurationPredictor(hp, "duration_predictor", lambda n: dev[n], gemm_mode=2, dropout_rate=0.5,
                                          drop_of=lambda name, p: Drop(p, 11, hash(name) % 100000, seed_dev))
    nd = predd.nll(x.cuda(), tl, maskf, ds.cuda(), e_q.cuda())
    nd.sum().backward()
    assert torch.isfinite(nd).all() and (nd.detach() - nll.detach() * sum(tlens)).abs().max().item() > 1e-3
    zn = S.randn((64, 2, 192), "cuda", 5, seed_dev, 3)
    assert abs(zn.mean().item()) < 0.02 and abs(zn.std().item() - 1.0) < 0.02
