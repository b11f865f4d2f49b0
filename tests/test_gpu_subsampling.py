"""Conv2dSubsampling2 / 6 / 8 on the GPU through the C ABI: the generic patch gather / scatter kernels vs their contracts
(tests/fake_ops.py; the gather is a copy and must be bit-exact, <gather(x), y> == <x, scatter(y)>), and the drop-in modules vs golden
vectors dumped from the live reference (tests/golden/subsampling_tiny.npz): fp32 on the CUDA-core GEMM and on the fp32-accurate
tcgen05 GEMM within 1e-4, bf16 drift-bounded."""
import os

import numpy as np
import pytest
import torch

import fake_ops as F

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "subsampling_tiny.npz")


@pytest.fixture(scope="module")
def ops():
    from seq2seq_vc_b200 import _lib, ops

    _lib.device_check()
    return ops


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,T1,F1,C,k,s", [(2, 30, 19, 16, 3, 1), (2, 30, 19, 16, 5, 3), (3, 29, 17, 24, 3, 2), (1, 7, 5, 12, 5, 3), (2, 11, 9, 7, 3, 2)])
def test_im2col2d_col2im2d(ops, dt, B, T1, F1, C, k, s):
    g = torch.Generator().manual_seed(T1 * k + s)
    y = torch.randn(B, T1, F1, C, generator=g).to(dt)
    T2, F2 = (T1 - k) // s + 1, (F1 - k) // s + 1
    col = F.im2col2d(y, torch.empty(B * T2 * F2, k * k * C, dtype=dt), k, s)
    gcol = ops.im2col2d(y.cuda(), torch.full((B * T2 * F2, k * k * C), 7.0, dtype=dt, device="cuda"), k, s)
    assert torch.equal(gcol.cpu(), col)
    dcol = torch.randn(B * T2 * F2, k * k * C, generator=g).to(dt)
    for gate in (None, y):
        ref = F.col2im2d(dcol, gate, torch.empty(B, T1, F1, C, dtype=dt), k, s)
        got = ops.col2im2d(dcol.cuda(), None if gate is None else gate.cuda(), torch.full((B, T1, F1, C), 7.0, dtype=dt, device="cuda"), k, s)
        tol = 1e-5 if dt == torch.float32 else 4e-2
        assert (got.cpu().float() - ref.float()).abs().max().item() <= tol
    if dt == torch.float32:                                     # adjointness
        lhs = (col.double() * dcol.double()).sum()
        rhs = (y.double() * F.col2im2d(dcol, None, torch.empty(B, T1, F1, C), k, s).double()).sum()
        assert abs(float(lhs - rhs)) <= 1e-4 * max(1.0, abs(float(lhs)))


def _case(z, n):
    p = f"s{n}."
    sd = {k[len(p) + 3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith(p + "sd.")}
    grads = {k[len(p) + 5:]: z[k] for k in z.files if k.startswith(p + "grad.")}
    return sd, grads, torch.from_numpy(z[p + "x"]), z[p + "y"], torch.from_numpy(z[p + "r"]), torch.from_numpy(z[p + "mask_in"]), z[p + "mask_out"]


@pytest.mark.parametrize("mode", ["float32_simt", "float32", "bf16"])
@pytest.mark.parametrize("n", [2, 6, 8])
def test_dropin_modules_match_reference_golden(n, mode):
    import seq2seq_vc_b200

    sd, grads, x, y, r, mask, mask_out = _case(np.load(GOLDEN), n)
    m = getattr(seq2seq_vc_b200, f"Conv2dSubsampling{n}")(40, 16, 0.0, compute_dtype=mode).cuda()
    m.load_state_dict(sd)
    m.train()
    out, mo = m(x.cuda(), mask.cuda())
    np.testing.assert_array_equal(mo.cpu().numpy(), mask_out)
    err = np.abs(out.detach().cpu().numpy() - y)
    (out * r.cuda()).sum().backward()
    if mode == "bf16":
        assert err.mean() <= 3e-2 * max(1.0, np.abs(y).mean())
        a = torch.cat([p.grad.reshape(-1).cpu() for _, p in m.named_parameters()])
        b = torch.cat([torch.from_numpy(grads[k]).reshape(-1) for k, _ in m.named_parameters()])
        assert torch.nn.functional.cosine_similarity(a, b, dim=0).item() >= 0.99
    else:
        assert err.mean() <= 1e-4 and err.max() <= 1e-3
        for k, p in m.named_parameters():
            g = grads[k]
            assert np.abs(p.grad.cpu().numpy() - g).max() <= 2e-3 * np.abs(g).max() + 1e-5, k


def test_default_positional_encoding_dropout_and_input_too_short():
    import seq2seq_vc_b200
    from seq2seq_vc_b200._lib import S2SError

    m = seq2seq_vc_b200.Conv2dSubsampling8(40, 16, 0.5).cuda()
    m.train()
    x = torch.randn(2, 90, 40, device="cuda")
    a, _ = m(x, None)
    b, _ = m(x, None)
    frac = (a == 0).float().mean().item()
    assert 0.4 < frac < 0.6 and not torch.equal(a, b)            # the default pos_enc applies dropout, fresh masks per call
    a.sum().backward()
    assert all(torch.isfinite(p.grad).all() for p in m.parameters())
    m.eval()
    c, _ = m(x, None)
    assert (c == 0).float().mean().item() < 0.01
    with pytest.raises(S2SError):
        m(torch.randn(1, 10, 40, device="cuda"), None)
