"""Conv2dSubsampling2 / 6 / 8 drop-ins (modules/transformer/subsampling.py:108-279): the oracle vs golden vectors from the live
reference, the drop-in modules (kernels replaced by their CPU contracts, tests/fake_ops.py) vs the golden vectors and -- in the build
container -- vs the live reference modules side by side (state-dict keys, outputs, masks, every parameter gradient, a custom
pos_enc module)."""
import os

import numpy as np
import pytest
import torch

import fake_ops
from oracle import subsampling_oracle as so

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "subsampling_tiny.npz")


def _case(z, n):
    p = f"s{n}."
    sd = {k[len(p) + 3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith(p + "sd.")}
    grads = {k[len(p) + 5:]: z[k] for k in z.files if k.startswith(p + "grad.")}
    return sd, grads, torch.from_numpy(z[p + "x"]), z[p + "y"], torch.from_numpy(z[p + "r"]), torch.from_numpy(z[p + "mask_in"]), z[p + "mask_out"]


@pytest.mark.parametrize("n", [2, 6, 8])
def test_oracle_matches_golden(n):
    sd, grads, x, y, r, mask, mask_out = _case(np.load(GOLDEN), n)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    out, m = so.conv2d_subsampling(params, n, x, mask)
    assert np.abs(out.detach().numpy() - y).max() <= 2e-5
    np.testing.assert_array_equal(m.numpy(), mask_out)
    (out * r).sum().backward()
    for k, g in grads.items():
        assert np.abs(params[k].grad.numpy() - g).max() <= 1e-4 * max(1.0, np.abs(g).max()), k


@pytest.mark.parametrize("n", [2, 6, 8])
def test_dropin_matches_golden(monkeypatch, n):
    import seq2seq_vc_b200

    fake_ops.install(monkeypatch)
    sd, grads, x, y, r, mask, mask_out = _case(np.load(GOLDEN), n)
    m = getattr(seq2seq_vc_b200, f"Conv2dSubsampling{n}")(40, 16, 0.0)
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == {k: tuple(v.shape) for k, v in sd.items()}
    m.load_state_dict(sd)
    m.train()
    out, mo = m(x, mask)
    assert np.abs(out.detach().numpy() - y).max() <= 2e-5
    np.testing.assert_array_equal(mo.numpy(), mask_out)
    (out * r).sum().backward()
    for k, p in m.named_parameters():
        assert np.abs(p.grad.numpy() - grads[k]).max() <= 1e-4 * max(1.0, np.abs(grads[k]).max()), k
    assert m(x, None)[1] is None
    with pytest.raises(NotImplementedError):
        m[0]
    assert m[-1] is m.out[1] if hasattr(m.out, "__getitem__") else m[-1] is getattr(m.out, "1")


@pytest.mark.parametrize("n", [2, 6, 8])
def test_dropin_matches_live_reference_with_custom_pos_enc(monkeypatch, n):
    from oracle import ref_shim

    if not ref_shim.available():
        pytest.skip("reference tree not present")
    fake_ops.install(monkeypatch)
    ref_shim.install()
    import seq2seq_vc.modules.transformer.subsampling as rs
    from seq2seq_vc.layers.positional_encoding import ScaledPositionalEncoding

    import seq2seq_vc_b200

    torch.manual_seed(n)
    ref = getattr(rs, f"Conv2dSubsampling{n}")(33, 24, 0.0, ScaledPositionalEncoding(24, 0.0))
    ours = getattr(seq2seq_vc_b200, f"Conv2dSubsampling{n}")(33, 24, 0.0, ScaledPositionalEncoding(24, 0.0))
    assert [k for k, _ in ours.named_parameters()] == [k for k, _ in ref.named_parameters()]
    ours.load_state_dict(ref.state_dict())
    x = torch.randn(3, 47, 33)
    mask = torch.ones(3, 1, 47, dtype=torch.bool)
    mask[2, :, 30:] = False
    a, am = ref(x, mask)
    b, bm = ours(x, mask)
    assert (a - b).abs().max().item() <= 2e-5 and torch.equal(am, bm)
    r = torch.randn(a.shape)
    (a * r).sum().backward()
    (b * r).sum().backward()
    gr = dict(ref.named_parameters())
    for k, p in ours.named_parameters():
        assert (p.grad - gr[k].grad).abs().max().item() <= 1e-4 * max(1.0, gr[k].grad.abs().max().item()), k
    assert ours[-1] is ours.out[1] if hasattr(ours.out, "__getitem__") else True
