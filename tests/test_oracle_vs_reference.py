"""Pins the oracle against the LIVE reference (build container only; skipped where /root/reference is absent)."""
import numpy as np
import pytest
import torch

from oracle import mas_oracle, ref_shim, vtn_oracle

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref():
    ref_shim.install()
    import seq2seq_vc.models  # noqa: F401
    return True


def test_vtn_c1_shape_against_live_reference(ref):
    from seq2seq_vc.losses import Seq2SeqLoss
    from seq2seq_vc.models import VTN

    hp = vtn_oracle.default_hparams(adim=64, aheads=4, elayers=2, dlayers=2, eunits=128, dunits=128, dprenet_units=32,
                                    postnet_chans=32)
    torch.manual_seed(3)
    model = VTN(dprenet_dropout_rate=0.0, **hp)
    ref_shim.disable_dropout(model)
    model.train()
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    xs, ilens, ys, labels, olens = vtn_oracle.synthetic_batch(2, 60, 50, ilens=[60, 47], olens=[50, 33], seed=5)
    out = model(xs, torch.tensor(ilens), ys, labels, torch.tensor(olens))
    l1, bce = Seq2SeqLoss()(*out[:6])
    o = vtn_oracle.vtn_forward(sd, hp, xs, ilens, ys, labels, olens, training=True)
    assert (o["after_outs"] - out[0]).abs().max() <= 2e-5
    assert (o["logits"] - out[2]).abs().max() <= 2e-5
    l1o, bceo = vtn_oracle.seq2seq_loss(o["after_outs"], o["before_outs"], o["logits"], o["ys"], o["labels"], o["olens"])
    assert abs(float(l1o) - float(l1)) <= 1e-6 and abs(float(bceo) - float(bce)) <= 1e-6


def test_mas_fuzz_against_numba(ref):
    from seq2seq_vc.modules.alignments import _monotonic_alignment_search

    rng = np.random.default_rng(0)
    for n in range(40):
        tm = int(rng.integers(1, 90))
        ti = int(rng.integers(1, min(tm, 40) + 1))
        lp = torch.log_softmax(torch.from_numpy(rng.standard_normal((tm, ti)).astype(np.float32)), -1).numpy()
        if n % 3 == 1:
            lp = (np.round(lp * 2) / 2).astype(np.float32)
        ref_path = _monotonic_alignment_search(lp)
        paths, _ = mas_oracle.mas_batch_c(lp[None], [ti], [tm])
        np.testing.assert_array_equal(paths[0], ref_path)
