"""The C-ABI library loads and exports every symbol include/s2svc_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "s2svc_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(s2s_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_entry_points():
    syms = declared_symbols()
    assert "s2s_gemm" in syms and "s2s_mas" in syms and "s2s_logmel" in syms and len(syms) >= 35


def test_library_exports_every_declared_symbol():
    from seq2seq_vc_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__

        __graft_entry__.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_python_binding_covers_header():
    from seq2seq_vc_b200 import _lib

    assert sorted(_lib.SIGNATURES) == declared_symbols()
    lib = _lib.load()
    assert lib.s2s_abi_version() == _lib.ABI_VERSION
    assert _lib.launch_count() == 0 or _lib.launch_count() > 0


def test_no_cpu_fallback():
    import torch

    from seq2seq_vc_b200 import S2SError, VTN

    m = VTN(idim=80, odim=80, adim=32, aheads=2, elayers=1, dlayers=1, eunits=32, dunits=32, dprenet_units=16,
            postnet_layers=2, postnet_chans=16)
    xs, ys = torch.zeros(1, 20, 80), torch.zeros(1, 8, 80)
    with pytest.raises(S2SError):
        m(xs, [20], ys, torch.zeros(1, 8), [8])


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "seq2seq_vc_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            assert "oracle" not in open(os.path.join(pkg, fn)).read().replace("no oracle", ""), fn


def test_argument_validation_returns_error_codes_without_a_gpu():
    """Entry points validate before launching: bad arguments give S2S_ERR_INVALID / S2S_ERR_UNSUPPORTED and a message
    through s2s_last_error(), on any machine (no kernel is launched)."""
    from seq2seq_vc_b200 import _lib

    lib = _lib.load()
    lib.s2s_last_error.restype = ctypes.c_char_p
    assert lib.s2s_forward_sum(None, None, None, None, 1, 8, 4, -1.0, None, None, None, 1.0, None) == -1
    assert b"forward_sum" in lib.s2s_last_error()
    one = ctypes.c_void_p(16)          # never dereferenced: validation fails first
    assert lib.s2s_forward_sum(one, one, one, one, 2, 100, 600, -1.0, one, one, None, 1.0, None) == -1
    assert b"510" in lib.s2s_last_error()
    assert lib.s2s_dwconv_fwd(one, one, None, one, 1, 10, 8, 4, 0, None) == -1            # even kernel size
    assert b"odd K" in lib.s2s_last_error()
    assert lib.s2s_attn_probs_fwd(one, 8, 8, 8, one, 8, 8, 8, one, None, 1, 1, 8, 8, 40, 8, 1.0, 0, None) == -2   # unsupported d_k
    assert b"d_k" in lib.s2s_last_error()
    assert lib.s2s_attn_probs_fwd(one, 8, 8, 8, one, 8, 8, 8, one, None, 1, 1, 8, 9, 48, 12, 1.0, 0, None) == -1  # ld % 8 != 0
    assert lib.s2s_relshift_add(one, one, 1, 1, 8, 8, 14, 0, None) == -1                  # ldB < 2T-1
    assert lib.s2s_glu_fwd(None, one, 4, 8, 0, None) == -1
    assert lib.s2s_gather_rows(one, None, one, one, 1, 4, 4, 8, 0, None) == -1
    assert lib.s2s_align_logp_fwd(one, one, one, one, one, 0, 4, 4, 8, 0, None) == -1     # B = 0
    assert lib.s2s_glu_fwd(one, one, 4, 8, 7, None) == -1 and b"dtype" in lib.s2s_last_error()
