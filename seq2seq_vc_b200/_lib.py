"""ctypes binding of libs2svc_b200.so (the C ABI declared in include/s2svc_b200.h).

This is the only place the package touches native code.  There is NO fallback: if the shared
library is missing or a call fails, an exception is raised ("no CPU
fallback").  PyTorch tensors are used purely as device-memory handles; every call passes raw
``data_ptr()`` values and the current CUDA stream.
"""
from __future__ import annotations

import contextlib
import ctypes
import gc
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_uint64, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libs2svc_b200.so")

S2S_F32, S2S_BF16 = 0, 1
ABI_VERSION = 22


class S2SError(RuntimeError):
    pass


class DropoutDesc(Structure):
    _fields_ = [("p", c_float), ("seed", c_uint64), ("stream", c_uint64), ("seed_dev", c_void_p)]


class GemmDesc(Structure):
    _fields_ = [
        ("M", c_int), ("N", c_int), ("K", c_int), ("taps", c_int),
        ("A", c_void_p), ("a_dtype", c_int), ("a_rs", c_int64), ("a_cs", c_int64), ("a_bs1", c_int64), ("a_bs2", c_int64),
        ("B", c_void_p), ("b_dtype", c_int), ("b_rs", c_int64), ("b_cs", c_int64), ("b_ts", c_int64), ("b_bs1", c_int64), ("b_bs2", c_int64),
        ("C", c_void_p), ("c_dtype", c_int), ("c_rs", c_int64), ("c_bs1", c_int64), ("c_bs2", c_int64),
        ("batch1", c_int), ("batch2", c_int),
        ("bias", c_void_p),
        ("R", c_void_p),
        ("alpha", c_float),
        ("relu", c_int),
        ("accumulate", c_int),
        ("drop", DropoutDesc),
        ("mask_period", c_int), ("mask_offset", c_int), ("mask_lo", c_int), ("mask_hi", c_int),
        ("r_mode", c_int), ("r_scale", c_float), ("split_terms", c_int), ("ws", c_void_p), ("ws_bytes", c_size_t),
    ]


class ColsumDesc(Structure):
    _fields_ = [("x", c_void_p), ("rows", c_int64), ("cols", c_int), ("ld", c_int64), ("out", c_void_p)]


# name -> (restype, argtypes).  Every symbol declared in include/s2svc_b200.h appears here; the CPU
# test-suite checks that the library exports each of them.
_P = c_void_p
_DP = POINTER(DropoutDesc)
SIGNATURES = {
    "s2s_last_error": (c_char_p, []),
    "s2s_abi_version": (c_int, []),
    "s2s_device_check": (c_int, []),
    "s2s_launch_count": (c_int64, []),
    "s2s_tc_fallback_count": (c_int64, []),
    "s2s_debug_gemm_tile": (None, [c_int]),
    "s2s_gemm": (c_int, [POINTER(GemmDesc), c_int, _P]),
    "s2s_gemm_workspace_bytes": (c_size_t, [POINTER(GemmDesc)]),
    "s2s_gemm_grouped": (c_int, [POINTER(GemmDesc), c_int, c_int, _P]),
    "s2s_colsum_multi": (c_int, [POINTER(ColsumDesc), c_int, c_int, _P]),
    "s2s_layernorm_fwd": (c_int, [_P, _P, _P, _P, _P, _P, c_int64, c_int, c_float, c_int, _P]),
    "s2s_layernorm_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int, c_int, _P]),
    "s2s_layernorm_bwd_drop": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _DP, _P, _P, c_int64, c_int, c_int, _P]),
    "s2s_skinny_linear_fwd": (c_int, [_P, _P, _P, _P, c_int64, c_int, c_int, c_int, _P]),
    "s2s_skinny_linear_bwd": (c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int64, c_int, c_int, c_int, _P]),
    "s2s_colsum": (c_int, [_P, c_int64, c_int, c_int64, _P, c_int, _P]),
    "s2s_relu_bwd": (c_int, [_P, _P, _P, c_int64, c_float, c_int, _P]),
    "s2s_dropout_bwd": (c_int, [_P, _P, c_int64, c_int, _DP, c_int, _P]),
    "s2s_add": (c_int, [_P, _P, _P, c_int64, c_int, _P]),
    "s2s_softmax_fwd": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int64, c_int, _DP, c_int, _P]),
    "s2s_softmax_bwd": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int64, c_float, _DP, c_int, _P]),
    "s2s_attn_probs_fwd": (c_int, [_P, c_int64, c_int64, c_int64, _P, c_int64, c_int64, c_int64, _P, _P, c_int, c_int, c_int, c_int, c_int,
                                   c_int64, c_float, c_int, _P]),
    "s2s_attn_probs_bwd": (c_int, [_P, c_int64, c_int64, c_int64, _P, c_int64, c_int64, c_int64, _P, _P, _P, c_int, c_int, c_int, c_int,
                                   c_int, c_int64, c_float, _P]),
    "s2s_attn_fwd_tc": (c_int, [_P, c_int64, c_int64, c_int64, _P, _P, c_int64, c_int64, c_int64, _P, c_int64, c_int64, c_int64, _P, _P, c_int64,
                                _P, c_int, c_int, c_int, c_int, c_int, c_float, c_int, _P]),
    "s2s_attn_bwd_tc": (c_int, [_P, c_int64, c_int64, c_int64, _P, _P, c_int64, c_int64, c_int64, _P, _P, c_int64, c_int64, c_int64, _P, _P,
                                _P, c_int64, c_int64, c_int64, _P, _P, c_int64, c_int64, c_int64, _P, c_int, c_int, c_int, c_int, c_int,
                                c_float, c_int, _P]),
    "s2s_scaled_pe_fwd": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, _DP, c_int, _P]),
    "s2s_scaled_pe_bwd": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, _DP, c_int, _P]),
    "s2s_embed_pe_fwd": (c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _DP, c_int, _P]),
    "s2s_embed_pe_bwd": (c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _DP, c_int, _P]),
    "s2s_conv1_fwd": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P]),
    "s2s_conv1_bwd": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P]),
    "s2s_im2col_s2": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, _P]),
    "s2s_col2im_s2": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, _P]),
    "s2s_col2im_s2_relu": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P]),
    "s2s_shift_thin": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P]),
    "s2s_fix_targets": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, _P]),
    "s2s_bn_stats": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, _P]),
    "s2s_bn_finalize": (c_int, [_P, _P, _P, _P, _P, c_int64, c_int, c_float, c_float, _P]),
    "s2s_bn_apply": (c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _DP, c_int, _P]),
    "s2s_bn_bwd_reduce": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _DP, c_int, _P]),
    "s2s_bn_bwd_apply": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _DP, c_int, _P]),
    "s2s_bn_eval_stats": (c_int, [_P, _P, _P, _P, c_int, c_float, _P]),
    "s2s_pack_conv1d_w": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, _P]),
    "s2s_pad_rows": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, _P]),
    "s2s_unpad_rows": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, _P]),
    "s2s_seq2seq_loss": (c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_float, _P, _P, _P, _P, _P, c_int, _P]),
    "s2s_guided_attn_loss": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_int64, c_float, c_float, _P, _P, _P, c_int, _P]),
    "s2s_sqnorm": (c_int, [_P, c_int64, _P, _P]),
    "s2s_adam_step": (c_int, [_P, _P, _P, _P, _P, c_int64, _P, c_float, c_float, c_float, c_float, _P, _P, c_float, c_float, _P]),
    "s2s_step_advance": (c_int, [_P, _P, _P]),
    "s2s_cast": (c_int, [_P, _P, c_int64, c_int, c_int, _P]),
    "s2s_transpose_last2": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P]),
    "s2s_mas_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "s2s_mas": (c_int, [_P, _P, _P, c_int, c_int, c_int, _P, _P, _P, _P, _P, c_size_t, _P]),
    "s2s_logmel": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_float, c_float, _P]),
    "s2s_gl_istft": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, _P]),
    "s2s_gl_stft": (c_int, [_P, _P, _P, c_int, c_int64, c_int, c_int, c_int, _P]),
    "s2s_gl_update": (c_int, [_P, _P, _P, c_int64, c_float, _P]),
    "s2s_row_sqnorm": (c_int, [_P, _P, c_int64, c_int, c_int, _P]),
    "s2s_align_logp_from_dot": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, _P]),
    "s2s_conv1_xcol": (c_int, [_P, _P, c_int, c_int, c_int, c_int, _P]),
    "s2s_conv1_dw_scatter": (c_int, [_P, _P, _P, c_int, _P]),
    "s2s_conv1_pack_w": (c_int, [_P, _P, _P, c_int, c_int, _P]),
    "s2s_im2col2d": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P]),
    "s2s_col2im2d": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P]),
    "s2s_feat_stats": (c_int, [_P, _P, _P, c_int, c_int, c_int, _P]),
    "s2s_logmel_norm": (c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_float, c_float, _P]),
    "s2s_bias_add2": (c_int, [_P, c_int64, _P, _P, _P, _P, c_int64, c_int, c_int, _P]),
    "s2s_add_strided": (c_int, [_P, _P, _P, c_int64, c_int64, c_int, c_int, _P]),
    "s2s_relshift_add": (c_int, [_P, _P, c_int, c_int, c_int, c_int64, c_int64, c_int, _P]),
    "s2s_relshift_bwd": (c_int, [_P, _P, c_int, c_int, c_int, c_int64, c_int64, c_int, _P]),
    "s2s_relshift_legacy_add": (c_int, [_P, _P, c_int, c_int, c_int, c_int64, c_int64, c_int, _P]),
    "s2s_relshift_legacy_bwd": (c_int, [_P, _P, c_int, c_int, c_int, c_int64, c_int64, c_int, _P]),
    "s2s_glu_fwd": (c_int, [_P, _P, c_int64, c_int, c_int, _P]),
    "s2s_glu_bwd": (c_int, [_P, _P, _P, c_int64, c_int, c_int, _P]),
    "s2s_dwconv_fwd": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P]),
    "s2s_dwconv_bwd": (c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P]),
    "s2s_swish_fwd": (c_int, [_P, _P, c_int64, _DP, c_int, _P]),
    "s2s_swish_bwd": (c_int, [_P, _P, _P, c_int64, _DP, c_int, _P]),
    "s2s_scale_dropout": (c_int, [_P, _P, c_int64, c_float, _DP, _DP, c_int, _P]),
    "s2s_axpy": (c_int, [_P, _P, c_int64, c_float, c_int, _P]),
    "s2s_rowscale": (c_int, [_P, _P, _P, c_int64, c_int, c_int, _P]),
    "s2s_gather_rows": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P]),
    "s2s_lr_cumsum": (c_int, [_P, _P, c_int, c_int, c_float, c_int, _P]),
    "s2s_lr_fwd": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_float, c_int, _P]),
    "s2s_lr_bwd": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P]),
    "s2s_align_logp_fwd": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P]),
    "s2s_align_logp_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int64, c_int, _P]),
    "s2s_forward_sum": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_float, _P, _P, _P, c_float, _P]),
    "s2s_gauss_weights": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int64, c_float, c_int, _P]),
    "s2s_gemv": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, _DP, _P, c_int, _P]),
    "s2s_decode_attn": (c_int, [_P, _P, _P, _P, _P, c_int64, c_int, c_int, c_int, c_int, _P, c_float, _P, _P, c_int, c_int64, c_int, _P]),
    "s2s_decode_pe": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, _P]),
    "s2s_decode_advance": (c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int, c_int, _P]),
    "s2s_duration_infer": (c_int, [_P, _P, c_int, c_float, c_float, c_int, _P]),
    "s2s_duration_loss": (c_int, [_P, _P, _P, c_int, c_int, c_float, c_float, c_float, _P, _P, _P, _P, c_int, _P]),
    "s2s_gelu_fwd": (c_int, [_P, _P, c_int64, _P]),
    "s2s_gelu_bwd": (c_int, [_P, _P, _P, c_int64, _P]),
    "s2s_outer_fwd": (c_int, [_P, _P, _P, _P, c_int64, c_int, _P]),
    "s2s_outer_bwd": (c_int, [_P, _P, _P, _P, _P, _P, c_int64, c_int, _P]),
    "s2s_dwconv_dilated_fwd": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P]),
    "s2s_dwconv_dilated_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P]),
    "s2s_rq_spline_fwd": (c_int, [_P, c_int64, _P, _P, _P, c_int64, _P, c_int, c_int, c_float, c_int, _P]),
    "s2s_rq_spline_bwd": (c_int, [_P, c_int64, _P, _P, _P, c_int64, _P, _P, c_int64, _P, c_int, c_int, c_float, _P]),
    "s2s_sdp_affine_fwd": (c_int, [_P, _P, _P, _P, _P, _P, c_float, c_int, c_int, c_int, _P]),
    "s2s_sdp_affine_bwd": (c_int, [_P, _P, _P, _P, _P, c_float, _P, _P, _P, c_int, c_int, _P]),
    "s2s_sdp_head_fwd": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, _P]),
    "s2s_sdp_head_bwd": (c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int, _P]),
    "s2s_sdp_gauss_fwd": (c_int, [_P, _P, _P, c_float, c_int, c_int, _P]),
    "s2s_sdp_gauss_bwd": (c_int, [_P, _P, _P, c_float, _P, c_int, c_int, c_int, _P]),
    "s2s_rowsum_acc": (c_int, [_P, _P, c_float, c_int, c_int, _P]),
    "s2s_rowbcast": (c_int, [_P, _P, c_float, c_int, c_int, _P]),
    "s2s_randn": (c_int, [_P, c_int64, c_uint64, _P, c_uint64, _P]),
    "s2s_sdp_durations": (c_int, [_P, _P, _P, c_float, c_int, c_int, _P]),
}

_lib = None


def load(path: str | None = None) -> ctypes.CDLL:
    """Load the shared library (once) and bind every signature.  Raises S2SError if absent."""
    global _lib
    if _lib is not None:
        return _lib
    path = path or os.environ.get("S2SVC_B200_LIB", LIB_PATH)
    if not os.path.exists(path):
        raise S2SError(f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "or seq2seq_vc_b200/csrc/build.sh -- there is no fallback path")
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    got = lib.s2s_abi_version()
    if got != ABI_VERSION:
        raise S2SError(f"libs2svc_b200.so ABI version {got} != expected {ABI_VERSION}: rebuild")
    _lib = lib
    return lib


@contextlib.contextmanager
def no_gc():
    """Collect cyclic garbage now and keep the collector off inside the block.  Every stream capture of this package runs under it:
    torch.cuda.graph() no longer calls gc.collect() itself (torch.compiler.config.force_cudagraph_gc), so a dead reference cycle that
    still owns CUDA graphs or device memory (an engine and its train step from an earlier batch of work) could be finalised by a
    collection that happens to trigger *inside* the capture -- its cudaFree / graph destruction invalidates a global-mode capture
    ("operation failed due to a previous error during capture", seen as an order-dependent test failure)."""
    gc.collect()
    was = gc.isenabled()
    gc.disable()
    try:
        yield
    finally:
        if was:
            gc.enable()


@contextlib.contextmanager
def graph_capture(graph, **kw):
    """`with torch.cuda.graph(graph, **kw)` under no_gc()."""
    with no_gc():
        with torch.cuda.graph(graph, **kw):
            yield


_DEBUG_CAPTURE = os.environ.get("S2S_DEBUG_CAPTURE", "0") == "1"
_last_ok = ""


def _capture_invalidated() -> bool:
    """Debug aid (S2S_DEBUG_CAPTURE=1): has the stream capture under way on torch's current stream been invalidated?"""
    from cuda.bindings import runtime as cudart

    err, status = cudart.cudaStreamIsCapturing(torch.cuda.current_stream().cuda_stream)
    return err != cudart.cudaError_t.cudaSuccess or status == cudart.cudaStreamCaptureStatus.cudaStreamCaptureStatusInvalidated


def check(rc: int, what: str = "") -> None:
    if _DEBUG_CAPTURE:
        global _last_ok
        if _capture_invalidated():
            raise S2SError(f"stream capture invalidated at or before {what} (rc={rc}); last call seen with a live capture: {_last_ok}")
        _last_ok = what
    if rc != 0:
        msg = load().s2s_last_error()
        raise S2SError(f"{what} failed (rc={rc}): {msg.decode() if msg else '?'}")


def dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return S2S_F32
    if t.dtype == torch.bfloat16:
        return S2S_BF16
    raise S2SError(f"unsupported dtype {t.dtype}")


def ptr(t) -> int | None:
    if t is None:
        return None
    if not t.is_cuda:
        raise S2SError("tensor is not on a CUDA device (no CPU fallback exists)")
    return t.data_ptr()


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def launch_count() -> int:
    return int(load().s2s_launch_count())


def device_check() -> None:
    check(load().s2s_device_check(), "s2s_device_check")


class Drop:
    """Python-side dropout descriptor (p, seed, per-site stream id, optional device seed)."""

    __slots__ = ("p", "seed", "site", "seed_dev")

    def __init__(self, p: float = 0.0, seed: int = 0, site: int = 0, seed_dev: torch.Tensor | None = None):
        self.p, self.seed, self.site, self.seed_dev = float(p), int(seed), int(site), seed_dev

    def c(self) -> DropoutDesc:
        return DropoutDesc(self.p, self.seed & 0xFFFFFFFFFFFFFFFF, self.site, ptr(self.seed_dev) if self.seed_dev is not None else None)

    @property
    def scale(self) -> float:
        return 1.0 / (1.0 - self.p) if 0.0 < self.p < 1.0 else 1.0


NO_DROP = Drop()
