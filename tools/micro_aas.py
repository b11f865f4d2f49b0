"""Micro-benchmarks of the HBM-bound AAS-VC kernels at the BASELINE configs[2] decoder shape
(B 64 x T 768 x C 1536, bf16): CUDA events, L2 flushed between repetitions, algorithmic bytes / measured HBM peak.
Usage: python tools/micro_aas.py [out.json]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from seq2seq_vc_b200 import ops
from seq2seq_vc_b200._lib import Drop


def timeit(fn, flush, reps=8):
    ts = []
    for i in range(reps + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    try:
        peak = json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    B, T, C, K, H = 64, 768, 1536, 15, 2
    bf = torch.bfloat16
    dev = "cuda"
    e = 2
    N = B * T * C
    x = torch.randn(B, T, C, device=dev).to(bf)
    y = torch.empty_like(x)
    g = torch.randn(B, T, C, device=dev).to(bf)
    w = torch.randn(C, K, device=dev) * 0.2
    bias = torch.randn(C, device=dev)
    dw, db = torch.zeros(C, K, device=dev), torch.zeros(C, device=dev)
    x2 = torch.randn(B * T, 2 * C, device=dev).to(bf)
    dx2 = torch.empty_like(x2)
    gam, bet = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    mean, rstd = torch.empty(B * T, device=dev), torch.empty(B * T, device=dev)
    sums = torch.zeros(2 * C, device=dev)
    bmean, binv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
    ld = T
    S = torch.randn(B, H, T, ld, device=dev).to(bf)
    dS = torch.randn(B, H, T, ld, device=dev).to(bf)
    BD = torch.randn(H, B, T, 2 * T, device=dev).to(bf)
    klens = torch.full((B,), T, dtype=torch.int32, device=dev)
    qkv = torch.randn(B * T, 3 * C, device=dev).to(bf)
    qu, qv = torch.empty(B * T, C, device=dev, dtype=bf), torch.empty(B * T, C, device=dev, dtype=bf)
    u, v = torch.randn(C, device=dev), torch.randn(C, device=dev)
    TT = T // 4
    feats, text = torch.randn(B, T, C, device=dev).to(bf), torch.randn(B, TT, C, device=dev).to(bf)
    tl = torch.full((B,), TT, dtype=torch.int32, device=dev)
    fl = torch.full((B,), T, dtype=torch.int32, device=dev)
    logp, lse = torch.empty(B, T, TT, device=dev), torch.empty(B, T, device=dev)
    prior = torch.zeros(B, T, TT, device=dev)
    loss, dlogp, aws = torch.zeros(1, device=dev), torch.empty(B, T, TT, device=dev), torch.empty(B, T, TT, device=dev)
    cases = {
        "dwconv_fwd": (lambda: ops.dwconv_fwd(x, w, bias, y), 2 * N * e),
        "dwconv_dx": (lambda: ops.dwconv_bwd(g, x, w, y, None), 2 * N * e),
        "dwconv_dw+dbias": (lambda: ops.dwconv_bwd(g, x, w, None, dw, db), 2 * N * e),
        "glu_fwd": (lambda: ops.glu_fwd(x2, y), 3 * N * e),
        "glu_bwd": (lambda: ops.glu_bwd(g, x2, dx2), 5 * N * e),
        "swish_fwd": (lambda: ops.swish_fwd(x, y), 2 * N * e),
        "swish_bwd": (lambda: ops.swish_bwd(g, x, y), 3 * N * e),
        "layernorm_fwd": (lambda: ops.layernorm_fwd(x, gam, bet, y, mean, rstd), 2 * N * e),
        "layernorm_bwd(dx+dgamma/dbeta)": (lambda: ops.layernorm_bwd(g, x, gam, mean, rstd, y, sums[:C], sums[C:]), 3 * N * e),
        "layernorm_bwd(dx only)": (lambda: ops.layernorm_bwd(g, x, gam, mean, rstd, y, None, None), 3 * N * e),
        "layernorm_bwd(dgamma/dbeta only)": (lambda: ops.layernorm_bwd(g, x, gam, mean, rstd, None, sums[:C], sums[C:]), 2 * N * e),
        "swish_fwd(p=0.2)": (lambda: ops.swish_fwd(x, y, Drop(0.2, 1, 3)), 2 * N * e),
        "scale_dropout(p=0.2)": (lambda: ops.scale_dropout(x, y, 2.0, Drop(0.2, 1, 4)), 2 * N * e),
        "dropout_bwd(p=0.2)": (lambda: ops.dropout_bwd(g, y, Drop(0.2, 1, 5)), 2 * N * e),
        "softmax_fwd(P and dropped copy, p=0.2)": (lambda: ops.softmax_fwd(S, klens, False, T, dS, Drop(0.2, 1, 6)), 3 * S.numel() * e),
        "colsum": (lambda: ops.colsum(x.view(B * T, C), db), N * e),
        "bn_stats": (lambda: ops.bn_stats(x, sums, T, 0), N * e),
        "bn_apply_swish": (lambda: ops.bn_apply(x, bmean, binv, gam, bet, y, T, 0, 2), 2 * N * e),
        "bn_bwd_reduce": (lambda: ops.bn_bwd_reduce(g, y, x, bmean, binv, gam, bet, sums, T, 0, 2), 2 * N * e),
        "bn_bwd_apply": (lambda: ops.bn_bwd_apply(g, y, x, bmean, binv, gam, bet, sums, y, sums[:C], sums[C:], T, 0, 2), 3 * N * e),
        "scale_dropout": (lambda: ops.scale_dropout(x, y, 2.0), 2 * N * e),
        "bias_add2": (lambda: ops.bias_add2(qkv[:, :C], u, v, qu, qv), 3 * N * e),
        "relshift_add": (lambda: ops.relshift_add(S, BD, T), (2 * S.numel() + S.numel()) * e),
        "relshift_bwd": (lambda: ops.relshift_bwd(dS, BD, T), (S.numel() + BD.numel()) * e),
        "softmax_fwd": (lambda: ops.softmax_fwd(S, klens, False, T), 2 * S.numel() * e),
        "softmax_bwd": (lambda: ops.softmax_bwd(S, dS, T, 0.1), 3 * S.numel() * e),
        "align_logp_fwd(pairdist+logsoftmax)": (lambda: ops.align_logp_fwd(feats, text, tl, logp, lse), (feats.numel() + text.numel()) * e + logp.numel() * 4),
        "forward_sum(loss+grad)": (lambda: ops.forward_sum(logp, prior, tl, fl, aws, loss, dlogp), logp.numel() * 4 * 3),
    }
    out = {}
    for name, (fn, nbytes) in cases.items():
        ms = timeit(fn, flush)
        out[name] = {"us": round(ms * 1e3, 1), "algorithmic_MB": round(nbytes / 1e6, 1), "GBps": round(nbytes / (ms * 1e-3) / 1e9, 1),
                     "frac_hbm_peak": round(nbytes / (ms * 1e-3) / 1e9 / peak, 3)}
        print(f"{name:40s} {ms * 1e3:9.1f} us  {nbytes / 1e6:8.1f} MB  {nbytes / (ms * 1e-3) / 1e9:8.1f} GB/s  {out[name]['frac_hbm_peak']:.3f} of {peak:.0f}")
    if len(sys.argv) > 1:
        json.dump({"shape": "B64 x T768 x C1536 bf16, K15, H2", "hbm_peak_GBps": peak, "kernels": out}, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
