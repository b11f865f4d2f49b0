"""Kernel-level timing of the attention paths at the BASELINE shapes (CUDA events, L2 flushed between calls):
tcgen05 flash attention (s2s_attn_fwd_tc / s2s_attn_bwd_tc) next to the stored-P path it replaces
(s2s_attn_probs_fwd + PV GEMM; dV / dP / softmax' / dQ / dK GEMMs).  Usage: python tools/micro_attn.py [out.json]"""
import json
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from seq2seq_vc_b200 import ops


def timeit(fn, flush, reps=12):
    ts = []
    for i in range(reps + 3):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2] * 1e3       # us


def main():
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    dt = torch.bfloat16
    shapes = [("c2 decoder self-attention (causal)", 32, 8, 512, 512, 48, True, False),
              ("c2 decoder source attention (P emitted)", 32, 8, 512, 127, 48, False, True),
              ("c2 decoder source attention", 32, 8, 512, 127, 48, False, False),
              ("c2 encoder self-attention", 32, 8, 127, 127, 48, False, False),
              ("c2b64 decoder self-attention (causal)", 64, 8, 512, 512, 48, True, False),
              ("c4 decoder self-attention (causal)", 64, 4, 500, 500, 96, True, False),
              ("c4 decoder source attention (P emitted)", 64, 4, 500, 161, 96, False, True),
              ("c1-like d_k 64 (causal)", 4, 4, 200, 200, 64, True, False)]
    out = []
    for name, B, H, T1, T2, dk, causal, emit in shapes:
        qkv = torch.randn(B, max(T1, T2), 3, H, dk, device="cuda", dtype=dt)
        q, k, v = qkv[:, :T1, 0], qkv[:, :T2, 1], qkv[:, :T2, 2]
        klens = torch.full((B,), T2, dtype=torch.int32, device="cuda")
        scale = 1.0 / math.sqrt(dk)
        ld = (T2 + 7) // 8 * 8
        ctx = torch.empty(B, T1, H, dk, device="cuda", dtype=dt)
        lse = torch.empty(ops.attn_lse_shape(B, H, T1), device="cuda")
        dvec = torch.empty_like(lse)
        P = torch.empty(B, H, T1, ld, device="cuda", dtype=dt)
        dP = torch.empty_like(P)
        dctx = torch.randn(B, T1, H, dk, device="cuda", dtype=dt)
        dqkv = torch.empty_like(qkv)
        dq, dk_, dv = dqkv[:, :T1, 0], dqkv[:, :T2, 1], dqkv[:, :T2, 2]
        r = {"shape": name, "B": B, "H": H, "T1": T1, "T2": T2, "dk": dk, "causal": causal}
        r["flash_fwd_us"] = timeit(lambda: ops.attn_fwd_tc(q, k, v, ctx, lse, klens, causal, scale, P if emit else None), flush)
        r["flash_bwd_us"] = timeit(lambda: ops.attn_bwd_tc(q, k, v, ctx, dctx, lse, dvec, dq, dk_, dv, klens, causal, scale), flush)

        def old_fwd():
            ops.attn_probs_fwd(q, k, P, klens, causal, T2, scale)
            ops.gemm(P[..., :T2], v.permute(0, 2, 3, 1), ctx.permute(0, 2, 1, 3), mode=1)

        def old_bwd():
            d4 = dctx.permute(0, 2, 1, 3)
            ops.gemm(P[..., :T2].transpose(-1, -2), d4.transpose(-1, -2), dv.permute(0, 2, 1, 3), mode=1)
            ops.gemm(d4, v.permute(0, 2, 1, 3), dP[..., :T2], mode=1)
            ops.softmax_bwd(P, dP, T2, scale)
            ops.gemm(dP[..., :T2], k.permute(0, 2, 3, 1), dq.permute(0, 2, 1, 3), mode=1)
            ops.gemm(dP[..., :T2].transpose(-1, -2), q.permute(0, 2, 3, 1), dk_.permute(0, 2, 1, 3), mode=1)

        r["stored_P_fwd_us"] = timeit(old_fwd, flush)
        r["stored_P_bwd_us"] = timeit(old_bwd, flush)
        f = (0.5 if causal else 1.0) * 4.0 * B * H * T1 * T2 * dk
        r["flash_fwd_tflops"] = f / r["flash_fwd_us"] / 1e6
        r["flash_bwd_tflops"] = 2.5 * f / r["flash_bwd_us"] / 1e6
        out.append(r)
        print(json.dumps(r), flush=True)
    if len(sys.argv) > 1:
        json.dump(out, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
