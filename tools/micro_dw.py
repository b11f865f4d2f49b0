"""Weight-gradient GEMMs of one C2 decoder / encoder layer: one grouped tcgen05 launch vs one launch per product
(CUDA events, L2 flushed).  Usage: python tools/micro_dw.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from seq2seq_vc_b200 import ops


def timeit(fn, flush, reps=10):
    ts = []
    for i in range(reps + 3):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    bf = torch.bfloat16
    sets = {"c2 decoder layer": [(384, 1536, 16384), (1536, 384, 16384), (384, 384, 16384), (384, 384, 16384), (768, 384, 4064), (384, 384, 16384),
                                 (1152, 384, 16384)],
            "c2 encoder layer": [(384, 1536, 4064), (1536, 384, 4064), (384, 384, 4064), (1152, 384, 4064)]}
    for name, shapes in sets.items():
        probs = []
        fl = 0.0
        for (M, N, K) in shapes:
            dy, x = torch.randn(K, M, device="cuda", dtype=bf), torch.randn(K, N, device="cuda", dtype=bf)
            probs.append((dy.t(), x.t(), torch.zeros(M, N, device="cuda"), dict(accumulate=True)))
            fl += 2.0 * M * N * K
        tg = timeit(lambda: ops.gemm_grouped(probs, mode=1), flush)
        ts = timeit(lambda: [ops.gemm(a, b, c, mode=1, **kw) for a, b, c, kw in probs], flush)
        each = [timeit(lambda p=p: ops.gemm(p[0], p[1], p[2], mode=1, **p[3]), flush) for p in probs]
        print(f"{name}: grouped {tg:7.1f} us ({fl / tg / 1e6:6.1f} TFLOP/s) | one by one {ts:7.1f} us ({fl / ts / 1e6:6.1f} TFLOP/s) | each "
              + " ".join(f"{e:.1f}" for e in each), flush=True)


if __name__ == "__main__":
    main()
