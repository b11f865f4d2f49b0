"""CUDA-event timing of s2s_feat_stats on the C5 log-mel output shape (256 clips x 1601 frames x 80 bins, fp32)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from seq2seq_vc_b200 import ops  # noqa: E402

B, T, D = 256, 1601, 80
x = torch.randn(B, T, D, device="cuda")
lens = torch.full((B,), T, dtype=torch.int32, device="cuda")
acc = torch.zeros(2 * D + 1, dtype=torch.float64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    ops.feat_stats(x, lens, acc)
ts = []
for _ in range(10):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.feat_stats(x, lens, acc)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = sorted(ts)[len(ts) // 2]
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))).get("hbm_gbs", 6437.6)
gbs = x.numel() * 4 / (ms * 1e-3) / 1e9
print(json.dumps({"kernel": "feat_stats_kernel", "shape": [B, T, D], "ms": ms, "GB/s": gbs, "frac_of_hbm_peak": gbs / peak, "l2": "flushed (256 MB write) before every launch"}))
