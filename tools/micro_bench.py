"""Kernel-level micro-benchmarks for the non-GEMM rows: STFT->log-mel (BASELINE configs[4]) and the
monotonic alignment search at the configs[2] shape.  CUDA events, inputs larger than L2 or L2 flushed."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from seq2seq_vc_b200 import api, ops


def timeit(fn, reps=10, flush=None):
    ts = []
    for i in range(reps + 3):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    out = {}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    # ---- C5: 256 clips x 10 s @ 48 kHz, n_fft 2048, hop 300, 80 mels
    B, ns = 256, 480000
    wav = (0.1 * torch.randn(B, ns, device="cuda")).clamp_(-1, 1)
    mel = torch.empty(B, 1 + ns // 300, 80, device="cuda")
    ms = timeit(lambda: api.logmel_batch(wav, 48000, fft_size=2048, hop_size=300, num_mels=80, out=mel), reps=5)
    nbytes = wav.numel() * 4 + mel.numel() * 4
    out["logmel_c5"] = {"ms": ms, "frames_per_s": mel.shape[0] * mel.shape[1] / (ms * 1e-3), "algorithmic_GB": nbytes / 1e9,
                        "achieved_GBps": nbytes / (ms * 1e-3) / 1e9, "frac_of_measured_hbm": nbytes / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"]}
    # CPU oracle on a bounded sample (one clip)
    from oracle import logmel_oracle
    w1 = wav[0].cpu().numpy()
    t0 = time.perf_counter()
    logmel_oracle.logmelfilterbank(w1, 48000, fft_size=2048, hop_size=300, num_mels=80)
    out["logmel_c5"]["cpu_oracle_frames_per_s_1core"] = (1 + ns // 300) / (time.perf_counter() - t0)
    # ---- MAS at the AAS-VC shape: 64 x (768 x 192)
    B, TF, TT = 64, 768, 192
    lp = torch.log_softmax(torch.randn(B, TF, TT, device="cuda"), -1)
    tl = torch.full((B,), TT, dtype=torch.int32, device="cuda")
    fl = torch.full((B,), TF, dtype=torch.int32, device="cuda")
    ms = timeit(lambda: ops.mas(lp, tl, fl, want_grad=False), reps=10, flush=flush)
    out["mas_c3"] = {"ms": ms, "utterances_per_s": B / (ms * 1e-3), "algorithmic_MB": lp.numel() * 4 / 1e6,
                     "achieved_GBps": lp.numel() * 4 / (ms * 1e-3) / 1e9, "serial_steps": TF}
    from oracle import mas_oracle
    lpc = lp.cpu().numpy()
    t0 = time.perf_counter()
    mas_oracle.mas_batch_c(lpc, [TT] * B, [TF] * B, threads=os.cpu_count())
    out["mas_c3"]["cpu_c_oracle_ms_all_cores"] = (time.perf_counter() - t0) * 1e3
    out["mas_c3"]["cpu_cores"] = os.cpu_count()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
