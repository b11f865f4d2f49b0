"""Stochastic duration predictor at the C3 shape (B 64 x T_text 192 x C 384): one nll forward + backward, CUDA events.
Usage: python tools/micro_sdp.py   (under `ncu --metrics gpu__time_duration.sum` for the per-kernel list)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from seq2seq_vc_b200 import sdp as S


def main():
    hp = dict(channels=384, kernel_size=3, dds_conv_layers=3, flows=4)
    sd = {k: v.cuda().requires_grad_(True) for k, v in S.init_params(hp, "duration_predictor", 3).items()}
    B, T, C = 64, 192, 384
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, T, C, generator=g).cuda()
    tl = torch.full((B,), T, dtype=torch.int32, device="cuda")
    maskf = torch.ones(B * T, device="cuda")
    ds = torch.randint(1, 9, (B, T), generator=g).float().cuda()
    e_q = torch.randn(B, 2, T, generator=g).cuda()
    pred = S.StochasticDurationPredictor(hp, "duration_predictor", lambda n: sd[n], gemm_mode=2, dropout_rate=0.0)
    for it in range(3):
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        nll = pred.nll(x, tl, maskf, ds, e_q)
        e1.record()
        nll.sum().backward()
        e2.record()
        torch.cuda.synchronize()
        print(f"fwd {e0.elapsed_time(e1):.2f} ms  bwd {e1.elapsed_time(e2):.2f} ms (eager: includes host launch gaps)", flush=True)


if __name__ == "__main__":
    main()
