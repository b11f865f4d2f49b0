"""Inference-path timing: autoregressive VTN.inference and non-autoregressive AASVC.inference on one utterance,
random-init weights of the BASELINE model shapes.  Prints generated mel frames / second (the number the reference logs as
"inference speed = %.1f frames / sec", trainers/ar_vc.py:185-192).  Usage: python tools/bench_inference.py [out.json]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from seq2seq_vc_b200 import AASVC, VTN


def main():
    dev = torch.device("cuda", 0)
    out = {}
    hp = bench.WORKLOADS["c2"][0]
    model = VTN(**hp, dprenet_dropout_rate=0.5, compute_dtype="bf16", device=dev)
    model.eval()
    x = torch.randn(512, 80, device=dev)
    args = dict(threshold=2.0, minlenratio=0.0, maxlenratio=2.0)          # never stops early: 127 decoder steps = 254 frames
    model.inference(x, args)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    outs, probs, att = model.inference(x, args)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out["vtn_base_inference"] = {"input_frames": 512, "decoder_steps": att.shape[2], "frames": outs.shape[0], "seconds": dt,
                                 "frames_per_s": outs.shape[0] / dt, "ms_per_step": dt * 1e3 / att.shape[2]}
    hp3 = bench.WORKLOADS["c3"][0]
    fixed = dict(positionwise_layer_type="linear", duration_predictor_use_encoder_outputs=False, encoder_normalize_before=True,
                 decoder_normalize_before=True, duration_predictor_type="deterministic", encoder_input_layer="linear")
    m3 = AASVC(**hp3, **fixed, compute_dtype="bf16", device=dev)
    with torch.no_grad():
        m3.duration_predictor.linear.bias.add_(1.4)                       # random init predicts ~0 frames per token; aim at ~4
    m3.eval()
    x3 = torch.randn(768, 80, device=dev)
    m3.inference(x3, dp_input=x3)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    o3, d3 = m3.inference(x3, dp_input=x3)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out["aasvc_inference"] = {"input_frames": 768, "frames": o3.shape[0], "seconds": dt, "frames_per_s": o3.shape[0] / dt}
    print(json.dumps(out, indent=1))
    if len(sys.argv) > 1:
        json.dump(out, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
