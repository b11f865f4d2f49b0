"""Times the STFT -> log-mel kernel on the reference recipes' settings next to BASELINE configs[4] (CUDA events, L2-exceeding input)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from seq2seq_vc_b200 import api

out = {}
for name, sr, n_fft, hop, secs in (("recipe 16 kHz n_fft 1024 hop 256", 16000, 1024, 256, 10), ("24 kHz n_fft 2048 hop 300 win 1200", 24000, 2048, 300, 10),
                                   ("c5 48 kHz n_fft 2048 hop 300", 48000, 2048, 300, 10), ("16 kHz n_fft 512 hop 128", 16000, 512, 128, 10)):
    B = 256
    wav = torch.randn(B, sr * secs, device="cuda") * 0.1
    mel = torch.empty(B, 1 + wav.shape[1] // hop, 80, device="cuda")
    kw = dict(fft_size=n_fft, hop_size=hop, num_mels=80, fmin=80 if sr <= 24000 else None, fmax=7600 if sr <= 24000 else None)
    for _ in range(3):
        api.logmel_batch(wav, sr, out=mel, **kw)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        api.logmel_batch(wav, sr, out=mel, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    frames = mel.shape[0] * mel.shape[1]
    out[name] = {"ms": ms, "frames": frames, "frames_per_s": frames / (ms * 1e-3), "ns_per_frame": ms * 1e6 / frames}
print(json.dumps(out, indent=1))
