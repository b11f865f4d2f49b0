"""One eager training step of a bench workload bracketed by cudaProfilerStart/Stop, for
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file X python tools/profile_step.py c3
(launch list) or `ncu --set full -k regex:<kernel> -c 1 ...` (one full capture).  Not a benchmark: numbers under ncu are never bench values."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
    from seq2seq_vc_b200 import AASVC, AASVCTrainStep, FastSpeechVC, NARVCTrainStep, VTN, TransformerTTS, VTNTrainStep

    hp, B, T, L, bf16, desc = bench.WORKLOADS[wl]
    dev = torch.device("cuda", 0)
    xs, ilens, ys, labels, olens = bench.synthetic_batch(B, T, L, 1234, wl == "c4")
    xs, ys, labels = xs.to(dev), ys.to(dev), labels.to(dev)
    cd = "bf16" if bf16 else "float32"
    if wl == "c6":
        fixed = dict(positionwise_layer_type="linear", duration_predictor_use_encoder_outputs=False, encoder_normalize_before=True,
                     decoder_normalize_before=True, encoder_type="conformer", decoder_type="conformer", encoder_input_layer="conv2d",
                     transformer_enc_dropout_rate=0.2, transformer_enc_positional_dropout_rate=0.2, transformer_enc_attn_dropout_rate=0.2,
                     transformer_dec_dropout_rate=0.2, transformer_dec_positional_dropout_rate=0.2, transformer_dec_attn_dropout_rate=0.2)
        st = NARVCTrainStep(FastSpeechVC(**hp, **fixed, compute_dtype=cd, device=dev), use_graph=False)
        fx, fy, fd = (t.to(dev) for t in bench.fs_batch(B, T, L, 1234))
        step = lambda: st(fx, ilens, fy, olens, fd, fx)
    elif bench.is_aas(wl):
        fixed = dict(positionwise_layer_type="linear", duration_predictor_use_encoder_outputs=False, encoder_normalize_before=True,
                     decoder_normalize_before=True, duration_predictor_type="deterministic", encoder_input_layer="linear",
                     transformer_enc_dropout_rate=0.2, transformer_enc_positional_dropout_rate=0.2, transformer_enc_attn_dropout_rate=0.2,
                     transformer_dec_dropout_rate=0.2, transformer_dec_positional_dropout_rate=0.2, transformer_dec_attn_dropout_rate=0.2)
        st = AASVCTrainStep(AASVC(**hp, **fixed, compute_dtype=cd, device=dev), use_graph=False)
        st.steps = 1
        step = lambda: st(xs, ilens, ys, olens, xs)
    else:
        st = VTNTrainStep((TransformerTTS if wl == "c4" else VTN)(**hp, compute_dtype=cd, device=dev), use_graph=False)
        step = lambda: st(xs, ilens, ys, labels, olens)
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    step()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("profiled one step of", desc)


if __name__ == "__main__":
    main()
