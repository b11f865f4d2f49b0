"""C1 (BASELINE configs[0]) parity of the float32 engine modes against the CPU oracle: CUDA-core GEMM ("simt"), fp32-accurate
tcgen05 GEMM with a 6-term and a 3-term bf16 split, and the bf16 path.  Prints mel / attention L1 and the worst gradient errors."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import vtn_oracle
from seq2seq_vc_b200 import VTNEngine, ops

C1_HP = dict(idim=80, odim=80, adim=256, aheads=4, elayers=2, dlayers=2, eunits=1024, dunits=1024, decoder_reduction_factor=2,
             dprenet_dropout_rate=0.0)
NO_DROPOUT = dict(transformer_enc_dropout_rate=0.0, enc_positional_dropout_rate=0.0, dec_dropout_rate=0.0,
                  dec_positional_dropout_rate=0.0, postnet_dropout_rate=0.0)


def main():
    hp = vtn_oracle.default_hparams(**C1_HP)
    sd = vtn_oracle.init_state_dict(hp, seed=2)
    batch = vtn_oracle.synthetic_batch(4, 200, 400, ilens=[200, 180, 160, 120], olens=[400, 380, 300, 250], seed=1234)
    out, (l1, bce), grads = vtn_oracle.vtn_loss_and_grads(sd, hp, *batch)
    xs, ilens, ys, labels, olens = batch
    res = {}
    for label, kw, terms in (("fp32 simt", dict(bf16=False, fp32_gemm="simt"), 6), ("fp32 tc x6", dict(bf16=False), 6),
                             ("fp32 tc x3", dict(bf16=False), 3), ("bf16", dict(bf16=True), 6)):
        ops.SPLIT_TERMS = terms
        eng = VTNEngine(dict(C1_HP, **NO_DROPOUT), device="cuda:0", **kw)
        eng.load_state_dict(sd)
        for it in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            after, before, logits = eng.forward(xs.cuda(), ys.cuda(), ilens, olens)
            losses = eng.loss(ys.cuda(), labels.cuda())
            eng.backward(eng.d_after, eng.d_before, eng.d_logits)
            torch.cuda.synchronize()
            ms = (time.perf_counter() - t0) * 1e3
        r = {"fwd_bwd_ms_eager": ms,
             "after_L1": (after.float().cpu() - out["after_outs"].detach()).abs().mean().item(),
             "before_L1": (before.float().cpu() - out["before_outs"].detach()).abs().mean().item(),
             "attn_L1_max": max((eng.attn[n].float().cpu() - ref.detach()).abs().mean().item() for n, ref in out["attn"].items() if n in eng.attn),
             "l1_loss_err": abs(losses[0].item() - float(l1)), "bce_err": abs(losses[1].item() - float(bce))}
        worst = []
        for name, ref in grads.items():
            got = eng.store.g(name).cpu()
            e = (got - ref).abs()
            worst.append((e.max().item() / (ref.abs().max().item() + 1e-5), e.mean().item() / (ref.abs().mean().item() + 1e-12), name))
        worst.sort(reverse=True)
        r["grad_max_rel_worst3"] = [(round(a, 6), n) for a, _, n in worst[:3]]
        r["grad_mean_rel_worst"] = max(b for _, b, _ in worst)
        res[label] = r
        print(label, json.dumps(r), flush=True)
    ops.SPLIT_TERMS = 6
    if len(sys.argv) > 1:
        json.dump(res, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
