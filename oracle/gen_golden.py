"""Generate tests/golden/*.npz from the LIVE reference (run in the build container only).

TEST INFRASTRUCTURE ONLY.  Usage:  python -m oracle.gen_golden
Imports /root/reference through oracle/ref_shim.py (nothing is copied) and dumps small seeded
input/output vectors so the GPU box -- which has no /root/reference -- can pin the oracle and the
CUDA path against the reference's actual outputs.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from oracle import aasvc_oracle, ref_shim, vtn_oracle

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

TINY_HP = dict(idim=80, odim=80, dprenet_layers=2, dprenet_units=16, adim=32, aheads=2, elayers=1, eunits=48,
               dlayers=2, dunits=48, postnet_layers=3, postnet_filts=5, postnet_chans=16,
               decoder_reduction_factor=2)


def gen_vtn_tiny():
    from seq2seq_vc.losses import Seq2SeqLoss
    from seq2seq_vc.models import VTN

    torch.manual_seed(7)
    model = VTN(dprenet_dropout_rate=0.0, **TINY_HP)
    ref_shim.disable_dropout(model)
    # de-trivialise LayerNorm / BatchNorm affine params and PE alphas so they are exercised
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith("alpha"):
                p.fill_(0.7 if "encoder" in n else 1.3)
            elif p.dim() == 1 and ("norm" in n or ".1." in n):
                p.add_(0.1 * torch.randn_like(p))
    model.train()
    sd0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    xs, ilens, ys, labels, olens = vtn_oracle.synthetic_batch(3, 48, 37, ilens=[48, 41, 30], olens=[37, 30, 21], seed=11)
    out = model(xs, torch.tensor(ilens), ys, labels, torch.tensor(olens))
    l1, bce = Seq2SeqLoss()(*out[:6])
    (l1 + bce).backward()
    dump = {"sd." + k: v.numpy() for k, v in sd0.items()}
    dump.update({"grad." + k: p.grad.numpy() for k, p in model.named_parameters()})
    dump.update({"bn_after." + k: v.numpy() for k, v in model.state_dict().items() if "running_" in k})
    dump.update(xs=xs.numpy(), ilens=np.array(ilens), ys=ys.numpy(), labels=labels.numpy(), olens=np.array(olens),
                after_outs=out[0].detach().numpy(), before_outs=out[1].detach().numpy(),
                logits=out[2].detach().numpy(), ys_out=out[3].numpy(), labels_out=out[4].numpy(),
                olens_out=out[5].numpy(), ilens_ds_st=out[6][1].numpy(), olens_in=out[6][2].numpy(),
                l1_loss=l1.detach().numpy(), bce_loss=bce.detach().numpy())
    for i, a in enumerate(out[6][0]):
        dump[f"att_ws.{i}"] = a.detach().numpy()
    # eval-mode (running-stat BatchNorm) forward on the *updated* running stats
    model.eval()
    with torch.no_grad():
        oute = model(xs, torch.tensor(ilens), ys, labels, torch.tensor(olens))
    dump["eval_after_outs"] = oute[0].numpy()
    # autoregressive inference (models/vtn.py:302-394) on the first utterance; runs to maxlen (threshold never reached first)
    with torch.no_grad():
        io, ip, ia = model.inference(xs[0, :ilens[0]], dict(threshold=0.9999, minlenratio=0.0, maxlenratio=1.6))
    dump.update(inf_outs=io.numpy(), inf_probs=ip.numpy(), inf_att_ws=ia.numpy())
    np.savez_compressed(os.path.join(GOLDEN, "vtn_tiny.npz"), **dump)
    print("vtn_tiny:", len(dump), "arrays")


def gen_vtn_rfactor():
    """decoder_reduction_factor 1 and 4 (the recipe's value, egs/arctic/vc1/conf/vtn.v1.yaml:43) with ragged target lengths that
    are NOT multiples of r: the trimming / thinning / label fix-up paths of models/vtn.py:227-243,262-274."""
    from seq2seq_vc.losses import Seq2SeqLoss
    from seq2seq_vc.models import VTN

    for r, olens in ((1, [37, 30, 21]), (4, [37, 30, 21]), (3, [35, 31, 22])):
        torch.manual_seed(70 + r)
        hp = dict(TINY_HP, decoder_reduction_factor=r)
        model = VTN(dprenet_dropout_rate=0.0, **hp)
        ref_shim.disable_dropout(model)
        with torch.no_grad():
            for n, p in model.named_parameters():
                if n.endswith("alpha"):
                    p.fill_(0.7 if "encoder" in n else 1.3)
                elif p.dim() == 1 and ("norm" in n or ".1." in n):
                    p.add_(0.1 * torch.randn_like(p))
        model.train()
        sd0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
        ilens = [48, 41, 30]
        xs, ilens, ys, labels, olens = vtn_oracle.synthetic_batch(3, 48, max(olens), ilens=ilens, olens=olens, seed=11 + r)
        out = model(xs, torch.tensor(ilens), ys, labels, torch.tensor(olens))
        l1, bce = Seq2SeqLoss()(*out[:6])
        (l1 + bce).backward()
        dump = {"sd." + k: v.numpy() for k, v in sd0.items()}
        dump.update({"grad." + k: p.grad.numpy() for k, p in model.named_parameters()})
        dump.update(xs=xs.numpy(), ilens=np.array(ilens), ys=ys.numpy(), labels=labels.numpy(), olens=np.array(olens), r=np.array(r),
                    after_outs=out[0].detach().numpy(), before_outs=out[1].detach().numpy(), logits=out[2].detach().numpy(),
                    ys_out=out[3].numpy(), labels_out=out[4].numpy(), olens_out=out[5].numpy(), ilens_ds_st=out[6][1].numpy(),
                    olens_in=out[6][2].numpy(), l1_loss=l1.detach().numpy(), bce_loss=bce.detach().numpy())
        for i, a in enumerate(out[6][0]):
            dump[f"att_ws.{i}"] = a.detach().numpy()
        with torch.no_grad():
            model.eval()
            io, ip, ia = model.inference(xs[0, :ilens[0]], dict(threshold=0.9999, minlenratio=0.0, maxlenratio=1.2))
        dump.update(inf_outs=io.numpy(), inf_probs=ip.numpy(), inf_att_ws=ia.numpy())
        np.savez_compressed(os.path.join(GOLDEN, f"vtn_r{r}_tiny.npz"), **dump)
        print(f"vtn_r{r}_tiny:", len(dump), "arrays", "out len", out[0].shape[1], "olens_out", out[5].tolist())


def gen_vtn_convffn_tiny():
    """VTN with the Transformer encoder's position-wise layer replaced by MultiLayeredConv1d / Conv1dLinear, kernel size 3
    (models/vtn.py:116-117, modules/transformer/encoder.py:143-175, multi_layer_conv.py:12-108)."""
    from seq2seq_vc.losses import Seq2SeqLoss
    from seq2seq_vc.models import VTN

    for tag, layer, seed in (("conv1d", "conv1d", 101), ("conv1d_linear", "conv1d-linear", 103)):
        torch.manual_seed(seed)
        model = VTN(dprenet_dropout_rate=0.0, positionwise_layer_type=layer, positionwise_conv_kernel_size=3, **dict(TINY_HP, elayers=2))
        ref_shim.disable_dropout(model)
        with torch.no_grad():
            for n, p in model.named_parameters():
                if n.endswith("alpha"):
                    p.fill_(0.7 if "encoder" in n else 1.3)
                elif p.dim() == 1 and ("norm" in n or ".1." in n):
                    p.add_(0.1 * torch.randn_like(p))
        model.train()
        sd0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
        xs, ilens, ys, labels, olens = vtn_oracle.synthetic_batch(3, 52, 37, ilens=[52, 45, 31], olens=[37, 30, 21], seed=seed)
        out = model(xs, torch.tensor(ilens), ys, labels, torch.tensor(olens))
        l1, bce = Seq2SeqLoss()(*out[:6])
        (l1 + bce).backward()
        dump = {"sd." + k: v.numpy() for k, v in sd0.items()}
        dump.update({"grad." + k: p.grad.numpy() for k, p in model.named_parameters()})
        dump.update(xs=xs.numpy(), ilens=np.array(ilens), ys=ys.numpy(), labels=labels.numpy(), olens=np.array(olens),
                    after_outs=out[0].detach().numpy(), before_outs=out[1].detach().numpy(), logits=out[2].detach().numpy(),
                    ys_out=out[3].numpy(), labels_out=out[4].numpy(), olens_out=out[5].numpy(), ilens_ds_st=out[6][1].numpy(),
                    olens_in=out[6][2].numpy(), l1_loss=l1.detach().numpy(), bce_loss=bce.detach().numpy())
        for i, a in enumerate(out[6][0]):
            dump[f"att_ws.{i}"] = a.detach().numpy()
        np.savez_compressed(os.path.join(GOLDEN, f"vtn_{tag}_k3_tiny.npz"), **dump)
        print(f"vtn_{tag}_k3_tiny:", len(dump), "arrays")


def gen_vtn_conformer_tiny():
    """VTN(encoder_type="conformer") (models/vtn.py:83-143): Conv2dSubsampling + LegacyRelPositionalEncoding, macaron conformer
    blocks with LegacyRelPositionMultiHeadedAttention and the convolution module (the class defaults), Transformer decoder;
    and the same model with conformer_rel_pos_type="latest"."""
    from seq2seq_vc.losses import Seq2SeqLoss
    from seq2seq_vc.models import VTN

    for tag, rel, seed in (("legacy", "legacy", 91), ("latest", "latest", 93)):
        torch.manual_seed(seed)
        model = VTN(dprenet_dropout_rate=0.0, encoder_type="conformer", conformer_rel_pos_type=rel, conformer_enc_kernel_size=7, **dict(TINY_HP, elayers=2))
        ref_shim.disable_dropout(model)
        with torch.no_grad():
            for n, p in model.named_parameters():
                if n.endswith("alpha"):
                    p.fill_(1.3)
                elif p.dim() == 1 and ("norm" in n or ".1." in n):
                    p.add_(0.1 * torch.randn_like(p))
        model.train()
        sd0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
        xs, ilens, ys, labels, olens = vtn_oracle.synthetic_batch(3, 52, 37, ilens=[52, 45, 31], olens=[37, 30, 21], seed=seed)
        out = model(xs, torch.tensor(ilens), ys, labels, torch.tensor(olens))
        l1, bce = Seq2SeqLoss()(*out[:6])
        (l1 + bce).backward()
        dump = {"sd." + k: v.numpy() for k, v in sd0.items()}
        dump.update({"grad." + k: p.grad.numpy() for k, p in model.named_parameters()})
        dump.update({"bn_after." + k: v.numpy() for k, v in model.state_dict().items() if "running_" in k})
        dump.update(xs=xs.numpy(), ilens=np.array(ilens), ys=ys.numpy(), labels=labels.numpy(), olens=np.array(olens),
                    after_outs=out[0].detach().numpy(), before_outs=out[1].detach().numpy(), logits=out[2].detach().numpy(),
                    ys_out=out[3].numpy(), labels_out=out[4].numpy(), olens_out=out[5].numpy(), ilens_ds_st=out[6][1].numpy(),
                    olens_in=out[6][2].numpy(), l1_loss=l1.detach().numpy(), bce_loss=bce.detach().numpy())
        for i, a in enumerate(out[6][0]):
            dump[f"att_ws.{i}"] = a.detach().numpy()
        for n, m in model.encoder.named_modules():
            if hasattr(m, "attn") and isinstance(getattr(m, "attn"), torch.Tensor):
                dump["attn.encoder." + n] = m.attn.detach().numpy()
        model.eval()
        with torch.no_grad():
            oute = model(xs, torch.tensor(ilens), ys, labels, torch.tensor(olens))
        dump["eval_after_outs"] = oute[0].numpy()
        with torch.no_grad():               # autoregressive inference on the updated running statistics (models/vtn.py:302-394)
            io, ip, ia = model.inference(xs[0, :ilens[0]], dict(threshold=0.9999, minlenratio=0.0, maxlenratio=1.3))
        dump.update(inf_outs=io.numpy(), inf_probs=ip.numpy(), inf_att_ws=ia.numpy())
        np.savez_compressed(os.path.join(GOLDEN, f"vtn_conformer_{tag}_tiny.npz"), **dump)
        print(f"vtn_conformer_{tag}_tiny:", len(dump), "arrays")


TTS_HP = dict(idim=40, odim=80, dprenet_layers=2, dprenet_units=16, adim=32, aheads=4, elayers=1, eunits=48,
              dlayers=2, dunits=48, postnet_layers=2, postnet_filts=5, postnet_chans=16, decoder_reduction_factor=2)


def gen_tts_tiny():
    """TransformerTTS (models/transformer_tts.py) + Seq2SeqLoss + GuidedMultiHeadAttentionLoss on token inputs."""
    from seq2seq_vc.losses import GuidedMultiHeadAttentionLoss, Seq2SeqLoss
    from seq2seq_vc.models.transformer_tts import TransformerTTS

    torch.manual_seed(11)
    model = TransformerTTS(dprenet_dropout_rate=0.0, use_guided_attn_loss=True, num_heads_applied_guided_attn=2,
                           num_layers_applied_guided_attn=2, **TTS_HP)
    ref_shim.disable_dropout(model)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith("alpha"):
                p.fill_(0.8 if "encoder" in n else 1.2)
    model.train()
    sd0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(21)
    ilens, olens = [13, 9, 5], [30, 23, 12]
    tokens = torch.randint(1, TTS_HP["idim"] - 1, (3, 13), generator=g)
    ys = torch.randn(3, 30, 80, generator=g)
    labels = torch.zeros(3, 30)
    for b in range(3):
        tokens[b, ilens[b]:] = 0
        ys[b, olens[b]:] = 0
        labels[b, olens[b] - 1:] = 1.0
    out = model(tokens, torch.tensor(ilens), ys, labels, torch.tensor(olens))
    l1, bce = Seq2SeqLoss()(*out[:6])
    ga = GuidedMultiHeadAttentionLoss(sigma=0.4, alpha=1.0)(out[6][0], out[6][1], out[6][2])
    (l1 + bce + ga).backward()
    dump = {"sd." + k: v.numpy() for k, v in sd0.items()}
    dump.update({"grad." + k: p.grad.numpy() for k, p in model.named_parameters()})
    dump.update(tokens=tokens.numpy(), ilens=np.array(ilens), ys=ys.numpy(), labels=labels.numpy(), olens=np.array(olens),
                after_outs=out[0].detach().numpy(), before_outs=out[1].detach().numpy(), logits=out[2].detach().numpy(),
                labels_out=out[4].numpy(), olens_out=out[5].numpy(), att_ws=out[6][0].detach().numpy(),
                ilens_out=out[6][1].numpy(), olens_in=out[6][2].numpy(), l1_loss=l1.detach().numpy(),
                bce_loss=bce.detach().numpy(), ga_loss=ga.detach().numpy())
    model.eval()
    with torch.no_grad():
        io, ip, ia = model.inference(tokens[0, :ilens[0]], dict(threshold=0.9999, minlenratio=0.0, maxlenratio=1.5))
    dump.update(inf_outs=io.numpy(), inf_probs=ip.numpy(), inf_att_ws=ia.numpy())
    dump.update({"bn_after." + k: v.numpy() for k, v in model.state_dict().items() if "running_" in k})
    np.savez_compressed(os.path.join(GOLDEN, "tts_tiny.npz"), **dump)
    print("tts_tiny:", len(dump), "arrays")


AAS_HP = dict(idim=80, odim=80, adim=32, aheads=2, elayers=1, eunits=48, dlayers=1, dunits=48,
              duration_predictor_input_dim=80, duration_predictor_layers=2, duration_predictor_chans=16,
              duration_predictor_kernel_size=3, postnet_layers=2, postnet_filts=5, postnet_chans=16,
              post_encoder_reduction_factor=4, conformer_enc_kernel_size=7, conformer_dec_kernel_size=7)
AAS_FIXED = dict(positionwise_layer_type="linear", positionwise_conv_kernel_size=1, duration_predictor_use_encoder_outputs=False,
                 encoder_normalize_before=True, decoder_normalize_before=True, duration_predictor_type="deterministic",
                 encoder_input_layer="linear")


def gen_aasvc_tiny():
    """AASVC (models/aas_vc.py) + L1Loss + ForwardSumLoss + bin loss + DurationPredictorLoss, assembled as
    AASVCTrainer._train_step does (trainers/aas_vc.py:73-134, lambda_align = 2.0)."""
    from seq2seq_vc.losses import DurationPredictorLoss, ForwardSumLoss, L1Loss
    from seq2seq_vc.models import AASVC

    torch.manual_seed(13)
    model = AASVC(**AAS_HP, **AAS_FIXED)
    ref_shim.disable_dropout(model)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if p.dim() == 1 and ("norm" in n or "embed.1" in n or ".2." in n or ".1." in n):
                p.add_(0.1 * torch.randn_like(p))
    model.train()
    sd0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    xs, ilens, ys, olens, dpi = aasvc_oracle.synthetic_batch(3, 50, 44, ilens=[50, 44, 37], olens=[44, 40, 30], seed=17)
    ret = model(xs, torch.tensor(ilens), ys, torch.tensor(olens), dpi, dp_lengths=torch.tensor(ilens))
    l1 = L1Loss()(ret["after_outs"], ret["before_outs"], ret["ys"], ret["olens"])
    fs = ForwardSumLoss()(ret["log_p_attn"], ret["ilens"], ret["olens_reduced"])
    dur = DurationPredictorLoss()(ret["d_outs"], ret["ds"], ret["ilens"])
    (l1 + 2.0 * (fs + ret["bin_loss"]) + dur).backward()
    dump = {"sd." + k: v.numpy() for k, v in sd0.items()}
    dump.update({"grad." + k: p.grad.numpy() for k, p in model.named_parameters() if p.grad is not None})
    dump.update({"bn_after." + k: v.numpy() for k, v in model.state_dict().items() if "running_" in k})
    dump.update(xs=xs.numpy(), ilens=np.array(ilens), ys=ys.numpy(), olens=np.array(olens), dp_inputs=dpi.numpy(),
                after_outs=ret["after_outs"].detach().numpy(), before_outs=ret["before_outs"].detach().numpy(),
                log_p_attn=ret["log_p_attn"].detach().numpy(), ds=ret["ds"].numpy(), d_outs=ret["d_outs"].detach().numpy(),
                ilens_out=ret["ilens"].numpy(), olens_out=ret["olens"].numpy(), l1_loss=l1.detach().numpy(),
                forward_sum_loss=fs.detach().numpy(), bin_loss=ret["bin_loss"].detach().numpy(), duration_loss=dur.detach().numpy())
    for n, m in model.named_modules():
        if hasattr(m, "attn") and isinstance(getattr(m, "attn"), torch.Tensor):
            dump["attn." + n] = m.attn.detach().numpy()
    model.eval()
    with torch.no_grad():
        rete = model(xs, torch.tensor(ilens), ys, torch.tensor(olens), dpi, dp_lengths=torch.tensor(ilens))
    dump["eval_after_outs"] = rete["after_outs"].numpy()
    # inference without ground truth (eval mode, running BatchNorm statistics as updated by the training step above);
    # the duration-predictor bias is raised so that the predicted durations are not all zero
    with torch.no_grad():
        model.duration_predictor.linear.bias.add_(1.0)
        outs, d_outs = model.inference(xs[0, :ilens[0]], dp_input=dpi[0, :ilens[0]])
        dump.update(inf_dp_bias=model.duration_predictor.linear.bias.detach().numpy().copy(), inf_outs=outs.numpy(), inf_d_outs=d_outs.numpy())
        model.duration_predictor.linear.bias.sub_(1.0)
        for k, v in model.state_dict().items():
            if "running_" in k:
                dump["inf_bn." + k] = v.numpy().copy()
    # ForwardSumLoss on its own (ragged lengths, incl. an infeasible utterance T < N that zero_infinity drops)
    g = torch.Generator().manual_seed(23)
    lp = torch.log_softmax(torch.randn(4, 40, 12, generator=g), -1)
    tl, fl = torch.tensor([12, 9, 5, 12]), torch.tensor([40, 31, 17, 8])
    lp.requires_grad_(True)
    fsl = ForwardSumLoss()(lp, tl, fl)
    fsl.backward()
    dump.update(fs_lp=lp.detach().numpy(), fs_tl=tl.numpy(), fs_fl=fl.numpy(), fs_loss=fsl.detach().numpy(), fs_grad=lp.grad.numpy())
    np.savez_compressed(os.path.join(GOLDEN, "aasvc_tiny.npz"), **dump)
    print("aasvc_tiny:", len(dump), "arrays")


def gen_aasvc_conv1d_tiny(name="aasvc_conv1d_tiny", layer="conv1d", k=1, seed=29):
    """Same step as gen_aasvc_tiny with the position-wise layer the AASVC class defaults to (models/aas_vc.py:52-53):
    MultiLayeredConv1d, kernel size 1, ReLU (modules/transformer/multi_layer_conv.py) instead of Linear + Swish.  Also the
    generator of the kernel-size-3 MultiLayeredConv1d and Conv1dLinear fixtures (multi_layer_conv.py:12-108)."""
    from seq2seq_vc.losses import DurationPredictorLoss, ForwardSumLoss, L1Loss
    from seq2seq_vc.models import AASVC

    torch.manual_seed(seed)
    model = AASVC(**AAS_HP, **dict(AAS_FIXED, positionwise_layer_type=layer, positionwise_conv_kernel_size=k))
    ref_shim.disable_dropout(model)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if p.dim() == 1 and ("norm" in n or "embed.1" in n or ".2." in n or ".1." in n):
                p.add_(0.1 * torch.randn_like(p))
    model.train()
    sd0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    xs, ilens, ys, olens, dpi = aasvc_oracle.synthetic_batch(3, 50, 44, ilens=[50, 41, 37], olens=[44, 38, 30], seed=31)
    ret = model(xs, torch.tensor(ilens), ys, torch.tensor(olens), dpi, dp_lengths=torch.tensor(ilens))
    l1 = L1Loss()(ret["after_outs"], ret["before_outs"], ret["ys"], ret["olens"])
    fs = ForwardSumLoss()(ret["log_p_attn"], ret["ilens"], ret["olens_reduced"])
    dur = DurationPredictorLoss()(ret["d_outs"], ret["ds"], ret["ilens"])
    (l1 + 2.0 * (fs + ret["bin_loss"]) + dur).backward()
    dump = {"sd." + k: v.numpy() for k, v in sd0.items()}
    dump.update({"grad." + k: p.grad.numpy() for k, p in model.named_parameters() if p.grad is not None})
    dump.update({"bn_after." + k: v.numpy() for k, v in model.state_dict().items() if "running_" in k})
    dump.update(xs=xs.numpy(), ilens=np.array(ilens), ys=ys.numpy(), olens=np.array(olens), dp_inputs=dpi.numpy(),
                after_outs=ret["after_outs"].detach().numpy(), before_outs=ret["before_outs"].detach().numpy(),
                log_p_attn=ret["log_p_attn"].detach().numpy(), ds=ret["ds"].numpy(), d_outs=ret["d_outs"].detach().numpy(),
                ilens_out=ret["ilens"].numpy(), olens_out=ret["olens"].numpy(), l1_loss=l1.detach().numpy(),
                forward_sum_loss=fs.detach().numpy(), bin_loss=ret["bin_loss"].detach().numpy(), duration_loss=dur.detach().numpy())
    for n, m in model.named_modules():
        if hasattr(m, "attn") and isinstance(getattr(m, "attn"), torch.Tensor):
            dump["attn." + n] = m.attn.detach().numpy()
    model.eval()
    with torch.no_grad():
        rete = model(xs, torch.tensor(ilens), ys, torch.tensor(olens), dpi, dp_lengths=torch.tensor(ilens))
    dump["eval_after_outs"] = rete["after_outs"].numpy()
    with torch.no_grad():                                # inference as in gen_aasvc_tiny (raised duration-predictor bias)
        model.duration_predictor.linear.bias.add_(1.0)
        outs, d_outs = model.inference(xs[0, :ilens[0]], dp_input=dpi[0, :ilens[0]])
        dump.update(inf_dp_bias=model.duration_predictor.linear.bias.detach().numpy().copy(), inf_outs=outs.numpy(), inf_d_outs=d_outs.numpy())
        model.duration_predictor.linear.bias.sub_(1.0)
        for k, v in model.state_dict().items():
            if "running_" in k:
                dump["inf_bn." + k] = v.numpy().copy()
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **dump)
    print(name + ":", len(dump), "arrays")


def gen_aasvc_conv1d_k3_tiny():
    gen_aasvc_conv1d_tiny("aasvc_conv1d_k3_tiny", "conv1d", 3, 37)


def gen_aasvc_conv1d_linear_k3_tiny():
    gen_aasvc_conv1d_tiny("aasvc_conv1d_linear_k3_tiny", "conv1d-linear", 3, 41)


FS_HP = dict(idim=80, odim=80, adim=32, aheads=2, elayers=2, eunits=48, dlayers=2, dunits=48, duration_predictor_input_dim=80,
             duration_predictor_layers=2, duration_predictor_chans=16, duration_predictor_kernel_size=3, postnet_layers=2, postnet_filts=5,
             postnet_chans=16, conformer_enc_kernel_size=7, conformer_dec_kernel_size=7)
FS_FIXED = dict(positionwise_layer_type="linear", positionwise_conv_kernel_size=1, duration_predictor_use_encoder_outputs=False,
                encoder_normalize_before=True, decoder_normalize_before=True, encoder_reduction_factor=1, decoder_reduction_factor=1,
                encoder_type="conformer", decoder_type="conformer", encoder_input_layer="conv2d", conformer_pos_enc_layer_type="rel_pos",
                conformer_self_attn_layer_type="rel_selfattn", use_macaron_style_in_conformer=True, use_cnn_in_conformer=True,
                teacher_model_decoder_reduction_factor=1)


def gen_fsvc_tiny():
    """FastSpeechVC in the configuration of egs/arctic/vc2/conf/fs2_vc.melmelmel.v1.yaml (conformer encoder / decoder, conv2d input
    layer, teacher durations) + L1Loss + DurationPredictorLoss as NARVCTrainer._train_step assembles them (trainers/nar_vc.py:53-99)."""
    from seq2seq_vc.losses import DurationPredictorLoss, L1Loss
    from seq2seq_vc.models import FastSpeechVC

    torch.manual_seed(51)
    model = FastSpeechVC(**FS_HP, **FS_FIXED)
    ref_shim.disable_dropout(model)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if p.dim() == 1 and ("norm" in n or ".2." in n or ".1." in n):
                p.add_(0.1 * torch.randn_like(p))
    model.train()
    sd0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(53)
    B, T = 3, 70
    ilens = [70, 61, 47]
    xs = torch.randn(B, T, 80, generator=g)
    tl = [((i - 2 + 1) // 2 - 2 + 1) // 2 for i in ilens]
    Tt = max(tl)
    ds = torch.randint(0, 5, (B, Tt), generator=g)
    for b in range(B):
        ds[b, tl[b]:] = 0
        ds[b, 0] = max(int(ds[b, 0]), 1)
    olens = ds.sum(1).tolist()
    ys = torch.randn(B, max(olens), 80, generator=g)
    for b in range(B):
        xs[b, ilens[b]:] = 0
        ys[b, olens[b]:] = 0
    out = model(xs, torch.tensor(ilens), ys, torch.tensor(olens), ds, torch.tensor(tl), xs, dp_lengths=torch.tensor(ilens))
    before, after, d_outs, ilens_o, olens_o, ys_o = out
    l1 = L1Loss()(after, before, ys_o, olens_o)
    dur = DurationPredictorLoss()(d_outs, ds, ilens_o)
    (l1 + dur).backward()
    dump = {"sd." + k: v.numpy() for k, v in sd0.items()}
    dump.update({"grad." + k: p.grad.numpy() for k, p in model.named_parameters() if p.grad is not None})
    dump.update({"bn_after." + k: v.numpy() for k, v in model.state_dict().items() if "running_" in k})
    dump.update(xs=xs.numpy(), ilens=np.array(ilens), ys=ys.numpy(), olens=np.array(olens), ds=ds.numpy(), dp_inputs=xs.numpy(),
                after_outs=after.detach().numpy(), before_outs=before.detach().numpy(), d_outs=d_outs.detach().numpy(),
                ilens_out=ilens_o.numpy(), olens_out=olens_o.numpy(), l1_loss=l1.detach().numpy(), duration_loss=dur.detach().numpy())
    for n, m in model.named_modules():
        if hasattr(m, "attn") and isinstance(getattr(m, "attn"), torch.Tensor):
            dump["attn." + n] = m.attn.detach().numpy()
    model.eval()
    with torch.no_grad():
        oe = model(xs, torch.tensor(ilens), ys, torch.tensor(olens), ds, torch.tensor(tl), xs, dp_lengths=torch.tensor(ilens))
    dump["eval_after_outs"] = oe[1].numpy()
    np.savez_compressed(os.path.join(GOLDEN, "fsvc_tiny.npz"), **dump)
    print("fsvc_tiny:", len(dump), "arrays", "olens", olens)


SDP_HP = dict(channels=16, kernel_size=3, dds_conv_layers=3, flows=4)


def gen_sdp_tiny():
    """StochasticDurationPredictor (modules/duration_predictor.py:131-304) called as AASVC._forward calls it
    (models/aas_vc.py:385-393,412-419).  The module draws its noise with torch.randn inside forward: the draw is recorded
    here and becomes an explicit input of the oracle / kernels.  ConvFlow.proj is zero-initialised by the reference (the
    splines start as the identity), so the weights are perturbed to exercise every spline branch incl. the linear tails."""
    from seq2seq_vc.modules.duration_predictor import StochasticDurationPredictor

    torch.manual_seed(41)
    m = StochasticDurationPredictor(channels=SDP_HP["channels"], kernel_size=SDP_HP["kernel_size"], dropout_rate=0.5,
                                    flows=SDP_HP["flows"], dds_conv_layers=SDP_HP["dds_conv_layers"], global_channels=-1)
    ref_shim.disable_dropout(m)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n.endswith(".proj.weight") and "flows" in n:
                p.add_(0.5 * torch.randn_like(p))
            elif n.endswith(".proj.bias") and "flows" in n:
                p.add_(0.3 * torch.randn_like(p))
            elif n.endswith(".m") or n.endswith(".logs"):
                p.add_(0.2 * torch.randn_like(p))
    m.train()
    B, T, C = 3, 13, SDP_HP["channels"]
    tl = [13, 9, 4]
    g = torch.Generator().manual_seed(43)
    dp = torch.randn(B, T, C, generator=g)
    ds = torch.randint(0, 9, (B, T), generator=g)
    mask = (torch.arange(T)[None, :] < torch.tensor(tl)[:, None])
    ds = ds * mask
    drawn = []
    real_randn = torch.randn

    def recording_randn(*a, **k):
        t = real_randn(*a, **k)
        drawn.append(t.clone())
        return t

    torch.randn = recording_randn
    try:
        torch.manual_seed(47)
        nll = m(dp.transpose(1, 2), mask.unsqueeze(1), w=ds.unsqueeze(1).float())
        dur_nll = nll / torch.sum(mask)
        dur_nll.sum().backward()
        e_q = drawn.pop()
        m.eval()
        with torch.no_grad():
            d = m(dp.transpose(1, 2), mask.unsqueeze(1), inverse=True, noise_scale=0.8).squeeze(1)
        z = drawn.pop()
        # a second inverse pass with large noise: drives elements into the linear tails of the splines
        with torch.no_grad():
            d_wide = m(dp.transpose(1, 2), mask.unsqueeze(1), inverse=True, noise_scale=4.0).squeeze(1)
        z_wide = drawn.pop()
    finally:
        torch.randn = real_randn
    dump = {"sd." + k: v.detach().numpy() for k, v in m.state_dict().items()}
    dump.update({"grad." + k: p.grad.numpy() for k, p in m.named_parameters() if p.grad is not None})
    dump.update(dp_inputs=dp.numpy(), text_lens=np.array(tl), ds=ds.numpy(), e_q=e_q.numpy(), dur_nll=dur_nll.detach().numpy(),
                z=z.numpy(), d_outs=torch.clamp(d, max=10).numpy(), z_wide=z_wide.numpy(), d_outs_wide=torch.clamp(d_wide, max=10).numpy())
    np.savez_compressed(os.path.join(GOLDEN, "sdp_tiny.npz"), **dump)
    print("sdp_tiny:", len(dump), "arrays; dur_nll", dur_nll.detach().numpy(), "d_outs", d[0].numpy())


def gen_mas():
    from seq2seq_vc.modules.alignments import _monotonic_alignment_search, viterbi_decode

    rng = np.random.default_rng(3)
    dump = {}
    shapes = [(1, 1), (5, 3), (7, 7), (17, 1), (33, 32), (64, 13), (96, 24), (50, 50)]
    for n, (tm, ti) in enumerate(shapes):
        for kind in ("rand", "ties", "inf"):
            lp = torch.log_softmax(torch.from_numpy(rng.standard_normal((tm, ti)).astype(np.float32)), -1).numpy()
            if kind == "ties":
                lp = (np.round(lp * 2) / 2).astype(np.float32)
            if kind == "inf" and ti > 1:
                lp[:, ti // 2:] = -np.inf
            dump[f"lp.{n}.{kind}"] = lp
            dump[f"path.{n}.{kind}"] = _monotonic_alignment_search(lp).astype(np.int64)
    B, TF, TT = 6, 72, 20
    lp = torch.log_softmax(torch.from_numpy(rng.standard_normal((B, TF, TT)).astype(np.float32)), -1)
    tl = torch.tensor([20, 17, 9, 20, 1, 12])
    fl = torch.tensor([72, 60, 40, 20, 33, 71])
    for b in range(B):  # the reference masks padded text columns with -inf before log_softmax
        lp[b, :, tl[b]:] = -np.inf
    ds, bin_loss = viterbi_decode(lp, tl, fl)
    dump.update(vd_lp=lp.numpy(), vd_tl=tl.numpy(), vd_fl=fl.numpy(), vd_ds=ds.numpy(), vd_bin_loss=np.float32(bin_loss))
    np.savez_compressed(os.path.join(GOLDEN, "mas.npz"), **dump)
    print("mas:", len(dump), "arrays")


def gen_kats():
    """Docstring known-answer examples of the reference."""
    from seq2seq_vc.layers.utils import make_non_pad_mask, make_pad_mask
    from seq2seq_vc.losses.guided_attention_loss import GuidedAttentionLoss
    from seq2seq_vc.modules.transformer.mask import subsequent_mask

    dump = dict(pad_mask_532=make_pad_mask([5, 3, 2]).numpy(), non_pad_mask_532=make_non_pad_mask([5, 3, 2]).numpy(),
                subsequent_mask_3=subsequent_mask(3).numpy())
    ga = GuidedAttentionLoss(sigma=0.4)
    dump["ga_5x5"] = ga._make_guided_attention_mask(torch.tensor(5), torch.tensor(5), 0.4).numpy()
    dump["ga_6x3"] = ga._make_guided_attention_mask(torch.tensor(3), torch.tensor(6), 0.4).numpy()
    dump["ga_masks_52_85"] = ga._make_masks(torch.tensor([5, 2]), torch.tensor([8, 5])).numpy()
    # a seeded multi-head guided-attention loss value
    from seq2seq_vc.losses.guided_attention_loss import GuidedMultiHeadAttentionLoss

    g = torch.Generator().manual_seed(5)
    att = torch.softmax(torch.randn(3, 4, 12, 9, generator=g), -1)
    il, ol = torch.tensor([9, 7, 4]), torch.tensor([12, 10, 5])
    dump.update(gmh_att=att.numpy(), gmh_ilens=il.numpy(), gmh_olens=ol.numpy(),
                gmh_loss=GuidedMultiHeadAttentionLoss(sigma=0.4, alpha=1.0)(att, il, ol).numpy())
    np.savez_compressed(os.path.join(GOLDEN, "kats.npz"), **dump)
    print("kats:", len(dump), "arrays")



def gen_subsampling_tiny():
    """Conv2dSubsampling2 / 6 / 8 (modules/transformer/subsampling.py:108-279) with their default PositionalEncoding, dropout 0:
    output, sliced mask and every parameter gradient of sum(y * r) for a fixed random r."""
    from seq2seq_vc.modules.transformer.subsampling import Conv2dSubsampling2, Conv2dSubsampling6, Conv2dSubsampling8

    out = {}
    g = torch.Generator().manual_seed(71)
    for n, cls in ((2, Conv2dSubsampling2), (6, Conv2dSubsampling6), (8, Conv2dSubsampling8)):
        torch.manual_seed(70 + n)
        m = cls(40, 16, 0.0)
        m.train()
        B, T = 2, 61
        x = torch.randn(B, T, 40, generator=g)
        mask = torch.ones(B, 1, T, dtype=torch.bool)
        mask[1, :, 50:] = False
        y, ym = m(x, mask)
        r = torch.randn(y.shape, generator=g)
        (y * r).sum().backward()
        out[f"s{n}.x"], out[f"s{n}.y"], out[f"s{n}.r"] = x.numpy(), y.detach().numpy(), r.numpy()
        out[f"s{n}.mask_in"], out[f"s{n}.mask_out"] = mask.numpy(), ym.numpy()
        for k, v in m.state_dict().items():
            out[f"s{n}.sd.{k}"] = v.numpy()
        for k, p in m.named_parameters():
            out[f"s{n}.grad.{k}"] = p.grad.numpy()
    np.savez_compressed(os.path.join(GOLDEN, "subsampling_tiny.npz"), **out)

if __name__ == "__main__":
    ref_shim.install()
    os.makedirs(GOLDEN, exist_ok=True)
    import sys

    gens = dict(vtn_tiny=gen_vtn_tiny, vtn_rfactor=gen_vtn_rfactor, vtn_conformer_tiny=gen_vtn_conformer_tiny, vtn_convffn_tiny=gen_vtn_convffn_tiny, fsvc_tiny=gen_fsvc_tiny, tts_tiny=gen_tts_tiny, aasvc_tiny=gen_aasvc_tiny, aasvc_conv1d_tiny=gen_aasvc_conv1d_tiny, aasvc_conv1d_k3_tiny=gen_aasvc_conv1d_k3_tiny,
                aasvc_conv1d_linear_k3_tiny=gen_aasvc_conv1d_linear_k3_tiny,
                mas=gen_mas, kats=gen_kats, sdp_tiny=gen_sdp_tiny, subsampling_tiny=gen_subsampling_tiny)
    for name in (sys.argv[1:] or list(gens)):      # e.g. `python oracle/gen_golden.py aasvc_conv1d_tiny` adds one fixture
        gens[name]()
