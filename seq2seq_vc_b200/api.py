"""Host-side mirror of the reference's plugin surface for the hot path.

Same names, constructor kwargs, call signatures, return structures and state-dict keys as
 * seq2seq_vc.models.VTN                                  (models/vtn.py:14-300)
 * seq2seq_vc.losses.Seq2SeqLoss                          (losses/seq2seq_loss.py:13-59)
 * seq2seq_vc.losses.GuidedMultiHeadAttentionLoss         (losses/guided_attention_loss.py:133-165)
 * seq2seq_vc.modules.alignments.viterbi_decode           (modules/alignments.py:281-310)
 * seq2seq_vc.bin.preprocess.logmelfilterbank             (bin/preprocess.py:30-92)
so that `getattr(module, config["model_type"])(**config["model_params"]).to(device)` and the
reference trainers work unchanged.  Every arithmetic step runs in libs2svc_b200.so; a missing
library or a CPU tensor raises (there is no fallback).
"""
from __future__ import annotations

import logging
import math
import os
import weakref
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib, ops
from ._lib import S2SError
from .vtn_engine import VTNEngine, default_hparams

_f32 = torch.float32
_i32 = torch.int32


def _require_cuda(t: torch.Tensor, who: str) -> None:
    """There is no CPU path: the modules refuse host tensors instead of falling back to anything."""
    if not t.is_cuda:
        raise S2SError(f"seq2seq_vc_b200.{who} runs on a B200 only (no CPU fallback): move the model and batch to cuda")


def _host_lens(v) -> List[int]:
    if isinstance(v, torch.Tensor):
        return [int(x) for x in v.detach().cpu().tolist()]
    return [int(x) for x in v]


# =================================================================================================
# module tree that reproduces the reference's parameter names and attribute seams
# =================================================================================================
class _Node(torch.nn.Module):
    """Parameter container; integer indexing walks digit-named children (Sequential/ModuleList-like)."""

    def __getitem__(self, idx):
        keys = sorted((k for k in self._modules if k.isdigit()), key=int)
        if not keys and "out" in self._modules:      # Conv2dSubsampling.__getitem__ (subsampling.py:96-105)
            if idx != -1:
                raise NotImplementedError("Support only `-1` (for `reset_parameters`).")
            return self._modules["out"][idx]
        return self._modules[keys[idx]]

    def __len__(self):
        return len([k for k in self._modules if k.isdigit()])

    def forward(self, *a, **k):
        raise S2SError("sub-modules of the B200 VTN are parameter containers; call the model itself")


class MultiHeadedAttention(_Node):
    """Holds linear_{q,k,v,out}; `.attn` is refreshed by every model forward (attention.py:81-85)."""
    attn = None


def _dropin_graph_forward(model, xs, ys, ilens, olens):
    """Opt-in CUDA-graph mode of the drop-in modules (`model.use_graph = True`, training only): the reference trainer's eager
    step around the drop-in is bound by the host cost of ~450 kernel launches, so from the second batch of a (B, T, L) shape on
    the engine's forward is replayed from a captured graph (the first batch runs eagerly and allocates every buffer outside
    the graph's pool; `_dropin_graph_backward` does the same for the backward).  Returns None while the shape is still new.
    Everything forward() assigns on the engine (shapes, output / attention views, dropout-site counter) is snapshotted at
    capture time and restored on replay, so backward() and the attribute seams read the state of THIS shape."""
    eng = model.engine
    B, T, L = xs.shape[0], xs.shape[1], ys.shape[1]
    eng.prepare(B, T, L, ilens, olens)
    graphs = eng._dropin_graphs
    key = (B, T, L)
    e = graphs.get(key)
    if e is None:
        graphs[key] = {"seen_fwd": True}
        return None
    if "gF" not in e:
        if not e.get("seen_bwd"):
            return None                     # no eager backward of this shape yet: its buffers are not all allocated
        sx, sy = torch.empty_like(xs), torch.empty_like(ys)
        torch.cuda.synchronize()
        before = dict(eng.__dict__)
        g = torch.cuda.CUDAGraph()
        with _lib.graph_capture(g):
            eng.p16_dirty = True            # an external optimizer owns the parameters: the bf16 shadow is refreshed inside the graph
            outs = eng.forward(sx, sy)
        e.update(gF=g, sx=sx, sy=sy, outs=outs,
                 pystate={k: v for k, v in eng.__dict__.items() if k not in before or before[k] is not v})
    e["sx"].copy_(xs, non_blocking=True)
    e["sy"].copy_(ys, non_blocking=True)
    e["gF"].replay()
    eng.__dict__.update(e["pystate"])
    return e


def _dropin_graph_backward(eng, e, grads, fresh: bool) -> None:
    if "d" not in e:
        e["d"] = tuple(torch.zeros_like(t) for t in e["outs"])
    for dst, g in zip(e["d"], grads):
        if g is None:
            dst.zero_()
        else:
            dst.copy_(g, non_blocking=True)         # engine.backward() uses its upstream gradients as scratch: always re-fill
    name = "gB0" if fresh else "gB1"                # zero the flat gradient buffer first / accumulate (gradient accumulation)
    if name not in e:
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with _lib.graph_capture(g):
            eng.backward(*e["d"], zero_grad=fresh)
        e[name] = g
    e[name].replay()


class _VTNFunction(torch.autograd.Function):
    """Whole-model autograd node: forward = engine.forward, backward = the engine's hand-written backward.
    When `n_att` > 0 the concatenated source-attention maps (last layers x first heads) are a differentiable
    output too, so a guided-attention loss on them back-propagates into the attention kernels."""

    @staticmethod
    def forward(ctx, model, xs, ys, ilens, olens, n_att_layers, n_att_heads, *params):
        eng = model.engine
        entry = None
        if model.use_graph and eng.training and n_att_layers == 0:
            entry = _dropin_graph_forward(model, xs, ys, ilens, olens)
        if entry is None:
            after, before, logits = eng.forward(xs, ys, ilens, olens)
        else:
            after, before, logits = entry["outs"]
        ctx.entry = entry
        ctx.model = model
        ctx.token = model._fwd_token
        ctx.att = (n_att_layers, n_att_heads)
        # outputs are engine-owned buffers: hand out copies so later forwards cannot clobber them
        outs = [after.clone(), before.clone(), logits.clone()]
        if n_att_layers > 0:
            names = [f"decoder.decoders.{l}.src_attn" for l in reversed(range(eng.hp["dlayers"]))][:n_att_layers]
            ctx.att_names = names
            outs.append(torch.cat([eng.attn[n][:, :n_att_heads] for n in names], dim=1).float())
        return tuple(outs)

    @staticmethod
    def backward(ctx, d_after, d_before, d_logits, d_att_ws=None):
        model = ctx.model
        if ctx.token != model._fwd_token:
            raise S2SError("backward() after a newer forward(): the engine keeps one set of activations")
        eng = model.engine
        fresh = all(p.grad is None for p in model.parameters())
        dt = eng.adt
        d_att = None
        if d_att_ws is not None and ctx.att[0] > 0:
            nh = ctx.att[1]
            d_att = {}
            for i, name in enumerate(ctx.att_names):
                P = eng.attn[name]                                            # (B, H, T1, T2) view of the (.., ld) buffer
                full = torch.zeros(P.shape[0], P.shape[1], P.shape[2], (P.shape[3] + 7) // 8 * 8, dtype=dt, device=P.device)
                full[:, :nh, :, :P.shape[3]] = d_att_ws[:, i * nh:(i + 1) * nh].to(dt)
                d_att[name] = full

        def z(g, like):
            return torch.zeros_like(like, dtype=dt) if g is None else g.to(dt).contiguous()

        if ctx.entry is not None:
            _dropin_graph_backward(eng, ctx.entry, (d_after, d_before, d_logits), fresh)
        else:
            eng.backward(z(d_after, eng.after), z(d_before, eng.before), z(d_logits, eng.logits), d_att=d_att, zero_grad=fresh)
            seen = eng._dropin_graphs.get((eng.shapes["B"], eng.shapes["T"], eng.shapes["L"]))
            if seen is not None:
                seen["seen_bwd"] = True
        model._sync_gradients()
        model._bind_grads()
        return (None,) * (7 + len(model._param_names))


class VTN(torch.nn.Module):
    """Drop-in for seq2seq_vc.models.VTN (transformer encoder/decoder variant)."""

    def __init__(self, idim, odim, dprenet_layers=2, dprenet_units=256, adim=384, aheads=4, encoder_type="transformer",
                 decoder_type="transformer", elayers=6, eunits=1536, dlayers=6, dunits=1536, postnet_layers=5, postnet_filts=5,
                 postnet_chans=256, positionwise_layer_type: str = "linear", positionwise_conv_kernel_size: int = 1,
                 dprenet_dropout_rate=0.5, transformer_enc_dropout_rate: float = 0.1,
                 transformer_enc_positional_dropout_rate: float = 0.1, transformer_enc_attn_dropout_rate: float = 0.1,
                 use_batch_norm=True, encoder_normalize_before=True, decoder_normalize_before=False,
                 encoder_concat_after=False, decoder_concat_after=False, decoder_reduction_factor=2, spk_embed_dim=None,
                 spk_embed_integration_type="add", initial_encoder_alpha=1.0, initial_decoder_alpha=1.0,
                 use_guided_attn_loss=False, num_heads_applied_guided_attn=2, num_layers_applied_guided_attn=2,
                 conformer_rel_pos_type: str = "legacy", conformer_pos_enc_layer_type: str = "rel_pos",
                 conformer_self_attn_layer_type: str = "rel_selfattn", use_macaron_style_in_conformer: bool = True,
                 use_cnn_in_conformer: bool = True, zero_triu: bool = False, conformer_enc_kernel_size: int = 7,
                 conformer_dec_kernel_size: int = 31, compute_dtype: str = "float32", device=None, seed: int = 0,
                 use_graph: bool = False):
        super().__init__()
        self.use_graph = bool(use_graph)     # replay captured CUDA graphs of forward / backward per batch shape (training mode)
        unsupported = []
        if encoder_type not in ("transformer", "conformer") or decoder_type != "transformer":
            unsupported.append("encoder_type not in ('transformer', 'conformer') / decoder_type != 'transformer'")
        conformer = encoder_type == "conformer"
        if conformer:
            # relative positional encoding compatibility (models/vtn.py:83-104): "legacy" turns the default rel_pos / rel_selfattn
            # into their legacy forms, "latest" keeps the new ones
            if conformer_rel_pos_type == "legacy":
                if conformer_pos_enc_layer_type == "rel_pos":
                    conformer_pos_enc_layer_type = "legacy_rel_pos"
                    logging.warning("Fallback to conformer_pos_enc_layer_type = 'legacy_rel_pos' due to the compatibility. "
                                    "If you want to use the new one, please use conformer_pos_enc_layer_type = 'latest'.")
                if conformer_self_attn_layer_type == "rel_selfattn":
                    conformer_self_attn_layer_type = "legacy_rel_selfattn"
                    logging.warning("Fallback to conformer_self_attn_layer_type = 'legacy_rel_selfattn' due to the compatibility. "
                                    "If you want to use the new one, please use conformer_pos_enc_layer_type = 'latest'.")
            elif conformer_rel_pos_type == "latest":
                assert conformer_pos_enc_layer_type != "legacy_rel_pos"
                assert conformer_self_attn_layer_type != "legacy_rel_selfattn"
            else:
                raise ValueError(f"Unknown rel_pos_type: {conformer_rel_pos_type}")
            pair = (conformer_pos_enc_layer_type, conformer_self_attn_layer_type)
            if pair not in (("legacy_rel_pos", "legacy_rel_selfattn"), ("rel_pos", "rel_selfattn")):
                unsupported.append(f"conformer positional / attention pair {pair} (rel_pos + rel_selfattn and their legacy forms are covered)")
            if not use_macaron_style_in_conformer or not use_cnn_in_conformer or zero_triu:
                unsupported.append("conformer without macaron / CNN module, zero_triu")
            if positionwise_layer_type not in ("linear", "conv1d", "conv1d-linear"):
                unsupported.append("positionwise_layer_type not in ('linear', 'conv1d', 'conv1d-linear')")
        elif positionwise_layer_type not in ("linear", "conv1d", "conv1d-linear"):
            unsupported.append("positionwise_layer_type not in ('linear', 'conv1d', 'conv1d-linear')")
        if not use_batch_norm or not encoder_normalize_before or decoder_normalize_before:
            unsupported.append("non-default normalisation wiring")
        if encoder_concat_after or decoder_concat_after or spk_embed_dim is not None:
            unsupported.append("concat_after / speaker embeddings")
        if unsupported:
            raise NotImplementedError("B200 VTN hot path does not cover: " + ", ".join(unsupported))
        if use_guided_attn_loss and getattr(self, "_encoder_input", "conv2d") != "embed":
            # the reference's VTN stores the flag and never uses it (models/vtn.py:73; ARVCTrainer has no guided-attention
            # criterion); the maps VTN.forward returns are detached here, so a loss on them would train nothing
            raise NotImplementedError("use_guided_attn_loss on VTN: only TransformerTTS returns differentiable attention maps")
        self.idim, self.odim = idim, odim
        self.spk_embed_dim = None
        self.decoder_reduction_factor = decoder_reduction_factor
        self.use_guided_attn_loss = use_guided_attn_loss
        self.num_heads_applied_guided_attn = num_heads_applied_guided_attn
        self.num_layers_applied_guided_attn = num_layers_applied_guided_attn
        self.encoder_type, self.decoder_type = encoder_type, decoder_type
        self.hp = default_hparams(idim=idim, odim=odim, dprenet_layers=dprenet_layers, dprenet_units=dprenet_units, adim=adim,
                                  aheads=aheads, elayers=elayers, eunits=eunits, dlayers=dlayers, dunits=dunits,
                                  postnet_layers=postnet_layers, postnet_filts=postnet_filts, postnet_chans=postnet_chans,
                                  dprenet_dropout_rate=dprenet_dropout_rate,
                                  transformer_enc_dropout_rate=transformer_enc_dropout_rate,
                                  decoder_reduction_factor=decoder_reduction_factor,
                                  initial_encoder_alpha=initial_encoder_alpha, initial_decoder_alpha=initial_decoder_alpha,
                                  encoder_input=getattr(self, "_encoder_input", "conv2d"))
        if not conformer and positionwise_layer_type != "linear":     # the Transformer ENCODER only (the decoder is built without it)
            self.hp.update(positionwise_layer_type=positionwise_layer_type, positionwise_conv_kernel_size=positionwise_conv_kernel_size)
            default_hparams(**self.hp)
        if conformer:       # models/vtn.py:122-143: the conformer encoder takes the positional / attention dropout rates of the constructor
            self.hp.update(encoder_type="conformer", conformer_enc_kernel_size=conformer_enc_kernel_size,
                           conformer_rel_pos_type="legacy" if conformer_self_attn_layer_type == "legacy_rel_selfattn" else "latest",
                           enc_positional_dropout_rate=transformer_enc_positional_dropout_rate,
                           enc_attn_dropout_rate=transformer_enc_attn_dropout_rate, positionwise_layer_type=positionwise_layer_type,
                           positionwise_conv_kernel_size=positionwise_conv_kernel_size)
            default_hparams(**self.hp)      # validates the combination
        # compute_dtype: "bf16" (tcgen05, bf16 activations) | "float32" (float32 activations, fp32-accurate tcgen05 GEMMs through a
        # bf16 split: the parity mode) | "float32_simt" (float32 on the CUDA cores: the numerical yard-stick)
        self._bf16 = compute_dtype in ("bf16", "bfloat16", torch.bfloat16)
        self._fp32_gemm = "simt" if compute_dtype == "float32_simt" else "tc"
        self._seed = seed
        self._fwd_token = 0
        self.engine: Optional[VTNEngine] = None
        dev = torch.device(device) if device is not None else torch.device("cpu")
        self._build(dev)

    # ---- construction / device movement -------------------------------------------------------
    def _build(self, device, state: Optional[Dict[str, torch.Tensor]] = None) -> None:
        self.engine = VTNEngine(self.hp, device=device, bf16=self._bf16, seed=self._seed, fp32_gemm=self._fp32_gemm)
        if state is not None:
            self.engine.load_state_dict(state)
        self._attach_graph_cache()
        self._modules.clear()
        self._param_names: List[str] = []
        st = self.engine.store
        for name in st.names():
            node, leaf = self._node_for(name)
            node.register_parameter(leaf, torch.nn.Parameter(st.p(name), requires_grad=True))
            self._param_names.append(name)
        for name, buf in self.engine.buffers.items():
            node, leaf = self._node_for(name)
            node.register_buffer(leaf, buf)

    use_graph = False

    def _attach_graph_cache(self) -> None:
        eng = self.engine
        eng._dropin_graphs = {}

        def on_evict(sig, graphs=eng._dropin_graphs):       # the engine dropped this shape's buffers: its graphs go with them
            graphs.pop(tuple(sig[:3]), None)

        eng._evict_listeners.append(on_evict)

    def _node_for(self, dotted: str):
        parts = dotted.split(".")
        node = self
        for i, part in enumerate(parts[:-1]):
            nxt = node._modules.get(part)
            if nxt is None:
                nxt = MultiHeadedAttention() if part in ("self_attn", "src_attn") else _Node()
                node.add_module(part, nxt)
                if isinstance(nxt, MultiHeadedAttention):
                    # registration order of the reference (attention.py:27-31), not the storage order (K / V of a source
                    # attention are adjacent in the flat buffer): optimizer state dicts are keyed by parameter ORDER, so a
                    # reference checkpoint's Adam moments must land on the same parameters (trainers/base.py:108-121)
                    for child in ("linear_q", "linear_k", "linear_v", "linear_out"):
                        nxt.add_module(child, _Node())
            node = nxt
        return node, parts[-1]

    def _apply(self, fn, recurse=True):
        """.to(device) / .cuda(): move the flat stores as a whole and re-bind the parameter views."""
        probe = fn(torch.zeros(1, dtype=_f32, device=self.engine.device))
        if probe.dtype != _f32:
            raise NotImplementedError("parameters stay float32; select bf16 compute with compute_dtype='bf16'")
        if probe.device != self.engine.device:
            state = {k: v.detach().cpu() for k, v in self.engine.state_dict().items()}
            self._build(probe.device, state)
        return self

    _ddp_group = False      # set by DistributedDataParallel (None = the default process group)

    def _sync_gradients(self) -> None:
        """Data-parallel mean of the flat gradient buffer at the end of backward, when wrapped by this package's
        DistributedDataParallel (the path shards by utterance batch: bin/vc_train.py:423-431).  Under gradient accumulation
        the buffer already holds the rank-invariant mean of earlier micro-steps, which the mean leaves unchanged."""
        pg = self._ddp_group
        if pg is False or not (torch.distributed.is_available() and torch.distributed.is_initialized()):
            return
        world = torch.distributed.get_world_size(pg)
        if world > 1:
            G = self.engine.store.G
            torch.distributed.all_reduce(G, group=pg)
            G.mul_(1.0 / world)

    def _bind_grads(self, unused: tuple = ()) -> None:
        """Point every parameter's .grad at its slice of the flat gradient buffer.  Parameters under the `unused` name
        prefixes took no part in this backward: like torch autograd, leave their .grad as None if it is None (optimizers
        skip such parameters, which keeps e.g. Adam's per-parameter step count identical to the reference's)."""
        st = self.engine.store
        for name, p in self.named_parameters():
            if not p.requires_grad:          # utils/model_io.py:95-111 freeze_modules: frozen parameters never get a gradient
                continue
            if unused and p.grad is None and name.startswith(unused):
                continue
            g = st.g(name)
            if p.grad is None or p.grad.data_ptr() != g.data_ptr():
                p.grad = g

    def _publish_attn(self) -> None:
        """`.attn` of the attention modules (attention.py:81-85).  The bf16 flash-attention path keeps only the maps the
        engine's `attn_emit` policy asks for in HBM (default: the source-attention maps, the ones VTN.forward returns);
        the others read None instead of a stale tensor.  `model.engine.attn_emit = "all"` restores every map."""
        eng = self.engine
        for name, mod in self.named_modules():
            if isinstance(mod, MultiHeadedAttention):
                P = eng.attn.get(name)
                mod.attn = None if P is None else (P.float() if P.dtype != _f32 else P)

    def train(self, mode: bool = True):
        super().train(mode)
        if self.engine is not None:
            self.engine.training = bool(mode)
        return self

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        out = super().load_state_dict(state_dict, strict=strict)
        self.engine.p16_dirty = True
        return out

    # ---- forward (vtn.py:207-300) ----------------------------------------------------------------
    def forward(self, xs, ilens, ys, labels, olens, spembs=None, *args, **kwargs):
        _require_cuda(xs, "VTN")
        eng = self.engine
        eng.p16_dirty = True       # parameters may have been updated by an external optimizer
        il, ol = _host_lens(ilens), _host_lens(olens)
        r = self.decoder_reduction_factor
        max_ilen, max_olen = max(il), max(ol)
        xs = xs[:, :max_ilen].to(_f32).contiguous()
        ys = ys[:, :max_olen].to(_f32).contiguous()
        labels = labels[:, :max_olen].to(_f32).contiguous()
        if r > 1:
            assert all(o >= r for o in ol), "Output length must be greater than or equal to reduction factor."
        self._fwd_token += 1
        if eng.training:
            ops.step_advance(None, eng.seed_dev)     # fresh dropout masks per forward (the fused steps advance it in their optimizer tail)
        after, before, logits = _VTNFunction.apply(self, xs, ys, il, ol, 0, 0, *self.parameters())
        Lo = after.shape[1]
        # target fix-ups (vtn.py:262-274)
        olens_out = torch.tensor(eng.olens_fix_host, dtype=torch.int64, device=xs.device)
        if r > 1:
            labels_out = torch.empty(labels.shape[0], Lo, dtype=_f32, device=xs.device)
            ops.fix_targets(labels, eng.olens_fix, labels_out, None, r)
        else:
            labels_out = labels
        ys_out = ys[:, :Lo]
        ilens_ds_st = torch.tensor(eng.ilens_ds_st, dtype=torch.int64, device=xs.device)
        olens_in = torch.tensor(eng.olens_in_host, dtype=torch.int64, device=xs.device)
        att_ws = []
        self._publish_attn()
        for l in reversed(range(self.hp["dlayers"])):           # vtn.py:280-287 (list, last layer first)
            att_ws.append(self.decoder.decoders[l].src_attn.attn)
        return after.float(), before.float(), logits.float(), ys_out, labels_out, olens_out, (att_ws, ilens_ds_st, olens_in)


class DistributedDataParallel(torch.nn.Module):
    """Stand-in for the wrapper the reference puts around the model under --distributed (apex DistributedDataParallel,
    bin/vc_train.py:423-431) for the drop-in modules of this package.  Their backward is one hand-written pass that fills a
    flat gradient buffer, so autograd-hook based wrappers (torch / apex DDP) never see a gradient; this one broadcasts rank 0's
    parameters and buffers at construction (as DDP constructors do) and makes the module average that buffer over the process
    group at the end of every backward.  Exposes `.module` (trainers/base.py:98-101,115-118 use it for checkpoints).
    The fused VTNTrainStep / AASVCTrainStep do their own all-reduce and do not need it."""

    def __init__(self, module, process_group=None, **ignored):
        super().__init__()
        if not isinstance(module, VTN):
            raise TypeError("seq2seq_vc_b200.DistributedDataParallel wraps the drop-in modules of this package")
        self.module = module
        module._ddp_group = process_group
        dist = torch.distributed
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(process_group) > 1:
            src = dist.get_global_rank(process_group, 0) if process_group is not None else 0
            eng = module.engine
            dist.broadcast(eng.store.P, src=src, group=process_group)
            for buf in eng.buffers.values():
                if buf.dtype.is_floating_point:
                    dist.broadcast(buf, src=src, group=process_group)
            eng.p16_dirty = True

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    def inference(self, *args, **kwargs):
        return self.module.inference(*args, **kwargs)


def _vtn_inference(self, x, inference_args, spemb=None, *args, **kwargs):
    """Drop-in for VTN.inference (models/vtn.py:302-394): x (T, idim) -> (outs (L, odim), probs (L,), att_ws (#layers, #heads, L/r, T'))."""
    if spemb is not None:
        raise NotImplementedError("speaker embeddings are outside the hot path")
    _require_cuda(x, "VTN")
    self.engine.p16_dirty = True
    outs, probs, att_ws = self.engine.inference(x, inference_args["threshold"], inference_args["minlenratio"], inference_args["maxlenratio"])
    for l in range(self.hp["dlayers"]):                    # `.attn` of the source-attention modules, as the reference leaves it
        self.decoder.decoders[l].src_attn.attn = att_ws[l].unsqueeze(0)
    return outs, probs, att_ws


VTN.inference = _vtn_inference          # TransformerTTS inherits it: the engine appends <eos> and embeds the tokens


class TransformerTTS(VTN):
    """Drop-in for seq2seq_vc.models.TransformerTTS (models/transformer_tts.py:13-229): token-embedding encoder,
    same decoder / heads / postnet as VTN, guided-attention maps returned as one differentiable tensor."""

    def __init__(self, idim, odim, dprenet_layers=2, dprenet_units=256, adim=384, aheads=4, elayers=6, eunits=1536, dlayers=6,
                 dunits=1536, postnet_layers=5, postnet_filts=5, postnet_chans=256, dprenet_dropout_rate=0.5, use_batch_norm=True,
                 encoder_normalize_before=True, decoder_normalize_before=False, encoder_concat_after=False,
                 decoder_concat_after=False, decoder_reduction_factor=2, spk_embed_dim=None, spk_embed_integration_type="add",
                 initial_encoder_alpha=1.0, initial_decoder_alpha=1.0, use_guided_attn_loss=False,
                 num_heads_applied_guided_attn=2, num_layers_applied_guided_attn=2, compute_dtype: str = "float32", device=None,
                 seed: int = 0, use_graph: bool = False):
        self._encoder_input = "embed"
        super().__init__(idim, odim, dprenet_layers=dprenet_layers, dprenet_units=dprenet_units, adim=adim, aheads=aheads,
                         elayers=elayers, eunits=eunits, dlayers=dlayers, dunits=dunits, postnet_layers=postnet_layers,
                         postnet_filts=postnet_filts, postnet_chans=postnet_chans, dprenet_dropout_rate=dprenet_dropout_rate,
                         use_batch_norm=use_batch_norm, encoder_normalize_before=encoder_normalize_before,
                         decoder_normalize_before=decoder_normalize_before, encoder_concat_after=encoder_concat_after,
                         decoder_concat_after=decoder_concat_after, decoder_reduction_factor=decoder_reduction_factor,
                         spk_embed_dim=spk_embed_dim, spk_embed_integration_type=spk_embed_integration_type,
                         initial_encoder_alpha=initial_encoder_alpha, initial_decoder_alpha=initial_decoder_alpha,
                         use_guided_attn_loss=use_guided_attn_loss, num_heads_applied_guided_attn=num_heads_applied_guided_attn,
                         num_layers_applied_guided_attn=num_layers_applied_guided_attn, compute_dtype=compute_dtype,
                         device=device, seed=seed, use_graph=use_graph)
        self.eos = idim - 1
        self.padding_idx = 0

    def forward(self, xs, ilens, ys, labels, olens, spembs=None, *args, **kwargs):
        _require_cuda(xs, "TransformerTTS")
        eng = self.engine
        eng.p16_dirty = True
        il, ol = _host_lens(ilens), _host_lens(olens)
        r = self.decoder_reduction_factor
        xs = xs[:, :max(il)].to(torch.int64).contiguous()
        ys = ys[:, :max(ol)].to(_f32).contiguous()
        labels = labels[:, :max(ol)].to(_f32).contiguous()
        if r > 1:
            assert all(o >= r for o in ol), "Output length must be greater than or equal to reduction factor."
        self._fwd_token += 1
        nl = self.num_layers_applied_guided_attn if self.use_guided_attn_loss else 0
        nh = self.num_heads_applied_guided_attn if self.use_guided_attn_loss else 0
        if eng.training:
            ops.step_advance(None, eng.seed_dev)     # fresh dropout masks per forward
        outs = _VTNFunction.apply(self, xs, ys, il, ol, nl, nh, *self.parameters())
        after, before, logits = outs[:3]
        att_ws = outs[3] if nl > 0 else []
        Lo = after.shape[1]
        olens_out = torch.tensor(eng.olens_fix_host, dtype=torch.int64, device=xs.device)
        if r > 1:
            labels_out = torch.empty(labels.shape[0], Lo, dtype=_f32, device=xs.device)
            ops.fix_targets(labels, eng.olens_fix, labels_out, None, r)
        else:
            labels_out = labels
        self._publish_attn()
        ilens_out = torch.tensor(eng.ilens_ds_st, dtype=torch.int64, device=xs.device)      # ilens + 1 (transformer_tts.py:142)
        olens_in = torch.tensor(eng.olens_in_host, dtype=torch.int64, device=xs.device)
        return after.float(), before.float(), logits.float(), ys[:, :Lo], labels_out, olens_out, (att_ws, ilens_out, olens_in)


# =================================================================================================
# losses
# =================================================================================================
class _Seq2SeqLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, after, before, logits, ys, labels, olens_dev, pos_weight):
        B, L, odim = after.shape
        dev = after.device
        losses = torch.empty(2, dtype=_f32, device=dev)
        ws = torch.empty(4, dtype=_f32, device=dev)
        d_after, d_before, d_logits = torch.empty_like(after), torch.empty_like(before), torch.empty_like(logits)
        ops.seq2seq_loss(after, before, logits, ys, labels, olens_dev, pos_weight, losses, d_after, d_before, d_logits, ws)
        ctx.save_for_backward(d_after, d_before, d_logits)
        return losses[0], losses[1]

    @staticmethod
    def backward(ctx, g_l1, g_bce):
        d_after, d_before, d_logits = ctx.saved_tensors
        return d_after * g_l1, d_before * g_l1, d_logits * g_bce, None, None, None, None


class Seq2SeqLoss(torch.nn.Module):
    """Drop-in for seq2seq_vc.losses.Seq2SeqLoss: one fused pass gives both losses and their gradients."""

    def __init__(self, use_masking=True, use_weighted_masking=False, bce_pos_weight=10.0):
        super().__init__()
        assert (use_masking != use_weighted_masking) or not use_masking
        if not use_masking or use_weighted_masking:
            raise NotImplementedError("only use_masking=True (the reference default and every shipped recipe)")
        self.bce_pos_weight = float(bce_pos_weight)

    def forward(self, after_outs, before_outs, logits, ys, labels, olens):
        olens_dev = torch.as_tensor(_host_lens(olens), dtype=_i32).to(after_outs.device) if not (
            isinstance(olens, torch.Tensor) and olens.is_cuda) else olens.to(_i32)
        c = lambda t: t.to(_f32).contiguous()
        return _Seq2SeqLossFn.apply(c(after_outs), c(before_outs), c(logits), c(ys), c(labels), olens_dev, self.bce_pos_weight)


class _GuidedAttnFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, att, ilens_dev, olens_dev, sigma, alpha):
        B, H, T_out, T_in = att.shape
        loss = torch.empty(1, dtype=_f32, device=att.device)
        ws = torch.empty(2, dtype=_f32, device=att.device)
        d_att = torch.empty_like(att)
        ops.guided_attn_loss(att, ilens_dev, olens_dev, T_in, sigma, alpha, loss, d_att, ws)
        ctx.save_for_backward(d_att)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        (d_att,) = ctx.saved_tensors
        return d_att * g, None, None, None, None


class GuidedMultiHeadAttentionLoss(torch.nn.Module):
    """Drop-in for losses.GuidedMultiHeadAttentionLoss: att_ws (B, H, T_out, T_in) -> scalar."""

    def __init__(self, sigma=0.4, alpha=1.0, reset_always=True):
        super().__init__()
        self.sigma, self.alpha = float(sigma), float(alpha)

    def forward(self, att_ws, ilens, olens):
        dev = att_ws.device
        il = torch.as_tensor(_host_lens(ilens), dtype=_i32).to(dev)
        ol = torch.as_tensor(_host_lens(olens), dtype=_i32).to(dev)
        if att_ws.dim() == 3:
            att_ws = att_ws.unsqueeze(1)
        return _GuidedAttnFn.apply(att_ws.to(_f32).contiguous(), il, ol, self.sigma, self.alpha)


GuidedAttentionLoss = GuidedMultiHeadAttentionLoss


# =================================================================================================
# monotonic alignment search (operator seam `self.viterbi_func`, aas_vc.py:132,402-404)
# =================================================================================================
class _ViterbiFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, log_p_attn, text_lens_dev, feats_lens_dev):
        paths, ds, bin_loss, d_log_p = ops.mas(log_p_attn, text_lens_dev, feats_lens_dev, want_grad=True)
        ctx.save_for_backward(d_log_p)
        ctx.mark_non_differentiable(ds)
        return ds, bin_loss[0], paths

    @staticmethod
    def backward(ctx, g_ds, g_bin, g_paths):
        (d_log_p,) = ctx.saved_tensors
        return d_log_p * g_bin, None, None


def viterbi_decode(log_p_attn, text_lengths, feats_lengths, return_paths: bool = False):
    """(B, T_feats, T_text) float32 -> (ds float32 (B, T_text), bin_loss 0-d); bit-exact paths/durations."""
    dev = log_p_attn.device
    tl = torch.as_tensor(_host_lens(text_lengths), dtype=_i32).to(dev)
    fl = torch.as_tensor(_host_lens(feats_lengths), dtype=_i32).to(dev)
    ds, bin_loss, paths = _ViterbiFn.apply(log_p_attn.to(_f32).contiguous(), tl, fl)
    return (ds, bin_loss, paths) if return_paths else (ds, bin_loss)


# =================================================================================================
# STFT -> log-mel (bin/preprocess.py:30-92)
# =================================================================================================
def _slaney_mel_points(n: int, fmin: float, fmax: float) -> np.ndarray:
    """n mel-spaced frequencies (Hz) between fmin and fmax on the Slaney scale (linear < 1 kHz, log above)."""
    f_sp, brk = 200.0 / 3.0, 1000.0
    brk_mel, step = brk / f_sp, math.log(6.4) / 27.0

    def to_mel(f):
        return brk_mel + math.log(f / brk) / step if f >= brk else f / f_sp

    m = np.linspace(to_mel(fmin), to_mel(fmax), n)
    return np.where(m >= brk_mel, brk * np.exp(step * (m - brk_mel)), f_sp * m)


_BASIS_CACHE: Dict[tuple, np.ndarray] = {}


def mel_filterbank(sr: int, n_fft: int, n_mels: int, fmin: float, fmax: float) -> np.ndarray:
    """Slaney-normalised triangular filters on the rFFT bin frequencies, float32 (n_mels, 1 + n_fft/2)."""
    key = (sr, n_fft, n_mels, float(fmin), float(fmax))
    if key not in _BASIS_CACHE:
        edges = _slaney_mel_points(n_mels + 2, fmin, fmax)
        bins = np.arange(1 + n_fft // 2, dtype=np.float64) * (sr / n_fft)
        lo, ce, hi = edges[:-2, None], edges[1:-1, None], edges[2:, None]
        tri = np.minimum((bins[None] - lo) / (ce - lo), (hi - bins[None]) / (hi - ce))
        tri = np.maximum(tri, 0.0) * (2.0 / (hi - lo))
        _BASIS_CACHE[key] = tri.astype(np.float32)
    return _BASIS_CACHE[key]


def hann_window(n_fft: int, win_length: Optional[int]) -> np.ndarray:
    wl = n_fft if win_length is None else int(win_length)
    w = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(wl) / wl)
    out = np.zeros(n_fft, dtype=np.float32)
    off = (n_fft - wl) // 2
    out[off:off + wl] = w
    return out


def logmel_batch(wav: torch.Tensor, sampling_rate: int, fft_size=1024, hop_size=256, win_length=None, num_mels=80,
                 fmin=None, fmax=None, eps=1e-10, log_base=10.0, out: Optional[torch.Tensor] = None, mean=None, scale=None) -> torch.Tensor:
    """Batched device entry: wav (B, n_samples) float32 CUDA -> (B, 1 + n_samples // hop, num_mels) float32.
    mean / scale (num_mels,): the global mean-variance normalisation of bin/normalize.py:173-193 (StandardScaler.transform with
    the statistics of bin/compute_statistics.py) fused into the kernel's store -- the model then sees the kernel's output directly."""
    fmin = 0.0 if fmin is None else fmin
    fmax = sampling_rate / 2.0 if fmax is None else fmax
    dev = wav.device
    key = (dev, sampling_rate, fft_size, win_length, num_mels, fmin, fmax)
    c = _DEV_CACHE.get(key)
    if c is None:
        c = (torch.from_numpy(hann_window(fft_size, win_length)).to(dev),
             torch.from_numpy(mel_filterbank(sampling_rate, fft_size, num_mels, fmin, fmax)).to(dev))
        _DEV_CACHE[key] = c
    B, ns = wav.shape
    if out is None:
        out = torch.empty(B, 1 + ns // hop_size, num_mels, dtype=_f32, device=dev)
    if mean is not None:
        mean = torch.as_tensor(mean, dtype=_f32).to(dev).contiguous()
        scale = torch.as_tensor(scale, dtype=_f32).to(dev).contiguous()
    return ops.logmel(wav, c[0], c[1], out, fft_size, hop_size, eps, log_base, mean, scale)


_DEV_CACHE: Dict[tuple, tuple] = {}


def logmelfilterbank(audio, sampling_rate, fft_size=1024, hop_size=256, win_length=None, window="hann", num_mels=80,
                     fmin=None, fmax=None, eps=1e-10, log_base=10.0):
    """Reference signature (preprocess.py:30-42): 1-D numpy audio -> ndarray (frames, num_mels)."""
    if window != "hann":
        raise NotImplementedError("only the hann window (every shipped recipe) is implemented")
    if log_base not in (None, 2.0, 10.0):
        raise ValueError(f"{log_base} is not supported.")
    wav = torch.from_numpy(np.ascontiguousarray(audio, dtype=np.float32)).cuda().unsqueeze(0)
    mel = logmel_batch(wav, sampling_rate, fft_size, hop_size, win_length, num_mels, fmin, fmax, eps, log_base)
    return mel[0].cpu().numpy()


# =================================================================================================
# Conv2dSubsampling2 / 6 / 8 (modules/transformer/subsampling.py:108-279): stand-alone drop-in operators
# =================================================================================================
class _Holder(torch.nn.Module):
    """Parameter container that reproduces the reference's `conv.N.*` / `out.0.*` state-dict keys."""


class _PositionalEncoding(torch.nn.Module):
    """Default `out.1` (layers/positional_encoding.py:14-70 PositionalEncoding): dropout(x * sqrt(d) + pe[:T]) through s2s_scaled_pe_fwd / _bwd."""

    def __init__(self, d_model, dropout_rate, max_len=5000):
        super().__init__()
        from .engine_base import sinusoid_table

        self.d_model, self.rate, self.max_len = d_model, float(dropout_rate), max_len
        self.xscale = math.sqrt(d_model)
        self._table = sinusoid_table(max_len, d_model, torch.device("cpu"))
        self._seed = 0

    def forward(self, x):
        return _PosEncFn.apply(self, x)


class _PosEncFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, x):
        B, T, d = x.shape
        if T > mod.max_len:
            raise S2SError(f"sequence of {T} frames exceeds the positional table ({mod.max_len})")
        pe = mod._table.to(x.device)
        if pe is not mod._table:
            mod._table = pe
        mod._seed += 1
        drop = _lib.Drop(mod.rate, seed=mod._seed, site=0) if (mod.training and mod.rate > 0) else _lib.NO_DROP
        xs = torch.empty_like(x)
        ops.scale_dropout(x.contiguous(), xs, mod.xscale)
        one = torch.ones(1, dtype=_f32, device=x.device)
        y = torch.empty_like(x)
        ops.scaled_pe_fwd(xs, pe, one, y, drop)
        ctx.mod, ctx.drop, ctx.pe = mod, drop, pe
        return y

    @staticmethod
    def backward(ctx, g):
        dx = torch.empty_like(g)
        dalpha = torch.zeros(1, dtype=_f32, device=g.device)
        ops.scaled_pe_bwd(g.contiguous(), ctx.pe, dx, dalpha, ctx.drop)
        out = torch.empty_like(dx)
        ops.scale_dropout(dx, out, ctx.mod.xscale)
        return None, out


class _Conv2dSubFn(torch.autograd.Function):
    """Conv2d(1 -> C, 3, 2) + ReLU, then Conv2d(C -> C, k, s) + ReLU per entry of `mod.LATER`, then Linear(C * F -> odim): the first
    convolution by the direct kernel, the later ones as patch matrix x weights on the GEMM path (ReLU in the epilogue), every
    parameter gradient by the hand-written backward (the input gets none: it is a feature tensor, as in the training path)."""

    @staticmethod
    def forward(ctx, mod, x, *params):
        adt, mode = mod._adt, mod._mode
        B, T, idim = x.shape
        C = params[0].shape[0]
        x = x.to(_f32).contiguous()
        dev = x.device
        T1, F1 = (T - 1) // 2, (idim - 1) // 2
        if T1 < 1 or F1 < 1:
            raise S2SError("input too short for the first convolution")
        maps = [torch.empty(B, T1, F1, C, dtype=adt, device=dev)]
        ops.conv1_fwd(x, params[0].contiguous(), params[1], maps[0])
        wps = []
        for i, (k, s) in enumerate(mod.LATER):
            w, b = params[2 + 2 * i], params[3 + 2 * i]
            Tp, Fp = maps[-1].shape[1], maps[-1].shape[2]
            if Tp < k or Fp < k:
                raise S2SError("input too short for this subsampling module")
            Tn, Fn = (Tp - k) // s + 1, (Fp - k) // s + 1
            wp = torch.empty(C, k * k, C, dtype=adt, device=dev)
            ops.transpose_last2(w.detach().contiguous(), wp, C, C, k * k)          # fp32 parameter -> compute dtype
            col = torch.empty(B * Tn * Fn, k * k * C, dtype=adt, device=dev)
            ops.im2col2d(maps[-1], col, k, s)
            y = torch.empty(B * Tn * Fn, C, dtype=adt, device=dev)
            ops.gemm(col, wp.view(C, k * k * C), y, bias=b, relu=True, mode=mode)
            maps.append(y.view(B, Tn, Fn, C))
            wps.append(wp)
        Tn, Fn = maps[-1].shape[1], maps[-1].shape[2]
        wo, bo = params[-2], params[-1]
        odim = wo.shape[0]
        wop = torch.empty(odim, Fn, C, dtype=adt, device=dev)
        ops.transpose_last2(wo.detach().contiguous(), wop, odim, C, Fn)
        lin = torch.empty(B * Tn, odim, dtype=adt, device=dev)
        ops.gemm(maps[-1].view(B * Tn, Fn * C), wop.view(odim, Fn * C), lin, bias=bo, mode=mode)
        ctx.mod, ctx.x, ctx.maps, ctx.wps, ctx.wop = mod, x, maps, wps, wop
        ctx.shapes = [tuple(p.shape) for p in params]
        return lin.view(B, Tn, odim).float()

    @staticmethod
    def backward(ctx, g):
        mod, x, maps, wps, wop = ctx.mod, ctx.x, ctx.maps, ctx.wps, ctx.wop
        adt, mode = mod._adt, mod._mode
        dev = x.device
        B = x.shape[0]
        C = maps[0].shape[3]
        Tn, Fn = maps[-1].shape[1], maps[-1].shape[2]
        odim = wop.shape[0]
        grads = [torch.zeros(s, dtype=_f32, device=dev) for s in ctx.shapes]
        dlin = g.reshape(B * Tn, odim).to(adt).contiguous()
        ylast = maps[-1].view(B * Tn, Fn * C)
        gwop = torch.empty(odim, Fn * C, dtype=_f32, device=dev)
        ops.gemm(dlin.t(), ylast.t(), gwop, mode=mode)
        ops.transpose_last2(gwop, grads[-2].view(odim, C, Fn), odim, Fn, C)
        ops.colsum(dlin, grads[-1])
        dy = torch.empty(B * Tn, Fn * C, dtype=adt, device=dev)
        ops.gemm(dlin, wop.view(odim, Fn * C).t(), dy, mode=mode, gate=ylast)            # ReLU' of the last convolution
        dy = dy.view(B * Tn * Fn, C)
        for i in reversed(range(len(mod.LATER))):
            k, s = mod.LATER[i]
            prev = maps[i]
            col = torch.empty(dy.shape[0], k * k * C, dtype=adt, device=dev)
            ops.im2col2d(prev, col, k, s)
            gwp = torch.empty(C, k * k * C, dtype=_f32, device=dev)
            ops.gemm(dy.t(), col.t(), gwp, mode=mode)
            ops.transpose_last2(gwp, grads[2 + 2 * i].view(C, C, k * k), C, k * k, C)
            ops.colsum(dy, grads[3 + 2 * i])
            ops.gemm(dy, wps[i].view(C, k * k * C).t(), col, mode=mode)                   # d(patches), in place of the patches
            dprev = torch.empty_like(prev)
            ops.col2im2d(col, prev, dprev, k, s)                                            # scatter-add + ReLU' of the map below
            dy = dprev.view(-1, C)
        ops.conv1_bwd(x, dy.view(maps[0].shape), grads[0], grads[1])
        return (None, None) + tuple(grads)


class _Conv2dSubsamplingN(torch.nn.Module):
    LATER: tuple = ()
    MASKS: tuple = ()

    def __init__(self, idim, odim, dropout_rate, pos_enc=None, compute_dtype: str = "float32"):
        super().__init__()
        self._adt = torch.bfloat16 if compute_dtype in ("bf16", "bfloat16", torch.bfloat16) else _f32
        self._mode = 1 if self._adt == torch.bfloat16 else (0 if compute_dtype == "float32_simt" else 2)
        f = (idim - 1) // 2
        self.conv = _Holder()
        convs = [torch.nn.Conv2d(1, odim, 3, 2)] + [torch.nn.Conv2d(odim, odim, k, s) for k, s in self.LATER]   # reference initialisation
        for i, c in enumerate(convs):
            h = _Holder()
            h.weight, h.bias = c.weight, c.bias
            self.conv.add_module(str(2 * i), h)
        for k, s in self.LATER:
            f = (f - k) // s + 1
        lin = torch.nn.Linear(odim * f, odim)
        h = _Holder()
        h.weight, h.bias = lin.weight, lin.bias
        self.out = _Holder()
        self.out.add_module("0", h)
        self.out.add_module("1", pos_enc if pos_enc is not None else _PositionalEncoding(odim, dropout_rate))

    def _params(self):
        ps = []
        for i in range(1 + len(self.LATER)):
            h = getattr(self.conv, str(2 * i))
            ps += [h.weight, h.bias]
        h = getattr(self.out, "0")
        return ps + [h.weight, h.bias]

    def forward(self, x, x_mask):
        _require_cuda(x, type(self).__name__)
        y = _Conv2dSubFn.apply(self, x, *self._params())
        y = getattr(self.out, "1")(y)
        if x_mask is None:
            return y, None
        for sl in self.MASKS:
            x_mask = x_mask[:, :, sl]
        return y, x_mask

    def __getitem__(self, key):
        if key != -1:
            raise NotImplementedError("Support only `-1` (for `reset_parameters`).")
        return getattr(self.out, "1")


class Conv2dSubsampling2(_Conv2dSubsamplingN):
    """Drop-in for subsampling.py:108-165 (time / 2): Conv2d(1, C, 3, 2) + ReLU + Conv2d(C, C, 3, 1) + ReLU + Linear + pos_enc."""
    LATER = ((3, 1),)
    MASKS = (slice(None, -2, 2), slice(None, -2, 1))


class Conv2dSubsampling6(_Conv2dSubsamplingN):
    """Drop-in for subsampling.py:167-213 (time / 6): Conv2d(1, C, 3, 2) + ReLU + Conv2d(C, C, 5, 3) + ReLU + Linear + pos_enc."""
    LATER = ((5, 3),)
    MASKS = (slice(None, -2, 2), slice(None, -4, 3))


class Conv2dSubsampling8(_Conv2dSubsamplingN):
    """Drop-in for subsampling.py:215-279 (time / 8): three Conv2d(., C, 3, 2) + ReLU, Linear, pos_enc."""
    LATER = ((3, 2), (3, 2))
    MASKS = (slice(None, -2, 2), slice(None, -2, 2), slice(None, -2, 2))


class FeatureStatistics:
    """Global mean / scale of a feature set, accumulated on the device: what bin/compute_statistics.py:128-152 obtains from
    sklearn's ``StandardScaler().partial_fit(mel)`` over every utterance -- same attribute names (``mean_``, ``var_``, ``scale_``,
    ``n_samples_seen_``), float64, population variance, zero variance -> scale 1 -- so that ``stats()`` is the (2, D) float32
    array the reference saves as ``stats.npy`` and ``logmel_batch(..., mean=, scale=)`` / bin/normalize.py consume.
    ``partial_fit`` takes one utterance (T, D) like the reference loop, or a zero-padded batch (B, T, D) with ``lens``
    (e.g. the output of ``logmel_batch`` with its per-clip frame counts), numpy or CUDA tensors."""

    def __init__(self, device="cuda:0"):
        self.device = torch.device(device)
        self._acc = None

    def partial_fit(self, feats, lens=None):
        x = torch.as_tensor(np.ascontiguousarray(feats, dtype=np.float32) if isinstance(feats, np.ndarray) else feats)
        x = x.to(device=self.device, dtype=_f32)
        if x.dim() == 2:
            x = x.unsqueeze(0)
        if x.dim() != 3:
            raise ValueError("feats must be (T, D) or (B, T, D)")
        x = x.contiguous()
        D = x.shape[2]
        if self._acc is None:
            self._acc = torch.zeros(2 * D + 1, dtype=torch.float64, device=self.device)
        elif self._acc.numel() != 2 * D + 1:
            raise ValueError(f"feature dimension changed: {(self._acc.numel() - 1) // 2} -> {D}")
        if lens is not None:
            lens = torch.as_tensor(lens).to(device=self.device, dtype=torch.int32).contiguous()
            if lens.numel() != x.shape[0]:
                raise ValueError("lens must hold one length per utterance")
        ops.feat_stats(x, lens, self._acc)
        return self

    def _host(self):
        if self._acc is None:
            raise S2SError("FeatureStatistics: partial_fit() has not been called")
        a = self._acc.cpu().numpy()
        D = (a.size - 1) // 2
        return a[:D], a[D:2 * D], a[2 * D]

    @property
    def n_samples_seen_(self) -> int:
        return int(self._host()[2])

    @property
    def mean_(self) -> np.ndarray:
        s1, _, n = self._host()
        return s1 / n

    @property
    def var_(self) -> np.ndarray:
        s1, s2, n = self._host()
        return np.maximum(s2 / n - (s1 / n) ** 2, 0.0)

    @property
    def scale_(self) -> np.ndarray:
        s1, _, n = self._host()
        v, m = self.var_, s1 / n
        eps = np.finfo(np.float64).eps
        sc = np.sqrt(v)
        # sklearn (_is_constant_feature + _handle_zeros_in_scale): features whose variance is at round-off level are left unscaled
        sc[(v <= n * eps * v + (n * m * eps) ** 2) | (sc == 0.0)] = 1.0
        return sc

    def stats(self) -> np.ndarray:
        """(2, D) float32 [mean_, scale_]: the array compute_statistics.py writes to stats.npy."""
        return np.stack([self.mean_, self.scale_], axis=0).astype(np.float32)


# =================================================================================================
# Griffin-Lim vocoder path (vocoder/griffin_lim.py:20-222): the inverse of the log-mel front end, same mel basis and window
# =================================================================================================
GL_EPS = 1e-10


def logmel2linear(lmspc, fs, n_fft, n_mels, fmin=None, fmax=None) -> np.ndarray:
    """griffin_lim.py:20-50: log-mel (T, n_mels) -> linear magnitudes (T, n_fft // 2 + 1) = max(EPS, pinv(mel_basis) @ 10 ** lmspc).
    The pseudo-inverse of the (n_mels x bins) basis is a one-off host computation (numpy, as in the reference); the product of
    every frame with it runs through s2s_gemm."""
    lmspc = np.asarray(lmspc)
    assert lmspc.shape[1] == n_mels
    fmin = 0 if fmin is None else fmin
    fmax = fs / 2 if fmax is None else fmax
    key = ("pinv", fs, n_fft, n_mels, fmin, fmax)
    inv = _DEV_CACHE.get(key)
    if inv is None:
        inv = torch.from_numpy(np.linalg.pinv(mel_filterbank(fs, n_fft, n_mels, fmin, fmax)).astype(np.float32)).cuda().contiguous()
        _DEV_CACHE[key] = inv                                         # (bins, n_mels)
    mspc = torch.from_numpy(np.power(10.0, lmspc).astype(np.float32)).cuda().contiguous()
    out = torch.empty(mspc.shape[0], inv.shape[0], dtype=_f32, device=mspc.device)
    ops.gemm(mspc, inv, out, mode=0)                                  # (T, n_mels) x (bins, n_mels)^T, fp32 on the CUDA cores
    return np.maximum(GL_EPS, out.cpu().numpy())


def griffin_lim(spc, n_fft, n_shift, win_length=None, window="hann", n_iter=32, init_angles=None, momentum=0.99,
                pad_mode="constant", seed=None) -> np.ndarray:
    """griffin_lim.py:52-106 (-> librosa.griffinlim, center = True): linear magnitudes (T, n_fft // 2 + 1) -> waveform
    (n_shift * (T - 1),).  The whole iteration runs on the device (s2s_gl_istft / s2s_gl_stft / s2s_gl_update).  librosa draws the
    initial phases from numpy's global generator; here they come from `init_angles` (T, bins) complex, or from
    numpy.random.default_rng(seed)."""
    if window != "hann":
        raise NotImplementedError("only the hann window (every shipped recipe) is implemented")
    if pad_mode not in ("constant", "reflect"):
        raise ValueError(f"pad_mode {pad_mode!r}")
    S = np.abs(np.asarray(spc, dtype=np.float32))
    T, bins = S.shape
    assert bins == n_fft // 2 + 1 and T >= 2
    if init_angles is None:
        init_angles = np.exp(2j * np.pi * np.random.default_rng(seed).random(S.shape))
    ang = np.asarray(init_angles, dtype=np.complex64)
    dev = torch.device("cuda", torch.cuda.current_device())
    mag = torch.from_numpy(S).to(dev).contiguous()
    angles = torch.view_as_real(torch.from_numpy(ang).to(dev)).contiguous()
    win = torch.from_numpy(hann_window(n_fft, win_length)).to(dev)
    frames = torch.empty(T, n_fft, dtype=_f32, device=dev)
    y = torch.empty(n_shift * (T - 1), dtype=_f32, device=dev)
    rebuilt = torch.empty(T, bins, 2, dtype=_f32, device=dev)
    tprev = torch.zeros(T, bins, 2, dtype=_f32, device=dev)
    c = momentum / (1.0 + momentum)
    for _ in range(n_iter):
        ops.gl_istft(mag, angles, win, frames, y, n_fft, n_shift)
        ops.gl_stft(y, win, rebuilt, n_fft, n_shift, pad_mode == "reflect")
        ops.gl_update(rebuilt, tprev, angles, c)
    ops.gl_istft(mag, angles, win, frames, y, n_fft, n_shift)
    return y.cpu().numpy()


class Spectrogram2Waveform:
    """vocoder/griffin_lim.py:110-222: log-mel (or linear) spectrogram -> waveform by (pseudo-inverse mel basis +) Griffin-Lim;
    same constructor arguments, `decode(spc)` takes and returns torch tensors."""

    def __init__(self, n_fft, n_shift, stats=None, fs=None, n_mels=None, win_length=None, window="hann", fmin=None, fmax=None,
                 griffin_lim_iters=8, take_norm_feat=True):
        self.take_norm_feat = take_norm_feat
        self.stats = stats
        if self.take_norm_feat:
            assert self.stats is not None, "must specify stats if take_norm_feat=True."
        self.fs = fs
        self.n_mels = n_mels
        self.params = dict(n_fft=n_fft, n_shift=n_shift, win_length=win_length, window=window, n_iter=griffin_lim_iters)
        self.mel_params = dict(fs=fs, n_fft=n_fft, n_mels=n_mels, fmin=fmin, fmax=fmax)
        if n_mels is not None:
            self.params.update(fs=fs, n_mels=n_mels, fmin=fmin, fmax=fmax)

    def __repr__(self):
        return f"{self.__class__.__name__}(" + "".join(f"{k}={v}, " for k, v in self.params.items()) + ")"

    def decode(self, spc: torch.Tensor, init_angles=None, seed=None) -> torch.Tensor:
        device, dtype = spc.device, spc.dtype
        spc = spc.detach().cpu().numpy()
        if self.take_norm_feat:
            spc = spc * self.stats["scale"] + self.stats["mean"]
        if self.n_mels is not None:
            spc = logmel2linear(spc, **self.mel_params)
        gl = {k: self.params[k] for k in ("n_fft", "n_shift", "win_length", "window", "n_iter")}
        wav = griffin_lim(spc, init_angles=init_angles, seed=seed, **gl)
        return torch.tensor(wav).to(device=device, dtype=dtype)

    def __call__(self, spc):
        return self.decode(spc)


# =================================================================================================
# fused training step (trainers/ar_vc.py:59-112 without the host round trips)
# =================================================================================================
class _ReferenceCheckpoint:
    """Checkpoints of the fused steps in the reference's layout (trainers/base.py:85-121):
    ``{"model": state_dict, "optimizer": torch.optim.Adam state dict, "scheduler": WarmupLR state dict, "steps", "epochs"}``
    so a run can move between the reference trainer and the fused step in either direction (`--resume`).  The Adam moments
    live in the engine's flat M / V buffers; the optimizer state is keyed by parameter ORDER, which the drop-in modules
    register exactly as the reference does.  Needs the drop-in module (not a bare engine) for that order."""

    # ---- input staging: the next batch's host -> device copy overlapped with the running step ------------------------------
    _stage = None

    def prefetch(self, *host_tensors) -> None:
        """Start copying the NEXT batch's pinned host tensors (the same objects, in the order the step takes them: xs, ys, labels
        for VTNTrainStep; xs, ys, dp_inputs for AASVCTrainStep) to the device on a side stream.  The next __call__ that is handed
        these tensors moves the staged copies into place device-to-device instead of waiting for PCIe, so with CUDA graphs the
        copy of batch n + 1 runs under the kernels of batch n (what a pin_memory DataLoader worker gives the reference trainer).
        Optional: without it __call__ copies on the compute stream as before."""
        if any(t.is_cuda for t in host_tensors):
            return
        st = self._stage
        if st is None:
            st = self._stage = {"stream": torch.cuda.Stream(), "ready": torch.cuda.Event(), "consumed": None, "bufs": {}, "key": None}
        shape_key = tuple((tuple(t.shape), t.dtype) for t in host_tensors)
        bufs = st["bufs"].get(shape_key)
        if bufs is None:
            bufs = st["bufs"][shape_key] = [torch.empty(t.shape, dtype=t.dtype, device=self.engine.device) for t in host_tensors]
            if len(st["bufs"]) > 8:
                st["bufs"].pop(next(iter(st["bufs"])))
        side = st["stream"]
        if st["consumed"] is not None:
            side.wait_event(st["consumed"])            # the previous staged batch has been moved out of these buffers
        with torch.cuda.stream(side):
            for b, t in zip(bufs, host_tensors):
                b.copy_(t, non_blocking=True)
            st["ready"].record(side)
        st["key"] = tuple(t.data_ptr() for t in host_tensors) + (shape_key,)
        st["cur"] = bufs

    def _take_staged(self, *tensors):
        """The device copies `prefetch` staged for exactly these host tensors (the compute stream then waits for the copy), or None."""
        st = self._stage
        if st is None or st["key"] is None:
            return None
        key = tuple(t.data_ptr() for t in tensors) + (tuple((tuple(t.shape), t.dtype) for t in tensors),)
        if key != st["key"]:
            return None
        st["key"] = None
        torch.cuda.current_stream().wait_event(st["ready"])
        return st["cur"]

    def _staged_consumed(self) -> None:
        st = self._stage
        if st is not None:
            ev = torch.cuda.Event()
            ev.record()
            st["consumed"] = ev

    def _on_evict(self, sig) -> None:
        """The engine dropped the activation buffers of batch shape `sig` (least recently used): graphs captured for that
        shape address freed memory and go with them."""
        for key in [k for k in self._graphs if tuple(k[:3]) == tuple(sig[:3])]:
            del self._graphs[key]

    def _named_params(self):
        if self._model is None:
            raise S2SError("state_dict() / load_state_dict() need the drop-in module (VTN / TransformerTTS / AASVC), not a bare engine")
        return [n for n, _ in self._model.named_parameters()]

    def _clock_of(self, name: str) -> torch.Tensor:
        return self.engine.step_dev

    def state_dict(self, epochs: int = 0) -> dict:
        eng, st = self.engine, self.engine.store
        names = self._named_params()
        skeleton = torch.optim.Adam([torch.nn.Parameter(torch.zeros(1)) for _ in names], lr=self.lr, betas=tuple(self.betas),
                                    eps=self.eps, weight_decay=self.wd).state_dict()["param_groups"]
        next_lr = self.lr_at(self.steps + 1)              # what scheduler.step() leaves in the group for the next step
        skeleton[0].update(lr=next_lr, initial_lr=self.lr)
        state = {}
        for i, n in enumerate(names):
            t = float(self._clock_of(n))
            if t > 0:                                      # torch keeps no state for parameters that never had a gradient
                state[i] = {"step": torch.tensor(t), "exp_avg": st.m(n).detach().clone(),
                            "exp_avg_sq": st.v(n).detach().clone()}
        sched = {"warmup_steps": self.warmup, "base_lrs": [self.lr], "last_epoch": self.steps, "_step_count": self.steps + 1,
                 "_is_initial": False, "_get_lr_called_within_step": False, "_last_lr": [next_lr]}
        return {"model": eng.state_dict(), "optimizer": {"state": state, "param_groups": skeleton}, "scheduler": sched,
                "steps": self.steps, "epochs": epochs}

    def load_state_dict(self, ckpt: dict, load_only_params: bool = False) -> None:
        eng, st = self.engine, self.engine.store
        eng.load_state_dict(ckpt["model"])
        if load_only_params:
            return
        names = self._named_params()
        grp = ckpt["optimizer"]["param_groups"]
        if len(grp) != 1 or len(grp[0]["params"]) != len(names):
            raise S2SError("optimizer state does not match this model (one param group over every parameter, in order)")
        self.lr = float(grp[0].get("initial_lr", ckpt["scheduler"].get("base_lrs", [self.lr])[0]))
        self.betas, self.eps, self.wd = tuple(grp[0]["betas"]), float(grp[0]["eps"]), float(grp[0]["weight_decay"])
        self.warmup = int(ckpt["scheduler"].get("warmup_steps", self.warmup))
        st.M.zero_()
        st.V.zero_()
        clocks = {}
        for i, n in enumerate(names):
            ent = ckpt["optimizer"]["state"].get(grp[0]["params"][i])
            if ent is None:
                continue
            st.m(n).copy_(ent["exp_avg"].to(eng.device, _f32))
            st.v(n).copy_(ent["exp_avg_sq"].to(eng.device, _f32))
            clk = self._clock_of(n)
            t = float(ent["step"])
            if clocks.setdefault(id(clk), t) != t:
                raise S2SError(f"parameters sharing one Adam clock carry different step counts ({n}: {t})")
            clk.fill_(t)
        self.steps = int(ckpt["steps"])
        if hasattr(self, "backward_steps"):
            self.backward_steps = self.steps * self.accum
        self._graphs.clear()


class VTNTrainStep(_ReferenceCheckpoint):
    """forward + Seq2SeqLoss + backward (+ gradient all-reduce) + clip + Adam + WarmupLR, device-resident.

    Mirrors ARVCTrainer._train_step (trainers/ar_vc.py:59-112) for a VTN model.  Lengths arrive as
    host ints (the collater's CPU tensors) so the step never synchronises.  With ``use_graph`` the
    launch sequence of one (B, T, L) batch shape is captured once into two CUDA graphs
    (forward+loss+backward | clip+Adam) and replayed; the NCCL gradient all-reduce of the flat
    gradient buffer runs between them (the path shards by utterance batch, bin/vc_train.py:423-431).
    """

    def __init__(self, model, lr: float = 8e-5, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 grad_norm: float = 1.0, warmup_steps: int = 4000, bce_pos_weight: float = 10.0, use_graph: bool = False,
                 process_group=None, guided_attn: Optional[dict] = None):
        self.engine: VTNEngine = model.engine if hasattr(model, "engine") else model
        self._model = model if hasattr(model, "engine") else None
        self.lr, self.betas, self.eps, self.wd = lr, betas, eps, weight_decay
        self.grad_norm, self.warmup = grad_norm, warmup_steps
        self.pos_weight = bce_pos_weight
        # guided_attn: dict(sigma=0.4, alpha=1.0, n_layers=2, n_heads=2) adds GuidedMultiHeadAttentionLoss (trainers/ar_tts.py:49-53)
        self.guided_attn = guided_attn
        self.ga_loss = None
        self.steps = 0
        self.use_graph = use_graph
        self.pg = process_group
        self.world = 1
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            self.world = torch.distributed.get_world_size(process_group)
        self._graphs: Dict[tuple, tuple] = {}
        self.replayed_launches = 0      # kernels of this library launched through graph replays
        self.engine._evict_listeners.append(weakref.WeakMethod(self._on_evict))
        if guided_attn is not None:     # the guided-attention loss reads (and back-propagates into) these maps: keep them in HBM
            nl = self.engine.hp["dlayers"]
            self.engine.attn_emit_names = frozenset(
                f"decoder.decoders.{l}.src_attn" for l in list(reversed(range(nl)))[:guided_attn.get("n_layers", 2)])
        self.emit_attention = False     # True: the fused step also writes the source-attention maps (engine.attn) for monitoring
        # Bucketed overlap of the gradient all-reduce with the encoder's backward (two graphs, the decoder-side 3/4 of the flat
        # buffer reduced on NCCL's stream meanwhile).  Measured at 2 GPUs on C2 (profiles/r02_ddp_overlap_ab.txt): 9.95 ms with,
        # 9.83 ms without -- the NCCL kernel takes SMs from the persistent GEMMs it overlaps and the second collective adds its
        # own launch latency, so the single all-reduce between the graphs stays the default.
        self.overlap_allreduce = os.environ.get("S2S_OVERLAP_ALLREDUCE", "0") == "1"
        self._comm_stream = None
        self._cut_event = None

    def lr_at(self, step: int) -> float:
        """WarmupLR (schedulers/warmup_lr.py:54-61): lr * warmup^0.5 * min(step^-0.5, step * warmup^-1.5)."""
        s = max(step, 1)
        return self.lr * self.warmup ** 0.5 * min(s ** -0.5, s * self.warmup ** -1.5)

    # -- the two halves of a step (each is a fixed launch sequence for a given batch shape)
    def _fwd_bwd(self, xs, ys, labels, on_decoder_done=None):
        eng = self.engine
        saved = eng.attn_emit
        if not self.emit_attention:
            eng.attn_emit = "none"      # nothing in the fused step reads an attention map (guided layers: attn_emit_names)
        try:
            eng.forward(xs, ys)
            eng.loss(ys, labels, self.pos_weight)
            d_att = None
            if self.guided_attn is not None:
                self.ga_loss, d_att = eng.guided_attention(**self.guided_attn)
            eng.backward(eng.d_after, eng.d_before, eng.d_logits, d_att=d_att, on_decoder_done=on_decoder_done)
        finally:
            eng.attn_emit = saved

    def _allreduce(self):
        if self.world > 1:
            torch.distributed.all_reduce(self.engine.store.G, group=self.pg)

    def _update(self):
        self.engine.optimizer_step(self.grad_norm, self.betas, self.eps, self.wd, grad_scale=1.0 / self.world)

    def __call__(self, xs, ilens, ys, labels, olens):
        """xs (B,T,idim), ys (B,L,odim), labels (B,L): float32 tensors trimmed to the batch maxima, either
        CUDA-resident or pinned host memory (copied with non-blocking H2D).  Returns the device
        tensor (l1_loss, bce_loss) of this step without synchronising."""
        eng = self.engine
        self.steps += 1
        eng.lr_dev.fill_(self.lr_at(self.steps))
        B, T, L = xs.shape[0], xs.shape[1], ys.shape[1]
        eng.training = True
        eng.prepare(B, T, L, ilens, olens)
        staged = self._take_staged(xs, ys, labels)
        if staged is not None:
            xs, ys, labels = staged
        if not self.use_graph:
            if not xs.is_cuda:
                xs, ys, labels = (t.to(eng.device, non_blocking=True) for t in (xs, ys, labels))
            self._fwd_bwd(xs, ys, labels)
            self._allreduce()
            self._update()
            if staged is not None:
                self._staged_consumed()
            return eng.losses
        key = (B, T, L)
        entry = self._graphs.get(key)
        if entry is None:
            sx = torch.empty(xs.shape, dtype=xs.dtype, device=eng.device)
            sy = torch.empty(ys.shape, dtype=_f32, device=eng.device)
            sl = torch.empty(labels.shape, dtype=_f32, device=eng.device)
            for dst, src in ((sx, xs), (sy, ys), (sl, labels)):
                dst.copy_(src, non_blocking=True)
            if staged is not None:
                self._staged_consumed()
            # one eager step first: allocates every activation buffer outside the graph's private pool
            self._fwd_bwd(sx, sy, sl)
            self._allreduce()
            self._update()
            torch.cuda.synchronize()
            g1, g2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            g1b = None
            n0 = _lib.launch_count()
            if self.world > 1 and self.overlap_allreduce:
                # data parallel: the step is cut where the decoder-side gradients are final (two graphs sharing one pool), so
                # that their all-reduce runs on NCCL's stream while the encoder's backward is still computing
                g1b = torch.cuda.CUDAGraph()
                cap = torch.cuda.Stream()
                cap.wait_stream(torch.cuda.current_stream())
                with _lib.no_gc(), torch.cuda.stream(cap):
                    g1.capture_begin()

                    def cut():
                        g1.capture_end()
                        g1b.capture_begin(pool=g1.pool())

                    self._fwd_bwd(sx, sy, sl, on_decoder_done=cut)
                    g1b.capture_end()
                torch.cuda.current_stream().wait_stream(cap)
            else:
                with _lib.graph_capture(g1):
                    self._fwd_bwd(sx, sy, sl)
            with _lib.graph_capture(g2):
                self._update()
            self._graphs[key] = (g1, g2, sx, sy, sl, _lib.launch_count() - n0, g1b)
            return eng.losses
        g1, g2, sx, sy, sl, n_kernels, g1b = entry
        if eng.p16_dirty:
            eng.sync_shadow()
        self.replayed_launches += n_kernels
        for dst, src in ((sx, xs), (sy, ys), (sl, labels)):
            if src.data_ptr() != dst.data_ptr():
                dst.copy_(src, non_blocking=True)
        if staged is not None:
            self._staged_consumed()         # the staging buffers are free again: the next prefetch may run under this step's kernels
        g1.replay()
        if g1b is None:
            self._allreduce()
        else:
            G, cut = eng.store.G, eng.encoder_span()
            if self._comm_stream is None:
                self._comm_stream, self._cut_event = torch.cuda.Stream(), torch.cuda.Event()
            self._cut_event.record()
            with torch.cuda.stream(self._comm_stream):
                self._comm_stream.wait_event(self._cut_event)
                w_tail = torch.distributed.all_reduce(G[cut:], group=self.pg, async_op=True)      # overlaps the encoder's backward
            g1b.replay()
            w_head = torch.distributed.all_reduce(G[:cut], group=self.pg, async_op=True)
            w_tail.wait()
            w_head.wait()
        g2.replay()
        return eng.losses


# =================================================================================================
# AAS-VC (models/aas_vc.py, trainers/aas_vc.py): drop-in model, losses and the fused training step
# =================================================================================================
from .aasvc_engine import AASVCEngine, beta_binomial_log_prior  # noqa: E402
from .aasvc_engine import default_hparams as aasvc_default_hparams  # noqa: E402


class L1Loss(torch.nn.Module):
    """Drop-in for seq2seq_vc.losses.L1Loss (losses/l1_loss.py:5-49): masked mean-L1(before) + mean-L1(after)."""

    def __init__(self, use_masking=True, reduction="mean"):
        super().__init__()
        if not use_masking or reduction != "mean":
            raise NotImplementedError("only use_masking=True, reduction='mean' (the reference defaults)")

    def forward(self, after_outs, before_outs, ys, olens):
        dev = before_outs.device
        olens_dev = torch.as_tensor(_host_lens(olens), dtype=_i32).to(dev)
        c = lambda t: t.to(_f32).contiguous()
        if after_outs is None:
            raise NotImplementedError("after_outs=None (diffusion decoders) is outside the hot path")
        B, L = before_outs.shape[0], before_outs.shape[1]
        zeros = torch.zeros(B, L, dtype=_f32, device=dev)
        l1, _ = _Seq2SeqLossFn.apply(c(after_outs), c(before_outs), zeros, c(ys), zeros, olens_dev, 1.0)
        return l1


class _ForwardSumFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, log_p_attn, prior, tl, fl, blank_logp):
        loss = torch.zeros(1, dtype=_f32, device=log_p_attn.device)
        grad = torch.empty_like(log_p_attn)
        ops.forward_sum(log_p_attn, prior, tl, fl, torch.empty_like(log_p_attn), loss, grad, 1.0, blank_logp)
        ctx.save_for_backward(grad)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None, None, None, None


class ForwardSumLoss(torch.nn.Module):
    """Drop-in for seq2seq_vc.losses.ForwardSumLoss (losses/forward_sum_loss.py:12-116).  The beta-binomial prior is
    built on the host (float64 lgamma, cached per (T, N)) as the reference does with scipy; the alpha/beta recursions
    of all utterances run in one kernel launch.  The gradient is the one torch's ctc_loss backward gives the reference."""

    def __init__(self, cache_prior: bool = True):
        super().__init__()
        self.cache_prior = cache_prior
        self._cache: Dict[tuple, torch.Tensor] = {}

    def forward(self, log_p_attn, ilens, olens, blank_prob: float = math.e ** -1):
        il, ol = _host_lens(ilens), _host_lens(olens)
        B, TF, TT = log_p_attn.shape
        dev = log_p_attn.device
        prior = torch.zeros(B, TF, TT, dtype=_f32)
        for b in range(B):
            key = (ol[b], il[b])
            tab = self._cache.get(key)
            if tab is None:
                tab = beta_binomial_log_prior(il[b], ol[b])
                if self.cache_prior:
                    self._cache[key] = tab
            prior[b, :ol[b], :il[b]] = tab
        tl = torch.tensor(il, dtype=_i32).to(dev)
        fl = torch.tensor(ol, dtype=_i32).to(dev)
        return _ForwardSumFn.apply(log_p_attn.to(_f32).contiguous(), prior.to(dev), tl, fl, float(math.log(blank_prob)))


class _DurationLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, d_outs, ds, tl, offset):
        loss = torch.zeros(1, dtype=_f32, device=d_outs.device)
        grad = torch.empty_like(d_outs)
        ops.duration_loss(d_outs, ds, tl, None, loss, grad, 1.0, offset, float("inf"))
        ctx.save_for_backward(grad)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None, None, None


class DurationPredictorLoss(torch.nn.Module):
    """Drop-in for seq2seq_vc.losses.DurationPredictorLoss (losses/duration_predictor_loss.py:5-50)."""

    def __init__(self, use_masking=True, offset=1.0, reduction="mean"):
        super().__init__()
        if not use_masking or reduction != "mean":
            raise NotImplementedError("only use_masking=True, reduction='mean' (the reference defaults)")
        self.offset = float(offset)

    def forward(self, d_outs, ds, ilens):
        tl = torch.tensor(_host_lens(ilens), dtype=_i32).to(d_outs.device)
        return _DurationLossFn.apply(d_outs.to(_f32).contiguous(), ds.to(_f32).contiguous(), tl, self.offset)


class _LengthRegulatorFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xs, cum, Lmax, pad_value):
        y = torch.empty(xs.shape[0], Lmax, xs.shape[2], dtype=xs.dtype, device=xs.device)
        ops.lr_fwd(xs, cum, y, pad_value)
        ctx.save_for_backward(cum)
        ctx.shape = xs.shape
        return y

    @staticmethod
    def backward(ctx, g):
        (cum,) = ctx.saved_tensors
        dx = torch.empty(ctx.shape, dtype=g.dtype, device=g.device)
        ops.lr_bwd(g.contiguous(), cum, dx)
        return dx, None, None, None


class LengthRegulator(torch.nn.Module):
    """modules/length_regulator.py:46-97: repeat every row of xs (B, Tmax, D) ds[b, i] times and pad the ragged result with
    `pad_value` -> (B, max_b sum_i ds[b, i], D).  Same call (`forward(xs, ds, alpha=1.0)`, ds a LongTensor), same rounding of
    `ds * alpha` (torch.round), same rescue when every predicted duration is 0 (all durations become 1).  One host read of the
    output lengths sizes the result, as the reference's pad_list does; the repeat itself is a device gather and its adjoint."""

    def __init__(self, pad_value=0.0):
        super().__init__()
        self.pad_value = pad_value

    def forward(self, xs, ds, alpha=1.0):
        _require_cuda(xs, "LengthRegulator")
        if alpha != 1.0:
            assert alpha > 0
        B, T, _ = xs.shape
        ds = ds.to(torch.int64).contiguous()
        cum = torch.empty(B, T + 1, dtype=torch.int32, device=xs.device)
        ops.lr_cumsum(ds, cum, alpha)
        lens = cum[:, T].tolist()
        if sum(lens) == 0:
            logging.warning("predicted durations includes all 0 sequences. fill the first element with 1.")
            ops.lr_cumsum(ds, cum, alpha, all_ones=True)
            lens = [T] * B
        x = xs.contiguous()
        if x.dtype not in (_f32, torch.bfloat16):
            x = x.to(_f32)
        return _LengthRegulatorFn.apply(x, cum, max(lens), float(self.pad_value))


class _AASVCFunction(torch.autograd.Function):
    """Whole-model autograd node over AASVCEngine: differentiable outputs are after / before / log_p_attn / d_outs /
    bin_loss; ds (the integer MAS durations) is not."""

    @staticmethod
    def forward(ctx, model, xs, ys, dp_inputs, ilens, olens, *params):
        eng = model.engine
        after, before = eng.forward(xs, ys, dp_inputs, ilens, olens)
        d_outs = eng.dur_nll if eng.stochastic else eng.forward_d_outs()      # (B,) dur_nll | (B, T_text) d_outs
        ctx.model, ctx.token = model, model._fwd_token
        ctx.set_materialize_grads(False)      # an output the loss does not use arrives as None (d_outs before dp_train_start_steps)
        ds = eng.ds.clone()
        ctx.mark_non_differentiable(ds)
        return after.clone(), before.clone(), eng.log_p_attn.clone(), d_outs.clone(), eng.losses[2].clone(), ds

    @staticmethod
    def backward(ctx, g_after, g_before, g_logp, g_douts, g_bin, g_ds=None):
        model = ctx.model
        if ctx.token != model._fwd_token:
            raise S2SError("backward() after a newer forward(): the engine keeps one set of activations")
        eng = model.engine
        fresh = all(p.grad is None for p in model.parameters())
        dt = eng.adt
        B, Tt = eng.shapes["B"], eng.shapes["Tt"]
        z = lambda g, like, d: torch.zeros_like(like, dtype=d) if g is None else g.to(d).contiguous()
        d_after, d_before = z(g_after, eng.after, dt), z(g_before, eng.before, dt)
        d_logp = z(g_logp, eng.log_p_attn, _f32).clone()
        if g_bin is not None:
            # d bin_loss / d log_p_attn was produced by the MAS kernel; scale by the incoming scalar on the device
            gb = torch.empty_like(eng.d_logp_mas)
            ops.rowscale(eng.d_logp_mas.view(1, -1), g_bin.to(_f32).reshape(1), gb.view(1, -1))
            ops.axpy(gb, d_logp, 1.0)
        if eng.stochastic:
            eng._sdp_weight = 0.0          # no gradient arrived for dur_nll (before dp_train_start_steps): the predictor is skipped
            d_pre = g_douts                # (B,) d loss / d dur_nll, or None
        else:
            d_pre = torch.empty(B * Tt, 1, dtype=dt, device=eng.device)
            ops.duration_loss(eng.dp_pre, eng.ds, eng.tlens_dev, None, None, d_pre, g_douts=z(g_douts, eng.d_outs, _f32))
        eng.backward(d_after, d_before, d_logp, d_pre, zero_grad=fresh)
        model._sync_gradients()
        # trainers/aas_vc.py:119-133: without the duration loss the predictor's parameters are not in the graph
        # the stochastic predictor detaches its input (duration_predictor.py:236): its projection never sees a gradient
        unused = ("duration_predictor.", "duration_predictor_projection.") if g_douts is None else (
            ("duration_predictor_projection.",) if eng.stochastic else ())
        model._bind_grads(unused=unused)
        return (None,) * (6 + len(model._param_names))


class AASVC(VTN):
    """Drop-in for seq2seq_vc.models.AASVC (models/aas_vc.py:38-603), teacher-forced training path of the
    configuration family of egs/*/vc2/conf/aas_vc.*.yaml with the deterministic duration predictor.  Same constructor
    kwargs (the ones the reference accepts and ignores are accepted and ignored), forward signature, returned dict keys
    and state-dict names."""

    def __init__(self, idim, odim, adim: int = 384, aheads: int = 4, elayers: int = 6, eunits: int = 1536, dlayers: int = 6,
                 dunits: int = 1536, postnet_layers: int = 5, postnet_chans: int = 512, postnet_filts: int = 5,
                 positionwise_layer_type: str = "conv1d", positionwise_conv_kernel_size: int = 1, use_scaled_pos_enc: bool = True,
                 use_batch_norm: bool = True, encoder_input_layer: str = "linear", encoder_input_conv_kernel_size: int = 3,
                 encoder_normalize_before: bool = False, decoder_normalize_before: bool = False, encoder_concat_after: bool = False,
                 decoder_concat_after: bool = False, duration_predictor_use_encoder_outputs: bool = True,
                 duration_predictor_input_dim: int = None, duration_predictor_layers: int = 2, duration_predictor_chans: int = 384,
                 duration_predictor_kernel_size: int = 3, encoder_reduction_factor: int = 1, post_encoder_reduction_factor: int = 1,
                 decoder_reduction_factor: int = 1, encoder_type: str = "conformer", decoder_type: str = "conformer",
                 duration_predictor_type: str = "deterministic", conformer_pos_enc_layer_type: str = "rel_pos",
                 conformer_self_attn_layer_type: str = "rel_selfattn", use_macaron_style_in_conformer: bool = True,
                 use_cnn_in_conformer: bool = True, conformer_enc_kernel_size: int = 7, conformer_dec_kernel_size: int = 31,
                 spk_embed_dim: int = None, spk_embed_integration_type: str = "add", transformer_enc_dropout_rate: float = 0.1,
                 transformer_enc_positional_dropout_rate: float = 0.1, transformer_enc_attn_dropout_rate: float = 0.1,
                 transformer_dec_dropout_rate: float = 0.1, transformer_dec_positional_dropout_rate: float = 0.1,
                 transformer_dec_attn_dropout_rate: float = 0.1, duration_predictor_dropout_rate: float = 0.1,
                 postnet_dropout_rate: float = 0.5, init_type: str = "xavier_uniform", use_masking: bool = False,
                 use_weighted_masking: bool = False, stochastic_duration_predictor_kernel_size: int = 3,
                 stochastic_duration_predictor_dropout_rate: float = 0.5, stochastic_duration_predictor_flows: int = 4,
                 stochastic_duration_predictor_dds_conv_layers: int = 3, stochastic_duration_predictor_noise_scale: float = 0.8,
                 compute_dtype: str = "float32", device=None, seed: int = 0, **ignored):
        torch.nn.Module.__init__(self)
        unsupported = []
        if encoder_type != "conformer" or decoder_type != "conformer":
            unsupported.append("encoder_type/decoder_type != 'conformer'")
        if positionwise_layer_type not in ("linear", "conv1d", "conv1d-linear"):
            unsupported.append("positionwise_layer_type not in ('linear', 'conv1d', 'conv1d-linear')")
        if encoder_input_layer != "linear":
            unsupported.append("encoder_input_layer != 'linear'")
        if not (encoder_normalize_before and decoder_normalize_before):
            unsupported.append("post-LN conformer blocks")
        if conformer_pos_enc_layer_type != "rel_pos" or conformer_self_attn_layer_type != "rel_selfattn":
            unsupported.append("non rel_pos / rel_selfattn attention")
        if not (use_macaron_style_in_conformer and use_cnn_in_conformer and use_batch_norm):
            unsupported.append("conformer blocks without macaron FFN / CNN module / BatchNorm")
        if duration_predictor_type not in ("deterministic", "stochastic"):
            raise ValueError(f"Duration predictor type: {duration_predictor_type} is not supported.")      # models/aas_vc.py:187-190
        if duration_predictor_use_encoder_outputs or duration_predictor_input_dim is None:
            unsupported.append("duration_predictor_use_encoder_outputs=True")
        if encoder_reduction_factor != 1 or decoder_reduction_factor != 1:
            unsupported.append("encoder/decoder reduction factor != 1")
        if encoder_concat_after or decoder_concat_after or spk_embed_dim is not None:
            unsupported.append("concat_after / speaker embeddings")
        if unsupported:
            raise NotImplementedError("B200 AASVC hot path does not cover: " + ", ".join(unsupported))
        self.idim, self.odim = idim, odim
        self.spk_embed_dim = None
        self.encoder_reduction_factor, self.decoder_reduction_factor = 1, 1
        self.post_encoder_reduction_factor = post_encoder_reduction_factor
        self.encoder_type, self.decoder_type = encoder_type, decoder_type
        self.duration_predictor_type = duration_predictor_type
        self.viterbi_func = viterbi_decode          # operator seam of the reference (aas_vc.py:132); the engine runs s2s_mas itself
        self.hp = aasvc_default_hparams(
            idim=idim, odim=odim, adim=adim, aheads=aheads, elayers=elayers, eunits=eunits, dlayers=dlayers, dunits=dunits,
            duration_predictor_input_dim=duration_predictor_input_dim, duration_predictor_layers=duration_predictor_layers,
            duration_predictor_chans=duration_predictor_chans, duration_predictor_kernel_size=duration_predictor_kernel_size,
            postnet_layers=postnet_layers, postnet_filts=postnet_filts, postnet_chans=postnet_chans,
            post_encoder_reduction_factor=post_encoder_reduction_factor, conformer_enc_kernel_size=conformer_enc_kernel_size,
            conformer_dec_kernel_size=conformer_dec_kernel_size, transformer_enc_dropout_rate=transformer_enc_dropout_rate,
            transformer_enc_positional_dropout_rate=transformer_enc_positional_dropout_rate,
            transformer_enc_attn_dropout_rate=transformer_enc_attn_dropout_rate, transformer_dec_dropout_rate=transformer_dec_dropout_rate,
            transformer_dec_positional_dropout_rate=transformer_dec_positional_dropout_rate,
            transformer_dec_attn_dropout_rate=transformer_dec_attn_dropout_rate,
            duration_predictor_dropout_rate=duration_predictor_dropout_rate, postnet_dropout_rate=postnet_dropout_rate,
            positionwise_layer_type=positionwise_layer_type, positionwise_conv_kernel_size=positionwise_conv_kernel_size,
            duration_predictor_type=duration_predictor_type,
            stochastic_duration_predictor_kernel_size=stochastic_duration_predictor_kernel_size,
            stochastic_duration_predictor_dropout_rate=stochastic_duration_predictor_dropout_rate,
            stochastic_duration_predictor_flows=stochastic_duration_predictor_flows,
            stochastic_duration_predictor_dds_conv_layers=stochastic_duration_predictor_dds_conv_layers,
            stochastic_duration_predictor_noise_scale=stochastic_duration_predictor_noise_scale)
        self.stochastic_duration_predictor_noise_scale = stochastic_duration_predictor_noise_scale
        # compute_dtype: "bf16" (tcgen05, bf16 activations) | "float32" (float32 activations, fp32-accurate tcgen05 GEMMs through a
        # bf16 split: the parity mode) | "float32_simt" (float32 on the CUDA cores: the numerical yard-stick)
        self._bf16 = compute_dtype in ("bf16", "bfloat16", torch.bfloat16)
        self._fp32_gemm = "simt" if compute_dtype == "float32_simt" else "tc"
        self._seed = seed
        self._fwd_token = 0
        self.engine = None
        self._build(torch.device(device) if device is not None else torch.device("cpu"))

    def _build(self, device, state=None) -> None:
        self.engine = AASVCEngine(self.hp, device=device, bf16=self._bf16, seed=self._seed, fp32_gemm=self._fp32_gemm)
        if state is not None:
            self.engine.load_state_dict(state)
        self._modules.clear()
        self._param_names = []
        st = self.engine.store
        for name in st.names():
            node, leaf = self._node_for(name)
            node.register_parameter(leaf, torch.nn.Parameter(st.p(name), requires_grad=True))
            self._param_names.append(name)
        for name, buf in self.engine.buffers.items():
            node, leaf = self._node_for(name)
            node.register_buffer(leaf, buf)

    def forward(self, src_speech, src_speech_lengths, tgt_speech, tgt_speech_lengths, dp_inputs=None, dp_lengths=None, spembs=None):
        _require_cuda(src_speech, "AASVC")
        if dp_inputs is None:
            raise S2SError("dp_inputs is required (duration_predictor_use_encoder_outputs=False)")
        eng = self.engine
        eng.p16_dirty = True
        il, ol = _host_lens(src_speech_lengths), _host_lens(tgt_speech_lengths)
        xs = src_speech[:, :max(il)].to(_f32).contiguous()          # aas_vc.py:503-504
        ys = tgt_speech[:, :max(ol)].to(_f32).contiguous()
        dpi = dp_inputs.to(_f32).contiguous()
        self._fwd_token += 1
        if eng.training:
            ops.step_advance(None, eng.seed_dev)     # fresh dropout masks (and duration-predictor noise) per forward
        after, before, logp, d_outs, bin_loss, ds = _AASVCFunction.apply(self, xs, ys, dpi, il, ol, *self.parameters())
        for name, P in eng.attn.items():
            self.get_submodule(name).attn = P.float() if P.dtype != _f32 else P
        dev = xs.device
        ilens_out = torch.tensor(eng.tlens_host, dtype=torch.int64, device=dev)
        olens_out = torch.tensor(ol, dtype=torch.int64, device=dev)
        ret = dict(before_outs=before.float(), after_outs=after.float(), ds=ds, ilens=ilens_out, bin_loss=bin_loss,
                   log_p_attn=logp, olens_reduced=olens_out, olens=olens_out, ys=ys)
        ret["dur_nll" if eng.stochastic else "d_outs"] = d_outs          # models/aas_vc.py:408-419
        return ret


def _aasvc_inference(self, src_speech, tgt_speech=None, spembs=None, dp_input=None, use_teacher_forcing=False):
    """Drop-in for AASVC.inference (models/aas_vc.py:531-603): (T, idim) -> (outs (L, odim), d_outs (T_text,)); with a
    ground-truth target (the form AASVCTrainer's evaluation hook uses, trainers/aas_vc.py:243-245) the reference's 5-tuple
    (outs, d_outs, ds, log_p_attn, ilens_) where ds / log_p_attn come from the alignment module + MAS on the target."""
    if use_teacher_forcing or spembs is not None:
        raise NotImplementedError("inference with ground-truth durations / speaker embeddings is outside the hot path")
    _require_cuda(src_speech, "AASVC")
    if dp_input is None:
        raise S2SError("dp_input is required (duration_predictor_use_encoder_outputs=False)")
    self.engine.p16_dirty = True
    if tgt_speech is None:
        return self.engine.inference(src_speech, dp_input)
    outs, d_outs, ds, log_p_attn = self.engine.inference(src_speech, dp_input, tgt_speech)
    ilens_ = torch.tensor(src_speech.shape[0] // self.post_encoder_reduction_factor, dtype=torch.int64, device=src_speech.device)
    return outs, d_outs, ds, log_p_attn, ilens_


AASVC.inference = _aasvc_inference


# =================================================================================================
# FastSpeechVC (models/fastspeech_vc.py, trainers/nar_vc.py): drop-in model over FastSpeechVCEngine
# =================================================================================================
from .fsvc_engine import FastSpeechVCEngine  # noqa: E402
from .fsvc_engine import default_hparams as fsvc_default_hparams  # noqa: E402


class _FastSpeechVCFunction(torch.autograd.Function):
    """Whole-model autograd node over FastSpeechVCEngine: differentiable outputs are before / after / d_outs."""

    @staticmethod
    def forward(ctx, model, xs, ys, ds, dp_inputs, ilens, olens, *params):
        eng = model.engine
        after, before = eng.forward(xs, ys, ds, dp_inputs, ilens, olens)
        ctx.model, ctx.token = model, model._fwd_token
        ctx.set_materialize_grads(False)
        return before.clone(), after.clone(), eng.forward_d_outs()

    @staticmethod
    def backward(ctx, g_before, g_after, g_douts):
        model = ctx.model
        if ctx.token != model._fwd_token:
            raise S2SError("backward() after a newer forward(): the engine keeps one set of activations")
        eng = model.engine
        fresh = all(p.grad is None for p in model.parameters())
        dt = eng.adt
        B, Tt = eng.shapes["B"], eng.shapes["Tt"]
        z = lambda g, like, d: torch.zeros_like(like, dtype=d) if g is None else g.to(d).contiguous()
        d_pre = torch.zeros(B * Tt, 1, dtype=dt, device=eng.device)
        if g_douts is not None:          # d_outs = pre * mask: the gradient of the predictor's pre-activation is the masked incoming one
            m = (torch.arange(Tt, device=eng.device)[None, :] < eng.tlens_dev[:, None])
            d_pre.copy_((g_douts.to(_f32) * m).reshape(B * Tt, 1))
        eng.backward(z(g_after, eng.after, dt), z(g_before, eng.before, dt), d_pre, zero_grad=fresh)
        model._sync_gradients()
        model._bind_grads(unused=("duration_predictor.", "duration_predictor_projection.") if g_douts is None else ())
        return (None,) * (7 + len(model._param_names))


class FastSpeechVC(AASVC):
    """Drop-in for seq2seq_vc.models.FastSpeechVC (models/fastspeech_vc.py:21-513) in the configuration family of
    egs/arctic/vc2/conf/fs2_vc.melmelmel.v1.yaml: conformer encoder / decoder (rel_pos / rel_selfattn, macaron, CNN module), `conv2d`
    encoder input layer, duration predictor on a projected side input, LengthRegulator with the teacher's durations.  Same
    constructor kwargs, `forward(src_speech, src_speech_lengths, tgt_speech, tgt_speech_lengths, durations, durations_lengths,
    dp_inputs, dp_lengths)` -> `(before_outs, after_outs, d_outs, ilens, olens, ys)`, `inference(...)`, state-dict keys and
    parameter registration order."""

    def __init__(self, idim, odim, adim: int = 384, aheads: int = 4, elayers: int = 6, eunits: int = 1536, dlayers: int = 6,
                 dunits: int = 1536, postnet_layers: int = 5, postnet_chans: int = 512, postnet_filts: int = 5,
                 postnet_dropout_rate: float = 0.5, positionwise_layer_type: str = "conv1d", positionwise_conv_kernel_size: int = 1,
                 use_scaled_pos_enc: bool = True, use_batch_norm: bool = True, encoder_input_layer: str = "linear", encoder_input_conv_kernel_size: int = 3,
                 encoder_normalize_before: bool = False, decoder_normalize_before: bool = False, encoder_concat_after: bool = False,
                 decoder_concat_after: bool = False, duration_predictor_use_encoder_outputs: bool = True,
                 duration_predictor_input_dim: Optional[int] = None, duration_predictor_layers: int = 2,
                 duration_predictor_chans: int = 384, duration_predictor_kernel_size: int = 3, duration_predictor_dropout_rate: float = 0.1,
                 encoder_reduction_factor: int = 1, decoder_reduction_factor: int = 1, teacher_model_decoder_reduction_factor: int = 4,
                 encoder_type: str = "transformer", decoder_type: str = "transformer", transformer_enc_dropout_rate: float = 0.1,
                 transformer_enc_positional_dropout_rate: float = 0.1, transformer_enc_attn_dropout_rate: float = 0.1,
                 transformer_dec_dropout_rate: float = 0.1, transformer_dec_positional_dropout_rate: float = 0.1,
                 transformer_dec_attn_dropout_rate: float = 0.1, conformer_pos_enc_layer_type: str = "rel_pos",
                 conformer_self_attn_layer_type: str = "rel_selfattn", use_macaron_style_in_conformer: bool = True,
                 use_cnn_in_conformer: bool = True, conformer_enc_kernel_size: int = 7, conformer_dec_kernel_size: int = 31,
                 spk_embed_dim: Optional[int] = None, spk_embed_integration_type: str = "add",
                 compute_dtype: str = "float32", device=None, seed: int = 0, **ignored):
        torch.nn.Module.__init__(self)
        unsupported = []
        if encoder_type != "conformer" or decoder_type != "conformer":
            unsupported.append("encoder_type / decoder_type != 'conformer' (the reference's transformer decoder branch does not construct: fastspeech_vc.py:183)")
        if encoder_input_layer != "conv2d":
            unsupported.append("encoder_input_layer != 'conv2d'")
        if conformer_pos_enc_layer_type != "rel_pos" or conformer_self_attn_layer_type != "rel_selfattn":
            unsupported.append("conformer layers other than rel_pos / rel_selfattn")
        if not use_macaron_style_in_conformer or not use_cnn_in_conformer:
            unsupported.append("conformer without macaron / CNN module")
        if duration_predictor_use_encoder_outputs or duration_predictor_input_dim is None:
            unsupported.append("duration_predictor_use_encoder_outputs=True")
        if encoder_reduction_factor != 1 or decoder_reduction_factor != 1:
            unsupported.append("encoder / decoder reduction factors != 1")
        if not encoder_normalize_before or not decoder_normalize_before or encoder_concat_after or decoder_concat_after or not use_batch_norm:
            unsupported.append("non-default normalisation wiring")
        if spk_embed_dim is not None:
            unsupported.append("speaker embeddings")
        if positionwise_layer_type not in ("linear", "conv1d", "conv1d-linear"):
            unsupported.append("positionwise_layer_type not in ('linear', 'conv1d', 'conv1d-linear')")
        if unsupported:
            raise NotImplementedError("B200 FastSpeechVC hot path does not cover: " + ", ".join(unsupported))
        self.idim, self.odim = idim, odim
        self.spk_embed_dim = None
        self.encoder_reduction_factor, self.decoder_reduction_factor = 1, 1
        self.teacher_model_decoder_reduction_factor = teacher_model_decoder_reduction_factor
        self.encoder_type, self.decoder_type, self.encoder_input_layer = encoder_type, decoder_type, encoder_input_layer
        self.duration_predictor_use_encoder_outputs = False
        self.hp = fsvc_default_hparams(
            idim=idim, odim=odim, adim=adim, aheads=aheads, elayers=elayers, eunits=eunits, dlayers=dlayers, dunits=dunits,
            duration_predictor_input_dim=duration_predictor_input_dim, duration_predictor_layers=duration_predictor_layers,
            duration_predictor_chans=duration_predictor_chans, duration_predictor_kernel_size=duration_predictor_kernel_size,
            postnet_layers=postnet_layers, postnet_filts=postnet_filts, postnet_chans=postnet_chans,
            conformer_enc_kernel_size=conformer_enc_kernel_size, conformer_dec_kernel_size=conformer_dec_kernel_size,
            transformer_enc_dropout_rate=transformer_enc_dropout_rate,
            transformer_enc_positional_dropout_rate=transformer_enc_positional_dropout_rate,
            transformer_enc_attn_dropout_rate=transformer_enc_attn_dropout_rate, transformer_dec_dropout_rate=transformer_dec_dropout_rate,
            transformer_dec_positional_dropout_rate=transformer_dec_positional_dropout_rate,
            transformer_dec_attn_dropout_rate=transformer_dec_attn_dropout_rate,
            duration_predictor_dropout_rate=duration_predictor_dropout_rate, postnet_dropout_rate=postnet_dropout_rate,
            positionwise_layer_type=positionwise_layer_type, positionwise_conv_kernel_size=positionwise_conv_kernel_size,
            teacher_model_decoder_reduction_factor=teacher_model_decoder_reduction_factor)
        self._bf16 = compute_dtype in ("bf16", "bfloat16", torch.bfloat16)
        self._fp32_gemm = "simt" if compute_dtype == "float32_simt" else "tc"
        self._seed = seed
        self._fwd_token = 0
        self.engine = None
        self._build(torch.device(device) if device is not None else torch.device("cpu"))

    def _build(self, device, state=None) -> None:
        self.engine = FastSpeechVCEngine(self.hp, device=device, bf16=self._bf16, seed=self._seed, fp32_gemm=self._fp32_gemm)
        if state is not None:
            self.engine.load_state_dict(state)
        self._modules.clear()
        self._param_names = []
        st = self.engine.store
        for name in st.names():
            node, leaf = self._node_for(name)
            node.register_parameter(leaf, torch.nn.Parameter(st.p(name), requires_grad=True))
            self._param_names.append(name)
        for name, buf in self.engine.buffers.items():
            node, leaf = self._node_for(name)
            node.register_buffer(leaf, buf)

    def forward(self, src_speech, src_speech_lengths, tgt_speech, tgt_speech_lengths, durations, durations_lengths, dp_inputs=None,
                dp_lengths=None, spembs=None):
        _require_cuda(src_speech, "FastSpeechVC")
        if dp_inputs is None:
            raise S2SError("dp_inputs is required (duration_predictor_use_encoder_outputs=False)")
        eng = self.engine
        eng.p16_dirty = True
        il, ol = _host_lens(src_speech_lengths), _host_lens(tgt_speech_lengths)
        dl = _host_lens(durations_lengths)
        xs = src_speech[:, :max(il)].to(_f32).contiguous()                     # fastspeech_vc.py:413-415
        ys = tgt_speech[:, :max(ol)].to(_f32).contiguous()
        T2 = (((xs.shape[1] - 1) // 2) - 1) // 2
        ds = durations[:, :max(dl)].to(device=xs.device, dtype=torch.int64)
        if ds.shape[1] != T2:
            raise S2SError(f"durations cover {ds.shape[1]} encoder frames, the conv2d input layer yields {T2}")
        ds = ds.contiguous()
        # the regulated length is data: the reference's LengthRegulator reads it on the host too (repeat_interleave + pad_list)
        L = int(ds.sum(1).max().item()) * self.teacher_model_decoder_reduction_factor
        if L != ys.shape[1]:
            raise S2SError(f"sum of durations ({L}) != target length ({ys.shape[1]}): the L1 loss needs them equal (trainers/nar_vc.py:74)")
        dpi = dp_inputs.to(_f32).contiguous()
        self._fwd_token += 1
        if eng.training:
            ops.step_advance(None, eng.seed_dev)
        before, after, d_outs = _FastSpeechVCFunction.apply(self, xs, ys, ds, dpi, il, ol, *self.parameters())
        for name, P in eng.attn.items():
            self.get_submodule(name).attn = P.float() if P.dtype != _f32 else P
        ilens_out = torch.tensor(eng.tlens_host, dtype=torch.int64, device=xs.device)
        olens_out = torch.as_tensor(tgt_speech_lengths).to(xs.device)
        return before.float(), after.float(), d_outs, ilens_out, olens_out, ys

    def inference(self, src_speech, tgt_speech=None, spembs=None, dp_input=None, alpha: float = 1.0, use_teacher_forcing: bool = False):
        """fastspeech_vc.py:427-470 without teacher forcing: (outs (L, odim), d_outs (T',))."""
        if use_teacher_forcing:
            raise NotImplementedError("inference(use_teacher_forcing=True)")
        _require_cuda(src_speech, "FastSpeechVC")
        outs, d_outs = self.engine.inference(src_speech, dp_input, alpha)
        return outs, d_outs



class AASVCTrainStep(_ReferenceCheckpoint):
    """forward + L1 / forward-sum / bin / duration losses + backward (+ gradient all-reduce) + clip + Adam + WarmupLR,
    device-resident: mirrors AASVCTrainer._train_step (trainers/aas_vc.py:56-159).
    With ``use_graph`` one (B, T, L) batch shape is captured into two CUDA graphs (forward+losses+backward | clip+Adam);
    the NCCL all-reduce of the flat gradient buffer runs between them.

    ``gradient_accumulate_steps`` = k (trainers/base.py:65, aas_vc.py:141-149): every call is one micro-step whose
    gradients are summed into the flat buffer; only the k-th one all-reduces, clips, runs Adam and advances ``steps`` /
    the LR schedule.  The reference's ``loss / k`` is applied as one factor 1/(k * world) inside the Adam kernel (before
    the clip), so non-boundary micro-steps do no collective and no optimizer work at all."""

    def __init__(self, model, lr: float = 8e-5, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 grad_norm: float = 1.0, warmup_steps: int = 4000, dp_train_start_steps: int = 0, use_graph: bool = False,
                 process_group=None, gradient_accumulate_steps: int = 1):
        if int(gradient_accumulate_steps) < 1:
            raise ValueError("gradient_accumulate_steps must be >= 1")
        self.accum = int(gradient_accumulate_steps)
        self.backward_steps = 0                            # micro-steps so far (trainers/base.py:69)
        self.engine: AASVCEngine = model.engine if hasattr(model, "engine") else model
        self._model = model if hasattr(model, "engine") else None
        self.lr, self.betas, self.eps, self.wd = lr, betas, eps, weight_decay
        self.grad_norm, self.warmup = grad_norm, warmup_steps
        self.dp_start = dp_train_start_steps
        self.steps = 0
        self.use_graph = use_graph
        self.pg = process_group
        self.world = 1
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            self.world = torch.distributed.get_world_size(process_group)
        self._graphs: Dict[tuple, tuple] = {}
        self.replayed_launches = 0
        self.engine._evict_listeners.append(weakref.WeakMethod(self._on_evict))

    lr_at = VTNTrainStep.lr_at

    def _clock_of(self, name: str) -> torch.Tensor:        # the duration predictor keeps torch Adam's per-parameter clock
        dp = name.startswith(("duration_predictor.", "duration_predictor_projection."))
        return self.engine.dp_step_dev if dp else self.engine.step_dev

    def _fwd_bwd(self, xs, ys, dpi, with_duration, fresh=True, boundary=True):
        eng = self.engine
        eng.forward(xs, ys, dpi)
        eng.loss(ys, duration_loss=with_duration)
        eng.backward(zero_grad=fresh)
        if not boundary:
            ops.step_advance(None, eng.seed_dev)           # no optimizer tail follows: still draw fresh dropout masks next time

    def _allreduce(self):
        if self.world > 1:
            torch.distributed.all_reduce(self.engine.store.G, group=self.pg)

    def _update(self, with_duration=True):
        self.engine.optimizer_step(self.grad_norm, self.betas, self.eps, self.wd, grad_scale=1.0 / (self.world * self.accum),
                                   duration_predictor_active=with_duration)

    def __call__(self, xs, ilens, ys, olens, dp_inputs):
        """xs (B,T,idim), ys (B,L,odim), dp_inputs (B,T_dp,dp_idim): float32, CUDA-resident or pinned host memory.
        Returns the device tensor (l1, forward_sum, bin, duration) of this step without synchronising."""
        eng = self.engine
        with_dur = self.steps > self.dp_start          # trainers/aas_vc.py:113 (the reference skips the duration loss at step 0)
        fresh = self.backward_steps % self.accum == 0  # first micro-step of an accumulation window: zero the gradient buffer
        self.backward_steps += 1
        boundary = self.backward_steps % self.accum == 0
        if boundary:
            self.steps += 1
            eng.lr_dev.fill_(self.lr_at(self.steps))
        B, T, L = xs.shape[0], xs.shape[1], ys.shape[1]
        eng.training = True
        eng.prepare(B, T, L, ilens, olens)
        staged = self._take_staged(xs, ys, dp_inputs)
        if staged is not None:
            xs, ys, dp_inputs = staged
        if not self.use_graph:
            if not xs.is_cuda:
                xs, ys, dp_inputs = (t.to(eng.device, non_blocking=True) for t in (xs, ys, dp_inputs))
            self._fwd_bwd(xs, ys, dp_inputs, with_dur, fresh, boundary)
            if boundary:
                self._allreduce()
                self._update(with_dur)
            if staged is not None:
                self._staged_consumed()
            return eng.losses
        key = (B, T, L, dp_inputs.shape[1], with_dur, fresh, boundary)
        entry = self._graphs.get(key)
        if entry is None:
            statics = [torch.empty(t.shape, dtype=_f32, device=eng.device) for t in (xs, ys, dp_inputs)]
            for dst, src in zip(statics, (xs, ys, dp_inputs)):
                dst.copy_(src, non_blocking=True)
            if staged is not None:
                self._staged_consumed()
            self._fwd_bwd(*statics, with_dur, fresh, boundary)   # eager step: allocates every buffer outside the graph pool
            if boundary:
                self._allreduce()
                self._update(with_dur)
            torch.cuda.synchronize()
            g1, g2 = torch.cuda.CUDAGraph(), (torch.cuda.CUDAGraph() if boundary else None)
            n0 = _lib.launch_count()
            with _lib.graph_capture(g1):
                self._fwd_bwd(*statics, with_dur, fresh, boundary)
            if boundary:
                with _lib.graph_capture(g2):
                    self._update(with_dur)
            self._graphs[key] = (g1, g2, statics, _lib.launch_count() - n0)
            return eng.losses
        g1, g2, statics, n_kernels = entry
        if eng.p16_dirty:
            eng.sync_shadow()
        self.replayed_launches += n_kernels
        for dst, src in zip(statics, (xs, ys, dp_inputs)):
            if src.data_ptr() != dst.data_ptr():
                dst.copy_(src, non_blocking=True)
        if staged is not None:
            self._staged_consumed()
        g1.replay()
        if boundary:
            self._allreduce()
            g2.replay()
        return eng.losses


class NARVCTrainStep(AASVCTrainStep):
    """forward + L1 / duration losses + backward (+ gradient all-reduce) + clip + Adam + WarmupLR, device-resident: mirrors
    NARVCTrainer._train_step (trainers/nar_vc.py:52-103) for the FastSpeechVC drop-in.  Same options as AASVCTrainStep (CUDA graphs
    per batch shape, ``prefetch`` of the next batch, gradient accumulation, process group); the teacher's durations are a fourth
    input tensor (int64) and there is no duration-predictor warm-up."""

    def __init__(self, model, lr: float = 8e-5, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0, grad_norm: float = 1.0,
                 warmup_steps: int = 4000, use_graph: bool = False, process_group=None, gradient_accumulate_steps: int = 1):
        super().__init__(model, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, grad_norm=grad_norm, warmup_steps=warmup_steps,
                         dp_train_start_steps=0, use_graph=use_graph, process_group=process_group,
                         gradient_accumulate_steps=gradient_accumulate_steps)

    def _clock_of(self, name: str) -> torch.Tensor:
        return self.engine.step_dev

    def _fwd_bwd(self, xs, ys, ds, dpi, fresh=True, boundary=True):
        eng = self.engine
        eng.forward(xs, ys, ds, dpi)
        eng.loss(ys)
        eng.backward(zero_grad=fresh)
        if not boundary:
            ops.step_advance(None, eng.seed_dev)

    def _update(self, with_duration=True):
        self.engine.optimizer_step(self.grad_norm, self.betas, self.eps, self.wd, grad_scale=1.0 / (self.world * self.accum))

    def __call__(self, xs, ilens, ys, olens, durations, dp_inputs):
        """xs (B,T,idim), ys (B,L,odim), dp_inputs (B,T_dp,dp_idim) float32 and durations (B,T') int64 with
        T' = ((T - 1) // 2 - 1) // 2 and max_b sum(durations[b]) * teacher factor == L: CUDA-resident or pinned host memory.
        Returns the device tensor (l1, duration) of this step without synchronising."""
        eng = self.engine
        fresh = self.backward_steps % self.accum == 0
        self.backward_steps += 1
        boundary = self.backward_steps % self.accum == 0
        if boundary:
            self.steps += 1
            eng.lr_dev.fill_(self.lr_at(self.steps))
        B, T, L = xs.shape[0], xs.shape[1], ys.shape[1]
        if durations.dtype != torch.int64 or tuple(durations.shape) != (B, ((T - 1) // 2 - 1) // 2):
            raise S2SError(f"durations must be int64 of shape (B, ((T - 1) // 2 - 1) // 2), got {durations.dtype} {tuple(durations.shape)}")
        eng.training = True
        eng.prepare(B, T, L, ilens, olens)
        ins = (xs, ys, durations, dp_inputs)
        staged = self._take_staged(*ins)
        if staged is not None:
            ins = tuple(staged)
        if not self.use_graph:
            if not ins[0].is_cuda:
                ins = tuple(t.to(eng.device, non_blocking=True) for t in ins)
            self._fwd_bwd(*ins, fresh, boundary)
            if boundary:
                self._allreduce()
                self._update()
            if staged is not None:
                self._staged_consumed()
            return eng.losses
        key = (B, T, L, dp_inputs.shape[1], True, fresh, boundary)
        entry = self._graphs.get(key)
        if entry is None:
            statics = [torch.empty(t.shape, dtype=t.dtype, device=eng.device) for t in ins]
            for dst, src in zip(statics, ins):
                dst.copy_(src, non_blocking=True)
            if staged is not None:
                self._staged_consumed()
            self._fwd_bwd(*statics, fresh, boundary)            # eager step: allocates every buffer outside the graph pool
            if boundary:
                self._allreduce()
                self._update()
            torch.cuda.synchronize()
            g1, g2 = torch.cuda.CUDAGraph(), (torch.cuda.CUDAGraph() if boundary else None)
            n0 = _lib.launch_count()
            with _lib.graph_capture(g1):
                self._fwd_bwd(*statics, fresh, boundary)
            if boundary:
                with _lib.graph_capture(g2):
                    self._update()
            self._graphs[key] = (g1, g2, statics, _lib.launch_count() - n0)
            return eng.losses
        g1, g2, statics, n_kernels = entry
        if eng.p16_dirty:
            eng.sync_shadow()
        self.replayed_launches += n_kernels
        for dst, src in zip(statics, ins):
            if src.data_ptr() != dst.data_ptr():
                dst.copy_(src, non_blocking=True)
        if staged is not None:
            self._staged_consumed()
        g1.replay()
        if boundary:
            self._allreduce()
            g2.replay()
        return eng.losses
