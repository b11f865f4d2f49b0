"""Shared machinery of the explicit forward/backward engines (VTNEngine, AASVCEngine).

Owns the flat parameter store, the named activation buffers, dropout-site bookkeeping and the building blocks every
model on the hot path is made of (Linear fwd/bwd over s2s_gemm, LayerNorm, the attention core, the optimizer tail).
PyTorch only owns device memory here; all arithmetic happens in libs2svc_b200.so.
"""
from __future__ import annotations

import math
import os
import weakref
from collections import OrderedDict
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch

from . import ops
from ._lib import NO_DROP, Drop
from .params import ParamStore

_f32 = torch.float32
_i32 = torch.int32


def _r8(n: int) -> int:
    return (n + 7) // 8 * 8


def sinusoid_table(length: int, d_model: int, device) -> torch.Tensor:
    """PE table of seq2seq_vc/layers/positional_encoding.py:36-57 (host-built once, float32)."""
    pos = torch.arange(0, length, dtype=_f32).unsqueeze(1)
    div = torch.exp(torch.arange(0, d_model, 2, dtype=_f32) * -(math.log(10000.0) / d_model))
    pe = torch.zeros(length, d_model)
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe.to(device)


class EngineBase:
    """Parameters, buffers and building blocks shared by the engines."""

    def _setup(self, groups, buffer_specs, device, bf16: bool, seed: int) -> None:
        self.device = torch.device(device)
        self.bf16 = bool(bf16)
        self.adt = torch.bfloat16 if bf16 else _f32
        # s2s_gemm mode: 1 = bf16 tcgen05; float32 engines run the fp32-accurate tensor-core mode (2: bf16-split operands,
        # fp32 TMEM accumulation) unless fp32_gemm = "simt" asks for the CUDA-core yard-stick (0)
        self.mode = 1 if bf16 else (0 if getattr(self, "fp32_gemm", "tc") == "simt" else 2)
        self.store = ParamStore(groups, self.device, bf16_shadow=bf16)
        self.buffers: Dict[str, torch.Tensor] = {}
        for name, shape, dt in buffer_specs:
            self.buffers[name] = (torch.ones if name.endswith("running_var") else torch.zeros)(shape, dtype=dt, device=self.device)
        self.training = True
        self.base_seed = int(seed)
        self.seed_dev = torch.zeros(1, dtype=torch.int64, device=self.device)   # advanced once per step
        self.step_dev = torch.zeros(1, dtype=_f32, device=self.device)
        self.lr_dev = torch.zeros(1, dtype=_f32, device=self.device)
        self._pe: Dict[int, torch.Tensor] = {}
        self._bufs: Dict[Tuple, torch.Tensor] = {}
        self._sig: Optional[Tuple] = None
        self._site = 0
        self._site_ids: Dict[str, int] = {}
        self.attn: Dict[str, torch.Tensor] = {}      # every attention map of the last forward
        self._sqn = torch.zeros(1, dtype=_f32, device=self.device)
        self.p16_dirty = True
        self._lens_ring: Dict[Tuple[int, int], dict] = {}
        self._prepared = None
        self._retired: List[torch.Tensor] = []        # outgrown scratch / tables that captured CUDA graphs may still address
        self._sig_lru: "OrderedDict[Tuple, bool]" = OrderedDict()
        self._evict_listeners: List[Callable[[Tuple], None]] = []
        self._xcol_ready = set()

    def named_drop(self, name: str, p: float) -> Drop:
        """Dropout site addressed by name (stable id per engine): forward and backward just ask for the same name."""
        sid = self._site_ids.setdefault(name, 1000 + len(self._site_ids))
        if not self.training or p <= 0.0:
            return NO_DROP
        return Drop(p, self.base_seed, sid, self.seed_dev)

    def load_state_dict(self, sd: Dict[str, torch.Tensor]) -> None:
        for name in self.store.names():
            self.store.p(name).copy_(sd[name].to(self.device, _f32).reshape(self.store.offsets[name][1]))
        for name, buf in self.buffers.items():
            if name in sd:
                buf.copy_(sd[name].to(self.device, buf.dtype))
        self.p16_dirty = True

    def state_dict(self) -> Dict[str, torch.Tensor]:
        sd = {name: self.store.p(name).detach().clone() for name in self.store.names()}
        sd.update({k: v.detach().clone() for k, v in self.buffers.items()})
        return sd

    def W(self, name: str) -> torch.Tensor:
        """GEMM operand view of a weight: float32 master in parity mode, bf16 shadow otherwise."""
        return self.store.p16(name) if self.bf16 else self.store.p(name)

    def Wspan(self, names: Sequence[str], shape) -> torch.Tensor:
        return self.store.span(self.store.P16 if self.bf16 else self.store.P, list(names), shape)

    def sync_shadow(self) -> None:
        if self.bf16 and self.p16_dirty:
            ops.cast(self.store.P, self.store.P16)
        self.p16_dirty = False

    def buf(self, name: str, shape, dtype=None, zero: bool = False) -> torch.Tensor:
        dtype = dtype or self.adt
        key = (self._sig, name)
        t = self._bufs.get(key)
        if t is None:
            t = (torch.zeros if zero else torch.empty)(tuple(shape), dtype=dtype, device=self.device)
            self._bufs[key] = t
        else:
            assert tuple(t.shape) == tuple(shape) and t.dtype == dtype, (name, t.shape, shape)
        return t

    def pe(self, d: int, length: int) -> torch.Tensor:
        t = self._pe.get(d)
        if t is None or t.shape[0] < length:
            if t is not None:
                self._retired.append(t)      # graphs captured for shorter batches keep reading the old table
            t = sinusoid_table(max(length, 2048), d, self.device)
            self._pe[d] = t
        return t

    # ------------------------------------------------------------------ per-shape caches
    # Activation buffers are cached per batch shape (B, T, L, training) and the fused steps capture one CUDA graph per
    # shape; the reference collater pads to the batch maxima, so a long run sees many shapes.  At most
    # `max_cached_shapes` training shapes stay resident (least recently used first out); evicting one drops its buffers
    # and tells every listener (the train steps) to drop the graphs that address them.  Bucket T / L (e.g. to multiples
    # of 64) in the collater to keep the set small -- see INTEGRATION.md.
    max_cached_shapes = 8

    def _use_sig(self, sig: Tuple) -> None:
        self._sig = sig
        lru = self._sig_lru
        lru.pop(sig, None)
        lru[sig] = True
        while len(lru) > max(1, int(self.max_cached_shapes)):
            old, _ = lru.popitem(last=False)
            self._evict_sig(old)

    def _evict_sig(self, sig: Tuple) -> None:
        for cb in list(self._evict_listeners):
            if isinstance(cb, weakref.WeakMethod):      # train steps register weakly: engine -> step -> engine must not be a cycle
                fn = cb()
                if fn is None:
                    self._evict_listeners.remove(cb)
                    continue
                cb = fn
            cb(sig)
        for key in [k for k in self._bufs if k[0] == sig]:
            del self._bufs[key]

    def _ship_lens(self, rows: Sequence[Sequence[int]]) -> torch.Tensor:
        """Per-utterance length vectors (host ints from the collater) -> the device `lens` buffer of the current shape in
        ONE small non-blocking copy.  The pinned staging rows form a ring guarded by CUDA events: the step never
        synchronises, so the host may run several steps ahead of the copy engine and must not overwrite a staging row
        whose copy has not executed yet."""
        n, B = len(rows), len(rows[0])
        dst = self.buf("lens", (n, B), _i32)
        vals = torch.tensor(rows, dtype=_i32)
        if self.device.type != "cuda":
            dst.copy_(vals)
            return dst
        ring = self._lens_ring.get((n, B))
        if ring is None:
            ring = {"slots": [[torch.empty(n, B, dtype=_i32).pin_memory(), None] for _ in range(4)], "next": 0}
            self._lens_ring[(n, B)] = ring
        slot = ring["slots"][ring["next"]]
        ring["next"] = (ring["next"] + 1) % len(ring["slots"])
        if slot[1] is not None:
            slot[1].synchronize()          # returns at once unless the host is a full ring ahead of the device
        slot[0].copy_(vals)
        dst.copy_(slot[0], non_blocking=True)
        if slot[1] is None:
            slot[1] = torch.cuda.Event()
        slot[1].record()
        return dst

    def drop(self, p: float) -> Drop:
        """Next dropout site of the step (forward and backward enumerate sites in the same order)."""
        self._site += 1
        if not self.training or p <= 0.0:
            return NO_DROP
        return Drop(p, self.base_seed, self._site, self.seed_dev)

    def _lin_fwd(self, x2d, w, bias, out, relu=False, drop=NO_DROP, residual=None):
        if w.shape[0] <= 4 and not relu and residual is None and drop.p == 0.0:
            return ops.skinny_linear_fwd(x2d, w, bias, out)
        return ops.gemm(x2d, w, out, bias=bias, relu=relu, drop=drop, residual=residual, mode=self.mode)

    def _lin_bwd(self, dy2d, x2d, w, gw, gb, dx=None, dx_residual=None, dx_accumulate=False, dx_gate=None, dx_gate_scale=1.0):
        """dW += dy^T x ; db += colsum(dy) ; dx = dy W (+ residual | += ); with dx_gate (the relu output that produced x2d's
        layer input) dx leaves as relu'(.) * scale of it: the ReLU / dropout backward rides in the GEMM epilogue."""
        mode = self.mode
        if w.shape[0] <= 4 and dx_residual is None and dx_gate is None:
            return ops.skinny_linear_bwd(dy2d, x2d, w, gw, gb, dx, dx_accumulate)
        if gw is not None:
            if self._defer is not None:
                self._defer[0].append((dy2d.t(), x2d.t(), gw, dict(accumulate=True)))
            else:
                ops.gemm(dy2d.t(), x2d.t(), gw, accumulate=True, mode=mode)
        if gb is not None:
            if self._defer is not None:
                self._defer[1].append((dy2d, gb))
            else:
                ops.colsum(dy2d, gb)      # (a side-stream overlap with the two GEMMs was measured: no gain, the persistent GEMM owns the SMs)
        if dx is not None:
            if dx_gate is not None and not self.fuse_relu_gate:        # A/B switch: separate relu' pass
                ops.gemm(dy2d, w.t(), dx, residual=dx_residual, accumulate=dx_accumulate, mode=mode)
                ops.relu_bwd(dx, dx_gate, dx, dx_gate_scale)
            else:
                ops.gemm(dy2d, w.t(), dx, residual=dx_residual, accumulate=dx_accumulate, mode=mode, gate=dx_gate, gate_scale=dx_gate_scale)
        return dx

    # Deferred weight gradients: inside a layer's backward the dW = dy^T x products and the bias-gradient column sums are only
    # recorded; _flush_defer() at the end of the layer issues them as ONE grouped tcgen05 launch and ONE multi-colsum launch
    # (7 + 7 launches of a decoder layer become 2).  Every recorded dy / x must stay untouched until the flush: the engines
    # give the per-layer gradient temporaries distinct buffers while deferring (see VTNEngine.backward).
    _defer = None
    group_dw = os.environ.get("S2S_GROUP_DW", "1") != "0"
    conv1_dw_tc = os.environ.get("S2S_CONV1_DW_TC", "1") != "0"      # A/B switch: first-convolution weight gradient on the GEMM path

    def _begin_defer(self) -> bool:
        if self.mode == 1 and self.group_dw and self.device.type == "cuda":
            self._defer = ([], [])
            return True
        return False

    def _flush_defer(self) -> None:
        d, self._defer = self._defer, None
        if d is not None:
            ops.gemm_grouped(d[0], mode=self.mode)
            ops.colsum_multi(d[1])

    def _ln_fwd(self, x, name, tag, eps: float = 1e-12):
        B_, T_, d = x.shape
        y = self.buf(tag + ".y", x.shape)
        mean = self.buf(tag + ".mean", (B_ * T_,), _f32)
        rstd = self.buf(tag + ".rstd", (B_ * T_,), _f32)
        ops.layernorm_fwd(x, self.store.p(name + ".weight"), self.store.p(name + ".bias"), y, mean, rstd, eps)
        return y

    def _ln_bwd(self, dy, x, name, tag, dx, dres=None, dx_drop=None, drop: Drop = NO_DROP):
        """dx_drop: second output dropout'(dx) -- the gradient behind the dropout of the branch that fed the normalised sum."""
        if dx_drop is not None and drop.p <= 0.0:
            dx_drop = None
        if dx_drop is not None and not self.fuse_ln_dropout:           # A/B switch: separate dropout' pass
            self._ln_bwd(dy, x, name, tag, dx, dres=dres)
            ops.dropout_bwd(dx, dx_drop, drop)
            return dx
        ops.layernorm_bwd(dy, x, self.store.p(name + ".weight"), self.buf(tag + ".mean", (x.shape[0] * x.shape[1],), _f32),
                          self.buf(tag + ".rstd", (x.shape[0] * x.shape[1],), _f32), dx, self.store.g(name + ".weight"),
                          self.store.g(name + ".bias"), dres=dres, dx_drop=dx_drop, drop=drop)
        return dx

    def _fused_attn(self, dk: int, T2: int) -> bool:
        """bf16 path with a small head dimension: scores + softmax (and dP + softmax') run as one kernel each."""
        return self.mode == 1 and dk in ops.FUSED_ATTN_DK and getattr(self, "fused_attention", True)

    # Which attention maps are written to HBM.  The reference keeps `self.attn` of every MultiHeadedAttention
    # (attention.py:81-85) but only the decoder's source-attention maps leave VTN.forward (models/vtn.py:280-287), so:
    #   "src"  (default) source-attention maps only      "all"  every map (as the reference module tree holds them)
    #   "none" no map: the fused training steps, which consume none (maps named in `attn_emit_names` are still written,
    #          e.g. the layers a guided-attention loss reads)
    attn_emit = "src"
    # A/B switches (tools/ab_step.py): fold relu' into the dX GEMM epilogue / dropout' into the LayerNorm backward
    fuse_relu_gate = os.environ.get("S2S_FUSE_GATE", "1") != "0"
    fuse_ln_dropout = os.environ.get("S2S_FUSE_LNDROP", "1") != "0"
    attn_emit_names: frozenset = frozenset()
    flash_attention = True

    def _flash_attn(self, dk: int) -> bool:
        """bf16 path, d_k a multiple of 16 up to 128: one tcgen05 kernel per direction, S / P never leave the chip."""
        return self.mode == 1 and dk % 16 == 0 and 16 <= dk <= 128 and self.flash_attention

    def _want_P(self, store_name: str) -> bool:
        if store_name in self.attn_emit_names or self.attn_emit == "all":
            return True
        return self.attn_emit == "src" and store_name.endswith("src_attn")

    def _attn_core_fwd(self, q, k, v, klens, causal, tag, store_name):
        """q (B,T1,H,dk) / k, v (B,T2,H,dk) strided views -> ctx (B,T1,d); keeps P (or the row statistics) for backward."""
        B_, T1, H, dk = q.shape
        T2 = k.shape[1]
        ld = _r8(T2)
        ctx = self.buf(tag + ".ctx", (B_, T1, H * dk))
        if self._flash_attn(dk):
            P = self.buf(tag + ".P", (B_, H, T1, ld)) if self._want_P(store_name) else None
            lse = self.buf(tag + ".lse", ops.attn_lse_shape(B_, H, T1), _f32)
            ops.attn_fwd_tc(q, k, v, ctx.view(B_, T1, H, dk), lse, klens, causal, 1.0 / math.sqrt(dk), P)
            if P is not None:
                self.attn[store_name] = P[..., :T2]
            return ctx
        P = self.buf(tag + ".P", (B_, H, T1, ld))
        if self._fused_attn(dk, T2):
            ops.attn_probs_fwd(q, k, P, klens, causal, T2, 1.0 / math.sqrt(dk))
        else:
            # S[b,h] = q_bh k_bh^T / sqrt(dk)   (reference: attention.py:95-104)
            ops.gemm(q.permute(0, 2, 1, 3), k.permute(0, 2, 1, 3), P[..., :T2], alpha=1.0 / math.sqrt(dk), mode=self.mode)
            ops.softmax_fwd(P, klens, causal, T2)
        self.attn[store_name] = P[..., :T2]
        # ctx[b,t,h,:] = sum_s P[b,h,t,s] v[b,s,h,:]
        ops.gemm(P[..., :T2], v.permute(0, 2, 3, 1), ctx.view(B_, T1, H, dk).permute(0, 2, 1, 3), mode=self.mode)
        return ctx

    def _attn_core_bwd(self, dctx, q, k, v, dq, dk_, dv, tag, d_att=None, klens=None, causal=False):
        """Gradients of the attention core; dq/dk_/dv are (B,T,H,dk) strided views to be filled.  klens / causal: the
        forward's mask (the flash backward recomputes P from them; the stored-P path does not need them)."""
        B_, T1, H, dk = q.shape
        T2 = k.shape[1]
        ld = _r8(T2)
        if self._flash_attn(dk) and d_att is None:
            lse = self.buf(tag + ".lse", ops.attn_lse_shape(B_, H, T1), _f32)
            dvec = self._scratch("attn.D", ops.attn_lse_shape(B_, H, T1), _f32)
            ctx = self.buf(tag + ".ctx", (B_, T1, H * dk))
            ops.attn_bwd_tc(q, k, v, ctx.view(B_, T1, H, dk), dctx.view(B_, T1, H, dk), lse, dvec, dq, dk_, dv, klens, causal,
                            1.0 / math.sqrt(dk))
            return
        P = self.buf(tag + ".P", (B_, H, T1, ld))       # a d_att gradient needs the stored map: its site is in attn_emit_names
        dP = self._scratch("dP", (B_, H, T1, ld))
        dctx4 = dctx.view(B_, T1, H, dk).permute(0, 2, 1, 3)
        # dv[b,s,h,j] = sum_t P[b,h,t,s] dctx[b,t,h,j]
        ops.gemm(P[..., :T2].transpose(-1, -2), dctx4.transpose(-1, -2), dv.permute(0, 2, 1, 3), mode=self.mode)
        if self._fused_attn(dk, T2) and T2 <= 256:
            # measured (C2, d_k 48): fused 49 us vs GEMM + softmax' 67 us at T2 = 127, but 305 us vs 146 us at T2 = 512, where the
            # CTA's P slice no longer fits in shared memory next to enough resident CTAs
            ops.attn_probs_bwd(dctx.view(B_, T1, H, dk), v, P, d_att, dP, T2, 1.0 / math.sqrt(dk))
        else:
            # dP[b,h,t,s] = sum_j dctx[b,t,h,j] v[b,s,h,j]
            ops.gemm(dctx4, v.permute(0, 2, 1, 3), dP[..., :T2], mode=self.mode)
            if d_att is not None:
                ops.add(dP, d_att, dP)
            ops.softmax_bwd(P, dP, T2, 1.0 / math.sqrt(dk))
        dS = dP
        # dq[b,t,h,j] = sum_s dS[t,s] k[s,j] ; dk[b,s,h,j] = sum_t dS[t,s] q[t,j]
        ops.gemm(dS[..., :T2], k.permute(0, 2, 3, 1), dq.permute(0, 2, 1, 3), mode=self.mode)
        ops.gemm(dS[..., :T2].transpose(-1, -2), q.permute(0, 2, 3, 1), dk_.permute(0, 2, 1, 3), mode=self.mode)

    def _scratch(self, name, shape, dtype=None):
        """Step-local scratch, shared between layers (sized to the largest request)."""
        dtype = dtype or self.adt
        n = 1
        for s in shape:
            n *= s
        key = ("scratch", name, dtype)
        t = self._bufs.get(key)
        if t is None or t.numel() < n:
            if t is not None:
                # CUDA graphs captured for smaller batch shapes have this pointer baked in: keep the old allocation
                # alive (their replays keep using it) instead of handing it back to the caching allocator
                self._retired.append(t)
            t = torch.empty(max(n, 1), dtype=dtype, device=self.device)
            self._bufs[key] = t
        return t[:n].view(shape)

    def _drop_bwd(self, dy2d, drop: Drop, out):
        if drop.p <= 0.0:
            return dy2d
        o = out[: dy2d.shape[0]].view(dy2d.shape) if out.shape != dy2d.shape else out
        return ops.dropout_bwd(dy2d, o, drop)

    def optimizer_step(self, max_norm: float = 1.0, betas=(0.9, 0.999), eps: float = 1e-8,
                       weight_decay: float = 0.0, grad_scale: float = 1.0) -> None:
        """clip_grad_norm_(max_norm) + Adam over the flat buffers (trainers/ar_vc.py:99-107).

        The learning rate is read from the device scalar ``self.lr_dev`` (set by the caller)."""
        st = self.store
        ops.step_advance(self.step_dev, self.seed_dev)
        self._sqn.zero_()
        ops.sqnorm(st.G, self._sqn)
        ops.adam_step(st.P, st.G, st.M, st.V, st.P16, self.lr_dev, betas[0], betas[1], eps, weight_decay, self.step_dev,
                      self._sqn, max_norm, grad_scale)
        self.p16_dirty = False  # adam_step refreshed the bf16 shadow

    # ------------------------------------------------------------------ postnet (shared by VTN / TransformerTTS / AASVC)
    def _postnet_fwd(self, before: torch.Tensor, drop_of) -> torch.Tensor:
        """after = before + Postnet(before) (modules/pre_postnets.py:105-185): Conv1d(k) as a taps-GEMM over haloed
        channels-last rows, training/eval BatchNorm (+tanh) and dropout.  drop_of(i) -> Drop of layer i."""
        hp, st = self.hp, self.store
        B, Lo, odim = before.shape
        n_post, k = hp["postnet_layers"], hp["postnet_filts"]
        after = self.buf("out.after", (B, Lo, odim))
        if n_post == 0:
            after.copy_(before)
            return after
        halo = (k - 1) // 2
        Lp = Lo + 2 * halo
        ypad = self.buf("post.in", (B, Lp, odim))
        ops.pad_rows(before, ypad, halo)
        for i in range(n_post):
            pn = f"postnet.postnet.{i}"
            w = st.p(pn + ".0.weight")
            oc, ic = w.shape[0], w.shape[1]
            wp = self.buf(f"w.post{i}p", (oc, k, ic))
            wpt = self.buf(f"w.post{i}pt", (ic, k, oc))
            ops.pack_conv1d_w(w, wp, wpt)
            z = self.buf(f"post.z{i}", (B, Lp, oc))
            M = B * Lp - 2 * halo
            ops.gemm(ypad.view(B * Lp, ic), wp, z.view(B * Lp, oc)[halo:], taps=k, row_mask=(Lp, halo, halo, halo + Lo),
                     mode=self.mode, M=M)
            mean = self.buf(f"post.mean{i}", (oc,), _f32)
            invstd = self.buf(f"post.invstd{i}", (oc,), _f32)
            if self.training:
                sums = self.buf(f"post.sums{i}", (2 * oc,), _f32)
                sums.zero_()
                ops.bn_stats(z, sums, Lo, halo)
                ops.bn_finalize(sums, mean, invstd, self.buffers[pn + ".1.running_mean"], self.buffers[pn + ".1.running_var"],
                                B * Lo)
                self.buffers[pn + ".1.num_batches_tracked"] += 1
            else:
                ops.bn_eval_stats(self.buffers[pn + ".1.running_mean"], self.buffers[pn + ".1.running_var"], mean, invstd)
            y = self.buf(f"post.y{i}", (B, Lp, oc))
            ops.bn_apply(z, mean, invstd, st.p(pn + ".1.weight"), st.p(pn + ".1.bias"), y, Lo, halo, i != n_post - 1, drop_of(i))
            ypad = y
        post = self._scratch("post.out", (B, Lo, odim))
        ops.unpad_rows(ypad, post, halo)
        ops.add(before, post, after)
        return after

    def _postnet_bwd(self, d_after: torch.Tensor, d_before: torch.Tensor, drop_of) -> torch.Tensor:
        """Gradient of (before -> after) w.r.t. `before`, summed with d_before; accumulates the postnet parameter gradients."""
        hp, st = self.hp, self.store
        B, Lo, odim = d_after.shape
        n_post, k = hp["postnet_layers"], hp["postnet_filts"]
        dbefore_tot = self._scratch("dbefore", (B, Lo, odim))
        if n_post == 0:
            ops.add(d_after, d_before, dbefore_tot)
            return dbefore_tot
        halo = (k - 1) // 2
        Lp = Lo + 2 * halo
        dy = self._scratch("post.dy_a", (B, Lp, odim))
        ops.pad_rows(d_after, dy, halo)
        for i in reversed(range(n_post)):
            pn = f"postnet.postnet.{i}"
            w = st.p(pn + ".0.weight")
            oc, ic = w.shape[0], w.shape[1]
            z = self.buf(f"post.z{i}", (B, Lp, oc))
            y = self.buf(f"post.y{i}", (B, Lp, oc))
            xin = self.buf(f"post.y{i - 1}", (B, Lp, ic)) if i > 0 else self.buf("post.in", (B, Lp, odim))
            mean = self.buf(f"post.mean{i}", (oc,), _f32)
            invstd = self.buf(f"post.invstd{i}", (oc,), _f32)
            drop = drop_of(i)
            dz = self._scratch(f"post.dz{i % 2}", (B, Lp, oc))
            gam, bet = st.p(pn + ".1.weight"), st.p(pn + ".1.bias")
            sums = self._scratch("post.bsums", (2 * oc,), _f32)
            sums.zero_()
            ops.bn_bwd_reduce(dy, y, z, mean, invstd, gam, bet, sums, Lo, halo, i != n_post - 1, drop)
            if self.training:
                ops.bn_bwd_apply(dy, y, z, mean, invstd, gam, bet, sums, dz, st.g(pn + ".1.weight"), st.g(pn + ".1.bias"),
                                 Lo, halo, i != n_post - 1, drop)
            else:
                ops.bn_bwd_apply(dy, y, z, mean, invstd, gam, bet, None, dz, None, None, Lo, halo, i != n_post - 1, drop)
                ops.add(st.g(pn + ".1.bias"), sums[:oc], st.g(pn + ".1.bias"))
                ops.add(st.g(pn + ".1.weight"), sums[oc:], st.g(pn + ".1.weight"))
            M = B * Lp - 2 * halo
            # dWp[oc][t][ic] = sum_m dz[m + halo][oc] * xin[m + t][ic]
            gwp = self._scratch("post.gwp", (oc, k, ic), _f32)
            gwp.zero_()
            dzt = dz.view(B * Lp, oc)[halo:halo + M].t()
            self._taps_dw(dzt, xin.view(B * Lp, ic), gwp, M)
            ops.transpose_last2(gwp, st.g(pn + ".0.weight"), oc, k, ic, accumulate=True)
            # dxin = conv_transpose(dz): taps-GEMM with the flipped, transposed kernel
            dxin = self._scratch(f"post.dx{i % 2}", (B, Lp, ic))
            wpt = self.buf(f"w.post{i}pt", (ic, k, oc))
            if halo > 0:
                dxin.view(B * Lp, ic)[:halo].zero_()
                dxin.view(B * Lp, ic)[B * Lp - halo:].zero_()
            ops.gemm(dz.view(B * Lp, oc), wpt, dxin.view(B * Lp, ic)[halo:], taps=k, row_mask=(Lp, halo, halo, halo + Lo),
                     mode=self.mode, M=M)
            dy = dxin
        dpost_in = self._scratch("post.dunpad", (B, Lo, odim))
        ops.unpad_rows(dy, dpost_in, halo)
        ops.add(d_after, d_before, dbefore_tot)
        ops.add(dbefore_tot, dpost_in, dbefore_tot)
        return dbefore_tot

    def _taps_dw(self, dzt: torch.Tensor, xrows: torch.Tensor, gwp: torch.Tensor, M: int) -> None:
        """Weight gradient of a Conv1d-as-taps-GEMM in ONE launch: gwp[oc][t][ic] += sum_m dz[m + halo][oc] * x[m + t][ic].
        The k shifted copies of the input are one matrix of overlapping rows: element (m, t * ic + c) of it sits at
        x_flat[m * ic + (t * ic + c)], i.e. an (M, k * ic) view with row stride ic -- legal for a TMA tensor map, so all taps
        form the N dimension of a single (oc x k*ic x M) GEMM instead of k skinny ones (zeroed gwp, split-K inside the kernel)."""
        oc, k, ic = gwp.shape
        if self.device.type != "cuda" or (self.mode != 0 and ic % 8 != 0):
            for t in range(k):          # CPU contracts of the host-logic tests; unaligned rows (TMA needs 16-byte strides)
                ops.gemm(dzt, xrows[t:t + M].t(), gwp[:, t, :], accumulate=True, mode=self.mode)
            return
        win = xrows.as_strided((M, k * ic), (ic, 1))
        ops.gemm(dzt, win.t(), gwp.view(oc, k * ic), accumulate=True, mode=self.mode)

    # ------------------------------------------------------------------ Conv1d(k) over time as a taps-GEMM (haloed rows)
    def _conv1d_fwd(self, xpad: torch.Tensor, name: str, L: int, relu: bool, tag: str) -> torch.Tensor:
        """xpad (B, L + 2*halo, ic) with zero halos -> z (B, L + 2*halo, oc) = [relu](Conv1d(x) + bias), halos zero.
        Weight `name.weight` (oc, ic, k) / `name.bias` (torch.nn.Conv1d, padding (k-1)/2)."""
        st = self.store
        w = st.p(name + ".weight")
        oc, ic, k = w.shape
        B, Lp, _ = xpad.shape
        halo = (k - 1) // 2
        assert Lp == L + 2 * halo and xpad.shape[2] == ic
        wp = self.buf(f"w.{tag}.p", (oc, k, ic))
        wpt = self.buf(f"w.{tag}.pt", (ic, k, oc))
        ops.pack_conv1d_w(w, wp, wpt)
        z = self.buf(tag + ".z", (B, Lp, oc))
        M = B * Lp - 2 * halo
        if halo > 0:
            z.view(B * Lp, oc)[:halo].zero_()
            z.view(B * Lp, oc)[B * Lp - halo:].zero_()
        ops.gemm(xpad.view(B * Lp, ic), wp, z.view(B * Lp, oc)[halo:], bias=st.p(name + ".bias"), relu=relu, taps=k,
                 row_mask=(Lp, halo, halo, halo + L), mode=self.mode, M=M)
        return z

    def _conv1d_bwd(self, dz: torch.Tensor, xpad: torch.Tensor, name: str, L: int, tag: str, dx: Optional[torch.Tensor]):
        """dz (B, Lp, oc) with zero halos = gradient at the conv output (after relu'); accumulates dW / dbias and
        writes dx (B, Lp, ic) (zero halos) when given."""
        st = self.store
        w = st.p(name + ".weight")
        oc, ic, k = w.shape
        B, Lp, _ = xpad.shape
        halo = (k - 1) // 2
        M = B * Lp - 2 * halo
        gwp = self._scratch("conv1d.gwp", (oc, k, ic), _f32)
        gwp.zero_()
        dzt = dz.view(B * Lp, oc)[halo:halo + M].t()
        self._taps_dw(dzt, xpad.view(B * Lp, ic), gwp, M)
        ops.transpose_last2(gwp, st.g(name + ".weight"), oc, k, ic, accumulate=True)
        ops.colsum(dz.view(B * Lp, oc), st.g(name + ".bias"))
        if dx is not None:
            wpt = self.buf(f"w.{tag}.pt", (ic, k, oc))
            if halo > 0:
                dx.view(B * Lp, ic)[:halo].zero_()
                dx.view(B * Lp, ic)[B * Lp - halo:].zero_()
            ops.gemm(dz.view(B * Lp, oc), wpt, dx.view(B * Lp, ic)[halo:], taps=k, row_mask=(Lp, halo, halo, halo + L),
                     mode=self.mode, M=M)
        return dx

    # ------------------------------------------------------------------ first convolution of Conv2dSubsampling
    conv1_fwd_tc = os.environ.get("S2S_CONV1_FWD_TC", "1") != "0"      # A/B switch: forward on the GEMM path (bf16 engines)
    _xcol_ready: set = set()

    def _conv1_fwd(self, xs, prefix: str, y1, tag: str):
        """y1 (B, T1, F1, d) = relu(Conv2d(1 -> d, 3, 2)(xs)).  bf16 engines: a 16-column patch matrix of the input (kept for the
        weight gradient) times the packed weights on the tcgen05 GEMM; float32 engines: the direct kernel."""
        st = self.store
        w, b = st.p(prefix + ".conv.0.weight"), st.p(prefix + ".conv.0.bias")
        if self.mode == 1 and self.conv1_fwd_tc and self.device.type == "cuda":
            B, T1, F1, d = y1.shape
            xcol = self.buf(tag + ".xcol", (B * T1 * F1, 16))
            ops.conv1_fwd_tc(xs, w, b, y1, xcol, self.buf(f"w.{tag}.conv1p", (d, 16)), mode=1)
            self._xcol_ready.add((self._sig, tag))
        else:
            ops.conv1_fwd(xs, w, b, y1)
        return y1

    def _conv1_bwd(self, xs, dy1, prefix: str, tag: str) -> None:
        st = self.store
        gw, gb = st.g(prefix + ".conv.0.weight"), st.g(prefix + ".conv.0.bias")
        if self.mode == 1 and self.conv1_dw_tc and self.device.type == "cuda":
            # conv.0's weight / bias gradient as dy1^T x (16-column patch matrix of the input) on the tcgen05 GEMM
            B, T1, F1, d = dy1.shape
            ready = (self._sig, tag) in self._xcol_ready          # the forward of this batch shape left the patch matrix in its buffer
            xcol = self.buf(tag + ".xcol", (B * T1 * F1, 16))
            ops.conv1_bwd_tc(xs, dy1, gw, gb, xcol, self._scratch("conv1.g16", (d, 16), _f32), mode=1, xcol_ready=ready)
        else:
            ops.conv1_bwd(xs, dy1, gw, gb)

    # ------------------------------------------------------------------ Conv2dSubsampling without positional encoding
    def _conv2d_sub_fwd(self, xs: torch.Tensor, prefix: str, out_name: str, tag: str) -> torch.Tensor:
        """(B, T, idim) float32 -> (B*T2, d): Conv2d(1->d,3,s2)+ReLU, Conv2d(d->d,3,s2)+ReLU, Linear(d*F2 -> d)
        (modules/transformer/subsampling.py:58-94)."""
        st = self.store
        B, T, idim = xs.shape
        d = st.p(prefix + ".conv.0.weight").shape[0]
        T1, F1 = (T - 1) // 2, (idim - 1) // 2
        T2, F2 = (T1 - 1) // 2, (F1 - 1) // 2
        w2p = self.buf(f"w.{tag}.conv2p", (d, 9, d))
        ops.transpose_last2(st.p(prefix + ".conv.2.weight"), w2p, d, d, 9)
        woutp = self.buf(f"w.{tag}.outp", (d, F2, d))
        ops.transpose_last2(st.p(out_name + ".weight"), woutp, d, d, F2)
        y1 = self.buf(tag + ".y1", (B, T1, F1, d))
        self._conv1_fwd(xs, prefix, y1, tag)
        col = self._scratch("col", (B * T2 * F2, 9 * d))
        ops.im2col_s2(y1, col)
        self._col_of = (self._sig, tag)         # the patch matrix stays valid until a backward turns it into dcol (or another forward reuses it)
        y2 = self.buf(tag + ".y2", (B * T2 * F2, d))
        ops.gemm(col, w2p.view(d, 9 * d), y2, bias=st.p(prefix + ".conv.2.bias"), relu=True, mode=self.mode)
        elin = self.buf(tag + ".elin", (B * T2, d))
        ops.gemm(y2.view(B * T2, F2 * d), woutp.view(d, F2 * d), elin, bias=st.p(out_name + ".bias"), mode=self.mode)
        return elin

    def _conv2d_sub_bwd(self, delin: torch.Tensor, xs: torch.Tensor, prefix: str, out_name: str, tag: str) -> None:
        st = self.store
        B, T, idim = xs.shape
        d = st.p(prefix + ".conv.0.weight").shape[0]
        T1, F1 = (T - 1) // 2, (idim - 1) // 2
        T2, F2 = (T1 - 1) // 2, (F1 - 1) // 2
        mode = self.mode
        y2 = self.buf(tag + ".y2", (B * T2 * F2, d))
        woutp = self.buf(f"w.{tag}.outp", (d, F2, d))
        gwoutp = self._scratch("g.woutp", (d, F2 * d), _f32)
        dy2 = self._scratch("g.y2", (B * T2 * F2, d))
        ops.gemm(delin.view(B * T2, d).t(), y2.view(B * T2, F2 * d).t(), gwoutp, mode=mode)
        ops.transpose_last2(gwoutp, st.g(out_name + ".weight"), d, F2, d, accumulate=True)
        ops.colsum(delin.view(B * T2, d), st.g(out_name + ".bias"))
        ops.gemm(delin.view(B * T2, d), woutp.view(d, F2 * d).t(), dy2.view(B * T2, F2 * d), mode=mode, gate=y2.view(B * T2, F2 * d))
        w2p = self.buf(f"w.{tag}.conv2p", (d, 9, d))
        col = self._scratch("col", (B * T2 * F2, 9 * d))
        y1 = self.buf(tag + ".y1", (B, T1, F1, d))
        if getattr(self, "_col_of", None) != (self._sig, tag):
            ops.im2col_s2(y1, col)              # normally still there from this step's forward
        self._col_of = None
        gw2p = self._scratch("g.w2p", (d, 9 * d), _f32)
        ops.gemm(dy2.t(), col.t(), gw2p, mode=mode)
        ops.transpose_last2(gw2p, st.g(prefix + ".conv.2.weight"), d, 9, d, accumulate=True)
        ops.colsum(dy2, st.g(prefix + ".conv.2.bias"))
        dcol = col
        ops.gemm(dy2, w2p.view(d, 9 * d).t(), dcol, mode=mode)
        dy1 = self._scratch("g.y1", (B, T1, F1, d))
        ops.col2im_s2_relu(dcol, y1, dy1)          # scatter-add + conv.0's ReLU' in one pass
        self._conv1_bwd(xs, dy1, prefix, tag)
