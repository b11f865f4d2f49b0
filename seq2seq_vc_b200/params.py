"""Flat parameter / gradient / optimiser-state store with reference state-dict names.

All trainable tensors of a model live in ONE float32 buffer ``P`` (gradients ``G``, Adam moments
``M``/``V``, optional bf16 shadow ``P16``), so that gradient clipping, Adam and the data-parallel
gradient all-reduce are single flat-buffer operations (reference: trainers/ar_vc.py:99-107,
bin/vc_train.py:423-431).  Named views keep the reference's state-dict keys and shapes
(SURVEY.md section 8b "state dict"), so reference checkpoints load unchanged.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch

ALIGN = 64  # elements; 256 B for float32, 128 B for the bf16 shadow


class ParamStore:
    def __init__(self, groups: Sequence[Sequence[Tuple[str, Tuple[int, ...]]]], device, bf16_shadow: bool):
        """groups: lists of (name, shape); tensors of one group are laid out back to back (no
        padding inside a group) so that e.g. linear_q/k/v weights form one (3d, d) matrix."""
        self.offsets: Dict[str, Tuple[int, Tuple[int, ...]]] = {}
        off = 0
        for grp in groups:
            off = (off + ALIGN - 1) // ALIGN * ALIGN
            for name, shape in grp:
                n = 1
                for s in shape:
                    n *= s
                self.offsets[name] = (off, tuple(shape))
                off += n
        self.numel = (off + ALIGN - 1) // ALIGN * ALIGN
        f32 = dict(dtype=torch.float32, device=device)
        self.P = torch.zeros(self.numel, **f32)
        self.G = torch.zeros(self.numel, **f32)
        self.M = torch.zeros(self.numel, **f32)
        self.V = torch.zeros(self.numel, **f32)
        self.P16 = torch.zeros(self.numel, dtype=torch.bfloat16, device=device) if bf16_shadow else None

    def _view(self, buf, name):
        off, shape = self.offsets[name]
        n = 1
        for s in shape:
            n *= s
        return buf[off:off + n].view(shape)

    def p(self, name):
        return self._view(self.P, name)

    def g(self, name):
        return self._view(self.G, name)

    def m(self, name):
        """Adam first-moment view of one parameter."""
        return self._view(self.M, name)

    def v(self, name):
        """Adam second-moment view of one parameter."""
        return self._view(self.V, name)

    def p16(self, name):
        return self._view(self.P16, name)

    def span(self, buf, names: List[str], shape):
        """One view over several consecutive tensors of a group (e.g. fused QKV weight)."""
        off0, _ = self.offsets[names[0]]
        n = 1
        for s in shape:
            n *= s
        return buf[off0:off0 + n].view(shape)

    def names(self):
        return list(self.offsets.keys())
