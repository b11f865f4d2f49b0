"""CPU restatement of Conv2dSubsampling2 / 6 / 8 (seq2seq_vc/modules/transformer/subsampling.py:108-279) with the default
PositionalEncoding (layers/positional_encoding.py:14-70), dropout off.

TEST INFRASTRUCTURE ONLY.  Pinned against the live reference modules (tests/test_subsampling_host_logic.py, build container) and
against golden vectors dumped from them (oracle/gen_golden.py subsampling_tiny -> tests/golden/subsampling_tiny.npz)."""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

# (kernel, stride) of the convolutions after the first Conv2d(1, C, 3, 2), and the mask slicing of forward()
LATER = {2: ((3, 1),), 6: ((5, 3),), 8: ((3, 2), (3, 2))}
MASKS = {2: ((None, -2, 2), (None, -2, 1)), 6: ((None, -2, 2), (None, -4, 3)), 8: ((None, -2, 2), (None, -2, 2), (None, -2, 2))}


def sinusoid(T: int, d: int) -> torch.Tensor:
    pos = torch.arange(0, T, dtype=torch.float32).unsqueeze(1)
    div = torch.exp(torch.arange(0, d, 2, dtype=torch.float32) * -(math.log(10000.0) / d))
    pe = torch.zeros(T, d)
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe


def conv2d_subsampling(sd, n: int, x: torch.Tensor, x_mask=None):
    """sd: conv.0.* , conv.2.* (, conv.4.*), out.0.*; x (B, T, idim) -> (y (B, T', odim), mask')."""
    h = F.relu(F.conv2d(x.unsqueeze(1), sd["conv.0.weight"], sd["conv.0.bias"], stride=2))            # subsampling.py:129 / 187 / 236
    for i, (k, s) in enumerate(LATER[n]):
        h = F.relu(F.conv2d(h, sd[f"conv.{2 * i + 2}.weight"], sd[f"conv.{2 * i + 2}.bias"], stride=s))
    b, c, t, f = h.shape
    y = F.linear(h.transpose(1, 2).contiguous().view(b, t, c * f), sd["out.0.weight"], sd["out.0.bias"])   # :153-156
    d = y.shape[2]
    y = y * math.sqrt(d) + sinusoid(t, d)[None]                                                            # positional_encoding.py:59-70, dropout off
    if x_mask is not None:
        for a, e, st in MASKS[n]:
            x_mask = x_mask[:, :, a:e:st]
    return y, x_mask
