"""Drop-in mirror of ``MT4MTLKD/Temporal_mstct/MSTCT/TS_Mixer.py``."""
from __future__ import annotations

import torch
from torch import nn

from .. import ops
from ..layout import SeqLayout


class linear_layer(nn.Module):
    """TS_Mixer.py:6-15."""

    def __init__(self, input_dim=2048, embed_dim=512):
        super().__init__()
        self.proj = nn.Linear(input_dim, embed_dim)

    def forward(self, x):
        """x: (B, C, T) -> (B, T, embed)."""
        B, Cc, T = x.shape
        lay = SeqLayout.uniform(B, T, x.device)
        y = ops.tap_linear(lay.pad_bct(x), self.proj.weight, self.proj.bias, lay)
        return lay.as_btc(y, self.proj.out_features)


class Temporal_Mixer(nn.Module):
    """TS_Mixer.py:28-84.  With equal lengths the linear interpolations are identities (all stages have stride 1),
    and only the ivt branch reaches the output, so
        _f{s}_ivt = (linear7/8/9 + linear1/2/3 + linear4/5/6)(_f4) + 3 * _f{s}
    which is evaluated as ONE GEMM per scale with summed weights and a residual (autograd distributes the
    gradient back onto the nine Conv1d parameters)."""

    def __init__(self, inter_channels, embedding_dim):
        super().__init__()
        c1, c2, c3, c4 = inter_channels
        self.linear_f4 = linear_layer(input_dim=c4, embed_dim=embedding_dim)
        self.linear_f3 = linear_layer(input_dim=c3, embed_dim=embedding_dim)
        self.linear_f2 = linear_layer(input_dim=c2, embed_dim=embedding_dim)
        self.linear_f1 = linear_layer(input_dim=c1, embed_dim=embedding_dim)
        for i in range(1, 10):
            setattr(self, f"linear{i}", nn.Conv1d(embedding_dim, embedding_dim, kernel_size=1))
        self.embedding_dim = embedding_dim

    def _packed(self, f_rows, lay):
        """f_rows: four packed (rows, C_s).  Returns the packed concat (rows, 4 * embedding_dim)."""
        f1, f2, f3, f4 = f_rows
        _f4 = ops.tap_linear(f4, self.linear_f4.proj.weight, self.linear_f4.proj.bias, lay)
        pieces = [_f4]
        for f, lin, (a, b, c) in ((f3, self.linear_f3, (1, 4, 7)), (f2, self.linear_f2, (2, 5, 8)),
                                  (f1, self.linear_f1, (3, 6, 9))):
            three_f = ops.tap_linear(f, 3.0 * lin.proj.weight, 3.0 * lin.proj.bias, lay)
            la, lb, lc = getattr(self, f"linear{a}"), getattr(self, f"linear{b}"), getattr(self, f"linear{c}")
            w = (la.weight + lb.weight + lc.weight)
            bias = la.bias + lb.bias + lc.bias
            pieces.append(ops.tap_linear(_f4, w, bias, lay, residual=three_f))
        return torch.cat(pieces, dim=1)

    def forward(self, x):
        """x: four (B, C_s, T) -> (B, 4 * embedding_dim, T)."""
        B, _, T = x[0].shape
        if any(t.shape[2] != T for t in x):
            raise NotImplementedError("scales of different length do not occur (every stage has stride 1)")
        lay = SeqLayout.uniform(B, T, x[0].device)
        cat = self._packed([lay.pad_bct(t) for t in x], lay)
        return lay.as_bct(cat, cat.shape[1])
