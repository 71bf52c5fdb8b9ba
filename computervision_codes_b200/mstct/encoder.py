"""Drop-in mirror of ``MT4MTLKD/Temporal_mstct/MSTCT/Temporal_Encoder.py``: same class names, constructor
arguments, parameter names / shapes (state_dict) and initialisation order; forward runs on the CUDA kernels
(Linear / Conv1d on the tcgen05 tap GEMM, LayerNorm / attention / depthwise-conv+GELU in csrc/mstct.cu).
Internally activations are packed time-major rows; the (B, N, C) / (B, C, T) tensors of the reference are
converted at the module boundary only."""
from __future__ import annotations

import math

import torch
from torch import nn

from .. import ops
from ..layout import SeqLayout
from . import functional as Fn

_stream = [1000]


def _sid():
    _stream[0] += 1
    return _stream[0]


def _init_weights(m):
    """Same initialisation rule (and RNG consumption) as the reference's ``_init_weights`` (:19-32)."""
    if isinstance(m, nn.Linear):
        nn.init.trunc_normal_(m.weight, std=.02)
        if m.bias is not None:
            nn.init.constant_(m.bias, 0)
    elif isinstance(m, nn.LayerNorm):
        nn.init.constant_(m.bias, 0)
        nn.init.constant_(m.weight, 1.0)
    elif isinstance(m, nn.Conv1d):
        fan_out = (m.kernel_size[0] * m.out_channels) // m.groups
        m.weight.data.normal_(0, math.sqrt(2.0 / fan_out))
        if m.bias is not None:
            m.bias.data.zero_()


def _rows_from_bnc(x):
    """(B, N, C) -> (layout, packed rows)."""
    B, N, Cc = x.shape
    lay = SeqLayout.uniform(B, N, x.device)
    Tp = lay.rows // B
    buf = torch.zeros(B, Tp, Cc, device=x.device, dtype=torch.float32)
    buf[:, :N] = x
    return lay, buf.view(lay.rows, Cc)


class Local_Relational_Block(nn.Module):
    """Temporal_Encoder.py:5-43: Linear(d, 8d) -> depthwise Conv1d(k=3) -> GELU -> Linear(8d, d)."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.linear1 = nn.Linear(in_features, hidden_features)
        self.TC = nn.Conv1d(hidden_features, hidden_features, 3, 1, 1, bias=True, groups=hidden_features)
        self.act = act_layer()
        self.linear2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)
        if drop != 0.:
            raise NotImplementedError("the reference always builds this block with drop = 0")
        self.apply(_init_weights)

    def _packed(self, x_rows, lay, residual=None):
        h = ops.tap_linear(x_rows, self.linear1.weight, self.linear1.bias, lay)
        h = Fn.dwconv_gelu(h, self.TC.weight, self.TC.bias, lay)
        return ops.tap_linear(h, self.linear2.weight, self.linear2.bias, lay, residual=residual)

    def forward(self, x):
        lay, rows = _rows_from_bnc(x)
        return lay.as_btc(self._packed(rows, lay), x.shape[2])


class Global_Relational_Block(nn.Module):
    """Temporal_Encoder.py:46-88: multi-head self-attention over the frames of a window."""

    def __init__(self, dim, num_heads=8):
        super().__init__()
        assert dim % num_heads == 0, f"dim {dim} should be divided by num_heads {num_heads}."
        self.dim = dim
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.q = nn.Linear(dim, dim)
        self.kv = nn.Linear(dim, dim * 2)
        self.proj = nn.Linear(dim, dim)
        self.apply(_init_weights)

    def _packed(self, x_rows, lay, residual=None):
        q = ops.tap_linear(x_rows, self.q.weight, self.q.bias, lay)
        kv = ops.tap_linear(x_rows, self.kv.weight, self.kv.bias, lay)
        o = Fn.attention(q, kv, lay, self.num_heads)
        return ops.tap_linear(o, self.proj.weight, self.proj.bias, lay, residual=residual)

    def forward(self, x):
        lay, rows = _rows_from_bnc(x)
        return lay.as_btc(self._packed(rows, lay), x.shape[2])


class GLRBlock(nn.Module):
    """Temporal_Encoder.py:91-126: x + GRB(LN(x)); x + LRB(LN(x))."""

    def __init__(self, dim, num_heads, mlp_ratio=4., drop=0., act_layer=nn.GELU, norm_layer=nn.LayerNorm):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.Global_Relational_Block = Global_Relational_Block(dim, num_heads=num_heads)
        self.norm2 = norm_layer(dim)
        self.Local_Relational_Block = Local_Relational_Block(in_features=dim, hidden_features=int(dim * mlp_ratio),
                                                             act_layer=act_layer, drop=drop)
        self.apply(_init_weights)

    def _packed(self, x_rows, lay):
        n1 = Fn.layer_norm(x_rows, self.norm1.weight, self.norm1.bias, lay, self.norm1.eps)
        x_rows = self.Global_Relational_Block._packed(n1, lay, residual=x_rows)
        n2 = Fn.layer_norm(x_rows, self.norm2.weight, self.norm2.bias, lay, self.norm2.eps)
        return self.Local_Relational_Block._packed(n2, lay, residual=x_rows)

    def forward(self, x):
        lay, rows = _rows_from_bnc(x)
        return lay.as_btc(self._packed(rows, lay), x.shape[2])


class Temporal_Merging_Block(nn.Module):
    """Temporal_Encoder.py:129-161: Conv1d(k=3, stride 1, pad 1) over time, then LayerNorm."""

    def __init__(self, kernel_size=3, stride=1, in_chans=1024, embed_dim=256):
        super().__init__()
        if kernel_size != 3 or stride != 1:
            raise NotImplementedError("the reference instantiates kernel_size=3, stride=1 only (:171-195)")
        self.proj = nn.Conv1d(in_chans, embed_dim, kernel_size=kernel_size, stride=stride, padding=kernel_size // 2)
        self.norm = nn.LayerNorm(embed_dim)
        self.apply(_init_weights)

    def _packed(self, x_rows, lay, in_drop_p=0.0, seed=0, stream_id=0):
        y = ops.tap_linear(x_rows, self.proj.weight, self.proj.bias, lay, shifts=(-1, 0, 1), in_drop_p=in_drop_p,
                           seed=seed, stream_id=stream_id)
        return Fn.layer_norm(y, self.norm.weight, self.norm.bias, lay, self.norm.eps)

    def forward(self, x):
        """x: (B, C_in, T) -> (B, T, embed_dim)."""
        B, Cc, T = x.shape
        lay = SeqLayout.uniform(B, T, x.device)
        return lay.as_btc(self._packed(lay.pad_bct(x), lay), self.proj.out_channels)


class TemporalEncoder(nn.Module):
    """Temporal_Encoder.py:164-256: four stages of (merging block, num_block GLR blocks, LayerNorm)."""

    def __init__(self, in_feat_dim=1024, embed_dims=[256, 384, 576, 864], num_head=8, mlp_ratio=8,
                 norm_layer=nn.LayerNorm, num_block=3):
        super().__init__()
        dims = [in_feat_dim] + list(embed_dims)
        for s in range(1, 5):
            setattr(self, f"Temporal_Merging_Block{s}",
                    Temporal_Merging_Block(kernel_size=3, stride=1, in_chans=dims[s - 1], embed_dim=dims[s]))
            setattr(self, f"block{s}", nn.ModuleList([GLRBlock(dim=dims[s], num_heads=num_head, mlp_ratio=mlp_ratio,
                                                                norm_layer=norm_layer) for _ in range(num_block)]))
            setattr(self, f"norm{s}", norm_layer(dims[s]))
        self.embed_dims = list(embed_dims)
        self._in_stream = _sid()
        self.apply(_init_weights)

    def freeze_init_emb(self):
        self.Temporal_Merging_Block1.requires_grad = False

    def _packed(self, x_rows, lay, in_drop_p=0.0):
        """Returns the four stage outputs as packed rows (rows, C_s)."""
        outs = []
        seed = ops.new_seed() if in_drop_p > 0 else 0
        for s in range(1, 5):
            merge = getattr(self, f"Temporal_Merging_Block{s}")
            x_rows = merge._packed(x_rows, lay, in_drop_p if s == 1 else 0.0, seed, self._in_stream)
            for blk in getattr(self, f"block{s}"):
                x_rows = blk._packed(x_rows, lay)
            norm = getattr(self, f"norm{s}")
            x_rows = Fn.layer_norm(x_rows, norm.weight, norm.bias, lay, norm.eps)
            outs.append(x_rows)
        return outs

    def forward(self, x):
        """x: (B, in_feat_dim, T) -> list of four (B, C_s, T)."""
        B, Cc, T = x.shape
        lay = SeqLayout.uniform(B, T, x.device)
        outs = self._packed(lay.pad_bct(x), lay)
        return [lay.as_bct(o, o.shape[1]).contiguous() for o in outs]
