"""autograd Functions over the MS-TCT kernels (csrc/mstct.cu); tensors are packed time-major rows."""
from __future__ import annotations

import ctypes as C

import torch

from .. import _lib
from ..layout import SeqLayout


def _like(x, lay):
    """Output buffer shaped like x: the kernels write every frame of every sequence, so zero-filling is needed only
    when the layout has padding rows."""
    return torch.empty_like(x) if lay.frames == lay.rows else torch.zeros_like(x)


def _c(t):
    t = t if t.dtype == torch.float32 else t.float()
    return t if t.is_contiguous() else t.contiguous()


class LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, lay, eps):
        lib = _lib.load()
        x = _c(x)
        rows, Cc = x.shape
        y = _like(x, lay)
        mean = torch.empty(rows, device=x.device, dtype=torch.float32)
        rstd = torch.empty(rows, device=x.device, dtype=torch.float32)
        w, b = _c(weight.detach()), _c(bias.detach())
        _lib.check(lib.tcn_layernorm_fwd(_lib.ptr(x), Cc, _lib.ptr(y), Cc, _lib.ptr(w), _lib.ptr(b), _lib.ptr(mean),
                                         _lib.ptr(rstd), _lib.ptr(lay.meta), rows, Cc, float(eps), _lib.stream_ptr()),
                   "tcn_layernorm_fwd")
        ctx.save_for_backward(x, w, mean, rstd)
        ctx.lay = lay
        return y

    @staticmethod
    def backward(ctx, gy):
        lib = _lib.load()
        x, w, mean, rstd = ctx.saved_tensors
        gy = _c(gy)
        rows, Cc = x.shape
        dx = _like(x, ctx.lay)
        dg = torch.zeros(Cc, device=x.device, dtype=torch.float32)
        db = torch.zeros(Cc, device=x.device, dtype=torch.float32)
        _lib.check(lib.tcn_layernorm_bwd(_lib.ptr(x), Cc, _lib.ptr(gy), Cc, _lib.ptr(dx), Cc, _lib.ptr(w),
                                         _lib.ptr(mean), _lib.ptr(rstd), _lib.ptr(dg), _lib.ptr(db),
                                         _lib.ptr(ctx.lay.meta), rows, Cc, _lib.stream_ptr()), "tcn_layernorm_bwd")
        return dx, dg, db, None, None


def layer_norm(x, weight, bias, lay: SeqLayout, eps=1e-5):
    return LayerNormFn.apply(x, weight, bias, lay, eps)


def _attn_args(q, kv, o, lse, lay, heads, dout=None, dq=None, dkv=None, delta=None):
    d = q.shape[1]
    hd = d // heads
    a = _lib.AttnArgs()
    a.q, a.ldq = _lib.ptr(q), d
    a.k, a.ldk = kv.data_ptr(), 2 * d
    a.v, a.ldv = kv.data_ptr() + d * 4, 2 * d
    a.o, a.ldo, a.lse = _lib.ptr(o), d, _lib.ptr(lse)
    if dout is not None:
        a.dout, a.lddo = _lib.ptr(dout), d
        a.dq, a.lddq = _lib.ptr(dq), d
        a.dk, a.lddk = dkv.data_ptr(), 2 * d
        a.dv, a.lddv = dkv.data_ptr() + d * 4, 2 * d
        a.delta = _lib.ptr(delta)
    a.seq_lo, a.seq_len, a.nseq, a.max_len = _lib.ptr(lay.seq_lo), _lib.ptr(lay.seq_len), lay.num_seqs, lay.max_len
    a.heads, a.head_dim, a.scale = heads, hd, float(hd) ** -0.5
    return a


def _attn_tc_args(q, kv, o, p, lay, heads, tmax, dout=None, dq=None, dkv=None, dp=None):
    d = q.shape[1]
    hd = d // heads
    a = _lib.AttnTcArgs()
    a.q, a.ldq = _lib.ptr(q), d
    a.k, a.ldk = kv.data_ptr(), 2 * d
    a.v, a.ldv = kv.data_ptr() + d * 4, 2 * d
    a.o, a.ldo, a.p = _lib.ptr(o), d, _lib.ptr(p)
    if dout is not None:
        a.dout, a.lddo = _lib.ptr(dout), d
        a.dq, a.lddq = _lib.ptr(dq), d
        a.dk, a.lddk = dkv.data_ptr(), 2 * d
        a.dv, a.lddv = dkv.data_ptr() + d * 4, 2 * d
        a.dp = _lib.ptr(dp)
    a.seq_lo, a.seq_len, a.nseq, a.rows = _lib.ptr(lay.seq_lo), _lib.ptr(lay.seq_len), lay.num_seqs, q.shape[0]
    a.heads, a.head_dim, a.tmax, a.scale = heads, hd, tmax, float(hd) ** -0.5
    return a


def _attn_tc_ok(q, lay, heads):
    """tcgen05 path (csrc/attention_tc.cu): windows of at most 256 frames, head dim a multiple of 4."""
    import os

    if os.environ.get("TCN_NO_ATTN_TC") is not None:
        return False
    d = q.shape[1]
    return bool(_lib.load().tcn_attn_tc_supported(lay.max_len, heads, d // heads, d, 2 * d))


class AttentionTcFn(torch.autograd.Function):
    """o = softmax(q k^T / sqrt(hd)) v per (window, head) on tcgen05: S = q k^T, P = softmax(S) (kept for the backward
    pass: nseq * heads blocks of tmax x tmax), o = P v; backward dV = P^T dO, dP = dO v^T, dS, dQ = dS k, dK = dS^T q."""

    @staticmethod
    def forward(ctx, q, kv, lay, heads):
        lib = _lib.load()
        q, kv = _c(q), _c(kv)
        o = _like(q, lay)
        tmax = (lay.max_len + 31) // 32 * 32
        p = torch.empty(lay.num_seqs * heads * tmax, tmax, device=q.device, dtype=torch.float32)
        a = _attn_tc_args(q, kv, o, p, lay, heads, tmax)
        _lib.check(lib.tcn_attn_fwd_tc(C.byref(a), _lib.stream_ptr()), "tcn_attn_fwd_tc")
        ctx.save_for_backward(q, kv, p)
        ctx.lay, ctx.heads, ctx.tmax = lay, heads, tmax
        return o

    @staticmethod
    def backward(ctx, go):
        lib = _lib.load()
        q, kv, p = ctx.saved_tensors
        go = _c(go)
        dq, dkv = _like(q, ctx.lay), _like(kv, ctx.lay)
        dp = torch.empty_like(p)
        a = _attn_tc_args(q, kv, dq, p, ctx.lay, ctx.heads, ctx.tmax, go, dq, dkv, dp)
        _lib.check(lib.tcn_attn_bwd_tc(C.byref(a), _lib.stream_ptr()), "tcn_attn_bwd_tc")
        return dq, dkv, None, None


class AttentionFn(torch.autograd.Function):
    """o = softmax(q k^T / sqrt(hd)) v per (sequence, head); q (rows, d), kv (rows, 2d) = [k | v]."""

    @staticmethod
    def forward(ctx, q, kv, lay, heads):
        lib = _lib.load()
        q, kv = _c(q), _c(kv)
        o = _like(q, lay)
        lse = torch.zeros(q.shape[0], heads, device=q.device, dtype=torch.float32)
        a = _attn_args(q, kv, o, lse, lay, heads)
        _lib.check(lib.tcn_attn_fwd(C.byref(a), _lib.stream_ptr()), "tcn_attn_fwd")
        ctx.save_for_backward(q, kv, o, lse)
        ctx.lay, ctx.heads = lay, heads
        return o

    @staticmethod
    def backward(ctx, go):
        lib = _lib.load()
        q, kv, o, lse = ctx.saved_tensors
        go = _c(go)
        dq, dkv = _like(q, ctx.lay), _like(kv, ctx.lay)
        delta = torch.empty_like(lse)
        a = _attn_args(q, kv, o, lse, ctx.lay, ctx.heads, go, dq, dkv, delta)
        _lib.check(lib.tcn_attn_bwd(C.byref(a), _lib.stream_ptr()), "tcn_attn_bwd")
        return dq, dkv, None, None


def attention(q, kv, lay: SeqLayout, heads: int):
    if _attn_tc_ok(q, lay, heads):
        return AttentionTcFn.apply(q, kv, lay, heads)
    return AttentionFn.apply(q, kv, lay, heads)   # longer windows / odd head dims: the mma.sync kernels


class DwConvGeluFn(torch.autograd.Function):
    """gelu(depthwise Conv1d(k=3, pad=1) over time); x (rows, C), weight (C, 1, 3), bias (C,)."""

    @staticmethod
    def forward(ctx, x, weight, bias, lay):
        lib = _lib.load()
        x = _c(x)
        rows, Cc = x.shape
        w, b = _c(weight.detach()).view(Cc, 3), _c(bias.detach())
        y = _like(x, lay)
        _lib.check(lib.tcn_dwconv_gelu_fwd(_lib.ptr(x), _lib.ptr(y), _lib.ptr(w), _lib.ptr(b), _lib.ptr(lay.meta), rows,
                                           Cc, _lib.stream_ptr()), "tcn_dwconv_gelu_fwd")
        ctx.save_for_backward(x, w, b)
        ctx.lay, ctx.wshape = lay, weight.shape
        return y

    @staticmethod
    def backward(ctx, gy):
        lib = _lib.load()
        x, w, b = ctx.saved_tensors
        gy = _c(gy)
        rows, Cc = x.shape
        du, dx = _like(x, ctx.lay), _like(x, ctx.lay)
        dw = torch.zeros(Cc, 3, device=x.device, dtype=torch.float32)
        db = torch.zeros(Cc, device=x.device, dtype=torch.float32)
        _lib.check(lib.tcn_dwconv_gelu_bwd(_lib.ptr(x), _lib.ptr(gy), _lib.ptr(du), _lib.ptr(dx), _lib.ptr(w),
                                           _lib.ptr(b), _lib.ptr(dw), _lib.ptr(db), _lib.ptr(ctx.lay.meta), rows, Cc,
                                           _lib.stream_ptr()), "tcn_dwconv_gelu_bwd")
        return dx, dw.view(ctx.wshape), db, None


def dwconv_gelu(x, weight, bias, lay: SeqLayout):
    return DwConvGeluFn.apply(x, weight, bias, lay)


class DropoutRowsFn(torch.autograd.Function):
    """nn.Dropout on packed rows with the counter-based mask of the kernels (regenerated in backward)."""

    @staticmethod
    def forward(ctx, x, p, seed, stream_id):
        from .. import ops

        ctx.key = (p, seed, stream_id)
        return ops.dropout_apply(_c(x), p, seed, stream_id)

    @staticmethod
    def backward(ctx, gy):
        from .. import ops

        p, seed, sid = ctx.key
        return ops.dropout_apply(_c(gy), p, seed, sid), None, None, None


def dropout_rows(x, p, training, stream_id):
    if not training or p <= 0:
        return x
    from .. import ops

    return DropoutRowsFn.apply(x, float(p), ops.new_seed(), int(stream_id))
