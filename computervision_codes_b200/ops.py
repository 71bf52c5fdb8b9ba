"""Thin Python wrappers over the C ABI (one call = one kernel launch) and the autograd Functions
built from them.  All tensors are fp32 CUDA tensors in the packed time-major layout (layout.py)."""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _lib
from .layout import SeqLayout, round_up


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


# ------------------------------------------------------------------------------------ raw launches
def _rows_out(lay: SeqLayout, cols: int, device, written_cols=None) -> torch.Tensor:
    """Output buffer (lay.rows, cols).  The kernels write the frames of every sequence and nothing else, so the
    buffer is zero-filled only when the layout has padding rows or the kernel leaves pad columns untouched."""
    if lay.frames == lay.rows and (written_cols is None or written_cols == cols):
        return torch.empty(lay.rows, cols, device=device, dtype=torch.float32)
    return torch.zeros(lay.rows, cols, device=device, dtype=torch.float32)


def prep_weight(w: torch.Tensor, transpose: bool = False) -> torch.Tensor:
    """torch Conv1d / Linear weight (n_out, c_in[, ntaps]) -> fragment-ordered hi/lo buffer."""
    lib = _lib.load()
    w = _f32c(w.detach())
    n_out, c_in = w.shape[0], w.shape[1]
    ntaps = w.shape[2] if w.dim() == 3 else 1
    n = lib.tcn_prep_weight_floats(n_out, c_in, ntaps, int(transpose))
    wf = torch.empty(n, device=w.device, dtype=torch.float32)
    _lib.check(lib.tcn_prep_weight(_lib.ptr(w), n_out, c_in, ntaps, int(transpose), _lib.ptr(wf), _lib.stream_ptr()),
               "tcn_prep_weight")
    return wf


def tapgemm(x, wf, lay: SeqLayout, c_in, n_out, shifts=(0,), bias=None, out=None, ldy=None, residual=None,
            relu_mask=None, relu=False, drop_p=0.0, seed=0, stream_id=0, x_unpadded=False, colscale=None,
            in_drop_p=0.0):
    lib = _lib.load()
    assert x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 2
    if ldy is None:
        ldy = round_up(n_out, 4)
    if out is None:
        out = _rows_out(lay, ldy, x.device, written_cols=n_out)
    a = _lib.TapGemmArgs()
    a.x, a.ldx, a.x_unpadded = _lib.ptr(x), x.shape[1], int(x_unpadded)
    a.colscale, a.colscale_ld = _lib.ptr(colscale), (colscale.shape[1] if colscale is not None else 0)
    a.wf, a.bias = _lib.ptr(wf), _lib.ptr(bias)
    a.y, a.ldy = _lib.ptr(out), out.shape[1]
    a.residual, a.ldr = _lib.ptr(residual), (residual.shape[1] if residual is not None else 0)
    a.relu_mask, a.ldm = _lib.ptr(relu_mask), (relu_mask.shape[1] if relu_mask is not None else 0)
    a.meta, a.nblk = _lib.ptr(lay.meta), lay.nblk
    a.c_in, a.n_out, a.ntaps = c_in, n_out, len(shifts)
    for i, s in enumerate(shifts):
        a.shift[i] = int(s)
    a.relu = int(relu)
    a.drop_p, a.drop_seed, a.drop_stream = float(drop_p), int(seed) & 0xFFFFFFFF, int(stream_id) & 0xFFFFFFFF
    a.in_drop_p = float(in_drop_p)
    _lib.check(lib.tcn_tapgemm(C.byref(a), _lib.stream_ptr()), "tcn_tapgemm")
    return out


def wgrad(g, x, lay: SeqLayout, n_out, c_in, shifts, dw, db=None, x_unpadded=False, colscale=None, g_drop_p=0.0,
          seed=0, stream_id=0):
    """dw (n_out, c_in, ntaps) += ..., db (n_out,) += ...  (accumulating, fp32 atomics)."""
    lib = _lib.load()
    assert g.is_contiguous() and x.is_contiguous() and dw.is_contiguous()
    a = _lib.WgradArgs()
    a.g, a.ldg, a.g_cols = _lib.ptr(g), g.shape[1], min(round_up(n_out, 4), g.shape[1])
    a.x, a.ldx, a.x_unpadded = _lib.ptr(x), x.shape[1], int(x_unpadded)
    a.colscale, a.colscale_ld = _lib.ptr(colscale), (colscale.shape[1] if colscale is not None else 0)
    a.meta, a.nblk = _lib.ptr(lay.meta), lay.nblk
    a.n_out, a.c_in, a.ntaps = n_out, c_in, len(shifts)
    for i, s in enumerate(shifts):
        a.shift[i] = int(s)
    a.dw, a.db = _lib.ptr(dw), _lib.ptr(db)
    a.g_drop_p, a.drop_seed, a.drop_stream = float(g_drop_p), int(seed) & 0xFFFFFFFF, int(stream_id) & 0xFFFFFFFF
    _lib.check(lib.tcn_wgrad(C.byref(a), _lib.stream_ptr()), "tcn_wgrad")


def _wgrad_tc_args(g, x, lay: SeqLayout, n_out, c_in, shifts, dw, db=None, x_unpadded=False, colscale=None,
                   g_drop_p=0.0, seed=0, stream_id=0):
    assert g.is_contiguous() and x.is_contiguous() and dw.is_contiguous()
    a = _lib.WgradTcArgs()
    a.g, a.ldg, a.g_cols, a.g_rows = _lib.ptr(g), g.shape[1], min(round_up(n_out, 4), g.shape[1]), g.shape[0]
    a.x, a.ldx, a.x_rows, a.x_unpadded = _lib.ptr(x), x.shape[1], x.shape[0], int(x_unpadded)
    a.colscale, a.colscale_ld = _lib.ptr(colscale), (colscale.shape[1] if colscale is not None else 0)
    a.meta, a.nblk = _lib.ptr(lay.meta), lay.nblk
    a.n_out, a.c_in, a.ntaps = n_out, c_in, len(shifts)
    for i, s in enumerate(shifts):
        a.shift[i] = int(s)
    a.dw, a.db = _lib.ptr(dw), _lib.ptr(db)
    a.g_drop_p, a.drop_seed, a.drop_stream = float(g_drop_p), int(seed) & 0xFFFFFFFF, int(stream_id) & 0xFFFFFFFF
    return a


def wgrad_tc(g, x, lay: SeqLayout, n_out, c_in, shifts, dw, db=None, x_unpadded=False, colscale=None, g_drop_p=0.0,
             seed=0, stream_id=0):
    """tcgen05/TMA weight gradient (same contract as wgrad)."""
    lib = _lib.load()
    a = _wgrad_tc_args(g, x, lay, n_out, c_in, shifts, dw, db, x_unpadded, colscale, g_drop_p, seed, stream_id)
    _lib.check(lib.tcn_wgrad_tc(C.byref(a), _lib.stream_ptr()), "tcn_wgrad_tc")


def wgrad_tc_layer_pair(gu, x, gy, h, lay: SeqLayout, shifts, gw1, gb1, gw2, gb2, drop_p=0.0, seed=0, stream_id=0):
    """Both weight gradients of a residual layer in one launch: gW1 / gb1 += (gu, x over the taps), gW2 / gb2 += (gv, h)
    with gv = keep * gy / (1 - p) regenerated from (seed, stream_id)."""
    lib = _lib.load()
    Cc = gu.shape[1]
    a1 = _wgrad_tc_args(gu, x, lay, Cc, Cc, shifts, gw1, gb1)
    a2 = _wgrad_tc_args(gy, h, lay, Cc, Cc, (0,), gw2, gb2, g_drop_p=drop_p, seed=seed, stream_id=stream_id)
    _lib.check(lib.tcn_wgrad_tc_pair(C.byref(a1), C.byref(a2), _lib.stream_ptr()), "tcn_wgrad_tc_pair")


def layer_wgrad_kernel_name() -> str:
    """Name of the kernel layer_wgrad() launches (for bench.py's roofline block)."""
    return "wgrad_layer_kernel"


_wl_workspace: dict = {}


def layer_wgrad(gu, x, gy, h, lay: SeqLayout, shifts, gw1, gb1, gw2, gb2, drop_p=0.0, seed=0, stream_id=0, masks=None,
                skip_reduce=False):
    """All weight / bias gradients of one 64-channel residual layer, deterministic (csrc/wgrad_layer.cu): accumulates
    gW1 (C, C, 3) / gb1 over the taps of (gu, x) and gW2 (C, C, 1) / gb2 of (gv, h), gv = keep * gy / (1 - p) with the
    keep bits taken from `masks` (the words layer_fwd_tc saved) or regenerated from (seed, stream_id)."""
    lib = _lib.load()
    assert gu.shape[1] == 64 and all(t.is_contiguous() for t in (gu, x, gy, h, gw1, gw2))
    dev = gu.device
    ws = _wl_workspace.get(dev)
    need = int(lib.tcn_wgrad_layer_workspace_bytes(lay.nblk))
    if ws is None or ws.numel() < need:
        ws = _wl_workspace[dev] = torch.empty(need, device=dev, dtype=torch.uint8)
    a = _lib.WgradLayerArgs()
    a.gu, a.x, a.gy, a.h, a.rows = _lib.ptr(gu), _lib.ptr(x), _lib.ptr(gy), _lib.ptr(h), gu.shape[0]
    a.masks = _lib.ptr(masks)
    a.meta, a.nblk, a.channels = _lib.ptr(lay.meta), lay.nblk, 64
    for i, s in enumerate(shifts):
        a.shift[i] = int(s)
    a.drop_p, a.drop_seed, a.drop_stream = float(drop_p), int(seed) & 0xFFFFFFFF, int(stream_id) & 0xFFFFFFFF
    a.dw1, a.db1, a.dw2, a.db2 = _lib.ptr(gw1), _lib.ptr(gb1), _lib.ptr(gw2), _lib.ptr(gb2)
    a.workspace, a.workspace_bytes = _lib.ptr(ws), ws.numel()
    a.flags = 1 if skip_reduce else 0   # benchmarks time the main kernel alone
    _lib.check(lib.tcn_wgrad_layer(C.byref(a), _lib.stream_ptr()), "tcn_wgrad_layer")


def layer_fwd(x, w1f, w2f, b1, b2, lay: SeqLayout, shifts, save_h=True, drop_p=0.0, seed=0, stream_id=0):
    """Fused residual layer forward (64 channels): returns (y, h)."""
    lib = _lib.load()
    assert x.is_contiguous() and x.shape[1] == 64
    y = torch.zeros_like(x)
    h = torch.zeros_like(x) if save_h else None
    a = _lib.LayerFwdArgs()
    a.x, a.y, a.h = _lib.ptr(x), _lib.ptr(y), _lib.ptr(h)
    a.w1f, a.w2f, a.b1, a.b2 = _lib.ptr(w1f), _lib.ptr(w2f), _lib.ptr(b1), _lib.ptr(b2)
    a.meta, a.nblk, a.channels = _lib.ptr(lay.meta), lay.nblk, x.shape[1]
    for i, s in enumerate(shifts):
        a.shift[i] = int(s)
    a.drop_p, a.drop_seed, a.drop_stream = float(drop_p), int(seed) & 0xFFFFFFFF, int(stream_id) & 0xFFFFFFFF
    _lib.check(lib.tcn_layer_fwd(C.byref(a), _lib.stream_ptr()), "tcn_layer_fwd")
    return y, h


def layer_fwd_tc(x, w1, w2, b1, b2, lay: SeqLayout, shifts, save_h=True, drop_p=0.0, seed=0, stream_id=0,
                 save_masks=False):
    """Fused residual layer forward on tcgen05 / TMEM (64 channels): returns (y, h) or (y, h, masks).  w1 (64, 64, 3),
    w2 (64, 64, 1) are the torch weights (split into TF32 hi / lo halves here); masks: (rows, 4) int32 bit words
    (ReLU / dropout masks) for layer_bwd_tc."""
    lib = _lib.load()
    assert x.is_contiguous() and x.shape[1] == 64
    y = torch.zeros_like(x)
    h = torch.zeros_like(x) if save_h else None
    masks = torch.zeros(x.shape[0], 4, device=x.device, dtype=torch.int32) if save_masks else None
    w1h, w1l = split_weight(w1)
    w2h, w2l = split_weight(w2)
    a = _lib.LayerFwdTcArgs()
    a.x, a.x_rows, a.y, a.h = _lib.ptr(x), x.shape[0], _lib.ptr(y), _lib.ptr(h)
    a.w1_hi, a.w1_lo, a.w2_hi, a.w2_lo = _lib.ptr(w1h), _lib.ptr(w1l), _lib.ptr(w2h), _lib.ptr(w2l)
    a.b1, a.b2 = _lib.ptr(b1), _lib.ptr(b2)
    a.meta, a.nblk, a.channels = _lib.ptr(lay.meta), lay.nblk, x.shape[1]
    for i, s in enumerate(shifts):
        a.shift[i] = int(s)
    a.drop_p, a.drop_seed, a.drop_stream = float(drop_p), int(seed) & 0xFFFFFFFF, int(stream_id) & 0xFFFFFFFF
    a.masks = _lib.ptr(masks)
    _lib.check(lib.tcn_layer_fwd_tc(C.byref(a), _lib.stream_ptr()), "tcn_layer_fwd_tc")
    return (y, h, masks) if save_masks else (y, h)


def layer_bwd_tc(gy, masks, w1, w2, lay: SeqLayout, shifts, drop_p=0.0):
    """Fused input gradient of a residual layer (64 channels): returns (gu, gx); gu feeds the weight gradients."""
    lib = _lib.load()
    assert gy.is_contiguous() and gy.shape[1] == 64 and masks.is_contiguous()
    gu, gx = torch.zeros_like(gy), torch.zeros_like(gy)
    w2h, w2l = split_weight(w2, transpose=True)
    w1h, w1l = split_weight(w1, transpose=True)
    a = _lib.LayerBwdTcArgs()
    a.gy, a.g_rows, a.gu, a.gx, a.masks = _lib.ptr(gy), gy.shape[0], _lib.ptr(gu), _lib.ptr(gx), _lib.ptr(masks)
    a.w2t_hi, a.w2t_lo, a.w1t_hi, a.w1t_lo = _lib.ptr(w2h), _lib.ptr(w2l), _lib.ptr(w1h), _lib.ptr(w1l)
    a.meta, a.nblk, a.channels = _lib.ptr(lay.meta), lay.nblk, gy.shape[1]
    for i, s in enumerate(shifts):
        a.shift[i] = int(s)
    a.drop_p = float(drop_p)
    _lib.check(lib.tcn_layer_bwd_tc(C.byref(a), _lib.stream_ptr()), "tcn_layer_bwd_tc")
    return gu, gx


def split_weight(w, transpose: bool = False):
    """torch weight (n_out, c_in[, ntaps]) -> (w_hi, w_lo) in the layout tcn_gemm_tc reads: rows = output column
    (zero-padded to a multiple of 64), columns = tap * ceil(c_in/32)*32 + c; w_hi is exact in TF32, w_hi + w_lo == w.
    transpose=True builds the operand of the input-gradient pass."""
    lib = _lib.load()
    w = _f32c(w.detach())
    n_out, c_in = w.shape[0], w.shape[1]
    ntaps = w.shape[2] if w.dim() == 3 else 1
    n = lib.tcn_split_weight_floats(n_out, c_in, ntaps, int(transpose))
    kcols = ntaps * ((((n_out if transpose else c_in) + 31) // 32) * 32)
    hi = torch.empty(n // kcols, kcols, device=w.device, dtype=torch.float32)
    lo = torch.empty_like(hi)
    _lib.check(lib.tcn_split_weight(_lib.ptr(w), n_out, c_in, ntaps, int(transpose), _lib.ptr(hi), _lib.ptr(lo),
                                    _lib.stream_ptr()), "tcn_split_weight")
    return hi, lo


def gemm_tc(x, w_hi, w_lo, lay: SeqLayout, c_in, n_out, shifts=(0,), bias=None, out=None, ldy=None, residual=None,
            relu_mask=None, relu=False, x_unpadded=False, colscale=None, in_drop_p=0.0, in_drop_rescale=False,
            drop_p=0.0, seed=0, stream_id=0):
    """tcgen05/TMA tap GEMM over the packed rows (same contract as tapgemm)."""
    lib = _lib.load()
    assert x.is_contiguous() and x.dim() == 2
    if ldy is None:
        ldy = round_up(n_out, 4)
    if out is None:
        out = _rows_out(lay, ldy, x.device, written_cols=n_out)
    a = _lib.GemmTcArgs()
    a.x, a.ldx, a.x_rows, a.x_unpadded = _lib.ptr(x), x.shape[1], x.shape[0], int(x_unpadded)
    a.w_hi, a.w_lo, a.bias = _lib.ptr(w_hi), _lib.ptr(w_lo), _lib.ptr(bias)
    a.y, a.ldy = _lib.ptr(out), out.shape[1]
    a.residual, a.ldr = _lib.ptr(residual), (residual.shape[1] if residual is not None else 0)
    a.relu_mask, a.ldm = _lib.ptr(relu_mask), (relu_mask.shape[1] if relu_mask is not None else 0)
    a.meta, a.nblk = _lib.ptr(lay.meta), lay.nblk
    a.c_in, a.n_out, a.ntaps = c_in, n_out, len(shifts)
    for i, sft in enumerate(shifts):
        a.shift[i] = int(sft)
    a.relu = int(relu)
    a.colscale, a.colscale_ld = _lib.ptr(colscale), (colscale.shape[1] if colscale is not None else 0)
    a.in_drop_p, a.in_drop_rescale = float(in_drop_p), int(in_drop_rescale)
    a.drop_p, a.drop_seed, a.drop_stream = float(drop_p), int(seed) & 0xFFFFFFFF, int(stream_id) & 0xFFFFFFFF
    _lib.check(lib.tcn_gemm_tc(C.byref(a), _lib.stream_ptr()), "tcn_gemm_tc")
    return out


def dropout_apply(x, p, seed, stream_id):
    lib = _lib.load()
    y = torch.empty_like(x)
    _lib.check(lib.tcn_dropout_apply(_lib.ptr(x), x.shape[1], _lib.ptr(y), y.shape[1], x.shape[0], x.shape[1],
                                     float(p), int(seed) & 0xFFFFFFFF, int(stream_id) & 0xFFFFFFFF,
                                     _lib.stream_ptr()), "tcn_dropout_apply")
    return y


def dropout_keep_mask(nrows, ncols, p, seed, stream_id, device):
    """The keep-mask the kernels use for (seed, stream_id): uint8 (nrows, ncols).  For tests."""
    lib = _lib.load()
    keep = torch.empty(nrows, ncols, device=device, dtype=torch.uint8)
    _lib.check(lib.tcn_dropout_mask(_lib.ptr(keep), nrows, ncols, float(p), int(seed) & 0xFFFFFFFF,
                                    int(stream_id) & 0xFFFFFFFF, _lib.stream_ptr()), "tcn_dropout_mask")
    return keep


def sgd_step(params_flat, grads_flat, lr, weight_decay=0.0, grad_scale=1.0):
    lib = _lib.load()
    _lib.check(lib.tcn_sgd_step(_lib.ptr(params_flat), _lib.ptr(grads_flat), params_flat.numel(), float(lr),
                                float(weight_decay), float(grad_scale), _lib.stream_ptr()), "tcn_sgd_step")


def sgd_step_dev(params_flat, grads_flat, hyper):
    """SGD step with hyper = device tensor (lr, weight_decay, grad_scale): graph-capturable under an LR schedule."""
    lib = _lib.load()
    assert hyper.is_cuda and hyper.dtype == torch.float32 and hyper.numel() >= 3
    _lib.check(lib.tcn_sgd_step_dev(_lib.ptr(params_flat), _lib.ptr(grads_flat), params_flat.numel(), _lib.ptr(hyper),
                                    _lib.stream_ptr()), "tcn_sgd_step_dev")


def tap_shifts(dilation: int, causal: bool):
    d = int(dilation)
    return (-2 * d, -d, 0) if causal else (-d, 0, d)


def new_seed() -> int:
    return int(torch.randint(0, 2 ** 31 - 1, (1,)).item())


# ------------------------------------------------------------------------------------ autograd
def _tc_ok(x, c_in):
    return x.shape[1] % 4 == 0 and c_in % 4 == 0


class TapGemmFn(torch.autograd.Function):
    """y = sum_tap drop(x)[r + s_tap] W[:, :, tap]^T + bias (+ residual), packed time-major rows.
    Runs on the tcgen05 kernels when the row pitch allows TMA (multiple of 16 bytes), else on the mma.sync kernels."""

    @staticmethod
    def forward(ctx, x, weight, bias, residual, lay, shifts, x_unpadded, colscale, in_drop_p, seed, stream_id):
        x = _f32c(x)
        n_out, c_in = weight.shape[0], weight.shape[1]
        res = _f32c(residual) if residual is not None else None
        b = _f32c(bias.detach()) if bias is not None else None
        tc = _tc_ok(x, c_in)
        if tc:
            hi, lo = split_weight(weight)
            y = gemm_tc(x, hi, lo, lay, c_in, n_out, shifts, bias=b, residual=res, x_unpadded=x_unpadded,
                        colscale=colscale, in_drop_p=in_drop_p, in_drop_rescale=True, seed=seed, stream_id=stream_id)
        else:
            y = tapgemm(x, prep_weight(weight), lay, c_in, n_out, shifts, bias=b, residual=res, x_unpadded=x_unpadded,
                        colscale=colscale, in_drop_p=in_drop_p, seed=seed, stream_id=stream_id)
        ctx.save_for_backward(x, weight)
        ctx.lay, ctx.shifts, ctx.x_unpadded, ctx.colscale = lay, tuple(shifts), x_unpadded, colscale
        ctx.has_bias, ctx.has_res = bias is not None, residual is not None
        ctx.res_ld = res.shape[1] if res is not None else 0
        ctx.drop = (in_drop_p, seed, stream_id)
        ctx.tc = tc
        return y

    @staticmethod
    def backward(ctx, gy):
        x, weight = ctx.saved_tensors
        lay, shifts = ctx.lay, ctx.shifts
        gy = _f32c(gy)
        n_out, c_in = weight.shape[0], weight.shape[1]
        p, seed, sid = ctx.drop
        gx = gw = gb = gres = None
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            gw = torch.zeros_like(weight, dtype=torch.float32).contiguous()
            gb = torch.zeros(n_out, device=gy.device, dtype=torch.float32) if ctx.has_bias else None
            xin = dropout_apply(x, p, seed, sid) if p > 0 else x  # the weight gradient sees the dropped input
            if ctx.tc and gy.shape[1] % 4 == 0:
                wgrad_tc(gy, xin, lay, n_out, c_in, shifts, gw.view(n_out, c_in, -1), gb, x_unpadded=ctx.x_unpadded,
                         colscale=ctx.colscale)
            else:
                wgrad(gy, xin, lay, n_out, c_in, shifts, gw.view(n_out, c_in, -1), gb, x_unpadded=ctx.x_unpadded,
                      colscale=ctx.colscale)
        if ctx.needs_input_grad[0]:
            assert not ctx.x_unpadded, "input gradient for unpadded inputs is not needed on this path"
            kin = min(round_up(n_out, 4), gy.shape[1])
            nshifts = tuple(-s for s in shifts)
            if ctx.tc and gy.shape[1] % 4 == 0:
                hit, lot = split_weight(weight, transpose=True)
                gx = gemm_tc(gy, hit, lot, lay, kin, c_in, nshifts, ldy=x.shape[1], drop_p=p, seed=seed, stream_id=sid)
            else:
                gx = tapgemm(gy, prep_weight(weight, transpose=True), lay, kin, c_in, nshifts, ldy=x.shape[1], drop_p=p,
                             seed=seed, stream_id=sid)
        if ctx.has_res and ctx.needs_input_grad[3]:
            gres = gy if gy.shape[1] == ctx.res_ld else gy[:, :ctx.res_ld]
        return gx, gw, gb, gres, None, None, None, None, None, None, None


def tap_linear(x, weight, bias, lay, shifts=(0,), residual=None, x_unpadded=False, colscale=None, in_drop_p=0.0, seed=0,
               stream_id=0):
    return TapGemmFn.apply(x, weight, bias, residual, lay, tuple(shifts), x_unpadded, colscale, float(in_drop_p),
                           int(seed), int(stream_id))


class DilatedResidualFn(torch.autograd.Function):
    """y = x + Dropout_p(W2 relu(W1 (*)_d x + b1) + b2)   (network.py:178-198), packed rows x C.

    Forward keeps x and h = relu(u); backward regenerates the dropout mask from (seed, stream id).
    """

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, lay, dilation, causal, p, seed, stream_id):
        x = _f32c(x)
        Cc = w1.shape[0]
        shifts = tap_shifts(dilation, causal)
        masks = None
        if Cc == 64 and os.environ.get("TCN_NO_TCGEN05") is None:  # one fused launch, tcgen05 / TMEM
            y, h, masks = layer_fwd_tc(x, w1, w2, _f32c(b1.detach()), _f32c(b2.detach()), lay, shifts, True, p, seed,
                                       stream_id, save_masks=True)
        elif Cc == 64:  # one fused launch, mma.sync
            y, h = layer_fwd(x, prep_weight(w1), prep_weight(w2), _f32c(b1.detach()), _f32c(b2.detach()), lay, shifts,
                             True, p, seed, stream_id)
        else:
            h = tapgemm(x, prep_weight(w1), lay, Cc, Cc, shifts, bias=_f32c(b1.detach()), relu=True)
            y = tapgemm(h, prep_weight(w2), lay, Cc, Cc, (0,), bias=_f32c(b2.detach()), residual=x, drop_p=p,
                        seed=seed, stream_id=stream_id)
        ctx.save_for_backward(x, h, w1, w2)
        ctx.lay, ctx.shifts, ctx.p, ctx.seed, ctx.stream_id = lay, shifts, p, seed, stream_id
        ctx.masks = masks if os.environ.get("TCN_NO_FUSED_BWD") is None else None
        return y

    @staticmethod
    def backward(ctx, gy):
        x, h, w1, w2 = ctx.saved_tensors
        lay, shifts, p = ctx.lay, ctx.shifts, ctx.p
        gy = _f32c(gy)
        Cc = w1.shape[0]
        if ctx.masks is not None:  # 64 channels on tcgen05: fused input gradient + ONE deterministic weight-gradient pass
            gu, gx = layer_bwd_tc(gy, ctx.masks, w1, w2, lay, shifts, p)
            z = lambda *shape: torch.zeros(*shape, device=gy.device, dtype=torch.float32)
            gw1, gb1, gw2, gb2 = z(Cc, Cc, 3), z(Cc), z(Cc, Cc, 1), z(Cc)
            layer_wgrad(gu, x, gy, h, lay, shifts, gw1, gb1, gw2, gb2, drop_p=p, seed=ctx.seed, stream_id=ctx.stream_id,
                        masks=ctx.masks)
            return gx, gw1, gb1, gw2.view_as(w2), gb2, None, None, None, None, None, None
        # gv = keep * gy / (1 - p) is never materialised: the mask is regenerated as gy is loaded
        gw2 = torch.zeros(Cc, Cc, 1, device=gy.device, dtype=torch.float32)
        gb2 = torch.zeros(Cc, device=gy.device, dtype=torch.float32)
        wgrad(gy, h, lay, Cc, Cc, (0,), gw2, gb2, g_drop_p=p, seed=ctx.seed, stream_id=ctx.stream_id)
        gx = None
        gu = tapgemm(gy, prep_weight(w2, transpose=True), lay, Cc, Cc, (0,), relu_mask=h, in_drop_p=p,
                     seed=ctx.seed, stream_id=ctx.stream_id)
        gw1 = torch.zeros(Cc, Cc, 3, device=gy.device, dtype=torch.float32)
        gb1 = torch.zeros(Cc, device=gy.device, dtype=torch.float32)
        wgrad(gu, x, lay, Cc, Cc, shifts, gw1, gb1)
        if ctx.needs_input_grad[0] and gx is None:
            gx = tapgemm(gu, prep_weight(w1, transpose=True), lay, Cc, Cc, tuple(-s for s in shifts), residual=gy)
        return gx, gw1, gb1, gw2.view_as(w2), gb2, None, None, None, None, None, None


def dilated_residual(x, w1, b1, w2, b2, lay, dilation, causal=False, p=0.0, seed=0, stream_id=0):
    return DilatedResidualFn.apply(x, w1, b1, w2, b2, lay, int(dilation), bool(causal), float(p), int(seed),
                                   int(stream_id))
