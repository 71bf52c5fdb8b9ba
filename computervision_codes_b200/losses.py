"""Loss kernels' Python face: sigmoid-BCE over concatenated heads, DistillKL, MSE, softmax-CE.

Mirrors the loss arithmetic the reference inlines in its run scripts:
  MT4MTLKD/Temporal_tenco/run.py:159-212, TERL/0_5fold_TCN_black/run.py:273-343  (temporal student)
  MT4MTLKD/Spatial_cnn/run.py:159-192,284-295,306-328                            (multi-teacher KD)
"""
from __future__ import annotations

import ctypes as C

import torch
from torch import nn

from . import _lib
from .layout import SeqLayout, round_up

# Spatial_cnn/run.py:306-310 == TERL/0_5fold_TCN_black/run.py:432-436
TOOL_WEIGHT = [0.93487068, 0.94234964, 0.93487068, 1.18448115, 1.02368339, 0.97974447]
VERB_WEIGHT = [0.60002400, 0.60002400, 0.60002400, 0.61682467, 0.67082683, 0.80163207, 0.70562823, 2.11208448,
               2.69230769, 0.60062402]
TARGET_WEIGHT = [0.49752894, 0.52041527, 0.49752894, 0.51394739, 2.71899565, 1.75577963, 0.58509403, 1.25228034,
                 0.49752894, 2.42993134, 0.49802647, 0.87266576, 1.36074165, 0.50150917, 0.49802647]

_col_cache: dict = {}


def _columns(head_sizes, head_weights, pos_weights, device):
    """Per-column tables for heads concatenated as columns.  head index = position in head_sizes."""
    key = (tuple(head_sizes), tuple(float(w) for w in head_weights),
           None if pos_weights is None else tuple(None if p is None else tuple(float(v) for v in p)
                                                  for p in pos_weights), str(device))
    hit = _col_cache.get(key)
    if hit is not None:
        return hit
    unit, scale, head, pw = [], [], [], []
    for h, (k, w) in enumerate(zip(head_sizes, head_weights)):
        unit += [1.0 / k] * k
        scale += [w / k] * k
        head += [h] * k
        p = None if pos_weights is None else pos_weights[h]
        pw += [1.0] * k if p is None else [float(v) for v in p]
    t = lambda v, dt: torch.tensor(v, dtype=dt, device=device)
    out = (t(unit, torch.float32), t(scale, torch.float32), t(head, torch.int32),
           None if pos_weights is None else t(pw, torch.float32))
    _col_cache[key] = out
    return out


def _bce_launch(logits, labels_u8, ncols, cols, row_scale, loss8, dl, meta=None, nrows=None, lab_unpadded=False,
                grad_scale=1.0):
    lib = _lib.load()
    unit, scale, head, pw = cols
    a = _lib.BceArgs()
    a.logits, a.ldl = _lib.ptr(logits), logits.shape[1]
    a.labels, a.ldlab, a.lab_unpadded = _lib.ptr(labels_u8), labels_u8.shape[1], int(lab_unpadded)
    a.meta = _lib.ptr(meta)
    a.nrows = logits.shape[0] if nrows is None else nrows
    a.ncols = ncols
    a.zero_cols = dl.shape[1] if dl is not None else ncols
    a.pos_w, a.col_scale, a.col_unit, a.col_head = _lib.ptr(pw), _lib.ptr(scale), _lib.ptr(unit), _lib.ptr(head)
    a.row_scale = float(row_scale)
    a.loss = _lib.ptr(loss8)
    a.dl, a.lddl, a.grad_scale = _lib.ptr(dl), (dl.shape[1] if dl is not None else 0), float(grad_scale)
    _lib.check(lib.tcn_bce_rows(C.byref(a), _lib.stream_ptr()), "tcn_bce_rows")


class _MultiHeadBceFn(torch.autograd.Function):
    """Sum over levels of the per-video mean BCE of every head; forward also writes dLogits."""

    @staticmethod
    def forward(ctx, lay, labels_u8, head_sizes, head_weights, pos_weights, lab_unpadded, *logit_rows):
        dev = logit_rows[0].device
        ncols = sum(head_sizes)
        cols = _columns(head_sizes, head_weights, pos_weights, dev)
        loss8 = torch.zeros(8, device=dev, dtype=torch.float32)
        need_grad = any(ctx.needs_input_grad[6:])
        dls = []
        for lg in logit_rows:
            lg = lg if lg.is_contiguous() else lg.contiguous()
            dl = torch.zeros_like(lg) if need_grad else None   # the kernel skips pad rows: they must read as zero
            _bce_launch(lg, labels_u8, ncols, cols, 1.0 / lay.num_seqs, loss8, dl, meta=lay.meta, nrows=lay.rows,
                        lab_unpadded=lab_unpadded)
            dls.append(dl)
        if need_grad:
            ctx.save_for_backward(*dls)
        nh = len(head_sizes)
        w = torch.tensor(list(head_weights), device=dev, dtype=torch.float32)
        total = (loss8[:nh] * w).sum()
        ctx.mark_non_differentiable(loss8)
        return total, loss8

    @staticmethod
    def backward(ctx, gtotal, _g8):
        dls = ctx.saved_tensors
        return (None,) * 6 + tuple(dl * gtotal for dl in dls)


def multi_head_bce(logit_rows, labels_u8, lay: SeqLayout, head_sizes, head_weights, pos_weights=None,
                   lab_unpadded=True):
    """logit_rows: list (levels) of packed (rows, ld) logits with heads concatenated along columns.
    labels_u8: uint8 (frames, >= sum(head_sizes)) in the same column order.
    Returns (total, per_head[8]) -- per_head[h] = sum over levels of head h's mean BCE."""
    return _MultiHeadBceFn.apply(lay, labels_u8, tuple(head_sizes), tuple(head_weights), pos_weights, lab_unpadded,
                                 *logit_rows)


_LOSS_TYPE_WEIGHTS = {  # column order ivt | i | v | t
    "all": (1.0, 0.1, 0.1, 0.1), "i": (0.0, 1.0, 0.0, 0.0), "v": (0.0, 0.0, 1.0, 0.0), "t": (0.0, 0.0, 0.0, 1.0),
    "ivt": (1.0, 0.0, 0.0, 0.0), "single": (0.0, 1 / 3, 1 / 3, 1 / 3),
}


def pack_labels(y_i, y_v, y_t, y_ivt):
    """Four (T, K) integer label matrices -> uint8 (T, 132) in column order ivt | i | v | t (+1 pad)."""
    lab = torch.cat([y_ivt, y_i, y_v, y_t], dim=1).to(torch.uint8)
    pad = round_up(lab.shape[1], 4) - lab.shape[1]
    if pad:
        lab = torch.nn.functional.pad(lab, (0, pad))
    return lab.contiguous()


def tenco_loss(logit_rows, labels_u8, lay, head_sizes=(100, 6, 10, 15), loss_type="all", terl_pos_weight=False):
    """Temporal_tenco/run.py:190-212 (unweighted) / TERL run.py:307-343 (pos_weight on i/v/t).
    Returns (loss, loss_i, loss_v, loss_t, loss_ivt) as 0-d tensors (one device buffer, no syncs)."""
    pws = (None, TOOL_WEIGHT, VERB_WEIGHT, TARGET_WEIGHT) if terl_pos_weight else None
    total, per = multi_head_bce(logit_rows, labels_u8, lay, head_sizes, _LOSS_TYPE_WEIGHTS[loss_type], pws)
    return total, per[1], per[2], per[3], per[0]


# ---------------------------------------------------------------------------------------------- (N, K) losses
class _BceFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, pos_weight):
        lg = logits.contiguous().float()
        N, K = lg.shape
        cols = _columns((K,), (1.0,), None if pos_weight is None else (tuple(float(v) for v in pos_weight),),
                        lg.device)
        loss8 = torch.zeros(8, device=lg.device, dtype=torch.float32)
        dl = torch.empty_like(lg) if ctx.needs_input_grad[0] else None
        lab = labels.to(torch.uint8).contiguous()
        _bce_launch(lg, lab, K, cols, 1.0 / N, loss8, dl)
        if dl is not None:
            ctx.save_for_backward(dl)
        return loss8[0]

    @staticmethod
    def backward(ctx, g):
        (dl,) = ctx.saved_tensors
        return dl * g, None, None


def bce_with_logits(logits, labels, pos_weight=None):
    """nn.BCEWithLogitsLoss(pos_weight)(logits (N, K), labels (N, K) in {0,1})."""
    return _BceFn.apply(logits, labels, pos_weight)


class _KdKlFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y_s, y_t, T, teacher_sigmoid):
        lib = _lib.load()
        ys, yt = y_s.contiguous().float(), y_t.detach().contiguous().float()
        N, K = ys.shape
        loss = torch.zeros(1, device=ys.device, dtype=torch.float32)
        g = torch.empty_like(ys) if ctx.needs_input_grad[0] else None
        _lib.check(lib.tcn_kd_kl_rows(_lib.ptr(ys), K, _lib.ptr(yt), K, int(teacher_sigmoid), N, K, float(T),
                                      _lib.ptr(loss), 1.0, _lib.ptr(g), K, 1.0, _lib.stream_ptr()), "tcn_kd_kl_rows")
        if g is not None:
            ctx.save_for_backward(g)
        return loss[0]

    @staticmethod
    def backward(ctx, gout):
        (g,) = ctx.saved_tensors
        return g * gout, None, None, None


class DistillKL(nn.Module):
    """MT4MTLKD/Spatial_cnn/run.py:284-295 -- same constructor and forward(y_s, y_t)."""

    def __init__(self, T):
        super().__init__()
        self.T = T

    def forward(self, y_s, y_t, teacher_is_logits=False):
        """y_t: teacher *probabilities* sigmoid(teacher logits) as the reference passes them
        (run.py:180-182); with teacher_is_logits=True the sigmoid is fused into the kernel."""
        return _KdKlFn.apply(y_s, y_t, self.T, teacher_is_logits)


class _MseFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        lib = _lib.load()
        a_, b_ = a.contiguous().float(), b.detach().contiguous().float()
        loss = torch.zeros(1, device=a_.device, dtype=torch.float32)
        g = torch.empty_like(a_) if ctx.needs_input_grad[0] else None
        _lib.check(lib.tcn_mse(_lib.ptr(a_), _lib.ptr(b_), a_.numel(), _lib.ptr(loss), 1.0, _lib.ptr(g), 1.0,
                               _lib.stream_ptr()), "tcn_mse")
        if g is not None:
            ctx.save_for_backward(g)
        return loss[0]

    @staticmethod
    def backward(ctx, gout):
        (g,) = ctx.saved_tensors
        return g * gout, None


def mse_loss(a, b):
    """nn.MSELoss()(a, b) (Spatial_cnn/run.py:328)."""
    return _MseFn.apply(a, b)


class _CeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target):
        lib = _lib.load()
        x = logits.contiguous().float()
        N, K = x.shape
        tg = target.to(torch.int32).contiguous()
        loss = torch.zeros(1, device=x.device, dtype=torch.float32)
        g = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        _lib.check(lib.tcn_ce_rows(_lib.ptr(x), K, _lib.ptr(tg), None, 0, N, K, 1.0 / N, _lib.ptr(loss), _lib.ptr(g),
                                   K, 1.0, _lib.stream_ptr()), "tcn_ce_rows")
        if g is not None:
            ctx.save_for_backward(g)
        return loss[0]

    @staticmethod
    def backward(ctx, gout):
        (g,) = ctx.saved_tensors
        return g * gout, None


def phase_cross_entropy(logits, target):
    """Mean softmax-CE for the 7-way phase head (logits (N, K), target (N,))."""
    return _CeFn.apply(logits, target)


class MultiTeacherKDLoss(nn.Module):
    """Loss composition of MT4MTLKD/Spatial_cnn/run.py:159-192:
    rates[0] * (BCE_i + BCE_v + BCE_t + BCE_ivt) + rates[1] * mean_k DistillKL(logit_k, sigmoid(teacher_k))
    + rates[2] * mean_k MSE(student_feat_k, teacher_feat_k), pos_weight BCE on i/v/t (run.py:323-326)."""

    def __init__(self, temp=4.0, rates=(1.0, 1.0, 1.0), pos_weight=True):
        super().__init__()
        self.kl = DistillKL(temp)
        self.rates = tuple(rates)
        self.pw = (TOOL_WEIGHT, VERB_WEIGHT, TARGET_WEIGHT) if pos_weight else (None, None, None)

    def forward(self, logits, labels, teacher_logits, feats=None, teacher_feats=None):
        """logits/labels: (i, v, t, ivt) tuples of (N, K); teacher_logits: (i, v, t) raw logits.
        Returns (loss, hard, soft, kd)."""
        hard = (bce_with_logits(logits[0], labels[0], self.pw[0]) + bce_with_logits(logits[1], labels[1], self.pw[1])
                + bce_with_logits(logits[2], labels[2], self.pw[2]) + bce_with_logits(logits[3], labels[3], None))
        soft = sum(self.kl(logits[k], teacher_logits[k], teacher_is_logits=True) for k in range(3)) / 3
        kd = hard.new_zeros(())
        if feats is not None:
            kd = sum(mse_loss(feats[k], teacher_feats[k]) for k in range(3)) / 3
        loss = self.rates[0] * hard + self.rates[1] * soft + self.rates[2] * kd
        return loss, hard, soft, kd


# ---------------------------------------------------------------------------------------------- feature KD (row f3)
class _KdAttnFn(torch.autograd.Function):
    """(s, m_i(t_i), m_v(t_v), m_t(t_t)) -> (s * attn_i, s * attn_v, s * attn_t); all (N, F)."""

    @staticmethod
    def forward(ctx, s, tea_i, tea_v, tea_t):
        lib = _lib.load()
        s = s.contiguous().float()
        teas = [t.contiguous().float() for t in (tea_i, tea_v, tea_t)]
        N, Fd = s.shape
        assert all(t.shape == s.shape for t in teas), "projected teacher features must be (N, student_dim)"
        zs = [torch.empty_like(s) for _ in range(3)]
        tsum = torch.empty(N, 3, device=s.device, dtype=torch.float32)
        a = _lib.KdAttnArgs()
        a.s, a.lds, a.ldt, a.ldz, a.tsum = _lib.ptr(s), Fd, Fd, Fd, _lib.ptr(tsum)
        for n in range(3):
            a.tea[n], a.z[n] = _lib.ptr(teas[n]), _lib.ptr(zs[n])
        a.n_rows, a.feat_dim = N, Fd
        _lib.check(lib.tcn_kd_attn_fwd(C.byref(a), _lib.stream_ptr()), "tcn_kd_attn_fwd")
        ctx.save_for_backward(s, tsum, *teas)
        return tuple(zs)

    @staticmethod
    def backward(ctx, g_i, g_v, g_t):
        lib = _lib.load()
        s, tsum, *teas = ctx.saved_tensors
        N, Fd = s.shape
        gz = [g.contiguous().float() for g in (g_i, g_v, g_t)]
        gs = torch.empty_like(s)
        gteas = [torch.empty_like(s) for _ in range(3)]
        a = _lib.KdAttnArgs()
        a.s, a.lds, a.ldt, a.ldz, a.tsum = _lib.ptr(s), Fd, Fd, Fd, _lib.ptr(tsum)
        a.gs, a.ldgs = _lib.ptr(gs), Fd
        for n in range(3):
            a.tea[n], a.z[n] = _lib.ptr(teas[n]), _lib.ptr(gz[n])   # z is not written by the backward
            a.gz[n], a.gtea[n] = _lib.ptr(gz[n]), _lib.ptr(gteas[n])
        a.n_rows, a.feat_dim = N, Fd
        _lib.check(lib.tcn_kd_attn_bwd(C.byref(a), _lib.stream_ptr()), "tcn_kd_attn_bwd")
        return gs, gteas[0], gteas[1], gteas[2]


class MultiTeacherFeatureAttention(nn.Module):
    """The feature-KD head of the spatial student, ``MT4MTLKD/Spatial_cnn/network.py:26-32`` (parameters) and
    ``:47-71`` (forward); identical block in ``Spatial_transformer/network.py:102-124``.

    Parameters keep the reference's names, shapes and creation order -- ``wi, wv, wt: Conv1d(student_dim,
    teacher_dim, 1)``, ``mi, mv, mt: Conv1d(teacher_dim, student_dim, 1)`` -- so a reference ``state_dict`` loads with
    ``strict=False`` and equal seeds give equal initial weights.  ``forward(s, tool, verb, target)``: s (N, student_dim)
    pooled student feature, the three teacher features (N, teacher_dim); returns (stus_fi, stus_fv, stus_ft), each
    (N, teacher_dim), which ``MultiTeacherKDLoss(feats=..., teacher_feats=...)`` compares with the teachers by MSE
    (``Spatial_cnn/run.py:187-191``).  The 1x1 projections run on the tap-GEMM kernels, the re-weighting on
    ``tcn_kd_attn_{fwd,bwd}``."""

    def __init__(self, student_dim=512, teacher_dim=1536):
        super().__init__()
        self.feat_dim, self.num_f_mstct = student_dim, teacher_dim
        self.wi = nn.Conv1d(student_dim, teacher_dim, kernel_size=1, stride=1, padding=0)
        self.wv = nn.Conv1d(student_dim, teacher_dim, kernel_size=1, stride=1, padding=0)
        self.wt = nn.Conv1d(student_dim, teacher_dim, kernel_size=1, stride=1, padding=0)
        self.mi = nn.Conv1d(teacher_dim, student_dim, kernel_size=1, stride=1, padding=0)
        self.mv = nn.Conv1d(teacher_dim, student_dim, kernel_size=1, stride=1, padding=0)
        self.mt = nn.Conv1d(teacher_dim, student_dim, kernel_size=1, stride=1, padding=0)

    def forward(self, s, tool, verb, target):
        from . import ops

        if not s.is_cuda:
            raise RuntimeError("MultiTeacherFeatureAttention runs on the CUDA kernels only (no CPU path)")
        N = s.shape[0]
        lay = SeqLayout.get([N], s.device)
        pad = lay.rows - N

        def rows(t):
            return torch.nn.functional.pad(t.float(), (0, 0, 0, pad)) if pad else t.float().contiguous()

        teas = [ops.tap_linear(rows(t), m.weight, m.bias, lay)[:N, :self.feat_dim]
                for t, m in zip((tool, verb, target), (self.mi, self.mv, self.mt))]
        zs = _KdAttnFn.apply(s, *teas)
        return tuple(ops.tap_linear(rows(z), w.weight, w.bias, lay)[:N, :self.num_f_mstct]
                     for z, w in zip(zs, (self.wi, self.wv, self.wt)))
