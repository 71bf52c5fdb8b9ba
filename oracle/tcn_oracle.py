"""Functional CPU restatement of the dilated-TCN path and its losses.  TEST INFRASTRUCTURE ONLY.

Every function restates one reference symbol (cited as file:line, paths relative to
``/root/reference``) as explicit shifted matrix products on plain tensors, so that
it is an independent statement of the arithmetic (it never calls ``conv1d``) and
runs in float32 or float64 (pass float64 tensors to attribute rounding error).
Gradients come from autograd over these explicit formulas; the closed forms the
CUDA kernels implement are in ``layer_backward_closed_form`` and are checked
against autograd in ``tests/test_oracle_golden.py``.

Pinned against ``tests/golden/*.npz`` (outputs of the reference modules themselves,
written by ``oracle/gen_golden.py``).
"""
from __future__ import annotations

import torch

# --------------------------------------------------------------------------- helpers


def shift_time(x: torch.Tensor, s: int) -> torch.Tensor:
    """out[..., t] = x[..., t + s] with zeros outside [0, T).  x: (B, C, T)."""
    T = x.shape[-1]
    out = torch.zeros_like(x)
    if s == 0:
        return x.clone()
    if abs(s) >= T:
        return out
    if s > 0:
        out[..., : T - s] = x[..., s:]
    else:
        out[..., -s:] = x[..., : T + s]
    return out


def tap_offsets(dilation: int, causal: bool, padding: int | None = None):
    """Tap offsets s_k such that conv(x)[t] = sum_k W[:, :, k] x[t + s_k].

    Acausal: Conv1d(padding=d, dilation=d)  -> (-d, 0, +d)
             (MT4MTLKD/Temporal_tenco/network.py:189)
    Causal:  F.pad(x, [P, 0]) + Conv1d(padding=0, dilation=d), P = 2d by default
             -> (-P, -P + d, -P + 2d); the output then has T + P - 2d frames
             (network.py:166-179).  Only P == 2d keeps the residual add legal.
    """
    d = int(dilation)
    if not causal:
        return (-d, 0, d)
    P = 2 * d if padding is None else int(padding)
    return (-P, -P + d, -P + 2 * d)


def conv1x1(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor | None) -> torch.Tensor:
    """Conv1d(kernel_size=1): (B, Cin, T) -> (B, Cout, T).  w: (Cout, Cin, 1) or (Cout, Cin)."""
    w2 = w.reshape(w.shape[0], w.shape[1])
    y = torch.einsum("oc,bct->bot", w2, x)
    if b is not None:
        y = y + b.view(1, -1, 1)
    return y


def conv_taps(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor | None, offsets) -> torch.Tensor:
    """sum_k W[:, :, k] @ x[t + s_k] + b.  w: (Cout, Cin, len(offsets))."""
    y = None
    for k, s in enumerate(offsets):
        term = torch.einsum("oc,bct->bot", w[:, :, k], shift_time(x, s))
        y = term if y is None else y + term
    if b is not None:
        y = y + b.view(1, -1, 1)
    return y


# ------------------------------------------------------------------- a1 / a2 : layer


def dilated_residual_layer(x, w1, b1, w2, b2, dilation, causal=False, keep=None, p=0.5):
    """DilatedResidualLayer.forward (network.py:193-198) / DilatedResidualCausalLayer
    .forward (network.py:178-183).

    keep: None (eval, dropout is the identity) or a {0,1} tensor shaped like x (train:
    ``y = x + keep * v / (1 - p)``, nn.Dropout() default p = 0.5, network.py:191).
    """
    u = conv_taps(x, w1, b1, tap_offsets(dilation, causal))
    h = torch.relu(u)
    v = conv1x1(h, w2, b2)
    if keep is not None:
        v = v * keep / (1.0 - p)
    return x + v


def layer_backward_closed_form(x, gy, w1, b1, w2, b2, dilation, causal=False, keep=None, p=0.5):
    """Closed-form backward of the layer (SURVEY.md section 8a 'a1/a2 math').

    Returns (gx, gw1, gb1, gw2, gb2).  This is what the CUDA backward kernels compute.
    """
    offs = tap_offsets(dilation, causal)
    u = conv_taps(x, w1, b1, offs)
    h = torch.relu(u)
    gv = gy if keep is None else gy * keep / (1.0 - p)
    w2m = w2.reshape(w2.shape[0], w2.shape[1])
    gw2 = torch.einsum("bot,bct->oc", gv, h).reshape(w2.shape)
    gb2 = gv.sum(dim=(0, 2))
    gh = torch.einsum("oc,bot->bct", w2m, gv)
    gu = gh * (u > 0).to(gh.dtype)
    gw1 = torch.stack(
        [torch.einsum("bot,bct->oc", gu, shift_time(x, s)) for s in offs], dim=2
    )
    gb1 = gu.sum(dim=(0, 2))
    gx = gy.clone()
    for k, s in enumerate(offs):
        gx = gx + torch.einsum("oc,bot->bct", w1[:, :, k], shift_time(gu, -s))
    return gx, gw1, gb1, gw2, gb2


# -------------------------------------------------------------- a3 / a4 : the stages


def _layer_params(params, prefix, i):
    return (
        params[f"{prefix}.layers.{i}.conv_dilated.weight"],
        params[f"{prefix}.layers.{i}.conv_dilated.bias"],
        params[f"{prefix}.layers.{i}.conv_1x1.weight"],
        params[f"{prefix}.layers.{i}.conv_1x1.bias"],
    )


def _count_layers(params, prefix):
    n = 0
    while f"{prefix}.layers.{n}.conv_dilated.weight" in params:
        n += 1
    return n


def base_tcn(x_bdt, params, prefix="PG", mask=None, chan_keep=None, layer_keeps=None,
             causal=False, p=0.5):
    """BaseCausalTCN.forward (network.py:120-135).

    x_bdt: (B, D, T).  mask: optional {0,1} tensor like x (network.py:122-123).
    chan_keep: None (eval) or {0,1} (B, D): Dropout2d over whole input channels
    (network.py:125-127), scaled by 1/(1-p).  layer_keeps: None or list of per-layer
    dropout keep tensors (B, C, T).  Returns (features (B,C,T), conv_out(features)).
    """
    x = x_bdt
    if mask is not None:
        x = x * mask
    if chan_keep is not None:
        x = x * (chan_keep / (1.0 - p)).unsqueeze(-1)
    out = conv1x1(x, params[f"{prefix}.conv_1x1.weight"], params[f"{prefix}.conv_1x1.bias"])
    for i in range(_count_layers(params, prefix)):
        keep = None if layer_keeps is None else layer_keeps[i]
        out = dilated_residual_layer(out, *_layer_params(params, prefix, i), 2 ** i,
                                     causal=causal, keep=keep, p=p)
    logits = conv1x1(out, params[f"{prefix}.conv_out.weight"], params[f"{prefix}.conv_out.bias"])
    return out, logits


def refinement(x, params, prefix, use_output=False, hier=False, layer_keeps=None,
               causal=False, p=0.5):
    """Refinement.forward (network.py:149-162)."""
    out = x
    if use_output:
        out = conv1x1(x, params[f"{prefix}.conv_1x1.weight"], params[f"{prefix}.conv_1x1.bias"])
    for i in range(_count_layers(params, prefix)):
        keep = None if layer_keeps is None else layer_keeps[i]
        out = dilated_residual_layer(out, *_layer_params(params, prefix, i), 2 ** i,
                                     causal=causal, keep=keep, p=p)
    f = out
    if hier:  # AvgPool1d(kernel_size=7, stride=3), network.py:145,158-159
        f = out.unfold(2, 7, 3).mean(dim=-1)
    logits = conv1x1(f, params[f"{prefix}.conv_out.weight"], params[f"{prefix}.conv_out.bias"])
    return f, logits


# ----------------------------------------------------------------------- a5 : FPN


def linear_resize(x, size):
    """F.interpolate(x, size=size, mode='linear', align_corners=False) on (B, C, W)
    (network.py:95-96).  Exact identity when W == size."""
    W = x.shape[-1]
    if W == size:
        return x
    scale = W / size
    dst = torch.arange(size, dtype=x.dtype)
    src = torch.clamp((dst + 0.5) * scale - 0.5, min=0.0)
    i0 = torch.clamp(src.floor().long(), max=W - 1)
    i1 = torch.clamp(i0 + 1, max=W - 1)
    lam = (src - i0.to(x.dtype)).view(1, 1, -1)
    return x[..., i0] * (1 - lam) + x[..., i1] * lam


def fpn(f_list, params, prefix="fpn"):
    """FPN.forward (network.py:98-106): only latlayer1 is ever applied (3 times)."""
    w, b = params[f"{prefix}.latlayer1.weight"], params[f"{prefix}.latlayer1.bias"]
    c1, c2, c3, p4 = f_list
    p3 = linear_resize(p4, c3.shape[-1]) + conv1x1(c3, w, b)
    p2 = linear_resize(p3, c2.shape[-1]) + conv1x1(c2, w, b)
    p1 = linear_resize(p2, c1.shape[-1]) + conv1x1(c1, w, b)
    return [p1, p2, p3, p4]


# -------------------------------------------------------------------- a6 : VideoNas


def videonas_forward(x_btd, params, fpn_on=True, use_output=False, hier=False, mask=None,
                     chan_keep=None, layer_keeps=None, causal=False, p=0.5):
    """VideoNas.forward (network.py:36-68).

    x_btd: (B, T, D).  layer_keeps: None or dict stage-prefix -> list of keep tensors.
    Returns (out_list, out_list_i, out_list_v, out_list_t, f_list, f_list).
    """
    x = x_btd.permute(0, 2, 1)
    lk = (lambda pre: None) if layer_keeps is None else (lambda pre: layer_keeps.get(pre))
    f, out1 = base_tcn(x, params, "PG", mask=mask, chan_keep=chan_keep, layer_keeps=lk("PG"),
                       causal=causal, p=p)
    f_list = [f]
    out_list, out_i, out_v, out_t = [], [], [], []
    if not fpn_on:
        out_list.append(out1)
    s = 0
    while f"Rs.{s}.conv_out.weight" in params:
        f, out1 = refinement(f, params, f"Rs.{s}", use_output=use_output, hier=hier,
                             layer_keeps=lk(f"Rs.{s}"), causal=causal, p=p)
        f_list.append(f)
        s += 1
    if fpn_on:
        f_list = fpn(f_list, params)
        for f in f_list:
            out_list.append(conv1x1(f, params["conv_out.weight"], params["conv_out.bias"]))
            out_i.append(conv1x1(f, params["conv_out_i.weight"], params["conv_out_i.bias"]))
            out_v.append(conv1x1(f, params["conv_out_v.weight"], params["conv_out_v.bias"]))
            out_t.append(conv1x1(f, params["conv_out_t.weight"], params["conv_out_t.bias"]))
    return out_list, out_i, out_v, out_t, f_list, f_list


# ------------------------------------------------------------------ a7 / a8 : losses

# Spatial_cnn/run.py:306-310 == TERL/0_5fold_TCN_black/run.py:432-436
TOOL_WEIGHT = [0.93487068, 0.94234964, 0.93487068, 1.18448115, 1.02368339, 0.97974447]
VERB_WEIGHT = [0.60002400, 0.60002400, 0.60002400, 0.61682467, 0.67082683, 0.80163207,
               0.70562823, 2.11208448, 2.69230769, 0.60062402]
TARGET_WEIGHT = [0.49752894, 0.52041527, 0.49752894, 0.51394739, 2.71899565, 1.75577963,
                 0.58509403, 1.25228034, 0.49752894, 2.42993134, 0.49802647, 0.87266576,
                 1.36074165, 0.50150917, 0.49802647]


def log_sigmoid(x):
    return torch.minimum(x, torch.zeros_like(x)) - torch.log1p(torch.exp(-x.abs()))


def bce_with_logits(x, y, pos_weight=None):
    """nn.BCEWithLogitsLoss(pos_weight)(x, y), mean reduction.  x, y: (N, K)."""
    pw = 1.0 if pos_weight is None else pos_weight.view(1, -1)
    loss = -(pw * y * log_sigmoid(x) + (1.0 - y) * log_sigmoid(-x))
    return loss.mean()


def tenco_loss(outs, labels, pos_weights=None, loss_type="all"):
    """Loss composition of train_loop.

    MT4MTLKD/Temporal_tenco/run.py:190-212 (unweighted BCE on all four heads) and
    TERL/0_5fold_TCN_black/run.py:307-343 (pos_weight on i/v/t, --loss_type switch).
    outs = (out_list, out_list_i, out_list_v, out_list_t) with logits (B, K, T);
    labels = (y_i, y_v, y_t, y_ivt), each (T, K) -- the reference takes sample 0 only
    (``pd[0]``, ``y4[0]``).  ``fusion`` (run.py:159-179) is the identity when the
    lengths agree, which they always do (hier=False).
    Returns (loss, loss_i, loss_v, loss_t, loss_ivt).
    """
    out_ivt, out_i, out_v, out_t = outs
    y_i, y_v, y_t, y_ivt = labels
    pw = pos_weights or (None, None, None)

    def head(lst, y, w):
        tot = 0.0
        for pd in lst:
            tot = tot + bce_with_logits(pd[0].transpose(0, 1), y.to(pd.dtype), w)
        return tot

    li, lv, lt = head(out_i, y_i, pw[0]), head(out_v, y_v, pw[1]), head(out_t, y_t, pw[2])
    livt = head(out_ivt, y_ivt, None)
    if loss_type == "i":
        loss = li
    elif loss_type == "v":
        loss = lv
    elif loss_type == "t":
        loss = lt
    elif loss_type == "ivt":
        loss = livt
    elif loss_type == "single":
        loss = (li + lv + lt) / 3
    else:
        loss = 0.1 * (li + lv + lt) + livt
    return loss, li, lv, lt, livt


def distill_kl(y_s, y_t, T):
    """DistillKL.forward (MT4MTLKD/Spatial_cnn/run.py:291-295).

    KL(softmax(y_t/T) || softmax(y_s/T)) summed over everything, times T^2 / N.
    Note the caller passes y_t = sigmoid(teacher_logits) (run.py:180-182).
    """
    ls = torch.log_softmax(y_s / T, dim=1)
    lt = torch.log_softmax(y_t / T, dim=1)
    pt = lt.exp()
    return (pt * (lt - ls)).sum() * (T ** 2) / y_s.shape[0]


def mse(a, b):
    return ((a - b) ** 2).mean()


def multi_teacher_kd_loss(logits, labels, teacher_logits, feats=None, teacher_feats=None,
                          T=4.0, rates=(1.0, 1.0, 1.0), pos_weights=None):
    """MT4MTLKD/Spatial_cnn/run.py:159-192.

    logits = (i, v, t, ivt) student (N, K); labels likewise; teacher_logits = (i, v, t)
    raw teacher logits.  hard = sum of four BCE (pos_weight on i/v/t, run.py:323-326);
    soft = mean of three DistillKL against sigmoid(teacher); kd = mean of three MSE.
    Returns (loss, hard, soft, kd).
    """
    pw = pos_weights or (None, None, None)
    li = bce_with_logits(logits[0], labels[0], pw[0])
    lv = bce_with_logits(logits[1], labels[1], pw[1])
    lt = bce_with_logits(logits[2], labels[2], pw[2])
    livt = bce_with_logits(logits[3], labels[3], None)
    hard = li + lv + lt + livt
    soft = sum(distill_kl(logits[k], torch.sigmoid(teacher_logits[k]), T) for k in range(3)) / 3
    kd = torch.zeros((), dtype=hard.dtype)
    if feats is not None:
        kd = sum(mse(feats[k], teacher_feats[k]) for k in range(3)) / 3
    loss = rates[0] * hard + rates[1] * soft + rates[2] * kd
    return loss, hard, soft, kd


def feature_kd_attention(s, teachers, params):
    """Multi-teacher attention re-weighting of the student feature, MT4MTLKD/Spatial_cnn/network.py:47-71
    (same block: Spatial_transformer/network.py:102-124).

    s (B, F) student feature; teachers = (t_i, t_v, t_t), each (B, M); params: '{wi,wv,wt}.weight' (M, F, 1),
    '{mi,mv,mt}.weight' (F, M, 1) and biases.  The reference stacks F copies of s into stus[b, c, d] = s[b, c] and the
    three projected teachers into teas[b, d, n] = m_n(t_n)[b, d] (:56-59), so that
    attn[b, c, :] = softmax_n( sum_d stus[b, c, d] / sqrt(F) * teas[b, d, n] ) (:61-62) and
    stus_f_n = w_n(s * attn[:, :, n]) (:63-65).  Returns (stus_fi, stus_fv, stus_ft), each (B, M)."""
    F = s.shape[1]
    teas = torch.stack([conv1x1(t.unsqueeze(-1), params[f"{m}.weight"], params[f"{m}.bias"]).squeeze(-1)
                        for t, m in zip(teachers, ("mi", "mv", "mt"))], dim=-1)          # (B, F, 3)
    stus = s.unsqueeze(-1).expand(-1, -1, F)                                               # (B, F, F)
    attn = torch.einsum("bcd,bdn->bcn", stus / (F ** 0.5), teas).softmax(dim=-1)
    return tuple(conv1x1((s * attn[:, :, n]).unsqueeze(-1), params[f"{w}.weight"], params[f"{w}.bias"]).squeeze(-1)
                 for n, w in enumerate(("wi", "wv", "wt")))


def feature_kd_loss(stus_f, teachers):
    """Spatial_cnn/run.py:187-191: mean of the three MSE losses."""
    return sum(mse(a, b) for a, b in zip(stus_f, teachers)) / 3


def phase_ce(logits, target):
    """7-way 'phase' head: mean softmax cross-entropy over frames.  logits (N, K), target (N,) int.

    PARITY UNPINNED: the reference has no phase head (SURVEY.md 'Facts'); this is the
    textbook definition, equal to torch.nn.functional.cross_entropy.
    """
    ls = torch.log_softmax(logits, dim=1)
    return -ls.gather(1, target.view(-1, 1).long()).mean()
