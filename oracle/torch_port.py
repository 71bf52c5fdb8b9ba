"""CPU port of the temporal student train step on torch's own CPU convolutions (oneDNN) -- the CPU
arm of bench.py (`cpu_baseline`, `--impl reference`).  TEST / BENCH INFRASTRUCTURE ONLY.

Same arithmetic as the reference modules (MT4MTLKD/Temporal_tenco/network.py:14-198) and loss
(Temporal_tenco/run.py:190-212) executed the way the reference executes it on a CPU: F.conv1d per
layer, F.dropout in train mode, BCE-with-logits per head and level, autograd backward.  Written
functionally over a parameter dict with the reference's state_dict keys; checked against
tests/golden/tcn_videonas.npz in tests/test_oracle_golden.py.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _layer(x, p, pre, d, causal, train):
    if causal:
        u = F.conv1d(F.pad(x, [2 * d, 0]), p[pre + ".conv_dilated.weight"], p[pre + ".conv_dilated.bias"], dilation=d)
    else:
        u = F.conv1d(x, p[pre + ".conv_dilated.weight"], p[pre + ".conv_dilated.bias"], padding=d, dilation=d)
    v = F.conv1d(F.relu(u), p[pre + ".conv_1x1.weight"], p[pre + ".conv_1x1.bias"])
    return x + F.dropout(v, 0.5, train)


def _stage(x, p, pre, causal, train):
    i = 0
    while f"{pre}.layers.{i}.conv_dilated.weight" in p:
        x = _layer(x, p, f"{pre}.layers.{i}", 2 ** i, causal, train)
        i += 1
    return x


def videonas_forward(x_btd, p, causal=False, train=False, mask=None):
    x = x_btd.permute(0, 2, 1)
    if mask is not None:
        x = x * mask
    if train:
        x = F.dropout2d(x.unsqueeze(3), 0.5, True).squeeze(3)
    f = F.conv1d(x, p["PG.conv_1x1.weight"], p["PG.conv_1x1.bias"])
    f = _stage(f, p, "PG", causal, train)
    fs = [f]
    s = 0
    while f"Rs.{s}.conv_out.weight" in p:
        f = _stage(f, p, f"Rs.{s}", causal, train)
        fs.append(f)
        s += 1
    lw, lb = p["fpn.latlayer1.weight"], p["fpn.latlayer1.bias"]
    p4 = fs[3]
    p3 = p4 + F.conv1d(fs[2], lw, lb)
    p2 = p3 + F.conv1d(fs[1], lw, lb)
    p1 = p2 + F.conv1d(fs[0], lw, lb)
    ps = [p1, p2, p3, p4]
    heads = {n: [F.conv1d(q, p[f"conv_out{n}.weight"], p[f"conv_out{n}.bias"]) for q in ps] for n in ("", "_i", "_v", "_t")}
    return heads[""], heads["_i"], heads["_v"], heads["_t"], ps


def train_step_loss(x_btd, p, labels, causal=False, train=True, mask=None):
    """labels = (y_i, y_v, y_t, y_ivt) float (T, K).  Returns the tenco total loss (run.py:212)."""
    o, oi, ov, ot, _ = videonas_forward(x_btd, p, causal, train, mask)

    def head(lst, y):
        return sum(F.binary_cross_entropy_with_logits(q[0].transpose(0, 1), y) for q in lst)

    return 0.1 * (head(oi, labels[0]) + head(ov, labels[1]) + head(ot, labels[2])) + head(o, labels[3])
