"""Temporal-student train step (forward + loss + backward [+ all-reduce] + SGD) over a ragged batch of
videos -- the hot loop of MT4MTLKD/Temporal_tenco/run.py:182-235 (TERL/0_5fold_TCN_black/run.py:297-366).

Parameters that receive a gradient live in one flat fp32 buffer with a matching flat gradient buffer,
so data-parallel training needs exactly one all-reduce per step (torch.distributed / NCCL over
NVLink) and one fused SGD launch.  Videos shard across ranks; nothing else crosses GPUs.
"""
from __future__ import annotations

import torch

from . import losses, ops
from .layout import SeqLayout

# Parameters the reference leaves without a gradient on the --fpn path: they stay
# outside the flat trainable buffer, their .grad stays None, weight decay never touches them.
_NO_GRAD_FPN = ("PG.conv_out.", ".conv_1x1.", ".conv_out.", "fpn.latlayer2.", "fpn.latlayer3.")


def _is_trainable(name: str) -> bool:
    if name.startswith("PG.conv_1x1.") or ".layers." in name:
        return True
    return not any(tok in name for tok in _NO_GRAD_FPN)


class TemporalTrainer:
    def __init__(self, model, lr=1e-2, weight_decay=1e-5, loss_type="all", terl_pos_weight=False,
                 process_group=None, world_size=1):
        self.model = model
        self.lr, self.weight_decay = lr, weight_decay
        self.loss_type, self.terl_pos_weight = loss_type, terl_pos_weight
        self.pg, self.world = process_group, world_size
        named = [(n, p) for n, p in model.named_parameters() if _is_trainable(n)]
        self.names = [n for n, _ in named]
        total = sum(p.numel() for _, p in named)
        dev = named[0][1].device
        self.flat_p = torch.empty(total, device=dev, dtype=torch.float32)
        self.flat_g = torch.zeros(total, device=dev, dtype=torch.float32)
        off = 0
        for _, p in named:
            n = p.numel()
            self.flat_p[off:off + n].copy_(p.data.reshape(-1))
            p.data = self.flat_p[off:off + n].view(p.shape)
            p.grad = self.flat_g[off:off + n].view(p.shape)
            off += n
        self.num_params = total

    def forward_backward(self, x_rows, labels_u8, lengths):
        """x_rows: (frames, D) fp32 device; labels_u8: (frames, >=131) uint8 device, columns ivt|i|v|t.
        Returns the device tensor (loss, loss_i, loss_v, loss_t, loss_ivt)."""
        lay = SeqLayout.get(lengths, x_rows.device)
        self.flat_g.zero_()
        f_rows, logit_rows = self.model.forward_packed(x_rows, lay)
        total, li, lv, lt, livt = losses.tenco_loss(logit_rows, labels_u8, lay, self.model.head_sizes,
                                                    self.loss_type, self.terl_pos_weight)
        total.backward()
        return torch.stack([total.detach(), li, lv, lt, livt])

    def step(self, x_rows, labels_u8, lengths):
        out = self.forward_backward(x_rows, labels_u8, lengths)
        scale = 1.0
        if self.world > 1:
            torch.distributed.all_reduce(self.flat_g, group=self.pg)
            scale = 1.0 / self.world
        ops.sgd_step(self.flat_p, self.flat_g, self.lr, self.weight_decay, grad_scale=scale)
        return out


def lpt_assign(lengths, world):
    """Longest-processing-time-first assignment of videos to ranks (balances frames per rank)."""
    order = sorted(range(len(lengths)), key=lambda i: -lengths[i])
    loads = [0] * world
    shards = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: loads[k])
        shards[r].append(i)
        loads[r] += lengths[i]
    return shards
