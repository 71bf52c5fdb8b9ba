"""Learning-rate schedule of the temporal heads.

``MT4MTLKD/Temporal_tenco/run.py:345-350``: SGD starts from ``lr / power`` and is driven by
``SequentialLR([LinearLR(start_factor=power, total_iters=warmup), ExponentialLR(gamma=decay_rate)],
milestones=[warmup + 1])``, stepped once per epoch (``:235-236``).  With the script defaults
(lr 0.01, power 0.1, warmup 58, decay 0.99) the rate climbs linearly 0.01 -> 0.1 over 58 epochs and then decays by
1 % per epoch.  ``WarmupExponentialLR.lr(epoch)`` is the closed form of that composition; the trainer writes it into
the device-side hyper-parameter block read by ``tcn_sgd_step_dev`` so the captured CUDA graph follows the schedule.
"""
from __future__ import annotations


class WarmupExponentialLR:
    def __init__(self, lr=0.01, power=0.1, warmup=58, decay_rate=0.99):
        self.base_lr = lr / power          # wp_lr (run.py:345)
        self.power, self.warmup, self.gamma = float(power), int(warmup), float(decay_rate)
        self.milestone = self.warmup + 1   # run.py:350
        self.epoch = 0

    def lr(self, epoch=None) -> float:
        """Rate used DURING epoch ``epoch`` (0-based; epoch e has seen e scheduler steps)."""
        e = self.epoch if epoch is None else int(epoch)
        if e < self.milestone:
            f = self.power + (1.0 - self.power) * min(e, self.warmup) / self.warmup
            return self.base_lr * f
        return self.base_lr * self.gamma ** (e - self.milestone)

    def step(self) -> float:
        self.epoch += 1
        return self.lr()
