#!/usr/bin/env python
"""Kernel-time breakdown of one MS-TCT (cfg3) train step with torch.profiler (CUPTI, no replay)."""
import os
import sys
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from computervision_codes_b200 import losses  # noqa: E402
from computervision_codes_b200.mstct import VideoNas  # noqa: E402

dev = "cuda"
torch.manual_seed(0)
B, T, D = 31, 256, 768
m = VideoNas(types.SimpleNamespace(loss_type="ivt"), [256, 384, 576, 864], 2, 8, 8, D, 512).to(dev).train()
x = torch.randn(B, D, T, device=dev)
lab = (torch.rand(B * T, 100, device=dev) < 0.05).float()


def step():
    for p in m.parameters():
        p.grad = None
    y = m(x)[3][0]
    loss = losses.bce_with_logits(y.reshape(B * T, 100), lab)
    loss.backward()


for _ in range(2):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:18]
tot = sum(e.device_time_total for e in prof.key_averages())
for e in rows:
    print(f"{e.key[:70]:70s} n={e.count:4d} total={e.device_time_total / 1e3:8.2f} ms  share={e.device_time_total / tot:.3f}")
print("total device ms", tot / 1e3)
