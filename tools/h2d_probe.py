#!/usr/bin/env python
"""Pinned-host -> device copy bandwidth of this box (the ceiling of bench.py's end-to-end arm: 8 KB of fp32 features
per frame at D = 2048)."""
import json

import torch

dev = "cuda"
out = {}
for mb in (16, 136, 512):
    n = mb * (1 << 20) // 4
    h = torch.empty(n, dtype=torch.float32).pin_memory()
    d = torch.empty(n, dtype=torch.float32, device=dev)
    for _ in range(2):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        d.copy_(h, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    out[f"{mb}MB_GBps"] = round(mb * (1 << 20) * 10 / (e0.elapsed_time(e1) * 1e-3) / 1e9, 2)
out["frames_per_s_ceiling_D2048"] = round(out["136MB_GBps"] * 1e9 / (2048 * 4 + 132))
print(json.dumps(out))
