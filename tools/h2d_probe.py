#!/usr/bin/env python
"""Pinned-host -> device copy bandwidth of this box (the ceiling of bench.py's end-to-end arm: 8 KB of fp32 features
per frame at D = 2048): one big transfer, and a step's worth of per-video transfers (8 x 19 MB + 8 x 0.3 MB) spread
over 1 / 2 / 4 copy streams, with and without a compute kernel running."""
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

T, D = 2325, 2048
hx = [torch.empty(T, D).pin_memory() for _ in range(8)]
hl = [torch.empty(T, 132, dtype=torch.uint8).pin_memory() for _ in range(8)]
dx = torch.empty(8 * T, D, device=dev)
dl = torch.empty(8 * T, 132, dtype=torch.uint8, device=dev)
nbytes = 8 * T * (D * 4 + 132)
a = torch.randn(4096, 4096, device=dev)


def step_copies(streams):
    cur = torch.cuda.current_stream()
    ev = torch.cuda.Event()
    ev.record(cur)
    for s in streams:
        s.wait_event(ev)
    for i in range(8):
        with torch.cuda.stream(streams[i % len(streams)]):
            dx[i * T:(i + 1) * T].copy_(hx[i], non_blocking=True)
        with torch.cuda.stream(streams[(i + 1) % len(streams)]):
            dl[i * T:(i + 1) * T].copy_(hl[i], non_blocking=True)
    for s in streams:
        e = torch.cuda.Event()
        e.record(s)
        cur.wait_event(e)


for busy in (False, True):
    for ns in (1, 2, 4):
        streams = [torch.cuda.Stream() for _ in range(ns)]
        comp = torch.cuda.Stream()
        for _ in range(2):
            step_copies(streams)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            if busy:
                with torch.cuda.stream(comp):
                    for _ in range(6):
                        a @ a
            step_copies(streams)
        e1.record()
        torch.cuda.synchronize()
        out[f"step_copies_{ns}streams_{'busy' if busy else 'idle'}_GBps"] = round(
            nbytes * 10 / (e0.elapsed_time(e1) * 1e-3) / 1e9, 2)
print(json.dumps(out))
