#!/usr/bin/env python
"""H2D of one bench step's inputs (8 x 19 MB + 8 x 0.3 MB, pinned) while the TCN train step runs on the compute stream:
copy-stream count, stream priority, one big copy."""
import json
import os
import sys
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402

from computervision_codes_b200.tcn import VideoNas  # noqa: E402
from computervision_codes_b200.trainer import TemporalTrainer  # noqa: E402

dev = "cuda"
T, D = 2325, 2048
hx = [torch.empty(T, D).pin_memory() for _ in range(8)]
hl = [torch.empty(T, 132, dtype=torch.uint8).pin_memory() for _ in range(8)]
hbig = torch.empty(8 * T, D).pin_memory()
dx = torch.empty(8 * T, D, device=dev)
dl = torch.empty(8 * T, 132, dtype=torch.uint8, device=dev)
nbytes = 8 * T * (D * 4 + 132)
args = types.SimpleNamespace(fpn=True, output=False, feature=False, trans=False, mask=False, hier=False)
torch.manual_seed(0)
m = VideoNas(args, 11, 10, 3, 64, D, 100).to(dev).train()
tr = TemporalTrainer(m, max_frames=8 * T, max_seqs=8, input_mask_p=0.25)
x = torch.randn(8 * T, D, device=dev)
lab = torch.zeros(8 * T, 132, device=dev, dtype=torch.uint8)
for _ in range(3):
    tr.step(x, lab, [T] * 8)
torch.cuda.synchronize()
out = {}


def run(name, streams, big=False, busy=True):
    def copies():
        if big:
            with torch.cuda.stream(streams[0]):
                dx.copy_(hbig, non_blocking=True)
            return
        for i in range(8):
            with torch.cuda.stream(streams[i % len(streams)]):
                dx[i * T:(i + 1) * T].copy_(hx[i], non_blocking=True)
            with torch.cuda.stream(streams[(i + 1) % len(streams)]):
                dl[i * T:(i + 1) * T].copy_(hl[i], non_blocking=True)

    for _ in range(2):
        copies()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 12
    e0.record()
    for _ in range(n):
        if busy:
            tr.step(x, lab, [T] * 8)
        copies()
    for s in streams:
        torch.cuda.current_stream().wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    out[name] = {"ms_per_step": round(e0.elapsed_time(e1) / n, 3), "copy_GBps_if_copy_bound": round(nbytes * n / (e0.elapsed_time(e1) * 1e-3) / 1e9, 1)}


lo, hi = torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream, "priority_range") else (0, -1)
run("idle_2streams", [torch.cuda.Stream(), torch.cuda.Stream()], busy=False)
run("busy_1stream", [torch.cuda.Stream()])
run("busy_2streams", [torch.cuda.Stream(), torch.cuda.Stream()])
run("busy_3streams", [torch.cuda.Stream() for _ in range(3)])
run("busy_2streams_high_priority", [torch.cuda.Stream(priority=-1), torch.cuda.Stream(priority=-1)])
run("busy_4streams_high_priority", [torch.cuda.Stream(priority=-1) for _ in range(4)])
run("busy_1big_copy", [torch.cuda.Stream()], big=True)
run("busy_1big_copy_high_priority", [torch.cuda.Stream(priority=-1)], big=True)
run("compute_only", [torch.cuda.Stream()], big=True, busy=True) if False else None
print(json.dumps(out))
