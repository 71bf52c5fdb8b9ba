#!/usr/bin/env python
"""Timeline of the graph-replayed TCN train step (bench workload, 8 videos / step) with torch.profiler (CUPTI):
per-kernel totals, the busy time of the main chain vs the step span, and the overlap won by the weight-gradient stream."""
import os
import sys
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from computervision_codes_b200.tcn import VideoNas  # noqa: E402
from computervision_codes_b200.trainer import TemporalTrainer  # noqa: E402

dev = "cuda"
torch.manual_seed(0)
D = 2048
lens = [2250] * 8
args = types.SimpleNamespace(fpn=True, output=False, feature=False, trans=False, mask=True, hier=False)
m = VideoNas(args, 11, 10, 3, 64, D, 100).to(dev).train()
if os.environ.get("PROF_CHAN") == "0":
    m.PG.channel_dropout.p = 0.0
tr = TemporalTrainer(m, lr=1e-2, weight_decay=1e-5, max_frames=sum(lens), max_seqs=8,
                     input_mask_p=float(os.environ.get("PROF_MASK", "0.25")))
x = torch.randn(sum(lens), D, device=dev)
lab = (torch.rand(sum(lens), 132, device=dev) < 0.05).to(torch.uint8)
for _ in range(4):
    tr.step(x, lab, lens)
torch.cuda.synchronize()
N = 4
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(N):
        tr.step(x, lab, lens)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_time_total > 0]
evs.sort(key=lambda e: e.time_range.start)
t0, t1 = evs[0].time_range.start, max(e.time_range.end for e in evs)
# union of busy intervals
busy, cur_s, cur_e = 0.0, None, None
for e in evs:
    s, en = e.time_range.start, e.time_range.end
    if cur_s is None:
        cur_s, cur_e = s, en
    elif s <= cur_e:
        cur_e = max(cur_e, en)
    else:
        busy += cur_e - cur_s
        cur_s, cur_e = s, en
busy += cur_e - cur_s
tot = sum(e.device_time_total for e in evs)
agg = {}
for e in evs:
    k = e.name.split("(")[0].replace("tcn::", "").replace("void ", "")[:44]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += e.device_time_total
print(f"steps {N}: span {(t1 - t0) / N / 1e3:.3f} ms/step, GPU busy (union) {busy / N / 1e3:.3f} ms/step, "
      f"sum of kernel times {tot / N / 1e3:.3f} ms/step")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
    print(f"{k:46s} n/step={n / N:6.1f} total/step={us / N:8.1f} us avg={us / n:7.2f} us share={us / tot:.3f}")

# ---- one step as a timeline: every kernel with its start offset, duration and stream; gaps on the busiest stream
if os.environ.get("TCN_TIMELINE"):
    step_evs = [e for e in evs if e.time_range.start >= t0 + (t1 - t0) * 2 / N and e.time_range.end <= t0 + (t1 - t0) * 3 / N + 50]
    base = step_evs[0].time_range.start
    last_end = {}
    print("# start_us dur_us gap_after_prev_on_stream  stream  kernel")
    for e in step_evs:
        sid = getattr(e, "stream", None) if hasattr(e, "stream") else None
        sid = sid if sid is not None else getattr(e, "device_resource_id", 0)
        gap = e.time_range.start - last_end.get(sid, e.time_range.start)
        last_end[sid] = e.time_range.end
        print(f"{e.time_range.start - base:9.1f} {e.time_range.end - e.time_range.start:7.1f} {gap:7.1f}  {sid}  {e.name.split('(')[0].replace('tcn::', '')[:40]}")
