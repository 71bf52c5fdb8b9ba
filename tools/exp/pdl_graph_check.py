#!/usr/bin/env python
"""Run-to-run identity of graph-replayed training steps: R trainers started from the same weights and seeds must hold
bit-identical parameters after every step (every reduction of the executor is fixed-order).  Eight different batches
(ragged lengths, the BASELINE width) are cycled so that a stale read of an earlier step's activations changes the result.
    python tools/exp/pdl_graph_check.py [steps] [repeats]          # prints one digest line per mark and repeat
Environment switches of the library (TCN_NO_PDL, TCN_NO_WGRAD_STREAM, TCN_NO_FUSED_TC, ...) select what is compared."""
import hashlib
import os
import sys
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402

from computervision_codes_b200.tcn import VideoNas  # noqa: E402
from computervision_codes_b200.trainer import TemporalTrainer  # noqa: E402

DEV = "cuda"
STEPS = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
REPEATS = int(sys.argv[2]) if len(sys.argv) > 2 else 2
args = types.SimpleNamespace(fpn=True, output=False, feature=False, trans=False, mask=True, hier=False)
g = torch.Generator().manual_seed(2)
batches = []
for b in range(8):
    lens = [int(t) for t in torch.randint(900, 3600, (8,), generator=g)]
    x = torch.randn(sum(lens), 2048, generator=g).to(DEV)
    lab = (torch.rand(sum(lens), 132, generator=g) < 0.05).to(torch.uint8).to(DEV)
    batches.append((x, lab, lens))
marks = sorted({1, 10, 100, 200, 400, 800, 1500, 3000, STEPS} & set(range(1, STEPS + 1)))
digests = []
for rep in range(REPEATS):
    torch.manual_seed(1)
    m = VideoNas(args, 11, 10, 3, 64, 2048, 100).to(DEV).train()
    tr = TemporalTrainer(m, lr=1e-3, weight_decay=1e-5, max_frames=8 * 3600 + 8 * 128, max_seqs=8, seed=5, input_mask_p=0.25)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out = []
    for s in range(1, STEPS + 1):
        if s == max(1, STEPS - 499):
            e0.record()
        x, lab, lens = batches[s % 8]
        tr.step(x, lab, lens)
        if s in marks:
            torch.cuda.synchronize()
            out.append(hashlib.sha256(tr.flat_p.cpu().numpy().tobytes()).hexdigest()[:16])
    e1.record()
    torch.cuda.synchronize()
    digests.append(out)
    sys.stderr.write(f"repeat {rep}: {e0.elapsed_time(e1) / min(500, STEPS):.4f} ms/step\n")
    tr.close()
    del tr, m
for i, s in enumerate(marks):
    row = [d[i] for d in digests]
    print(s, " ".join(row), "same" if len(set(row)) == 1 else "DIFFERENT")
