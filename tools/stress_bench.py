#!/usr/bin/env python
"""BASELINE cfg5 (TERL long-sequence stress) on one GPU's share: 8 sequences x 8,000 frames x 768-d features,
VideoNas(fpn, 11/10/3, C=64), full train step (fwd + loss + bwd + SGD) through the trainer's CUDA graph.
Reports frames/s and the step's algorithmic HBM bytes (70.5 KB/frame at C=64, D=768) against the
measured HBM peak."""
import json
import os
import sys
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from computervision_codes_b200.tcn import VideoNas  # noqa: E402
from computervision_codes_b200.trainer import TemporalTrainer  # noqa: E402


def main():
    dev = "cuda"
    nseq = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    T, D = 8000, 768
    args = types.SimpleNamespace(fpn=True, output=False, feature=False, trans=False, mask=False, hier=False)
    torch.manual_seed(5)
    m = VideoNas(args, 11, 10, 3, 64, D, 100).to(dev).train()
    tr = TemporalTrainer(m, lr=1e-2, weight_decay=1e-5, max_frames=nseq * T, max_seqs=nseq, input_mask_p=0.25)
    x = torch.randn(nseq * T, D, device=dev)
    lab = (torch.rand(nseq * T, 132, device=dev) < 0.05).to(torch.uint8)
    lens = [T] * nseq
    for _ in range(3):
        tr.step(x, lab, lens)
    torch.cuda.synchronize()
    n = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        out = tr.step(x, lab, lens)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    peak = 6530.0
    if os.path.exists("MEASURED_PEAKS.json"):
        peak = float(json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"])
    alg = 70.5e3 * nseq * T
    print(json.dumps({"workload": f"cfg5 share: {nseq} x {T} frames x {D}-d, VideoNas(fpn,11/10/3,C=64), train step",
                      "ms_per_step": round(ms, 3), "frames_per_s": round(nseq * T / ms * 1e3),
                      "algorithmic_GB_per_step": round(alg / 1e9, 2), "alg_GBps": round(alg / ms / 1e6, 1),
                      "frac_of_measured_hbm": round(alg / ms / 1e6 / peak, 4), "loss": float(out[4])}))


if __name__ == "__main__":
    main()
