#!/usr/bin/env python
"""BASELINE cfg3: MS-TCT temporal head (TemporalEncoder(768,[256,384,576,864],8 heads,mlp 8,2 blocks) + mixer +
classifier) over 31 windows x 256 frames of 768-d features: forward + BCE loss + backward, CUDA events."""
import json
import os
import sys
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from computervision_codes_b200 import losses  # noqa: E402
from computervision_codes_b200.mstct import VideoNas  # noqa: E402


def main():
    dev = "cuda"
    torch.manual_seed(0)
    B, T, D = 31, 256, 768
    m = VideoNas(types.SimpleNamespace(loss_type="ivt"), [256, 384, 576, 864], 2, 8, 8, D, 512).to(dev).train()
    nparams = sum(p.numel() for p in m.parameters())
    x = torch.randn(B, D, T, device=dev)
    lab = (torch.rand(B * T, 100, device=dev) < 0.05).float()

    def step():
        for p in m.parameters():
            p.grad = None
        y = m(x)[3][0]
        loss = losses.bce_with_logits(y.reshape(B * T, 100), lab)
        loss.backward()
        return loss

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    n = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    flops = 2.93e12  # fwd + bwd, SURVEY 8(d)
    print(json.dumps({"workload": "cfg3 MS-TCT 31 x 256 x 768, fwd+loss+bwd (per-op Python path)", "params": nparams,
                      "ms_per_step": round(ms, 2), "frames_per_s": round(B * T / ms * 1e3),
                      "algorithmic_TFLOPs": round(flops / ms / 1e9, 1), "executed_TFLOPs_3xtf32": round(3 * flops / ms / 1e9, 1),
                      "loss": float(loss)}))


if __name__ == "__main__":
    main()
