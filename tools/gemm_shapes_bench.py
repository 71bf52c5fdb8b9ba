#!/usr/bin/env python
"""Per-shape timing of the tcgen05 GEMM (tcn_gemm_tc) and weight-gradient (tcn_wgrad_tc) kernels at the cfg3 MS-TCT
shapes (31 windows x 256 frames at every stage, channels grow per stage; qkv / proj / fc1 / fc2 of each block).  CUDA events over a graph."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from computervision_codes_b200 import ops  # noqa: E402
from computervision_codes_b200.layout import SeqLayout  # noqa: E402


def timed(fn, iters=20):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    dev = "cuda"
    torch.manual_seed(0)
    out = []
    B = 31
    for T, C in ((256, 256), (256, 384), (256, 576), (256, 864)):
        lay = SeqLayout([T] * B, dev)
        for name, ci, no in (("qkv", C, 3 * C), ("proj", C, C), ("fc1", C, 8 * C), ("fc2", 8 * C, C)):
            x = torch.randn(lay.rows, ci, device=dev)
            w = torch.randn(no, ci, device=dev) * 0.05
            hi, lo = ops.split_weight(w)
            y = torch.zeros(lay.rows, no, device=dev)
            ms = timed(lambda: ops.gemm_tc(x, hi, lo, lay, ci, no, out=y))
            fl = 2.0 * B * T * ci * no
            g = torch.randn(lay.rows, no, device=dev)
            dw = torch.zeros(no, ci, 1, device=dev)
            ms_w = timed(lambda: ops.wgrad_tc(g, x, lay, no, ci, (0,), dw))
            out.append({"T": T, "C": C, "op": name, "rows": B * T, "c_in": ci, "n_out": no,
                        "gemm_us": round(ms * 1e3, 1), "gemm_alg_TFLOPs": round(fl / ms / 1e9, 1),
                        "wgrad_us": round(ms_w * 1e3, 1), "wgrad_alg_TFLOPs": round(fl / ms_w / 1e9, 1)})
            print(json.dumps(out[-1]), flush=True)
    tot_g = sum(o["gemm_us"] for o in out)
    tot_w = sum(o["wgrad_us"] for o in out)
    print(json.dumps({"sum_gemm_us": round(tot_g, 1), "sum_wgrad_us": round(tot_w, 1)}))


if __name__ == "__main__":
    main()
