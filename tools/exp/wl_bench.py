#!/usr/bin/env python
"""Isolated timing of tcn_wgrad_layer (kernel + slab reduction) at the stress shape and at the bench step's shape."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402

from computervision_codes_b200 import ops  # noqa: E402
from computervision_codes_b200.layout import SeqLayout  # noqa: E402

DEV, C = "cuda", 64
NBUF = int(os.environ.get("WL_NBUF", "6"))
for lens in ([8000] * 64, [8000] * 8, [2250] * 8):
    lay = SeqLayout.get(lens, DEV)
    shifts = (-32, -16, 0)
    bufs = [[torch.randn(lay.rows, C, device=DEV) for _ in range(4)] for _ in range(NBUF)]
    masks = torch.randint(-2 ** 31, 2 ** 31 - 1, (lay.rows, 4), device=DEV, dtype=torch.int32)
    gw1, gb1 = torch.zeros(C, C, 3, device=DEV), torch.zeros(C, device=DEV)
    gw2, gb2 = torch.zeros(C, C, 1, device=DEV), torch.zeros(C, device=DEV)
    it = [0]

    def fn():
        gu, x, gy, h = bufs[it[0] % NBUF]
        it[0] += 1
        ops.layer_wgrad(gu, x, gy, h, lay, shifts, gw1, gb1, gw2, gb2, drop_p=0.5, seed=1, stream_id=2, masks=masks)

    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    n = 24
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"frames {sum(lens)}: {ms * 1e3:.1f} us per call, {1024 * sum(lens) / ms / 1e6:.0f} GB/s algorithmic (16 C B/frame)")
