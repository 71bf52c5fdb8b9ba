#!/usr/bin/env python
"""The three per-layer kernels (tcn_layer_fwd_tc, tcn_layer_bwd_tc, tcn_wgrad_layer) alone at the TERL stress shape
(64 x 8000 frames), in turn, for an `ncu --set full --import-source on` capture of one launch of each:
  ncu --set full --import-source on --clock-control none -k "regex:layer_(fwd|bwd)_tc_kernel|wgrad_layer_kernel" \\
      --launch-skip 6 --launch-count 3 -o gpurun_out/fused_stress python tools/exp/fused_stress.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402

from computervision_codes_b200 import ops  # noqa: E402
from computervision_codes_b200.layout import SeqLayout  # noqa: E402

DEV = "cuda"
C = 64
lens = [8000] * 64 if len(sys.argv) < 2 else [int(sys.argv[1])] * int(sys.argv[2])
d = 16
shifts = (-d, 0, d)
lay = SeqLayout.get(lens, DEV)
torch.manual_seed(0)
w1 = torch.randn(C, C, 3, device=DEV) / (3 * C) ** 0.5
w2 = torch.randn(C, C, 1, device=DEV) / C ** 0.5
b = torch.zeros(C, device=DEV)
xs = [torch.randn(lay.rows, C, device=DEV) for _ in range(3)]
gys = [torch.randn(lay.rows, C, device=DEV) for _ in range(3)]
gw1, gb1 = torch.zeros(C, C, 3, device=DEV), torch.zeros(C, device=DEV)
gw2, gb2 = torch.zeros(C, C, 1, device=DEV), torch.zeros(C, device=DEV)
for i in range(4):
    y, h, masks = ops.layer_fwd_tc(xs[i % 3], w1, w2, b, b, lay, shifts, True, 0.5, 1, 2, save_masks=True)
    gu, gx = ops.layer_bwd_tc(gys[i % 3], masks, w1, w2, lay, shifts, 0.5)
    ops.layer_wgrad(gu, xs[i % 3], gys[i % 3], h, lay, shifts, gw1, gb1, gw2, gb2, drop_p=0.5, seed=1, stream_id=2, masks=masks)
    del y, h, gu, gx
torch.cuda.synchronize()
print("ok")
