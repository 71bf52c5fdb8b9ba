#!/usr/bin/env python
"""Input-gradient pass of one residual layer: the fused kernel (tcn_layer_bwd_tc, one launch) against the two
tcgen05 tap-GEMM launches it replaces, at the three BASELINE shapes.  CUDA events over a graph of 24 launches;
algorithmic bytes per frame of the fused kernel: 12 * C (read gy, write gu, write gx) + 16 (mask words)."""
import ctypes as Ct
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from computervision_codes_b200 import _lib, ops  # noqa: E402
from computervision_codes_b200.layout import SeqLayout  # noqa: E402
from tools.kernel_bench import PEAK, timeit  # noqa: E402

DEV = "cuda"


def main():
    C = 64
    lib = _lib.load()
    shapes = {"1x1800": [1800], "8x2250": [2250] * 8, "64x8000": [8000] * 64}
    for name in sys.argv[1:] or list(shapes):
        lens = shapes[name]
        lay = SeqLayout.get(lens, DEV)
        frames = sum(lens)
        nbuf = 6 if frames > 100000 else 12   # rotate buffers: > L2 for the stress shape
        gys = [torch.randn(lay.rows, C, device=DEV) for _ in range(nbuf)]
        gus = [torch.zeros(lay.rows, C, device=DEV) for _ in range(nbuf)]
        gxs = [torch.zeros(lay.rows, C, device=DEV) for _ in range(nbuf)]
        hs = [torch.randn(lay.rows, C, device=DEV) for _ in range(nbuf)]
        masks = [torch.randint(-2 ** 31, 2 ** 31 - 1, (lay.rows, 4), device=DEV, dtype=torch.int32) for _ in range(nbuf)]
        w1 = torch.randn(C, C, 3, device=DEV) / (3 * C) ** 0.5
        w2 = torch.randn(C, C, 1, device=DEV) / C ** 0.5
        h1t, l1t = ops.split_weight(w1, True)
        h2t, l2t = ops.split_weight(w2, True)
        for d in (1, 16, 512):
            shifts = (-d, 0, d)
            it = [0]

            def fused():
                i = it[0] % nbuf
                it[0] += 1
                a = _lib.LayerBwdTcArgs()
                a.gy, a.g_rows, a.gu, a.gx, a.masks = (gys[i].data_ptr(), lay.rows, gus[i].data_ptr(), gxs[i].data_ptr(),
                                                       masks[i].data_ptr())
                a.w2t_hi, a.w2t_lo, a.w1t_hi, a.w1t_lo = h2t.data_ptr(), l2t.data_ptr(), h1t.data_ptr(), l1t.data_ptr()
                a.meta, a.nblk, a.channels = lay.meta.data_ptr(), lay.nblk, C
                for k, s in enumerate(shifts):
                    a.shift[k] = s
                a.drop_p = 0.5
                _lib.check(lib.tcn_layer_bwd_tc(Ct.byref(a), _lib.stream_ptr()))

            def two_launches():
                i = it[0] % nbuf
                it[0] += 1
                ops.gemm_tc(gys[i], h2t, l2t, lay, C, C, (0,), out=gus[i], relu_mask=hs[i], in_drop_p=0.5,
                            in_drop_rescale=True, seed=1, stream_id=2)
                ops.gemm_tc(gus[i], h1t, l1t, lay, C, C, tuple(-s for s in shifts), out=gxs[i], residual=gys[i])

            for nm, fn, byt in (("layer_bwd_tc (fused, gu via TMEM)", fused, 12 * C + 16),
                                ("gemm_tc dgrad1 + dgrad2 (two launches)", two_launches, 24 * C)):
                ms = timeit(fn)
                print(json.dumps({"kernel": nm, "shape": name, "dilation": d, "ms": round(ms, 4),
                                  "alg_bytes_per_frame": byt, "alg_GBps": round(byt * frames / ms / 1e6, 1),
                                  "frac_hbm": round(byt * frames / ms / 1e6 / PEAK, 4)}), flush=True)


if __name__ == "__main__":
    main()
