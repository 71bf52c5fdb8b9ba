#!/usr/bin/env python
"""Per-kernel timings (CUDA events) at the BASELINE shapes: one 1,800-frame video and the TERL stress
shape (64 x 8,000 frames); prints one JSON line per kernel with algorithmic GB/s and executed TFLOP/s."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from computervision_codes_b200 import ops  # noqa: E402
from computervision_codes_b200.layout import SeqLayout  # noqa: E402

DEV = "cuda"
PEAK = 6530.0
if os.path.exists("MEASURED_PEAKS.json"):
    PEAK = float(json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"])


def timeit(fn, n=24, warm=3):
    """Average GPU time per launch: n launches captured in one CUDA graph (no host launch overhead)."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    C = 64
    shapes = {"1x1800": [1800], "8x2250": [2250] * 8, "64x8000": [8000] * 64}
    only = sys.argv[1:] or list(shapes)
    for name in only:
        lens = shapes[name]
        lay = SeqLayout.get(lens, DEV)
        frames = sum(lens)
        nbuf = 6 if frames > 100000 else 12   # rotate buffers: > L2 for the stress shape
        xs = [torch.randn(lay.rows, C, device=DEV) for _ in range(nbuf)]
        ys = [torch.zeros(lay.rows, C, device=DEV) for _ in range(nbuf)]
        hs = [torch.zeros(lay.rows, C, device=DEV) for _ in range(nbuf)]
        w1 = torch.randn(C, C, 3, device=DEV) / (3 * C) ** 0.5
        w2 = torch.randn(C, C, 1, device=DEV) / C ** 0.5
        b = torch.zeros(C, device=DEV)
        w1f, w2f = ops.prep_weight(w1), ops.prep_weight(w2)
        w1t, w2t = ops.prep_weight(w1, True), ops.prep_weight(w2, True)
        from computervision_codes_b200 import _lib
        import ctypes as Ct
        lib = _lib.load()
        for d in (1, 16, 512):
            shifts = (-d, 0, d)
            it = [0]

            def fused():
                i = it[0] % nbuf
                it[0] += 1
                a = _lib.LayerFwdArgs()
                a.x, a.y, a.h = xs[i].data_ptr(), ys[i].data_ptr(), hs[i].data_ptr()
                a.w1f, a.w2f, a.b1, a.b2 = w1f.data_ptr(), w2f.data_ptr(), b.data_ptr(), b.data_ptr()
                a.meta, a.nblk, a.channels = lay.meta.data_ptr(), lay.nblk, C
                for k, s in enumerate(shifts):
                    a.shift[k] = s
                a.drop_p, a.drop_seed, a.drop_stream = 0.5, 1, 2
                _lib.check(lib.tcn_layer_fwd(Ct.byref(a), _lib.stream_ptr()))

            ms = timeit(fused)
            gbs = 8 * C * frames / ms / 1e6
            tf = 8 * C * C * frames * 3 / ms / 1e9
            print(json.dumps({"kernel": "layer_fwd64", "shape": name, "dilation": d, "ms": round(ms, 4),
                              "alg_GBps": round(gbs, 1), "frac_hbm": round(gbs / PEAK, 4),
                              "executed_TFLOPs_3xtf32": round(tf, 1)}))
            w1h, w1l = ops.split_weight(w1)
            w2h, w2l = ops.split_weight(w2)

            def fused_tc():
                i = it[0] % nbuf
                it[0] += 1
                a = _lib.LayerFwdTcArgs()
                a.x, a.x_rows, a.y, a.h = xs[i].data_ptr(), xs[i].shape[0], ys[i].data_ptr(), hs[i].data_ptr()
                a.w1_hi, a.w1_lo, a.w2_hi, a.w2_lo = w1h.data_ptr(), w1l.data_ptr(), w2h.data_ptr(), w2l.data_ptr()
                a.b1, a.b2 = b.data_ptr(), b.data_ptr()
                a.meta, a.nblk, a.channels = lay.meta.data_ptr(), lay.nblk, C
                for k, s in enumerate(shifts):
                    a.shift[k] = s
                a.drop_p, a.drop_seed, a.drop_stream = 0.5, 1, 2
                _lib.check(lib.tcn_layer_fwd_tc(Ct.byref(a), _lib.stream_ptr()))

            ms = timeit(fused_tc)
            gbs = 12 * C * frames / ms / 1e6   # read x, write h, write y
            tf = 8 * C * C * frames * 3 / ms / 1e9
            print(json.dumps({"kernel": "layer_fwd_tc (tcgen05, h via TMEM)", "shape": name, "dilation": d,
                              "ms": round(ms, 4), "alg_GBps": round(gbs, 1), "frac_hbm": round(gbs / PEAK, 4),
                              "executed_TFLOPs_3xtf32": round(tf, 1)}))
        # backward pieces of one layer (d = 16)
        shifts = (-16, 0, 16)
        it = [0]

        def dgrad1():
            i = it[0] % nbuf; it[0] += 1
            ops.tapgemm(xs[i], w2t, lay, C, C, (0,), out=ys[i], relu_mask=hs[i], in_drop_p=0.5, seed=1, stream_id=2)

        def dgrad2():
            i = it[0] % nbuf; it[0] += 1
            ops.tapgemm(xs[i], w1t, lay, C, C, tuple(-s for s in shifts), out=ys[i], residual=hs[i])

        gw1, gb1 = torch.zeros(C, C, 3, device=DEV), torch.zeros(C, device=DEV)
        gw2, gb2 = torch.zeros(C, C, 1, device=DEV), torch.zeros(C, device=DEV)

        def wg1():
            i = it[0] % nbuf; it[0] += 1
            ops.wgrad(xs[i], hs[i], lay, C, C, shifts, gw1, gb1)

        def wg2():
            i = it[0] % nbuf; it[0] += 1
            ops.wgrad(xs[i], hs[i], lay, C, C, (0,), gw2, gb2, g_drop_p=0.5, seed=1, stream_id=2)

        for nm, fn, flops, byt in (("tapgemm dgrad1 (K=64, relu-mask, drop-on-load)", dgrad1, 2 * C * C, 12 * C),
                                   ("tapgemm dgrad2 (3 taps, residual)", dgrad2, 6 * C * C, 12 * C),
                                   ("wgrad W1 (3 taps)", wg1, 6 * C * C, 8 * C), ("wgrad W2", wg2, 2 * C * C, 8 * C)):
            ms = timeit(fn)
            print(json.dumps({"kernel": nm, "shape": name, "ms": round(ms, 4),
                              "alg_GBps": round(byt * frames / ms / 1e6, 1),
                              "executed_TFLOPs_3xtf32": round(flops * frames * 3 / ms / 1e9, 1)}))
        # the same four shapes on the tcgen05 tap GEMM
        h1, l1 = ops.split_weight(w1)
        h2, l2 = ops.split_weight(w2)
        h1t, l1t = ops.split_weight(w1, True)
        h2t, l2t = ops.split_weight(w2, True)

        def tc_conv3():
            i = it[0] % nbuf; it[0] += 1
            ops.gemm_tc(xs[i], h1, l1, lay, C, C, shifts, bias=b, out=ys[i], relu=True)

        def tc_conv1():
            i = it[0] % nbuf; it[0] += 1
            ops.gemm_tc(xs[i], h2, l2, lay, C, C, (0,), bias=b, out=ys[i], residual=hs[i], drop_p=0.5, seed=1, stream_id=2)

        def tc_dgrad1():
            i = it[0] % nbuf; it[0] += 1
            ops.gemm_tc(xs[i], h2t, l2t, lay, C, C, (0,), out=ys[i], relu_mask=hs[i], in_drop_p=0.5, in_drop_rescale=True,
                        seed=1, stream_id=2)

        def tc_dgrad2():
            i = it[0] % nbuf; it[0] += 1
            ops.gemm_tc(xs[i], h1t, l1t, lay, C, C, tuple(-s for s in shifts), out=ys[i], residual=hs[i])

        for nm, fn, flops, byt in (("gemm_tc conv3+relu", tc_conv3, 6 * C * C, 8 * C),
                                   ("gemm_tc conv1x1+drop+res", tc_conv1, 2 * C * C, 12 * C),
                                   ("gemm_tc dgrad1", tc_dgrad1, 2 * C * C, 12 * C),
                                   ("gemm_tc dgrad2", tc_dgrad2, 6 * C * C, 12 * C)):
            ms = timeit(fn)
            print(json.dumps({"kernel": nm, "shape": name, "ms": round(ms, 4),
                              "alg_GBps": round(byt * frames / ms / 1e6, 1),
                              "executed_TFLOPs_3xtf32": round(flops * frames * 3 / ms / 1e9, 1)}))
        def tc_wg1():
            i = it[0] % nbuf; it[0] += 1
            ops.wgrad_tc(xs[i], hs[i], lay, C, C, shifts, gw1, gb1)

        def tc_wg2():
            i = it[0] % nbuf; it[0] += 1
            ops.wgrad_tc(xs[i], hs[i], lay, C, C, (0,), gw2, gb2, g_drop_p=0.5, seed=1, stream_id=2)

        for nm, fn, flops, byt in (("wgrad_tc W1 (3 taps)", tc_wg1, 6 * C * C, 8 * C), ("wgrad_tc W2", tc_wg2, 2 * C * C, 8 * C)):
            ms = timeit(fn)
            print(json.dumps({"kernel": nm, "shape": name, "ms": round(ms, 4),
                              "alg_GBps": round(byt * frames / ms / 1e6, 1),
                              "executed_TFLOPs_3xtf32": round(flops * frames * 3 / ms / 1e9, 1)}))
        # stage-input projection 2048 -> 64
        if True:
            D = 2048 if frames <= 100000 else 768
            x = torch.randn(frames, D, device=DEV)
            wp = ops.prep_weight(torch.randn(C, D, device=DEV) / D ** 0.5)
            out = torch.zeros(lay.rows, C, device=DEV)
            ms = timeit(lambda: ops.tapgemm(x, wp, lay, D, C, (0,), bias=b, out=out, x_unpadded=True))
            print(json.dumps({"kernel": "tapgemm projection 2048->64", "shape": name, "ms": round(ms, 4),
                              "alg_GBps": round(4 * (D + C) * frames / ms / 1e6, 1),
                              "frac_hbm": round(4 * (D + C) * frames / ms / 1e6 / PEAK, 4),
                              "executed_TFLOPs_3xtf32": round(2 * D * C * frames * 3 / ms / 1e9, 1)}))
            whi, wlo = ops.split_weight(torch.randn(C, D, device=DEV) / D ** 0.5)
            ms = timeit(lambda: ops.gemm_tc(x, whi, wlo, lay, D, C, bias=b, out=out, x_unpadded=True))
            print(json.dumps({"kernel": "gemm_tc (tcgen05+TMA) projection 2048->64", "shape": name, "ms": round(ms, 4),
                              "alg_GBps": round(4 * (D + C) * frames / ms / 1e6, 1),
                              "frac_hbm": round(4 * (D + C) * frames / ms / 1e6 / PEAK, 4),
                              "executed_TFLOPs_3xtf32": round(2 * D * C * frames * 3 / ms / 1e9, 1)}))
            gw = torch.zeros(C, D, 1, device=DEV)
            g = torch.randn(lay.rows, C, device=DEV)
            ms = timeit(lambda: ops.wgrad_tc(g, x, lay, C, D, (0,), gw, None, x_unpadded=True))
            print(json.dumps({"kernel": "wgrad_tc projection", "shape": name, "ms": round(ms, 4),
                              "alg_GBps": round(4 * (D + C) * frames / ms / 1e6, 1),
                              "frac_hbm": round(4 * (D + C) * frames / ms / 1e6 / PEAK, 4),
                              "executed_TFLOPs_3xtf32": round(2 * D * C * frames * 3 / ms / 1e9, 1)}))
            ms = timeit(lambda: ops.wgrad(g, x, lay, C, D, (0,), gw, None, x_unpadded=True))
            print(json.dumps({"kernel": "wgrad projection", "shape": name, "ms": round(ms, 4),
                              "alg_GBps": round(4 * (D + C) * frames / ms / 1e6, 1),
                              "executed_TFLOPs_3xtf32": round(2 * D * C * frames * 3 / ms / 1e9, 1)}))


if __name__ == "__main__":
    main()
