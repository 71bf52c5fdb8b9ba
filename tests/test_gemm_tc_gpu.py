"""tcgen05 + TMA GEMM (csrc/gemm_tc.cu) against an fp64 reference.  GPU only."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _maxabs(a, b):
    return float((a.detach().double().cpu() - b.detach().double().cpu()).abs().max())


@pytest.mark.parametrize("lengths,K,N,unpadded", [([1800], 2048, 64, True), ([300, 170, 129], 96, 64, True),
                                                   ([256, 40], 768, 128, False), ([130], 32, 192, False),
                                                   ([1000, 999], 512, 256, True), ([200, 77], 132, 64, False),
                                                   ([140], 64, 131, False), ([90], 100, 7, True)])
def test_gemm_tc_matches_fp64(lengths, K, N, unpadded):
    from computervision_codes_b200 import ops
    from computervision_codes_b200.layout import SeqLayout

    torch.manual_seed(K + N)
    lay = SeqLayout.get(lengths, DEV)
    xs = [torch.randn(T, K) for T in lengths]
    w = torch.randn(N, K) / K ** 0.5
    b = torch.randn(N)
    if unpadded:
        x = torch.cat(xs).to(DEV)
    else:
        x = torch.zeros(lay.rows, K)
        for s, T in enumerate(lengths):
            x[lay.starts[s]:lay.starts[s] + T] = xs[s]
        x = x.to(DEV)
    hi, lo = ops.split_weight(w.to(DEV))
    assert torch.equal((hi + lo).cpu()[:N, :K], w)
    y = ops.gemm_tc(x, hi, lo, lay, K, N, bias=b.to(DEV), x_unpadded=unpadded)
    torch.cuda.synchronize()
    for s, T in enumerate(lengths):
        ref = xs[s].double() @ w.double().t() + b.double()
        got = y[lay.starts[s]:lay.starts[s] + T, :N]
        err = _maxabs(got, ref)
        assert err <= 4e-6 * max(1.0, float(ref.abs().max())) + 1e-6 * K ** 0.5, (lengths, K, N, err)
    # rows between sequences are never written
    for s, T in enumerate(lengths):
        end = lay.starts[s + 1] if s + 1 < len(lengths) else lay.rows
        assert float(y[lay.starts[s] + T:end].abs().max() if end > lay.starts[s] + T else 0.0) == 0.0
    if y.shape[1] > N:
        assert float(y[:, N:].abs().max()) == 0.0


def test_gemm_tc_channel_scale_and_input_mask():
    from computervision_codes_b200 import ops
    from computervision_codes_b200.layout import SeqLayout

    torch.manual_seed(1)
    lengths, K, N = [200, 150], 128, 64
    lay = SeqLayout.get(lengths, DEV)
    xs = [torch.randn(T, K) for T in lengths]
    w = torch.randn(N, K) / K ** 0.5
    hi, lo = ops.split_weight(w.to(DEV))
    cs = (torch.rand(len(lengths), K) > 0.5).float() * 2.0
    seed, sid = 99, 0x7fff0002
    y = ops.gemm_tc(torch.cat(xs).to(DEV), hi, lo, lay, K, N, x_unpadded=True, colscale=cs.to(DEV), in_drop_p=0.25,
                    in_drop_rescale=False, seed=seed, stream_id=sid)
    keep = ops.dropout_keep_mask(lay.rows, K, 0.25, seed, sid, DEV).cpu().double()
    assert 0.72 < float(keep.mean()) < 0.78
    for s, T in enumerate(lengths):
        r0 = lay.starts[s]
        xm = xs[s].double() * cs[s].double() * keep[r0:r0 + T]
        ref = xm @ w.double().t()
        assert _maxabs(y[r0:r0 + T], ref) <= 1e-5


@pytest.mark.parametrize("shifts", [(-1, 0, 1), (-4, 0, 4), (-512, 0, 512), (-2, -1, 0), (-64, -32, 0), (0,)])
def test_gemm_tc_taps_and_epilogue_match_mma_sync_path_and_fp64(shifts):
    """Taps, sequence-boundary zeroing and the full epilogue: tcgen05 kernel == mma.sync kernel == fp64."""
    from computervision_codes_b200 import ops
    from computervision_codes_b200.layout import SeqLayout
    from oracle import tcn_oracle as O

    torch.manual_seed(5 + len(shifts) + abs(shifts[0]))
    lengths, C = [300, 129, 1000], 64
    lay = SeqLayout.get(lengths, DEV)
    ntaps = len(shifts)
    w = torch.randn(C, C, ntaps) / (C * ntaps) ** 0.5
    b = torch.randn(C)
    x = torch.zeros(lay.rows, C)
    res = torch.zeros(lay.rows, C)
    msk = torch.zeros(lay.rows, C)
    xs = []
    for s, T in enumerate(lengths):
        xs.append(torch.randn(T, C))
        x[lay.starts[s]:lay.starts[s] + T] = xs[-1]
        res[lay.starts[s]:lay.starts[s] + T] = torch.randn(T, C)
        msk[lay.starts[s]:lay.starts[s] + T] = torch.randn(T, C)
    x, res, msk = x.to(DEV), res.to(DEV), msk.to(DEV)
    hi, lo = ops.split_weight(w.to(DEV))
    y = ops.gemm_tc(x, hi, lo, lay, C, C, shifts, bias=b.to(DEV), residual=res, relu_mask=msk, drop_p=0.5, seed=3,
                    stream_id=9)
    y2 = ops.tapgemm(x, ops.prep_weight(w.to(DEV)), lay, C, C, shifts, bias=b.to(DEV), residual=res, relu_mask=msk,
                     drop_p=0.5, seed=3, stream_id=9)
    assert _maxabs(y, y2) <= 5e-5
    keep = ops.dropout_keep_mask(lay.rows, C, 0.5, 3, 9, DEV).cpu().double()
    for s, T in enumerate(lengths):
        r0 = lay.starts[s]
        u = O.conv_taps(xs[s].double().t().unsqueeze(0), w.double(), b.double(), shifts)[0].t()
        u = u * (msk[r0:r0 + T].cpu().double() > 0) * keep[r0:r0 + T] * 2.0 + res[r0:r0 + T].cpu().double()
        assert _maxabs(y[r0:r0 + T], u) <= 5e-5
    # relu + dropout applied to the loaded operand (the backward of a dropout), transposed weights
    hit, lot = ops.split_weight(w.to(DEV), transpose=True)
    g = ops.gemm_tc(x, hit, lot, lay, C, C, tuple(-s for s in shifts), relu=True, in_drop_p=0.5, in_drop_rescale=True,
                    seed=3, stream_id=9)
    g2 = ops.tapgemm(x, ops.prep_weight(w.to(DEV), transpose=True), lay, C, C, tuple(-s for s in shifts), relu=True,
                     in_drop_p=0.5, seed=3, stream_id=9)
    assert _maxabs(g, g2) <= 5e-5


@pytest.mark.parametrize("lengths,c_in,n_out,shifts,unpadded", [([300, 129], 64, 64, (-4, 0, 4), False),
                                                                 ([1000], 64, 64, (0,), False),
                                                                 ([200, 77, 513], 64, 131, (0,), False),
                                                                 ([257], 64, 64, (-512, -256, 0), False),
                                                                 ([140, 260], 2048, 64, (0,), True),
                                                                 ([90], 100, 24, (-1, 0, 1), False)])
def test_wgrad_tc_matches_fp64(lengths, c_in, n_out, shifts, unpadded):
    from computervision_codes_b200 import ops
    from computervision_codes_b200.layout import SeqLayout
    from oracle import tcn_oracle as O

    torch.manual_seed(c_in + n_out + len(shifts))
    lay = SeqLayout.get(lengths, DEV)
    ntaps = len(shifts)
    xs = [torch.randn(T, c_in) for T in lengths]
    gs = [torch.randn(T, n_out) for T in lengths]
    ldg = (n_out + 3) // 4 * 4
    g_rows = torch.zeros(lay.rows, ldg)
    x_rows = torch.zeros(lay.rows, c_in)
    for s, T in enumerate(lengths):
        g_rows[lay.starts[s]:lay.starts[s] + T, :n_out] = gs[s]
        x_rows[lay.starts[s]:lay.starts[s] + T] = xs[s]
    x_dev = torch.cat(xs).to(DEV) if unpadded else x_rows.to(DEV)
    g_dev = g_rows.to(DEV)
    dw = torch.zeros(n_out, c_in, ntaps, device=DEV)
    db = torch.zeros(n_out, device=DEV)
    p = 0.5 if ntaps == 1 and not unpadded else 0.0
    ops.wgrad_tc(g_dev, x_dev, lay, n_out, c_in, shifts, dw, db, x_unpadded=unpadded, g_drop_p=p, seed=11, stream_id=4)
    keep = ops.dropout_keep_mask(lay.rows, ldg, 0.5, 11, 4, DEV).cpu().double() * 2.0 if p > 0 else None
    dw_ref = torch.zeros(n_out, c_in, ntaps, dtype=torch.float64)
    db_ref = torch.zeros(n_out, dtype=torch.float64)
    for s, T in enumerate(lengths):
        gd = gs[s].double()
        if keep is not None:
            gd = gd * keep[lay.starts[s]:lay.starts[s] + T, :n_out]
        gb, xb = gd.t().unsqueeze(0), xs[s].double().t().unsqueeze(0)
        for k, sh in enumerate(shifts):
            dw_ref[:, :, k] += torch.einsum("bot,bct->oc", gb, O.shift_time(xb, sh))
        db_ref += gd.sum(0)
    assert _maxabs(dw, dw_ref) <= 2e-5 * max(1.0, float(dw_ref.abs().max())), (lengths, c_in, n_out, shifts)
    assert _maxabs(db, db_ref) <= 2e-5 * max(1.0, float(db_ref.abs().max()))
    # accumulates: a second call doubles the result
    ops.wgrad_tc(g_dev, x_dev, lay, n_out, c_in, shifts, dw, db, x_unpadded=unpadded, g_drop_p=p, seed=11, stream_id=4)
    assert _maxabs(dw, 2 * dw_ref) <= 4e-5 * max(1.0, float(dw_ref.abs().max()))


@pytest.mark.parametrize("lengths,shifts,p", [([300, 129, 1], (-4, 0, 4), 0.5), ([2250] * 8, (-64, -32, 0), 0.0),
                                              ([5000, 77], (-1024, 0, 1024), 0.3)])
def test_wgrad_tc_layer_pair_matches_two_launches(lengths, shifts, p):
    """tcn_wgrad_tc_pair (both weight gradients of a residual layer in one launch, the executor's per-layer launch)
    against two tcn_wgrad_tc launches on the same operands (same kernel body: equal up to the order of the fp32 atomics)."""
    from computervision_codes_b200 import ops
    from computervision_codes_b200.layout import SeqLayout

    torch.manual_seed(sum(lengths) % 1000)
    C = 64
    lay = SeqLayout.get(lengths, DEV)
    bufs = [torch.zeros(lay.rows, C, device=DEV) for _ in range(4)]
    for s, T in enumerate(lengths):
        for b in bufs:
            b[lay.starts[s]:lay.starts[s] + T] = torch.randn(T, C, device=DEV)
    gu, x, gy, h = bufs
    z = lambda *shape: torch.zeros(*shape, device=DEV)
    gw1, gb1, gw2, gb2 = z(C, C, 3), z(C), z(C, C, 1), z(C)
    ops.wgrad_tc_layer_pair(gu, x, gy, h, lay, shifts, gw1, gb1, gw2, gb2, drop_p=p, seed=5, stream_id=9)
    rw1, rb1, rw2, rb2 = z(C, C, 3), z(C), z(C, C, 1), z(C)
    ops.wgrad_tc(gu, x, lay, C, C, shifts, rw1, rb1)
    ops.wgrad_tc(gy, h, lay, C, C, (0,), rw2, rb2, g_drop_p=p, seed=5, stream_id=9)
    for got, ref in ((gw1, rw1), (gb1, rb1), (gw2, rw2), (gb2, rb2)):
        assert _maxabs(got, ref) <= 2e-5 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("lengths,shifts,p,use_masks", [([300, 129, 1], (-4, 0, 4), 0.5, True),
                                                        ([2250] * 8, (-64, -32, 0), 0.0, False),
                                                        ([5000, 77], (-1024, 0, 1024), 0.3, False),
                                                        ([1800], (-512, 0, 512), 0.5, True),
                                                        ([17, 16, 15, 128, 129], (-2, -1, 0), 0.5, True),
                                                        ([40000], (-8, 0, 8), 0.5, True)])
def test_wgrad_layer_matches_fp64_and_is_deterministic(lengths, shifts, p, use_masks):
    """tcn_wgrad_layer (csrc/wgrad_layer.cu: the four weight-gradient products of a residual layer from one pass, slab
    reduction in fixed order) against fp64 -- SURVEY 8a closed forms gW1[:, :, k] = sum_t gu[t] x[t + s_k]^T,
    gW2 = sum_t gv[t] h[t]^T -- with the dropout keep bits either handed over as the forward kernel's bit words or
    regenerated from the key; two runs must agree bit for bit."""
    from computervision_codes_b200 import ops
    from computervision_codes_b200.layout import SeqLayout
    from oracle import tcn_oracle as O

    torch.manual_seed(sum(lengths) % 1000 + len(lengths))
    C = 64
    lay = SeqLayout.get(lengths, DEV)
    host = [[torch.randn(T, C) for T in lengths] for _ in range(4)]   # gu, x, gy, h per sequence
    bufs = [torch.full((lay.rows, C), float("nan")) for _ in range(4)]  # pad rows poisoned: the kernel must ignore them
    for b, hs in zip(bufs, host):
        for s, T in enumerate(lengths):
            b[lay.starts[s]:lay.starts[s] + T] = hs[s]
    gu, x, gy, h = [b.to(DEV) for b in bufs]
    keep = ops.dropout_keep_mask(lay.rows, C, p, 5, 9, DEV) if p > 0 else torch.ones(lay.rows, C, device=DEV, dtype=torch.uint8)
    masks = None
    if use_masks:
        w = (keep.view(lay.rows, 2, 32).long() << torch.arange(32, device=DEV)).sum(-1)   # bit c of word c // 32
        w = torch.where(w >= 2 ** 31, w - 2 ** 32, w).to(torch.int32)
        masks = torch.zeros(lay.rows, 4, device=DEV, dtype=torch.int32)
        masks[:, 2:] = w
    z = lambda *shape: torch.zeros(*shape, device=DEV)
    outs = []
    for _ in range(2):
        gw1, gb1, gw2, gb2 = z(C, C, 3), z(C), z(C, C, 1), z(C)
        ops.layer_wgrad(gu, x, gy, h, lay, shifts, gw1, gb1, gw2, gb2, drop_p=p, seed=5, stream_id=9, masks=masks)
        outs.append((gw1, gb1, gw2, gb2))
    torch.cuda.synchronize()
    for a, b in zip(*outs):
        assert torch.equal(a, b), "the slab reduction must be bit-identical from run to run"
    kd = keep.cpu().double() / (1.0 - p)
    rw1, rb1 = torch.zeros(C, C, 3, dtype=torch.float64), torch.zeros(C, dtype=torch.float64)
    rw2, rb2 = torch.zeros(C, C, dtype=torch.float64), torch.zeros(C, dtype=torch.float64)
    for s, T in enumerate(lengths):
        gud, xd, gyd, hd = (host[i][s].double() for i in range(4))
        gvd = gyd * kd[lay.starts[s]:lay.starts[s] + T]
        xb = xd.t().unsqueeze(0)
        for k, sh in enumerate(shifts):
            rw1[:, :, k] += torch.einsum("to,ct->oc", gud, O.shift_time(xb, sh)[0])
        rb1 += gud.sum(0)
        rw2 += gvd.t() @ hd
        rb2 += gvd.sum(0)
    gw1, gb1, gw2, gb2 = outs[0]
    for got, ref in ((gw1, rw1), (gb1, rb1), (gw2.view(C, C), rw2), (gb2, rb2)):
        assert torch.isfinite(got).all()
        assert _maxabs(got, ref) <= 2e-5 * max(1.0, float(ref.abs().max())), (lengths, shifts, p)
    # accumulates into dW like the other weight-gradient entry points
    ops.layer_wgrad(gu, x, gy, h, lay, shifts, gw1, gb1, gw2, gb2, drop_p=p, seed=5, stream_id=9, masks=masks)
    assert _maxabs(gw1, 2 * rw1) <= 4e-5 * max(1.0, float(rw1.abs().max()))
