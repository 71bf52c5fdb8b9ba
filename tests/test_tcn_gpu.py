"""Parity of the CUDA path (through the C ABI) against the oracle and the golden fixtures.  GPU only.

Bars (BASELINE.json north_star): per-frame logits within 1e-3 max-abs at fp32 accumulation, loss
within 1e-4 relative, argmax identical.  The 3xTF32 kernels sit far inside: activations are held
to 2e-5 here so that a regression shows up long before the contractual bar.
"""
import os
import types

import numpy as np
import pytest
import torch

from oracle import tcn_oracle as O

from gradcheck import assert_grad_close

pytestmark = pytest.mark.gpu

DEV = "cuda"


@pytest.fixture(autouse=True, scope="module")
def _no_tf32_in_any_torch_reference():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def _t(a):
    return torch.from_numpy(np.asarray(a)).float()


def _maxabs(a, b):
    return float((a.detach().double().cpu() - b.detach().double().cpu()).abs().max())


def test_tapgemm_and_wgrad_against_fp64():
    from computervision_codes_b200 import ops
    from computervision_codes_b200.layout import SeqLayout

    torch.manual_seed(0)
    for (lengths, c_in, n_out, shifts) in [([70], 24, 16, (0,)), ([300, 17], 64, 131, (0,)),
                                           ([129, 128, 5], 40, 64, (-3, 0, 3)), ([200], 16, 16, (-256, -128, 0)),
                                           ([50, 260], 2048, 64, (0,)), ([140], 100, 72, (-1, 0, 1))]:
        lay = SeqLayout(lengths, DEV)
        ntaps = len(shifts)
        xs = [torch.randn(T, c_in, dtype=torch.float64) for T in lengths]
        w = torch.randn(n_out, c_in, ntaps, dtype=torch.float64) / (c_in * ntaps) ** 0.5
        b = torch.randn(n_out, dtype=torch.float64)
        gs = [torch.randn(T, n_out, dtype=torch.float64) for T in lengths]
        x_rows = torch.zeros(lay.rows, c_in)
        g_rows = torch.zeros(lay.rows, (n_out + 3) // 4 * 4)
        for s, T in enumerate(lengths):
            x_rows[lay.starts[s]:lay.starts[s] + T] = xs[s].float()
            g_rows[lay.starts[s]:lay.starts[s] + T, :n_out] = gs[s].float()
        x_rows, g_rows = x_rows.to(DEV), g_rows.to(DEV)
        wf = ops.prep_weight(w.float().to(DEV))
        y = ops.tapgemm(x_rows, wf, lay, c_in, n_out, shifts, bias=b.float().to(DEV))
        dw = torch.zeros(n_out, c_in, ntaps, device=DEV)
        db = torch.zeros(n_out, device=DEV)
        ops.wgrad(g_rows, x_rows, lay, n_out, c_in, shifts, dw, db)
        wft = ops.prep_weight(w.float().to(DEV), transpose=True)
        gx = ops.tapgemm(g_rows, wft, lay, g_rows.shape[1], c_in, tuple(-s for s in shifts), ldy=c_in)
        dw_ref = torch.zeros_like(w)
        db_ref = torch.zeros_like(b)
        for s, T in enumerate(lengths):
            xb = xs[s].t().unsqueeze(0)
            ref = O.conv_taps(xb, w, b, shifts)[0].t()
            got = y[lay.starts[s]:lay.starts[s] + T, :n_out]
            scale = float(ref.abs().max())
            assert _maxabs(got, ref) <= 3e-6 * max(scale, 1.0) + 2e-6 * c_in ** 0.5, (lengths, c_in, n_out, shifts)
            gb_ = gs[s].t().unsqueeze(0)
            for k, sh in enumerate(shifts):
                dw_ref[:, :, k] += torch.einsum("bot,bct->oc", gb_, O.shift_time(xb, sh))
            db_ref += gs[s].sum(0)
            gx_ref = sum(torch.einsum("oc,bot->bct", w[:, :, k], O.shift_time(gb_, -sh))
                         for k, sh in enumerate(shifts))[0].t()
            assert _maxabs(gx[lay.starts[s]:lay.starts[s] + T], gx_ref) <= 1e-5 * max(float(gx_ref.abs().max()), 1.0)
        assert _maxabs(dw, dw_ref) <= 2e-5 * max(float(dw_ref.abs().max()), 1.0)
        assert _maxabs(db, db_ref) <= 2e-5 * max(float(db_ref.abs().max()), 1.0)
        # pad rows / pad columns of the output stay zero
        if y.shape[1] > n_out:
            assert float(y[:, n_out:].abs().max()) == 0.0


def test_layers_against_golden(golden_dir):
    from computervision_codes_b200.tcn import DilatedResidualCausalLayer, DilatedResidualLayer

    z = _load(golden_dir, "tcn_layers.npz")
    for i in range(int(z["num_cases"])):
        tag = f"c{i}."
        causal = str(z[tag + "kind"]) == "causal"
        d = int(z[tag + "dilation"])
        cls = DilatedResidualCausalLayer if causal else DilatedResidualLayer
        m = cls(d, 16, 16).to(DEV).eval()
        m.load_state_dict({k[len(tag) + 3:]: _t(z[k]) for k in z.files if k.startswith(tag + "sd.")})
        x = _t(z[tag + "x"]).to(DEV).requires_grad_(True)
        y = m(x)
        assert y.shape == x.shape
        assert _maxabs(y, _t(z[tag + "y"])) <= 2e-5, (tag, _maxabs(y, _t(z[tag + "y"])))
        y.backward(_t(z[tag + "gy"]).to(DEV))
        assert _maxabs(x.grad, _t(z[tag + "gx"])) <= 2e-5
        for k, v in m.named_parameters():
            ref = _t(z[tag + "grad." + k])
            assert _maxabs(v.grad, ref) <= 2e-5 * max(1.0, float(ref.abs().max())), (tag, k)


@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("C,T,B,d", [(64, 1800, 1, 512), (64, 999, 2, 1), (64, 10, 3, 1024), (32, 300, 2, 16),
                                     (512, 200, 1, 4)])
def test_layer_against_oracle_fp64(causal, C, T, B, d):
    from computervision_codes_b200.tcn import DilatedResidualCausalLayer, DilatedResidualLayer

    torch.manual_seed(C + T + d)
    cls = DilatedResidualCausalLayer if causal else DilatedResidualLayer
    m = cls(d, C, C).to(DEV).eval()
    x = torch.randn(B, C, T)
    gy = torch.randn(B, C, T)
    xg = x.to(DEV).requires_grad_(True)
    y = m(xg)
    y.backward(gy.to(DEV))
    ws = [m.conv_dilated.weight, m.conv_dilated.bias, m.conv_1x1.weight, m.conv_1x1.bias]
    wd = [w.detach().double().cpu() for w in ws]
    yr = O.dilated_residual_layer(x.double(), *wd, d, causal=causal)
    refs = O.layer_backward_closed_form(x.double(), gy.double(), *wd, d, causal=causal)
    assert _maxabs(y, yr) <= 2e-5
    assert _maxabs(xg.grad, refs[0]) <= 2e-5
    for w, r in zip(ws, refs[1:]):
        assert _maxabs(w.grad, r.reshape(w.shape)) <= 3e-5 * max(1.0, float(r.abs().max())), (C, T, d)


def test_layer_train_mode_dropout_matches_oracle_with_same_mask():
    from computervision_codes_b200 import ops
    from computervision_codes_b200.layout import SeqLayout

    torch.manual_seed(5)
    B, C, T, d, p = 2, 64, 300, 4, 0.5
    lay = SeqLayout.uniform(B, T, DEV)
    x = torch.randn(B, C, T)
    gy = torch.randn(B, C, T)
    w1, b1 = torch.randn(C, C, 3) / (3 * C) ** 0.5, torch.randn(C) * 0.1
    w2, b2 = torch.randn(C, C, 1) / C ** 0.5, torch.randn(C) * 0.1
    prm = [t.to(DEV).requires_grad_(True) for t in (w1, b1, w2, b2)]
    xr = lay.pad_bct(x.to(DEV)).requires_grad_(True)
    seed, sid = 1234, 7
    y = ops.dilated_residual(xr, *prm, lay, d, False, p, seed, sid)
    y.backward(lay.pad_bct(gy.to(DEV)))
    keep_rows = ops.dropout_keep_mask(lay.rows, C, p, seed, sid, DEV)
    keep = lay.as_bct(keep_rows.float(), C).double().cpu()
    frac = float(keep.mean())
    assert 0.47 < frac < 0.53
    wd = [t.double() for t in (w1, b1, w2, b2)]
    yr = O.dilated_residual_layer(x.double(), *wd, d, keep=keep, p=p)
    refs = O.layer_backward_closed_form(x.double(), gy.double(), *wd, d, keep=keep, p=p)
    assert _maxabs(lay.as_bct(y, C), yr) <= 3e-5
    assert _maxabs(lay.as_bct(xr.grad, C), refs[0]) <= 3e-5
    for w, r in zip(prm, refs[1:]):
        assert _maxabs(w.grad, r.reshape(w.shape)) <= 3e-5 * max(1.0, float(r.abs().max()))
    # a different stream id gives a different mask
    other = ops.dropout_keep_mask(lay.rows, C, p, seed, sid + 1, DEV)
    assert float((other != keep_rows).float().mean()) > 0.4


@pytest.mark.parametrize("lengths,shifts,p", [([300, 129, 1, 128], (-4, 0, 4), 0.0), ([999, 7], (-2, -1, 0), 0.5),
                                              ([1800], (-512, 0, 512), 0.5), ([130, 260, 20000], (-1024, -512, 0), 0.3),
                                              ([40000], (-1, 0, 1), 0.5)])
def test_fused_layer_backward_matches_unfused_kernels_and_fp64(lengths, shifts, p):
    """tcn_layer_bwd_tc (one launch: gu recomputed per tap in tensor memory, ReLU / dropout masks as bit words saved by
    tcn_layer_fwd_tc) against (1) the two tap-GEMM launches it replaces and (2) the fp64 closed form of the layer's
    backward pass (oracle).  Ragged batches, taps leaving the sequence, more 128-frame tiles than SMs."""
    from computervision_codes_b200 import ops
    from computervision_codes_b200.layout import SeqLayout

    torch.manual_seed(len(lengths) + shifts[2])
    C = 64
    lay = SeqLayout(lengths, DEV)
    w1 = (torch.randn(C, C, 3) / (3 * C) ** 0.5).to(DEV)
    w2 = (torch.randn(C, C, 1) / C ** 0.5).to(DEV)
    b1, b2 = (torch.randn(C) * 0.1).to(DEV), (torch.randn(C) * 0.1).to(DEV)
    x = torch.zeros(lay.rows, C, device=DEV)
    gy = torch.zeros(lay.rows, C, device=DEV)
    for s, T in enumerate(lengths):
        x[lay.starts[s]:lay.starts[s] + T] = torch.randn(T, C, device=DEV)
        gy[lay.starts[s]:lay.starts[s] + T] = torch.randn(T, C, device=DEV)
    seed, sid = 77, 3
    y, h, masks = ops.layer_fwd_tc(x, w1, w2, b1, b2, lay, shifts, True, p, seed, sid, save_masks=True)
    y0, h0 = ops.layer_fwd_tc(x, w1, w2, b1, b2, lay, shifts, True, p, seed, sid)
    assert torch.equal(y, y0) and torch.equal(h, h0)          # saving the masks does not change the forward result
    # the bit words are the masks the unfused kernels recompute
    keep_rows = ops.dropout_keep_mask(lay.rows, C, p, seed, sid, DEV) if p > 0 else torch.ones(lay.rows, C, device=DEV)
    bits = torch.arange(32, device=DEV)
    unpack = lambda wds: ((wds.long().unsqueeze(-1) >> bits) & 1).reshape(wds.shape[0], -1)
    for s, T in enumerate(lengths):
        sl = slice(lay.starts[s], lay.starts[s] + T)
        assert torch.equal(unpack(masks[sl, 0:2]) != 0, h[sl] > 0)
        assert torch.equal(unpack(masks[sl, 2:4]) != 0, keep_rows[sl] != 0)
    gu, gx = ops.layer_bwd_tc(gy, masks, w1, w2, lay, shifts, p)
    gu_ref = ops.tapgemm(gy, ops.prep_weight(w2, transpose=True), lay, C, C, (0,), relu_mask=h, in_drop_p=p, seed=seed,
                         stream_id=sid)
    gx_ref = ops.tapgemm(gu_ref, ops.prep_weight(w1, transpose=True), lay, C, C, tuple(-s for s in shifts), residual=gy)
    torch.cuda.synchronize()
    for s, T in enumerate(lengths):
        sl = slice(lay.starts[s], lay.starts[s] + T)
        assert _maxabs(gu[sl], gu_ref[sl]) <= 2e-5, (s, T)
        assert _maxabs(gx[sl], gx_ref[sl]) <= 2e-5, (s, T)
        # fp64: gv = keep * gy / (1 - p); gu = (gv W2) * [h > 0]; gx[t] = gy[t] + sum_k W1_k^T gu[t - s_k]
        gyd, kd = gy[sl].double(), keep_rows[sl].double()
        gud = ((gyd * kd / (1.0 - p)) @ w2[:, :, 0].double()) * (h[sl] > 0).double()
        gxd = gyd.clone()
        for k, sh in enumerate(shifts):
            contrib = gud @ w1[:, :, k].double()              # lands on frame t + s_k
            if sh >= 0:
                if sh < T:
                    gxd[sh:] += contrib[:T - sh]
            elif -sh < T:
                gxd[:T + sh] += contrib[-sh:]
        assert _maxabs(gu[sl], gud) <= 2e-5 and _maxabs(gx[sl], gxd) <= 3e-5, (s, T)


def test_stage_against_golden(golden_dir):
    from computervision_codes_b200.tcn import BaseCausalTCN, Refinement

    z = _load(golden_dir, "tcn_stage.npz")
    args = types.SimpleNamespace(output=False, hier=False)
    for tag, causal in (("acausal", False), ("causal", True)):
        pg = BaseCausalTCN(6, 32, 40, 7, causal=causal).to(DEV).eval()
        rf = Refinement(args, 6, 32, 7, 7, None, causal=causal).to(DEV).eval()
        pg.load_state_dict({k[len(tag) + 7:]: _t(z[k]) for k in z.files if k.startswith(tag + ".sd.PG.")})
        rf.load_state_dict({k[len(tag) + 9:]: _t(z[k]) for k in z.files if k.startswith(tag + ".sd.Rs.0.")})
        x = _t(z[tag + ".x"]).to(DEV)
        f0, l0 = pg(x.permute(0, 2, 1))
        f1, l1 = rf(f0)
        for got, name in ((f0, "f0"), (l0, "l0"), (f1, "f1"), (l1, "l1")):
            ref = _t(z[f"{tag}.{name}"])
            assert got.shape == ref.shape
            assert _maxabs(got, ref) <= 5e-5, (tag, name, _maxabs(got, ref))
        assert torch.equal(l1.argmax(1).cpu(), _t(z[tag + ".l1"]).argmax(1))  # phase predictions identical


@pytest.mark.parametrize("fixture", ["tcn_videonas.npz", "tcn_videonas_c64.npz"])
def test_videonas_against_golden(golden_dir, fixture):
    """tcn_videonas.npz has 16 channels (mma.sync kernels); tcn_videonas_c64.npz (C = 64, D = 96, T = 300) pins the
    tcgen05 fused-layer / weight-gradient kernels against the reference's own outputs in one hop."""
    from computervision_codes_b200.tcn import VideoNas

    z = _load(golden_dir, fixture)
    nl_pg, nl_r, n_r, C, D, K, T, B = [int(v) for v in z["cfg"]]
    args = types.SimpleNamespace(fpn=True, output=False, feature=False, trans=False, mask=False, hier=False)
    m = VideoNas(args, nl_pg, nl_r, n_r, C, D, K).to(DEV).eval()
    m.load_state_dict({k[3:]: _t(z[k]) for k in z.files if k.startswith("sd.")})
    x = _t(z["x"]).to(DEV)
    outs = m(x, False)  # goes through the native executor (two C calls)
    assert m._executor is not None
    for name, lst in zip(("ivt", "i", "v", "t", "f"), (outs[0], outs[1], outs[2], outs[3], outs[4])):
        assert len(lst) == 4
        for lvl, t in enumerate(lst):
            ref = _t(z[f"out_{name}.{lvl}"])
            assert t.shape == ref.shape
            assert _maxabs(t, ref) <= 1e-4, (name, lvl, _maxabs(t, ref))
            if name != "f":
                assert torch.equal(t.argmax(1).cpu(), ref.argmax(1))
    # the reference's own loss glue (torch BCEWithLogitsLoss on our outputs), then backward
    labels = [_t(z["label_" + n]).to(DEV) for n in "ivtq"]
    bce = torch.nn.BCEWithLogitsLoss()
    terms = [sum(bce(pd[0].transpose(0, 1), y) for pd in lst)
             for lst, y in zip((outs[1], outs[2], outs[3], outs[0]), labels)]
    loss = 0.1 * (terms[0] + terms[1] + terms[2]) + terms[3]
    assert abs(float(loss) - float(z["loss"])) <= 1e-4 * abs(float(z["loss"]))
    loss.backward()
    nograd = set(str(s) for s in z["nograd"])
    for k, v in m.named_parameters():
        if "grad." + k in z.files:
            ref = _t(z["grad." + k])
            assert v.grad is not None, k
            assert_grad_close(v.grad, ref, k, strict=2e-5)
        else:
            assert k in nograd and v.grad is None, k  # stays None so SGD weight decay skips it, as in the reference

    # fused loss path == the reference composition (tenco and TERL variants)
    from computervision_codes_b200 import losses
    from computervision_codes_b200.layout import SeqLayout

    lay = SeqLayout.uniform(B, T, DEV)
    m.zero_grad(set_to_none=True)
    f_rows, logit_rows = m.forward_packed(x.contiguous(), lay)
    lab = losses.pack_labels(*[l.long() for l in labels])
    total, li, lv, lt, livt = losses.tenco_loss(logit_rows, lab, lay)
    assert abs(float(total) - float(z["loss"])) <= 1e-4 * abs(float(z["loss"]))
    np.testing.assert_allclose([float(li), float(lv), float(lt), float(livt)], z["loss_terms"], rtol=1e-4)
    total.backward()
    for k, v in m.named_parameters():
        if "grad." + k in z.files:
            ref = _t(z["grad." + k])
            assert_grad_close(v.grad, ref, k, strict=2e-5)
    with torch.no_grad():
        _, li, lv, lt, livt = losses.tenco_loss(logit_rows, lab, lay, terl_pos_weight=True)
    np.testing.assert_allclose([float(li), float(lv), float(lt), float(livt)], z["terl_loss_terms"], rtol=1e-4)


def test_kd_losses_against_golden(golden_dir):
    from computervision_codes_b200 import losses

    z = _load(golden_dir, "kd_loss.npz")
    kl = losses.DistillKL(4.0)
    for i in range(int(z["num_kl"])):
        ys = _t(z[f"kl{i}.ys"]).to(DEV).requires_grad_(True)
        ytl = _t(z[f"kl{i}.yt_logits"]).to(DEV)
        ref = float(z[f"kl{i}.loss"])
        for fused in (False, True):
            ys.grad = None
            loss = kl(ys, ytl if fused else torch.sigmoid(ytl), teacher_is_logits=fused)
            assert abs(float(loss) - ref) <= 1e-4 * abs(ref) + 1e-8
            loss.backward()
            assert _maxabs(ys.grad, _t(z[f"kl{i}.gys"])) <= 1e-6
    logits = [_t(z[f"comp.logits{k}"]).to(DEV).requires_grad_(True) for k in range(4)]
    labels = [_t(z[f"comp.labels{k}"]).to(DEV) for k in range(4)]
    teach = [_t(z[f"comp.teach{k}"]).to(DEV) for k in range(3)]
    feats = [_t(z[f"comp.feat{k}"]).to(DEV).requires_grad_(True) for k in range(3)]
    tfeats = [_t(z[f"comp.tfeat{k}"]).to(DEV) for k in range(3)]
    crit = losses.MultiTeacherKDLoss(temp=4.0, rates=(1.0, 1.0, 1.0))
    loss, hard, soft, kd = crit(logits, labels, teach, feats, tfeats)
    np.testing.assert_allclose([float(loss), float(hard), float(soft), float(kd)], z["comp.loss"], rtol=1e-4)
    loss.backward()
    for k in range(4):
        assert _maxabs(logits[k].grad, _t(z[f"comp.glogits{k}"])) <= 1e-6
    for k in range(3):
        assert _maxabs(feats[k].grad, _t(z[f"comp.gfeat{k}"])) <= 1e-6
    # phase head (parity unpinned: textbook definition)
    torch.manual_seed(3)
    x = torch.randn(1800, 7, device=DEV, requires_grad=True)
    y = torch.randint(0, 7, (1800,), device=DEV)
    got = losses.phase_cross_entropy(x, y)
    ref = O.phase_ce(x.detach().double().cpu(), y.cpu())
    assert abs(float(got) - float(ref)) <= 1e-5 * float(ref)
    got.backward()
    xr = x.detach().double().cpu().requires_grad_(True)
    O.phase_ce(xr, y.cpu()).backward()
    assert _maxabs(x.grad, xr.grad) <= 1e-7


def test_videonas_cfg_baseline_shape_vs_oracle():
    """BASELINE-shaped instantiation (C=64, D=2048, one 1,800-frame video): logits <= 1e-3, argmax identical."""
    from computervision_codes_b200.tcn import VideoNas

    args = types.SimpleNamespace(fpn=True, output=False, feature=False, trans=False, mask=False, hier=False)
    torch.manual_seed(0)
    m = VideoNas(args, 11, 10, 3, 64, 2048, 100).eval()
    x = torch.randn(1, 1800, 2048)
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    with torch.no_grad():
        ref = O.videonas_forward(x, sd)
    m = m.to(DEV)
    with torch.no_grad():
        outs = m(x.to(DEV), False)
    for a_list, r_list in zip(outs[:4], ref[:4]):
        for a, r in zip(a_list, r_list):
            assert _maxabs(a, r) <= 1e-3
            assert torch.equal(a.argmax(1).cpu(), r.argmax(1))
            flips = (a > 0).cpu() != (r > 0)   # sigmoid > 0.5 decisions: only ties (|logit| ~ 0) may differ
            assert float(r[flips].abs().max()) < 1e-4 if bool(flips.any()) else True


def test_drop_in_training_loop_matches_cpu_port_for_two_sgd_steps():
    """The reference's train_loop pattern (run.py:182-235: model(img, True) -> 16 x BCEWithLogitsLoss -> grads = None ->
    backward -> torch.optim.SGD(weight_decay)) on the drop-in module, against the CPU port doing the same
    (dropout off in both so that the trajectories are comparable)."""
    from computervision_codes_b200.tcn import VideoNas
    from oracle import torch_port as P

    args = types.SimpleNamespace(fpn=True, output=False, feature=False, trans=False, mask=False, hier=False)
    torch.manual_seed(5)
    m = VideoNas(args, 4, 3, 3, 64, 96, 100)
    ref = {k: v.detach().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    m = m.to(DEV).eval()
    g = torch.Generator().manual_seed(6)
    videos = [(torch.randn(1, T, 96, generator=g), [(torch.rand(T, k, generator=g) < 0.1).float() for k in (6, 10, 15, 100)])
              for T in (150, 333)]
    opt = torch.optim.SGD(m.parameters(), lr=0.05, weight_decay=1e-5)
    opt_ref = torch.optim.SGD(list(ref.values()), lr=0.05, weight_decay=1e-5)
    bce = torch.nn.BCEWithLogitsLoss()
    for x, labels in videos:
        outs = m(x.to(DEV), True)
        terms = [sum(bce(pd[0].transpose(0, 1), y.to(DEV)) for pd in lst)
                 for lst, y in zip((outs[1], outs[2], outs[3], outs[0]), labels)]
        loss = 0.1 * (terms[0] + terms[1] + terms[2]) + terms[3]
        for p in m.parameters():
            p.grad = None
        loss.backward()
        opt.step()
        for p in ref.values():
            p.grad = None
        loss_ref = P.train_step_loss(x, ref, tuple(labels), train=False)
        loss_ref.backward()
        opt_ref.step()
        assert abs(float(loss) - float(loss_ref)) <= 1e-4 * abs(float(loss_ref))
    for k, v in m.state_dict().items():
        assert _maxabs(v, ref[k]) <= 2e-5, (k, _maxabs(v, ref[k]))


@pytest.mark.parametrize("tag", ["small", "wide"])
def test_multi_teacher_feature_attention_matches_reference_golden(golden_dir, tag):
    """Row f3: the drop-in block loads the reference's weights by name and reproduces its outputs, the KD term of
    Spatial_cnn/run.py:187-191 and every gradient (fixtures: oracle/gen_golden_kdattn.py)."""
    import numpy as np

    from computervision_codes_b200.losses import MultiTeacherFeatureAttention, mse_loss

    z = np.load(os.path.join(golden_dir, "kd_attn.npz"))
    s = torch.from_numpy(z[f"{tag}.s"]).to(DEV).requires_grad_(True)
    teachers = [torch.from_numpy(z[f"{tag}.teacher_{n}"]).to(DEV) for n in "ivt"]
    blk = MultiTeacherFeatureAttention(s.shape[1], teachers[0].shape[1])
    sd = {k[len(tag) + 4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith(f"{tag}.sd.")}
    assert set(sd) == set(blk.state_dict())
    blk.load_state_dict(sd)
    blk = blk.to(DEV)
    outs = blk(s, *teachers)
    for n, y in zip("ivt", outs):
        ref = z[f"{tag}.stus_f{n}"]
        assert y.shape == ref.shape
        assert np.abs(y.detach().cpu().numpy() - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max())
    kd = sum(mse_loss(a, b) for a, b in zip(outs, teachers)) / 3
    assert abs(float(kd.detach()) - float(z[f"{tag}.kd_loss"])) <= 1e-5 * float(z[f"{tag}.kd_loss"])
    kd.backward()
    ref = z[f"{tag}.grad.s"]
    assert np.abs(s.grad.cpu().numpy() - ref).max() <= 1e-4 * np.abs(ref).max() + 1e-8
    for name, prm in blk.named_parameters():
        ref = z[f"{tag}.grad.{name}"]
        got = prm.grad.cpu().numpy()
        assert np.abs(got - ref).max() <= 1e-4 * np.abs(ref).max() + 1e-8, name


def test_video_ap_matches_sklearn_average_precision():
    """f4: device per-video AP == sklearn.metrics.average_precision_score (the routine ivtmetrics calls) per class,
    ties from a saturating fp32 sigmoid included; classes without positives are NaN and drop out of the means."""
    import warnings

    from sklearn.metrics import average_precision_score

    from computervision_codes_b200.evaluation import VideoAP, video_ap

    g = torch.Generator().manual_seed(5)
    ap_meter = VideoAP(20)
    per_video = []
    for T in (1, 37, 5000):
        logits = torch.randn(T, 20, generator=g) * 6
        logits[:, 3] = torch.round(logits[:, 3])            # many exact ties
        logits[:, 4] = logits[:, 4].abs() + 18.0            # sigmoid saturates to 1.0f: everything ties
        y = (torch.rand(T, 20, generator=g) < 0.2).to(torch.uint8)
        y[:, 7] = 0                                         # class absent from this video
        got = video_ap(y.to(DEV), logits.to(DEV)).cpu().numpy()
        s = torch.sigmoid(logits).numpy()
        want = np.full(20, np.nan)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for c in range(20):
                if y[:, c].any():
                    want[c] = average_precision_score(y[:, c].numpy(), s[:, c])
        assert np.array_equal(np.isnan(got), np.isnan(want))
        assert np.nanmax(np.abs(got - want)) <= 2e-6, (T, np.nanmax(np.abs(got - want)))
        per_video.append(want)
        ap_meter.update(y.to(DEV), logits.to(DEV))
        ap_meter.video_end()
    res = ap_meter.compute_video_AP()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        classwise = np.nanmean(np.stack(per_video), axis=0)
    assert np.allclose(res["AP"], classwise, atol=2e-6, equal_nan=True)
    assert abs(res["mAP"] - np.nanmean(classwise)) <= 2e-6


def test_refinement_with_args_output_against_golden(golden_dir):
    """Refinement(args.output=True) (network.py:150-151): its own conv_1x1 (K -> C) on a K-channel input; outputs and
    every parameter gradient against the reference's run (K = 10: the row pitch is not a multiple of 16 bytes, so the
    projection takes the mma.sync tap GEMM)."""
    from computervision_codes_b200.tcn import Refinement

    z = _load(golden_dir, "tcn_refine_output.npz")
    L, C, K, T, B = [int(v) for v in z["cfg"]]
    rf = Refinement(types.SimpleNamespace(output=True, hier=False), L, C, K, K, None).to(DEV).eval()
    rf.load_state_dict({k[3:]: _t(z[k]) for k in z.files if k.startswith("sd.")})
    f, lg = rf(_t(z["x"]).to(DEV))
    assert _maxabs(f, _t(z["f"])) <= 1e-4 and _maxabs(lg, _t(z["logits"])) <= 1e-4
    ((f * _t(z["gf"]).to(DEV)).sum() + (lg * _t(z["gl"]).to(DEV)).sum()).backward()
    for k, v in rf.named_parameters():
        ref = _t(z["grad." + k])
        assert _maxabs(v.grad, ref) <= 3e-5 * max(1.0, float(ref.abs().max())), k


def test_args_hier_is_refused_loudly():
    """args.hier (AvgPool1d(7, 3) between stages, network.py:145,156-157) is not on the CUDA path: constructing a stage
    with it must raise instead of silently running a different network (no reference script enables it)."""
    from computervision_codes_b200.tcn import Refinement, VideoNas

    with pytest.raises(NotImplementedError):
        Refinement(types.SimpleNamespace(output=False, hier=True), 2, 16, 16, 7, None)
    with pytest.raises(NotImplementedError):
        VideoNas(types.SimpleNamespace(fpn=True, output=False, feature=False, trans=False, mask=False, hier=True),
                 2, 2, 3, 16, 24, 100)


def test_second_forward_before_backward_is_refused():
    """The native executor keeps one set of saved activations: backward of a forward that a later forward has overwritten
    must raise, not return gradients of the wrong batch."""
    from computervision_codes_b200.tcn import VideoNas

    args = types.SimpleNamespace(fpn=True, output=False, feature=False, trans=False, mask=False, hier=False)
    torch.manual_seed(0)
    m = VideoNas(args, 3, 2, 3, 64, 32, 100).to(DEV).train()
    x1, x2 = torch.randn(1, 200, 32, device=DEV), torch.randn(1, 200, 32, device=DEV)
    o1 = m(x1, False)
    o2 = m(x2, False)
    with pytest.raises(RuntimeError, match="one outstanding forward"):
        o1[0][0].sum().backward()
    o2[0][0].sum().backward()   # the latest forward is fine
    assert m.PG.conv_1x1.weight.grad is not None


def test_drop_in_forward_with_mask_runs_the_device_side_generator():
    """forward(x, ismask=True) with args.mask in train mode: the 25 % input mask is drawn inside the projection kernel
    (no host randperm / H2D); two calls draw different masks, the gradient is finite, eval mode is unaffected."""
    from computervision_codes_b200.tcn import VideoNas

    args = types.SimpleNamespace(fpn=True, output=False, feature=False, trans=False, mask=True, hier=False)
    torch.manual_seed(0)
    m = VideoNas(args, 3, 2, 3, 64, 32, 100).to(DEV).train()
    x = torch.randn(1, 300, 32, device=DEV)
    a = m(x, True)[0][0].detach().clone()
    out = m(x, True)
    assert not torch.equal(a, out[0][0])
    out[0][0].sum().backward()
    assert bool(torch.isfinite(m.PG.conv_1x1.weight.grad).all())
    m.eval()
    with torch.no_grad():
        e1, e2 = m(x, False)[0][0].clone(), m(x, False)[0][0].clone()
    assert torch.equal(e1, e2)


@pytest.mark.parametrize("backlog", [0, 1, 3, 6])
def test_first_forward_behind_queued_work_equals_the_second(backlog):
    """The very first forward of a fresh module (executor created, buffers zero-filled) issued while earlier work is
    still draining on the stream must equal a repeat of it bit for bit.  Round 2 found eager programmatic dependent
    launches letting a layer read 1-40 frames of its input before the previous layer's last stores had landed -- visible
    only here, where the stale values are the zero fill (later forwards re-read identical values); the launchers now use
    programmatic launch on capturing streams only (csrc/runtime.cu: pdl_allowed_on)."""
    from computervision_codes_b200.tcn import VideoNas

    args = types.SimpleNamespace(fpn=True, output=False, feature=False, trans=False, mask=False, hier=False)
    torch.manual_seed(backlog)
    m = VideoNas(args, 11, 10, 3, 64, 2048, 100).to(DEV).eval()
    x = torch.randn(1, 1800, 2048, device=DEV)
    big = torch.randn(4096, 4096, device=DEV)
    torch.cuda.synchronize()
    for _ in range(backlog):  # ~1 ms each: the forward's ~50 launches are submitted while this drains
        big = (big @ big).clamp_(-1, 1)
    first = m(x, False)
    second = m(x, False)
    for a_list, b_list in zip(first[:5], second[:5]):
        for a, b in zip(a_list, b_list):
            assert torch.equal(a, b)
