"""Parity at the BASELINE.json sizes.  GPU only.

The fixtures under tests/golden/ (*_full, *_cfg3, *_t1800) were produced by the reference's own modules
(oracle/gen_golden_full.py): a full-size run cannot ship its weights, so the fixture carries the seeds, per-tensor
state_dict sums (the test first proves that this repo's mirrors draw the very same weights), output samples, the
loss and per-parameter gradient digests.  Where the unmodified reference modules travelled to the box
(baseline/_ref, written by oracle/vendor_ref.py) they are also run live on the GPU, TF32 off, as the checker.
"""
import os
import types
import warnings

import numpy as np
import pytest
import torch

from gradcheck import assert_grad_close

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(autouse=True, scope="module")
def _no_tf32_in_any_torch_reference():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.deterministic = True  # the live-reference legs: same cuDNN algorithms every run
    yield


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def _t(a):
    return torch.from_numpy(np.asarray(a)).float()


def _maxabs(a, b):
    return float((a.detach().double().cpu() - b.detach().double().cpu()).abs().max())


def _assert_argmax_consistent(got, ref, dim, tie=2e-3):
    """Identical predictions, except where the live reference's own top two logits are closer than `tie` (twice the
    logit tolerance): the reference on cuDNN is not run-to-run deterministic, so a near-tie may resolve either way."""
    ga, ra = got.argmax(dim), ref.argmax(dim)
    bad = ga != ra
    if bool(bad.any()):
        top2 = ref.detach().topk(2, dim=dim).values
        gap = (top2.select(dim, 0) - top2.select(dim, 1)).abs()
        assert float(gap[bad].max()) <= tie, f"{int(bad.sum())} predictions differ, widest reference gap {float(gap[bad].max()):.3e}"


def _check_digest(grad, ref_dig, idx, rel, name):
    """ref_dig = [sum, abs-sum, max-abs, <g, r_idx>] of the reference gradient (float64).  Bulk agreement at `rel` moves
    each sum by at most rel * abs-sum; a few ReLU flips (tests/gradcheck.py) add rows of ~1e-3 * max-abs, bounded here by
    2e-3 of the abs-sum / max-abs.  A structural bug changes these digests by tens of percent."""
    from oracle.gen_golden_full import digest

    got = digest(grad, idx)
    asum, gmax = max(float(ref_dig[1]), 1e-30), max(float(ref_dig[2]), 1e-30)
    assert abs(got[2] - ref_dig[2]) <= 2e-3 * gmax, (name, "max-abs", got[2], ref_dig[2])
    assert abs(got[1] - ref_dig[1]) <= (rel + 2e-3) * asum, (name, "abs-sum", got[1], ref_dig[1])
    n = grad.numel()
    tol = (rel + 2e-3) * asum / max(1.0, n ** 0.5 / 8.0) + 4.0 * rel * gmax * n ** 0.5   # signed sums: errors partly cancel
    for j, what in ((0, "sum"), (3, "projection")):
        assert abs(got[j] - ref_dig[j]) <= max(tol, 2e-3 * gmax * 256), (name, what, got[j], ref_dig[j], tol)


def _sd_sums(mods):
    from oracle.gen_golden_full import sd_sums

    return np.concatenate([sd_sums(m.state_dict().items()) for m in mods])


# ------------------------------------------------------------------------------------------------ cfg2, full size
def test_tcn_cfg2_full_size_forward_loss_backward_against_reference_golden(golden_dir):
    """VideoNas(fpn, 11/10/3, C = 64, D = 2048) on one 1,800-frame video: logits, argmax, loss and ALL gradients against
    the reference's own run (tests/golden/tcn_cfg2_full.npz)."""
    from computervision_codes_b200 import losses
    from computervision_codes_b200.layout import SeqLayout
    from computervision_codes_b200.tcn import VideoNas

    z = _load(golden_dir, "tcn_cfg2_full.npz")
    nl_pg, nl_r, n_r, C, D, K, T, B, seed = [int(v) for v in z["cfg"]]
    args = types.SimpleNamespace(fpn=True, output=False, feature=False, trans=False, mask=False, hier=False)
    torch.manual_seed(seed)
    m = VideoNas(args, nl_pg, nl_r, n_r, C, D, K).eval()
    x = torch.randn(B, T, D)
    np.testing.assert_allclose(_sd_sums([m]), z["sd_sums"], rtol=1e-12, atol=1e-12,
                               err_msg="the mirror does not draw the reference's initial weights under the same seed")
    g = torch.Generator().manual_seed(1)
    labels = [(torch.rand(T, k, generator=g) < 0.05).long() for k in (6, 10, 15, K)]
    m = m.to(DEV)
    outs = m(x.to(DEV), False)   # native executor
    for name, lst in zip(("ivt", "i", "v", "t"), outs[:4]):
        for lvl, t in enumerate(lst):
            ref = _t(z[f"out_{name}.{lvl}.sample"])
            assert _maxabs(t[:, :, ::8], ref) <= 1e-3, (name, lvl)
            assert torch.equal(t.argmax(1).cpu().to(torch.uint8), torch.from_numpy(z[f"out_{name}.{lvl}.argmax"]))
    for lvl, t in enumerate(outs[4]):
        assert _maxabs(t[:, ::4, ::8], _t(z[f"out_f.{lvl}.sample"])) <= 1e-3
    bce = torch.nn.BCEWithLogitsLoss()
    terms = [sum(bce(pd[0].transpose(0, 1), y.float().to(DEV)) for pd in lst)
             for lst, y in zip((outs[1], outs[2], outs[3], outs[0]), labels)]
    loss = 0.1 * (terms[0] + terms[1] + terms[2]) + terms[3]
    assert abs(float(loss) - float(z["loss"])) <= 1e-4 * abs(float(z["loss"]))
    loss.backward()
    names = [str(s) for s in z["grad_names"]]
    params = dict(m.named_parameters())
    for i, k in enumerate(names):
        assert params[k].grad is not None, k
        _check_digest(params[k].grad, z[f"gdig.{i}"], i, 3e-5, k)
        if f"grad.{k}" in z.files:
            ref = _t(z[f"grad.{k}"])
            assert_grad_close(params[k].grad, ref, k)
    for k in (str(s) for s in z["nograd"]):
        assert params[k].grad is None, k
    # the fused loss kernel on the same logits
    lay = SeqLayout.uniform(B, T, DEV)
    with torch.no_grad():
        _, logit_rows = m.forward_packed(x.to(DEV).contiguous(), lay)
        lab = losses.pack_labels(*[l.to(DEV) for l in labels])
        total, li, lv, lt, livt = losses.tenco_loss(logit_rows, lab, lay)
    assert abs(float(total) - float(z["loss"])) <= 1e-4 * abs(float(z["loss"]))
    np.testing.assert_allclose([float(li), float(lv), float(lt), float(livt)], z["loss_terms"], rtol=1e-4)


# ------------------------------------------------------------------------------------------------ cfg3, full size
def _mstct_modules(z):
    from computervision_codes_b200.mstct import Classifier, TemporalEncoder, Temporal_Mixer
    from oracle.gen_golden_full import perturb_1d

    in_dim, d1, d2, d3, d4, heads, ratio, nblk, emb, K, B, T, seed = [int(v) for v in z["cfg"]]
    torch.manual_seed(seed)
    enc = TemporalEncoder(in_dim, [d1, d2, d3, d4], heads, ratio, torch.nn.LayerNorm, nblk)
    mix = Temporal_Mixer([d1, d2, d3, d4], emb)
    cls = Classifier(emb, K)
    perturb_1d((enc, mix, cls))
    x = torch.randn(B, in_dim, T)
    lab = (torch.rand(B, T, K) < 0.05).float()
    return (enc, mix, cls), x, lab, (B, T, K)


def test_mstct_cfg3_full_size_forward_loss_backward_against_reference_golden(golden_dir):
    """BASELINE cfg3: TemporalEncoder(768 -> 256/384/576/864, head dims 32/48/72/108, hidden 2048..6912) + mixer +
    classifier on (31, 768, 256): outputs, argmax, loss and every parameter gradient against the reference's own run."""
    from computervision_codes_b200 import losses

    z = _load(golden_dir, "mstct_cfg3.npz")
    mods, x, lab, (B, T, K) = _mstct_modules(z)
    np.testing.assert_allclose(_sd_sums(mods), z["sd_sums"], rtol=1e-12, atol=1e-12,
                               err_msg="the mirror does not draw the reference's initial weights under the same seed")
    enc, mix, cls = [m.to(DEV).eval() for m in mods]
    feats = enc(x.to(DEV))
    for i, f in enumerate(feats):
        assert _maxabs(f[:, ::8, ::8], _t(z[f"enc_out_sample.{i}"])) <= 1e-3, i
    concat = mix(feats)
    assert _maxabs(concat[:, ::16, ::8], _t(z["concat_sample"])) <= 1e-3
    y, _ = cls(concat)
    assert _maxabs(y[:, ::8, :], _t(z["y_sample"])) <= 1e-3
    ref_arg = torch.from_numpy(z["y_argmax"]).long()
    assert torch.equal(y.argmax(-1).cpu(), ref_arg)
    loss = losses.bce_with_logits(y.reshape(B * T, K), lab.to(DEV).reshape(B * T, K))
    assert abs(float(loss) - float(z["loss"])) <= 1e-4 * abs(float(z["loss"]))
    loss.backward()
    names = [str(s) for s in z["grad_names"]]
    params = {}
    for pre, mod in (("TemporalEncoder.", enc), ("Temporal_Mixer.", mix), ("classifier.", cls)):
        params.update({pre + k: v for k, v in mod.named_parameters()})
    for i, k in enumerate(names):
        assert params[k].grad is not None, k
        _check_digest(params[k].grad, z[f"gdig.{i}"], i, 2e-4, k)
        if f"grad.{k}" in z.files:
            ref = _t(z[f"grad.{k}"])
            assert_grad_close(params[k].grad, ref, k, strict=2e-4, atol=1e-8)


# ------------------------------------------------------------------------------------------------ cfg4, T = 1800
def test_kd_loss_at_1800_frames_against_reference_golden(golden_dir):
    """DistillKL (the reference class, Spatial_cnn/run.py:284-295) at (T = 1800, K in {100, 6, 10, 15, 7})."""
    from computervision_codes_b200 import losses

    z = _load(golden_dir, "kd_t1800.npz")
    g = torch.Generator().manual_seed(4)
    kl = losses.DistillKL(4.0)
    for K in (100, 6, 10, 15, 7):
        ys = torch.randn(1800, K, generator=g)
        ytl = torch.randn(1800, K, generator=g) * 2
        np.testing.assert_allclose([float(ys.double().sum()), float(ytl.double().sum())], z[f"K{K}.seed_check"], rtol=1e-12)
        for fused in (False, True):
            yd = ys.to(DEV).requires_grad_(True)
            loss = kl(yd, ytl.to(DEV) if fused else torch.sigmoid(ytl.to(DEV)), teacher_is_logits=fused)
            ref = float(z[f"K{K}.loss"])
            assert abs(float(loss) - ref) <= 1e-4 * abs(ref)
            loss.backward()
            assert _maxabs(yd.grad[::16], _t(z[f"K{K}.gys_sample"])) <= 1e-7
            _check_digest(yd.grad, z[f"K{K}.gdig"], K, 1e-5, f"K{K}")


# ------------------------------------------------------------------------------------------------ cfg5 share
def test_cfg5_share_whole_model_step_against_cpu_port():
    """One GPU's share of BASELINE cfg5 (8 sequences x 8,000 frames x 768-d, VideoNas(fpn, 11/10/3, C = 64)): the
    executor's forward + loss + backward in eval-mode arithmetic against the oracle's torch port on the host, loss and
    gradients (per-video mean BCE averaged over the 8 sequences)."""
    from computervision_codes_b200 import losses
    from computervision_codes_b200.executor import ModelExecutor
    from computervision_codes_b200.layout import SeqLayout
    from computervision_codes_b200.tcn import VideoNas
    from oracle import torch_port as P

    nseq, T, D, C = 8, 8000, 768, 64
    args = types.SimpleNamespace(fpn=True, output=False, feature=False, trans=False, mask=False, hier=False)
    torch.manual_seed(5)
    m = VideoNas(args, 11, 10, 3, C, D, 100)
    x = torch.randn(nseq, T, D)
    g = torch.Generator().manual_seed(6)
    lab = (torch.rand(nseq * T, 132, generator=g) < 0.05).to(torch.uint8)
    lab[:, 131] = 0
    params = {k: v.detach().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    y = lab[:, :131].float().view(nseq, T, 131)
    total = 0.0
    for s in range(nseq):   # batch 1 per pass keeps the host memory of the autograd graph small
        ys = y[s]
        loss_s = P.train_step_loss(x[s:s + 1], params, (ys[:, 100:106], ys[:, 106:116], ys[:, 116:131], ys[:, 0:100]),
                                   train=False) / nseq
        loss_s.backward()
        total += float(loss_s)
    m = m.to(DEV)
    lay = SeqLayout.uniform(nseq, T, DEV)
    ex = ModelExecutor(m, max_rows=lay.rows, max_seqs=nseq)
    ex.set_batch(lay, seed=1)
    out = ex.train_step(x.reshape(nseq * T, D).to(DEV), lab.to(DEV), training=False).cpu()
    assert abs(float(out[4]) - total) <= 1e-4 * abs(total)
    for k, v in m.named_parameters():
        ref = params[k].grad
        if ref is None:
            continue
        assert v.grad is not None, k
        assert_grad_close(v.grad, ref, k, strict=5e-5)


# ------------------------------------------------------------------------------------------------ live reference
def _live_ref():
    from oracle import ref_import

    if not ref_import.available():
        pytest.skip("reference modules not on this box (baseline/_ref is written by oracle/vendor_ref.py)")
    return ref_import


@pytest.mark.parametrize("C,D,T", [(64, 2048, 1800), (512, 512, 1800)])
def test_videonas_against_live_reference_on_gpu(C, D, T):
    """The unmodified reference VideoNas on this GPU (fp32, TF32 off) against the drop-in on the same weights and input:
    every output, the loss and every gradient -- at the BASELINE width (64, 2048) and at the reference scripts' own
    width (512, 512: Temporal_tenco/run.py:89,313)."""
    ref_import = _live_ref()
    from computervision_codes_b200.tcn import VideoNas

    net = ref_import.tenco_network()
    args = types.SimpleNamespace(fpn=True, output=False, feature=False, trans=False, mask=False, hier=False)
    torch.manual_seed(C)
    ref = net.VideoNas(args, 11, 10, 3, C, D, 100).to(DEV).eval()
    m = VideoNas(args, 11, 10, 3, C, D, 100).to(DEV).eval()
    m.load_state_dict(ref.state_dict())
    x = torch.randn(1, T, D, device=DEV)
    labels = [(torch.rand(T, k, device=DEV) < 0.05).float() for k in (6, 10, 15, 100)]
    bce = torch.nn.BCEWithLogitsLoss()

    def run(model):
        outs = model(x, False)
        terms = [sum(bce(pd[0].transpose(0, 1), y) for pd in lst) for lst, y in zip((outs[1], outs[2], outs[3], outs[0]), labels)]
        loss = 0.1 * (terms[0] + terms[1] + terms[2]) + terms[3]
        loss.backward()
        return outs, loss

    o_ref, l_ref = run(ref)
    o_got, l_got = run(m)
    for a_list, r_list in zip(o_got[:5], o_ref[:5]):
        for a, r in zip(a_list, r_list):
            assert _maxabs(a, r) <= 1e-3
    for a_list, r_list in zip(o_got[:4], o_ref[:4]):
        for a, r in zip(a_list, r_list):
            _assert_argmax_consistent(a, r, 1)
    assert abs(float(l_got) - float(l_ref)) <= 1e-4 * abs(float(l_ref))
    pr = dict(ref.named_parameters())
    for k, v in m.named_parameters():
        if pr[k].grad is None:
            assert v.grad is None, k
            continue
        assert_grad_close(v.grad, pr[k].grad, k, strict=5e-5)


def test_mstct_cfg3_against_live_reference_on_gpu():
    """cfg3 at full size against the reference's TemporalEncoder / Temporal_Mixer / Classifier running on this GPU."""
    ref_import = _live_ref()
    from computervision_codes_b200 import losses
    from computervision_codes_b200.mstct import Classifier, TemporalEncoder, Temporal_Mixer

    enc_mod, mix_mod = ref_import.mstct_encoder(), ref_import.mstct_mixer()
    RefClassifier = ref_import.mstct_classifier_class()
    dims, B, T, K, D = [256, 384, 576, 864], 31, 256, 100, 768
    torch.manual_seed(3)
    r_enc = enc_mod.TemporalEncoder(in_feat_dim=D, embed_dims=dims, num_head=8, mlp_ratio=8, norm_layer=torch.nn.LayerNorm,
                                    num_block=2)
    r_mix = mix_mod.Temporal_Mixer(inter_channels=dims, embedding_dim=512)
    r_cls = RefClassifier(512, K)
    enc, mix, cls = TemporalEncoder(D, dims, 8, 8, torch.nn.LayerNorm, 2), Temporal_Mixer(dims, 512), Classifier(512, K)
    with torch.no_grad():
        for mod in (r_enc, r_mix, r_cls):
            for _, p in mod.named_parameters():
                if p.dim() == 1:
                    p.add_(0.1 * torch.randn_like(p))
    for a, b in ((enc, r_enc), (mix, r_mix), (cls, r_cls)):
        a.load_state_dict(b.state_dict())
        a.to(DEV).eval(), b.to(DEV).eval()
    x = torch.randn(B, D, T, device=DEV)
    lab = (torch.rand(B, T, K, device=DEV) < 0.05).float()
    bce = torch.nn.BCEWithLogitsLoss()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        yr, _ = r_cls(r_mix(r_enc(x)))
        lr = sum(bce(yr[i], lab[i]) for i in range(B)) / B
        lr.backward()
    y, _ = cls(mix(enc(x)))
    lg = losses.bce_with_logits(y.reshape(B * T, K), lab.reshape(B * T, K))
    lg.backward()
    assert _maxabs(y, yr) <= 1e-3
    _assert_argmax_consistent(y, yr, -1)
    assert abs(float(lg) - float(lr)) <= 1e-4 * abs(float(lr))
    for (a, b) in ((enc, r_enc), (mix, r_mix), (cls, r_cls)):
        pr = dict(b.named_parameters())
        for k, v in a.named_parameters():
            if pr[k].grad is None:
                continue
            assert_grad_close(v.grad, pr[k].grad, k, strict=2e-4, atol=1e-8)
