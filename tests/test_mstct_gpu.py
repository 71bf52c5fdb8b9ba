"""MS-TCT blocks (a9-a13) against golden fixtures produced by the reference's own modules.  GPU only."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _t(a):
    return torch.from_numpy(np.asarray(a)).float()


def _maxabs(a, b):
    return float((a.detach().double().cpu() - b.detach().double().cpu()).abs().max())


@pytest.mark.parametrize("name", ["mstct_small.npz", "mstct_mid.npz"])
def test_mstct_forward_backward_against_golden(golden_dir, name):
    from computervision_codes_b200.mstct import Classifier, TemporalEncoder, Temporal_Mixer

    z = np.load(os.path.join(golden_dir, name), allow_pickle=False)
    in_dim, d1, d2, d3, d4, heads, ratio, nblk, emb, K, B, T = [int(v) for v in z["cfg"]]
    enc = TemporalEncoder(in_dim, [d1, d2, d3, d4], heads, ratio, torch.nn.LayerNorm, nblk)
    mix = Temporal_Mixer([d1, d2, d3, d4], emb)
    cls = Classifier(emb, K)
    mods = (("TemporalEncoder.", enc), ("Temporal_Mixer.", mix), ("classifier.", cls))
    for pre, mod in mods:
        mod.load_state_dict({k[len("sd." + pre):]: _t(z[k]) for k in z.files if k.startswith("sd." + pre)})
        mod.to(DEV).eval()
    x = _t(z["x"]).to(DEV)
    feats = enc(x)
    for i, f in enumerate(feats):
        ref = _t(z[f"enc_out.{i}"])
        assert f.shape == ref.shape
        assert _maxabs(f, ref) <= 2e-4, (i, _maxabs(f, ref))
    concat = mix(feats)
    assert _maxabs(concat, _t(z["concat"])) <= 3e-4
    y, feat = cls(concat)
    ref_y = _t(z["y"])
    assert y.shape == ref_y.shape
    assert _maxabs(y, ref_y) <= 1e-3
    assert _maxabs(feat, _t(z["feat"])) <= 5e-4
    assert torch.equal(y.argmax(-1).cpu(), ref_y.argmax(-1))
    (y * _t(z["gy"]).to(DEV)).sum().backward()
    worst = 0.0
    for pre, mod in mods:
        for k, v in mod.named_parameters():
            key = "grad." + pre + k
            if key not in z.files:
                assert v.grad is None or float(v.grad.abs().max()) == 0.0, k
                continue
            ref = _t(z[key])
            assert v.grad is not None, k
            err = _maxabs(v.grad, ref) / max(1.0, float(ref.abs().max()))
            worst = max(worst, err)
            assert err <= 2e-4, (pre + k, err)


def test_mstct_videonas_wrapper_shapes_and_train_mode():
    import types

    from computervision_codes_b200.mstct import VideoNas

    torch.manual_seed(0)
    args = types.SimpleNamespace(loss_type="ivt")
    m = VideoNas(args, [32, 48, 64, 80], 2, 8, 2, 40, 32).to(DEV).train()
    x = torch.randn(3, 40, 70, device=DEV)
    (yi, fi), (yv, fv), (yt, ft), (yivt, cat) = m(x)
    assert yivt.shape == (3, 70, 100) and cat.shape == (3, 128, 70)
    assert yi.shape == (3, 70, 6) and float(yi.abs().max()) == 0.0
    loss = torch.nn.functional.binary_cross_entropy_with_logits(yivt, (torch.rand_like(yivt) < 0.1).float())
    loss.backward()
    g = m.TemporalEncoder.Temporal_Merging_Block1.proj.weight.grad
    assert g is not None and bool(torch.isfinite(g).all()) and float(g.abs().max()) > 0
    m.eval()
    with torch.no_grad():
        a = m(x)[3][0]
        b = m(x)[3][0]
    assert torch.equal(a, b)


def test_mstct_loss_composition_a14():
    """Temporal_mstct/run.py:155-196: per-sample BCE (pos_weight on i/v/t) averaged over the batch == one fused
    launch over the flattened (B*T, K) logits (all windows have the same length)."""
    from computervision_codes_b200 import losses

    torch.manual_seed(2)
    B, T = 5, 64
    for K, pw in ((6, losses.TOOL_WEIGHT), (10, losses.VERB_WEIGHT), (15, losses.TARGET_WEIGHT), (100, None)):
        y = torch.randn(B, T, K, device=DEV, requires_grad=True)
        lab = (torch.rand(B, T, K, device=DEV) < 0.1).float()
        fn = torch.nn.BCEWithLogitsLoss(pos_weight=None if pw is None else torch.tensor(pw, device=DEV))
        ref = sum(fn(y[i], lab[i]) for i in range(B)) / B           # the reference's python loop
        (gref,) = torch.autograd.grad(ref, y)
        got = losses.bce_with_logits(y.reshape(B * T, K), lab.reshape(B * T, K), pw)
        (ggot,) = torch.autograd.grad(got, y)
        assert abs(float(got) - float(ref)) <= 1e-5 * abs(float(ref))
        assert _maxabs(ggot, gref) <= 1e-7


@pytest.mark.parametrize("heads,hd", [(2, 3), (3, 5), (8, 32), (2, 72), (1, 108), (1, 128)])
def test_attention_kernels_ragged_windows_and_odd_head_dims(heads, hd):
    """Global_Relational_Block attention (Temporal_Encoder.py:76-88) on the tensor-core kernels, forward and both
    backward passes, against fp64 softmax attention per (window, head): windows of 1, 5, 63, 64, 65 and 200 frames
    packed in one call (chunk boundaries at 32 / 64 frames, a single-frame window), head dims that are not multiples
    of 4 or 8 (scalar loads, zero-padded k-slices) up to the 128 limit."""
    from computervision_codes_b200.layout import SeqLayout
    from computervision_codes_b200.mstct import functional as Fn

    lengths = [1, 5, 63, 64, 65, 200]
    lay = SeqLayout.get(lengths, DEV)
    d = heads * hd
    g = torch.Generator().manual_seed(heads * 1000 + hd)
    q = torch.zeros(lay.rows, d)
    kv = torch.zeros(lay.rows, 2 * d)
    go = torch.zeros(lay.rows, d)
    for s, T in enumerate(lengths):
        r0 = lay.starts[s]
        q[r0:r0 + T] = torch.randn(T, d, generator=g)
        kv[r0:r0 + T] = torch.randn(T, 2 * d, generator=g)
        go[r0:r0 + T] = torch.randn(T, d, generator=g)
    qd, kvd = q.to(DEV).requires_grad_(True), kv.to(DEV).requires_grad_(True)
    o = Fn.attention(qd, kvd, lay, heads)
    (o * go.to(DEV)).sum().backward()
    for s, T in enumerate(lengths):
        r0 = lay.starts[s]
        q64 = q[r0:r0 + T].double().requires_grad_(True)
        kv64 = kv[r0:r0 + T].double().requires_grad_(True)
        qh = q64.view(T, heads, hd).transpose(0, 1)
        kh = kv64[:, :d].reshape(T, heads, hd).transpose(0, 1)
        vh = kv64[:, d:].reshape(T, heads, hd).transpose(0, 1)
        att = torch.softmax(qh @ kh.transpose(1, 2) * hd ** -0.5, dim=-1)
        ref = (att @ vh).transpose(0, 1).reshape(T, d)
        (ref * go[r0:r0 + T].double()).sum().backward()
        assert _maxabs(o[r0:r0 + T], ref) <= 2e-5 * max(1.0, float(ref.abs().max())), (T, "o")
        assert _maxabs(qd.grad[r0:r0 + T], q64.grad) <= 5e-5 * max(1.0, float(q64.grad.abs().max())), (T, "dq")
        assert _maxabs(kvd.grad[r0:r0 + T], kv64.grad) <= 5e-5 * max(1.0, float(kv64.grad.abs().max())), (T, "dkv")
