"""The oracle restatement against the reference's own outputs (tests/golden/*.npz) -- CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import tcn_oracle as O

torch.set_num_threads(1)


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def _t(a, dtype=torch.float32):
    return torch.from_numpy(np.asarray(a)).to(dtype)


def _sd(z, prefix):
    return {k[len(prefix):]: _t(z[k]) for k in z.files if k.startswith(prefix)}


def test_layers_forward_backward(golden_dir):
    z = _load(golden_dir, "tcn_layers.npz")
    for i in range(int(z["num_cases"])):
        tag = f"c{i}."
        causal = str(z[tag + "kind"]) == "causal"
        d = int(z[tag + "dilation"])
        sd = _sd(z, tag + "sd.")
        x = _t(z[tag + "x"]).requires_grad_(True)
        ws = [sd["conv_dilated.weight"], sd["conv_dilated.bias"], sd["conv_1x1.weight"], sd["conv_1x1.bias"]]
        ws = [w.requires_grad_(True) for w in ws]
        y = O.dilated_residual_layer(x, *ws, d, causal=causal)
        assert torch.allclose(y, _t(z[tag + "y"]), atol=2e-6, rtol=1e-5)
        gy = _t(z[tag + "gy"])
        y.backward(gy)
        assert torch.allclose(x.grad, _t(z[tag + "gx"]), atol=5e-6, rtol=1e-5)
        names = ["conv_dilated.weight", "conv_dilated.bias", "conv_1x1.weight", "conv_1x1.bias"]
        for w, n in zip(ws, names):
            assert torch.allclose(w.grad, _t(z[tag + "grad." + n]), atol=2e-5, rtol=1e-4), n
        # the closed forms the CUDA kernels implement == autograd, in float64
        xd, gyd = x.detach().double(), gy.double()
        wd = [w.detach().double() for w in ws]
        gx, gw1, gb1, gw2, gb2 = O.layer_backward_closed_form(xd, gyd, *wd, d, causal=causal)
        assert torch.allclose(gx.float(), _t(z[tag + "gx"]), atol=5e-6, rtol=1e-5)
        for got, n in zip((gw1, gb1, gw2, gb2), names):
            assert torch.allclose(got.float(), _t(z[tag + "grad." + n]), atol=2e-5, rtol=1e-4), n


def test_closed_form_backward_with_dropout_fp64():
    torch.manual_seed(0)
    B, C, T, d = 2, 8, 30, 4
    for causal in (False, True):
        x = torch.randn(B, C, T, dtype=torch.float64, requires_grad=True)
        w1 = torch.randn(C, C, 3, dtype=torch.float64, requires_grad=True)
        b1 = torch.randn(C, dtype=torch.float64, requires_grad=True)
        w2 = torch.randn(C, C, 1, dtype=torch.float64, requires_grad=True)
        b2 = torch.randn(C, dtype=torch.float64, requires_grad=True)
        keep = (torch.rand(B, C, T) < 0.5).double()
        gy = torch.randn(B, C, T, dtype=torch.float64)
        y = O.dilated_residual_layer(x, w1, b1, w2, b2, d, causal=causal, keep=keep)
        y.backward(gy)
        got = O.layer_backward_closed_form(x.detach(), gy, w1.detach(), b1.detach(), w2.detach(),
                                           b2.detach(), d, causal=causal, keep=keep)
        for a, b in zip(got, (x.grad, w1.grad, b1.grad, w2.grad, b2.grad)):
            assert torch.allclose(a, b, atol=1e-12, rtol=1e-12)


def test_stage_pg_refinement(golden_dir):
    z = _load(golden_dir, "tcn_stage.npz")
    args = dict(use_output=False, hier=False)
    for tag, causal in (("acausal", False), ("causal", True)):
        sd = _sd(z, tag + ".sd.")
        x = _t(z[tag + ".x"])
        f0, l0 = O.base_tcn(x.permute(0, 2, 1), sd, "PG", causal=causal)
        f1, l1 = O.refinement(f0, sd, "Rs.0", causal=causal, **args)
        for got, name in ((f0, "f0"), (l0, "l0"), (f1, "f1"), (l1, "l1")):
            ref = _t(z[f"{tag}.{name}"])
            assert torch.allclose(got, ref, atol=2e-5, rtol=1e-4), (tag, name, (got - ref).abs().max())


def test_videonas_forward_loss_grads(golden_dir):
    z = _load(golden_dir, "tcn_videonas.npz")
    sd = {k: v.requires_grad_(True) for k, v in _sd(z, "sd.").items()}
    x = _t(z["x"])
    outs = O.videonas_forward(x, sd)
    for name, lst in zip(("ivt", "i", "v", "t", "f"), (outs[0], outs[1], outs[2], outs[3], outs[4])):
        for lvl, t in enumerate(lst):
            ref = _t(z[f"out_{name}.{lvl}"])
            assert torch.allclose(t, ref, atol=3e-5, rtol=1e-4), (name, lvl)
            if name != "f":  # argmax over classes identical per frame
                assert torch.equal(t.argmax(1), ref.argmax(1))
    labels = tuple(_t(z["label_" + n]) for n in "ivtq")
    loss, li, lv, lt, livt = O.tenco_loss(outs[:4], labels)
    assert abs(float(loss) - float(z["loss"])) <= 1e-5 * abs(float(z["loss"]))
    np.testing.assert_allclose([float(li), float(lv), float(lt), float(livt)], z["loss_terms"], rtol=2e-5)
    loss.backward()
    nograd = set(str(s) for s in z["nograd"])
    for k, v in sd.items():
        if "grad." + k in z.files:
            ref = _t(z["grad." + k])
            assert torch.allclose(v.grad, ref, atol=2e-6, rtol=2e-3), (k, (v.grad - ref).abs().max())
        elif v.grad is not None:
            assert k in nograd or float(v.grad.abs().max()) == 0.0, k
    # parameters the reference leaves without a gradient
    assert {"PG.conv_out.weight", "fpn.latlayer2.weight", "Rs.0.conv_1x1.weight"} <= nograd
    # TERL pos-weighted variant
    pws = tuple(torch.tensor(w) for w in (O.TOOL_WEIGHT, O.VERB_WEIGHT, O.TARGET_WEIGHT))
    with torch.no_grad():
        _, li, lv, lt, livt = O.tenco_loss(outs[:4], labels, pos_weights=pws)
    np.testing.assert_allclose([float(li), float(lv), float(lt), float(livt)], z["terl_loss_terms"], rtol=2e-5)


def test_kd_losses(golden_dir):
    z = _load(golden_dir, "kd_loss.npz")
    for i in range(int(z["num_kl"])):
        ys = _t(z[f"kl{i}.ys"]).requires_grad_(True)
        yt = torch.sigmoid(_t(z[f"kl{i}.yt_logits"]))
        loss = O.distill_kl(ys, yt, 4.0)
        assert abs(float(loss) - float(z[f"kl{i}.loss"])) <= 2e-5 * abs(float(z[f"kl{i}.loss"])) + 1e-9
        loss.backward()
        assert torch.allclose(ys.grad, _t(z[f"kl{i}.gys"]), atol=1e-7, rtol=1e-4)
        # closed form used by the kernel: T * (softmax(ys/T) - p_t) / N
        T, N = 4.0, ys.shape[0]
        closed = T * (torch.softmax(ys.detach() / T, 1) - torch.softmax(yt / T, 1)) / N
        assert torch.allclose(closed, _t(z[f"kl{i}.gys"]), atol=1e-7, rtol=1e-4)
    logits = [_t(z[f"comp.logits{k}"]).requires_grad_(True) for k in range(4)]
    labels = [_t(z[f"comp.labels{k}"]) for k in range(4)]
    teach = [_t(z[f"comp.teach{k}"]) for k in range(3)]
    feats = [_t(z[f"comp.feat{k}"]).requires_grad_(True) for k in range(3)]
    tfeats = [_t(z[f"comp.tfeat{k}"]) for k in range(3)]
    pws = tuple(torch.tensor(w) for w in (O.TOOL_WEIGHT, O.VERB_WEIGHT, O.TARGET_WEIGHT))
    loss, hard, soft, kd = O.multi_teacher_kd_loss(logits, labels, teach, feats, tfeats, T=4.0,
                                                   rates=(1.0, 1.0, 1.0), pos_weights=pws)
    np.testing.assert_allclose([float(loss), float(hard), float(soft), float(kd)], z["comp.loss"], rtol=2e-5)
    loss.backward()
    for k in range(4):
        assert torch.allclose(logits[k].grad, _t(z[f"comp.glogits{k}"]), atol=1e-7, rtol=1e-4)
    for k in range(3):
        assert torch.allclose(feats[k].grad, _t(z[f"comp.gfeat{k}"]), atol=1e-7, rtol=1e-4)


@pytest.mark.parametrize("tag", ["small", "wide"])
def test_feature_kd_attention_block(golden_dir, tag):
    """Row f3 against the reference's own forward/backward (oracle/gen_golden_kdattn.py)."""
    z = _load(golden_dir, "kd_attn.npz")
    params = {k: v.double().requires_grad_(True) for k, v in _sd(z, f"{tag}.sd.").items()}
    s = _t(z[f"{tag}.s"]).double().requires_grad_(True)
    teachers = [_t(z[f"{tag}.teacher_{n}"]).double() for n in "ivt"]
    outs = O.feature_kd_attention(s, teachers, params)
    for n, y in zip("ivt", outs):
        ref = z[f"{tag}.stus_f{n}"]
        assert np.abs(y.detach().numpy() - ref).max() <= 2e-6 * max(1.0, np.abs(ref).max())
    loss = O.feature_kd_loss(outs, teachers)
    assert abs(float(loss) - float(z[f"{tag}.kd_loss"])) <= 1e-6 * float(z[f"{tag}.kd_loss"])
    loss.backward()
    ref = z[f"{tag}.grad.s"]
    assert np.abs(s.grad.numpy() - ref).max() <= 1e-5 * np.abs(ref).max() + 1e-9
    for k, v in params.items():
        ref = z[f"{tag}.grad.{k}"]
        assert np.abs(v.grad.numpy() - ref).max() <= 1e-5 * np.abs(ref).max() + 1e-9, k


def test_phase_ce_matches_textbook():
    torch.manual_seed(3)
    x = torch.randn(50, 7)
    y = torch.randint(0, 7, (50,))
    assert torch.allclose(O.phase_ce(x, y), torch.nn.functional.cross_entropy(x, y), atol=1e-6)


def test_linear_resize_identity_and_general():
    torch.manual_seed(4)
    x = torch.randn(2, 3, 17)
    assert torch.equal(O.linear_resize(x, 17), x)
    ref = torch.nn.functional.interpolate(x, size=29, mode="linear")
    assert torch.allclose(O.linear_resize(x, 29), ref, atol=1e-5)


def test_torch_port_matches_golden(golden_dir):
    from oracle import torch_port as P

    z = _load(golden_dir, "tcn_videonas.npz")
    sd = {k: v.requires_grad_(True) for k, v in _sd(z, "sd.").items()}
    labels = tuple(_t(z["label_" + n]) for n in "ivtq")
    loss = P.train_step_loss(_t(z["x"]), sd, labels, train=False)
    assert abs(float(loss.detach()) - float(z["loss"])) <= 1e-5 * abs(float(z["loss"]))
    loss.backward()
    for k, v in sd.items():
        if "grad." + k in z.files:
            assert torch.allclose(v.grad, _t(z["grad." + k]), atol=2e-6, rtol=2e-3), k
