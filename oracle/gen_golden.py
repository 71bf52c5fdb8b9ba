"""Write tests/golden/*.npz by running the reference's own modules.  TEST INFRASTRUCTURE ONLY.

Run in the build container (needs /root/reference):  python -m oracle.gen_golden
Every fixture holds the seeded inputs, the reference module's state_dict, and the
reference's outputs / loss / gradients in float32 on CPU (the oracle of record:
CPU fp32, no TF32 anywhere).  Fixtures are small (a few hundred KB in total).
"""
from __future__ import annotations

import os
import types

import numpy as np
import torch

from . import ref_import
from . import tcn_oracle as O

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _np(t):
    return t.detach().cpu().numpy()


def _sd(mod, prefix="sd."):
    return {prefix + k: _np(v) for k, v in mod.state_dict().items()}


def gen_layers(net):
    """a1 / a2: DilatedResidualLayer and DilatedResidualCausalLayer, eval mode, fwd + bwd."""
    out = {}
    cases = [("acausal", 1), ("acausal", 4), ("acausal", 64), ("causal", 1), ("causal", 8),
             ("causal", 32)]
    for idx, (kind, d) in enumerate(cases):
        torch.manual_seed(100 + idx)
        C, T, B = 16, 45, 2
        cls = net.DilatedResidualLayer if kind == "acausal" else net.DilatedResidualCausalLayer
        m = cls(d, C, C).eval()
        x = torch.randn(B, C, T, requires_grad=True)
        y = m(x)
        gy = torch.randn_like(y)
        y.backward(gy)
        tag = f"c{idx}."
        out[tag + "kind"] = np.array(kind)
        out[tag + "dilation"] = np.array(d)
        out[tag + "x"] = _np(x)
        out[tag + "gy"] = _np(gy)
        out[tag + "y"] = _np(y)
        out[tag + "gx"] = _np(x.grad)
        for k, v in m.named_parameters():
            out[tag + "sd." + k] = _np(v)
            out[tag + "grad." + k] = _np(v.grad)
    out["num_cases"] = np.array(len(cases))
    np.savez_compressed(os.path.join(OUT, "tcn_layers.npz"), **out)


def gen_videonas(net, fname="tcn_videonas.npz", C=16, D=24, T=70, B=1, seed=7):
    """a3-a7: VideoNas(fpn) eval forward, tenco / TERL loss composition, parameter grads.
    The default fixture has 16 channels (mma.sync kernels); tcn_videonas_c64.npz (C = 64) pins the tcgen05 fused-layer
    kernels against the reference in one hop."""
    out = {}
    args = types.SimpleNamespace(fpn=True, output=False, feature=False, trans=False,
                                 mask=False, hier=False)
    torch.manual_seed(seed)
    m = net.VideoNas(args, 5, 4, 3, C, D, 100).eval()
    x = torch.randn(B, T, D)
    g = torch.Generator().manual_seed(1)
    labels = [(torch.rand(T, k, generator=g) < 0.1).long() for k in (6, 10, 15, 100)]
    outs = m(x, False)
    # Temporal_tenco/run.py:190-212 restated by the oracle (run.py is not importable); the
    # per-head terms use torch.nn.BCEWithLogitsLoss itself, as the reference does.
    bce = torch.nn.BCEWithLogitsLoss()
    terms = []
    for lst, y in zip((outs[1], outs[2], outs[3], outs[0]), labels):
        terms.append(sum(bce(pd[0].transpose(0, 1), y.float()) for pd in lst))
    loss = 0.1 * (terms[0] + terms[1] + terms[2]) + terms[3]
    loss.backward()
    out["cfg"] = np.array([5, 4, 3, C, D, 100, T, B])
    out["x"] = _np(x)
    for n, y in zip("ivtq", labels):
        out["label_" + n] = _np(y).astype(np.uint8)
    for name, lst in zip(("ivt", "i", "v", "t", "f"), (outs[0], outs[1], outs[2], outs[3], outs[4])):
        for lvl, t in enumerate(lst):
            out[f"out_{name}.{lvl}"] = _np(t)
    out["loss"] = _np(loss)
    out["loss_terms"] = np.array([float(t) for t in terms], dtype=np.float64)
    out.update(_sd(m))
    for k, v in m.named_parameters():
        if v.grad is not None:
            out["grad." + k] = _np(v.grad)
    out["nograd"] = np.array([k for k, v in m.named_parameters() if v.grad is None])

    # TERL variant: pos_weight BCE on i/v/t (TERL/0_5fold_TCN_black/run.py:320-343,481-485)
    pws = [torch.tensor(w) for w in (O.TOOL_WEIGHT, O.VERB_WEIGHT, O.TARGET_WEIGHT)]
    fns = [torch.nn.BCEWithLogitsLoss(pos_weight=w) for w in pws] + [bce]
    outs = m(x, False)
    terms = []
    for fn, lst, y in zip(fns, (outs[1], outs[2], outs[3], outs[0]), labels):
        terms.append(sum(fn(pd[0].transpose(0, 1), y.float()) for pd in lst))
    out["terl_loss_terms"] = np.array([float(t) for t in terms], dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, fname), **out)


def gen_stage(net):
    """a3 / a4 composed as the BASELINE cfg1 shape family: BaseCausalTCN -> Refinement,
    acausal (as shipped) and with the causal layer swapped in; phase (7) head logits."""
    out = {}
    args = types.SimpleNamespace(output=False, hier=False)
    for tag, causal in (("acausal", False), ("causal", True)):
        torch.manual_seed(11)
        C, D, T, K = 32, 40, 130, 7
        pg = net.BaseCausalTCN(6, C, D, K)
        rf = net.Refinement(args, 6, C, K, K, None)
        if causal:
            for st in (pg, rf):
                for i in range(len(st.layers)):
                    new = net.DilatedResidualCausalLayer(2 ** i, C, C)
                    new.load_state_dict(st.layers[i].state_dict())
                    st.layers[i] = new
        pg.eval(), rf.eval()
        x = torch.randn(2, T, D)
        f0, l0 = pg(x.permute(0, 2, 1))
        f1, l1 = rf(f0)
        out[f"{tag}.x"] = _np(x)
        out[f"{tag}.f0"], out[f"{tag}.l0"] = _np(f0), _np(l0)
        out[f"{tag}.f1"], out[f"{tag}.l1"] = _np(f1), _np(l1)
        out.update({f"{tag}.sd.PG." + k: _np(v) for k, v in pg.state_dict().items()})
        out.update({f"{tag}.sd.Rs.0." + k: _np(v) for k, v in rf.state_dict().items()})
    np.savez_compressed(os.path.join(OUT, "tcn_stage.npz"), **out)


def gen_refine_output(net):
    """a4 with args.output = True: Refinement applies its own conv_1x1 (K -> C) to the previous stage's K-channel output
    (network.py:150-151).  Forward + all parameter gradients."""
    out = {}
    torch.manual_seed(13)
    C, K, T, B, L = 32, 10, 90, 2, 4
    rf = net.Refinement(types.SimpleNamespace(output=True, hier=False), L, C, K, K, None).eval()
    x = torch.randn(B, K, T)
    f, lg = rf(x)
    gf, gl = torch.randn_like(f), torch.randn_like(lg)
    ((f * gf).sum() + (lg * gl).sum()).backward()
    out["cfg"] = np.array([L, C, K, T, B])
    out["x"], out["f"], out["logits"], out["gf"], out["gl"] = _np(x), _np(f), _np(lg), _np(gf), _np(gl)
    out.update({"sd." + k: _np(v) for k, v in rf.state_dict().items()})
    out.update({"grad." + k: _np(v.grad) for k, v in rf.named_parameters() if v.grad is not None})
    np.savez_compressed(os.path.join(OUT, "tcn_refine_output.npz"), **out)


def gen_kd():
    """a8: DistillKL (the reference class itself) + BCE / MSE composition of Spatial_cnn/run.py."""
    out = {}
    DistillKL = ref_import.distill_kl_class()
    g = torch.Generator().manual_seed(2)
    idx = 0
    for N in (8, 33):
        for K in (6, 10, 15, 100, 7):
            ys = torch.randn(N, K, generator=g, requires_grad=True)
            yt_logits = torch.randn(N, K, generator=g) * 2
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                loss = DistillKL(4.0)(ys, torch.sigmoid(yt_logits))
            loss.backward()
            out[f"kl{idx}.ys"], out[f"kl{idx}.yt_logits"] = _np(ys), _np(yt_logits)
            out[f"kl{idx}.loss"], out[f"kl{idx}.gys"] = _np(loss), _np(ys.grad)
            idx += 1
    out["num_kl"] = np.array(idx)
    # full composition, Spatial_cnn/run.py:159-192 with --rates 1 1 1 --temp 4
    N = 8
    pws = [torch.tensor(w) for w in (O.TOOL_WEIGHT, O.VERB_WEIGHT, O.TARGET_WEIGHT)]
    fns = [torch.nn.BCEWithLogitsLoss(pos_weight=w) for w in pws] + [torch.nn.BCEWithLogitsLoss()]
    logits = [torch.randn(N, k, generator=g, requires_grad=True) for k in (6, 10, 15, 100)]
    labels = [(torch.rand(N, k, generator=g) < 0.1).float() for k in (6, 10, 15, 100)]
    teach = [torch.randn(N, k, generator=g) * 2 for k in (6, 10, 15)]
    feats = [torch.randn(N, 48, generator=g, requires_grad=True) for _ in range(3)]
    tfeats = [torch.randn(N, 48, generator=g) for _ in range(3)]
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        hard = sum(fn(l, y) for fn, l, y in zip(fns, logits, labels))
        soft = sum(DistillKL(4.0)(logits[k], torch.sigmoid(teach[k])) for k in range(3)) / 3
    kd = sum(torch.nn.MSELoss()(feats[k], tfeats[k]) for k in range(3)) / 3
    loss = 1.0 * hard + 1.0 * soft + 1.0 * kd
    loss.backward()
    for k in range(4):
        out[f"comp.logits{k}"], out[f"comp.labels{k}"] = _np(logits[k]), _np(labels[k])
        out[f"comp.glogits{k}"] = _np(logits[k].grad)
    for k in range(3):
        out[f"comp.teach{k}"], out[f"comp.feat{k}"] = _np(teach[k]), _np(feats[k])
        out[f"comp.tfeat{k}"], out[f"comp.gfeat{k}"] = _np(tfeats[k]), _np(feats[k].grad)
    out["comp.loss"] = np.array([float(loss), float(hard), float(soft), float(kd)])
    np.savez_compressed(os.path.join(OUT, "kd_loss.npz"), **out)


def main():
    assert ref_import.available(), "needs /root/reference (build container only)"
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)
    net = ref_import.tenco_network()
    gen_layers(net)
    gen_videonas(net)
    gen_videonas(net, "tcn_videonas_c64.npz", C=64, D=96, T=300, B=1, seed=8)
    gen_stage(net)
    gen_refine_output(net)
    gen_kd()
    from . import gen_golden_mstct
    gen_golden_mstct.main()
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
