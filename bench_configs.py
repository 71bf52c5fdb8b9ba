"""bench.py --config cfg1 | cfg3 | cfg4 | cfg5: the other BASELINE.json configurations on ONE GPU.

Each prints one JSON line in bench.py's format (metric = frames/s of the configuration's train step) with, next to it,
the unmodified reference modules timed in the same run: on the host cores (`cpu_baseline`) and in PyTorch eager on the
same B200 with TF32 off / on (`eager_gpu_baseline`) -- SURVEY.md 8(d).  The default bench line (cfg2) lives in bench.py.
"""
from __future__ import annotations

import json
import os
import time
import types
import warnings

import torch

from bench import METRIC, ClockSampler, _quiet, hbm_peak, model_args


def _events(fn, steps, warmup, min_seconds=1.0):
    """CUDA-event timing of fn(): warm-up, a probe that sizes the repeat count, then >= min_seconds of calls."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()

    def run(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    probe = run(steps)
    reps = max(1, min(int(min_seconds * 1e3 / max(probe, 1e-3)) + 1, 500))
    ms = run(steps * reps)
    return ms / (steps * reps), reps, ms * 1e-3


def _graphed(step):
    """The whole step (forward + loss + backward through the autograd Functions) captured in ONE CUDA graph; returns
    a replay callable, or None when capture is not possible."""
    try:
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):
                step()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            step()
        g.replay()
        torch.cuda.synchronize()
        return g.replay
    except Exception:  # noqa: BLE001
        torch.cuda.synchronize()
        return None


def _cpu_time(fn, min_iters=2, budget_s=25.0):
    ncpu = os.cpu_count() or 1
    best = None
    t_all = time.time()
    for n in sorted({c for c in (ncpu, 16, 8, 4, 1) if c <= ncpu}, reverse=True):
        torch.set_num_threads(n)
        fn()
        t0 = time.time()
        it = 0
        while it < min_iters:
            fn()
            it += 1
        dt = (time.time() - t0) / it
        if best is None or dt < best[0]:
            best = (dt, n)
        if time.time() - t_all > budget_s:
            break
    return best


def _tf32(on):
    torch.backends.cudnn.allow_tf32 = on
    torch.backends.cuda.matmul.allow_tf32 = on


def _line(a, workload, frames, ms, reps, region_s, clocks, extra):
    line = {"metric": METRIC, "value": frames / (ms * 1e-3), "unit": "frames/s", "n_gpus": 1, "steps": a.steps,
            "warmup": a.warmup, "inner_repeats": reps, "timed_region_s": region_s, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (3xTF32 tensor-core products, fp32 accumulate)", "data": "synthetic",
            "config": {"workload": workload, "l2": "working set < 126 MB L2 unless stated; a 256 MB buffer is written "
                       "between timed regions, not between steps"}, "clocks": clocks}
    line.update(extra)
    return line


# ------------------------------------------------------------------------------------------------ cfg1
def _ref_stage_models(net, K, causal, dev):
    pg = _quiet(net.BaseCausalTCN, 10, 64, 2048, K)
    rf = net.Refinement(types.SimpleNamespace(output=False, hier=False), 10, 64, K, K, None)
    if causal:   # the layer the reference defines (network.py:165-183) but never instantiates
        for st in (pg, rf):
            for i in range(len(st.layers)):
                new = net.DilatedResidualCausalLayer(2 ** i, 64, 64)
                new.load_state_dict(st.layers[i].state_dict())
                st.layers[i] = new
    return pg.to(dev).train(), rf.to(dev).train()


def run_cfg1(a):
    """BASELINE configs[0]: TeCNO 2-stage TCN (10 layers, 64 ch) fwd + bwd on one 1,800 x 2048 sequence; the reference's
    acausal layer as shipped and the causal layer; K = 7 (phase, softmax-CE per stage) and K = 100 (triplet, BCE)."""
    from computervision_codes_b200 import losses
    from computervision_codes_b200.tcn import BaseCausalTCN, Refinement
    from oracle import ref_import

    dev = torch.device("cuda", 0)
    clocks = ClockSampler(0)
    T, D = 1800, 2048
    torch.manual_seed(0)
    x_host = torch.randn(1, T, D).pin_memory()
    x = x_host.to(dev)
    xt = x.permute(0, 2, 1)   # (B, D, T) view, as VideoNas.forward hands it to PG
    variants = {}
    net = ref_import.tenco_network() if ref_import.available() else None
    for causal in (True, False):
        for K in (7, 100):
            torch.manual_seed(0)
            pg = BaseCausalTCN(10, 64, D, K, causal=causal).to(dev).train()
            rf = Refinement(types.SimpleNamespace(output=False, hier=False), 10, 64, K, K, None, causal=causal).to(dev).train()
            params = list(pg.parameters()) + list(rf.parameters())
            g = torch.Generator().manual_seed(1)
            if K == 7:
                tgt = torch.randint(0, 7, (T,), generator=g).to(dev)
                lossf = lambda lg: losses.phase_cross_entropy(lg[0].transpose(0, 1).contiguous(), tgt)
                ref_lossf = lambda lg: torch.nn.functional.cross_entropy(lg[0].transpose(0, 1), tgt)
            else:
                lab = (torch.rand(T, K, generator=g) < 0.05).float().to(dev)
                lossf = lambda lg: losses.bce_with_logits(lg[0].transpose(0, 1).contiguous(), lab)
                ref_lossf = lambda lg: torch.nn.functional.binary_cross_entropy_with_logits(lg[0].transpose(0, 1), lab)

            def step(pg=pg, rf=rf, params=params, lossf=lossf):
                for p in params:
                    p.grad = None
                f0, l0 = pg(xt)
                f1, l1 = rf(f0)
                loss = lossf(l0) + lossf(l1)
                loss.backward()
                return loss

            name = f"{'causal' if causal else 'acausal'}_K{K}"
            with clocks.window():
                ms_eager, reps, region = _events(step, a.steps, a.warmup, a.min_seconds)
            rec = {"frames_per_s_eager_dispatch": T / (ms_eager * 1e-3), "ms_per_step_eager_dispatch": ms_eager}
            replay = _graphed(step)
            if replay is not None:
                with clocks.window():
                    ms_g, reps, region = _events(replay, a.steps, a.warmup, a.min_seconds)
                rec.update({"frames_per_s": T / (ms_g * 1e-3), "ms_per_step": ms_g, "cuda_graph": True})
            else:
                rec.update({"frames_per_s": rec["frames_per_s_eager_dispatch"], "ms_per_step": ms_eager, "cuda_graph": False})
            rec["inner_repeats"], rec["timed_region_s"] = reps, region

            # end to end: pinned host input -> device inside the step, loss read back
            loss_host = torch.empty((), dtype=torch.float32).pin_memory()

            def step_e2e(step=step):
                x.copy_(x_host, non_blocking=True)
                loss_host.copy_(step().detach(), non_blocking=True)
                torch.cuda.current_stream().synchronize()

            ms_e, _, _ = _events(step_e2e, a.steps, a.warmup, a.min_seconds)
            rec["e2e_frames_per_s"] = T / (ms_e * 1e-3)
            if net is not None:   # the reference modules: eager on this GPU, and on the host cores
                eg = {}
                for tag, on in (("tf32_off", False), ("tf32_on", True)):
                    _tf32(on)
                    torch.manual_seed(0)
                    rpg, rrf = _ref_stage_models(net, K, causal, dev)
                    rparams = list(rpg.parameters()) + list(rrf.parameters())

                    def rstep(rpg=rpg, rrf=rrf, rparams=rparams, inp=xt, lf=ref_lossf):
                        for p in rparams:
                            p.grad = None
                        f0, l0 = rpg(inp)
                        f1, l1 = rrf(f0)
                        (lf(l0) + lf(l1)).backward()

                    ms_r, _, _ = _events(rstep, 5, 2, 0.3)
                    eg[tag] = T / (ms_r * 1e-3)
                _tf32(False)
                rec["eager_gpu_baseline"] = eg
                if not a.no_cpu_baseline:
                    torch.manual_seed(0)
                    cpg, crf = _ref_stage_models(net, K, causal, "cpu")
                    cparams = list(cpg.parameters()) + list(crf.parameters())
                    xc = x_host.permute(0, 2, 1)
                    lf_c = ((lambda lg: torch.nn.functional.cross_entropy(lg[0].transpose(0, 1), tgt.cpu())) if K == 7 else
                            (lambda lg: torch.nn.functional.binary_cross_entropy_with_logits(lg[0].transpose(0, 1), lab.cpu())))

                    def cstep():
                        for p in cparams:
                            p.grad = None
                        f0, l0 = cpg(xc)
                        f1, l1 = crf(f0)
                        (lf_c(l0) + lf_c(l1)).backward()

                    dt, n = _cpu_time(cstep, budget_s=8.0)
                    rec["cpu_baseline"] = {"value": T / dt, "unit": "frames/s", "cores": n, "kind": "reference"}
            variants[name] = rec
    ck = clocks.stop()
    head = variants["causal_K7"]
    extra = {"variants": variants,
             "e2e": {"value": head["e2e_frames_per_s"], "unit": "frames/s", "h2d_bytes_per_step": T * D * 4, "d2h_bytes_per_step": 4},
             "gpu_launches": None,
             "note": "value = causal layer, K = 7 (TeCNO phase head), whole step replayed from one CUDA graph when capture "
                     "succeeds; 41 launches-deep dependency chain on a 1,800-frame sequence: latency-bound (15 tiles of 128 frames "
                     "on 148 SMs), SURVEY 8(d)"}
    if "cpu_baseline" in head:
        cb = dict(head["cpu_baseline"])
        cb["sample"] = "reference BaseCausalTCN(10,64,2048,7) -> Refinement(10) with the causal layer, fwd + CE + bwd, train mode, one 1,800-frame sequence, best thread count"
        extra["cpu_baseline"] = cb
    if "eager_gpu_baseline" in head:
        extra["eager_gpu_baseline"] = dict(head["eager_gpu_baseline"], unit="frames/s")
    print(json.dumps(_line(a, "cfg1: TeCNO 2-stage dilated TCN (BaseCausalTCN(10,64,2048,K) -> Refinement(10 layers)), fwd + loss + bwd "
                           "on one 1,800 x 2048 sequence, causal / acausal x K = 7 (CE) / 100 (BCE)", T, head["ms_per_step"],
                           head["inner_repeats"], head["timed_region_s"], ck, extra)))


# ------------------------------------------------------------------------------------------------ cfg3
def run_cfg3(a):
    """BASELINE configs[2]: MS-TCT temporal head over 768-d features, (31, 768, 256), forward + BCE ('ivt') + backward."""
    from computervision_codes_b200 import losses
    from computervision_codes_b200.mstct import VideoNas
    from oracle import ref_import

    dev = torch.device("cuda", 0)
    clocks = ClockSampler(0)
    B, T, D, K = 31, 256, 768, 100
    dims = [256, 384, 576, 864]
    torch.manual_seed(0)
    m = VideoNas(types.SimpleNamespace(loss_type="ivt"), dims, 2, 8, 8, D, 512).to(dev).train()
    x = torch.randn(B, D, T, device=dev)
    lab = (torch.rand(B * T, K, device=dev) < 0.05).float()
    params = list(m.parameters())

    def step():
        for p in params:
            p.grad = None
        y = m(x)[3][0]
        loss = losses.bce_with_logits(y.reshape(B * T, K), lab)
        loss.backward()
        return loss

    with clocks.window():
        ms, reps, region = _events(step, a.steps, a.warmup, a.min_seconds)
    flops = 2.93e12   # forward + backward, SURVEY 8(d)
    extra = {"algorithmic_TFLOPs": flops / ms / 1e9, "executed_TFLOPs_3xtf32": 3 * flops / ms / 1e9,
             "e2e": None, "gpu_launches": None}
    if ref_import.available():
        enc_mod, mix_mod = ref_import.mstct_encoder(), ref_import.mstct_mixer()
        RefClassifier = ref_import.mstct_classifier_class()

        def build(device):
            torch.manual_seed(0)
            enc = enc_mod.TemporalEncoder(in_feat_dim=D, embed_dims=dims, num_head=8, mlp_ratio=8,
                                          norm_layer=torch.nn.LayerNorm, num_block=2)
            mix = mix_mod.Temporal_Mixer(inter_channels=dims, embedding_dim=512)
            cls = RefClassifier(512, K)
            drop = torch.nn.Dropout(0.5)
            mods = [mm.to(device).train() for mm in (enc, mix, cls)]
            ps = [p for mm in mods for p in mm.parameters()]
            bce = torch.nn.BCEWithLogitsLoss()

            def rstep(xi, li):
                for p in ps:
                    p.grad = None
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    y, _ = mods[2](mods[1](mods[0](drop(xi))))      # Temporal_mstct/network.py:75-101 composed by hand
                loss = sum(bce(y[i], li[i]) for i in range(B)) / B  # run.py:185-196
                loss.backward()
            return rstep

        eg = {}
        l3 = lab.view(B, T, K)
        for tag, on in (("tf32_off", False), ("tf32_on", True)):
            _tf32(on)
            rstep = build(dev)
            ms_r, _, _ = _events(lambda: rstep(x, l3), 5, 2, 0.5)
            eg[tag] = B * T / (ms_r * 1e-3)
        _tf32(False)
        eg["unit"] = "frames/s"
        extra["eager_gpu_baseline"] = eg
        if not a.no_cpu_baseline:
            rstep = build("cpu")
            xc, lc = x.cpu(), l3.cpu()
            dt, n = _cpu_time(lambda: rstep(xc, lc), min_iters=1, budget_s=40.0)
            extra["cpu_baseline"] = {"value": B * T / dt, "unit": "frames/s", "cores": n, "kind": "reference",
                                     "sample": "reference TemporalEncoder + Temporal_Mixer + Classifier, one (31, 768, 256) step, fwd + BCE + bwd, train mode"}
    print(json.dumps(_line(a, "cfg3: MS-TCT temporal head, TemporalEncoder(768,[256,384,576,864],8 heads,mlp 8,2 blocks) + "
                           "Temporal_Mixer + Classifier(512,100), 31 windows x 256 frames, fwd + BCE + bwd (per-op path)",
                           B * T, ms, reps, region, clocks.stop(), extra)))


# ------------------------------------------------------------------------------------------------ cfg4
def run_cfg4(a):
    """BASELINE configs[3]: the multi-teacher KD loss of Spatial_cnn/run.py:159-192 on student logits of one 1,800-frame
    video: heads 100 / 6 / 10 / 15 (BCE, pos_weight on i/v/t; KL against sigmoid(teacher) at T = 4 on i/v/t) + phase (7, CE),
    forward + backward w.r.t. the student logits."""
    from computervision_codes_b200 import losses
    from oracle import ref_import

    dev = torch.device("cuda", 0)
    clocks = ClockSampler(0)
    T = 1800
    g = torch.Generator().manual_seed(2)
    Ks = (6, 10, 15, 100)
    logits = [torch.randn(T, k, generator=g).to(dev).requires_grad_(True) for k in Ks]
    labels = [(torch.rand(T, k, generator=g) < 0.05).float().to(dev) for k in Ks]
    teach = [(torch.randn(T, k, generator=g) * 2).to(dev) for k in Ks[:3]]
    ph = torch.randn(T, 7, generator=g).to(dev).requires_grad_(True)
    ph_t = torch.randint(0, 7, (T,), generator=g).to(dev)
    crit = losses.MultiTeacherKDLoss(temp=4.0, rates=(1.0, 1.0, 1.0))

    def step():
        for t in logits + [ph]:
            t.grad = None
        loss = crit(logits, labels, teach)[0] + losses.phase_cross_entropy(ph, ph_t)
        loss.backward()
        return loss

    with clocks.window():
        ms_eager, reps, region = _events(step, a.steps, a.warmup, a.min_seconds)
    replay = _graphed(step)
    ms = ms_eager
    if replay is not None:
        with clocks.window():
            ms, reps, region = _events(replay, a.steps, a.warmup, a.min_seconds)
    extra = {"ms_per_step_eager_dispatch": ms_eager, "cuda_graph": replay is not None, "e2e": None, "gpu_launches": None}
    if ref_import.available():
        DistillKL = ref_import.distill_kl_class()
        pws = [torch.tensor(w) for w in (losses.TOOL_WEIGHT, losses.VERB_WEIGHT, losses.TARGET_WEIGHT)]

        def build(device):
            lg = [t.detach().to(device).requires_grad_(True) for t in logits]
            lb = [t.to(device) for t in labels]
            tc = [t.to(device) for t in teach]
            p7, t7 = ph.detach().to(device).requires_grad_(True), ph_t.to(device)
            fns = [torch.nn.BCEWithLogitsLoss(pos_weight=w.to(device)) for w in pws] + [torch.nn.BCEWithLogitsLoss()]
            kl = DistillKL(4.0)

            def rstep():
                for t in lg + [p7]:
                    t.grad = None
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    hard = sum(fn(l, y) for fn, l, y in zip(fns, lg, lb))
                    soft = sum(kl(lg[k], torch.sigmoid(tc[k])) for k in range(3)) / 3
                (hard + soft + torch.nn.functional.cross_entropy(p7, t7)).backward()
            return rstep

        rstep = build(dev)
        ms_r, _, _ = _events(rstep, 10, 3, 0.3)
        extra["eager_gpu_baseline"] = {"tf32_off": T / (ms_r * 1e-3), "tf32_on": T / (ms_r * 1e-3), "unit": "frames/s",
                                       "what": "reference DistillKL + torch BCEWithLogitsLoss / cross_entropy, eager on this GPU (no matmul: TF32 irrelevant)"}
        if not a.no_cpu_baseline:
            cstep = build("cpu")
            dt, n = _cpu_time(cstep, min_iters=5, budget_s=8.0)
            extra["cpu_baseline"] = {"value": T / dt, "unit": "frames/s", "cores": n, "kind": "reference",
                                     "sample": "the same loss composition with the reference DistillKL class on the host, fwd + bwd"}
    print(json.dumps(_line(a, "cfg4: multi-teacher KD loss (BCE + pos_weight, DistillKL T=4 against sigmoid(teacher) on i/v/t, phase CE) on "
                           "student logits of one 1,800-frame video, K = 100/6/10/15/7, fwd + bwd", T, ms, reps, region,
                           clocks.stop(), extra)))


# ------------------------------------------------------------------------------------------------ cfg5
def run_cfg5(a):
    """One GPU's share of BASELINE configs[4] (64 x 8,000 frames over 8 GPUs = 8 sequences per GPU per step), D = 768."""
    from computervision_codes_b200.tcn import VideoNas
    from computervision_codes_b200.trainer import TemporalTrainer

    dev = torch.device("cuda", 0)
    clocks = ClockSampler(0)
    nseq, T, D = 8, 8000, 768
    torch.manual_seed(5)
    m = VideoNas(model_args(), 11, 10, 3, 64, D, 100).to(dev).train()
    tr = TemporalTrainer(m, lr=1e-2, weight_decay=1e-5, max_frames=nseq * T, max_seqs=nseq, input_mask_p=0.25)
    x = torch.randn(nseq * T, D, device=dev)
    lab = (torch.rand(nseq * T, 132, device=dev) < 0.05).to(torch.uint8)
    lens = [T] * nseq
    with clocks.window():
        ms, reps, region = _events(lambda: tr.step(x, lab, lens), a.steps, a.warmup, a.min_seconds)
    peak, src = hbm_peak()
    alg = 70.5e3 * nseq * T   # 70.5 KB / frame at C = 64, D = 768
    extra = {"gpu_launches": tr.launches_per_step() * a.steps * reps, "gpu_launches_per_step": tr.launches_per_step(), "e2e": None,
             "roofline": {"bound": "hbm", "kernel": "whole train step (SURVEY 8(d) algorithmic bytes of the full model: 70.5 KB/frame)",
                          "achieved": alg / (ms * 1e-3) / 1e9, "peak": peak, "peak_source": src, "unit": "GB/s",
                          "frac": alg / (ms * 1e-3) / 1e9 / peak, "traffic": None}}
    tr.close()
    print(json.dumps(_line(a, "cfg5 share: 8 sequences x 8,000 frames x 768-d (one GPU's part of 64 x 8,000 on 8 GPUs), "
                           "VideoNas(fpn,11/10/3,C=64), full train step (fwd + loss + bwd + SGD), inputs resident (197 MB > L2)",
                           nseq * T, ms, reps, region, clocks.stop(), extra)))


def run(a):
    {"cfg1": run_cfg1, "cfg3": run_cfg3, "cfg4": run_cfg4, "cfg5": run_cfg5}[a.config](a)
