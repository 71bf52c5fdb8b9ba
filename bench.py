#!/usr/bin/env python
"""bench.py -- temporal-head train frames/s (BASELINE.json metric) on N B200s of one node.

Workload (BASELINE.json configs[1]): one 5-fold training epoch schedule over 45 synthetic
CholecT45-shaped videos (seeded ragged lengths 900..3600 frames, 2048-d fp32 features, 100/6/10/15
multi-label heads), VideoNas(fpn, 11/10/3 layers, 64 channels), data-parallel by video.  One "step"
= every rank runs forward + loss + backward over its next batch of `--videos-per-step` videos, then
one all-reduce of the flat gradient buffer and one SGD update.

  python bench.py [--gpus N --steps K --warmup W]          this repo's CUDA path
  python bench.py --impl reference ...                     CPU arm: the oracle port on the host cores
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

CROSSVAL = {  # cholect45-crossval test folds, MT4MTLKD/Temporal_tenco/dataloader.py:131-137
    1: [79, 2, 51, 6, 25, 14, 66, 23, 50], 2: [80, 32, 5, 15, 40, 47, 26, 48, 70],
    3: [31, 57, 36, 18, 52, 68, 10, 8, 73], 4: [42, 29, 60, 27, 65, 75, 22, 49, 12],
    5: [78, 43, 62, 35, 74, 1, 56, 4, 13],
}
D_FEAT, C_MAPS, HEADS = 2048, 64, (100, 6, 10, 15)
LAYERS = (11, 10, 3)


def video_table():
    vids = sorted(v for f in CROSSVAL.values() for v in f)
    rng = np.random.RandomState(45)
    lengths = rng.randint(900, 3600, size=len(vids))
    return vids, {v: int(t) for v, t in zip(vids, lengths)}


def fold_schedule():
    """155 training passes: for each fold the 31 videos that are neither its test fold nor its
    5-video validation subset (first five of the next fold), in fold order."""
    vids, lengths = video_table()
    passes = []
    for k in range(1, 6):
        test = set(CROSSVAL[k])
        val = set(CROSSVAL[k % 5 + 1][:5])
        passes += [v for v in vids if v not in test and v not in val]
    return passes, lengths


def make_video(v, T, pinned):
    g = torch.Generator().manual_seed(1000 + v)
    x = torch.randn(T, D_FEAT, generator=g)
    lab = (torch.rand(T, 132, generator=g) < 0.05).to(torch.uint8)
    lab[:, 131] = 0
    if pinned:
        x, lab = x.pin_memory(), lab.pin_memory()
    return x, lab


def build_model(device):
    from computervision_codes_b200.tcn import VideoNas

    args = types.SimpleNamespace(fpn=True, output=False, feature=False, trans=False, mask=False, hier=False)
    torch.manual_seed(0)
    return VideoNas(args, *LAYERS, C_MAPS, D_FEAT, HEADS[0]).to(device).train()


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(prefix="clocks_", suffix=".csv")
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            f = [s.strip() for s in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.path)
        except OSError:
            pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------- CPU arm
def cpu_arm(steps, warmup, budget_s=25.0):
    """The oracle's torch-CPU port (oracle/torch_port.py) of the same train step, on the host cores:
    forward + tenco loss + backward, train mode, one video per step, thread count swept."""
    from oracle import torch_port as P

    passes, lengths = fold_schedule()
    model = build_model("cpu")
    params = {k: v.detach().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    vids = passes[: max(1, min(3, steps))]
    data = []
    for v in vids:
        x, lab = make_video(v, lengths[v], pinned=False)
        y = lab[:, :131].float()
        data.append((x.unsqueeze(0), (y[:, 100:106], y[:, 106:116], y[:, 116:131], y[:, 0:100]), lengths[v]))

    def one(x, labels):
        for p in params.values():
            p.grad = None
        loss = P.train_step_loss(x, params, labels, train=True)
        loss.backward()
        return float(loss)

    ncpu = os.cpu_count() or 1
    cands = sorted({n for n in (1, 2, 4, 8, 16, 32, 64, ncpu) if n <= ncpu})
    best = None
    t_start = time.time()
    for n in cands:
        torch.set_num_threads(n)
        one(*data[0][:2])  # warm-up
        t0 = time.time()
        frames = 0
        for x, labels, T in data:
            one(x, labels)
            frames += T
        dt = time.time() - t0
        fps = frames / dt
        if best is None or fps > best[0]:
            best = (fps, n, dt / len(data))
        if time.time() - t_start > budget_s:
            break
    return {"value": best[0], "unit": "frames/s", "cores": best[1], "kind": "port",
            "sample": f"{len(data)} videos ({sum(d[2] for d in data)} frames) of the same schedule, fwd+loss+bwd, "
                      f"train mode, torch {torch.__version__} CPU (oneDNN), best of threads {cands}, host has {ncpu} cpus",
            "ms_per_video": best[2] * 1e3}


# --------------------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--videos-per-step", type=int, default=8, help="videos per rank per step (ragged batch)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="plain stream launches instead of a CUDA graph")
    a = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": "cfg2: 5-fold epoch schedule over 45 synthetic CholecT45-shaped videos "
                          "(ragged 900..3600 frames, seed 45), VideoNas(fpn, 11/10/3, C=64, D=2048, heads 100/6/10/15), "
                          "tenco BCE loss, SGD(lr 1e-2, wd 1e-5)",
              "videos_per_rank_per_step": a.videos_per_step, "parallelism": f"dp{world}-by-video, videos of a global step assigned longest-first (LPT), equal count per rank",
              "l2": "inputs cycle through the 0.8 GB feature set (>> 126 MB L2); no explicit flush",
              "resident_inputs": "one HBM feature arena (FeatureCache.pack), read in place by the step"}

    if a.impl == "reference":
        if rank != 0:
            return
        cb = cpu_arm(a.steps, a.warmup, budget_s=60.0)
        line = {"impl": "reference", "metric": "temporal-head train frames/s", "value": cb["value"], "unit": "frames/s",
                "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": cb["ms_per_video"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config, "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": cb["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    assert torch.cuda.is_available(), "bench.py needs a GPU (use --impl reference for the CPU arm)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    pg = None
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
        pg = torch.distributed.group.WORLD

    from computervision_codes_b200 import _lib
    from computervision_codes_b200.trainer import TemporalTrainer

    _lib.load()
    model = build_model(dev)
    passes, lengths = fold_schedule()
    V = a.videos_per_step
    max_frames = max(lengths.values()) * V
    trainer = TemporalTrainer(model, lr=1e-2, weight_decay=1e-5, process_group=pg, world_size=world,
                              max_frames=max_frames, max_seqs=max(V, 1), use_graph=not a.no_graph,
                              input_mask_p=0.25)  # --mask of Scripts/train_fold1.sh:28

    nsteps_total = a.warmup + a.steps
    # Global step s takes the next V * world passes of the schedule (cycled) and assigns them to the ranks
    # longest-first, V videos each (length-balanced DP by video).  Every rank computes the same table.
    from computervision_codes_b200.trainer import lpt_assign
    batches = []
    for s in range(nsteps_total):
        vids = [passes[(s * V * world + j) % len(passes)] for j in range(V * world)]
        shard = lpt_assign([lengths[v] for v in vids], world, cap=V)[rank]
        batches.append([vids[i] for i in shard])
    needed = sorted({v for b in batches for v in b})
    host = {v: make_video(v, lengths[v], pinned=True) for v in needed}
    # resident arm: the feature set lives in HBM as one arena (data.FeatureCache.pack: SURVEY 8(f1)); a step reads its
    # videos in place -- the block table carries their positions -- so nothing is copied inside the timed region
    from computervision_codes_b200.data import FeatureCache
    cache = FeatureCache(dev)
    for v, (x, lab) in host.items():
        cache.add_packed(v, x, lab)
    cache.pack()

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(run_step):
        for s in range(a.warmup):
            run_step(batches[s])
        barrier()
        clocks = ClockSampler(local_rank)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(a.warmup, nsteps_total):
            run_step(batches[s])
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        ck = clocks.stop()
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t.item()), ck

    # ---- arm 1: inputs resident in HBM (the step starts from device tensors)
    def step_resident(b):
        trainer.step_cached(cache, [(v, 0, lengths[v]) for v in b])

    ms_total, clocks = timed(step_resident)
    frames_rank = sum(lengths[v] for b in batches[a.warmup:] for v in b)
    fr = torch.tensor([frames_rank], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(fr)
    frames_all = float(fr.item())
    value = frames_all / (ms_total * 1e-3)

    # ---- arm 2: end to end through the public API with HOST (pinned) inputs: H2D of this step's
    # features + labels and D2H of the loss vector inside the timed region, every step
    loss_host = torch.empty(8, dtype=torch.float32).pin_memory()
    h2d = [0]

    order = {id(b): i for i, b in enumerate(batches)}

    def stage(b):
        xs, ls = [host[v][0] for v in b], [host[v][1] for v in b]
        h2d[0] = sum(t.numel() * t.element_size() for t in xs + ls)
        trainer.prefetch(xs, ls, [lengths[v] for v in b])

    def step_e2e(b):
        # software pipeline over the public API: the H2D of the next step's inputs (pinned host -> device, copy
        # stream) is issued before this step's kernels; every step still pays its own H2D and its own D2H read.
        i = order[id(b)]
        if trainer._prefetched is None:
            stage(b)
        out = trainer.step()
        if i + 1 < len(batches):
            stage(batches[i + 1])
        loss_host.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    ms_e2e, _ = timed(step_e2e)
    trainer._prefetched = None
    e2e_value = frames_all / (ms_e2e * 1e-3)

    # ---- roofline of the dominant kernel, timed live with CUDA events
    roof = measure_layer_roofline(model, lengths, batches[a.warmup], dev)

    def shutdown():
        # graphs hold NCCL work: drop them before the process group, and never let a stuck teardown hang the job
        trainer.close()
        torch.cuda.synchronize()
        if world > 1:
            import threading
            threading.Timer(20.0, lambda: os._exit(0)).start()
            try:
                torch.distributed.barrier()
                torch.distributed.destroy_process_group()
            except Exception:
                pass

    if rank != 0:
        shutdown()
        os._exit(0)
    cb = None if a.no_cpu_baseline or world > 1 else cpu_arm(a.steps, a.warmup)
    launches = trainer.launches_per_step() * a.steps
    line = {"metric": "temporal-head train frames/s", "value": value, "unit": "frames/s", "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_total / a.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32 (3xTF32 tensor-core products, fp32 accumulate)",
            "data": "synthetic", "config": config, "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d[0], "d2h_bytes_per_step": 32,
                    "ms_per_step": ms_e2e / a.steps},
            "gpu_launches": launches, "roofline": roof}
    if cb is not None:
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line))
    sys.stdout.flush()
    shutdown()
    os._exit(0)


def _graph_time(fn, n=24):
    """Average GPU time of fn(): n calls captured in one CUDA graph, replayed once under CUDA events."""
    for _ in range(3):
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


def measure_layer_roofline(model, lengths, batch, dev):
    """Roofline of the step's dominant kernel, timed live with CUDA events (graph of 24 launches on the current stream).

    A train step is 41 residual layers x three launches -- `layer_fwd_tc_kernel` (fused forward), `layer_bwd_tc_kernel`
    (fused input gradient) and `wgrad_tc_pair_kernel` (both weight gradients) -- plus a few dozen one-off launches.
    All three are timed here on this step's batch shape and on the TERL stress shape (64 x 8000 frames, where the HBM
    roofline is the binding limit); the one with the largest time per step is the dominant kernel and fills `roofline`,
    the others are listed under `roofline.kernels`.  Algorithmic bytes per frame (DESIGN.md section 3, C = 64 fp32):
      layer_fwd_tc  12*C + 16  read x, write h, write y (+ 16 B of ReLU / dropout bit words)
      layer_bwd_tc  12*C + 16  read gy, write gu, write gx (+ the bit words)
      wgrad_tc_pair 16*C       read gu, x, gy, h (the 64 x 256 weight-gradient tile is negligible)"""
    import ctypes as Ct

    from computervision_codes_b200 import _lib, ops
    from computervision_codes_b200.layout import SeqLayout

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak, src = 6650.0, "fallback"
    if os.path.exists(peaks_path):
        peak, src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured"
    lib = _lib.load()
    layer = model.PG.layers[4]
    d = int(layer.conv_dilated.dilation[0])
    shifts = (-2 * d, -d, 0)
    C = C_MAPS
    w1, w2 = layer.conv_dilated.weight.detach(), layer.conv_1x1.weight.detach()
    b1, b2 = layer.conv_dilated.bias.detach().contiguous(), layer.conv_1x1.bias.detach().contiguous()
    w1h, w1l = ops.split_weight(w1)
    w2h, w2l = ops.split_weight(w2)
    w1th, w1tl = ops.split_weight(w1, transpose=True)
    w2th, w2tl = ops.split_weight(w2, transpose=True)
    p_drop = 0.5

    def run(lens, nbuf):
        lay = SeqLayout.get(lens, dev)
        rnd = lambda: [torch.randn(lay.rows, C, device=dev) for _ in range(nbuf)]
        xs, gys = rnd(), rnd()
        ys, hs, gus, gxs = ([torch.zeros(lay.rows, C, device=dev) for _ in range(nbuf)] for _ in range(4))
        masks = [torch.zeros(lay.rows, 4, device=dev, dtype=torch.int32) for _ in range(nbuf)]
        gw1, gb1 = torch.zeros(C, C, 3, device=dev), torch.zeros(C, device=dev)
        gw2, gb2 = torch.zeros(C, C, 1, device=dev), torch.zeros(C, device=dev)
        it = [0]

        def fwd():
            i = it[0] % nbuf
            it[0] += 1
            a = _lib.LayerFwdTcArgs()
            a.x, a.x_rows, a.y, a.h = xs[i].data_ptr(), lay.rows, ys[i].data_ptr(), hs[i].data_ptr()
            a.w1_hi, a.w1_lo, a.w2_hi, a.w2_lo = w1h.data_ptr(), w1l.data_ptr(), w2h.data_ptr(), w2l.data_ptr()
            a.b1, a.b2 = b1.data_ptr(), b2.data_ptr()
            a.meta, a.nblk, a.channels = lay.meta.data_ptr(), lay.nblk, C
            for k, s_ in enumerate(shifts):
                a.shift[k] = s_
            a.drop_p, a.drop_seed, a.drop_stream = p_drop, 7, 4
            a.masks = masks[i].data_ptr()
            _lib.check(lib.tcn_layer_fwd_tc(Ct.byref(a), _lib.stream_ptr()), "tcn_layer_fwd_tc")

        def bwd():
            i = it[0] % nbuf
            it[0] += 1
            a = _lib.LayerBwdTcArgs()
            a.gy, a.g_rows, a.gu, a.gx = gys[i].data_ptr(), lay.rows, gus[i].data_ptr(), gxs[i].data_ptr()
            a.masks = masks[i].data_ptr()
            a.w2t_hi, a.w2t_lo, a.w1t_hi, a.w1t_lo = w2th.data_ptr(), w2tl.data_ptr(), w1th.data_ptr(), w1tl.data_ptr()
            a.meta, a.nblk, a.channels = lay.meta.data_ptr(), lay.nblk, C
            for k, s_ in enumerate(shifts):
                a.shift[k] = s_
            a.drop_p = p_drop
            _lib.check(lib.tcn_layer_bwd_tc(Ct.byref(a), _lib.stream_ptr()), "tcn_layer_bwd_tc")

        def wg():
            i = it[0] % nbuf
            it[0] += 1
            ops.wgrad_tc_layer_pair(gus[i], xs[i], gys[i], hs[i], lay, shifts, gw1, gb1, gw2, gb2, drop_p=p_drop, seed=7,
                                    stream_id=4)

        frames = sum(lens)
        out = {}
        for name, fn, byt in (("layer_fwd_tc_kernel", fwd, 12 * C + 16), ("layer_bwd_tc_kernel", bwd, 12 * C + 16),
                              ("wgrad_tc_pair_kernel", wg, 16 * C)):
            ms = _graph_time(fn)
            out[name] = {"frames_per_launch": frames, "ms_per_launch": ms, "alg_bytes_per_frame": byt,
                         "achieved": byt * frames / (ms * 1e-3) / 1e9}
            out[name]["frac"] = out[name]["achieved"] / peak
        return out

    here = run([lengths[v] for v in batch], 8)       # fwd() fills the masks bwd() reads
    stress = run([8000] * 64, 6)
    what = {"layer_fwd_tc_kernel": "fused residual layer forward: dilated conv + ReLU + 1x1 conv + dropout + residual, "
                                   "both GEMMs with A from tensor memory",
            "layer_bwd_tc_kernel": "fused residual layer input gradient: gu recomputed per tap in tensor memory",
            "wgrad_tc_pair_kernel": "both weight gradients of a residual layer, contraction over frames on tcgen05"}
    dom = max(here, key=lambda k: here[k]["ms_per_launch"])   # each runs once per layer: largest time per step
    traffic = None   # dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, from the committed ncu capture
    tpath = os.path.join(ROOT, "profiles", "r1_roofline_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        if tj.get("kernel") == dom:
            traffic = tj.get("traffic_bytes_per_launch")
    kernels = {k: {"what": what[k], "step_shape": here[k], "stress_shape": stress[k]} for k in here}
    return {"bound": "hbm", "kernel": dom + " (" + what[dom] + ")",
            "achieved": here[dom]["achieved"], "peak": peak, "peak_source": src, "unit": "GB/s",
            "frac": here[dom]["frac"], "traffic": traffic,
            "frames_per_launch": here[dom]["frames_per_launch"], "ms_per_launch": here[dom]["ms_per_launch"],
            "alg_bytes_per_frame": here[dom]["alg_bytes_per_frame"],
            "stress_shape": {k: stress[dom][k] for k in ("frames_per_launch", "ms_per_launch", "achieved", "frac")},
            "kernels": kernels,
            "note": "dominant = largest time per step among the three per-layer kernels (41 launches each); at this "
                    "step's batch (a few MB per activation, L2-resident, one tile per CTA) every kernel is latency-bound; "
                    "the stress shape (64 x 8000 frames, 131 MB per activation) is where HBM binds"}


if __name__ == "__main__":
    main()
