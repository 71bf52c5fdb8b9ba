#!/usr/bin/env python
"""bench.py -- temporal-head train frames/s (BASELINE.json metric) on N B200s of one node.

Default workload (BASELINE.json configs[1], `--config cfg2`): one 5-fold training epoch schedule over 45 synthetic
CholecT45-shaped videos (seeded ragged lengths 900..3600 frames, 2048-d fp32 features, 100/6/10/15 multi-label
heads), VideoNas(fpn, 11/10/3 layers, 64 channels), data-parallel by video.  One "step" = every rank runs forward +
loss + backward over its next batch of `--videos-per-step` videos, then one all-reduce of the flat gradient buffer
and one SGD update.

  python bench.py [--gpus N --steps K --warmup W]          this repo's CUDA path (cfg2)
  python bench.py --impl reference ...                     the reference's own modules on the host cores (CPU arm)
  python bench.py --config cfg1|cfg2_512|cfg3|cfg4|cfg5 .. the other BASELINE configs (1 GPU; tools for DESIGN.md)
Prints ONE JSON line on rank 0.

Timed region: the driver's K steps are repeated R times back to back (`inner_repeats`) so that the region lasts
>= 1 s; ms_per_step = region / (K R).  Clocks are sampled through NVML for the whole process and reported for the
samples that fall inside the timed regions.
"""
from __future__ import annotations

import argparse
import contextlib
import json
import math
import os
import sys
import threading
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
HEADS = (100, 6, 10, 15)
LAYERS = (11, 10, 3)
METRIC = "temporal-head train frames/s"


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


def make_video(v, T, D, pinned):
    g = torch.Generator().manual_seed(1000 + v)
    x = torch.randn(T, D, generator=g)
    lab = (torch.rand(T, 132, generator=g) < 0.05).to(torch.uint8)
    lab[:, 131] = 0
    if pinned:
        x, lab = x.pin_memory(), lab.pin_memory()
    return x, lab


def model_args(**kw):
    a = dict(fpn=True, output=False, feature=False, trans=False, mask=False, hier=False)
    a.update(kw)
    return types.SimpleNamespace(**a)


def build_model(device, C, D):
    from computervision_codes_b200.tcn import VideoNas

    torch.manual_seed(0)
    return VideoNas(model_args(), *LAYERS, C, D, HEADS[0]).to(device).train()


def device_for_rank(local_rank, world):
    """Rank -> GPU.  On the 8-GPU box GPUs 0-3 and 4-7 hang off two host bridges (tools/h2d_concurrent.py,
    profiles/r2_h2d_concurrent_8gpu.json: 2 GPUs copy at 54 GB/s each, 4 GPUs of ONE bridge at 29 GB/s each), and the
    end-to-end arm is bound by pinned host -> device copies: with fewer ranks than GPUs the ranks are spread over both
    bridges (0, 4, 1, 5, ...) instead of filling GPUs 0..N-1."""
    n = torch.cuda.device_count()
    if n == 8 and 1 < world < 8 and os.environ.get("BENCH_NO_DEVICE_MAP") is None:
        order = [0, 4, 1, 5, 2, 6, 3, 7]
        return order[local_rank], "ranks spread over both host bridges: " + str(order[:world])
    return local_rank, "rank r -> GPU r"


def h2d_cap(world, bytes_per_frame):
    """Ceiling of the end-to-end arm from the measured concurrent pinned-copy bandwidth of the 8-GPU box."""
    p = os.path.join(ROOT, "profiles", "r2_h2d_concurrent_8gpu.json")
    if not os.path.exists(p):
        return None
    t = json.load(open(p))["concurrent_h2d"]
    if world == 8:
        per = min(t["8"]["per_gpu_GBps"])      # every step ends with an all-reduce: the slowest link sets the pace
        why = "8 ranks: GPUs 0-3 share one host bridge (23.7 GB/s each), GPUs 4-7 the other (36 GB/s each)"
    elif world == 4 and torch.cuda.device_count() != 8:
        per = min(t["4"]["per_gpu_GBps"])
        why = "4 ranks on one host bridge (28.9 GB/s each)"
    else:
        per = t["1"]["per_gpu_GBps"][0]
        why = "at most two ranks per host bridge: the full x16 link each (54 GB/s)"
    return {"frames_per_s": world * per * 1e9 / bytes_per_frame, "per_gpu_GBps": per, "why": why,
            "source": "profiles/r2_h2d_concurrent_8gpu.json (tools/h2d_concurrent.py: pinned copies of one step's inputs, all ranks at once)"}


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """NVML poller (thread) over the whole process; `with sampler.window():` marks the timed regions."""

    REASONS = (("hw_slowdown", 0x8), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40),
               ("sw_power_cap", 0x4))

    def __init__(self, gpu_index, period=0.02):
        self.samples, self.windows, self.err = [], [], None
        self._stop = threading.Event()
        self._t0 = None
        try:
            import pynvml

            pynvml.nvmlInit()
            # NVML enumerates physical devices: map the CUDA ordinal through CUDA_VISIBLE_DEVICES when it is a list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = gpu_index
            if vis:
                ids = [s.strip() for s in vis.split(",") if s.strip()]
                if gpu_index < len(ids) and ids[gpu_index].isdigit():
                    phys = int(ids[gpu_index])
            self._nv, self._h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
            self._thr = threading.Thread(target=self._run, args=(period,), daemon=True)
            self._thr.start()
        except Exception as e:  # noqa: BLE001
            self.err, self._thr = repr(e), None

    def _run(self, period):
        nv, h = self._nv, self._h
        while not self._stop.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.samples.append((time.time(), float(mhz), int(rs)))
            except Exception as e:  # noqa: BLE001
                self.err = repr(e)
                return
            self._stop.wait(period)

    @contextlib.contextmanager
    def window(self):
        t0 = time.time()
        try:
            yield
        finally:
            self.windows.append((t0, time.time()))

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [f"nvml unavailable: {self.err}"], "samples": 0}
        inside = [s for s in self.samples if any(a <= s[0] <= b for a, b in self.windows)]
        use = inside or self.samples
        bits = 0
        for s in use:
            bits |= s[2]
        reasons = [n for n, m in self.REASONS if bits & m]
        return {"sm_mhz": float(np.median([s[1] for s in use])), "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(inside), "samples_process": len(self.samples),
                "timed_window_s": round(sum(b - a for a, b in self.windows), 3)}


# --------------------------------------------------------------------------------------------- reference arms
def _ref_tenco():
    """The reference's own Temporal_tenco/network.py (from /root/reference here, from baseline/_ref on the GPU box)."""
    from oracle import ref_import

    if not ref_import.available():
        return None
    return ref_import.tenco_network()


def _quiet(fn, *args, **kw):
    """Run fn with stdout captured: the reference's BaseCausalTCN.__init__ prints its layer count (network.py:111), and
    this program's stdout carries exactly one JSON line."""
    import io

    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*args, **kw)


def _ref_loss(outs, labels):
    """train_loop's loss (Temporal_tenco/run.py:190-212): BCEWithLogitsLoss (mean over T x K) of sample 0, summed over
    the four FPN levels, 0.1 (i + v + t) + ivt.  labels = (y_i, y_v, y_t, y_ivt) float (T, K)."""
    bce = torch.nn.BCEWithLogitsLoss()
    o, oi, ov, ot = outs[:4]

    def head(lst, y):
        return sum(bce(q[0].transpose(0, 1), y) for q in lst)

    return 0.1 * (head(oi, labels[0]) + head(ov, labels[1]) + head(ot, labels[2])) + head(o, labels[3])


def _ref_mask(x_btd):
    """The --mask branch of VideoNas.forward (network.py:43-48), restated on the tensor's own device: 25 % of the
    flattened (B, D, T) input zeroed by a random permutation."""
    n = x_btd.numel()
    num_mask = int(n * 0.75)
    mask = torch.cat((torch.zeros(n - num_mask), torch.ones(num_mask)))
    mask = mask[torch.randperm(n)]
    B, T, D = x_btd.shape
    return mask.view(B, D, T).permute(0, 2, 1).to(x_btd.device)


def _split_labels(lab_u8):
    y = lab_u8[:, :131].float()
    return (y[:, 100:106], y[:, 106:116], y[:, 116:131], y[:, 0:100])


def cpu_arm(steps, V, C, D, budget_s=150.0, sweep_budget_s=20.0):
    """Reference arm: the reference's VideoNas (unmodified module file) under PyTorch CPU on the host cores, the same
    train step the reference's train_loop runs per video: --mask input masking, forward in train mode, BCE loss over
    4 heads x 4 levels, backward, SGD(lr 1e-2, wd 1e-5).  One bench step = V videos (the reference steps per video).
    Thread count swept first on one video; then `steps` steps timed (bounded by budget_s)."""
    net = _ref_tenco()
    passes, lengths = fold_schedule()
    kind = "reference"
    if net is not None:
        torch.manual_seed(0)
        model = _quiet(net.VideoNas, model_args(), *LAYERS, C, D, HEADS[0]).train()
        opt = torch.optim.SGD(model.parameters(), lr=1e-2, weight_decay=1e-5)

        def one(x, labels):
            xin = x * _ref_mask(x)     # args.mask branch (network.py:43-48); its .cuda() cannot run on a CPU model
            outs = model(xin, False)
            loss = _ref_loss(outs, labels)
            for p in model.parameters():
                p.grad = None
            loss.backward()
            opt.step()
            return float(loss)
    else:  # baseline/_ref missing: the oracle's torch port of the same step
        from oracle import torch_port as P

        kind = "port"
        m = build_model("cpu", C, D)
        params = {k: v.detach().clone().requires_grad_(True) for k, v in m.state_dict().items()}

        def one(x, labels):
            for p in params.values():
                p.grad = None
            loss = P.train_step_loss(x * _ref_mask(x), params, labels, train=True)
            loss.backward()
            return float(loss)

    def video(i):
        v = passes[i % len(passes)]
        x, lab = make_video(v, lengths[v], D, pinned=False)
        return x.unsqueeze(0), _split_labels(lab), lengths[v]

    ncpu = os.cpu_count() or 1
    cands = sorted({n for n in (1, 2, 4, 8, 16, 32, 64, ncpu) if n <= ncpu}, reverse=True)
    x0, l0, T0 = video(0)
    best_n, best_fps = cands[0], 0.0
    t_sweep = time.time()
    for n in cands:
        torch.set_num_threads(n)
        one(x0, l0)
        t0 = time.time()
        one(x0, l0)
        fps = T0 / (time.time() - t0)
        if fps > best_fps:
            best_n, best_fps = n, fps
        if time.time() - t_sweep > sweep_budget_s:
            break
    torch.set_num_threads(best_n)
    frames, nvid, done_steps = 0, 0, 0
    t0 = time.time()
    for s in range(steps):
        for j in range(V):
            x, labels, T = video(s * V + j)
            one(x, labels)
            frames += T
            nvid += 1
        done_steps += 1
        if time.time() - t0 > budget_s:
            break
    dt = time.time() - t0
    return {"value": frames / dt, "unit": "frames/s", "cores": best_n, "kind": kind,
            "sample": f"{done_steps} of {steps} steps x {V} videos ({nvid} videos, {frames} frames) of the same schedule; "
                      f"per video: --mask masking + forward (train mode) + BCE loss + backward + SGD, "
                      f"{'reference Temporal_tenco/network.py VideoNas' if kind == 'reference' else 'oracle torch port'} "
                      f"on torch {torch.__version__} CPU; threads swept over {cands} on one video, best {best_n}; "
                      f"host has {ncpu} cpus",
            "ms_per_step": dt / max(1, done_steps) * 1e3}


def eager_gpu_baseline(dev, C, D, nvid=12):
    """The practical bar: the reference's own VideoNas in PyTorch eager on this B200, one
    video per step as train_loop runs it (mask + forward + loss + backward + SGD), inputs resident, TF32 off / on."""
    net = _ref_tenco()
    if net is None:
        return {"unavailable": "baseline/_ref missing (run oracle/vendor_ref.py where /root/reference exists)"}
    passes, lengths = fold_schedule()
    vids = [passes[i] for i in range(nvid)]
    data = []
    for v in vids:
        x, lab = make_video(v, lengths[v], D, pinned=False)
        data.append((x.unsqueeze(0).to(dev), tuple(t.to(dev) for t in _split_labels(lab)), lengths[v]))
    out = {}
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    try:
        for name, tf32 in (("tf32_off", False), ("tf32_on", True)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.manual_seed(0)
            model = _quiet(net.VideoNas, model_args(mask=True), *LAYERS, C, D, HEADS[0]).to(dev).train()
            opt = torch.optim.SGD(model.parameters(), lr=1e-2, weight_decay=1e-5)

            def one(x, labels):
                outs = model(x, True)
                loss = _ref_loss(outs, labels)
                for p in model.parameters():
                    p.grad = None
                loss.backward()
                opt.step()

            for x, labels, _ in data[:3]:
                one(x, labels)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for x, labels, _ in data:
                one(x, labels)
            e1.record()
            torch.cuda.synchronize()
            out[name] = sum(d[2] for d in data) / (e0.elapsed_time(e1) * 1e-3)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    out["unit"] = "frames/s"
    out["what"] = (f"reference Temporal_tenco/network.py VideoNas(fpn, 11/10/3, C={C}, D={D}) in PyTorch "
                   f"{torch.__version__} eager on this GPU, batch 1 (one video per step as run.py does), {nvid} videos, "
                   "--mask + forward + loss + backward + SGD, inputs resident")
    return out


# --------------------------------------------------------------------------------------------- cfg2 (default)
def run_cfg2(a, rank, world, local_rank, C, D, tag):
    V = a.videos_per_step
    config = {"workload": f"{tag}: 5-fold epoch schedule over 45 synthetic CholecT45-shaped videos "
                          f"(ragged 900..3600 frames, seed 45), VideoNas(fpn, 11/10/3, C={C}, D={D}, heads 100/6/10/15), "
                          "tenco BCE loss, --mask input masking, SGD(lr 1e-2, wd 1e-5)",
              "videos_per_rank_per_step": V,
              "global_batch_videos": V * world,
              "global_batch_note": "the schedule (155 passes; 31 training videos per fold) is cycled: a global step takes "
                                   "the next V x N passes, so at N = 8 one step spans two folds' worth of videos",
              "parallelism": f"dp{world}-by-video, videos of a global step assigned longest-first (LPT), equal count per rank",
              "l2": "inputs cycle through the 0.8 GB feature set (>> 126 MB L2); no explicit flush",
              "resident_inputs": "one HBM feature arena (FeatureCache.pack), read in place by the step",
              "launch_mode": "one CUDA graph per step; programmatic dependent launch inside the graph with "
                             "griddepcontrol.wait and NO early trigger (DESIGN.md section 3: the early trigger, worth "
                             "4.6 % of the step, let a layer read frames its predecessor had not stored yet)"}

    if a.impl == "reference":
        if rank != 0:
            return
        cb = cpu_arm(a.steps, V, C, D)
        line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "frames/s",
                "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": cb["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config, "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": cb["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    assert torch.cuda.is_available(), "bench.py needs a GPU (use --impl reference for the CPU arm)"
    dev_index, dev_map = device_for_rank(local_rank, world)
    config["device_map"] = dev_map
    torch.cuda.set_device(dev_index)
    dev = torch.device("cuda", dev_index)
    clocks = ClockSampler(dev_index)
    pg = None
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
        pg = torch.distributed.group.WORLD

    from computervision_codes_b200 import _lib
    from computervision_codes_b200.data import FeatureCache
    from computervision_codes_b200.trainer import TemporalTrainer, lpt_assign

    _lib.load()
    model = build_model(dev, C, D)
    passes, lengths = fold_schedule()
    # capacity: V videos per rank, or up to Vcap when the end-to-end arm shards by host-link bandwidth (N > 1, below)
    Vcap = V if world == 1 else V + V // 2 + 1
    max_frames = max(lengths.values()) * Vcap
    trainer = TemporalTrainer(model, lr=1e-2, weight_decay=1e-5, process_group=pg, world_size=world,
                              max_frames=max_frames, max_seqs=max(Vcap, 1), use_graph=not a.no_graph,
                              input_mask_p=0.25, seed=1234 + rank)  # --mask of Scripts/train_fold1.sh:28

    # Global step s takes the next V * world passes of the schedule (cycled) and assigns them to the ranks
    # longest-first, V videos each (length-balanced DP by video).  Every rank computes the same table.
    _bcache = {}

    def batch_at(s):
        b = _bcache.get(s)
        if b is None:
            vids = [passes[(s * V * world + j) % len(passes)] for j in range(V * world)]
            shard = lpt_assign([lengths[v] for v in vids], world, cap=V)[rank]
            b = _bcache[s] = [vids[i] for i in shard]
        return b

    host = {v: make_video(v, lengths[v], D, pinned=True) for v in sorted(set(passes))}
    # resident arm: the feature set lives in HBM as one arena (data.FeatureCache.pack: SURVEY 8(f1)); a step reads its
    # videos in place -- the block table carries their positions -- so nothing is copied inside the timed region
    cache = FeatureCache(dev)
    for v, (x, lab) in host.items():
        cache.add_packed(v, x, lab)
    cache.pack()

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if world > 1:
            torch.distributed.all_reduce(t)
        return float(t.item())

    def region(run_step, s0, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with clocks.window():
            e0.record()
            for s in range(s0, s0 + n):
                run_step(s)
            e1.record()
            barrier()
        return allmax(e0.elapsed_time(e1))

    def timed(run_step, batch_fn):
        """W warm-up steps, K probe steps (also warm-up) that size the repeat count R, then K * R timed steps."""
        K, W = a.steps, a.warmup
        for s in range(W):
            run_step(s)
        probe_ms = region(run_step, W, K)
        # 5 % head-room: the probe region carries the tail of the warm-up and runs a little slower than the timed one
        R = max(1, min(int(math.ceil(1.05 * a.min_seconds * 1e3 / max(probe_ms, 1e-3))), 2000))
        s0 = W + K
        ms = region(run_step, s0, K * R)
        frames = allsum(sum(lengths[v] for s in range(s0, s0 + K * R) for v in batch_fn(s)))
        return ms, R, frames

    # ---- arm 1: inputs resident in HBM (the step starts from device tensors)
    def step_resident(s):
        trainer.step_cached(cache, [(v, 0, lengths[v]) for v in batch_at(s)])

    ms_total, R, frames_all = timed(step_resident, batch_at)
    value = frames_all / (ms_total * 1e-3)

    # ---- arm 2: end to end through the public API with HOST (pinned) inputs: H2D of this step's
    # features + labels and D2H of the loss vector inside the timed region, every step
    loss_host = torch.empty(8, dtype=torch.float32).pin_memory()
    h2d = [0]
    staged = [None]
    # N > 1: the ranks' host links differ (two host bridges, tools/h2d_concurrent.py) and every step ends with an
    # all-reduce, so the slowest link would set the pace.  The end-to-end arm measures each rank's pinned-copy bandwidth
    # with all ranks copying at once, shards the videos of a global step in proportion to it (weighted LPT, at most Vcap
    # per rank) and normalises the loss by the GLOBAL video count, so that the summed gradient is still the mean over
    # all V x N videos of the step.  The resident arm keeps equal shares.
    link_bw = None
    if world > 1:
        probe = sorted(host)[:8]
        dst = [torch.empty_like(host[v][0], device=dev) for v in probe]
        torch.cuda.synchronize()
        for rep in range(3):
            if rep == 1:
                barrier()
                t0 = time.perf_counter()
            for v, d_ in zip(probe, dst):
                d_.copy_(host[v][0], non_blocking=True)
            torch.cuda.synchronize()
        mine = 2 * sum(host[v][0].numel() * 4 for v in probe) / (time.perf_counter() - t0) / 1e9
        del dst
        t = torch.zeros(world, device=dev, dtype=torch.float64)
        t[rank] = mine
        torch.distributed.all_reduce(t)
        link_bw = [float(x) for x in t.tolist()]
        config["e2e_sharding"] = {"pinned_copy_GBps_per_rank_all_ranks_copying": [round(x, 1) for x in link_bw],
                                  "rule": "videos of a global step assigned longest-first in proportion to the rank's "
                                          f"link bandwidth (at most {Vcap} per rank); loss normalised by the global video count"}
    _ecache = {}

    def batch_e2e(s):
        if link_bw is None:
            return batch_at(s)
        b = _ecache.get(s)
        if b is None:
            vids = [passes[(s * V * world + j) % len(passes)] for j in range(V * world)]
            shard = lpt_assign([lengths[v] for v in vids], world, cap=Vcap, weights=link_bw)[rank]
            b = _ecache[s] = [vids[i] for i in shard]
        return b

    gseq = None if link_bw is None else V * world

    def stage(s):
        b = batch_e2e(s)
        xs, ls = [host[v][0] for v in b], [host[v][1] for v in b]
        h2d[0] = sum(t.numel() * t.element_size() for t in xs + ls)
        trainer.prefetch(xs, ls, [lengths[v] for v in b])
        staged[0] = s

    def step_e2e(s):
        # software pipeline over the public API: the H2D of the next step's inputs (pinned host -> device, copy
        # stream) is issued before this step's kernels; every step still pays its own H2D and its own D2H read.
        if trainer._prefetched is None or staged[0] != s:
            trainer._prefetched = None
            stage(s)
        out = trainer.step(global_seqs=gseq)
        stage(s + 1)
        loss_host.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    ms_e2e, R_e2e, frames_e2e = timed(step_e2e, batch_e2e)
    trainer._prefetched = None
    torch.cuda.synchronize()
    e2e_value = frames_e2e / (ms_e2e * 1e-3)

    # ---- roofline of the dominant kernel, timed live with CUDA events
    roof = None
    if C == 64:
        with clocks.window():
            roof = measure_layer_roofline(model, lengths, batch_at(a.warmup), dev, C)
    eager = None
    if rank == 0 and world == 1 and not a.no_eager_baseline:
        try:
            eager = eager_gpu_baseline(dev, C, D)
        except Exception as e:  # noqa: BLE001
            eager = {"unavailable": repr(e)[:200]}
    ck = clocks.stop()
    launches = trainer.launches_per_step() * a.steps * R

    # graphs hold NCCL work: drop them before the process group; a watchdog (daemon) only fires if teardown hangs
    trainer.close()
    torch.cuda.synchronize()
    if world > 1:
        wd = threading.Timer(60.0, lambda: os._exit(0))
        wd.daemon = True
        wd.start()
        try:
            torch.distributed.barrier()
            torch.distributed.destroy_process_group()
        except Exception:  # noqa: BLE001
            pass
        wd.cancel()
    if rank != 0:
        return
    cb = None if a.no_cpu_baseline or world > 1 else cpu_arm(min(a.steps, 3), V, C, D, budget_s=25.0, sweep_budget_s=10.0)
    line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "inner_repeats": R, "timed_region_s": ms_total * 1e-3,
            "ms_per_step": ms_total / (a.steps * R), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32 (3xTF32 tensor-core products, fp32 accumulate)",
            "data": "synthetic", "config": config, "clocks": ck,
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d[0], "d2h_bytes_per_step": 32,
                    "ms_per_step": ms_e2e / (a.steps * R_e2e), "inner_repeats": R_e2e, "timed_region_s": ms_e2e * 1e-3,
                    "host_link_cap": h2d_cap(world, D * 4 + 132)},
            "gpu_launches": launches, "gpu_launches_per_step": trainer.launches_per_step(), "roofline": roof}
    if eager is not None:
        line["eager_gpu_baseline"] = eager
    if cb is not None:
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line))
    sys.stdout.flush()


# --------------------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="cfg2", choices=["cfg1", "cfg2", "cfg2_512", "cfg3", "cfg4", "cfg5"])
    ap.add_argument("--videos-per-step", type=int, default=8, help="videos per rank per step (ragged batch)")
    ap.add_argument("--min-seconds", type=float, default=1.0, help="minimum length of each timed region")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="plain stream launches instead of a CUDA graph")
    a = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if a.config == "cfg2":
        run_cfg2(a, rank, world, local_rank, 64, 2048, "cfg2")
    elif a.config == "cfg2_512":
        run_cfg2(a, rank, world, local_rank, 512, 512, "cfg2 at the reference scripts' own width (--embed_num 512, "
                 "--input_dim 512: Temporal_tenco/run.py:89,313, Scripts/train_fold1.sh:28)")
    else:
        import bench_configs

        if rank == 0:
            bench_configs.run(a)


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


def measure_layer_roofline(model, lengths, batch, dev, C):
    """Roofline of the step's dominant kernel, timed live with CUDA events (graph of 24 launches on the current stream).

    A train step is 41 residual layers x three launches -- `layer_fwd_tc_kernel` (fused forward), `layer_bwd_tc_kernel`
    (fused input gradient) and the weight-gradient kernel of the layer -- plus a few dozen one-off launches.  All three
    are timed here on this step's batch shape and on the TERL stress shape (64 x 8000 frames, where the HBM roofline is
    the binding limit); the one with the largest time per step is the dominant kernel and fills `roofline`, the others
    are listed under `roofline.kernels`.  Two byte counts per frame are reported (C = 64, fp32):
      `alg_bytes_per_frame`  what this design's launch must move:
          layer_fwd_tc  12*C + 16  read x, write h, write y (+ 16 B of ReLU / dropout bit words)
          layer_bwd_tc  12*C + 16  read gy, write gu, write gx (+ the bit words)
          wgrad         16*C       read gu, x, gy, h (the 64 x 256 weight-gradient tile is negligible)
      `sec8d_bytes_per_frame`  SURVEY 8(d)'s accounting: forward 8*C (read x, write y; h recomputed in backward), the whole
          backward 12*C (read x, read gy, write gx) -- charged here to the two backward launches TOGETHER, so
          `sec8d_frac_backward` = 12*C * frames / (t_bwd + t_wgrad) / peak."""
    import ctypes as Ct

    from computervision_codes_b200 import _lib, ops
    from computervision_codes_b200.layout import SeqLayout

    peak, src = hbm_peak()
    lib = _lib.load()
    layer = model.PG.layers[4]
    d = int(layer.conv_dilated.dilation[0])
    shifts = (-2 * d, -d, 0)
    w1, w2 = layer.conv_dilated.weight.detach(), layer.conv_1x1.weight.detach()
    b1, b2 = layer.conv_dilated.bias.detach().contiguous(), layer.conv_1x1.bias.detach().contiguous()
    w1h, w1l = ops.split_weight(w1)
    w2h, w2l = ops.split_weight(w2)
    w1th, w1tl = ops.split_weight(w1, transpose=True)
    w2th, w2tl = ops.split_weight(w2, transpose=True)
    p_drop = 0.5
    wg_name = ops.layer_wgrad_kernel_name()

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
            q = _lib.LayerFwdTcArgs()
            q.x, q.x_rows, q.y, q.h = xs[i].data_ptr(), lay.rows, ys[i].data_ptr(), hs[i].data_ptr()
            q.w1_hi, q.w1_lo, q.w2_hi, q.w2_lo = w1h.data_ptr(), w1l.data_ptr(), w2h.data_ptr(), w2l.data_ptr()
            q.b1, q.b2 = b1.data_ptr(), b2.data_ptr()
            q.meta, q.nblk, q.channels = lay.meta.data_ptr(), lay.nblk, C
            for k, s_ in enumerate(shifts):
                q.shift[k] = s_
            q.drop_p, q.drop_seed, q.drop_stream = p_drop, 7, 4
            q.masks = masks[i].data_ptr()
            _lib.check(lib.tcn_layer_fwd_tc(Ct.byref(q), _lib.stream_ptr()), "tcn_layer_fwd_tc")

        def bwd():
            i = it[0] % nbuf
            it[0] += 1
            q = _lib.LayerBwdTcArgs()
            q.gy, q.g_rows, q.gu, q.gx = gys[i].data_ptr(), lay.rows, gus[i].data_ptr(), gxs[i].data_ptr()
            q.masks = masks[i].data_ptr()
            q.w2t_hi, q.w2t_lo, q.w1t_hi, q.w1t_lo = w2th.data_ptr(), w2tl.data_ptr(), w1th.data_ptr(), w1tl.data_ptr()
            q.meta, q.nblk, q.channels = lay.meta.data_ptr(), lay.nblk, C
            for k, s_ in enumerate(shifts):
                q.shift[k] = s_
            q.drop_p = p_drop
            _lib.check(lib.tcn_layer_bwd_tc(Ct.byref(q), _lib.stream_ptr()), "tcn_layer_bwd_tc")

        def wg():
            i = it[0] % nbuf
            it[0] += 1
            ops.layer_wgrad(gus[i], xs[i], gys[i], hs[i], lay, shifts, gw1, gb1, gw2, gb2, drop_p=p_drop, seed=7,
                            stream_id=4, masks=masks[i], skip_reduce=True)

        frames = sum(lens)
        out = {}
        for name, fn, byt, s8 in (("layer_fwd_tc_kernel", fwd, 12 * C + 16, 8 * C),
                                  ("layer_bwd_tc_kernel", bwd, 12 * C + 16, None), (wg_name, wg, 16 * C, None)):
            ms = _graph_time(fn)
            out[name] = {"frames_per_launch": frames, "ms_per_launch": ms, "alg_bytes_per_frame": byt,
                         "achieved": byt * frames / (ms * 1e-3) / 1e9}
            out[name]["frac"] = out[name]["achieved"] / peak
            if s8 is not None:
                out[name]["sec8d_bytes_per_frame"] = s8
                out[name]["sec8d_frac"] = s8 * frames / (ms * 1e-3) / 1e9 / peak
        t_b = out["layer_bwd_tc_kernel"]["ms_per_launch"] + out[wg_name]["ms_per_launch"]
        for k in ("layer_bwd_tc_kernel", wg_name):
            out[k]["sec8d_bytes_per_frame_backward_total"] = 12 * C
            out[k]["sec8d_frac_backward"] = 12 * C * frames / (t_b * 1e-3) / 1e9 / peak
        return out

    here = run([lengths[v] for v in batch], 8)       # fwd() fills the masks bwd() reads
    stress = run([8000] * 64, 6)
    what = {"layer_fwd_tc_kernel": "fused residual layer forward: dilated conv + ReLU + 1x1 conv + dropout + residual, "
                                   "both GEMMs with A from tensor memory",
            "layer_bwd_tc_kernel": "fused residual layer input gradient: gu recomputed per tap in tensor memory",
            wg_name: "all four weight-gradient products of a residual layer (gW1 over three taps, gW2) + both bias "
                     "gradients, contraction over frames on tcgen05"}
    # each of the three runs once per residual layer (the executor groups several layers' weight gradients into one launch
    # with proportionally fewer CTAs each): the largest time per layer is the largest time per step
    dom = max(here, key=lambda k: here[k]["ms_per_launch"])
    traffic = None   # dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, from the committed ncu capture
    for tname in ("r2_roofline_traffic.json", "r1_roofline_traffic.json"):
        tpath = os.path.join(ROOT, "profiles", tname)
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            ent = tj.get("kernels", {}).get(dom) or (tj if tj.get("kernel") == dom else None)
            if ent:
                traffic = ent.get("traffic_bytes_per_launch_step_shape", ent.get("traffic_bytes_per_launch"))
                break
    kernels = {k: {"what": what[k], "step_shape": here[k], "stress_shape": stress[k]} for k in here}
    return {"bound": "hbm", "kernel": dom + " (" + what[dom] + ")",
            "achieved": here[dom]["achieved"], "peak": peak, "peak_source": src, "unit": "GB/s",
            "frac": here[dom]["frac"], "traffic": traffic,
            "frames_per_launch": here[dom]["frames_per_launch"], "ms_per_launch": here[dom]["ms_per_launch"],
            "alg_bytes_per_frame": here[dom]["alg_bytes_per_frame"],
            "sec8d": {k: v for k, v in here[dom].items() if k.startswith("sec8d")},
            "stress_shape": {k: v for k, v in stress[dom].items() if k != "alg_bytes_per_frame"},
            "kernels": kernels,
            "note": "dominant = largest time per step among the three per-layer kernels (41 launches each); at this "
                    "step's batch (a few MB per activation, L2-resident, one tile per CTA) every kernel is latency-bound; "
                    "the stress shape (64 x 8000 frames, 131 MB per activation) is where HBM binds; alg_bytes = this "
                    "design's per-launch bytes, sec8d = SURVEY 8(d)'s accounting (8C forward, 12C for the whole backward)"}


if __name__ == "__main__":
    main()
