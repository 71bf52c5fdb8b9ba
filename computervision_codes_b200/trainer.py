"""Temporal-student train step (forward + loss + backward [+ all-reduce] + SGD) over a ragged batch of
videos -- the hot loop of MT4MTLKD/Temporal_tenco/run.py:182-235 (TERL/0_5fold_TCN_black/run.py:297-366).

Parameters that receive a gradient live in one flat fp32 buffer with a matching flat gradient buffer,
so data-parallel training needs exactly one all-reduce per step (torch.distributed / NCCL over
NVLink) and one fused SGD launch.  Videos shard across ranks; nothing else crosses GPUs.
"""
from __future__ import annotations

import torch

from . import losses, ops
from .layout import SeqLayout

# Parameters the reference leaves without a gradient on the --fpn path: they stay
# outside the flat trainable buffer, their .grad stays None, weight decay never touches them.
_NO_GRAD_FPN = ("PG.conv_out.", ".conv_1x1.", ".conv_out.", "fpn.latlayer2.", "fpn.latlayer3.")


def _is_trainable(name: str) -> bool:
    if name.startswith("PG.conv_1x1.") or ".layers." in name:
        return True
    return not any(tok in name for tok in _NO_GRAD_FPN)


class TemporalTrainerEager:
    """Reference composition through the per-op autograd Functions (one Python call per kernel):
    slow, kept as the cross-check of the native executor."""

    def __init__(self, model, lr=1e-2, weight_decay=1e-5, loss_type="all", terl_pos_weight=False,
                 process_group=None, world_size=1):
        self.model = model
        self.lr, self.weight_decay = lr, weight_decay
        self.loss_type, self.terl_pos_weight = loss_type, terl_pos_weight
        self.pg, self.world = process_group, world_size
        named = [(n, p) for n, p in model.named_parameters() if _is_trainable(n)]
        self.names = [n for n, _ in named]
        total = sum(p.numel() for _, p in named)
        dev = named[0][1].device
        self.flat_p = torch.empty(total, device=dev, dtype=torch.float32)
        self.flat_g = torch.zeros(total, device=dev, dtype=torch.float32)
        off = 0
        for _, p in named:
            n = p.numel()
            self.flat_p[off:off + n].copy_(p.data.reshape(-1))
            p.data = self.flat_p[off:off + n].view(p.shape)
            p.grad = self.flat_g[off:off + n].view(p.shape)
            off += n
        self.num_params = total

    def forward_backward(self, x_rows, labels_u8, lengths):
        """x_rows: (frames, D) fp32 device; labels_u8: (frames, >=131) uint8 device, columns ivt|i|v|t.
        Returns the device tensor (loss, loss_i, loss_v, loss_t, loss_ivt)."""
        lay = SeqLayout.get(lengths, x_rows.device)
        self.flat_g.zero_()
        f_rows, logit_rows = self.model.forward_packed(x_rows, lay)
        total, li, lv, lt, livt = losses.tenco_loss(logit_rows, labels_u8, lay, self.model.head_sizes,
                                                    self.loss_type, self.terl_pos_weight)
        total.backward()
        return torch.stack([total.detach(), li, lv, lt, livt])

    def step(self, x_rows, labels_u8, lengths):
        out = self.forward_backward(x_rows, labels_u8, lengths)
        if self.world > 1:
            torch.distributed.all_reduce(self.flat_g, group=self.pg)
        ops.sgd_step(self.flat_p, self.flat_g, self.lr, self.weight_decay, grad_scale=1.0 / max(1, self.world))
        return out


def lpt_assign(lengths, world, cap=None, weights=None):
    """Longest-processing-time-first assignment of videos to ranks (balances frames per rank).
    cap: at most this many videos per rank (equal counts keep "mean over ranks of the per-rank mean" == mean over all
    videos).  weights: relative speed of each rank (e.g. its host-link bandwidth when the inputs come from the host): rank
    k then gets frames in proportion to weights[k]; the shares are unequal, so the step must normalise by the global
    video count (TemporalTrainer.step(..., global_seqs=n))."""
    order = sorted(range(len(lengths)), key=lambda i: -lengths[i])
    loads = [0] * world
    w = [1.0] * world if weights is None else [float(x) for x in weights]
    shards = [[] for _ in range(world)]
    for i in order:
        open_ranks = [k for k in range(world) if cap is None or len(shards[k]) < cap]
        r = min(open_ranks, key=lambda k: (loads[k] + lengths[i]) / w[k])
        shards[r].append(i)
        loads[r] += lengths[i]
    return shards


class TemporalTrainer:
    """Native path: csrc/model.cu executor, optionally replayed from one CUDA graph per trainer
    (batch shape, block table and dropout seed are device-side data, so one graph serves every batch).

    step(x_rows, labels_u8, lengths): x_rows (frames, D) fp32 and labels_u8 (frames, >=131) uint8 may be
    device tensors or pinned host tensors (then the H2D copy is part of the step, on the same stream).
    """

    def __init__(self, model, lr=1e-2, weight_decay=1e-5, loss_type="all", terl_pos_weight=False,
                 process_group=None, world_size=1, max_frames=4096, max_seqs=8, use_graph=True,
                 input_mask_p=0.0, seed=0):
        from .executor import ModelExecutor
        from .losses import TARGET_WEIGHT, TOOL_WEIGHT, VERB_WEIGHT, _LOSS_TYPE_WEIGHTS

        self.model = model
        self.lr, self.weight_decay = lr, weight_decay
        self.pg, self.world = process_group, world_size
        max_rows = max_frames + 128 * max_seqs
        self.ex = ModelExecutor(model, max_rows, max_seqs)
        pw = None
        if terl_pos_weight:
            pw = [1.0] * model.head_sizes[0] + list(TOOL_WEIGHT) + list(VERB_WEIGHT) + list(TARGET_WEIGHT)
        self.ex.set_loss(_LOSS_TYPE_WEIGHTS[loss_type], pw)
        self.ex.set_dropout(input_mask_p, model.PG.channel_dropout.p, model.PG.layers[0].dropout.p)
        dev = self.ex.device
        D = self.ex.cfg.in_dim
        self.max_frames = max_frames
        # two input slots: while the graph of step i runs out of slot i % 2, prefetch() can already stage the
        # inputs of step i + 1 into the other slot on a copy stream (H2D overlapped with compute)
        self.x_slots = [torch.zeros(max_frames, D, device=dev, dtype=torch.float32) for _ in range(2)]
        self.lab_slots = [torch.zeros(max_frames, self.ex.ld_logits, device=dev, dtype=torch.uint8) for _ in range(2)]
        self.slot = 0
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.copy_stream2 = torch.cuda.Stream(device=dev)
        self.slot_ready2 = [torch.cuda.Event(), torch.cuda.Event()]
        self.slot_free = [torch.cuda.Event(), torch.cuda.Event()]   # compute finished reading the slot
        self.slot_ready = [torch.cuda.Event(), torch.cuda.Event()]  # copy into the slot finished
        self._prefetched = None
        self.flat_p, self.flat_g = self.ex.flat_p, self.ex.flat_g
        self.num_params = self.flat_p.numel()
        self.use_graph = use_graph
        self.graphs = [None, None]
        self._outs = [None, None]
        # data parallel: every rank must start from rank 0's weights, and draws its own dropout / input-mask stream
        rank = 0
        if world_size > 1:
            rank = torch.distributed.get_rank(process_group)
            torch.distributed.broadcast(self.flat_p, src=torch.distributed.get_global_rank(process_group, 0)
                                        if process_group is not None else 0, group=process_group)
        self.rng = torch.Generator().manual_seed(int(seed) * 1000003 + rank)
        self.training = True
        self._launches_per_step = None
        self._arena_graphs = {}
        # {lr, weight_decay, grad_scale} live on the device: the captured graph follows set_lr() / an LR schedule
        self._hyper_host = torch.tensor([lr, weight_decay, 1.0 / max(1, world_size)], dtype=torch.float32).pin_memory()
        self.hyper = self._hyper_host.to(dev)
        self._global_seqs = 0

    def set_lr(self, lr: float):
        """New learning rate for the following steps (no graph re-capture).  The value travels as a kernel argument of
        a stream-ordered fill, so steps already queued keep the rate they were enqueued with."""
        self.lr = float(lr)
        self._hyper_host[0] = self.lr
        self.hyper[0:1].fill_(self.lr)

    def _set_global_norm(self, global_seqs):
        """global_seqs: number of videos of the whole data-parallel step when the ranks hold unequal shares (every rank
        divides by it, the all-reduce sums: the global mean); None = equal shares (mean per rank, then mean over ranks)."""
        g = 0 if global_seqs is None else int(global_seqs)
        if g != self._global_seqs:
            self._global_seqs = g
            self.ex.set_loss_norm(g)
            self._hyper_host[2] = 1.0 if g > 0 else 1.0 / max(1, self.world)
            self.hyper[2:3].fill_(float(self._hyper_host[2]))   # stream-ordered: the captured SGD reads it from device memory

    def step_cached(self, cache, items, global_seqs=None):
        """One step on clips / videos of a ``data.FeatureCache``: items = [(video, start, length), ...].
        A packed cache (``cache.pack()``) is read in place: no copy, the block table addresses the arena."""
        self._set_global_norm(global_seqs)
        items = list(items)
        if getattr(cache, "arena_x", None) is None:
            xs, ls, lens = cache.batch(items)
            return self.step(xs, ls, lens, global_seqs=global_seqs)
        lens, starts = [], []
        for vid, start, n in items:
            assert 0 <= start and n > 0 and start + n <= cache.frames(vid), (vid, start, n)
            lens.append(int(n))
            starts.append(cache.offset[vid] + int(start))
        lay = SeqLayout.get(lens, self.ex.device, in_starts=starts)
        seed = int(torch.randint(0, 2 ** 31 - 1, (1,), generator=self.rng).item())
        self.ex.set_batch(lay, seed)
        x, lab = cache.arena_x, cache.arena_lab
        if not self.use_graph:
            return self._counted_body(None, x, lab)
        key = (x.data_ptr(), lab.data_ptr())
        if key not in self._arena_graphs:
            snap_p = self.flat_p.clone()
            self._body(None, x, lab)
            self.flat_p.copy_(snap_p)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = self._counted_body(None, x, lab)
            self.flat_p.copy_(snap_p)
            self._arena_graphs[key] = (g, out, x, lab)
        g, out = self._arena_graphs[key][:2]
        g.replay()
        return out

    def _body(self, slot, x=None, lab=None):
        if x is None:
            x, lab = self.x_slots[slot], self.lab_slots[slot]
        out = self.ex.train_step(x, lab, training=self.training)
        if self.world > 1:
            torch.distributed.all_reduce(self.flat_g, group=self.pg)
        ops.sgd_step_dev(self.flat_p, self.flat_g, self.hyper)
        return out

    def launches_per_step(self):
        """Kernels of this library one step enqueues: counted by the library itself (tcn_launch_count) around the
        step body -- under a CUDA graph, around its capture; every replay launches the same kernels."""
        assert self._launches_per_step is not None, "launches_per_step(): run a step first"
        return self._launches_per_step

    def _counted_body(self, slot, x=None, lab=None):
        from . import _lib

        lib = _lib.load()
        n0 = lib.tcn_launch_count()
        out = self._body(slot, x, lab)
        self._launches_per_step = int(lib.tcn_launch_count() - n0)
        return out

    def _stage(self, slot, x_rows, labels_u8, lengths):
        """Copy one batch into an input slot on the current stream (H2D when the sources are pinned host tensors)."""
        lay = SeqLayout.get(lengths, self.ex.device)
        assert lay.frames <= self.max_frames
        xs = x_rows if isinstance(x_rows, (list, tuple)) else [x_rows]
        ls = labels_u8 if isinstance(labels_u8, (list, tuple)) else [labels_u8]
        off = 0
        for x in xs:
            self.x_slots[slot][off:off + x.shape[0]].copy_(x, non_blocking=True)
            off += x.shape[0]
        assert off == lay.frames
        off = 0
        for lab in ls:
            self.lab_slots[slot][off:off + lab.shape[0], :lab.shape[1]].copy_(lab, non_blocking=True)
            off += lab.shape[0]
        return lay

    def prefetch(self, x_rows, labels_u8, lengths):
        """Stage the NEXT step's batch into the idle input slot on the copy streams; the following step() call
        (without arguments) consumes it.  Lets the H2D copy of step i + 1 overlap the kernels of step i.  The
        per-video copies alternate between two copy streams so that the set-up latency of one transfer hides behind
        the payload of the other (a single stream reaches ~42 GB/s on 16 MB transfers, the link does ~54 GB/s)."""
        nxt = self.slot ^ 1
        lay = SeqLayout.get(lengths, self.ex.device)
        assert lay.frames <= self.max_frames
        xs = x_rows if isinstance(x_rows, (list, tuple)) else [x_rows]
        ls = labels_u8 if isinstance(labels_u8, (list, tuple)) else [labels_u8]
        streams = (self.copy_stream, self.copy_stream2)
        for st in streams:
            st.wait_event(self.slot_free[nxt])
        off = 0
        for i, x in enumerate(xs):
            with torch.cuda.stream(streams[i & 1]):
                self.x_slots[nxt][off:off + x.shape[0]].copy_(x, non_blocking=True)
            off += x.shape[0]
        assert off == lay.frames
        off = 0
        for i, lab in enumerate(ls):
            with torch.cuda.stream(streams[(i + 1) & 1]):
                self.lab_slots[nxt][off:off + lab.shape[0], :lab.shape[1]].copy_(lab, non_blocking=True)
            off += lab.shape[0]
        self.slot_ready[nxt].record(self.copy_stream)
        self.slot_ready2[nxt].record(self.copy_stream2)
        self._prefetched = (nxt, lay)

    def step(self, x_rows=None, labels_u8=None, lengths=None, global_seqs=None):
        """One optimizer step.  x_rows / labels_u8: one tensor or a list with one tensor per video (device or pinned
        host); omit them to consume the batch staged by prefetch().  global_seqs: see _set_global_norm (unequal data-
        parallel shares).  Returns the device loss vector (loss_ivt, loss_i, loss_v, loss_t, total, 0, 0, 0)."""
        self._set_global_norm(global_seqs)
        cur = torch.cuda.current_stream()
        if x_rows is None:
            assert self._prefetched is not None, "step() without arguments needs a prefetch()"
            slot, lay = self._prefetched
            self._prefetched = None
            cur.wait_event(self.slot_ready[slot])
            cur.wait_event(self.slot_ready2[slot])
        else:
            slot = self.slot ^ 1 if self._prefetched is None else self.slot  # never the slot a prefetch is filling
            if self._prefetched is not None and self._prefetched[0] == slot:
                slot ^= 1
            lay = self._stage(slot, x_rows, labels_u8, lengths)
        self.slot = slot
        seed = int(torch.randint(0, 2 ** 31 - 1, (1,), generator=self.rng).item())
        self.ex.set_batch(lay, seed)
        if not self.use_graph:
            out = self._counted_body(slot)
        else:
            if self.graphs[slot] is None:
                # warm-up outside capture (lazy module loading, cudaFuncSetAttribute), then capture once per slot
                snap_p = self.flat_p.clone()
                self._body(slot)
                self.flat_p.copy_(snap_p)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._outs[slot] = self._counted_body(slot)
                self.flat_p.copy_(snap_p)
                self.graphs[slot] = g
            self.graphs[slot].replay()
            out = self._outs[slot]
        self.slot_free[slot].record(cur)
        return out

    def close(self):
        """Drop the captured graphs (do this before destroying a process group they reference)."""
        self.graphs = [None, None]
        self._outs = [None, None]
        self._arena_graphs = {}
