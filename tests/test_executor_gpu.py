"""Native whole-model executor (csrc/model.cu) against the oracle, eval and train mode.  GPU only."""
import os
import types

import pytest
import torch

from oracle import tcn_oracle as O

from gradcheck import assert_grad_close

pytestmark = pytest.mark.gpu
DEV = "cuda"
ARGS = types.SimpleNamespace(fpn=True, output=False, feature=False, trans=False, mask=False, hier=False)


def _maxabs(a, b):
    return float((a.detach().double().cpu() - b.detach().double().cpu()).abs().max())


def _oracle_step(sd64, xs, labs, head_sizes, keeps=None, causal=False):
    """Mean over videos of the tenco loss; returns (loss terms, grads dict) in float64."""
    params = {k: v.clone().requires_grad_(True) for k, v in sd64.items()}
    total = 0.0
    terms = [0.0, 0.0, 0.0, 0.0]
    k0, k1, k2, k3 = head_sizes
    for s, (x, lab) in enumerate(zip(xs, labs)):
        kw = {}
        if keeps is not None:
            kw = dict(mask=keeps["mask"][s], chan_keep=keeps["chan"][s], layer_keeps=keeps["layers"][s], p=0.5)
        outs = O.videonas_forward(x.double().unsqueeze(0), params, causal=causal, **kw)
        y = lab.double()
        labels = (y[:, k0:k0 + k1], y[:, k0 + k1:k0 + k1 + k2], y[:, k0 + k1 + k2:k0 + k1 + k2 + k3], y[:, :k0])
        loss, li, lv, lt, livt = O.tenco_loss(outs[:4], labels)
        total = total + loss / len(xs)
        for i, t in enumerate((livt, li, lv, lt)):
            terms[i] += float(t) / len(xs)
    total.backward()
    return float(total), terms, {k: v.grad for k, v in params.items()}


@pytest.mark.parametrize("C,causal", [(64, False), (64, True), (32, False)])
def test_executor_eval_matches_oracle_and_eager(C, causal):
    from computervision_codes_b200.executor import ModelExecutor
    from computervision_codes_b200.layout import SeqLayout
    from computervision_codes_b200.tcn import VideoNas

    torch.manual_seed(3)
    D, heads = 48, (100, 6, 10, 15)
    m = VideoNas(ARGS, 5, 4, 3, C, D, 100, causal=causal).to(DEV)
    sd64 = {k: v.detach().double().cpu() for k, v in m.state_dict().items()}
    lengths = [300, 170, 129]
    g = torch.Generator().manual_seed(1)
    xs = [torch.randn(T, D, generator=g) for T in lengths]
    labs = [(torch.rand(T, 132, generator=g) < 0.1).to(torch.uint8) for T in lengths]
    for lab in labs:
        lab[:, 131] = 0
    ref_total, ref_terms, ref_grads = _oracle_step(sd64, xs, labs, heads, causal=causal)

    ex = ModelExecutor(m, max_rows=1024, max_seqs=4)
    lay = SeqLayout.get(lengths, DEV)
    ex.set_batch(lay, seed=5)
    x_rows = torch.cat(xs).to(DEV)
    lab_rows = torch.cat(labs).to(DEV)
    loss = ex.train_step(x_rows, lab_rows, training=False).cpu()
    assert abs(float(loss[4]) - ref_total) <= 1e-4 * abs(ref_total)
    for i in range(4):
        assert abs(float(loss[i]) - ref_terms[i]) <= 1e-4 * abs(ref_terms[i])
    for name, p in m.named_parameters():
        r = ref_grads[name]
        if r is None:
            continue
        assert p.grad is not None, name
        assert_grad_close(p.grad, r, name)
    # inference outputs of the executor == oracle (logits <= 1e-3, argmax identical)
    feats, logits = ex.forward(x_rows, training=False)
    torch.cuda.synchronize()
    with torch.no_grad():
        for s, T in enumerate(lengths):
            outs = O.videonas_forward(xs[s].double().unsqueeze(0), sd64, causal=causal)
            r0 = lay.starts[s]
            for lv in range(4):
                got = logits[lv][r0:r0 + T, :100].cpu()
                ref = outs[0][lv][0].t()
                assert _maxabs(got, ref) <= 1e-3
                assert torch.equal(got.argmax(1), ref.float().argmax(1))
                assert _maxabs(feats[lv][r0:r0 + T].cpu(), outs[4][lv][0].t()) <= 1e-4


@pytest.mark.parametrize("D", [40, 96])
def test_executor_train_mode_with_the_kernels_own_masks(D):
    """Train mode: input mask (p=0.25, no rescale), Dropout2d over input channels, nn.Dropout in every layer.
    The oracle is run with the very keep-masks the kernels derive from (seed, stream id).  D = 96: the projection and its
    weight gradient run on the tcgen05 kernels (TMA needs a 16-byte row pitch); D = 40: mma.sync kernels."""
    from computervision_codes_b200 import ops
    from computervision_codes_b200.executor import ModelExecutor
    from computervision_codes_b200.layout import SeqLayout
    from computervision_codes_b200.tcn import VideoNas

    torch.manual_seed(4)
    C, heads = 64, (100, 6, 10, 15)
    m = VideoNas(ARGS, 3, 2, 3, C, D, 100).to(DEV)
    sd64 = {k: v.detach().double().cpu() for k, v in m.state_dict().items()}
    lengths = [150, 200]
    g = torch.Generator().manual_seed(2)
    xs = [torch.randn(T, D, generator=g) for T in lengths]
    labs = [(torch.rand(T, 132, generator=g) < 0.1).to(torch.uint8) for T in lengths]
    ex = ModelExecutor(m, max_rows=1024, max_seqs=4)
    ex.set_dropout(input_mask_p=0.25, chan_drop_p=0.5, layer_drop_p=0.5)
    lay = SeqLayout.get(lengths, DEV)
    seed = 777
    ex.set_batch(lay, seed=seed)
    loss = ex.train_step(torch.cat(xs).to(DEV), torch.cat(labs).to(DEV), training=True).cpu()

    nl = 3 + 2 * 3
    layer_masks = [ops.dropout_keep_mask(lay.rows, C, 0.5, seed, l, DEV).cpu().double() for l in range(nl)]
    chan = ops.dropout_keep_mask(len(lengths), D, 0.5, seed, 0x7fff0001, DEV).cpu().double()
    inmask = ops.dropout_keep_mask(lay.rows, D, 0.25, seed, 0x7fff0002, DEV).cpu().double()
    keeps = {"mask": [], "chan": [], "layers": []}
    prefixes = ["PG"] * 3 + ["Rs.0"] * 2 + ["Rs.1"] * 2 + ["Rs.2"] * 2
    for s, T in enumerate(lengths):
        r0 = lay.starts[s]
        keeps["mask"].append(inmask[r0:r0 + T].t().unsqueeze(0))
        keeps["chan"].append(chan[s:s + 1])
        d = {}
        for l, pre in enumerate(prefixes):
            d.setdefault(pre, []).append(layer_masks[l][r0:r0 + T].t().unsqueeze(0))
        keeps["layers"].append(d)
    ref_total, ref_terms, ref_grads = _oracle_step(sd64, xs, labs, heads, keeps)
    assert abs(float(loss[4]) - ref_total) <= 1e-4 * abs(ref_total)
    for name, p in m.named_parameters():
        r = ref_grads[name]
        if r is None:
            continue
        assert_grad_close(p.grad, r, name)


def test_trainer_graph_replay_equals_eager_steps():
    """The CUDA-graph trainer (one graph, ragged batches of different shapes) follows the same SGD
    trajectory as stepping the executor without a graph."""
    from computervision_codes_b200.tcn import VideoNas
    from computervision_codes_b200.trainer import TemporalTrainer

    D = 32
    g = torch.Generator().manual_seed(9)
    batches = []
    for lengths in ([200], [90, 140], [257]):
        xs = torch.cat([torch.randn(T, D, generator=g) for T in lengths]).to(DEV)
        lab = (torch.rand(sum(lengths), 132, generator=g) < 0.1).to(torch.uint8).to(DEV)
        batches.append((xs, lab, lengths))
    results = []
    for use_graph in (False, True):
        torch.manual_seed(11)
        m = VideoNas(ARGS, 3, 2, 3, 64, D, 100).to(DEV).train()
        tr = TemporalTrainer(m, lr=0.05, weight_decay=1e-5, max_frames=512, max_seqs=4, use_graph=use_graph, seed=3)
        losses = [float(tr.step(*b)[4]) for b in batches for _ in range(2)]
        results.append((losses, tr.flat_p.clone()))
    la, lb = results[0][0], results[1][0]
    assert all(abs(a - b) <= 2e-4 * abs(a) for a, b in zip(la, lb)), (la, lb)
    assert _maxabs(results[0][1], results[1][1]) <= 1e-4


def test_trainer_follows_lr_schedule_inside_the_captured_graph():
    """f2: the SGD hyper-parameters live on the device, so set_lr() changes the update of an already captured graph;
    the update itself is torch.optim.SGD's p -= lr * (g + wd * p) (Temporal_tenco/run.py:345-346)."""
    from computervision_codes_b200.schedule import WarmupExponentialLR
    from computervision_codes_b200.tcn import VideoNas
    from computervision_codes_b200.trainer import TemporalTrainer

    D = 32
    g = torch.Generator().manual_seed(2)
    lengths = [150, 131]
    xs = torch.cat([torch.randn(T, D, generator=g) for T in lengths]).to(DEV)
    lab = (torch.rand(sum(lengths), 132, generator=g) < 0.1).to(torch.uint8).to(DEV)
    sched = WarmupExponentialLR(lr=0.01, power=0.1, warmup=3, decay_rate=0.5)
    finals = []
    for use_graph in (True, False):
        torch.manual_seed(4)
        m = VideoNas(ARGS, 3, 2, 3, 64, D, 100).to(DEV).train()
        tr = TemporalTrainer(m, lr=123.0, weight_decay=1e-3, max_frames=512, max_seqs=4, use_graph=use_graph, seed=8)
        for epoch in range(7):
            lr = sched.lr(epoch)
            tr.set_lr(lr)
            before = tr.flat_p.clone()
            tr.step(xs, lab, lengths)
            want = before - lr * (tr.flat_g + 1e-3 * before)
            assert _maxabs(tr.flat_p, want) <= 1e-6, (use_graph, epoch)
        finals.append(tr.flat_p.clone())
    assert _maxabs(finals[0], finals[1]) <= 1e-4


def test_step_cached_reads_clips_from_the_gpu_resident_feature_cache():
    """f1: a step on (video, start, length) items of the FeatureCache equals a step on the same frames passed from the
    host, and the reference-rule ClipSampler drives it."""
    import random

    import numpy as np

    from computervision_codes_b200.data import ClipSampler, FeatureCache
    from computervision_codes_b200.losses import pack_labels
    from computervision_codes_b200.tcn import VideoNas
    from computervision_codes_b200.trainer import TemporalTrainer

    D = 32
    rng = np.random.default_rng(6)
    cache = FeatureCache(DEV)
    host = {}
    for vid, T in (("01", 400), ("02", 260), ("03", 1100)):
        f = rng.standard_normal((T, D)).astype(np.float32)
        ys = [(rng.random((T, k)) < 0.1).astype(np.int64) for k in (6, 10, 15, 100)]
        cache.add_video(vid, f, *ys)
        host[vid] = (torch.from_numpy(f), pack_labels(*[torch.from_numpy(y) for y in ys]))
    sampler = ClipSampler("train", random.Random(1))
    items = []
    while len({n == cache.frames(v) for v, _, n in items}) < 2:   # make sure both a clip and a whole video occur
        items = [(v, *sampler.sample(cache.frames(v))) for v in ("03", "01", "02")]
    out = []
    for cached in (True, False, "packed", "packed-eager"):
        torch.manual_seed(4)
        m = VideoNas(ARGS, 3, 2, 3, 64, D, 100).to(DEV).train()
        tr = TemporalTrainer(m, lr=0.05, weight_decay=1e-5, max_frames=2048, max_seqs=4, seed=8,
                             use_graph=cached != "packed-eager")
        if cached == "packed":
            # one arena for all videos: the step reads clips in place (no copy), the block table addresses the arena
            cache.pack()
            assert cache.arena_x.shape[0] == 400 + 260 + 1100 and cache.feats["02"].data_ptr() == \
                cache.arena_x[400:].data_ptr()
        for _ in range(2):
            if cached:
                loss = tr.step_cached(cache, items)
            else:
                xs = [host[v][0][s:s + n].pin_memory() for v, s, n in items]
                ls = [host[v][1][s:s + n].pin_memory() for v, s, n in items]
                loss = tr.step(xs, ls, [n for _, _, n in items])
        out.append((loss.clone(), tr.flat_p.clone()))
    # same frames, same seeds, same arithmetic: only the addressing of the inputs differs
    for k in (1, 2, 3):
        assert _maxabs(out[0][0], out[k][0]) <= 5e-5 and _maxabs(out[0][1], out[k][1]) <= 5e-5, k


def test_trainer_steps_are_bit_reproducible():
    """Every weight-gradient reduction of the executor is a fixed-order slab sum (no fp32 atomics): two trainers started
    from the same weights and seed end with bit-identical parameters after several train-mode steps."""
    from computervision_codes_b200.tcn import VideoNas
    from computervision_codes_b200.trainer import TemporalTrainer

    D, lens = 96, [700, 129, 333]
    g = torch.Generator().manual_seed(3)
    x = torch.randn(sum(lens), D, generator=g).to(DEV)
    lab = (torch.rand(sum(lens), 132, generator=g) < 0.1).to(torch.uint8).to(DEV)
    outs = []
    for _ in range(2):
        torch.manual_seed(11)
        m = VideoNas(ARGS, 4, 3, 3, 64, D, 100).to(DEV).train()
        tr = TemporalTrainer(m, lr=0.05, weight_decay=1e-5, max_frames=2048, max_seqs=4, seed=5, input_mask_p=0.25)
        for _ in range(3):
            tr.step(x, lab, lens)
        torch.cuda.synchronize()
        outs.append(tr.flat_p.clone())
    assert torch.equal(outs[0], outs[1])


@pytest.mark.skipif(os.environ.get("TCN_LONG_TESTS") is None,
                    reason="long run (set TCN_LONG_TESTS=1): the detector of DESIGN.md section 3")
def test_graph_replayed_training_is_bit_reproducible_over_many_steps():
    """tools/exp/pdl_graph_check.py as a test: two trainers, same weights / seeds / eight cycling ragged batches at the
    BASELINE width, 400 graph-replayed steps each -- identical parameters at every mark.  With an early programmatic-launch
    trigger in the kernels this failed within 10-200 steps; the shipped build (no trigger) passed 800+ in every probe, but
    a rarer unexplained divergence (one per 2000-3000 steps) exists, hence opt-in."""
    import hashlib

    from computervision_codes_b200.tcn import VideoNas
    from computervision_codes_b200.trainer import TemporalTrainer

    args = types.SimpleNamespace(fpn=True, output=False, feature=False, trans=False, mask=True, hier=False)
    g = torch.Generator().manual_seed(2)
    batches = []
    for _ in range(8):
        lens = [int(t) for t in torch.randint(900, 3600, (8,), generator=g)]
        x = torch.randn(sum(lens), 2048, generator=g).to(DEV)
        lab = (torch.rand(sum(lens), 132, generator=g) < 0.05).to(torch.uint8).to(DEV)
        batches.append((x, lab, lens))
    marks, digests = (1, 10, 100, 200, 400), []
    for _ in range(2):
        torch.manual_seed(1)
        m = VideoNas(args, 11, 10, 3, 64, 2048, 100).to(DEV).train()
        tr = TemporalTrainer(m, lr=1e-3, weight_decay=1e-5, max_frames=8 * 3600 + 8 * 128, max_seqs=8, seed=5,
                             input_mask_p=0.25)
        out = []
        for s in range(1, marks[-1] + 1):
            x, lab, lens = batches[s % 8]
            tr.step(x, lab, lens)
            if s in marks:
                torch.cuda.synchronize()
                out.append(hashlib.sha256(tr.flat_p.cpu().numpy().tobytes()).hexdigest())
        digests.append(out)
        tr.close()
    assert digests[0] == digests[1]


@pytest.mark.parametrize("causal", [False, True])
def test_executor_edge_lengths_one_frame_to_block_boundaries(causal):
    """Ragged edge cases of the reference's zero padding: sequences of 1, 2 and 7 frames (every dilation tap of the
    deeper layers falls outside the video), exactly one 128-frame block, one frame more, and a sequence shorter than
    the largest dilation (d = 512 in an 11-layer stage).  Loss and every gradient against the fp64 oracle."""
    from computervision_codes_b200.executor import ModelExecutor
    from computervision_codes_b200.layout import SeqLayout
    from computervision_codes_b200.tcn import VideoNas

    torch.manual_seed(13)
    D, heads = 32, (100, 6, 10, 15)
    m = VideoNas(ARGS, 11, 3, 3, 64, D, 100, causal=causal).to(DEV)
    sd64 = {k: v.detach().double().cpu() for k, v in m.state_dict().items()}
    lengths = [1, 2, 7, 128, 129, 300]
    g = torch.Generator().manual_seed(2)
    xs = [torch.randn(T, D, generator=g) for T in lengths]
    labs = [(torch.rand(T, 132, generator=g) < 0.1).to(torch.uint8) for T in lengths]
    for lab in labs:
        lab[:, 131] = 0
    ref_total, ref_terms, ref_grads = _oracle_step(sd64, xs, labs, heads, causal=causal)
    ex = ModelExecutor(m, max_rows=1280, max_seqs=8)
    lay = SeqLayout.get(lengths, DEV)
    ex.set_batch(lay, seed=1)
    loss = ex.train_step(torch.cat(xs).to(DEV), torch.cat(labs).to(DEV), training=False).cpu()
    assert abs(float(loss[4]) - ref_total) <= 1e-4 * abs(ref_total)
    for name, p in m.named_parameters():
        r = ref_grads[name]
        if r is None:
            continue
        assert_grad_close(p.grad, r, name)
    feats, logits = ex.forward(torch.cat(xs).to(DEV), training=False)
    torch.cuda.synchronize()
    with torch.no_grad():
        for s, T in enumerate(lengths):
            outs = O.videonas_forward(xs[s].double().unsqueeze(0), sd64, causal=causal)
            r0 = lay.starts[s]
            for lv in range(4):
                got, ref = logits[lv][r0:r0 + T, :100].cpu(), outs[0][lv][0].t()
                assert _maxabs(got, ref) <= 1e-3
                assert torch.equal(got.argmax(1), ref.float().argmax(1))
