"""f1 / f2 host logic on CPU: LR schedule vs torch's own schedulers, clip sampler and duplicate-frame filter vs the
restated reference rules (oracle/data_oracle.py), feature cache views and label packing."""
import random
import warnings

import numpy as np
import pytest
import torch

from computervision_codes_b200.data import ClipSampler, FeatureCache, terl_keep_index
from computervision_codes_b200.schedule import WarmupExponentialLR
from oracle import data_oracle


@pytest.mark.parametrize("lr,power,warm,decay", [(0.01, 0.1, 58, 0.99), (0.02, 0.25, 9, 0.9)])
def test_schedule_matches_torch_sequential_lr(lr, power, warm, decay):
    """Temporal_tenco/run.py:345-350 built from torch's schedulers, stepped once per epoch (:235-236)."""
    p = [torch.nn.Parameter(torch.zeros(1))]
    opt = torch.optim.SGD(p, lr=lr / power, weight_decay=1e-5)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        a = torch.optim.lr_scheduler.LinearLR(opt, start_factor=power, total_iters=warm)
        b = torch.optim.lr_scheduler.ExponentialLR(opt, gamma=decay)
        s = torch.optim.lr_scheduler.SequentialLR(opt, schedulers=[a, b], milestones=[warm + 1])
        mine = WarmupExponentialLR(lr, power, warm, decay)
        for e in range(200):
            ref = opt.param_groups[0]["lr"]
            assert abs(mine.lr(e) - ref) <= 1e-12 * ref, (e, ref, mine.lr(e))
            assert abs(mine.lr() - ref) <= 1e-12 * ref
            opt.step()
            s.step()
            mine.step()


@pytest.mark.parametrize("T", [11, 500, 1000, 1001, 3577])
def test_clip_sampler_draws_the_reference_clips(T):
    for seed in (0, 7):
        ref_rng, my = random.Random(seed), ClipSampler("train", random.Random(seed))
        n_clip = 0
        for _ in range(300):
            idx = data_oracle.clip_indices(T, "train", ref_rng)
            start, n = my.sample(T)
            assert idx[0] == start and len(idx) == n and idx[-1] == start + n - 1
            n_clip += n != T
        assert 40 < n_clip < 140   # p = 0.3
    assert ClipSampler("val", random.Random(0)).sample(T) == (0, T)
    assert len(data_oracle.clip_indices(T, "test", random.Random(0))) == T


def test_terl_duplicate_filter_matches_reference_rule():
    g = np.random.default_rng(3)
    f = g.standard_normal((200, 16)).astype(np.float32)
    for i in (0, 5, 6, 50, 120, 198):      # duplicate runs, including both ends
        f[i + 1] = f[i]
    want = data_oracle.terl_kept_rows(f)
    got = terl_keep_index(torch.from_numpy(f)).tolist()
    assert got == want and len(want) < 200
    cache = FeatureCache("cpu", terl_filter=True)
    zeros = [np.zeros((200, k), dtype=np.int64) for k in (6, 10, 15, 100)]
    cache.add_video("01", f, *zeros)
    assert cache.frames("01") == len(want)
    assert torch.equal(cache.feats["01"], torch.from_numpy(f[want]))


def test_feature_cache_views_and_label_packing():
    g = np.random.default_rng(4)
    cache = FeatureCache("cpu")
    vids = {}
    for vid, T in (("01", 40), ("02", 75)):
        f = g.standard_normal((T, 8)).astype(np.float32)
        ids = np.arange(T)[:, None]
        ys = [np.concatenate([ids, (g.random((T, k)) < 0.3).astype(np.int64)], axis=1) for k in (6, 10, 15, 100)]
        cache.add_video(vid, f, *ys, drop_id_column=True)   # label files carry the frame id in column 0
        vids[vid] = (f, ys)
    assert len(cache) == 2 and "02" in cache and cache.nbytes() == (40 + 75) * (8 * 4 + 132)
    xs, ls, lens = cache.batch([("02", 10, 30), ("01", 0, 40)])
    assert lens == [30, 40]
    f, ys = vids["02"]
    assert torch.equal(xs[0], torch.from_numpy(f[10:40]))
    lab = ls[0].numpy()
    assert lab.shape == (30, 132) and lab.dtype == np.uint8
    # column order ivt | i | v | t (+1 pad), losses.pack_labels
    assert np.array_equal(lab[:, :100], ys[3][10:40, 1:])
    assert np.array_equal(lab[:, 100:106], ys[0][10:40, 1:])
    assert np.array_equal(lab[:, 106:116], ys[1][10:40, 1:])
    assert np.array_equal(lab[:, 116:131], ys[2][10:40, 1:])
    assert not lab[:, 131].any()
    xs2, ls2, lens2 = cache.sample_batch(["01", "02"], ClipSampler("val"))
    assert lens2 == [40, 75] and xs2[1].data_ptr() == cache.feats["02"].data_ptr()   # views, not copies
    with pytest.raises(AssertionError):
        cache.batch([("01", 30, 20)])


def test_artefact_pickles_round_trip(tmp_path):
    """f4: the k{fold}[_{task}]_{feats,pred}.pkl tables (dict video id -> float32 (T, D)) written here load with plain
    pickle exactly like the reference's readers do (Spatial_cnn/dataloader.py:216-238), and back."""
    import pickle

    from computervision_codes_b200.evaluation import artefact_name, read_artefact, write_artefact

    assert artefact_name(3, "feats") == "k3_feats.pkl" and artefact_name(1, "pred", "v") == "k1_v_pred.pkl"
    g = np.random.default_rng(0)
    table = {"01": torch.from_numpy(g.standard_normal((7, 5))), "12": g.standard_normal((3, 5))}
    path = tmp_path / artefact_name(2, "feats", "i")
    write_artefact(str(path), table)
    with open(path, "rb") as fh:
        raw = pickle.load(fh)
    assert set(raw) == {"01", "12"} and raw["01"].dtype == np.float32 and raw["01"].shape == (7, 5)
    assert np.allclose(raw["12"], table["12"].astype(np.float32))
    back = read_artefact(str(path), device="cpu")
    assert torch.equal(back["01"], table["01"].float())
    cache = FeatureCache("cpu")
    zeros = [np.zeros((7, k), dtype=np.int64) for k in (6, 10, 15, 100)]
    cache.add_pickle(str(path), {"01": zeros}, drop_id_column=False)
    assert len(cache) == 1 and cache.frames("01") == 7


def test_packed_cache_is_one_arena_and_the_block_table_addresses_it():
    """f1, zero-copy steps: ``FeatureCache.pack()`` moves every video into one arena (views stay valid) and a layout
    built with ``in_starts`` points each 128-row block at its clip inside the arena: block b of sequence s covers padded
    rows [lo, hi) and reads unpadded rows ``row + in_delta`` -- exactly the clip's frames."""
    from computervision_codes_b200.layout import SeqLayout

    rng = np.random.default_rng(3)
    cache = FeatureCache("cpu")
    ref = {}
    for vid, T in (("a", 300), ("b", 129), ("c", 1000)):
        f = rng.standard_normal((T, 8)).astype(np.float32)
        ys = [(rng.random((T, k)) < 0.2).astype(np.int64) for k in (6, 10, 15, 100)]
        cache.add_video(vid, f, *ys)
        ref[vid] = (torch.from_numpy(f), cache.labels[vid].clone())
    assert cache.arena_x is None
    cache.pack()
    assert cache.arena_x.shape == (1429, 8) and cache.arena_lab.shape == (1429, 132)
    assert [cache.offset[v] for v in "abc"] == [0, 300, 429]
    for vid in "abc":   # the per-video tensors are now views of the arena with the same contents
        assert torch.equal(cache.feats[vid], ref[vid][0]) and torch.equal(cache.labels[vid], ref[vid][1])
        assert cache.feats[vid].data_ptr() == cache.arena_x[cache.offset[vid]:].data_ptr()
    items = [("c", 200, 500), ("a", 0, 300), ("b", 100, 29)]
    lens = [n for _, _, n in items]
    starts = [cache.offset[v] + s for v, s, _ in items]
    lay = SeqLayout.get(lens, "cpu", in_starts=starts)
    assert lay.frames == sum(lens) and lay.rows == 512 + 384 + 128 and lay.nblk == 8
    for lo, hi, delta, s in lay.meta_np.tolist():
        vid, start, n = items[s]
        assert hi - lo == n and lo % 128 == 0
        rows = torch.arange(lo, hi)
        assert torch.equal(cache.arena_x[rows + delta], ref[vid][0][start:start + n])
    # the default layout (concatenated inputs) is unchanged and cached under a different key
    cat = SeqLayout.get(lens, "cpu")
    assert cat.in_starts == [0, 500, 800] and cat is not lay and SeqLayout.get(lens, "cpu", in_starts=starts) is lay
    # adding a video invalidates the arena
    cache.add_video("d", rng.standard_normal((5, 8)).astype(np.float32),
                    *[np.zeros((5, k), dtype=np.int64) for k in (6, 10, 15, 100)])
    assert cache.arena_x is None
