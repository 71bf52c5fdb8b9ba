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
