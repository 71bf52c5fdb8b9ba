"""Feature-sequence input pipeline: GPU-resident feature cache + clip sampler + packed labels.

The reference feeds the temporal heads one video per step from a pickled dict of pre-computed CNN features
(``MT4MTLKD/Temporal_tenco/dataloader.py:200-233``): ``T50.__getitem__`` returns either the whole video or, in the
train split with probability 0.3, a random clip of 10..999 consecutive frames; labels are four int64 matrices whose
first column (the frame id) is dropped; ``run.py:185`` then copies features and labels to the GPU every step.
The whole CholecT45 feature set is ~0.74 GB (90 k frames x 2048 x 4 B), so here it lives in HBM once:

* ``FeatureCache``   -- per video one (T, D) fp32 device tensor and one (T, 132) uint8 label tensor
                        (``losses.pack_labels`` order ivt | i | v | t); optional TERL duplicate-frame filter
                        (``TERL/0_5fold_TCN_black/dataloader.py:252-261``).
* ``ClipSampler``    -- the reference's clip rule, drawing from a ``random.Random`` in the same order as the
                        reference draws from the ``random`` module, so equal seeds give equal clips.
* ``FeatureCache.batch`` -- device views for a list of (video, start, length): what ``TemporalTrainer.step`` stages
                        with device-to-device copies (no host traffic in the step).
"""
from __future__ import annotations

import pickle
import random
from typing import Dict, Iterable, List, Sequence, Tuple

import numpy as np
import torch

from .losses import pack_labels


def terl_keep_index(feats: torch.Tensor) -> torch.Tensor:
    """Rows the TERL loader keeps (dataloader.py:252-257): a frame is dropped when the element-sum of its difference
    to the previous OR the next frame is exactly zero (duplicated feature rows)."""
    if feats.shape[0] < 2:
        return torch.arange(feats.shape[0], device=feats.device)
    sub = feats[1:] - feats[:-1]
    dup = (sub.sum(dim=-1) == 0).nonzero().flatten()
    drop = torch.zeros(feats.shape[0], dtype=torch.bool, device=feats.device)
    drop[dup] = True
    drop[dup + 1] = True
    return (~drop).nonzero().flatten()


class ClipSampler:
    """dataloader.py:220-225.  sample(T) -> (start, length); whole video outside the train split."""

    def __init__(self, split: str = "train", rng: random.Random | None = None):
        self.split = split
        self.rng = rng if rng is not None else random.Random()

    def sample(self, T: int) -> Tuple[int, int]:
        if self.split == "train" and self.rng.random() > 0.7:
            n = self.rng.choice(range(10, 1000 if T > 1000 else T))
            start = self.rng.choice(range(0, T - n))
            return start, n
        return 0, T


class FeatureCache:
    def __init__(self, device="cuda", terl_filter: bool = False):
        self.device = torch.device(device)
        self.terl_filter = terl_filter
        self.feats: Dict[str, torch.Tensor] = {}
        self.labels: Dict[str, torch.Tensor] = {}
        self.kept: Dict[str, torch.Tensor] = {}
        self.arena_x = None        # pack(): every video in one (frames, D) tensor / one (frames, 132) label tensor
        self.arena_lab = None
        self.offset: Dict[str, int] = {}

    def add_video(self, vid: str, feats, y_i, y_v, y_t, y_ivt, drop_id_column: bool = False):
        """feats (T, D) float; y_* (T, K) integer matrices (with the leading frame-id column of the reference's
        label files when ``drop_id_column``, dataloader.py:226-229)."""
        f = torch.as_tensor(np.asarray(feats), dtype=torch.float32).to(self.device)
        ys = [torch.as_tensor(np.asarray(y)) for y in (y_i, y_v, y_t, y_ivt)]
        if drop_id_column:
            ys = [y[:, 1:] for y in ys]
        lab = pack_labels(*[y.to(self.device) for y in ys])
        assert lab.shape[0] == f.shape[0], "features and labels disagree on the number of frames"
        if self.terl_filter:
            keep = terl_keep_index(f)
            self.kept[vid] = keep
            f, lab = f[keep].contiguous(), lab[keep].contiguous()
        self.feats[vid], self.labels[vid] = f.contiguous(), lab
        self.arena_x = self.arena_lab = None   # a new video invalidates the arena: pack() again

    def pack(self):
        """Move every video into ONE arena.  A step on cached clips (``TemporalTrainer.step_cached``) then reads features
        and labels in place -- the block table carries each clip's position inside the arena -- instead of copying them
        device-to-device into the trainer's input slot.  ``feats[vid]`` / ``labels[vid]`` become views of the arena."""
        vids = list(self.feats)
        assert vids, "pack(): the cache is empty"
        D, L = self.feats[vids[0]].shape[1], self.labels[vids[0]].shape[1]
        total = sum(self.feats[v].shape[0] for v in vids)
        ax = torch.empty(total, D, device=self.device, dtype=torch.float32)
        al = torch.empty(total, L, device=self.device, dtype=torch.uint8)
        off = 0
        for v in vids:
            n = self.feats[v].shape[0]
            ax[off:off + n].copy_(self.feats[v])
            al[off:off + n].copy_(self.labels[v])
            self.offset[v] = off
            self.feats[v], self.labels[v] = ax[off:off + n], al[off:off + n]
            off += n
        self.arena_x, self.arena_lab = ax, al
        return self

    def add_packed(self, vid: str, feats: torch.Tensor, labels_u8: torch.Tensor):
        """A video whose features (T, D) fp32 and packed uint8 labels (T, 132; ``losses.pack_labels``) already exist."""
        assert feats.shape[0] == labels_u8.shape[0] and labels_u8.dtype == torch.uint8
        self.feats[vid] = feats.to(self.device, torch.float32).contiguous()
        self.labels[vid] = labels_u8.to(self.device).contiguous()
        self.arena_x = self.arena_lab = None

    def add_pickle(self, path: str, labels: Dict[str, Sequence], drop_id_column: bool = True):
        """``k{fold}_feats.pkl`` of the reference (dict: video id -> (T, D) ndarray, dataloader.py:212-214);
        ``labels[vid] = (y_i, y_v, y_t, y_ivt)``."""
        with open(path, "rb") as fh:
            table = pickle.load(fh)
        for vid, f in table.items():
            if vid in labels:
                self.add_video(vid, f, *labels[vid], drop_id_column=drop_id_column)

    def __len__(self):
        return len(self.feats)

    def __contains__(self, vid):
        return vid in self.feats

    def frames(self, vid: str) -> int:
        return int(self.feats[vid].shape[0])

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self.feats.values()) + \
            sum(t.numel() for t in self.labels.values())

    def batch(self, items: Iterable[Tuple[str, int, int]]):
        """items: (video, start, length) -> (feature views, label views, lengths), all on the device."""
        xs: List[torch.Tensor] = []
        ls: List[torch.Tensor] = []
        lens: List[int] = []
        for vid, start, n in items:
            f = self.feats[vid]
            assert 0 <= start and n > 0 and start + n <= f.shape[0], (vid, start, n, f.shape[0])
            xs.append(f[start:start + n])
            ls.append(self.labels[vid][start:start + n])
            lens.append(int(n))
        return xs, ls, lens

    def sample_batch(self, vids: Sequence[str], sampler: ClipSampler):
        return self.batch([(v, *sampler.sample(self.frames(v))) for v in vids])
