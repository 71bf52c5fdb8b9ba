"""CPU restatement of the reference's input-side rules.  TEST INFRASTRUCTURE ONLY: imported by
tests/ only; the product (computervision_codes_b200/data.py) never imports this.

Pinned by construction: the reference Dataset classes need the CholecT45 files at import, so these few lines are a
restatement (numpy / ``random``), each citing the lines it follows.
"""
from __future__ import annotations

import numpy as np


def clip_indices(num_frames: int, split: str, rng):
    """MT4MTLKD/Temporal_tenco/dataloader.py:220-225 (= TERL/0_5fold_TCN_black/dataloader.py:269-274): frame indices
    one ``__getitem__`` call returns; ``rng`` is the ``random`` module or a ``random.Random``."""
    if split == "train" and rng.random() > 0.7:
        upper = 1000 if num_frames > 1000 else num_frames
        num_clips = rng.choice(range(10, upper))
        first = rng.choice(range(0, num_frames - num_clips))
        return [first + i for i in range(num_clips)]
    return [i for i in range(num_frames)]


def terl_kept_rows(feats: np.ndarray):
    """TERL/0_5fold_TCN_black/dataloader.py:252-257: indices of the frames kept by the duplicate-row filter."""
    diff = feats[1:, :] - feats[:-1, :]
    same = np.where(np.sum(diff, axis=-1) == 0)[0]
    gone = np.unique(np.concatenate((same, same + 1)))
    return [i for i in range(len(feats)) if i not in set(gone.tolist())]
