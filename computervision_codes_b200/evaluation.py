"""Evaluation and artefact formats either side of the temporal head.

* ``VideoAP`` -- per-video average precision on the device: the accumulator protocol the reference drives on
  ``ivtmetrics.Recognition`` (``reset`` / ``update`` / ``video_end`` / ``compute_video_AP``; call sites
  ``MT4MTLKD/Temporal_tenco/run.py:238-269`` and ``:428-450``) for ONE component.  Logits stay on the GPU; the sigmoid
  (the reference's ``activation``) and the ranking run in ``tcn_ap_rows``.  ``ivtmetrics==0.0.6`` is absent here, so
  parity with it is **unpinned**; the AP arithmetic is checked against ``sklearn.metrics.average_precision_score``,
  the function ivtmetrics calls.  The triplet -> component disentangling of ``compute_video_AP('i'|'v'|'t'|'iv'|'it')``
  needs ivtmetrics' mapping table and is not provided.
* ``write_artefact`` / ``read_artefact`` / ``artefact_name`` -- the pickles that carry features and teacher predictions
  between the stages: ``k{fold}_feats.pkl`` (``Temporal_tenco/dataloader.py:212-214``), ``k{fold}_{task}_feats.pkl`` and
  ``k{fold}_{task}_pred.pkl`` (writers ``Temporal_mstct/test.py:335-366``, ``Spatial_cnn/test.py:265-283``; readers
  ``Spatial_cnn/dataloader.py:216-238``): a dict video id (two characters) -> float32 ndarray (T, D) / (T, K).
"""
from __future__ import annotations

import pickle
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib


def video_ap(labels_u8: torch.Tensor, logits: torch.Tensor, apply_sigmoid: bool = True) -> torch.Tensor:
    """labels (T, >=K) uint8, logits (T, K) fp32, both on the device -> (K,) AP per class, NaN without positives."""
    lib = _lib.load()
    if not (labels_u8.is_cuda and logits.is_cuda):
        raise RuntimeError("video_ap runs on the CUDA kernel only (no CPU path)")
    lg = logits if logits.dtype == torch.float32 and logits.stride(1) == 1 else logits.float().contiguous()
    lab = labels_u8 if labels_u8.dtype == torch.uint8 and labels_u8.stride(1) == 1 else labels_u8.to(torch.uint8).contiguous()
    T, K = lg.shape
    assert lab.shape[0] == T and lab.shape[1] >= K
    ap = torch.empty(K, device=lg.device, dtype=torch.float32)
    _lib.check(lib.tcn_ap_rows(_lib.ptr(lg), lg.stride(0), _lib.ptr(lab), lab.stride(0), T, K, int(apply_sigmoid),
                               _lib.ptr(ap), _lib.stream_ptr()), "tcn_ap_rows")
    return ap


class VideoAP:
    def __init__(self, num_class: int, apply_sigmoid: bool = True):
        self.num_class, self.apply_sigmoid = num_class, apply_sigmoid
        self.reset_global()

    def reset(self):
        self._lab, self._log = [], []

    def reset_global(self):
        self.reset()
        self._videos = []

    def update(self, targets: torch.Tensor, logits: torch.Tensor):
        """targets (T, K) {0, 1}, logits (T, K): frames of the current video (device tensors)."""
        self._lab.append(targets.to(torch.uint8))
        self._log.append(logits.float())

    def video_end(self):
        if self._lab:
            self._videos.append(video_ap(torch.cat(self._lab), torch.cat(self._log), self.apply_sigmoid))
        self.reset()

    def compute_video_AP(self) -> Dict[str, object]:
        """Mean over videos per class (NaN-aware), then mean over classes -- one device -> host read."""
        if not self._videos:
            return {"AP": np.full(self.num_class, np.nan), "mAP": float("nan")}
        per_video = torch.stack(self._videos)                       # (videos, K)
        classwise = torch.nanmean(per_video, dim=0)
        out = torch.cat([classwise, torch.nanmean(classwise).reshape(1)]).cpu().numpy()
        return {"AP": out[:-1], "mAP": float(out[-1])}


def artefact_name(fold: int, kind: str, task: Optional[str] = None) -> str:
    """kind in {'feats', 'pred'}; task in {None, 'i', 'v', 't', 'ivt'}."""
    assert kind in ("feats", "pred")
    return f"k{fold}_{kind}.pkl" if task is None else f"k{fold}_{task}_{kind}.pkl"


def write_artefact(path: str, table: Dict[str, object]):
    """dict video id -> (T, D) tensor / array, stored as float32 ndarrays like the reference's writers."""
    out = {}
    for vid, v in table.items():
        a = v.detach().float().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v, dtype=np.float32)
        out[str(vid)] = np.ascontiguousarray(a, dtype=np.float32)
    with open(path, "wb") as fh:
        pickle.dump(out, fh)


def read_artefact(path: str, device=None) -> Dict[str, object]:
    with open(path, "rb") as fh:
        table = pickle.load(fh)
    if device is None:
        return table
    return {k: torch.as_tensor(np.asarray(v), dtype=torch.float32).to(device) for k, v in table.items()}
