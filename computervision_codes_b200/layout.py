"""Packed, time-major activation layout shared by every kernel.

A batch of sequences (videos) is packed along the row axis; every sequence starts at a multiple of
128 rows so that a 128-row block never straddles two sequences.  ``meta`` is the int32 (nblk, 4)
table {lo, hi, in_delta, seq} the kernels read (include/tcn_b200.h).
"""
from __future__ import annotations

import numpy as np
import torch

BLK = 128
_cache: dict = {}


def round_up(a: int, b: int) -> int:
    return (a + b - 1) // b * b


class SeqLayout:
    def __init__(self, lengths, device, in_starts=None):
        """in_starts: first row of each sequence in the caller's unpadded arrays (features, labels).  Default: the
        sequences are concatenated (0, T0, T0 + T1, ...); a feature arena passes the clips' positions inside it."""
        self.lengths = [int(t) for t in lengths]
        assert len(self.lengths) > 0 and all(t > 0 for t in self.lengths)
        assert in_starts is None or len(in_starts) == len(self.lengths)
        self.device = torch.device(device)
        self.num_seqs = len(self.lengths)
        in_starts_arg = in_starts
        starts, in_starts, meta = [], [], []
        row = 0
        src = 0
        for s, T in enumerate(self.lengths):
            starts.append(row)
            if in_starts_arg is not None:
                src = int(in_starts_arg[s])
            in_starts.append(src)
            nb = round_up(T, BLK) // BLK
            for _ in range(nb):
                meta.append((row, row + T, src - row, s))
            row += nb * BLK
            src += T
        self.starts = starts
        self.in_starts = in_starts
        self.rows = row            # padded rows
        self.frames = sum(self.lengths)   # valid frames
        self.nblk = row // BLK
        self.meta_np = np.asarray(meta, dtype=np.int32).reshape(-1, 4)
        self.uniform_T = self.lengths[0] if len(set(self.lengths)) == 1 else None
        self.max_len = max(self.lengths)
        self._meta = self._seq_lo = self._seq_len = None

    # Device copies of the tables are made on first use: a pageable host-to-device copy blocks the host behind everything
    # queued on the stream, and the executor path (tcn_model_set_batch uploads meta_np itself) never needs them -- a
    # trainer that builds a new layout per step would otherwise serialise its input prefetch behind the previous step.
    @property
    def meta(self):
        if self._meta is None:
            self._meta = torch.from_numpy(self.meta_np).to(self.device)
        return self._meta

    @property
    def seq_lo(self):
        if self._seq_lo is None:
            self._seq_lo = torch.tensor(self.starts, dtype=torch.int32, device=self.device)
        return self._seq_lo

    @property
    def seq_len(self):
        if self._seq_len is None:
            self._seq_len = torch.tensor(self.lengths, dtype=torch.int32, device=self.device)
        return self._seq_len

    @staticmethod
    def get(lengths, device, in_starts=None) -> "SeqLayout":
        key = (tuple(int(t) for t in lengths), str(torch.device(device)),
               None if in_starts is None else tuple(int(t) for t in in_starts))
        lay = _cache.get(key)
        if lay is None:
            if len(_cache) > 4096:
                _cache.clear()
            lay = _cache[key] = SeqLayout(lengths, device, in_starts)
        return lay

    @staticmethod
    def uniform(B: int, T: int, device) -> "SeqLayout":
        return SeqLayout.get([T] * B, device)

    # ---- conversions for the module boundary (uniform batches only) -------------------------
    def pad_bct(self, x_bct: torch.Tensor) -> torch.Tensor:
        """(B, C, T) -> packed time-major (rows, C) fp32 buffer (pad rows zero)."""
        B, Cc, T = x_bct.shape
        assert self.uniform_T == T and self.num_seqs == B
        Tp = self.rows // B
        buf = torch.zeros(B, Tp, Cc, device=x_bct.device, dtype=torch.float32)
        buf[:, :T, :] = x_bct.permute(0, 2, 1)
        return buf.view(self.rows, Cc)

    def as_bct(self, buf: torch.Tensor, ncols: int, col0: int = 0) -> torch.Tensor:
        """Packed (rows, ld) buffer -> (B, ncols, T) strided view (no copy)."""
        B, T = self.num_seqs, self.uniform_T
        assert T is not None
        Tp = self.rows // B
        return buf.view(B, Tp, buf.shape[1])[:, :T, col0:col0 + ncols].permute(0, 2, 1)

    def as_btc(self, buf: torch.Tensor, ncols: int, col0: int = 0) -> torch.Tensor:
        B, T = self.num_seqs, self.uniform_T
        assert T is not None
        Tp = self.rows // B
        return buf.view(B, Tp, buf.shape[1])[:, :T, col0:col0 + ncols]
