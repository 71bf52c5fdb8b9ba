"""Python face of the native whole-model executor (csrc/model.cu, `tcn_model_*` in include/tcn_b200.h).

The executor owns its activation workspace; parameters and gradients live in two flat fp32 torch
buffers that the nn.Module's Parameters alias, so torch optimizers / state_dict / checkpoints of the
reference (Temporal_tenco/run.py:272-283,345-353) keep working unchanged.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .layout import BLK, SeqLayout, round_up


def canonical_param_names(layers_pg, layers_r, num_r):
    names = ["PG.conv_1x1.weight", "PG.conv_1x1.bias"]
    stages = [("PG", layers_pg)] + [(f"Rs.{s}", layers_r) for s in range(num_r)]
    for pre, n in stages:
        for i in range(n):
            names += [f"{pre}.layers.{i}.conv_dilated.weight", f"{pre}.layers.{i}.conv_dilated.bias",
                      f"{pre}.layers.{i}.conv_1x1.weight", f"{pre}.layers.{i}.conv_1x1.bias"]
    names += ["fpn.latlayer1.weight", "fpn.latlayer1.bias", "conv_out.weight", "conv_out_i.weight",
              "conv_out_v.weight", "conv_out_t.weight", "conv_out.bias", "conv_out_i.bias", "conv_out_v.bias",
              "conv_out_t.bias"]
    return names


class ModelExecutor:
    def __init__(self, model, max_rows, max_seqs=64):
        lib = _lib.load()
        # the configurations csrc/model.cu implements; anything else would silently train a different network
        if not model.use_fpn:
            raise _lib.TcnError("ModelExecutor implements the --fpn path (every reference script uses it)")
        if model.use_output:
            raise _lib.TcnError("ModelExecutor: args.output=True (Rs.*.conv_1x1 on the stage input, network.py:150-151) "
                                "is not on the executor path; use the module forward (forward_packed)")
        if len(model.Rs) != 3:
            raise _lib.TcnError("ModelExecutor: the FPN needs exactly 3 refinement stages (network.py:98-106)")
        if any(getattr(R, "hier", False) for R in model.Rs):
            raise _lib.TcnError("ModelExecutor: args.hier (AvgPool1d between stages) is not on the executor path")
        self.model = model
        dev = next(model.parameters()).device
        if dev.type != "cuda":
            raise _lib.TcnError("the executor needs the module on a CUDA device (no CPU fallback)")
        self.device = dev
        cfg = _lib.ModelConfig()
        cfg.layers_pg, cfg.layers_r, cfg.num_r = len(model.PG.layers), len(model.Rs[0].layers), len(model.Rs)
        cfg.channels, cfg.in_dim = model.PG.conv_1x1.out_channels, model.PG.conv_1x1.in_channels
        for i, k in enumerate(model.head_sizes):
            cfg.head_sizes[i] = k
        cfg.causal = int(model.PG.layers[0].causal)
        cfg.max_rows, cfg.max_seqs = round_up(max_rows, BLK), max_seqs
        self.cfg = cfg
        self.max_rows, self.max_seqs = cfg.max_rows, max_seqs
        h = C.c_void_p()
        with torch.cuda.device(dev):
            _lib.check(lib.tcn_model_create(C.byref(cfg), C.byref(h)), "tcn_model_create")
        self.h = h
        n = lib.tcn_model_num_tensors(h)
        offs, sizes = (C.c_longlong * n)(), (C.c_longlong * n)()
        _lib.check(lib.tcn_model_param_layout(h, offs, sizes, n), "tcn_model_param_layout")
        self.names = canonical_param_names(cfg.layers_pg, cfg.layers_r, cfg.num_r)
        assert len(self.names) == n
        total = lib.tcn_model_num_params(h)
        self.flat_p = torch.zeros(total, device=dev, dtype=torch.float32)
        self.flat_g = torch.zeros(total, device=dev, dtype=torch.float32)
        params = dict(model.named_parameters())
        for name, off, size in zip(self.names, offs, sizes):
            p = params[name]
            assert p.numel() == size, (name, p.shape, size)
            self.flat_p[off:off + size].copy_(p.data.reshape(-1))
            p.data = self.flat_p[off:off + size].view(p.shape)
            p.grad = self.flat_g[off:off + size].view(p.shape)
        _lib.check(lib.tcn_model_bind(h, _lib.ptr(self.flat_p), _lib.ptr(self.flat_g)), "tcn_model_bind")
        self.loss = torch.zeros(8, device=dev, dtype=torch.float32)
        self.ld_logits = round_up(sum(model.head_sizes), 4)
        self._lay = None

    def __del__(self):
        try:
            if getattr(self, "h", None):
                _lib.load().tcn_model_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # ------------------------------------------------------------------------------ configuration
    def set_loss(self, head_weights=(1.0, 0.1, 0.1, 0.1), pos_weight=None):
        """head_weights for (ivt, i, v, t); pos_weight: None or a flat list of sum(head_sizes) floats."""
        lib = _lib.load()
        hw = (C.c_float * 4)(*head_weights)
        pw = None
        if pos_weight is not None:
            assert len(pos_weight) == sum(self.model.head_sizes)
            pw = (C.c_float * len(pos_weight))(*pos_weight)
        _lib.check(lib.tcn_model_set_loss(self.h, hw, pw), "tcn_model_set_loss")

    def set_dropout(self, input_mask_p=0.0, chan_drop_p=0.5, layer_drop_p=0.5):
        _lib.check(_lib.load().tcn_model_set_dropout(self.h, input_mask_p, chan_drop_p, layer_drop_p),
                   "tcn_model_set_dropout")

    def set_loss_norm(self, norm_seqs: int = 0):
        """Average the loss over `norm_seqs` sequences (the global number of videos of a data-parallel step) instead of the
        sequences of this rank's batch; 0 = default.  Takes effect with the next set_batch."""
        _lib.check(_lib.load().tcn_model_set_loss_norm(self.h, int(norm_seqs)), "tcn_model_set_loss_norm")

    def set_batch(self, lay: SeqLayout, seed: int):
        assert lay.rows <= self.max_rows and lay.num_seqs <= self.max_seqs, "batch exceeds the executor capacity"
        meta = np.ascontiguousarray(lay.meta_np)
        _lib.check(_lib.load().tcn_model_set_batch(self.h, meta.ctypes.data_as(C.c_void_p), lay.nblk, lay.rows,
                                                   lay.num_seqs, lay.frames, int(seed) & 0xFFFFFFFF,
                                                   _lib.stream_ptr()), "tcn_model_set_batch")
        self._lay = lay

    # ------------------------------------------------------------------------------ execution
    def train_step(self, x_rows, labels_u8, training=True):
        """forward + loss + backward on the batch described by set_batch.  Returns the device tensor
        (loss_ivt, loss_i, loss_v, loss_t, total, 0, 0, 0); gradients are in flat_g / p.grad."""
        assert x_rows.is_contiguous() and labels_u8.is_contiguous() and labels_u8.dtype == torch.uint8
        _lib.check(_lib.load().tcn_model_train_step(self.h, _lib.ptr(x_rows), x_rows.shape[0], _lib.ptr(labels_u8),
                                                    labels_u8.shape[1], int(training), _lib.ptr(self.loss),
                                                    _lib.stream_ptr()), "tcn_model_train_step")
        return self.loss

    def forward(self, x_rows, training=False, keep_activations=False):
        """Returns (feature pointers, logits pointers) wrapped as torch views over executor memory:
        4 x (rows, C) and 4 x (rows, ld_logits) tensors, valid until the next executor call.
        keep_activations: keep what backward() needs even when dropout is off (eval-mode gradients)."""
        lib = _lib.load()
        feats, logits = (C.c_void_p * 4)(), (C.c_void_p * 4)()
        ld = C.c_int()
        mode = 1 if training else (2 if keep_activations else 0)
        _lib.check(lib.tcn_model_forward(self.h, _lib.ptr(x_rows), x_rows.shape[0], mode, feats, logits, C.byref(ld),
                                         _lib.stream_ptr()), "tcn_model_forward")
        rows = self._lay.rows
        Cc = self.cfg.channels
        return ([_view(feats[i], rows, Cc, self.device) for i in range(4)],
                [_view(logits[i], rows, ld.value, self.device) for i in range(4)])


    def backward(self, x_rows, glogits):
        """Backward of the last forward(training or keep_activations) from the gradients w.r.t. the four logit maps
        (rows, ld_logits; pad columns zero).  Gradients land in flat_g / p.grad (zeroed first)."""
        lib = _lib.load()
        gl = (C.c_void_p * 4)(*[_lib.ptr(g) for g in glogits])
        gf = (C.c_void_p * 4)(None, None, None, None)
        _lib.check(lib.tcn_model_backward(self.h, _lib.ptr(x_rows), x_rows.shape[0], gl, gf, _lib.stream_ptr()),
                   "tcn_model_backward")


class _RawCudaBuffer:
    """Minimal __cuda_array_interface__ holder so torch can wrap executor-owned memory without a copy."""

    def __init__(self, ptr, rows, cols):
        self.__cuda_array_interface__ = {"shape": (rows, cols), "typestr": "<f4", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def _view(ptr, rows, cols, device):
    return torch.as_tensor(_RawCudaBuffer(ptr, rows, cols), device=device)
