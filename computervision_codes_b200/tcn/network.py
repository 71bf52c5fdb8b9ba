"""Drop-in mirror of ``MT4MTLKD/Temporal_tenco/network.py`` (== ``TERL/0_5fold_TCN_black/network.py``).

Same class names, constructor arguments, forward signatures, return structure and ``state_dict``
keys / shapes as the reference, so ``import network`` in the reference's
``run.py`` can be pointed here and its checkpoints load unchanged.  The nn.Conv1d children exist
only to own the parameters (identical names, shapes and default initialisation); their forward is
never called -- all arithmetic runs in the CUDA kernels behind include/tcn_b200.h, on activations
kept time-major (frames x channels) and returned as strided (B, C, T) views.
"""
from __future__ import annotations

import copy

import torch
from torch import nn

from .. import ops
from ..layout import SeqLayout

_stream_counter = [0]


def _next_stream_id() -> int:
    _stream_counter[0] += 1
    return _stream_counter[0]


def _check_input(x: torch.Tensor):
    if not x.is_cuda:
        raise RuntimeError("computervision_codes_b200 runs on CUDA tensors only (no CPU fallback)")


class _LayerBase(nn.Module):
    causal = False

    def __deepcopy__(self, memo):
        """A plain deep copy with its own dropout stream id.  It must not draw from the RNG: the reference builds its
        stages as copy.deepcopy(Layer(...)) (network.py:112,142), so under the same seed this mirror has to consume the
        generator exactly as nn.Module's default deepcopy does -- not at all."""
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = copy.deepcopy(v, memo)
        new._stream_id = _next_stream_id()
        return new

    def _run_packed(self, x_rows, lay):
        """x_rows: packed (rows, C).  Returns packed (rows, C)."""
        p = self.dropout.p if self.training else 0.0
        seed = ops.new_seed() if p > 0 else 0
        return ops.dilated_residual(x_rows, self.conv_dilated.weight, self.conv_dilated.bias,
                                    self.conv_1x1.weight, self.conv_1x1.bias, lay, self.dilation, self.causal,
                                    p, seed, self._stream_id)

    def forward(self, x):
        """x: (B, C, T) -> (B, C, T)   (reference layout; converted at the module boundary)."""
        _check_input(x)
        B, Cc, T = x.shape
        lay = SeqLayout.uniform(B, T, x.device)
        y = self._run_packed(lay.pad_bct(x), lay)
        return lay.as_bct(y, Cc)


class DilatedResidualLayer(_LayerBase):
    """network.py:186-198 -- taps x[t-d], x[t], x[t+d]."""

    def __init__(self, dilation, in_channels, out_channels):
        super().__init__()
        assert in_channels == out_channels, "the residual add needs in_channels == out_channels"
        self.conv_dilated = nn.Conv1d(in_channels, out_channels, 3, padding=dilation, dilation=dilation)
        self.conv_1x1 = nn.Conv1d(out_channels, out_channels, 1)
        self.dropout = nn.Dropout()
        self.dilation = dilation
        self._stream_id = _next_stream_id()



class DilatedResidualCausalLayer(_LayerBase):
    """network.py:165-183 -- front padding 2d: taps x[t-2d], x[t-d], x[t] (TeCNO causal)."""
    causal = True

    def __init__(self, dilation, in_channels, out_channels, padding=None):
        super().__init__()
        assert in_channels == out_channels
        if padding is not None and padding != 2 * dilation:
            raise NotImplementedError("only padding == 2 * dilation keeps x + out shape-consistent")
        self.padding = 2 * dilation
        self.conv_dilated = nn.Conv1d(in_channels, out_channels, 3, padding=0, dilation=dilation)
        self.conv_1x1 = nn.Conv1d(out_channels, out_channels, 1)
        self.dropout = nn.Dropout()
        self.dilation = dilation
        self._stream_id = _next_stream_id()



def _make_layers(num_layers, num_f_maps, causal):
    cls = DilatedResidualCausalLayer if causal else DilatedResidualLayer
    return nn.ModuleList([copy.deepcopy(cls(2 ** i, num_f_maps, num_f_maps)) for i in range(num_layers)])


class BaseCausalTCN(nn.Module):
    """network.py:109-135.  ``causal=True`` (extension) swaps in the causal layer the reference
    defines but never instantiates."""

    def __init__(self, num_layers, num_f_maps, dim, num_classes, causal=False):
        super().__init__()
        self.conv_1x1 = nn.Conv1d(dim, num_f_maps, 1)
        self.layers = _make_layers(num_layers, num_f_maps, causal)
        self.conv_out = nn.Conv1d(num_f_maps, num_classes, 1)
        self.channel_dropout = nn.Dropout2d()
        self.num_classes = num_classes

    def _features_packed(self, x_rows, lay, mask_rows=None):
        """x_rows: (frames, D) contiguous frames of the batch, unpadded.  Returns packed (rows, C) features."""
        D = x_rows.shape[-1]
        x_rows = x_rows.reshape(-1, D)
        if mask_rows is not None:
            x_rows = x_rows * mask_rows.reshape(-1, D)
        colscale = None
        if self.training and self.channel_dropout.p > 0:
            p = self.channel_dropout.p
            keep = (torch.rand(lay.num_seqs, D, device=x_rows.device) >= p).float()
            colscale = (keep / (1.0 - p)).contiguous()
        out = ops.tap_linear(x_rows, self.conv_1x1.weight, self.conv_1x1.bias, lay, x_unpadded=True,
                             colscale=colscale)
        for layer in self.layers:
            out = layer._run_packed(out, lay)
        return out

    def forward(self, x, mask=None, labels=None, mask_labels=None, test=False):
        """x: (B, D, T) -> (features (B, C, T), conv_out(features) (B, K, T))."""
        _check_input(x)
        B, D, T = x.shape
        lay = SeqLayout.uniform(B, T, x.device)
        x_btd = x.permute(0, 2, 1).contiguous().float()
        mask_btd = mask.permute(0, 2, 1) if mask is not None else None
        f = self._features_packed(x_btd, lay, mask_btd)
        logits = ops.tap_linear(f, self.conv_out.weight, self.conv_out.bias, lay)
        return lay.as_bct(f, f.shape[1]), lay.as_bct(logits, self.num_classes)


class Refinement(nn.Module):
    """network.py:138-162."""

    def __init__(self, args, num_layers, num_f_maps, dim, num_classes, conv_out, causal=False):
        super().__init__()
        self.conv_1x1 = nn.Conv1d(dim, num_f_maps, 1)
        self.layers = _make_layers(num_layers, num_f_maps, causal)
        self.conv_out = nn.Conv1d(num_f_maps, num_classes, 1)
        self.max_pool_1x1 = nn.AvgPool1d(kernel_size=7, stride=3)
        self.use_output = args.output
        self.hier = args.hier
        self.num_classes = num_classes
        if self.hier:
            raise NotImplementedError("args.hier (AvgPool1d(7,3) between stages) is out of scope: no reference "
                                      "script enables it (DESIGN.md)")

    def _features_packed(self, f_rows, lay):
        out = f_rows
        if self.use_output:
            w = self.conv_1x1.weight
            k = out.shape[1]
            if k % 4 != 0:   # the kernels need a 16-byte row pitch: zero-pad the K input channels (and the weight with them)
                pad = 4 - k % 4
                out = torch.nn.functional.pad(out, (0, pad))
                w = torch.nn.functional.pad(w, (0, 0, 0, pad))
            out = ops.tap_linear(out, w, self.conv_1x1.bias, lay)
        for layer in self.layers:
            out = layer._run_packed(out, lay)
        return out

    def forward(self, x):
        _check_input(x)
        B, Cc, T = x.shape
        lay = SeqLayout.uniform(B, T, x.device)
        f = self._features_packed(lay.pad_bct(x), lay)
        logits = ops.tap_linear(f, self.conv_out.weight, self.conv_out.bias, lay)
        return lay.as_bct(f, f.shape[1]), lay.as_bct(logits, self.num_classes)


class FPN(nn.Module):
    """network.py:71-106.  Only latlayer1 is applied (three times); latlayer2/3 are parameters the
    reference never uses.  With equal lengths F.interpolate(mode='linear') is the identity."""

    def __init__(self, num_f_maps):
        super().__init__()
        self.latlayer1 = nn.Conv1d(num_f_maps, num_f_maps, kernel_size=1, stride=1, padding=0)
        self.latlayer2 = nn.Conv1d(num_f_maps, num_f_maps, kernel_size=1, stride=1, padding=0)
        self.latlayer3 = nn.Conv1d(num_f_maps, num_f_maps, kernel_size=1, stride=1, padding=0)

    def _packed(self, f_rows_list, lay):
        c1, c2, c3, p4 = f_rows_list
        w, b = self.latlayer1.weight, self.latlayer1.bias
        p3 = ops.tap_linear(c3, w, b, lay, residual=p4)
        p2 = ops.tap_linear(c2, w, b, lay, residual=p3)
        p1 = ops.tap_linear(c1, w, b, lay, residual=p2)
        return [p1, p2, p3, p4]

    def forward(self, out_list):
        x0 = out_list[0]
        _check_input(x0)
        B, Cc, T = x0.shape
        if any(o.shape[-1] != T for o in out_list):
            raise NotImplementedError("FPN over levels of different length (args.hier) is out of scope")
        lay = SeqLayout.uniform(B, T, x0.device)
        ps = self._packed([lay.pad_bct(o) for o in out_list], lay)
        return [lay.as_bct(p, Cc) for p in ps]


class _VideoNasExecFn(torch.autograd.Function):
    """VideoNas forward / backward through the native executor (csrc/model.cu): two C calls per step instead of a
    few hundred Python-dispatched launches.  Outputs: 4 packed logit maps (rows, 132) + 4 packed feature maps."""

    @staticmethod
    def forward(ctx, module, x_rows, lay, training, need_grad, *params):
        ex = module._get_executor(lay)
        # the executor keeps ONE set of saved activations: stamp this forward so that a backward that follows a later
        # forward (two forwards before backward, a grad-enabled validation pass, a re-created executor) fails loudly
        module._exec_gen += 1
        ctx.exec_gen, ctx.ex = module._exec_gen, ex
        ex.set_batch(lay, ops.new_seed() if training else 0)
        feats, logits = ex.forward(x_rows, training=training, keep_activations=need_grad)
        outs = [t.clone() for t in logits] + [t.clone() for t in feats]  # executor memory is reused by the next call
        ctx.module, ctx.x_rows, ctx.nparams = module, x_rows, len(params)
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(*outs[4:])  # gradients w.r.t. the returned feature maps are not supported
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gouts):
        module = ctx.module
        ex = module._executor
        if ex is not ctx.ex or module._exec_gen != ctx.exec_gen:
            raise RuntimeError("VideoNas (native executor): backward() of a forward whose saved activations were "
                               "overwritten by a later forward -- the executor supports one outstanding forward "
                               "(INTEGRATION.md); call backward before the next forward, or run the later forward "
                               "under torch.no_grad() on a second module")
        glogits = []
        for i in range(4):
            g = gouts[i]
            if g is None:
                g = torch.zeros(ex._lay.rows, ex.ld_logits, device=ex.device, dtype=torch.float32)
            glogits.append(g.contiguous())
        ex.backward(ctx.x_rows, glogits)
        flat = ex.flat_g.clone()  # one copy: later backward calls reuse (and zero) the executor's buffer
        grads = []
        offs = module._exec_offsets
        for name, p in module._exec_params:
            if name in offs:
                o, n = offs[name]
                grads.append(flat[o:o + n].view(p.shape))
            else:
                grads.append(None)
        return (None, None, None, None, None) + tuple(grads)


class VideoNas(nn.Module):
    """network.py:14-68."""

    def __init__(self, args, num_layers_PG, num_layers_R, num_R, num_f_maps, dim, num_classes, num_i=6, num_v=10,
                 num_t=15, causal=False):
        super().__init__()
        self.PG = BaseCausalTCN(num_layers_PG, num_f_maps, dim, num_classes, causal=causal)
        self.conv_out = nn.Conv1d(num_f_maps, num_classes, 1)
        self.conv_out_i = nn.Conv1d(num_f_maps, num_i, 1)
        self.conv_out_v = nn.Conv1d(num_f_maps, num_v, 1)
        self.conv_out_t = nn.Conv1d(num_f_maps, num_t, 1)
        self.args = args
        self.Rs = nn.ModuleList(
            [copy.deepcopy(Refinement(args, num_layers_R, num_f_maps, num_classes, num_classes, self.conv_out,
                                      causal=causal)) for _ in range(num_R)])
        self.use_fpn = args.fpn
        self.use_output = args.output
        self.use_feature = getattr(args, "feature", False)
        self.use_trans = getattr(args, "trans", False)
        self.head_sizes = (num_classes, num_i, num_v, num_t)
        if args.fpn:
            self.fpn = FPN(num_f_maps)

    def _head_weights(self):
        w = torch.cat([self.conv_out.weight, self.conv_out_i.weight, self.conv_out_v.weight,
                       self.conv_out_t.weight], dim=0)
        b = torch.cat([self.conv_out.bias, self.conv_out_i.bias, self.conv_out_v.bias, self.conv_out_t.bias], dim=0)
        return w, b

    def forward_packed(self, x_rows, lay, mask_rows=None):
        """Packed core of forward.  x_rows: (frames, D) or (B, T, D) contiguous.  Returns (f_rows list,
        logits_rows list) with the four heads concatenated along the columns in the order ivt | i | v | t."""
        f = self.PG._features_packed(x_rows, lay, mask_rows)
        f_list = [f]
        for R in self.Rs:
            f = R._features_packed(f, lay)
            f_list.append(f)
        logits = []
        if self.use_fpn:
            f_list = self.fpn._packed(f_list, lay)
            w, b = self._head_weights()
            logits = [ops.tap_linear(p, w, b, lay) for p in f_list]
        return f_list, logits

    # ---- native executor path --------------------------------------------------------------------------
    _executor = None
    _exec_gen = 0

    def _get_executor(self, lay):
        from ..executor import ModelExecutor, canonical_param_names

        ex = self._executor
        if ex is None or ex.max_rows < lay.rows or ex.max_seqs < lay.num_seqs or ex.device != next(self.parameters()).device:
            rows = max(lay.rows, ex.max_rows if ex is not None else 0)
            self._executor = ex = ModelExecutor(self, max_rows=rows, max_seqs=max(lay.num_seqs, 8))
            ex.set_dropout(0.0, self.PG.channel_dropout.p, self.PG.layers[0].dropout.p)
            names = ex.names
            lib_offs = {}
            params = dict(self.named_parameters())
            base = ex.flat_p.data_ptr()
            for n in names:
                p = params[n]
                lib_offs[n] = ((p.data_ptr() - base) // 4, p.numel())
            self._exec_offsets = lib_offs
            self._exec_params = list(self.named_parameters())
            for _, p in self._exec_params:
                p.grad = None  # autograd assigns the gradients this path returns (the trainer binds flat views instead)
        return ex

    def _executor_ok(self):
        c = self.PG.conv_1x1.out_channels
        return (self.use_fpn and not self.use_output and len(self.Rs) == 3 and c % 4 == 0
                and self.PG.conv_1x1.in_channels % 4 == 0)

    def forward(self, x, ismask):
        """x: (B, T, D).  Returns (out_list, out_list_i, out_list_v, out_list_t, f_list, f_list)."""
        _check_input(x)
        B, T, D = x.shape
        lay = SeqLayout.uniform(B, T, x.device)
        x_btd = x.contiguous().float()
        mask_btd = None
        want_mask = bool(getattr(self.args, "mask", False) and ismask)
        # Executor path in train mode: the 25 % input mask (network.py:43-48) and Dropout2d are drawn on the device by
        # the projection kernel's counter-based generator (Bernoulli(0.25) per element instead of the reference's
        # exact-count permutation), as in TemporalTrainer -- no host randperm, no H2D, no eager multiply.
        device_mask = want_mask and self.training and self._executor_ok() and not x.requires_grad
        if want_mask and not device_mask:
            n = x_btd.numel()
            num_mask = int(n * 0.75)
            mask = torch.cat((torch.zeros(n - num_mask), torch.ones(num_mask)))
            mask = mask[torch.randperm(n)].view(B, D, T).to(x.device)  # network.py:43-48 (flat (B, D, T) order)
            mask_btd = mask.permute(0, 2, 1)
        if self._executor_ok() and not x.requires_grad:
            xin = x_btd if mask_btd is None else x_btd * mask_btd
            x_rows = xin.reshape(B * T, D).contiguous()
            lay_e = lay
            ex = self._get_executor(lay_e)
            ex.set_dropout(0.25 if device_mask else 0.0, self.PG.channel_dropout.p, self.PG.layers[0].dropout.p)
            plist = [p for _, p in self._exec_params]
            need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in plist)
            outs = _VideoNasExecFn.apply(self, x_rows, lay_e, self.training, need_grad, *plist)
            logit_rows, f_rows = list(outs[:4]), list(outs[4:])
        else:
            f_rows, logit_rows = self.forward_packed(x_btd, lay, mask_btd)
        Cc = f_rows[0].shape[1]
        out_list, out_i, out_v, out_t = [], [], [], []
        if not self.use_fpn:
            pg_logits = ops.tap_linear(f_rows[0], self.PG.conv_out.weight, self.PG.conv_out.bias, lay)
            out_list.append(lay.as_bct(pg_logits, self.PG.num_classes))
        else:
            k0, k1, k2, k3 = self.head_sizes
            for lg in logit_rows:
                out_list.append(lay.as_bct(lg, k0, 0))
                out_i.append(lay.as_bct(lg, k1, k0))
                out_v.append(lay.as_bct(lg, k2, k0 + k1))
                out_t.append(lay.as_bct(lg, k3, k0 + k1 + k2))
        f_list = [lay.as_bct(f, Cc) for f in f_rows]
        return out_list, out_i, out_v, out_t, f_list, f_list
