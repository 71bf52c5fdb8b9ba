"""Drop-in mirror of ``MT4MTLKD/Temporal_mstct/network.py`` (VideoNas :46-101, Classifier :104-118)."""
from __future__ import annotations

import torch
from torch import nn

from .. import ops
from ..layout import SeqLayout
from . import functional as Fn
from .encoder import TemporalEncoder, _sid
from .mixer import Temporal_Mixer


class Classifier(nn.Module):
    """network.py:104-118: Conv1d(4E -> E, 1) -> Dropout -> Conv1d(E -> K, 1); returns ((B, T, K), feat (B, E, T))."""

    def __init__(self, embedding_dim, num_classes):
        super().__init__()
        self.linear_fuse = nn.Conv1d(in_channels=embedding_dim * 4, out_channels=embedding_dim, kernel_size=1)
        self.linear_pred = nn.Conv1d(embedding_dim, num_classes, kernel_size=1)
        self.dropout = nn.Dropout()
        self.num_classes = num_classes
        self._stream = _sid()

    def _packed(self, cat_rows, lay):
        x = ops.tap_linear(cat_rows, self.linear_fuse.weight, self.linear_fuse.bias, lay)
        feat = Fn.dropout_rows(x, self.dropout.p, self.training, self._stream)
        return ops.tap_linear(feat, self.linear_pred.weight, self.linear_pred.bias, lay), feat

    def forward(self, concat_feature):
        B, Cc, T = concat_feature.shape
        lay = SeqLayout.uniform(B, T, concat_feature.device)
        y, feat = self._packed(lay.pad_bct(concat_feature), lay)
        return lay.as_btc(y, self.num_classes), lay.as_bct(feat, feat.shape[1])


class VideoNas(nn.Module):
    """network.py:46-101.  ``args.loss_type`` selects the one classifier that is built; the heads that are not
    built return zeros, as in the reference."""

    def __init__(self, args, inter_channels, num_block, head, mlp_ratio, in_feat_dim, final_embedding_dim, num_tool=6,
                 num_verb=10, num_target=15, num_triplet=100):
        super().__init__()
        self.args = args
        self.dropout = nn.Dropout()
        self.TemporalEncoder = TemporalEncoder(in_feat_dim=in_feat_dim, embed_dims=inter_channels, num_head=head,
                                               mlp_ratio=mlp_ratio, norm_layer=nn.LayerNorm, num_block=num_block)
        self.Temporal_Mixer = Temporal_Mixer(inter_channels=inter_channels, embedding_dim=final_embedding_dim)
        self.sizes = {"i": num_tool, "v": num_verb, "t": num_target, "ivt": num_triplet}
        if self.args.loss_type in self.sizes:
            setattr(self, f"classifier_{self.args.loss_type}",
                    Classifier(final_embedding_dim, self.sizes[self.args.loss_type]))

    def forward(self, inputs):
        """inputs: (B, D, T) -> ((y_i, feat_i), (y_v, feat_v), (y_t, feat_t), (y_ivt, concat_feature))."""
        if not inputs.is_cuda:
            raise RuntimeError("computervision_codes_b200 runs on CUDA tensors only (no CPU fallback)")
        B, D, T = inputs.shape
        lay = SeqLayout.uniform(B, T, inputs.device)
        p = self.dropout.p if self.training else 0.0
        feats = self.TemporalEncoder._packed(lay.pad_bct(inputs), lay, in_drop_p=p)  # input Dropout folded into the load
        cat_rows = self.Temporal_Mixer._packed(feats, lay)
        concat_feature = lay.as_bct(cat_rows, cat_rows.shape[1])
        out = {k: (torch.zeros(B, T, n, device=inputs.device), concat_feature) for k, n in self.sizes.items()}
        lt = self.args.loss_type
        if lt in self.sizes:
            y, feat = getattr(self, f"classifier_{lt}")._packed(cat_rows, lay)
            feat_out = concat_feature if lt == "ivt" else lay.as_bct(feat, feat.shape[1])
            out[lt] = (lay.as_btc(y, self.sizes[lt]), feat_out)
        return out["i"], out["v"], out["t"], out["ivt"]
