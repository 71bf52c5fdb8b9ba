"""MS-TCT golden fixtures (a9-a13) from the reference's own modules.  TEST INFRASTRUCTURE ONLY.

Composes Dropout(eval: identity) -> TemporalEncoder -> Temporal_Mixer -> Classifier by
hand following MT4MTLKD/Temporal_mstct/network.py:75-101 (that file's forward hard-codes
.cuda() at :85-88, so it cannot run on CPU as shipped).
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import ref_import

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _np(t):
    return t.detach().cpu().numpy()


def build(in_dim, dims, heads, mlp_ratio, num_block, emb, K, seed):
    enc_mod = ref_import.mstct_encoder()
    mix_mod = ref_import.mstct_mixer()
    Classifier = ref_import.mstct_classifier_class()
    torch.manual_seed(seed)
    enc = enc_mod.TemporalEncoder(in_feat_dim=in_dim, embed_dims=dims, num_head=heads,
                                  mlp_ratio=mlp_ratio, norm_layer=torch.nn.LayerNorm,
                                  num_block=num_block)
    mix = mix_mod.Temporal_Mixer(inter_channels=dims, embedding_dim=emb)
    cls = Classifier(emb, K)
    # default init leaves every bias 0 and LN affine at identity; perturb them so the
    # fixtures exercise those terms.
    with torch.no_grad():
        for mod in (enc, mix, cls):
            for n, p in mod.named_parameters():
                if p.dim() == 1:
                    p.add_(0.1 * torch.randn_like(p))
    return enc.eval(), mix.eval(), cls.eval()


def one(name, in_dim, dims, heads, ratio, nblk, emb, K, B, T, seed):
    out = {}
    enc, mix, cls = build(in_dim, dims, heads, ratio, nblk, emb, K, seed=seed)
    x = torch.randn(B, in_dim, T)
    feats = enc(x)
    concat = mix(feats)
    y, feat = cls(concat)
    g = torch.randn_like(y)
    (y * g).sum().backward()
    out["cfg"] = np.array([in_dim, *dims, heads, ratio, nblk, emb, K, B, T])
    out["x"], out["gy"], out["y"] = _np(x), _np(g), _np(y)
    out["concat"], out["feat"] = _np(concat), _np(feat)
    for i, f in enumerate(feats):
        out[f"enc_out.{i}"] = _np(f)
    for pre, mod in (("TemporalEncoder.", enc), ("Temporal_Mixer.", mix), ("classifier.", cls)):
        for k, v in mod.state_dict().items():
            out["sd." + pre + k] = _np(v)
        for k, v in mod.named_parameters():
            if v.grad is not None:
                out["grad." + pre + k] = _np(v.grad)
    np.savez_compressed(os.path.join(OUT, name), **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    one("mstct_small.npz", 24, [16, 24, 32, 40], 8, 2, 1, 16, 10, B=2, T=20, seed=21)
    # longer windows (three key chunks in the attention kernel), head dims 4..10, two blocks per stage
    one("mstct_mid.npz", 48, [32, 48, 64, 80], 8, 2, 2, 32, 15, B=2, T=150, seed=22)


if __name__ == "__main__":
    main()
