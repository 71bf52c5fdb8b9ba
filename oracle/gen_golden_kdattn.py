"""Golden fixtures for the multi-teacher attention feature-KD block (row f3) from the reference's own forward.
TEST INFRASTRUCTURE ONLY.

Runs ``VideoNas.forward`` of MT4MTLKD/Spatial_cnn/network.py:44-96 (train mode, loss_type 'all') with the ResNet
backbone replaced by a stub returning a given (B, F, 1, 1) feature, and the KD term of Spatial_cnn/run.py:187-191
(three MSE losses / 3) on its outputs; records inputs, the six projection weights, outputs and gradients.
"""
from __future__ import annotations

import os
import types

import numpy as np
import torch
from torch import nn

from . import ref_import

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


class _Stub(nn.Module):
    def __init__(self, feat):
        super().__init__()
        self.feat = feat

    def forward(self, x):
        return self.feat


def one(tag, B, F, M, seed, out):
    mod = ref_import.spatial_cnn_network()
    args = types.SimpleNamespace(network="resnet18", teacher_dim=M, student_dim=F, loss_type="all", train=True)
    torch.manual_seed(seed)
    model = mod.VideoNas(args)
    feat = torch.randn(B, F, 1, 1).relu_().requires_grad_(True)   # post-ReLU pooled CNN feature
    model.basemodel = _Stub(feat)
    teachers = [torch.randn(B, M) for _ in range(3)]
    cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self   # network.py:84-87 hard-codes .cuda()
    try:
        (fi, _), (fv, _), (ft, _), (out_feat, _) = model.train()(torch.zeros(B, 3, 8, 8), *teachers)
    finally:
        torch.Tensor.cuda = cuda
    mse = nn.MSELoss()
    kd = (mse(fi, teachers[0]) + mse(fv, teachers[1]) + mse(ft, teachers[2])) / 3   # run.py:187-191
    kd.backward()
    out[f"{tag}.s"] = feat.detach().reshape(B, F).numpy()
    for n, t in zip("ivt", teachers):
        out[f"{tag}.teacher_{n}"] = t.numpy()
    for n, y in zip("ivt", (fi, fv, ft)):
        out[f"{tag}.stus_f{n}"] = y.detach().numpy()
    out[f"{tag}.kd_loss"] = np.array(float(kd))
    out[f"{tag}.grad.s"] = feat.grad.reshape(B, F).numpy()
    for name in ("wi", "wv", "wt", "mi", "mv", "mt"):
        conv = getattr(model, name)
        out[f"{tag}.sd.{name}.weight"] = conv.weight.detach().numpy()
        out[f"{tag}.sd.{name}.bias"] = conv.bias.detach().numpy()
        out[f"{tag}.grad.{name}.weight"] = conv.weight.grad.numpy()
        out[f"{tag}.grad.{name}.bias"] = conv.bias.grad.numpy()


def main():
    os.makedirs(OUT, exist_ok=True)
    out = {}
    one("small", B=6, F=48, M=40, seed=31, out=out)
    one("wide", B=8, F=512, M=96, seed=32, out=out)    # the scripts' student_dim
    np.savez_compressed(os.path.join(OUT, "kd_attn.npz"), **out)
    print({k: v.shape for k, v in out.items() if k.startswith("small")})


if __name__ == "__main__":
    main()
