"""Import the reference's own modules from /root/reference.  TEST INFRASTRUCTURE ONLY.

In the build container the modules come from /root/reference; on the GPU box (which has no
/root/reference) from ``baseline/_ref`` -- the same files, copied unmodified by
``oracle/vendor_ref.py``.  Used by ``oracle/gen_golden*.py`` to write ``tests/golden/*.npz``, by the
tests that re-check against the live reference when it is present, and by bench.py's reference /
eager-GPU-baseline arms.
"""
from __future__ import annotations

import contextlib
import importlib.util
import io
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
_VENDORED = os.path.join(os.path.dirname(_HERE), "baseline", "_ref")   # written by oracle/vendor_ref.py (git-ignored)


def _default_root() -> str:
    env = os.environ.get("CVC_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isdir("/root/reference/MT4MTLKD"):
        return "/root/reference"
    return _VENDORED   # the GPU box: the unmodified module files copied by the recipe


REF_ROOT = _default_root()


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "MT4MTLKD", "Temporal_tenco"))


def _load(path: str, name: str):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    with contextlib.redirect_stdout(io.StringIO()):
        spec.loader.exec_module(mod)
    return mod


def tenco_network():
    """MT4MTLKD/Temporal_tenco/network.py (oracle of record for a1-a6)."""
    return _load(os.path.join(REF_ROOT, "MT4MTLKD/Temporal_tenco/network.py"), "_ref_tenco_network")


def terl_network():
    """TERL/0_5fold_TCN_black/network.py (same classes, a2/a3 duplicates)."""
    return _load(os.path.join(REF_ROOT, "TERL/0_5fold_TCN_black/network.py"), "_ref_terl_network")


def distill_kl_class():
    """DistillKL from MT4MTLKD/Spatial_cnn/run.py:284-295.

    run.py executes a training job at import, so only the class statement is
    executed: its source lines are read from the file and exec'd in a namespace
    that holds torch / nn / F.  No reference text is stored in this repository.
    """
    import torch
    from torch import nn
    import torch.nn.functional as F

    path = os.path.join(REF_ROOT, "MT4MTLKD/Spatial_cnn/run.py")
    lines = open(path).read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith("class DistillKL"))
    end = start + 1
    while end < len(lines) and (lines[end].startswith((" ", "\t")) or not lines[end].strip()):
        end += 1
    ns = {"torch": torch, "nn": nn, "F": F}
    exec("\n".join(lines[start:end]), ns)
    return ns["DistillKL"]


def _install_mstct_shims():
    """timm is absent (pinned timm==0.6.5 is used for trunc_normal_ only,
    MSTCT/Temporal_Encoder.py:1); map it to torch.nn.init.trunc_normal_."""
    import torch

    if "timm.models.layers" in sys.modules:
        return
    timm = types.ModuleType("timm")
    models = types.ModuleType("timm.models")
    layers = types.ModuleType("timm.models.layers")
    layers.trunc_normal_ = torch.nn.init.trunc_normal_
    layers.DropPath = torch.nn.Identity
    layers.to_2tuple = lambda x: (x, x)
    timm.models = models
    models.layers = layers
    sys.modules.update({"timm": timm, "timm.models": models, "timm.models.layers": layers})


def mstct_encoder():
    _install_mstct_shims()
    return _load(os.path.join(REF_ROOT, "MT4MTLKD/Temporal_mstct/MSTCT/Temporal_Encoder.py"),
                 "_ref_mstct_encoder")


def mstct_mixer():
    return _load(os.path.join(REF_ROOT, "MT4MTLKD/Temporal_mstct/MSTCT/TS_Mixer.py"),
                 "_ref_mstct_mixer")


def mstct_classifier_class():
    """Classifier from MT4MTLKD/Temporal_mstct/network.py:104-118 (the module itself
    fails at import on matplotlib / TSNE(n_iter=...), so only the class is exec'd)."""
    import torch
    from torch import nn

    path = os.path.join(REF_ROOT, "MT4MTLKD/Temporal_mstct/network.py")
    lines = open(path).read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith("class Classifier"))
    end = start + 1
    while end < len(lines) and (lines[end].startswith((" ", "\t")) or not lines[end].strip()):
        end += 1
    ns = {"torch": torch, "nn": nn}
    exec("\n".join(lines[start:end]), ns)
    return ns["Classifier"]


def spatial_cnn_network():
    """MT4MTLKD/Spatial_cnn/network.py (the student of the multi-teacher KD stage; row f3 uses lines 47-71).

    Its constructor asks torchvision for *pretrained* ResNet weights (a download) and its forward calls ``.cuda()``;
    callers patch ``basemodels.resnet18`` to the un-initialised architecture and swap ``model.basemodel`` for a stub
    that returns a given feature tensor, so that only the attention / projection lines of the reference run."""
    import torchvision.models as basemodels

    mod = _load(os.path.join(REF_ROOT, "MT4MTLKD/Spatial_cnn/network.py"), "_ref_spatial_cnn_network")
    orig = basemodels.resnet18
    mod.basemodels = types.SimpleNamespace(resnet18=lambda pretrained=True: orig(weights=None),
                                           resnet50=basemodels.resnet50)
    return mod
