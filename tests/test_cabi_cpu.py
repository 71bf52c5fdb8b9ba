"""The C-ABI library builds, loads and exports every symbol include/tcn_b200.h declares (no GPU)."""
import ctypes
import os
import re

from computervision_codes_b200 import _lib, build as build_mod

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = set()
    inc = os.path.join(ROOT, "include")
    for fn in os.listdir(inc):
        if fn.endswith(".h"):
            text = open(os.path.join(inc, fn)).read()
            text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
            names |= set(re.findall(r"\b(tcn_[a-z0-9_]+)\s*\(", text))
    return names


def test_library_builds_and_exports_header_symbols():
    build_mod.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 10
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/ but not exported"
    # and the Python binding table covers the header exactly
    assert set(_lib.SIGNATURES) == declared


def test_version_and_error_string_without_gpu():
    lib = _lib.load()
    assert lib.tcn_version() == 100
    # invalid arguments are rejected on the host before any CUDA call
    rc = lib.tcn_prep_weight(None, 0, 0, 0, 0, None, None)
    assert rc == -1
    assert b"tcn_prep_weight" in lib.tcn_last_error()
    assert lib.tcn_prep_weight_floats(64, 64, 3, 0) == 24 * 8 * 32 * 4


def test_cpu_tensors_are_rejected():
    import pytest
    import torch

    from computervision_codes_b200.tcn import DilatedResidualLayer

    with pytest.raises(RuntimeError):
        DilatedResidualLayer(1, 8, 8)(torch.zeros(1, 8, 16))


def test_layout_tables():
    import numpy as np
    import torch

    from computervision_codes_b200.layout import SeqLayout

    lay = SeqLayout([130, 5, 256], "cpu")
    assert lay.rows == 256 + 128 + 256 and lay.nblk == 5 and lay.frames == 391
    np.testing.assert_array_equal(lay.meta_np[:, 0], [0, 0, 256, 384, 384])
    np.testing.assert_array_equal(lay.meta_np[:, 1], [130, 130, 261, 640, 640])
    np.testing.assert_array_equal(lay.meta_np[:, 2], [0, 0, 130 - 256, 135 - 384, 135 - 384])
    np.testing.assert_array_equal(lay.meta_np[:, 3], [0, 0, 1, 2, 2])
    x = torch.arange(2 * 3 * 5, dtype=torch.float32).view(2, 3, 5)
    u = SeqLayout.uniform(2, 5, "cpu")
    buf = u.pad_bct(x)
    assert buf.shape == (256, 3)
    assert torch.equal(u.as_bct(buf, 3), x)
