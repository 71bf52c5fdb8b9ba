"""Recipe that makes the UNMODIFIED reference modules of the hot path available on the GPU box.
TEST / BENCH INFRASTRUCTURE ONLY.

The reference is a set of script directories without setup.py / pyproject, so
``pip install --target baseline/_ref /root/reference`` (the contract's install step) has nothing to
install.  This script is the equivalent: it copies the handful of module files the hot path lives in
 byte for byte from ``/root/reference`` into ``baseline/_ref/`` under their own
relative paths.  ``baseline/_ref/`` is git-ignored (no reference text enters the history) but NOT
gpurun-ignored, so it travels to the GPU box, where ``oracle/ref_import.py`` finds it.  It is run by
``__graft_entry__.build()`` whenever ``/root/reference`` is present (this container only).

Consumers: ``bench.py --impl reference`` (CPU arm), ``bench.py``'s ``eager_gpu_baseline`` leg, and the
``-m gpu`` tests that compare the CUDA path with the live reference at BASELINE sizes.
"""
from __future__ import annotations

import filecmp
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("CVC_REFERENCE_SRC", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")

FILES = (
    "MT4MTLKD/Temporal_tenco/network.py",                 # a1-a6 (oracle of record)
    "TERL/0_5fold_TCN_black/network.py",                  # a1-a6 duplicates (TERL signatures)
    "MT4MTLKD/Temporal_mstct/MSTCT/Temporal_Encoder.py",  # a9-a12
    "MT4MTLKD/Temporal_mstct/MSTCT/TS_Mixer.py",          # a13
    "MT4MTLKD/Temporal_mstct/network.py",                 # a13 (Classifier; class statement exec'd only)
    "MT4MTLKD/Spatial_cnn/run.py",                        # a8 (DistillKL; class statement exec'd only)
    "MT4MTLKD/Spatial_cnn/network.py",                    # f3 (multi-teacher attention block)
)


def vendor(verbose: bool = False) -> bool:
    """Copy FILES into baseline/_ref.  Returns False (and does nothing) when the reference is absent."""
    if not os.path.isdir(SRC):
        return False
    for rel in FILES:
        src, dst = os.path.join(SRC, rel), os.path.join(DST, rel)
        if not os.path.exists(src):
            raise FileNotFoundError(src)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not (os.path.exists(dst) and filecmp.cmp(src, dst, shallow=False)):
            shutil.copyfile(src, dst)
            if verbose:
                print("vendored", rel)
    return True


if __name__ == "__main__":
    ok = vendor(verbose=True)
    print("baseline/_ref ready" if ok else f"{SRC} absent: nothing vendored")
    sys.exit(0)
