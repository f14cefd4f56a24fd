"""CPU-side checks of the boundary: the C-ABI library loads, exports every symbol that
include/svdb200.h declares, and refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build()
    from svdsolver_b200 import capi
    return capi


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "svdb200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(svdb200_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported(built):
    lib = ctypes.CDLL(built.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_no_cpu_fallback(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(built.SvdB200Error) as ei:
        built.Handle(64, 4, np.float64)
    assert ei.value.status == -4


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under svdsolver_b200/ or include/ may reference it."""
    bad = []
    for base in ("svdsolver_b200", "include"):
        for dp, _, fs in os.walk(os.path.join(ROOT, base)):
            for f in fs:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                    s = open(os.path.join(dp, f), errors="ignore").read()
                    if "svd_oracle" in s or "libsvdref" in s or "oracle/" in s:
                        bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_dist_local_cols(built):
    lib = built.lib()
    Z = ctypes.c_size_t
    tot = sum(lib.svdb200_dist_local_cols(Z(1024), Z(32), ctypes.c_int(r), ctypes.c_int(3)) for r in range(3))
    assert tot == 1024
