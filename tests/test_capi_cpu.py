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


def test_list_plan_balances_the_chains():
    """svdb200_list_plan (host logic of svdb200_bidiagonalize_many_*): every matrix exactly once, chains balanced by
    longest-processing-time-first (the reference's ascending benchmark list dealt out i % 2 leaves one chain 8 % longer),
    ascending sizes inside a chain, alternating issue order; degenerate inputs keep the caller's order."""
    import ctypes
    from svdsolver_b200 import capi
    lib = capi.lib()

    def plan(lanes, sizes):
        cnt = len(sizes)
        n = (ctypes.c_size_t * cnt)(*sizes)
        order = (ctypes.c_size_t * cnt)()
        chain = (ctypes.c_int * cnt)()
        assert lib.svdb200_list_plan(ctypes.c_int(lanes), ctypes.c_size_t(cnt), n, order, chain) == 0
        return list(order), list(chain)

    sizes = list(range(320, 3841, 320))                           # BASELINE configs[1]
    order, chain = plan(2, sizes)
    assert sorted(order) == list(range(len(sizes)))
    loads = [sum(sizes[i] for i, c in zip(order, chain) if c == l) for l in (0, 1)]
    assert loads[0] == loads[1] == sum(sizes) // 2                # 12480 / 12480 (i % 2 gives 11520 / 13440)
    for l in (0, 1):
        mine = [sizes[i] for i, c in zip(order, chain) if c == l]
        assert mine == sorted(mine)
    assert chain[:4] == [0, 1, 0, 1]
    # all sizes equal: round robin in the caller's order
    order, chain = plan(2, [512] * 6)
    assert order == list(range(6)) and chain == [0, 1, 0, 1, 0, 1]
    # fewer matrices than chains, one chain, empty list
    assert plan(4, [640, 320]) == ([0, 1], [0, 1])
    assert plan(1, [640, 320, 960]) == ([0, 1, 2], [0, 0, 0])
    assert plan(2, []) == ([], [])
    assert lib.svdb200_list_plan(ctypes.c_int(0), ctypes.c_size_t(0), None, None, None) != 0
    # three chains, ragged sizes: every matrix once, loads within the largest matrix of each other
    sizes = [64, 1280, 96, 640, 128, 512, 320, 2048, 32]
    order, chain = plan(3, sizes)
    assert sorted(order) == list(range(len(sizes)))
    loads = [sum(sizes[i] for i, c in zip(order, chain) if c == l) for l in range(3)]
    assert max(loads) - min(loads) <= max(sizes)
