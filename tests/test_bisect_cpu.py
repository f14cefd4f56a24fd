"""Sturm-count core of the bisection solver (svdsolver_b200/csrc/bisect_core.h) compiled for the host and checked
against LAPACK: the same header is what bidiag_bisect.cu runs per thread on the GPU."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def bis(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("bis") / "libbis.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-I", os.path.join(ROOT, "svdsolver_b200", "csrc"),
                           os.path.join(ROOT, "tests", "bisect_host.cpp"), "-o", out])
    lib = ctypes.CDLL(out)

    def run(d, e):
        d = np.ascontiguousarray(d, dtype=np.float64)
        e = np.ascontiguousarray(e, dtype=np.float64)
        o = np.zeros(len(d))
        lib.bis_all(d.ctypes.data_as(ctypes.c_void_p), e.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(len(d)),
                    o.ctypes.data_as(ctypes.c_void_p))
        return o
    return run


CASES = {
    "random": lambda r, n: (r.standard_normal(n) * 3, r.standard_normal(n - 1)),
    "positive": lambda r, n: (r.random(n) * 5e3, r.random(n - 1) * 5e3),
    "graded": lambda r, n: (10.0 ** (-np.arange(n) / 10.0), 10.0 ** (-np.arange(n - 1) / 10.0) * 0.5),
    "zeros_e": lambda r, n: (r.standard_normal(n), np.where(np.arange(n - 1) % 7 == 0, 0.0, r.standard_normal(n - 1))),
    "zeros_d": lambda r, n: (np.where(np.arange(n) % 9 == 3, 0.0, r.standard_normal(n)), r.standard_normal(n - 1)),
    "clustered": lambda r, n: (np.ones(n), np.full(n - 1, 1e-3)),
    "huge": lambda r, n: (r.standard_normal(n) * 1e150, r.standard_normal(n - 1) * 1e150),
    "tiny": lambda r, n: (r.standard_normal(n) * 1e-150, r.standard_normal(n - 1) * 1e-150),
}


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("n", [2, 3, 17, 200])
def test_bisection_core_vs_lapack(bis, name, n):
    rng = np.random.default_rng(n)
    d, e = CASES[name](rng, n)
    ref = np.linalg.svd(np.diag(d) + np.diag(e, 1), compute_uv=False)
    got = bis(d, e)
    assert np.all(np.diff(got) <= 0)                     # descending by construction
    assert np.abs(got - ref).max() <= 1e-14 * ref[0]


def test_bisection_core_all_zero(bis):
    assert np.array_equal(bis(np.zeros(5), np.zeros(4)), np.zeros(5))
