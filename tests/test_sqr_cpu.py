"""Implicit shifted QR core (svdsolver_b200/csrc/sqr_core.h) compiled for the host: the multishift iteration the CUDA
kernel pipelines (P sweeps per pass, shifts = singular values of the trailing P x P block) converges and agrees with
LAPACK; the same header is what bidiag_sqr.cu runs on the GPU."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def sqr(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("sqr") / "libsqr.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-I", os.path.join(ROOT, "svdsolver_b200", "csrc"),
                           os.path.join(ROOT, "tests", "sqr_host.cpp"), "-o", out])
    lib = ctypes.CDLL(out)
    lib.sqr_all.restype = ctypes.c_longlong

    def run(d, e, P):
        d = np.ascontiguousarray(d, dtype=np.float64)
        e = np.ascontiguousarray(e, dtype=np.float64)
        o = np.zeros(len(d))
        passes = ctypes.c_longlong(0)
        sw = lib.sqr_all(d.ctypes.data_as(ctypes.c_void_p), e.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(len(d)),
                         o.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(P), ctypes.byref(passes))
        return o, int(sw), int(passes.value)
    return run


CASES = {
    "random": lambda r, n: (r.standard_normal(n) * 3, r.standard_normal(n - 1)),
    "positive": lambda r, n: (r.random(n) * 5, r.random(n - 1) * 5),
    "graded": lambda r, n: (10.0 ** (-np.arange(n) / 20.0), 10.0 ** (-np.arange(n - 1) / 20.0) * 0.5),
    "zeros_e": lambda r, n: (r.standard_normal(n), np.where(np.arange(n - 1) % 7 == 0, 0.0, r.standard_normal(n - 1))),
    "zeros_d": lambda r, n: (np.where(np.arange(n) % 9 == 3, 0.0, r.standard_normal(n)), r.standard_normal(n - 1)),
    "clustered": lambda r, n: (np.ones(n), np.full(n - 1, 1e-3)),
}


@pytest.mark.parametrize("P", [1, 8, 32, 256])
@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("n", [2, 3, 17, 200, 600])
def test_shifted_qr_core_vs_lapack(sqr, name, n, P):
    rng = np.random.default_rng(n)
    d, e = CASES[name](rng, n)
    ref = np.linalg.svd(np.diag(d) + np.diag(e, 1), compute_uv=False)
    got, sweeps, passes = sqr(d, e, P)
    assert sweeps >= 0, "did not converge"
    assert np.all(np.diff(got) <= 0)
    assert np.abs(got - ref).max() <= 1e-13 * ref[0]          # measured: 1e-15 .. 6e-14 (n = 2000), growing with the sweep count
    # a shifted iteration needs a few sweeps per singular value (zero-shift QR: ~n log(1/tol) sweeps in total)
    assert sweeps <= 12 * n + 64
