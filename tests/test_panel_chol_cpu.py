"""CPU checks of the algorithm behind the Cholesky-QR stage-1 panel (svdsolver_b200/csrc/stage1_panel_chol.cu) on its numpy
model (tools/panel_chol_model.py): the Householder factorisation rebuilt from the Cholesky factor is THE Householder QR
factorisation with the reference's sign rule (svd_serial.h:194-201), and the identities the kernel relies on hold:
    A1 - S R = L U~ (LU of (Q1 - S) R),   T^-1 = diag(Y^T Y)/2 + striu(Y^T Y) = -L^T S U^-1,   U^-1 = R U~^-1.
The GPU kernels are compared with the per-column kernels and the oracle in tests/test_gpu_parity_large.py."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from panel_blk_model import hh_ref  # noqa: E402
from panel_chol_model import chol_panel  # noqa: E402


@pytest.mark.parametrize("m,b,lo,hi", [(200, 32, 0, 5), (130, 64, 0, 5), (1000, 64, 1, 5), (96, 8, -1, 1), (3000, 16, -1, 1)])
@pytest.mark.parametrize("stored", [True, False])
def test_reconstructed_factorisation_is_the_householder_one(m, b, lo, hi, stored):
    rng = np.random.default_rng(m + b)
    a = rng.random((m, b)) * (hi - lo) + lo
    r_ref, v_ref, tau_ref = hh_ref(a)
    rhh, y, tau, t, v2, ratio = chol_panel(a, np.float64, stored)
    assert ratio > 1e-3                                                    # these panels are inside the kernel's guard
    scale = np.abs(r_ref[:b]).max()
    assert np.abs(np.triu(rhh) - np.triu(r_ref[:b])).max() <= 1e-12 * scale     # same R, every sign
    assert np.abs(y - v_ref).max() <= 1e-12                                 # same Householder vectors
    assert np.abs(tau - tau_ref).max() <= 1e-12
    # Q = I - Y T Y^T is orthogonal and maps [R; 0] back to the panel
    yy = y.T @ y
    assert np.abs(t + t.T - t.T @ yy @ t).max() <= 1e-12
    top = np.zeros((m, b)); top[:b] = np.triu(rhh)
    assert np.abs(top - y @ (t @ (y.T @ top)) - a).max() <= 1e-11 * np.abs(a).max()
    assert np.abs(v2 + y @ t.T).max() <= 1e-12                              # V2 = V S^T with S = -T


def test_identities_used_by_the_algebra_kernel():
    rng = np.random.default_rng(7)
    m, b = 500, 32
    a = rng.random((m, b)) * 5
    g = a.T @ a
    r = np.linalg.cholesky(g).T
    w = a[:b].copy(); s = np.zeros(b)
    for i in range(b):                                                      # LU of A1 - S R, sign chosen while eliminating
        s[i] = -np.copysign(1.0, w[i, i])
        w[i, i:] -= s[i] * r[i, i:]
        assert abs(w[i, i]) >= r[i, i]                                      # |pivot| >= R_ii: no growth
        w[i + 1:, i] /= w[i, i]
        w[i + 1:, i + 1:] -= np.outer(w[i + 1:, i], w[i, i + 1:])
    l = np.tril(w, -1) + np.eye(b); ut = np.triu(w)
    assert np.abs(l @ ut - (a[:b] - s[:, None] * r)).max() <= 1e-12 * np.abs(a).max()
    m1 = np.linalg.inv(ut)
    y = np.vstack([l, a[b:] @ m1])
    yy = y.T @ y
    tinv = np.diag(np.diag(yy) / 2) + np.triu(yy, 1)
    ui = r @ m1                                                             # U^-1 with U = U~ R^-1
    tinv2 = -(l.T * s) @ ui
    assert np.abs(np.tril(tinv2, -1)).max() == 0.0                          # product of two upper triangular matrices
    assert np.abs(tinv - tinv2).max() <= 1e-12 * np.abs(tinv).max()


def test_pivot_ratio_flags_the_first_row_panel_of_a_mean_shifted_matrix():
    """Why the guard trips once per matrix on the benchmark's U[0,5) inputs: after the first QR step the common
    mean of the columns sits in one row of the trailing matrix, and the first ROW panel comes out nearly rank deficient
    (pivot ratio ~ 3 / n); every other panel is far from the guard."""
    rng = np.random.default_rng(0)
    n, b = 512, 32
    a = rng.random((n, n)) * 5

    def ratio(p):
        g = p.T @ p
        r = np.linalg.cholesky(g).T
        return float((np.diag(r) ** 2 / np.diag(g)).min())

    q, _ = np.linalg.qr(a[:, :b], mode="complete")
    a = q.T @ a
    first_lq = ratio(a[:b, b:].T)
    assert first_lq < 0.05
    q2, _ = np.linalg.qr(a[:b, b:].T, mode="complete")
    a[:, b:] = a[:, b:] @ q2
    assert ratio(a[b:, b:2 * b]) > 0.5                                      # the next column panel is well conditioned
