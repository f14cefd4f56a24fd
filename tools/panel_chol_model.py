"""numpy model of the Cholesky-QR panel with reconstructed Householder vectors (svdsolver_b200/csrc/stage1_panel_chol.cu).

    G  = A^T A                     (double accumulation; one pass over the panel)
    R  = chol(G)                   (upper, positive diagonal)          A = Q R
    Q1 = A1 R^-1                   (top b x b block)
    Q1 - S = L U                   (LU without pivoting; s_i = -sign(pivot_i) chosen during the elimination)
    Y  = [L ; A2 (U R)^-1]         Householder vectors of the QR factorisation with R_hh = S R  (unique given the sign rule
                                   svd_serial.h:194: H x = -sign(x0) ||x|| e1)
    T^-1 = diag(1/tau) + striu(Y^T Y),  tau_j = 2 / (y_j^T y_j)       (from the Gram matrix of the STORED Y)
    V2 = -Y T^T

checked against a plain column-by-column Householder QR, in the working precision of the element type."""
import numpy as np
from panel_blk_model import hh_ref


def chol_panel(A, dtype=np.float64, gram_of_stored=True):
    A = A.astype(dtype); m, b = A.shape
    Ad = A.astype(np.float64)
    G = Ad.T @ Ad
    # Cholesky, upper
    R = np.linalg.cholesky(G).T
    ratio = (np.diag(R) ** 2 / np.diag(G)).min()
    A1 = Ad[:b]
    Rinv = np.linalg.inv(R)
    Q1 = A1 @ Rinv
    # modified LU
    W = Q1.copy(); S = np.zeros(b)
    for i in range(b):
        S[i] = -np.copysign(1.0, W[i, i])
        W[i, i] -= S[i]
        W[i + 1:, i] /= W[i, i]
        W[i + 1:, i + 1:] -= np.outer(W[i + 1:, i], W[i, i + 1:])
    L = np.tril(W, -1) + np.eye(b); U = np.triu(W)
    M1 = np.linalg.inv(U @ R)                       # (U R)^-1, upper triangular
    Y = np.zeros((m, b), dtype)
    Y[:b] = L.astype(dtype)
    Y[b:] = A[b:] @ M1.astype(dtype)                # working precision product
    Rhh = (S[:, None] * R).astype(dtype)
    if gram_of_stored:
        Yd = Y.astype(np.float64); YY = Yd.T @ Yd
    else:
        G2 = Ad[b:].T @ Ad[b:]
        YY = L.T @ L + M1.T @ G2 @ M1
    tau = 2.0 / np.diag(YY)
    Tinv = np.diag(1.0 / tau) + np.triu(YY, 1)
    T = np.linalg.inv(Tinv)
    V2 = (-(Y.astype(np.float64) @ T.T)).astype(dtype)
    return Rhh, Y, tau, T, V2, ratio


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for dtype in (np.float64, np.float32):
        for (m, b, lo_, hi_) in [(200, 32, 0, 5), (64, 64, 0, 5), (1000, 64, 1, 5), (96, 8, -1, 1), (4096, 64, 0, 5), (16384, 64, 0, 5),
                                 (65536, 64, 0, 5), (3000, 32, -1, 1)]:
            A0 = (rng.random((m, b)) * (hi_ - lo_) + lo_).astype(dtype)
            Rr, V, tr = hh_ref(A0.astype(np.float64))
            for gs in (True, False):
                Rhh, Y, tau, T, V2, ratio = chol_panel(A0, dtype, gs)
                Yd = Y.astype(np.float64)
                H = np.eye(m) - Yd @ T @ Yd.T if m <= 4096 else None
                orth = np.abs(H.T @ H - np.eye(m)).max() if H is not None else np.nan
                # orthogonality through the b x b identity  (I - Y T Y^T)^T (I - Y T Y^T) = I  <=>  T + T^T = T^T Y^T Y T
                YY = Yd.T @ Yd
                orth2 = np.abs(T + T.T - T.T @ YY @ T).max()
                # residual: A - H [R; 0]
                top = np.zeros((m, b)); top[:b] = np.triu(Rhh.astype(np.float64))
                HR = top - Yd @ (T @ (Yd.T @ top))
                resid = np.abs(HR - A0.astype(np.float64)).max() / np.abs(A0).max()
                print(f"{np.dtype(dtype).name} m={m:6d} b={b:2d} gram_of_stored={int(gs)} min pivot ratio {ratio:.3f}  "
                      f"|R-Rref| {np.abs(np.triu(Rhh) - np.triu(Rr[:b])).max() / np.abs(Rr[:b]).max():.2e}  "
                      f"|Y-Vref| {np.abs(Yd - V).max():.2e}  |tau-ref| {np.abs(tau - tr).max():.2e}  orth {orth:.2e}/{orth2:.2e}  resid {resid:.2e}")
