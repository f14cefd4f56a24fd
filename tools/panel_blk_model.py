"""numpy model of the blocked panel factorisation (svdsolver_b200/csrc/stage1_panel_blk.cu): same formulas, one "CTA",
checked against a plain column-by-column Householder QR with the reference's sign convention."""
import numpy as np


def hh_ref(A):
    A = A.copy(); m, b = A.shape
    V = np.zeros((m, b)); taus = np.zeros(b)
    for j in range(b):
        x = A[j:, j].copy()
        nrm = np.sqrt((x * x).sum())
        sgn = -np.copysign(1.0, x[0])
        u1 = x[0] - sgn * nrm
        alpha = 1.0 / u1; tau = -sgn * u1 / nrm
        v = x * alpha; v[0] = 1.0
        A[j:, j:] -= tau * np.outer(v, v @ A[j:, j:])
        A[j + 1:, j] = v[1:]
        V[j:, j] = v; taus[j] = tau
    return A, V, taus


def blk(A, C=8, guard=1.0 / 64, dtype=np.float64):
    A = A.astype(dtype).copy(); m, b = A.shape
    Gm = np.zeros((b, b), dtype); taus = np.zeros(b, dtype)
    j = 0; rounds = 0
    while j < b:
        lo = j + C
        D = A[lo:, j:j + C].T @ A[lo:, :] if lo < m else np.zeros((C, b), dtype)     # C x b
        if D.shape[0] < C:
            D = np.vstack([D, np.zeros((C - D.shape[0], b), dtype)])
        Top = np.zeros((C, b), dtype); nt = min(C, m - j); Top[:nt] = A[j:j + nt, :]
        Cm = np.zeros((C, b), dtype)
        done = 0
        for i in range(C):
            ji = j + i
            if ji >= b: break
            mv = -Cm[:, ji].copy(); mv[i] += 1; mv[i + 1:] = 0
            tj = Top[:, ji].copy()
            gl = mv @ D                                  # m^T D[:,k]
            gv = np.array([gl[min(j + p, b - 1)] if p <= i else 0 for p in range(C)], dtype)
            s = gl - gv @ Cm
            dT = tj[i + 1:] @ Top[i + 1:, :]
            sji = s[ji]; dii = D[i, ji]
            if i > 0 and not (sji >= guard * dii):
                break
            nrm = np.sqrt(sji + (tj[i:] ** 2).sum()); x0 = tj[i]
            sgn = -np.copysign(1.0, x0); u1 = float(x0) - sgn * float(nrm)
            alpha = dtype(1.0 / u1); tau = dtype(-sgn * u1 / float(nrm)); beta = dtype(sgn * float(nrm))
            dot = Top[i, :] + alpha * (dT + s)
            for col in range(b):
                if col > ji:
                    f = tau * dot[col]; fa = f * alpha
                    Top[i, col] -= f; Top[i + 1:, col] -= fa * tj[i + 1:]; Cm[:i + 1, col] += fa * mv[:i + 1]
                elif col == ji:
                    Top[i, col] = beta; Top[i + 1:, col] = alpha * tj[i + 1:]
                    e = np.zeros(C, dtype); e[i] = 1
                    Cm[:, col] = e - alpha * mv; Cm[i + 1:, col] = 0
                    taus[ji] = tau
                else:
                    Gm[col, ji] = dot[col]
            done = i + 1
        # pass
        K = -Cm.copy(); keep = np.ones(b, dtype)
        for col in range(j, j + done):
            K[:, col] = -Cm[:, col]; K[col - j, col] += 1; keep[col] = 0
        if lo < m:
            X = A[lo:, j:j + C].copy()
            if X.shape[1] < C:
                X = np.hstack([X, np.zeros((X.shape[0], C - X.shape[1]), dtype)])
            A[lo:, :] = A[lo:, :] * keep + X @ K
        A[j:j + nt, :] = Top[:nt]
        j += done; rounds += 1
    return A, Gm, taus, rounds


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for (m, b, lo_, hi_) in [(200, 32, 0, 5), (64, 64, 0, 5), (1000, 64, 1, 5), (96, 8, -1, 1), (32, 32, 0, 5), (500, 16, 0, 5)]:
        A0 = rng.random((m, b)) * (hi_ - lo_) + lo_
        Rr, V, tr = hh_ref(A0)
        Ab, Gm, tb, rounds = blk(A0)
        errR = np.abs(np.triu(Ab[:b]) - np.triu(Rr[:b])).max() / np.abs(Rr).max()
        errV = np.abs(np.tril(Ab, -1) - np.tril(Rr, -1)).max()
        errT = np.abs(tb - tr).max()
        G = np.triu(V.T @ V, 1)
        errG = np.abs(np.triu(Gm, 1) - G).max()
        print(f"m={m} b={b} rounds={rounds} errR={errR:.2e} errV={errV:.2e} errTau={errT:.2e} errGram={errG:.2e}")
    # float32 with the float guard
    A0 = (rng.random((4096, 64)) * 5).astype(np.float32)
    Rr, V, tr = hh_ref(A0.astype(np.float64))
    Ab, Gm, tb, rounds = blk(A0, guard=0.25, dtype=np.float32)
    print("f32 4096x64 rounds", rounds, "errR", np.abs(np.triu(Ab[:64]) - np.triu(Rr[:64])).max() / np.abs(Rr).max(), "errV", np.abs(np.tril(Ab, -1) - np.tril(Rr, -1)).max())
    # nearly dependent columns: the guard must kick in and the result stay accurate
    A0 = rng.random((300, 16)); A0[:, 3] = A0[:, 1] + 1e-9 * rng.random(300); A0[:, 9] = A0[:, 8] * 2 + 1e-7 * rng.random(300)
    Rr, V, tr = hh_ref(A0)
    Ab, Gm, tb, rounds = blk(A0)
    print("dependent cols rounds", rounds, "errR", np.abs(np.triu(Ab[:16]) - np.triu(Rr[:16])).max() / np.abs(Rr).max())
    Q = np.eye(300)
    # orthogonality of the blocked reflectors
    Vb = np.tril(Ab, -1)[:, :16] + np.eye(300, 16)
    for jj in range(16):
        Q = Q @ (np.eye(300) - tb[jj] * np.outer(Vb[:, jj], Vb[:, jj]))
    print("orth", np.abs(Q.T @ Q - np.eye(300)).max(), "recon", np.abs(Q @ np.vstack([np.triu(Ab[:16]), np.zeros((284, 16))]) - A0).max())


def blk2(A, C=8, guard=1.0 / 64, dtype=np.float64):
    """Same blocked factorisation with the DOWNDATING form of the small algebra (what the kernel runs): instead of
    re-deriving every dot product from the exchanged D through the coefficient matrix, the C x b matrix
    S[p][k] = (current column j+p, lo part)^T (current column k, lo part) is carried along and updated after each reflector
    (rank-one formulas, no dependent chains); norms and dots of the next pivot are then single entries of S."""
    A = A.astype(dtype).copy(); m, b = A.shape
    Gm = np.zeros((b, b), dtype); taus = np.zeros(b, dtype)
    j = 0; rounds = 0
    while j < b:
        lo = j + C
        D = A[lo:, j:j + C].T @ A[lo:, :] if lo < m else np.zeros((C, b), dtype)
        if D.shape[0] < C:
            D = np.vstack([D, np.zeros((C - D.shape[0], b), dtype)])
        Top = np.zeros((C, b), dtype); nt = min(C, m - j); Top[:nt] = A[j:j + nt, :]
        Cm = np.zeros((C, b), dtype)
        S = D.copy()
        d0 = np.array([D[p, min(j + p, b - 1)] for p in range(C)])
        done = 0
        for i in range(C):
            ji = j + i
            if ji >= b: break
            sji = S[i, ji]
            if i > 0 and not (sji >= guard * d0[i]):
                break
            tj = Top[:, ji].copy()
            nrm = np.sqrt(sji + (tj[i:] ** 2).sum()); x0 = tj[i]
            sgn = -np.copysign(1.0, x0); u1 = float(x0) - sgn * float(nrm)
            alpha = dtype(1.0 / u1); tau = dtype(-sgn * u1 / float(nrm)); beta = dtype(sgn * float(nrm))
            mv = -Cm[:, ji].copy(); mv[i] += 1; mv[i + 1:] = 0
            xa = S[i, :].copy()                              # x_lo^T (column k, lo)
            dT = tj[i + 1:] @ Top[i + 1:, :]
            dot = Top[i, :] + alpha * (dT + xa)
            fa = np.zeros(b, dtype)
            for col in range(b):
                if col > ji:
                    f = tau * dot[col]; fa[col] = f * alpha
                    Top[i, col] -= f; Top[i + 1:, col] -= fa[col] * tj[i + 1:]; Cm[:i + 1, col] += fa[col] * mv[:i + 1]
                elif col == ji:
                    Top[i, col] = beta; Top[i + 1:, col] = alpha * tj[i + 1:]
                    e = np.zeros(C, dtype); e[i] = 1
                    Cm[:, col] = e - alpha * mv; Cm[i + 1:, col] = 0
                    taus[ji] = tau
                else:
                    Gm[col, ji] = dot[col]
            # downdate S for the rows of the later pivots
            for p in range(i + 1, C):
                q = j + p
                if q >= b: break
                for col in range(b):
                    if col == ji:
                        S[p, col] = alpha * (xa[q] - fa[q] * sji)
                    else:
                        S[p, col] = S[p, col] - fa[col] * xa[q] - fa[q] * (xa[col] - fa[col] * sji)
            done = i + 1
        K = -Cm.copy(); keep = np.ones(b, dtype)
        for col in range(j, j + done):
            K[:, col] = -Cm[:, col]; K[col - j, col] += 1; keep[col] = 0
        if lo < m:
            X = A[lo:, j:j + C].copy()
            if X.shape[1] < C:
                X = np.hstack([X, np.zeros((X.shape[0], C - X.shape[1]), dtype)])
            A[lo:, :] = A[lo:, :] * keep + X @ K
        A[j:j + nt, :] = Top[:nt]
        j += done; rounds += 1
    return A, Gm, taus, rounds
