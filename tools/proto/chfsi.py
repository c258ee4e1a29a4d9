"""numpy prototype: Chebyshev-filtered subspace iteration for the leading k eigenpairs of a symmetric PSD matrix."""
import numpy as np


def chol_qr(W):
    """shifted CholQR + CholQR2 (three passes); returns orthonormal basis"""
    n, m = W.shape
    for it in range(3):
        G = W.T @ W
        if it == 0:
            G = G + np.eye(m) * (11 * (m * n + m * (m + 1)) * 1.1e-16 * np.trace(G))
        R = np.linalg.cholesky(G).T
        W = W @ np.linalg.inv(R)
    return W


def chfsi(A, k, m=None, tol=1e-13, maxouter=200, dmax=40, seed=1):
    n = A.shape[0]
    m = m or min(n, max(2 * k, k + 16))
    rng = np.random.default_rng(seed)
    V = chol_qr(rng.standard_normal((n, m)))
    lo = 0.0
    nmat = 0
    anorm = None
    for outer in range(maxouter):
        AV = A @ V; nmat += 1
        H = V.T @ AV; H = 0.5 * (H + H.T)
        th, Z = np.linalg.eigh(H); th = th[::-1]; Z = Z[:, ::-1]
        V = V @ Z; AV = AV @ Z
        R = AV - V * th
        res = np.linalg.norm(R, axis=0)
        anorm = max(abs(th[0]), abs(th[-1]))
        if res[:k].max() <= tol * anorm:
            break
        # filter: damp [lo, cut], cut = smallest Ritz value in the block
        cut = th[-1]; lo = min(lo, cut)
        c = 0.5 * (cut + lo); e = 0.5 * (cut - lo)
        if e <= 0: e = 1e-3 * anorm
        xi1 = (th[0] - c) / e
        # degree: keep the amplification spread T_d(xi1) below 1e8
        d = int(max(2, min(dmax, np.floor(np.arccosh(1e8) / max(np.arccosh(max(xi1, 1.0 + 1e-12)), 1e-6)))))
        sigma = e / (th[0] - c); sigma1 = sigma
        Y0 = V
        Y1 = (AV - c * V) * (sigma1 / e)
        for j in range(2, d + 1):
            sn = 1.0 / (2.0 / sigma1 - sigma)
            Y2 = (A @ Y1 - c * Y1) * (2.0 * sn / e) - (sigma * sn) * Y0; nmat += 1
            Y0, Y1, sigma = Y1, Y2, sn
        V = chol_qr(Y1)
    return th[:k], V[:, :k], outer + 1, nmat, res[:k].max() / anorm


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for n, M, npop in [(600, 4000, 1), (600, 4000, 5), (1500, 12000, 1)]:
        X = rng.standard_normal((n, M))
        if npop > 1:
            lab = rng.integers(0, npop, n)
            X += 0.5 * rng.standard_normal((npop, M))[lab]
        A = X @ X.T / M
        w, U = np.linalg.eigh(A); w = w[::-1]; U = U[:, ::-1]
        for m in (32, 48):
            th, V, outer, nmat, r = chfsi(A, 10, m=m)
            cos = np.abs(np.sum(V * U[:, :10], axis=0))
            print(n, M, npop, "m", m, "outer", outer, "matvecs", nmat, "res", r, "eval err", np.abs(th - w[:10]).max() / w[0], "1-cos", (1 - cos).max())
