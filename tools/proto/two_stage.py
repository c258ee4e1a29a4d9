"""numpy prototype of the two-stage tridiagonalisation (index conventions for the CUDA kernels).
stage 1: dense -> band (bandwidth b) with panel Householder QR + two-sided block update
stage 2: band -> tridiagonal, column-wise bulge chasing (Lang), lower band storage AB[col][d], d = row-col in [0, 2b)
"""
import numpy as np


def house(x):
    """dlarfg: returns v (v[0]=1), tau, beta with (I - tau v v^T) x = beta e1"""
    alpha = x[0]; s = float(x[1:] @ x[1:])
    v = x.copy(); v[0] = 1.0
    if s == 0.0:
        return v * 0 + np.eye(len(x))[0], 0.0, alpha
    nrm = np.sqrt(alpha * alpha + s)
    beta = -nrm if alpha >= 0 else nrm
    tau = (beta - alpha) / beta
    v[1:] = x[1:] / (alpha - beta)
    return v, tau, beta


def panel_qr_gramrow(P):
    """Householder QR with ONE reduction per column (row k of the Gram matrix of the current panel)."""
    P = P.copy(); npan, b = P.shape
    nr = min(b, npan - 1)
    V = np.zeros((npan, b)); taus = np.zeros(b)
    for k in range(nr):
        g = P[k + 1:, k] @ P[k + 1:, k:]          # the grid-wide reduction
        alpha = P[k, k]; sigma = g[0]
        if sigma == 0.0:
            V[k, k] = 1.0; continue                 # tau = 0
        nrm = np.sqrt(alpha * alpha + sigma); beta = -nrm if alpha >= 0 else nrm
        tau = (beta - alpha) / beta; scal = 1.0 / (alpha - beta)
        vtp = P[k, k + 1:] + scal * g[1:]
        P[k, k + 1:] -= tau * vtp
        P[k + 1:, k + 1:] -= np.outer(P[k + 1:, k] * (tau * scal), vtp)
        V[k, k] = 1.0; V[k + 1:, k] = P[k + 1:, k] * scal
        P[k, k] = beta; P[k + 1:, k] = 0.0
        taus[k] = tau
    # T factor (forward, columnwise): T[:k,k] = -tau_k T[:k,:k] V[:, :k]^T v_k
    G = V.T @ V
    T = np.zeros((b, b))
    for k in range(b):
        T[k, k] = taus[k]
        if k and taus[k] != 0.0:
            T[:k, k] = -taus[k] * (T[:k, :k] @ G[:k, k])
    return P, V, T


def stage1(A, b):
    A = A.copy(); n = A.shape[0]
    j = 0
    while n - j - b >= 2:
        r0 = j + b
        P = A[r0:, j:j + b]
        R, V, T = panel_qr_gramrow(P)
        A[r0:, j:j + b] = R; A[j:j + b, r0:] = R.T
        A22 = A[r0:, r0:]
        W = A22 @ V
        S = V.T @ W
        Y = W @ T
        M = T.T @ S @ T
        Z = Y - 0.5 * V @ M
        A[r0:, r0:] = A22 - V @ Z.T - Z @ V.T
        j += b
    return A


def to_band(A, b):
    n = A.shape[0]
    AB = np.zeros((n, 2 * b))
    for col in range(n):
        for d in range(0, min(b, n - 1 - col) + 1):
            AB[col, d] = A[col + d, col]
    return AB


def chase(AB, b):
    """AB[col][d] = A[col+d][col]; returns d, e"""
    AB = AB.copy(); n = AB.shape[0]

    def get(r, c):
        return AB[c, r - c]

    def blk(r0, nr, c0, nc):
        return np.array([[AB[c0 + c, r0 + i - (c0 + c)] for c in range(nc)] for i in range(nr)])

    def put(r0, nr, c0, nc, B, lower_only=False):
        for i in range(nr):
            for c in range(nc):
                if lower_only and r0 + i < c0 + c:
                    continue
                AB[c0 + c, r0 + i - (c0 + c)] = B[i, c]

    def symblk(r0, m):
        D = np.zeros((m, m))
        for i in range(m):
            for c in range(i + 1):
                D[i, c] = D[c, i] = AB[r0 + c, i - c]
        return D

    for s in range(n - 2):
        # type 1
        r0 = s + 1; la = min(b, n - r0)
        x = np.array([AB[s, 1 + i] for i in range(la)])
        v, tau, beta = house(x)
        AB[s, 1] = beta
        for i in range(1, la):
            AB[s, 1 + i] = 0.0
        D = symblk(r0, la)
        H = np.eye(la) - tau * np.outer(v, v)
        put(r0, la, r0, la, H @ D @ H, lower_only=True)
        while True:
            r1 = r0 + la
            lb = min(b, n - r1)
            if lb <= 0:
                break
            B = blk(r1, lb, r0, la)            # rows r1.., cols r0..
            B = B - tau * np.outer(B @ v, v)   # right-apply previous reflector
            v2, tau2, beta2 = house(B[:, 0].copy())
            B = B - tau2 * np.outer(v2, v2 @ B)
            B[0, 0] = beta2; B[1:, 0] = 0.0
            put(r1, lb, r0, la, B)
            D = symblk(r1, lb)
            H = np.eye(lb) - tau2 * np.outer(v2, v2)
            put(r1, lb, r1, lb, H @ D @ H, lower_only=True)
            r0, la, v, tau = r1, lb, v2, tau2
    return AB[:, 0].copy(), AB[:n - 1, 1].copy()


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for n, b in [(37, 4), (64, 8), (101, 8), (50, 16), (20, 8), (18, 8), (9, 4)]:
        X = rng.standard_normal((n, 3 * n)); A = X @ X.T / n
        w = np.linalg.eigvalsh(A)
        A1 = stage1(A, b)
        AB = to_band(A1, b)
        # band check: eigenvalues of the band matrix
        Bm = np.zeros((n, n))
        for col in range(n):
            for d in range(0, min(b, n - 1 - col) + 1):
                Bm[col + d, col] = Bm[col, col + d] = AB[col, d]
        print(n, b, "band err", np.abs(np.linalg.eigvalsh(Bm) - w).max(), end=" ")
        d, e = chase(AB, b)
        T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
        print("tri err", np.abs(np.linalg.eigvalsh(T) - w).max())
