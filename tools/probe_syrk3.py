import sys
import numpy as np
sys.path.insert(0, ".")
from eig_b200 import capi
c = capi.Context(0)
n = int(sys.argv[1])
rs = np.random.RandomState(n)
T = rs.randn(64, n); T -= T.mean(axis=1, keepdims=True)
want = T.T @ T
nt = (n + 127) // 128
shown = 0
for rep in range(3):
    y, X = c.grm_dense([T], n, want_xtx=True)
    got = X * y
    err = np.abs(got - want) > 1e-9 * np.abs(want).max()
    for I in range(nt):
        for J in range(2 * I + 2):
            blk = err[I * 128:(I + 1) * 128, J * 64:(J + 1) * 64]
            if not blk.any() or shown >= 6:
                continue
            shown += 1
            rows = np.flatnonzero(blk.any(axis=1)); cols = np.flatnonzero(blk.any(axis=0))
            r0, c0 = I * 128, J * 64
            # per k-row contributions for the bad sub-block: d = got - want = sum_k coef_k * T[k,row]*T[k,col]
            rr = rows[:16]; cc = cols[:32]
            D = (got - want)[np.ix_(r0 + rr, c0 + cc)].reshape(-1)
            B = np.stack([(T[k, r0 + rr][:, None] * T[k, c0 + cc][None, :]).reshape(-1) for k in range(64)], axis=1)
            coef, res, *_ = np.linalg.lstsq(B, D, rcond=None)
            resid = np.linalg.norm(B @ coef - D) / np.linalg.norm(D)
            nz = np.flatnonzero(np.abs(coef) > 1e-6)
            print("rep %d tile (I=%d,J=%d) rows %s cols %d-%d : k-rows with nonzero coef %s coefs %s fit resid %.2e" %
                  (rep, I, J, rows.tolist()[:12], cols.min(), cols.max(), nz.tolist(), np.round(coef[nz], 3).tolist(), resid), flush=True)
