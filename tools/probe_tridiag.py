"""GPU probe: which stage of the two-stage tridiagonalisation loses the spectrum as n grows."""
import sys, time, json
import numpy as np
import scipy.linalg as sl
sys.path.insert(0, ".")
from eig_b200 import capi
c = capi.Context(0)
for n in [int(v) for v in sys.argv[1:]]:
    rs = np.random.RandomState(n)
    M = 2 * n
    X = rs.randn(n, M)
    A = X @ X.T / M
    w = np.linalg.eigvalsh(A)[::-1]
    d, e, band = c.debug_tridiag(A)
    ab = np.zeros((65, n))
    for k in range(65):
        ab[k, :n - k] = band[:n - k, k]
    wb = sl.eigvals_banded(ab, lower=True)[::-1]
    wt = sl.eigvalsh_tridiagonal(d, e)[::-1] if np.isfinite(d).all() and np.isfinite(e).all() else np.full(n, np.nan)
    # locate the first band column whose content deviates: compare trace and Frobenius norm too
    fro_A = np.sqrt((A * A).sum()); fro_B = np.sqrt((ab[0] ** 2).sum() + 2 * (ab[1:] ** 2).sum())
    print(json.dumps(dict(n=n, band_err=float(np.abs(wb - w).max() / w[0]), tri_err=float(np.nanmax(np.abs(wt - w)) / w[0]),
                          tri_vs_band=float(np.nanmax(np.abs(wt - wb)) / w[0]), fro_A=fro_A, fro_band=fro_B,
                          beyond64=float(np.abs(band[:, 65:]).max()), nonfinite=int((~np.isfinite(d)).sum() + (~np.isfinite(e)).sum()))), flush=True)
