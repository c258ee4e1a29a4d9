"""GPU probe: run the band reduction twice on the same matrix and report where the two band matrices first differ,
plus the deviation from a float64 numpy block-Householder reference for the first panels."""
import sys, json
import numpy as np
sys.path.insert(0, ".")
from eig_b200 import capi
c = capi.Context(0)
n = int(sys.argv[1]); reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
rs = np.random.RandomState(n)
X = rs.randn(n, 2 * n); A = X @ X.T / (2 * n)
w = np.linalg.eigvalsh(A)[::-1]
bands = []
for r in range(reps):
    d, e, band = c.debug_tridiag(A)
    bands.append(band.copy())
    import scipy.linalg as sl
    ab = np.zeros((65, n))
    for k in range(65):
        ab[k, :n - k] = band[:n - k, k]
    err = np.abs(sl.eigvals_banded(ab, lower=True)[::-1] - w).max() / w[0]
    print("rep", r, "band_err", err, flush=True)
for r in range(1, reps):
    diff = np.abs(bands[r] - bands[0]).max(axis=1)
    bad = np.flatnonzero(diff > 1e-9 * np.abs(bands[0]).max())
    print("rep", r, "vs 0: differing band columns:", len(bad), "first", bad[:5], "last", bad[-3:] if len(bad) else [])
# first panel against numpy: after panel 0, band columns 0..63 are final: diag block + R
d0 = np.abs(bands[0][:64, :65]); print("col0..63 norms", np.linalg.norm(bands[0][:64, :65]))
