"""GPU probe for the two-stage eigensolver pieces (run under gpurun)."""
import sys, time, json
import numpy as np
import scipy.linalg as sl
sys.path.insert(0, ".")
from eig_b200 import capi

c = capi.Context(0)
rng = np.random.default_rng(0)
sizes = [int(x) for x in (sys.argv[1:] or ["700", "1537", "3000"])]
for n in sizes:
    M = 4 * n
    X = rng.standard_normal((n, M))
    lab = rng.integers(0, 4, n); X += 0.3 * rng.standard_normal((4, M))[lab]
    A = X @ X.T / M
    t0 = time.time(); w = np.linalg.eigvalsh(A)[::-1]; tnp = time.time() - t0
    rec = {"n": n, "numpy_eigvalsh_s": tnp}
    try:
        t0 = time.time(); d, e, band = c.debug_tridiag(A); rec["tridiag_s"] = time.time() - t0
        # band check
        ab = np.zeros((65, n))
        for k in range(65):
            ab[k, :n - k] = band[:n - k, k]
        wb = sl.eigvals_banded(ab, lower=True)[::-1]
        rec["band_err"] = float(np.abs(wb - w).max() / w[0])
        wt = sl.eigvalsh_tridiagonal(d, e)[::-1]
        rec["tri_err"] = float(np.abs(wt - w).max() / w[0])
    except Exception as ex:
        rec["tridiag_exc"] = str(ex)[:300]
    for method in (1, 2):
        c.set_option("eig_method", method)
        try:
            t0 = time.time(); lam, vec = c.eigvecs(A, nvec=10); dt = time.time() - t0
            t0 = time.time(); lam, vec = c.eigvecs(A, nvec=10); dt = time.time() - t0
            tm = c.timings()
            rec["m%d_s" % method] = dt
            rec["m%d_tm" % method] = {k: tm[k] for k in ("tridiag_ms", "bisect_ms", "vectors_ms", "eig_method", "chfsi_iters", "chfsi_matvecs")}
            rec["m%d_lam_err" % method] = float(np.abs(lam - w).max() / w[0])
            wv, U = np.linalg.eigh(A) if n <= 4000 else (None, None)
            if U is not None:
                U = U[:, ::-1][:, :10].T
                rec["m%d_cos_err" % method] = float(np.abs(np.abs(np.sum(U * vec, axis=1)) - 1).max())
            rec["m%d_resid" % method] = float(np.abs(A @ vec.T - vec.T * lam[:10]).max())
        except Exception as ex:
            rec["m%d_exc" % method] = str(ex)[:300]
    print(json.dumps(rec), flush=True)
