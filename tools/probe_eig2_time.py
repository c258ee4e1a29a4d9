"""two-stage path only, for ncu launch lists / timing at larger n"""
import sys, time, json
import numpy as np
sys.path.insert(0, ".")
import torch
from eig_b200 import capi
c = capi.Context(0)
c.set_option("eig_method", 2)
for n in [int(x) for x in sys.argv[1:]]:
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    X = torch.randn(n, 2 * n, device="cuda", dtype=torch.float64, generator=g)
    A = (X @ X.T / (2 * n)).cpu().numpy(); del X
    for rep in range(2):
        t0 = time.time(); lam, vec = c.eigvecs(A, nvec=10); dt = time.time() - t0
    tm = c.timings()
    r = float(np.abs(A @ vec.T - vec.T * lam[:10]).max())
    print(json.dumps({"n": n, "wall_s": dt, "resid": r, "lam0": lam[0], "trace_err": float(abs(lam.sum() - np.trace(A)) / np.trace(A)),
                      **{k: tm[k] for k in ("tridiag_ms", "bisect_ms", "vectors_ms", "chfsi_iters", "chfsi_matvecs")}}), flush=True)
