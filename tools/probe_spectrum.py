"""GPU probe: sanity of the full spectrum from the two-stage tridiagonalisation at scale.
sum(lambda) must equal trace(XTX/y) = n - 1; the top values must match the Ritz values of the subspace iteration."""
import json, sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from eig_b200 import capi, synth
c = capi.Context(0)
for spec in sys.argv[1:]:
    parts = spec.split(":")
    N, M = (int(v) for v in parts[0].split("x"))
    npops = int(parts[1]) if len(parts) > 1 else 1
    rl = synth.rlen_for(N)
    buf = torch.empty((M, rl), dtype=torch.uint8, device="cuda")
    c.synth_packed_device(buf.data_ptr(), M, rl, N, seed=1, missing=0.0, npops=npops, delta=0.05 if npops > 1 else 0.0)
    c.adopt_packed_device(buf.data_ptr(), M, rl, N); c.set_rows(None)
    r = c.grm(want_snp=False)
    lam, vec = c.eig(6)
    tm = c.timings()
    bad = int((~np.isfinite(lam)).sum())
    rec = dict(N=N, M=M, npops=npops, y=r["y"], sum_lam=float(np.nansum(lam)), want_sum=N - 1, top=lam[:5].tolist(), bottom=lam[-3:].tolist(), nonfinite=bad,
               tridiag_ms=tm["tridiag_ms"], method=tm["eig_method"])
    if N <= 8000:
        X = c.grm(want_xtx=True, want_snp=False)["XTX"]
        rec["ritz"] = np.einsum("ij,ij->i", vec @ X, vec)[:5].tolist()
        w = np.linalg.eigvalsh(X)[::-1]
        rec["max_rel_err"] = float((np.abs(lam - w) / np.maximum(np.abs(w), 1e-6 * w[0])).max())
    print(json.dumps(rec), flush=True)
    del buf
