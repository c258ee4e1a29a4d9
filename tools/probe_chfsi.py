"""GPU probe: leading eigenvectors on STRUCTURED data at scale (outlier eigenvalues far above the bulk)."""
import json, sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from eig_b200 import capi, synth
c = capi.Context(0)
for spec in sys.argv[1:]:
    N, M = (int(v) for v in spec.split("x"))
    rl = synth.rlen_for(N)
    buf = torch.empty((M, rl), dtype=torch.uint8, device="cuda")
    c.synth_packed_device(buf.data_ptr(), M, rl, N, seed=1, missing=0.0, npops=4, delta=0.05)
    c.adopt_packed_device(buf.data_ptr(), M, rl, N); c.set_rows(None)
    r = c.grm(want_snp=False)
    t0 = time.time(); lam, vec = c.eig(10, want_lambda=False); t1 = time.time() - t0
    tm = c.timings()
    print(json.dumps(dict(N=N, M=M, vec_s=t1, iters=tm["chfsi_iters"], matvecs=tm["chfsi_matvecs"], vectors_ms=tm["vectors_ms"])), flush=True)
    if N <= 6000:
        # residual check against the GRM itself
        rr = c.grm(want_xtx=True, want_snp=False)
        X = rr["XTX"]
        th = np.einsum("ij,ij->i", vec @ X, vec)
        res = np.linalg.norm(vec @ X - th[:, None] * vec, axis=1)
        print("theta", th[:6], "res", res.max(), "orth", np.abs(vec @ vec.T - np.eye(10)).max(), flush=True)
    del buf
