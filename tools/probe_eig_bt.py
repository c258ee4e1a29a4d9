"""Eigen phase at n rows on one GPU: spectrum + 10 vectors, back-transformation (default) against the subspace iteration.
usage: probe_eig_bt.py N M [structured]"""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from eig_b200 import capi, synth  # noqa: E402

N, M = int(sys.argv[1]), int(sys.argv[2])
structured = len(sys.argv) > 3 and sys.argv[3] == "1"
c = capi.Context(0)
rl = synth.rlen_for(N)
buf = torch.empty((M, rl), dtype=torch.uint8, device="cuda")
if structured:
    c.synth_packed_device(buf.data_ptr(), M, rl, N, seed=1, missing=0.05, npops=4, delta=0.05)
else:
    c.synth_packed_device(buf.data_ptr(), M, rl, N, seed=1)
c.adopt_packed_device(buf.data_ptr(), M, rl, N); c.set_rows(None)
c.grm(want_snp=False)
out = {}
for name, opt in (("backtransform", 0), ("subspace", 1), ("backtransform", 0)):
    c.set_option("eig_vectors", opt)
    t0 = time.time(); lam, vec = c.eig(10); t = time.time() - t0
    tm = c.timings()
    out[name] = (lam, vec)
    print(json.dumps(dict(method=name, N=N, eig_s=t, **{k: tm[k] for k in ("tridiag_ms", "band_ms", "chase_ms", "bisect_ms", "vectors_ms", "chfsi_matvecs")},
                          lam=lam[:3].tolist())), flush=True)
la, va = out["backtransform"]; lb, vb = out["subspace"]
print("max |lam diff| / lam0 = %.3e" % (np.abs(la - lb).max() / la[0]))
print("1 - |cos| per vector:", ["%.2e" % abs(abs(float(va[i] @ vb[i])) - 1) for i in range(10)])
print("orthonormality (backtransform): %.2e" % np.abs(va @ va.T - np.eye(10)).max())
