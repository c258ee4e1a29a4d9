"""fastmode (kjg_fpca, K = 10, L = 20, I = 10) on one GPU with the FP64 DMMA products and with the integer tensor-core products.
usage: probe_pg_i8.py N M"""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from eig_b200 import capi, synth  # noqa: E402

N, M = int(sys.argv[1]), int(sys.argv[2])
c = capi.Context(0)
rl = synth.rlen_for(N)
buf = torch.empty((M, rl), dtype=torch.uint8, device="cuda")
c.synth_packed_device(buf.data_ptr(), M, rl, N, seed=1, missing=0.02, npops=12, delta=0.3)
c.adopt_packed_device(buf.data_ptr(), M, rl, N); c.set_rows(None)
res = {}
runs = (("i8", 2), ("i8", 2)) if len(sys.argv) > 3 and sys.argv[3] == "i8" else (("dmma", 1), ("i8", 2), ("i8", 2))
for name, m in runs:
    c.set_option("pg_method", m)
    t0 = time.time(); ev, vec = c.fpca(10, 20, 10, seed=5); t = time.time() - t0
    res[name] = (ev, vec)
    print(json.dumps(dict(method=name, N=N, M=M, fpca_s=t, ev=ev[:3].tolist())), flush=True)
if "dmma" not in res:
    sys.exit(0)
e1, v1 = res["dmma"]; e2, v2 = res["i8"]
print("max rel eigenvalue diff %.3e; 1 - |cos|: %s" % ((np.abs(e1 - e2) / e1).max(), ["%.1e" % abs(abs(float(v1[:, j] @ v2[:, j])) - 1) for j in range(10)]))
