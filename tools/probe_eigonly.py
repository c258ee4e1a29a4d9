"""GPU probe for profilers: one GRM + one full eigen solve (two-stage + subspace iteration) + one fastmode run.
usage: probe_eigonly.py N M [fast]"""
import sys, time, json
import torch
sys.path.insert(0, ".")
from eig_b200 import capi, synth
N, M = int(sys.argv[1]), int(sys.argv[2])
c = capi.Context(0)
rl = synth.rlen_for(N)
buf = torch.empty((M, rl), dtype=torch.uint8, device="cuda")
c.synth_packed_device(buf.data_ptr(), M, rl, N, seed=1, missing=0.05, npops=4, delta=0.05)
c.adopt_packed_device(buf.data_ptr(), M, rl, N); c.set_rows(None)
if len(sys.argv) > 3 and sys.argv[3] == "fast":
    t0 = time.time(); ev, vec = c.fpca(10, 20, 10, seed=5); print(json.dumps(dict(fpca_s=time.time() - t0, ev=ev[:3].tolist())))
else:
    r = c.grm(want_snp=False)
    t0 = time.time(); lam, vec = c.eig(10); t = time.time() - t0
    print(json.dumps(dict(N=N, M=M, eig_s=t, **{k: v for k, v in c.timings().items() if k.endswith("_ms")}, lam=lam[:4].tolist())))
