"""GPU probe (run under gpurun): FP64 peaks, GRM kernel rate, eigensolver timings."""
import json
import sys
import time
import numpy as np
import torch

sys.path.insert(0, ".")
from eig_b200 import capi, synth

out = {}
c = capi.Context(0)
out["microbench_tflops"] = dict(zip(("dmma", "dfma"), c.microbench_fp64()))
# cuBLAS DGEMM (library roofline denominator for the FP64 pipe)
n = 8192
a = torch.randn(n, n, dtype=torch.float64, device="cuda"); b = torch.randn(n, n, dtype=torch.float64, device="cuda")
for _ in range(2): torch.matmul(a, b)
torch.cuda.synchronize(); best = 1e9
for _ in range(5):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
out["cublas_dgemm_tflops"] = 2 * n ** 3 / (best * 1e-3) / 1e12
del a, b

for spec in (sys.argv[1:] or ["5000x100000"]):
    noeig = spec.endswith(":noeig"); spec = spec.split(":")[0]
    N, M = (int(v) for v in spec.split("x"))
    rl = synth.rlen_for(N)
    buf = torch.empty((M, rl), dtype=torch.uint8, device="cuda")
    c.synth_packed_device(buf.data_ptr(), M, rl, N, seed=1)
    c.adopt_packed_device(buf.data_ptr(), M, rl, N)
    c.set_rows(None)
    for rep in range(2):
        t0 = time.time(); r = c.grm(want_snp=False); t1 = time.time()
        tm = c.timings()
    rec = dict(N=N, M=M, wall_s=t1 - t0, nused=r["nused"], y=r["y"], **tm)
    rec["grm_tflops"] = (N * (N + 1.0) * r["nused"]) / (tm["grm_ms"] * 1e-3) / 1e12
    rec["snp_indiv2_per_s"] = N * float(N) * r["nused"] / ((tm["grm_ms"] + tm["stats_ms"] + tm["finalize_ms"]) * 1e-3)
    if not noeig:
        t0 = time.time(); lam, vec = c.eig(10); rec["eig_wall_s"] = time.time() - t0
        rec.update({k: v for k, v in c.timings().items() if k.endswith("_ms")})
        rec["lam_top"] = lam[:4].tolist()
    out["%dx%d" % (N, M)] = rec
    print(json.dumps(rec), flush=True)
    del buf
print(json.dumps(out))
