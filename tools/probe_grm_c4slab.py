"""One bench-shaped GRM pass (50,000 individuals x 60,000 SNPs: one slab of BASELINE configs[3]) on one GPU, twice: the launch that
`ncu --set full -k regex:grm_syrk -s 1 -c 1` captures for profiles/ (DRAM traffic, DMMA pipe activity of the dominant kernel)."""
import sys
import time

import torch

sys.path.insert(0, ".")
from eig_b200 import capi, synth  # noqa: E402

nind = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
nsnp = int(sys.argv[2]) if len(sys.argv) > 2 else 60000
ctx = capi.Context(0)
rl = synth.rlen_for(nind)
slab = torch.empty((nsnp, rl), dtype=torch.uint8, device="cuda")
ctx.synth_packed_device(slab.data_ptr(), nsnp, rl, nind, seed=1, s0=0)
ctx.adopt_packed_device(slab.data_ptr(), nsnp, rl, nind)
ctx.set_rows(None)
for rep in range(2):
    t0 = time.perf_counter()
    r = ctx.grm(want_snp=False)
    tm = ctx.timings()
    print("pass %d: %.1f ms wall, kernel %.1f ms, %.2f TFLOP/s algorithmic, sm clock %.0f MHz on %d SMs" % (
        rep, (time.perf_counter() - t0) * 1e3, tm["grm_ms"], nind * (nind + 1.0) * r["nused"] / tm["grm_ms"] / 1e9, tm["grm_sm_mhz"], tm["grm_sms"]), flush=True)
