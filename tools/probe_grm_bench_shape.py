"""GPU probe for ncu: exactly one GRM pass at the bench shape (5000 x 600000, no missing)."""
import sys
import torch
sys.path.insert(0, ".")
from eig_b200 import capi, synth
N, M = 5000, 600000
c = capi.Context(0)
rl = synth.rlen_for(N)
buf = torch.empty((M, rl), dtype=torch.uint8, device="cuda")
c.synth_packed_device(buf.data_ptr(), M, rl, N, seed=1)
c.adopt_packed_device(buf.data_ptr(), M, rl, N); c.set_rows(None)
r = c.grm(want_snp=False)
print("nused", r["nused"], c.timings())
