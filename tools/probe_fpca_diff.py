import sys
import numpy as np
sys.path.insert(0, ".")
from eig_b200 import capi, synth
from oracle import bindings as ob
ctx = capi.Context(0)
nsnp, nind = 20000, 2000
P = synth.packed_genotypes(3, nsnp, nind, missing=0.02, npops=4, delta=0.1)
ctx.upload_packed(P, nind); ctx.set_rows(None)
for I in (2, 4, 10):
    er, ur, _ = ob.ref_fpca(P, nind, K=10, L=20, I=I, seed=7)
    eg, ug = ctx.fpca(10, 20, I, seed=7)
    cos = [abs(float(ur[:, k] @ ug[:, k])) for k in range(10)]
    print("I=%d rel eval diff per k:" % I, np.array2string(np.abs(er - eg) / er, precision=1), "1-|cos|:", np.array2string(1 - np.array(cos), precision=1), flush=True)
# exact eigenvalues for comparison
r = ctx.grm(want_snp=False); lam, _ = ctx.eig(10)
print("exact top-10 (full mode, different normalisation of missing data):", np.array2string(lam[:10], precision=4))
print("fastmode I=10 GPU:", np.array2string(eg, precision=4), " ref:", np.array2string(er, precision=4))
