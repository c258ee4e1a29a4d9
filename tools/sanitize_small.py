"""Tiny pass over the newer kernels for `compute-sanitizer --tool memcheck` (sizes chosen to finish in about a minute under the tool)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from eig_b200 import capi, synth
c = capi.Context(0)
nsnp, nind = 700, 150
P = synth.packed_genotypes(5, nsnp, nind, missing=0.1, npops=3, delta=0.3)
c.upload_packed(P, nind); c.set_rows(np.arange(3, nind - 2, dtype=np.int32))
r = c.grm(want_xtx=True)
lam, vec = c.eig(4)
co, es, ok = c.evec_coords(vec, indiv=np.arange(nind, dtype=np.int32))
a, l2, ok2 = c.shrink_coords(3, newshrink=False)
b, l3, ok3 = c.shrink_coords(3, newshrink=True)
rs = np.random.RandomState(1); X = rs.randn(130, 117); A = X @ X.T / 117
c.set_option("eig_method", 1); w, V = c.eigvecs(A); c.set_option("eig_method", 0)
G = c.debug_gemm(rs.randn(150, 70), rs.randn(90, 70))
ev, u = c.fpca(3, 6, 2, seed=3)
pc = c.pop_counts(np.asarray(synth.pop_of(nind - 5, 3), np.int32), 3)
print("ok", float(np.abs(V @ V.T - np.eye(130)).max()), ok.all(), ok2.all(), ok3.all())
