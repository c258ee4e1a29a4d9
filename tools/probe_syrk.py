"""GPU probe: the dense-block SYRK (same syr2k_lower_kernel as the band reduction) at large n, repeated."""
import sys
import numpy as np
sys.path.insert(0, ".")
from eig_b200 import capi
c = capi.Context(0)
for n in [int(v) for v in sys.argv[1:]]:
    rs = np.random.RandomState(n)
    T = rs.randn(192, n); T -= T.mean(axis=1, keepdims=True)
    want = T.T @ T
    for rep in range(4):
        y, X = c.grm_dense([T[:64], T[64:]], n, want_xtx=True)
        got = X * y
        err = np.abs(got - want)
        bad = np.argwhere(err > 1e-9 * np.abs(want).max())
        print("n", n, "rep", rep, "max err", err.max(), "bad elems", len(bad), "rows", sorted(set((bad[:, 0] // 128).tolist()))[:8], "cols64", sorted(set((bad[:, 1] // 64).tolist()))[:8], flush=True)
