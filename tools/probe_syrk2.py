import sys
import numpy as np
sys.path.insert(0, ".")
from eig_b200 import capi
c = capi.Context(0)
n = int(sys.argv[1]); grid = 296
rs = np.random.RandomState(n)
T = rs.randn(64, n); T -= T.mean(axis=1, keepdims=True)
want = T.T @ T
nt = (n + 127) // 128
def item_of(I, J):
    return I * (I + 1) + J
for rep in range(3):
    y, X = c.grm_dense([T], n, want_xtx=True)
    got = X * y
    err = np.abs(got - want) > 1e-9 * np.abs(want).max()
    out = []
    for I in range(nt):
        for J in range(2 * I + 2):
            blk = err[I * 128:(I + 1) * 128, J * 64:(J + 1) * 64]
            if blk.any():
                it = item_of(I, J)
                rows = np.flatnonzero(blk.any(axis=1)); cols = np.flatnonzero(blk.any(axis=0))
                # ratio got/want to see whether the contribution is missing (0) or doubled (2)
                g = got[I * 128:(I + 1) * 128, J * 64:(J + 1) * 64][blk]; w = want[I * 128:(I + 1) * 128, J * 64:(J + 1) * 64][blk]
                out.append((it, it % grid, it // grid, I, J, int(blk.sum()), rows.min(), rows.max(), cols.min(), cols.max(), float(np.median(g / w))))
    print("rep", rep, "bad tiles", len(out))
    for o in out[:14]:
        print("  item %d cta %d seq %d (I=%d,J=%d) nbad %d rows %d-%d cols %d-%d median got/want %.3f" % o)
    seqs = [o[2] for o in out]
    print("  seq histogram", np.bincount(seqs) if seqs else [])
