"""GPU parity: per-population genotype-class counts, bit-exact against the port (itself pinned bit for
bit to the reference's fstcolyy, qpsubs.c:1205-1346)."""
import numpy as np
import pytest

from eig_b200 import synth
from oracle import bindings as ob

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nsnp,nind,npops,missing", [(500, 150, 5, 0.1), (2000, 1037, 12, 0.3), (300, 40, 40, 0.0), (100, 70, 1, 0.5)])
def test_pop_counts_bit_exact(ctx, nsnp, nind, npops, missing):
    P = synth.pack(synth.genotypes(7, nsnp, nind, missing=missing, npops=min(npops, 12), delta=0.3))
    rs = np.random.RandomState(1)
    xi = np.sort(rs.choice(nind, size=nind - nind // 7, replace=False)).astype(np.int32)
    xt = rs.randint(0, npops, len(xi)).astype(np.int32)
    xt[::11] = -1; xt[5] = npops + 2                    # skipped labels (fstcolyy: k < 0 or k >= numeg)
    ctx.upload_packed(P, nind); ctx.set_rows(xi)
    got = ctx.pop_counts(xt, npops)
    want = ob.port_pop_counts(P, nind, xt, npops, xindex=xi)
    assert np.array_equal(got, want)
    # consistency with the per-SNP totals when every row has a label
    xt2 = rs.randint(0, npops, len(xi)).astype(np.int32)
    got2 = ctx.pop_counts(xt2, npops)
    c0, c1, nm = ctx.snp_counts()
    assert np.array_equal(got2[:, :, 1].sum(1) + 2 * got2[:, :, 2].sum(1), c0)
    assert np.array_equal(got2.sum((1, 2)), len(xi) - nm)
    if ob.ref() is not None and npops > 1:
        en, ed = ob.port_fstcol(got)
        ren, red = ob.ref_fstcol(P, nind, xt, npops, xindex=xi)
        assert np.array_equal(en, ren) and np.array_equal(ed, red)
