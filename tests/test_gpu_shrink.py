"""GPU parity for shrinkmode: eb_shrink_coords against the unmodified reference's doshrinkp / doshrinkp2
(smartpca.c:4223-4419 / 4022-4220, through oracle/_ref when built) and the numpy restatement in oracle/bindings.py; plus
the general FP64 tensor-core GEMM the path is built on."""
import numpy as np
import pytest

from eig_b200 import synth
from oracle import bindings as ob

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,N,K", [(128, 64, 32), (300, 130, 77), (1000, 517, 259), (64, 2048, 1500)])
@pytest.mark.parametrize("a_km,b_kn", [(False, False), (True, True), (True, False), (False, True)])
def test_general_gemm(ctx, M, N, K, a_km, b_kn):
    rs = np.random.RandomState(M + N + K)
    A = rs.randn(M, K); B = rs.randn(N, K)
    got = ctx.debug_gemm(A.T.copy() if a_km else A, B.T.copy() if b_kn else B, a_km=a_km, b_kn=b_kn)
    want = A @ B.T
    assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max() * np.sqrt(K)


def _case(nsnp, nind, missing, lo, hi, seed=3):
    P = synth.packed_genotypes(seed, nsnp, nind, missing=missing, npops=4, delta=0.35)
    P[17] = 0xFF                                  # all-missing SNP: dropped by the rule, zero column everywhere
    xi = np.arange(lo, nind - hi, dtype=np.int32)
    return P, xi


@pytest.mark.parametrize("newshrink", [False, True])
@pytest.mark.parametrize("nsnp,nind,missing,lo,hi,k", [(2500, 120, 0.15, 0, 0, 4), (3000, 170, 0.3, 6, 9, 6), (1200, 140, 0.0, 3, 0, 3),
                                                     (2000, 420, 0.1, 5, 5, 3)])
def test_shrink_coords_vs_reference(ctx, nsnp, nind, missing, lo, hi, k, newshrink):
    P, xi = _case(nsnp, nind, missing, lo, hi)
    ctx.upload_packed(P, nind); ctx.set_rows(xi)
    r = ctx.grm(want_xtx=True)
    got, lam, ok = ctx.shrink_coords(k, newshrink=newshrink)
    assert ok.all()
    o = (ob.ref_grm if ob.ref() is not None else ob.port_grm)(P, nind, xindex=xi)
    assert np.array_equal(o["used"], r["used"])
    X = o["XTX"] / o["y"]
    want, wlam = ob.port_shrink(P, nind, o["used"], o["xmean"], o["xfancy"], X, k, xindex=xi, newshrink=newshrink)
    if ob.ref() is not None:                      # the port itself against the unmodified reference
        rw = ob.ref_shrink(P, nind, o["used"], o["xmean"], o["xfancy"], X, k, xindex=xi, newshrink=newshrink)
        sg = np.sign((rw * want).sum(1))
        assert np.abs(rw - want * sg[:, None]).max() < 1e-10
        want = rw
    assert np.abs(lam - wlam).max() <= 1e-9 * wlam[0]
    sg = np.sign((got * want).sum(1))              # eigenvector signs are LAPACK's choice
    assert np.abs(got * sg[:, None] - want).max() < 1e-6, np.abs(got * sg[:, None] - want).max()     # north_star: .evec entries 1e-6 absolute
    assert np.abs(np.sqrt((got * got).sum(1)) - 1).max() < 1e-12


def test_shrink_coords_vs_committed_reference_vectors(ctx):
    """against the unmodified reference's outputs committed in tests/golden/ref_shrink.npz"""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_shrink.npz"))
    nsnp, nind, k = int(g["nsnp"]), int(g["nind"]), int(g["k"])
    P = synth.packed_genotypes(int(g["seed"]), nsnp, nind, missing=float(g["missing"]), npops=4, delta=0.35)
    ctx.upload_packed(P, nind); ctx.set_rows(g["xindex"])
    ctx.grm(want_snp=False)
    for new, key in ((False, "shrink_old"), (True, "shrink_new")):
        got, lam, ok = ctx.shrink_coords(k, newshrink=new)
        want = g[key]
        sg = np.sign((got * want).sum(1))
        assert ok.all() and np.abs(got * sg[:, None] - want).max() < 1e-6
