"""GPU parity: fastmode (kjg_fpca) and the projection passes against the reference / port."""
import numpy as np
import pytest

from eig_b200 import capi, synth
from oracle import bindings as ob

pytestmark = pytest.mark.gpu


def _ref_fpca(P, nind, **kw):
    if ob.ref() is not None:
        e, v, _ = ob.ref_fpca(P, nind, **kw)
        return e, v
    return ob.port_fpca(P, nind, **kw)


@pytest.mark.parametrize("nsnp,nind,miss,K,L,I,rows,alt", [
    (2000, 120, 0.05, 4, 8, 3, None, 1),
    (3000, 300, 0.0, 2, 5, 2, None, 0),          # odd L: last column of the start matrix is a uniform
    (5000, 400, 0.2, 5, 10, 4, "subset", 1),
    (1500, 260, 0.1, 3, 20, 2, None, 1),
])
def test_fpca_matches_reference(ctx, nsnp, nind, miss, K, L, I, rows, alt):
    g = synth.genotypes(31, nsnp, nind, missing=miss, npops=K + 1, delta=0.3)
    g[3, :] = 1                                    # monomorphic-in-counts SNP stays a row of X
    P = synth.pack(g)
    xi = None
    if rows == "subset":
        xi = np.sort(np.random.RandomState(1).choice(nind, nind - 37, replace=False)).astype(np.int32)
    ctx.upload_packed(P, nind); ctx.set_rows(xi)
    ev, vec = ctx.fpca(K, L, I, seed=99, altnormstyle=alt)
    re_, rv = _ref_fpca(P, nind, K=K, L=L, I=I, seed=99, xindex=xi, altnormstyle=alt)
    assert (np.abs(ev - re_) / re_).max() < 1e-9
    cos = np.abs((vec * rv).sum(0))
    assert np.abs(cos - 1).max() < 1e-9, cos
    assert np.abs(np.linalg.norm(vec, axis=0) - 1).max() < 1e-12


def test_project_matches_port(ctx):
    nsnp, nind = 3000, 200
    g = synth.genotypes(13, nsnp, nind, missing=0.1, npops=3, delta=0.3)
    g[5, :] = 0
    P = synth.pack(g)
    xi = np.arange(3, nind, dtype=np.int32)
    ctx.upload_packed(P, nind); ctx.set_rows(xi)
    r = ctx.grm()
    lam, vec = ctx.eig(3)
    ff, fx, sc = ctx.project(vec)
    pf, px, ps = ob.port_project(P, nind, r["used"], r["xmean"], r["xfancy"], vec, xindex=xi)
    assert np.abs(ff - pf).max() < 1e-10 * np.abs(pf).max()
    assert np.abs(fx - px).max() < 1e-10 * np.abs(px).max()
    assert np.abs(sc - ps).max() < 1e-10 * np.abs(ps).max()
    # fxvecs are (a multiple of) the eigenvectors when nothing is missing; with missing data they stay highly correlated
    for j in range(3):
        assert abs(np.corrcoef(fx[j], vec[j])[0, 1]) > 0.9
