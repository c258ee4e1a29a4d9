"""GPU parity for the .evec coordinate path: eb_lsqproj / eb_evec_coords against the oracle port and,
when built, the unmodified reference's own sequence smartpca.c:1440-1564 (lsqproj smartpca.c:4606-4757)."""
import numpy as np
import pytest

from eig_b200 import synth
from oracle import bindings as ob

pytestmark = pytest.mark.gpu

EVEC_ATOL = 1e-6      # north_star: .evec entries within 1e-6 absolute


def _case(seed, nsnp, nind, missing):
    g = synth.genotypes(seed, nsnp, nind, missing=missing, npops=3, delta=0.3)
    g[:, 7] = -1                      # an individual without any data
    g[5:, 11] = -1                    # an individual with 5 valid SNPs (<= numeigs): "insufficient data"
    return synth.pack(g)


@pytest.mark.parametrize("nsnp,nind,missing,k", [(3000, 120, 0.1, 6), (5000, 333, 0.3, 10), (2000, 64, 0.0, 3)])
def test_evec_coords_vs_oracle(ctx, nsnp, nind, missing, k):
    P = _case(3, nsnp, nind, missing)
    # PCA rows: a subset (the rest are projected, like populations outside poplistname)
    xi = np.array([i for i in range(nind) if i % 5 != 0 and i not in (7, 11)], dtype=np.int32)
    ctx.upload_packed(P, nind); ctx.set_rows(xi)
    r = ctx.grm()
    lam, vec = ctx.eig(k)
    co, es, ok = ctx.evec_coords(vec)
    pc, pes, pok, ff, sc = ob.port_evec_coords(P, nind, r["used"], r["xmean"], r["xfancy"], vec, xindex=xi)
    assert np.array_equal(ok, pok)
    assert set(np.flatnonzero(ok == 0)) == ({7, 11} if k >= 5 else {7})    # individual 11 has 5 valid SNPs
    assert np.abs(co - pc).max() <= 1e-9 * max(1.0, np.abs(pc).max())
    assert np.abs(es - pes).max() <= 1e-9 * np.abs(pes).max()
    assert np.abs(co[:, ok == 0]).max() == 0.0
    if ob.ref() is not None:
        rr = ob.ref_evec_coords(P, nind, r["used"], r["xmean"], r["xfancy"], vec, xindex=xi)
        assert np.array_equal(rr["ignored"], 1 - ok)
        assert np.abs(co - rr["coords"]).max() <= EVEC_ATOL * 1e-3
        assert np.abs(es - rr["eigscale"]).max() <= 1e-9 * np.abs(rr["eigscale"]).max()


def test_lsqproj_pieces(ctx):
    nsnp, nind, k = 2500, 200, 5
    P = _case(9, nsnp, nind, 0.2)
    ctx.upload_packed(P, nind); ctx.set_rows(None)
    r = ctx.grm()
    lam, vec = ctx.eig(k)
    ff, fx, sc = ctx.project(vec)
    lst = np.arange(0, nind, 3, dtype=np.int32)
    a, b, nv, ok = ctx.lsqproj(ff, sc, indiv=lst)
    pa, pb, pnv, pok = ob.port_lsqproj(P, lst, r["used"], r["xmean"], r["xfancy"], ff, sc)
    assert np.array_equal(nv, pnv) and np.array_equal(ok, pok)       # integer outputs: bit-exact
    assert np.abs(a - pa).max() <= 1e-10 * np.abs(pa).max()
    assert np.abs(b - pb).max() <= 1e-10 * np.abs(pb).max()


def test_evec_coords_requires_pca_rows_in_list(ctx):
    from eig_b200.capi import EigB200Error
    P = _case(1, 1000, 40, 0.0)
    ctx.upload_packed(P, 40); ctx.set_rows(None)
    ctx.grm(); lam, vec = ctx.eig(2)
    with pytest.raises(EigB200Error):
        ctx.evec_coords(vec, indiv=np.arange(1, 40, dtype=np.int32))


def test_evec_coords_with_snp_weights(ctx):
    """weightname (smartpca.c:1178-1180) scales the GRM columns only: the SNP loadings come from getcolxf (smartpca.c:1487,
    3564-3597), which never applies cupt->weight, so the .evec coordinates of a weighted run must match the reference's
    sequence fed with the same eigenvectors."""
    nsnp, nind, k = 3000, 150, 4
    P = _case(5, nsnp, nind, 0.1)
    w = 0.25 + 1.5 * np.random.default_rng(2).random(nsnp)
    ctx.upload_packed(P, nind); ctx.set_rows(None)
    r = ctx.grm(snp_weight=w)
    r0 = ctx.grm()
    assert np.abs(r["y"] - r0["y"]) > 1e-3 * r0["y"]          # the weights did change the GRM
    r = ctx.grm(snp_weight=w)
    lam, vec = ctx.eig(k)
    co, es, ok = ctx.evec_coords(vec)
    pc, pes, pok, ff, sc = ob.port_evec_coords(P, nind, r["used"], r["xmean"], r["xfancy"], vec)
    assert np.array_equal(ok, pok)
    assert np.abs(co - pc).max() <= 1e-9 * max(1.0, np.abs(pc).max())
    if ob.ref() is not None:
        rr = ob.ref_evec_coords(P, nind, r["used"], r["xmean"], r["xfancy"], vec)
        assert np.abs(co - rr["coords"]).max() <= EVEC_ATOL * 1e-3
        assert np.abs(es - rr["eigscale"]).max() <= 1e-9 * np.abs(rr["eigscale"]).max()
