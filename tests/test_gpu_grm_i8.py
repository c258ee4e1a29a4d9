"""GPU parity of the exact integer tensor-core GRM (grm_i8.cu: tcgen05.mma kind::i8, accumulators in TMEM) against the oracle, the
FP64 DMMA kernel and an exact rational-free check: the integer path has no rounding in the sum over SNPs, so it must sit at least as
close to a longdouble product of the reference's own normalised columns as the FP64 kernel does."""
import numpy as np
import pytest

from eig_b200 import synth
from oracle import bindings as ob

pytestmark = pytest.mark.gpu


def _oracle_grm(P, n, **kw):
    return ob.ref_grm(P, n, **kw) if ob.ref() is not None else ob.port_grm(P, n, **kw)


def _both(ctx, **kw):
    ctx.set_option("grm_method", 1)
    d = ctx.grm(want_xtx=True, **kw)
    ctx.set_option("grm_method", 2)
    i = ctx.grm(want_xtx=True, **kw)
    t = ctx.timings()
    # one CTA per 128 x 256 tile instead of CTA pairs on 256 x 256 tiles: the integer sums are exact and the FP64 accumulation order
    # is the same, so the two kernels must agree bit for bit
    ctx.set_option("i8_pair", 0)
    try:
        j = ctx.grm(want_xtx=True, **kw)
    finally:
        ctx.set_option("i8_pair", 1)
    assert np.array_equal(i["XTX"], j["XTX"])
    ctx.set_option("grm_method", 0)
    return d, i, t


CASES = [
    # nsnp, nind, missing, altnorm, fancynorm, rows, slab
    (7, 5, 0.0, 0, 1, None, 0),
    (300, 101, 0.10, 1, 1, None, 0),
    (1000, 333, 0.30, 0, 1, "subset", 256),
    (513, 128, 0.0, 1, 0, None, 0),
    (2049, 700, 0.05, 1, 1, "subset", 512),
    (1500, 1100, 0.0, 1, 1, None, 384),
]


@pytest.mark.parametrize("nsnp,nind,miss,alt,fancy,rows,slab", CASES)
def test_i8_grm_matches_oracle_and_dmma(ctx, nsnp, nind, miss, alt, fancy, rows, slab):
    g = synth.genotypes(11, nsnp, nind, missing=miss, npops=3, delta=0.2)
    g[0, :] = -1            # an all-missing SNP
    g[1, :] = 2             # a monomorphic SNP
    if nsnp > 600 and miss > 0:
        g[256:512, :] = np.where(g[256:512, :] < 0, 1, g[256:512, :])      # two 128-SNP blocks without a missing genotype
    P = synth.pack(g)
    xi = None
    if rows == "subset":
        rs = np.random.RandomState(5)
        xi = np.sort(rs.choice(nind, size=nind - nind // 5, replace=False)).astype(np.int32)
    ctx.upload_packed(P, nind)
    ctx.set_rows(xi)
    ctx.set_option("i8_slab", slab)
    try:
        d, r, t = _both(ctx, fancynorm=fancy, altnormstyle=alt)
    finally:
        ctx.set_option("i8_slab", 0)
    assert t["grm_method"] == 2 and 7 <= t["i8_slices"] <= 9
    assert t["i8_segments"] == (2 if miss > 0 else 1)
    o = _oracle_grm(P, nind, xindex=xi, fancynorm=fancy, altnormstyle=alt)
    for k in ("c0", "c1", "nmiss", "used"):
        assert np.array_equal(r[k], o[k]), k
    assert r["nused"] == int(o["used"].sum())
    assert abs(r["y"] - o["y"]) <= 1e-12 * abs(o["y"])
    ref = o["XTX"] / o["y"]
    scale = np.abs(ref).max()
    assert np.abs(r["XTX"] - ref).max() <= 1e-11 * scale
    assert np.abs(r["XTX"] - d["XTX"]).max() <= 1e-12 * scale
    assert np.array_equal(r["XTX"], r["XTX"].T)


def test_i8_grm_weights_ignore_and_repeatable(ctx):
    nsnp, nind = 900, 260
    g = synth.genotypes(9, nsnp, nind, missing=0.1)
    P = synth.pack(g)
    rs = np.random.RandomState(1)
    w = 0.5 + rs.rand(nsnp)
    ign = (rs.rand(nsnp) < 0.1).astype(np.uint8)
    ctx.upload_packed(P, nind); ctx.set_rows(None)
    d, r, t = _both(ctx, snp_weight=w, snp_ignore=ign, maxmissing=40)
    assert np.array_equal(r["used"], d["used"]) and not r["used"][ign.astype(bool)].any()
    assert np.abs(r["XTX"] - d["XTX"]).max() <= 1e-12 * np.abs(d["XTX"]).max()
    d1, r1, _ = _both(ctx, snp_weight=w, maxmissing=40)
    o = _oracle_grm(P, nind, weights=w, maxmissing=40)
    assert np.array_equal(r1["used"], o["used"])
    assert np.abs(r1["XTX"] - o["XTX"] / o["y"]).max() < 1e-11 * np.abs(o["XTX"] / o["y"]).max()
    ctx.set_option("grm_method", 2)
    r2 = ctx.grm(snp_weight=w, snp_ignore=ign, maxmissing=40, want_xtx=True)
    ctx.set_option("grm_method", 0)
    assert np.array_equal(r["XTX"], r2["XTX"])              # integer sums + fixed-order FP64: bit-identical from run to run


def test_i8_grm_closer_to_exact_than_fp64(ctx):
    """The sum over SNPs is exact on the integer path; only the 56-bit weights round.  Against a longdouble product of the same
    table values it must be at least as accurate as the DMMA kernel (whose FP64 sums round M times)."""
    nsnp, nind = 4096, 300
    g = synth.genotypes(21, nsnp, nind, missing=0.08, npops=4, delta=0.3)
    P = synth.pack(g)
    ctx.upload_packed(P, nind); ctx.set_rows(None)
    d, r, t = _both(ctx)
    used = r["used"].astype(bool)
    gg = g[used].astype(np.longdouble)
    valid = g[used] >= 0
    xm = r["xmean"][used].astype(np.longdouble); xf = r["xfancy"][used].astype(np.longdouble)
    X = np.where(valid, gg * xf[:, None] - xm[:, None], 0)        # (g - mean) * scale, mean stored as mean * scale
    exact = (X.T @ X)
    y = np.trace(exact) / (nind - 1)
    exact = (exact / y).astype(np.float64)
    e_i8 = np.abs(r["XTX"] - exact).max(); e_dm = np.abs(d["XTX"] - exact).max()
    assert e_i8 <= 2e-13 * np.abs(exact).max()
    assert e_i8 <= 4 * e_dm + 1e-15 * np.abs(exact).max()


def test_i8_grm_many_tiles_and_pca(ctx):
    """Several tiles per CTA, an odd number of 128-row tiles, outliers: eb_pca_full on the integer path == on the FP64 path."""
    nsnp, nind = 3000, 2700
    g = synth.genotypes(5, nsnp, nind, missing=0.02, npops=3, delta=0.25)
    P = synth.pack(g)
    ctx.upload_packed(P, nind); ctx.set_rows(None)
    d, r, t = _both(ctx)
    assert np.abs(r["XTX"] - d["XTX"]).max() <= 1e-12 * np.abs(d["XTX"]).max()
    ctx.set_option("grm_method", 1)
    a = ctx.pca_full(numeigs=4, numoutliter=2)
    ctx.set_option("grm_method", 2)
    b = ctx.pca_full(numeigs=4, numoutliter=2)
    ctx.set_option("grm_method", 0)
    assert np.array_equal(a["xindex"], b["xindex"])
    la, lb = a["lambda_"][:4], b["lambda_"][:4]
    assert np.abs(la - lb).max() <= 1e-10 * la[0]
    for i in range(3):
        assert abs(abs(a["evecs"][i] @ b["evecs"][i]) - 1) < 1e-9
