"""GPU parity: integer reductions, normalisation and the GRM kernel against the oracle (port and compiled reference)."""
import numpy as np
import pytest

from eig_b200 import synth
from oracle import bindings as ob

pytestmark = pytest.mark.gpu


def _oracle_grm(P, n, **kw):
    return ob.ref_grm(P, n, **kw) if ob.ref() is not None else ob.port_grm(P, n, **kw)


CASES = [
    # nsnp, nind, missing, altnorm, fancynorm, rows
    (7, 5, 0.0, 0, 1, None),
    (300, 101, 0.10, 1, 1, None),
    (1000, 333, 0.30, 0, 1, "subset"),
    (513, 128, 0.0, 1, 0, None),
    (2049, 700, 0.05, 1, 1, "subset"),
]


@pytest.mark.parametrize("nsnp,nind,miss,alt,fancy,rows", CASES)
def test_grm_matches_oracle(ctx, nsnp, nind, miss, alt, fancy, rows):
    g = synth.genotypes(11, nsnp, nind, missing=miss, npops=3, delta=0.2)
    g[0, :] = -1            # an all-missing SNP
    g[1, :] = 2             # a monomorphic SNP
    P = synth.pack(g)
    xi = None
    if rows == "subset":
        rs = np.random.RandomState(5)
        xi = np.sort(rs.choice(nind, size=nind - nind // 5, replace=False)).astype(np.int32)
    ctx.upload_packed(P, nind)
    ctx.set_rows(xi)
    r = ctx.grm(fancynorm=fancy, altnormstyle=alt, want_xtx=True)
    o = _oracle_grm(P, nind, xindex=xi, fancynorm=fancy, altnormstyle=alt)
    for k in ("c0", "c1", "nmiss", "used"):
        assert np.array_equal(r[k], o[k]), k                       # integer work: bit-exact
    assert np.array_equal(r["xmean"], o["xmean"]) and np.array_equal(r["xfancy"], o["xfancy"])   # IEEE div/sqrt: bit-exact
    assert r["nused"] == int(o["used"].sum())
    assert abs(r["y"] - o["y"]) <= 1e-12 * abs(o["y"])
    ref = o["XTX"] / o["y"]
    assert np.abs(r["XTX"] - ref).max() <= 1e-11 * np.abs(ref).max()   # FP64, summation order differs
    assert np.array_equal(r["XTX"], r["XTX"].T)


def test_snp_and_indiv_counts(ctx):
    nsnp, nind = 900, 257
    g = synth.genotypes(3, nsnp, nind, missing=0.2)
    P = synth.pack(g)
    ctx.upload_packed(P, nind)
    xi = np.arange(0, nind, 2, dtype=np.int32)
    ctx.set_rows(xi)
    c0, c1, nm = ctx.snp_counts()
    p0, p1, pm = ob.port_snp_counts(P, xi, nind)
    assert np.array_equal(c0, p0) and np.array_equal(c1, p1) and np.array_equal(nm, pm)
    keep = (np.arange(nsnp) % 3 != 0).astype(np.uint8)
    assert np.array_equal(ctx.indiv_valid_counts(keep), ob.port_indiv_valid_counts(P, nind, keep))
    assert np.array_equal(ctx.indiv_valid_counts(None), (g >= 0).sum(0))


def test_weights_and_ignore(ctx):
    nsnp, nind = 400, 96
    g = synth.genotypes(9, nsnp, nind, missing=0.1)
    P = synth.pack(g)
    w = 0.5 + np.random.RandomState(1).rand(nsnp)
    ctx.upload_packed(P, nind); ctx.set_rows(None)
    r = ctx.grm(snp_weight=w, want_xtx=True, maxmissing=12)
    o = _oracle_grm(P, nind, weights=w, maxmissing=12)
    assert np.array_equal(r["used"], o["used"])
    assert np.abs(r["XTX"] - o["XTX"] / o["y"]).max() < 1e-11 * np.abs(o["XTX"] / o["y"]).max()


def test_device_synth_matches_host(ctx):
    import torch
    nsnp, nind = 257, 1001
    rl = synth.rlen_for(nind)
    buf = torch.empty((nsnp, rl), dtype=torch.uint8, device="cuda:0")
    ctx.synth_packed_device(buf.data_ptr(), nsnp, rl, nind, seed=5, s0=10, missing=0.07, npops=4, delta=0.1)
    ctx.sync()
    host = synth.packed_genotypes(5, nsnp, nind, s0=10, missing=0.07, npops=4, delta=0.1)
    assert np.array_equal(buf.cpu().numpy(), host)


def test_grm_large_property(ctx):
    """size-independent checks at a larger shape: trace identity and row sums (centred columns => X^T 1 = 0 without missing)."""
    import torch
    nsnp, nind = 20000, 2000
    rl = synth.rlen_for(nind)
    buf = torch.empty((nsnp, rl), dtype=torch.uint8, device="cuda:0")
    ctx.synth_packed_device(buf.data_ptr(), nsnp, rl, nind, seed=2)
    ctx.adopt_packed_device(buf.data_ptr(), nsnp, rl, nind)
    ctx.set_rows(None)
    r = ctx.grm(want_xtx=True)
    X = r["XTX"]
    assert abs(np.trace(X) - (nind - 1)) < 1e-9 * nind
    assert np.abs(X.sum(1)).max() < 1e-8 * np.abs(X).max() * nind
    assert r["nused"] > 0.99 * nsnp


def test_grm_many_items_per_cta_exact_and_repeatable(ctx):
    """3000 x 60000 with missing data and structure: ~20 (tile, SNP-chunk) items per persistent CTA.  The result must be
    bit-identical from run to run (fixed-order split-K) and agree with an independent FP64 product of the decoded matrix."""
    import torch
    nsnp, nind = 60000, 3000
    rl = synth.rlen_for(nind)
    buf = torch.empty((nsnp, rl), dtype=torch.uint8, device="cuda")
    ctx.synth_packed_device(buf.data_ptr(), nsnp, rl, nind, seed=3, missing=0.1, npops=3, delta=0.1)
    ctx.adopt_packed_device(buf.data_ptr(), nsnp, rl, nind); ctx.set_rows(None)
    r = ctx.grm(want_xtx=True)
    a = r["XTX"] * r["y"]
    b = ctx.grm(want_xtx=True, want_snp=False)
    assert np.array_equal(a, b["XTX"] * b["y"])
    P = buf.cpu().numpy()
    ref = torch.zeros((nind, nind), dtype=torch.float64, device="cuda")
    xm = torch.from_numpy(r["xmean"]).cuda(); xf = torch.from_numpy(r["xfancy"]).cuda(); us = torch.from_numpy(r["used"].astype(np.float64)).cuda()
    for s0 in range(0, nsnp, 20000):
        g = torch.from_numpy(synth.unpack(P[s0:s0 + 20000], nind).astype(np.float64)).cuda()
        x = torch.where(g < 0, torch.zeros_like(g), g * xf[s0:s0 + 20000, None] - xm[s0:s0 + 20000, None]) * us[s0:s0 + 20000, None]
        ref += x.T @ x
    ref = ref.cpu().numpy()
    assert np.abs(a - ref).max() <= 1e-12 * np.abs(ref).max()
