"""GPU parity for usepopsformissing (getcolxz + fvadjust + domult_increment_normal, smartpca.c:3129-3216, 2236-2279,
3531-3561): eb_grm_popfill fills missing genotypes with population means, normalises, applies the drop rule and accumulates the GRM
on the device; checked against the UNMODIFIED reference's getcolxz (usepopsformissing = YES) and its dense accumulation."""
import numpy as np
import pytest

from eig_b200 import capi, synth
from oracle import bindings as ob

pytestmark = pytest.mark.gpu


def _case(seed, nsnp, nind, npops, missing):
    g = synth.genotypes(seed, nsnp, nind, missing=missing, npops=npops, delta=0.3)
    g[5, :] = -1                                        # a SNP without any data: dropped (-1)
    g[6, synth.pop_of(nind, npops) == 1] = -1           # a population without data at a SNP: its members stay missing
    g[7, :] = 1; g[7, 3] = -1                           # monomorphic apart from one missing genotype
    return g


@pytest.mark.parametrize("nsnp,nind,npops,missing,alt,rows", [(2500, 300, 4, 0.2, 1, None), (4000, 1100, 12, 0.3, 0, "subset"), (1500, 130, 1, 0.1, 1, None)])
def test_popfill_grm_vs_reference(ctx, nsnp, nind, npops, missing, alt, rows):
    if ob.ref() is None:
        pytest.skip("compiled reference not available")
    g = _case(11, nsnp, nind, npops, missing)
    P = synth.pack(g)
    xi = np.arange(nind, dtype=np.int32) if rows is None else np.array([i for i in range(nind) if i % 7 != 3], dtype=np.int32)
    xt = synth.pop_of(nind, npops)[xi].astype(np.int32)
    ctx.upload_packed(P, nind); ctx.set_rows(xi)
    r = ctx.grm_popfill(xt, npops, altnormstyle=alt, want_xtx=True)
    ref = ob.ref_popfill_cols(P, nind, xt, npops, xindex=xi, altnormstyle=alt)
    # integer outputs and the drop rule (smartpca.c:1131-1144): bit-exact
    assert np.array_equal(r["c0"], ref["c0"]) and np.array_equal(r["c1"], ref["c1"]) and np.array_equal(r["nmiss"], ref["nmiss"])
    t = np.minimum(ref["c0"], ref["c1"])
    used = ~((t < 1) | (ref["nmiss"] < 0) | (t == 0))
    assert np.array_equal(r["used"].astype(bool), used) and r["nused"] == used.sum()
    assert not used[5]
    if npops > 1:
        assert ref["nmiss"][6] > 0                      # the population without data keeps its missing genotypes
    # normalisation: to rounding (the filled sum is accumulated in a different order)
    assert np.abs(r["xmean"] - ref["xmean"]).max() <= 1e-12 * np.abs(ref["xmean"]).max()
    assert np.abs(r["xfancy"] - ref["xfancy"]).max() <= 1e-12 * np.abs(ref["xfancy"]).max()
    # GRM: the reference's own dense accumulation of its own columns
    want = ob.ref_dense_grm(ref["cols"][used])
    y = np.trace(want) / (len(xi) - 1)
    assert abs(r["y"] - y) <= 1e-11 * y
    assert np.abs(r["XTX"] - want / y).max() <= 1e-11 * np.abs(want / y).max()
    # the eigensolver runs on the resident matrix; the packed-table projection passes refuse
    lam, vec = ctx.eig(3)
    w = np.linalg.eigvalsh(want / y)[::-1]
    assert np.abs(lam - w).max() <= 1e-9 * w[0]
    with pytest.raises(capi.EigB200Error, match="usepopsformissing"):
        ctx.project(vec)
    ctx.grm()                                           # a plain pass clears the state
    ctx.project(ctx.eig(2)[1])


def test_popfill_without_missing_equals_plain_pass(ctx):
    """no missing genotype: nothing to fill, so the dense path must reproduce the packed path (same counts, same table, same GRM)"""
    nsnp, nind = 3000, 260
    P = synth.packed_genotypes(4, nsnp, nind, npops=3, delta=0.2)
    ctx.upload_packed(P, nind); ctx.set_rows(None)
    a = ctx.grm(want_xtx=True)
    b = ctx.grm_popfill(synth.pop_of(nind, 3).astype(np.int32), 3, want_xtx=True)
    for k in ("c0", "c1", "nmiss", "used"):
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(a["xmean"], b["xmean"]) and np.array_equal(a["xfancy"], b["xfancy"])
    assert np.abs(a["XTX"] - b["XTX"]).max() <= 1e-12 * np.abs(a["XTX"]).max()
