"""GPU parity: eigensolver, outlier loop -- against LAPACK (numpy), the port and the compiled reference."""
import numpy as np
import pytest

from eig_b200 import synth
from oracle import bindings as ob

pytestmark = pytest.mark.gpu

EVAL_RTOL = 1e-9      # north_star: eigenvalues within 1e-9 relative
COS_TOL = 1e-9        # eigenvectors within 1e-9 in |cos| after sign alignment


def _check(lam, vec, ref_lam, ref_vec, k, floor=1e-6):
    scale = np.abs(ref_lam).max()
    big = np.abs(ref_lam) > floor * scale
    assert np.abs(lam[big] - ref_lam[big]).max() / 1.0 <= EVAL_RTOL * np.abs(ref_lam[big]).max() or \
        (np.abs(lam[big] - ref_lam[big]) / np.abs(ref_lam[big])).max() <= EVAL_RTOL
    assert (np.abs(lam[big] - ref_lam[big]) / np.abs(ref_lam[big])).max() <= EVAL_RTOL
    assert np.abs(lam[~big] - ref_lam[~big]).max(initial=0.0) <= EVAL_RTOL * scale
    for i in range(k):
        c = abs(float(vec[i] @ ref_vec[i]))
        assert abs(c - 1.0) <= COS_TOL, (i, c)
        assert abs(np.linalg.norm(vec[i]) - 1.0) < 1e-12


@pytest.mark.parametrize("n", [1, 2, 3, 5, 33, 64, 130, 517])
def test_eigvecs_random_symmetric(ctx, n):
    rs = np.random.RandomState(n)
    A = rs.randn(n, n); A = (A + A.T) / 2 + np.diag(np.linspace(0, 3, n))
    lam, vec = ctx.eigvecs(A, nvec=min(n, 6))
    w, v = np.linalg.eigh(A)
    _check(lam, vec, w[::-1], v[:, ::-1].T, min(n, 6), floor=0.0 if n < 4 else 1e-6)


def test_eigvecs_example_golden(ctx):
    """the 5x5 GRM of POPGEN/par.example: eigenvalues must reproduce POPGEN/example.eval at print precision."""
    import os
    gold = os.path.join(os.path.dirname(__file__), "golden")
    P = np.fromfile(os.path.join(gold, "example.packedancestrymapgeno"), dtype=np.uint8)[48:].reshape(7, 48)
    ctx.upload_packed(P, 5); ctx.set_rows(None)
    r = ctx.grm(altnormstyle=0, want_xtx=True)
    lam, vec = ctx.eig(2)
    want = np.loadtxt(os.path.join(gold, "example.eval"))
    assert ["%12.6f" % x for x in lam] == ["%12.6f" % x for x in want] or np.abs(lam - want).max() < 5e-7


def test_grm_eig_vs_reference(ctx):
    nsnp, nind = 6000, 400
    g = synth.genotypes(21, nsnp, nind, missing=0.05, npops=5, delta=0.25)
    P = synth.pack(g)
    ctx.upload_packed(P, nind); ctx.set_rows(None)
    r = ctx.grm(want_xtx=True)
    lam, vec = ctx.eig(10)
    if ob.ref() is not None:
        o = ob.ref_grm(P, nind); rl, rv = ob.ref_eigvecs(o["XTX"] / o["y"])
    else:
        o = ob.port_grm(P, nind); rl, rv = ob.port_eigvecs(o["XTX"] / o["y"])
    _check(lam, vec, rl, rv, 4)      # 5 populations -> 4 structured axes with real gaps
    # eigenvector residuals for all 10 against the GPU's own GRM
    A = r["XTX"]
    for i in range(10):
        assert np.linalg.norm(A @ vec[i] - lam[i] * vec[i]) < 1e-10 * abs(lam[0])


def test_ridoutlier_bit_exact():
    from eig_b200 import capi
    rs = np.random.RandomState(0)
    E = rs.randn(6, 200) / np.sqrt(200); E[2, 17] = 0.9; E[4, 3] = -1.1
    for mode in (0, 1):
        b1, v1, s1 = capi.ridoutlier(E, 6, 6.0, mode)
        b2, v2, s2 = ob.port_ridoutlier(E, 6, 6.0, mode)
        assert np.array_equal(b1, b2) and np.array_equal(v1, v2) and np.array_equal(s1[b1], s2[b2])
        if ob.ref() is not None:
            b3, v3, s3 = ob.ref_ridoutlier(E, 6, 6.0, mode)
            assert np.array_equal(b1, b3) and np.array_equal(v1, v3) and np.array_equal(s1[b1], s3[b3])


def test_pca_full_outlier_loop(ctx):
    """outlier iterations that actually fire: two planted outlier individuals (own population, big drift)."""
    nsnp, nind = 5000, 300
    pd = np.array([0.05] * 3 + [1.5])
    g = synth.genotypes(4, nsnp, nind, missing=0.02, npops=4, pop_delta=pd)
    # make population 3 tiny: individuals 298,299 keep their drifted genotypes, the rest of pop 3 copy pop 0 behaviour
    pop = synth.pop_of(nind, 4)
    g0 = synth.genotypes(4, nsnp, nind, missing=0.02, npops=4, pop_delta=np.array([0.05, 0.05, 0.05, 0.05]))
    sel = (pop == 3) & (np.arange(nind) < 298)
    g[:, sel] = g0[:, sel]
    P = synth.pack(g)
    ctx.upload_packed(P, nind)
    res = ctx.pca_full(numeigs=5, numoutliter=5, numoutleigs=5, outlthresh=6.0)
    # reference loop (smartpca.c:1077-1265) with the oracle pieces
    xi = np.arange(nind, dtype=np.int32); removed = []
    ignore = np.zeros(nsnp, bool)
    use_ref = ob.ref() is not None
    for it in range(1, 7):
        Pk = P[~ignore]
        o = (ob.ref_grm if use_ref else ob.port_grm)(Pk, nind, xindex=xi)
        idx = np.flatnonzero(~ignore); ignore[idx[o["used"] == 0]] = True
        lam, vec = (ob.ref_eigvecs if use_ref else ob.port_eigvecs)(o["XTX"] / o["y"])
        if it > 5:
            break
        bad, vecno, score = ob.port_ridoutlier(vec[:5], 5, 6.0, 0)
        if len(bad) == 0:
            break
        removed += [(int(xi[j]), it, int(vecno[j])) for j in bad]
        xi = np.delete(xi, bad)
    assert len(removed) > 0, "test data must trigger outlier removal"
    assert [(int(a), int(b), int(c)) for a, b, c in zip(res["removed_index"], res["removed_iter"], res["removed_vecno"])] == removed
    assert np.array_equal(res["xindex"], xi)
    assert (np.abs(res["lambda_"] - lam) / np.maximum(np.abs(lam), 1e-6 * lam[0])).max() < 1e-9
    for i in range(2):
        assert abs(abs(res["evecs"][i] @ vec[i]) - 1) < 1e-9


@pytest.mark.parametrize("n", [130, 700, 2300])
def test_full_eigenbasis(ctx, n):
    """all n eigenvectors (what eigvecs() promises, eigsubs.c:39-55, and shrinkmode consumes): batched inverse iteration,
    Gram-Schmidt only inside numerically coincident groups, blocked back-transformation on the tensor cores.  The matrix has
    a rank-deficient tail (a group of ~n/10 coincident zero eigenvalues) and population structure on top."""
    rs = np.random.RandomState(n)
    m = n - n // 10
    X = rs.randn(n, m) + 0.5 * rs.randn(3, m)[rs.randint(0, 3, n)]
    A = X @ X.T / m
    ctx.set_option("eig_method", 1)
    try:
        lam, vec = ctx.eigvecs(A)
    finally:
        ctx.set_option("eig_method", 0)
    w = np.linalg.eigvalsh(A)[::-1]
    assert np.abs(lam - w).max() <= 1e-11 * w[0]
    assert vec.shape == (n, n)
    assert np.abs(vec @ vec.T - np.eye(n)).max() < 1e-9
    R = vec @ A - lam[:, None] * vec
    assert np.abs(R).max() <= 1e-11 * w[0]
