"""CPU tests (-m "not gpu"): pin the oracle against the reference's own golden files and, when it is available,
against the unmodified reference compiled into oracle/_ref (see oracle/Makefile).  No GPU, no product code paths."""
import os

import numpy as np
import pytest

from eig_b200 import synth
from oracle import bindings as ob

GOLD = os.path.join(os.path.dirname(__file__), "golden")
HAVE_REF = ob.ref() is not None
needs_ref = pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built (no /root/reference here)")


def _example():
    raw = np.fromfile(os.path.join(GOLD, "example.packedancestrymapgeno"), dtype=np.uint8)
    return raw[:48], raw[48:].reshape(7, 48)


def test_packed_layout_kat():
    """CONVERTF/example.packedancestrymapgeno vs CONVERTF/example.eigenstratgeno: admutils.c:718-735 layout."""
    hdr, P = _example()
    assert hdr.tobytes().startswith(b"GENO       5       7 d28ee92a 7a6c3691")
    assert P[:, 0].tolist()[:7] == [0x54, 0x19, 0x94, 0x06, 0x94, 0x05, 0xA5]
    txt = [l.strip() for l in open(os.path.join(GOLD, "example.eigenstratgeno"))]
    want = np.array([[int(ch) for ch in l] for l in txt], dtype=np.int8)
    assert np.array_equal(synth.unpack(P, 5), want)
    ours = synth.pack(want)                                      # the file pads with zero bits, synth pads with code 3
    assert np.array_equal(ours[:, 0], P[:, 0]) and np.array_equal(ours[:, 1] & 0xC0, P[:, 1] & 0xC0)


def test_port_reproduces_example_goldens():
    """POPGEN/par.example (altnormstyle NO): POPGEN/example.eval, POPGEN/grmjunk, POPGEN/example.evec directions."""
    _, P = _example()
    o = ob.port_grm(P, 5, altnormstyle=0)
    A = o["XTX"] / o["y"]
    lam, vec = ob.port_eigvecs(A)
    want = np.loadtxt(os.path.join(GOLD, "example.eval"))
    assert np.abs(lam - want).max() < 5e-7
    # grmjunk: lower triangle scaled to mean diagonal 1 (smartpca.c:3789-3805)
    G = A / (np.trace(A) / 5)
    for line in open(os.path.join(GOLD, "grmjunk")):
        i, j, ns, v = line.split()
        assert int(ns) == 7 and abs(G[int(i) - 1, int(j) - 1] - float(v)) < 5e-7
    ev = np.array([[float(x) for x in l.split()[1:3]] for l in open(os.path.join(GOLD, "example.evec")) if not l.lstrip().startswith("#")])
    for k in range(2):
        c = abs(ev[:, k] @ vec[k]) / np.linalg.norm(ev[:, k])
        assert abs(c - 1) < 1e-7       # file has 4 decimals


def test_mt19937_known_answers():
    """GSL manual, 'Random number environment variables': mt19937 seed 0 -> 4293858116, seed 123 -> 2991312382."""
    assert ob.port().orc_mt_first(0) == 4293858116
    assert ob.port().orc_mt_first(123) == 2991312382
    rs = np.random.RandomState(20140428)      # independent MT19937 (same init_genrand seeding)
    assert ob.port().orc_mt_first(20140428) == int(rs.randint(0, 2 ** 32, dtype=np.uint64))


def test_gauss_matrix_properties():
    G = ob.port_gauss(7, 4000, 5)
    assert abs(G[:, :4].mean()) < 0.03 and abs(G[:, :4].std() - 1) < 0.03
    assert (G[:, 4] > 0).all() and (G[:, 4] < 1).all()      # odd last column is a plain uniform (kjg_gsl.c:183-184)


def test_port_eigvecs_vs_lapack():
    rs = np.random.RandomState(3)
    for n in (1, 2, 3, 17, 120):
        A = rs.randn(n, n); A = A + A.T
        lam, vec = ob.port_eigvecs(A)
        w, v = np.linalg.eigh(A)
        assert np.abs(lam - w[::-1]).max() < 1e-12 * max(1, np.abs(w).max())
        for i in range(n):
            assert abs(abs(vec[i] @ v[:, n - 1 - i]) - 1) < 1e-9


def test_counts_and_normalisation_vs_numpy():
    g = synth.genotypes(5, 500, 77, missing=0.15)
    g[3, :] = -1; g[4, :] = 0
    P = synth.pack(g)
    xi = np.arange(1, 77, 3, dtype=np.int32)
    c0, c1, nm = ob.port_snp_counts(P, xi, 77)
    gs = g[:, xi]
    assert np.array_equal(c0, np.where(gs >= 0, gs, 0).sum(1))
    assert np.array_equal(c1, np.where(gs >= 0, 2 - gs, 0).sum(1))
    assert np.array_equal(nm, (gs < 0).sum(1))
    o = ob.port_grm(P, 77, xindex=xi)
    assert o["used"][3] == 0 and o["nmiss"][3] == -1 and o["c0"][3] == -1      # all missing
    assert o["used"][4] == 0                                                    # monomorphic
    valid = len(xi) - nm
    s = 10
    mean = c0[s] / valid[s]
    p = mean / 2
    assert o["xfancy"][s] == 1 / np.sqrt(p * (1 - p)) and o["xmean"][s] == mean * o["xfancy"][s]
    # dense restatement of XTX
    X = np.where(gs >= 0, (gs - (c0 / np.maximum(valid, 1))[:, None]) * o["xfancy"][:, None], 0.0) * o["used"][:, None]
    assert np.abs(X.T @ X - o["XTX"]).max() < 1e-9


@needs_ref
@pytest.mark.parametrize("alt,fancy,miss", [(1, 1, 0.1), (0, 1, 0.3), (1, 0, 0.0)])
def test_port_grm_vs_reference(alt, fancy, miss):
    g = synth.genotypes(8, 700, 150, missing=miss, npops=2, delta=0.2)
    g[0, :] = -1; g[5, :] = 2
    P = synth.pack(g)
    xi = np.sort(np.random.RandomState(2).choice(150, 120, replace=False)).astype(np.int32)
    w = 0.5 + np.random.RandomState(1).rand(700)
    a = ob.port_grm(P, 150, xindex=xi, fancynorm=fancy, altnormstyle=alt, weights=w, minallelecnt=2, maxmissing=40)
    b = ob.ref_grm(P, 150, xindex=xi, fancynorm=fancy, altnormstyle=alt, weights=w, minallelecnt=2, maxmissing=40, nthreads=3)
    for k in ("c0", "c1", "nmiss", "used"):
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(a["xmean"], b["xmean"]) and np.array_equal(a["xfancy"], b["xfancy"])
    assert np.abs(a["XTX"] - b["XTX"]).max() < 1e-11 * np.abs(b["XTX"]).max()
    assert abs(a["y"] - b["y"]) < 1e-12 * b["y"]


@needs_ref
def test_port_eig_ridoutlier_gauss_fpca_vs_reference():
    g = synth.genotypes(3, 2000, 120, missing=0.05, npops=3, delta=0.3)
    P = synth.pack(g)
    r = ob.ref_grm(P, 120)
    A = r["XTX"] / r["y"]
    l1, v1 = ob.port_eigvecs(A); l2, v2 = ob.ref_eigvecs(A)
    assert np.abs(l1 - l2).max() < 1e-12 * l2[0]
    for i in range(5):
        assert abs(abs(v1[i] @ v2[i]) - 1) < 1e-10
    E = v2[:6].copy(); E[1, 7] = 0.8
    for mode in (0, 1):
        a = ob.port_ridoutlier(E, 6, 6.0, mode); b = ob.ref_ridoutlier(E, 6, 6.0, mode)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2][a[0]], b[2][b[0]])
    assert np.array_equal(ob.port_gauss(11, 300, 8), ob.ref_gauss(11, 300, 8))
    assert np.array_equal(ob.port_gauss(0, 50, 5), ob.ref_gauss(0, 50, 5))
    e1, u1 = ob.port_fpca(P, 120, K=4, L=8, I=3, seed=11)
    e2, u2, _ = ob.ref_fpca(P, 120, K=4, L=8, I=3, seed=11)
    assert (np.abs(e1 - e2) / e2).max() < 1e-10
    assert np.abs(np.abs((u1 * u2).sum(0)) - 1).max() < 1e-10


def test_committed_reference_vectors():
    """tests/golden/ref_vectors.npz was produced by tests/golden/make_golden.py from oracle/_ref (the unmodified
    reference); it pins the port on boxes where /root/reference does not exist."""
    z = np.load(os.path.join(GOLD, "ref_vectors.npz"))
    P = synth.packed_genotypes(int(z["seed"]), int(z["nsnp"]), int(z["nind"]), missing=float(z["missing"]), npops=3, delta=0.3)
    xi = z["xindex"]
    o = ob.port_grm(P, int(z["nind"]), xindex=xi, altnormstyle=int(z["altnormstyle"]))
    for k in ("c0", "c1", "nmiss", "used"):
        assert np.array_equal(o[k], z[k]), k
    assert np.array_equal(o["xmean"], z["xmean"]) and np.array_equal(o["xfancy"], z["xfancy"])
    assert abs(o["y"] - float(z["y"])) < 1e-12 * float(z["y"])
    lam, vec = ob.port_eigvecs(o["XTX"] / o["y"])
    assert np.abs(lam - z["lambda_"]).max() < 1e-11 * z["lambda_"][0]
    for i in range(z["evecs"].shape[0]):
        assert abs(abs(vec[i] @ z["evecs"][i]) - 1) < 1e-10
    assert np.array_equal(ob.port_gauss(int(z["gseed"]), 64, 6), z["gauss"])
    ev, u = ob.port_fpca(P, int(z["nind"]), K=3, L=6, I=2, seed=int(z["gseed"]), xindex=xi, altnormstyle=int(z["altnormstyle"]))
    assert (np.abs(ev - z["fpca_eval"]) / z["fpca_eval"]).max() < 1e-10
    assert np.abs(np.abs((u * z["fpca_evec"]).sum(0)) - 1).max() < 1e-10


def test_port_lsqproj_against_reference():
    """pin of orc_lsqproj / orc_seteigscale: the unmodified reference's post-eigen sequence (smartpca.c:1440-1564)"""
    if ob.ref() is None:
        pytest.skip("oracle/_ref not built")
    from eig_b200 import synth
    nsnp, nind, k = 2000, 90, 4
    g = synth.genotypes(3, nsnp, nind, missing=0.15, npops=3, delta=0.3)
    g[:, 7] = -1; g[3:, 11] = -1
    P = synth.pack(g)
    xi = np.array([i for i in range(nind) if i % 4 != 1 and i not in (7, 11)], dtype=np.int32)
    o = ob.ref_grm(P, nind, xindex=xi)
    lam, vec = ob.ref_eigvecs(o["XTX"] / o["y"])
    r = ob.ref_evec_coords(P, nind, o["used"], o["xmean"], o["xfancy"], vec[:k], xindex=xi)
    c, es, ok, ff, sc = ob.port_evec_coords(P, nind, o["used"], o["xmean"], o["xfancy"], vec[:k], xindex=xi)
    assert np.array_equal(r["ignored"], 1 - ok) and set(np.flatnonzero(ok == 0)) == {7, 11}
    assert np.abs(ff - r["ffvecs"]).max() < 1e-11 and np.abs(sc - r["fxscal"]).max() < 1e-15
    assert np.abs(es - r["eigscale"]).max() <= 1e-12 * np.abs(es).max()
    assert np.abs(c - r["coords"]).max() < 1e-13


def test_port_pop_counts_and_fst_against_reference():
    """pin of orc_pop_counts + orc_fstcol: the reference's fstcolyy (qpsubs.c:1205-1346), bit for bit"""
    if ob.ref() is None:
        pytest.skip("oracle/_ref not built")
    from eig_b200 import synth
    nsnp, nind, npops = 300, 130, 4
    P = synth.pack(synth.genotypes(5, nsnp, nind, missing=0.12, npops=npops, delta=0.3))
    xi = np.arange(nind, dtype=np.int32)[3:]
    xt = synth.pop_of(nind, npops)[xi].astype(np.int32); xt[2] = -1; xt[9] = npops + 3
    en, ed = ob.port_fstcol(ob.port_pop_counts(P, nind, xt, npops, xindex=xi))
    ren, red = ob.ref_fstcol(P, nind, xt, npops, xindex=xi)
    assert np.array_equal(en, ren) and np.array_equal(ed, red)


def test_port_shrink_against_committed_reference_vectors():
    """numpy restatement of doshrinkp / doshrinkp2 (oracle/bindings.py: port_shrink) against outputs of the unmodified
    reference committed in tests/golden/ref_shrink.npz (tests/golden/make_golden_shrink.py), and against the compiled
    reference itself when it is present."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_shrink.npz"))
    nsnp, nind, k = int(g["nsnp"]), int(g["nind"]), int(g["k"])
    P = synth.packed_genotypes(int(g["seed"]), nsnp, nind, missing=float(g["missing"]), npops=4, delta=0.35)
    xi = g["xindex"]
    o = ob.port_grm(P, nind, xindex=xi)
    X = o["XTX"] / o["y"]
    for new, key in ((False, "shrink_old"), (True, "shrink_new")):
        got, lam = ob.port_shrink(P, nind, o["used"], o["xmean"], o["xfancy"], X, k, xindex=xi, newshrink=new)
        want = g[key]
        sg = np.sign((got * want).sum(1))
        assert np.abs(got * sg[:, None] - want).max() < 1e-9
        if ob.ref() is not None:
            rw = ob.ref_shrink(P, nind, o["used"], o["xmean"], o["xfancy"], X, k, xindex=xi, newshrink=new)
            sg = np.sign((rw * want).sum(1))
            assert np.abs(rw * sg[:, None] - want).max() < 1e-9
