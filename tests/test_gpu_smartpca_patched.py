"""Drop-in test of the whole seam: the reference's smartpca with integration/smartpca_b200.patch applied, linked against
libeigb200.so and NO LAPACK / BLAS / GSL (integration/Makefile), run as `smartpca -p parfile`.

  * POPGEN/par.example and EIGENSTRAT/example.pca.par: .evec / .eval / grmjunk byte-identical to the reference's checked-in goldens
  * generated data (missing genotypes, populations, planted outliers): full mode with outlier removal + lsqproject of a
    population left out of the PCA, shrinkmode, and fastmode -- against the UNMODIFIED reference binary (oracle/_ref/smartpca)
    run on the same par file: eigenvalues 1e-9 relative, .evec entries 1e-6 absolute after sign alignment (north_star)
Both binaries are built where /root/reference exists and travel to the GPU box as built artefacts; the test skips when they
are absent.
"""
import os
import subprocess

import numpy as np
import pytest

from eig_b200 import synth

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
PATCHED = os.path.join(ROOT, "integration", "_build", "smartpca_b200")
REFBIN = os.path.join(ROOT, "oracle", "_ref", "smartpca")


def _env():
    env = dict(os.environ)
    extra = ["/usr/local/cuda/lib64"]
    try:
        import nvidia.cuda_runtime as cr
        extra.insert(0, os.path.join(os.path.dirname(cr.__file__), "lib"))
    except Exception:
        pass
    env["LD_LIBRARY_PATH"] = ":".join(extra + [env.get("LD_LIBRARY_PATH", "")])
    return env


def _run(binary, parfile, cwd):
    r = subprocess.run([binary, "-p", parfile], cwd=cwd, capture_output=True, text=True, env=_env(), timeout=600)
    assert r.returncode == 0, "%s failed (%d):\n%s\n%s" % (binary, r.returncode, r.stdout[-3000:], r.stderr[-3000:])
    return r.stdout


def _eval_bytes(b):
    """.eval bytes with the sign of a printed zero removed: centred data make the smallest eigenvalue exactly 0 in exact arithmetic,
    so what is printed is the sign of its rounding error (LAPACK's happens to be negative in the reference's goldens)"""
    return b.replace(b"-0.000000", b" 0.000000")


def _need(*paths):
    for p in paths:
        if not os.path.exists(p):
            pytest.skip("%s was not built (needs /root/reference at build time)" % os.path.relpath(p, ROOT))


def test_par_example_byte_identical(tmp_path):
    """POPGEN/example.perl: smartpca -p par.example (PED input, altnormstyle NO, numoutevec 2, grmoutname)"""
    _need(PATCHED)
    g = os.path.join(GOLD, "popgen")
    par = tmp_path / "par.example"
    par.write_text("genotypename: %s/example.ped\nsnpname: %s/example.map\nindivname: %s/example.ped\nevecoutname: example.evec\n"
                   "evaloutname: example.eval\naltnormstyle: NO\nnumoutevec: 2\nfamilynames: NO\ngrmoutname: grmjunk\n" % (g, g, g))
    out = _run(PATCHED, str(par), str(tmp_path))
    assert "libeigb200:" in out                       # the GPU path ran
    for name in ("example.evec", "grmjunk", "grmjunk.id"):
        assert (tmp_path / name).read_bytes() == open(os.path.join(g, name), "rb").read(), name
    assert _eval_bytes((tmp_path / "example.eval").read_bytes()) == _eval_bytes(open(os.path.join(g, "example.eval"), "rb").read())


def test_eigenstrat_example_byte_identical(tmp_path):
    """EIGENSTRAT/example.perl -> example.pca.par (EIGENSTRAT input, numoutlieriter 5, numoutlierevec 2)"""
    _need(PATCHED)
    g = os.path.join(GOLD, "eigenstrat")
    par = tmp_path / "example.pca.par"
    txt = open(os.path.join(g, "example.pca.par")).read()
    for k in ("example.geno", "example.snp", "example.ind"):
        txt = txt.replace(": " + k, ": " + os.path.join(g, k))
    par.write_text(txt)
    _run(PATCHED, str(par), str(tmp_path))
    assert (tmp_path / "example.pca.evec").read_bytes() == open(os.path.join(g, "example.pca.evec"), "rb").read()
    assert _eval_bytes((tmp_path / "example.eval").read_bytes()) == _eval_bytes(open(os.path.join(g, "example.eval"), "rb").read())


def _read_evec(path):
    rows = open(path).read().split("\n")
    lam = np.array([float(x) for x in rows[0].split()[1:]])
    ids, vals, pops = [], [], []
    for r in rows[1:]:
        f = r.split()
        if not f:
            continue
        ids.append(f[0]); vals.append([float(x) for x in f[1:-1]]); pops.append(f[-1])
    return lam, ids, np.array(vals), pops


def _dataset(tmp_path, nind=240, nsnp=4000, missing=0.05, outliers=True, empty_indiv=True):
    pd = np.array([0.05, 0.3, 0.3, 1.6]) if outliers else np.array([0.05, 0.3, 0.3, 0.3])
    g = synth.genotypes(21, nsnp, nind, missing=missing, npops=4, pop_delta=pd)
    if outliers:                                     # only the last two members of pop 3 stay extreme: planted outliers
        g0 = synth.genotypes(21, nsnp, nind, missing=missing, npops=4, pop_delta=np.array([0.05, 0.3, 0.3, 0.3]))
        sel = (synth.pop_of(nind, 4) == 3) & (np.arange(nind) < nind - 2)
        g[:, sel] = g0[:, sel]
    if empty_indiv:
        g[:, 5] = -1                                 # an individual without data (ignored by smartpca.c:870-878)
    P = synth.pack(g)
    return synth.write_dataset(str(tmp_path / "syn"), P, nind, npops=4)


PAR = ("genotypename: {p}.geno\nsnpname: {p}.snp\nindivname: {p}.ind\nevecoutname: {o}.evec\nevaloutname: {o}.eval\n"
       "numoutevec: {k}\nhashcheck: NO\nhiprec: YES\n")


def _both(tmp_path, extra, k=4, **data):
    _need(PATCHED, REFBIN)
    prefix = _dataset(tmp_path, **data)
    outs = {}
    for tag, binary in (("gpu", PATCHED), ("ref", REFBIN)):
        par = tmp_path / ("par." + tag)
        par.write_text(PAR.format(p=prefix, o=tag, k=k) + extra)
        outs[tag] = _run(binary, str(par), str(tmp_path))
    return outs


def _hiprec_compare(tmp_path, k, atol=1e-6):
    la, ida, va, pa = _read_evec(str(tmp_path / "gpu.evec"))
    lb, idb, vb, pb = _read_evec(str(tmp_path / "ref.evec"))
    assert ida == idb and pa == pb, "different individuals in the two .evec files"
    assert np.abs(la - lb).max() <= 1.001e-3
    sg = np.sign((va * vb).sum(0)); sg[sg == 0] = 1
    err = np.abs(va - vb * sg).max()
    assert err <= atol + 1.0001e-6, err           # hiprec prints 6 decimals: 1e-6 absolute + one unit of print rounding


def test_full_mode_outliers_lsqproject_vs_reference_binary(tmp_path):
    """full mode, numoutlieriter 5 with outliers that fire, poplistname (pop 0 is projected by lsqproj), all eigenvalues"""
    (tmp_path / "poplist").write_text("Pop1\nPop2\nPop3\n")
    outs = _both(tmp_path, "numoutlieriter: 5\nnumoutlierevec: 3\npoplistname: %s\noutliername: x.outliers\n" % (tmp_path / "poplist"))
    assert "REMOVED outlier" in open(tmp_path / "x.outliers").read()
    ev_g = np.loadtxt(tmp_path / "gpu.eval"); ev_r = np.loadtxt(tmp_path / "ref.eval")
    assert ev_g.shape == ev_r.shape
    assert np.abs(ev_g - ev_r).max() <= 1.0001e-6                  # .eval prints 6 decimals
    _hiprec_compare(tmp_path, 4)
    # the log lines a user greps survive, in the same order
    keep = lambda s: [l for l in s.split("\n") if l.startswith(" snp ") or l.startswith("total number of snps killed") or
                      l.startswith("number of samples after outlier removal") or "ignored (insufficient data" in l]
    assert keep(outs["gpu"]) == keep(outs["ref"])


def test_shrinkmode_vs_reference_binary(tmp_path):
    outs = _both(tmp_path, "numoutlieriter: 0\nshrinkmode: YES\n", k=3)
    assert "doshrink called" in outs["gpu"]
    _hiprec_compare(tmp_path, 3)


def test_fastmode_vs_reference_binary(tmp_path):
    """fastmode: YES -> setgval (integration/eb_gval.c) + kjg_fpca (library symbol); K = 4, L = 8, I = 4, fixed seed"""
    outs = _both(tmp_path, "fastmode: YES\nseed: 77\n", k=4, empty_indiv=False, outliers=False)      # printevecs needs nrows == numindivs
    assert "end of smartpca(fastmode)" in outs["gpu"]
    _hiprec_compare(tmp_path, 4)


def test_usepopsformissing_vs_reference_binary(tmp_path):
    """usepopsformissing: YES -> the dense path: eb_grm_popfill on the GPU (population fill, normalisation, GRM) + eb_eig; the
    projection passes stay with the reference's host code (their columns are population dependent)"""
    outs = _both(tmp_path, "numoutlieriter: 2\nusepopsformissing: YES\n", k=3, missing=0.2)
    assert "libeigb200:" in outs["gpu"]
    ev_g = np.loadtxt(tmp_path / "gpu.eval"); ev_r = np.loadtxt(tmp_path / "ref.eval")
    assert ev_g.shape == ev_r.shape and np.abs(ev_g - ev_r).max() <= 1.0001e-6
    _hiprec_compare(tmp_path, 3)
