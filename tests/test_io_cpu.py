"""CPU tests of the host-side file layer (io.cu): PACKEDANCESTRYMAP header + ID hashes against the reference's own example
files (CONVERTF/example.*), and the .eval / .evec / grm writers byte for byte against POPGEN/example.{eval,evec}, grmjunk."""
import os

import numpy as np
import pytest

from eig_b200 import capi

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _ids():
    ind = [l.split() for l in open(os.path.join(GOLD, "example.ind"))]
    snp = [l.split() for l in open(os.path.join(GOLD, "example.snp"))]
    return [r[0] for r in ind], [r[2] for r in ind], [r[0] for r in snp]


def test_header_and_hashes_match_reference_file():
    ids, groups, snps = _ids()
    h = capi.packed_file_header(os.path.join(GOLD, "example.packedancestrymapgeno"))
    assert (h["nind"], h["nsnp"], h["rlen"], h["file_bytes"]) == (5, 7, 48, 48 * 8)
    # the header was written by the reference's convertf with hasharr over the same .ind / .snp (mcio.c:2402)
    assert capi.hash_ids(ids) == h["ihash"]
    assert capi.hash_ids(snps) == h["shash"]
    assert capi.hash_ids(ids[::-1]) != h["ihash"]          # order dependent


def test_header_rejects_other_files(tmp_path):
    p = tmp_path / "x.geno"
    p.write_bytes(b"0" * 100)
    with pytest.raises(capi.EigB200Error, match="GENO"):
        capi.packed_file_header(str(p))
    with pytest.raises(capi.EigB200Error, match="bad open"):
        capi.packed_file_header(str(tmp_path / "missing"))


def test_eval_writer_reproduces_golden_bytes(tmp_path):
    want = open(os.path.join(GOLD, "example.eval"), "rb").read()
    lam = np.array([3.145365, 0.479177, 0.279889, 0.095570, -0.0000001])
    out = tmp_path / "e.eval"
    capi.write_eval(str(out), lam)
    assert out.read_bytes() == want


def test_evec_writer_reproduces_golden_bytes(tmp_path):
    want = open(os.path.join(GOLD, "example.evec"), "rb").read()
    ids, groups, _ = _ids()
    rows = [l.split() for l in open(os.path.join(GOLD, "example.evec"))][1:]
    coords = np.array([[float(r[1]) for r in rows], [float(r[2]) for r in rows]])
    out = tmp_path / "e.evec"
    capi.write_evec(str(out), [3.145365, 0.479177], ids, groups, coords)
    assert out.read_bytes() == want
    capi.write_evec(str(out), [3.145365, 0.479177], ids, groups, coords, hiprec=True)
    assert out.read_bytes().splitlines()[1].split()[1] == b"0.650200"


def test_grm_writer_reproduces_golden_bytes(tmp_path):
    want = open(os.path.join(GOLD, "grmjunk"), "rb").read()
    rows = [l.split() for l in open(os.path.join(GOLD, "grmjunk"))]
    n = max(int(r[0]) for r in rows)
    X = np.zeros((n, n))
    for a, b, _, v in rows:
        X[int(a) - 1, int(b) - 1] = X[int(b) - 1, int(a) - 1] = float(v)
    # any positive scaling of XTX gives the same file (dumpgrm rescales to mean diagonal 1); the printed values already have it
    out = tmp_path / "g"
    capi.write_grm(str(out), X * 3.7, int(rows[0][2]))
    got = out.read_bytes()
    assert got.splitlines()[0].split()[:3] == want.splitlines()[0].split()[:3]
    g = np.array([float(l.split()[3]) for l in got.splitlines()]); w = np.array([float(l.split()[3]) for l in want.splitlines()])
    assert np.abs(g - w).max() <= 2e-6        # printed values carry 6 decimals


def test_grm_bin_writer_layout(tmp_path):
    """grmbinary: YES (dumpgrmbin, smartpca.c:3704-3766): int32 SNP count per lower-triangle entry, float32 entries / (trace / n)"""
    rng = np.random.default_rng(0)
    a = rng.standard_normal((7, 7)); x = a @ a.T
    capi.write_grm_bin(str(tmp_path / "g"), x, 1234)
    n = 7 * 8 // 2
    cnt = np.fromfile(str(tmp_path / "g.N.bin"), dtype=np.int32)
    val = np.fromfile(str(tmp_path / "g.bin"), dtype=np.float32)
    assert cnt.shape == (n,) and (cnt == 1234).all()
    want = np.array([(x[i, j] / (np.trace(x) / 7)) for i in range(7) for j in range(i + 1)], dtype=np.float64).astype(np.float32)
    assert np.array_equal(val, want)
