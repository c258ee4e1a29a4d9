"""CPU tests: the C-ABI library loads, exports every symbol include/eigb200.h declares, its host-side logic matches the
oracle bit for bit, and compute entry points fail loudly without a GPU (no fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from eig_b200 import capi
from oracle import bindings as ob

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "eigb200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    # eb_* entry points plus the drop-in symbols that keep the reference's own names (include/eigsubs.h:6-7, include/kjg_fpca.h:22)
    return sorted(set(re.findall(r"\b(eb_[a-z0-9_]+|eigvecs|eigvals|kjg_fpca)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    L = capi.lib()
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), n
    assert sorted(capi.EXPORTS) == names
    assert L.eb_version() >= 100


def test_gauss_matrix_bit_exact():
    for seed, n, Lw in [(0, 10, 4), (7, 333, 20), (123456, 64, 5)]:
        assert np.array_equal(capi.gauss_matrix(seed, n, Lw), ob.port_gauss(seed, n, Lw))
        if ob.ref() is not None:
            assert np.array_equal(capi.gauss_matrix(seed, n, Lw), ob.ref_gauss(seed, n, Lw))


def test_ridoutlier_host_logic_bit_exact():
    rs = np.random.RandomState(0)
    E = rs.randn(6, 200) / np.sqrt(200); E[2, 17] = 0.9; E[4, 3] = -1.1; E[0, 17] = 0.7
    for mode in (0, 1):
        b1, v1, s1 = capi.ridoutlier(E, 6, 6.0, mode)
        b2, v2, s2 = ob.port_ridoutlier(E, 6, 6.0, mode)
        assert len(b1) > 0 and np.array_equal(b1, b2) and np.array_equal(v1, v2) and np.array_equal(s1[b1], s2[b2])
    assert len(capi.ridoutlier(E, 6, 6.0, 2)[0]) == 0          # outliermode 2 = off (smartsubs.c:30)
    assert len(capi.ridoutlier(E[:, :2], 2, 6.0, 0)[0]) == 0   # n < 3 (smartsubs.c:32)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.EigB200Error):
        capi.Context(0)


def test_header_is_plain_c_and_shim_links(tmp_path):
    """include/eigb200.h compiles as C99 with -Wall -Werror and the integration shim (examples/smartpca_shim.c: the calls
    INTEGRATION.md adds to smartpca.c, single- and multi-GPU) links against libeigb200.so"""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    capi.lib()
    src = os.path.join(ROOT, "examples", "smartpca_shim.c")
    out = str(tmp_path / "libshim.so")
    cmd = [gcc, "-std=c99", "-D_POSIX_C_SOURCE=200809L", "-Wall", "-Werror", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"), src, "-o", out,
           "-L", os.path.join(ROOT, "eig_b200"), "-leigb200", "-lpthread", "-Wl,-rpath," + os.path.join(ROOT, "eig_b200"), "-Wl,--no-undefined"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    L = ctypes.CDLL(out)
    assert hasattr(L, "eb_shim_full_mode") and hasattr(L, "eb_shim_threads")


def test_tw_stats_reproduce_twexample_golden():
    """POPGEN/twexample.eval -> twexample.out (the reference's own fixture for twstats / the Tracy-Widom block of smartpca):
    twstat and effective n columns at print precision; suffix-sum formulation vs the reference's per-eigenvalue re-summation"""
    gold = os.path.join(ROOT, "tests", "golden")
    lam = np.loadtxt(os.path.join(gold, "twexample.eval"))
    lam = -np.sort(-lam)
    rows = [l.split() for l in open(os.path.join(gold, "twexample.out")) if l.strip() and not l.strip().startswith("#")]
    tw, zn = capi.tw_stats(lam, znval=-1.0, minm=10)
    assert len(rows) == len(tw) == capi.lib().eb_numgtz(lam.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(len(lam)))
    checked = 0
    for i, r in enumerate(rows):
        assert abs(float(r[1]) - lam[i]) < 5.1e-7
        if r[3] == "NA":
            assert tw[i] == -1.0 and zn[i] == -1.0
            continue
        assert abs(float(r[3]) - tw[i]) <= 5.1e-4, (i, r[3], tw[i])
        assert abs(float(r[5]) - zn[i]) <= 5.1e-4 * max(1.0, abs(zn[i]) * 1e-3), (i, r[5], zn[i])
        checked += 1
    assert checked > 100
    # p-value column through the table the reference ships (POPGEN/twtable == include/twtable.h), 6 significant digits
    table = np.loadtxt(os.path.join(gold, "twtable"), comments="#")
    assert table.shape == (161, 3)
    pv = capi.tw_tail(tw, table)
    for i, r in enumerate(rows):
        if r[3] == "NA":
            continue
        want = float(r[4])
        assert abs(pv[i] - want) <= 6e-6 * max(abs(want), 1e-300) + (1e-12 if want > 1e-6 else 0.0), (i, r[4], pv[i])
    # plain re-summation (the reference's order of operations) agrees to rounding
    m = len(tw)
    for i in (0, 1, 57, m - 11):
        ev = lam[i:m] * ((m - i) / lam[i:m].sum())
        z = (m - i) * (m - i + 2) / ((ev * ev).sum() - (m - i))
        assert abs(z - zn[i]) <= 1e-9 * abs(z)


def test_scripts_parse():
    """bench.py, the driver entry and every probe under tools/ are at least syntactically valid (they only run on a GPU box)"""
    import ast
    import glob
    files = [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")] + sorted(glob.glob(os.path.join(ROOT, "tools", "*.py"))) + \
        [os.path.join(ROOT, "tests", "multi_worker.py")]
    assert len(files) > 10
    for f in files:
        ast.parse(open(f).read(), filename=f)
