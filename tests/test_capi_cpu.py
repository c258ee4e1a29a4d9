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
    # eb_* entry points plus the two drop-in symbols that keep the reference's own names (include/eigsubs.h:6-7)
    return sorted(set(re.findall(r"\b(eb_[a-z0-9_]+|eigvecs|eigvals)\s*\(", txt)))


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
