"""GPU parity: the dense GRM path (domult_increment_normal, smartpca.c:3531-3561) and the eigvecs()/eigvals() drop-in
symbols (include/eigsubs.h:6-7)."""
import ctypes as C

import numpy as np
import pytest

from eig_b200 import capi
from oracle import bindings as ob

pytestmark = pytest.mark.gpu


def _columns(seed, ncols, nrows, missing=0.1):
    """centred, scaled columns with imputed (non table-valued) entries, like getcolxz under usepopsformissing"""
    rs = np.random.RandomState(seed)
    p = rs.uniform(0.05, 0.95, ncols)
    g = rs.binomial(2, p[:, None], size=(ncols, nrows)).astype(np.float64)
    miss = rs.rand(ncols, nrows) < missing
    g[miss] = (2 * p[:, None] + 0.1 * rs.randn(ncols, nrows))[miss]       # population-mean style imputation
    g -= g.mean(axis=1, keepdims=True)
    g /= np.sqrt(p * (1 - p))[:, None]
    return g


@pytest.mark.parametrize("ncols,nrows", [(700, 97), (2500, 300), (1030, 513)])
def test_dense_grm_vs_oracle(ctx, ncols, nrows):
    T = _columns(ncols, ncols, nrows)
    y, X = ctx.grm_dense([T[:1000], T[1000:]] if ncols > 1000 else [T], nrows, want_xtx=True)
    ref = ob.port_dense_grm(T)
    yr = np.trace(ref) / (nrows - 1)
    assert abs(y - yr) <= 1e-12 * yr
    assert np.abs(X - ref / yr).max() <= 1e-11 * np.abs(ref / yr).max()
    assert np.array_equal(X, X.T)
    if ob.ref() is not None:
        rr = ob.ref_dense_grm(T, blocksize=1024)
        assert np.abs(X - rr / yr).max() <= 1e-11 * np.abs(rr / yr).max()
    # the resident GRM feeds the eigensolver like the packed path
    lam, vec = ctx.eig(3)
    w = np.linalg.eigvalsh(ref / yr)[::-1]
    assert np.abs(lam - w).max() <= 1e-9 * w[0]


def test_dense_rejects_uncentred_column(ctx):
    T = _columns(1, 10, 50)
    T[3] += 1.0
    with pytest.raises(capi.EigB200Error, match="ycheck"):
        ctx.grm_dense([T], 50)


def test_dropin_eigvecs_symbols():
    L = capi.lib()
    for n in (2, 5, 40, 300):
        rs = np.random.RandomState(n)
        A = np.eye(2) if n == 2 else (lambda B: (B + B.T) / 2)(rs.randn(n, n))      # pcatoy.c: 2 x 2 identity
        mat = A.copy(); ev = np.empty(n); vec = np.empty((n, n))
        L.eigvecs(mat.ctypes.data_as(C.c_void_p), ev.ctypes.data_as(C.c_void_p), vec.ctypes.data_as(C.c_void_p), C.c_int(n))
        assert np.array_equal(mat, A)                                 # mat preserved (eigsubs.c:39-55)
        w, v = np.linalg.eigh(A)
        assert np.abs(ev - w[::-1]).max() <= 1e-12 * max(1.0, np.abs(w).max())
        assert np.abs(vec @ vec.T - np.eye(n)).max() < 1e-10          # all n vectors, orthonormal rows
        assert np.abs(vec @ A @ vec.T - np.diag(ev)).max() < 1e-10 * max(1.0, np.abs(w).max())
        ev2 = np.empty(n)
        L.eigvals(mat.ctypes.data_as(C.c_void_p), ev2.ctypes.data_as(C.c_void_p), C.c_int(n))
        assert np.array_equal(ev, ev2)
        if ob.ref() is not None:
            rl, rv = ob.ref_eigvecs(A)
            assert np.abs(ev - rl).max() <= 1e-12 * max(1.0, np.abs(w).max())


def test_dropin_eigvecs_fills_all_vectors_beyond_8192():
    """the reference's eigvecs() fills all n vectors at any n (eigsubs.c:39-55); round 1 returned 40 + zeros above n = 8192"""
    L = capi.lib()
    n = 8320
    rs = np.random.RandomState(1)
    B = rs.randn(n, 300)
    A = B @ B.T / 300 + np.diag(np.linspace(0.0, 1.0, n))
    mat = A.copy(); ev = np.empty(n); vec = np.empty((n, n))
    L.eigvecs(mat.ctypes.data_as(C.c_void_p), ev.ctypes.data_as(C.c_void_p), vec.ctypes.data_as(C.c_void_p), C.c_int(n))
    assert np.array_equal(mat, A)
    assert np.all(np.diff(ev) <= 1e-12)
    nrm = np.linalg.norm(vec, axis=1)
    assert np.abs(nrm - 1).max() < 1e-10                             # every one of the n rows is a unit vector (none left zero)
    idx = np.r_[0:4, n // 2:n // 2 + 4, n - 4:n]
    V = vec[idx]
    assert np.abs(V @ V.T - np.eye(len(idx))).max() < 1e-9
    R = A @ V.T - V.T * ev[idx]
    assert np.abs(R).max() <= 1e-9 * np.abs(ev).max()
    # sign convention: the entry of largest magnitude of every vector is positive
    assert np.all(vec[np.arange(n), np.abs(vec).argmax(axis=1)] > 0)


def test_dense_syrk_many_tiles_per_cta(ctx):
    """4608 rows: the tensor-core SYRK runs 4-5 output tiles per persistent CTA (see test_two_stage_tridiagonal_many_tiles_per_cta)"""
    n = 4608
    rs = np.random.RandomState(5)
    T = rs.randn(192, n); T -= T.mean(axis=1, keepdims=True)
    want = T.T @ T
    for rep in range(3):
        y, X = ctx.grm_dense([T[:64], T[64:]], n, want_xtx=True)
        assert np.abs(X * y - want).max() <= 1e-12 * np.abs(want).max()
