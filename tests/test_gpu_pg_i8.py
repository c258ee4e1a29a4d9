"""GPU parity of the integer tensor-core packed x skinny products (pg_i8.cu: in-kernel decode of the 2-bit matrix into byte bases,
7-bit digit rows of the FP64 operand, tcgen05.mma kind::i8) against the FP64 DMMA kernels they replace (which the other tests pin to
the reference): SNP loadings / projections, the .evec coordinate sequence with lsqproj, fastmode, shrinkmode."""
import numpy as np
import pytest

from eig_b200 import synth

pytestmark = pytest.mark.gpu


def _with(ctx, method, fn):
    ctx.set_option("pg_method", method)
    try:
        return fn()
    finally:
        ctx.set_option("pg_method", 0)


@pytest.mark.parametrize("nsnp,nind,miss,k", [(700, 260, 0.1, 3), (3000, 1000, 0.05, 10), (2500, 129, 0.0, 33)])
def test_projection_products(ctx, nsnp, nind, miss, k):
    P = synth.pack(synth.genotypes(17, nsnp, nind, missing=miss, npops=3, delta=0.25))
    ctx.upload_packed(P, nind); ctx.set_rows(None)
    ctx.grm(want_snp=False)
    rs = np.random.RandomState(3)
    ev = np.linalg.qr(rs.randn(nind, k))[0].T.copy()
    ev *= 10.0 ** rs.uniform(-3, 3, size=(k, 1))             # columns of very different scale: one digit scale per column
    f1, x1, s1 = _with(ctx, 1, lambda: ctx.project(ev))
    f2, x2, s2 = _with(ctx, 2, lambda: ctx.project(ev))
    for a, b in ((f1, f2), (x1, x2), (s1, s2)):
        scale = np.abs(a).max(axis=-1, keepdims=True) if a.ndim == 2 else np.abs(a)
        assert (np.abs(a - b) / scale).max() <= 1e-12


def test_evec_coords_and_lsqproj(ctx):
    nsnp, nind = 2600, 700
    P = synth.pack(synth.genotypes(5, nsnp, nind, missing=0.08, npops=4, delta=0.3))
    ctx.upload_packed(P, nind)
    xi = np.arange(0, nind - 50, dtype=np.int32)              # the last 50 individuals are projected (lsqproj)
    ctx.set_rows(xi)
    ctx.grm(want_snp=False)
    lam, vec = ctx.eig(5)
    c1, e1, k1 = _with(ctx, 1, lambda: ctx.evec_coords(vec))
    c2, e2, k2 = _with(ctx, 2, lambda: ctx.evec_coords(vec))
    assert np.array_equal(k1, k2)
    assert np.abs(e1 - e2).max() <= 1e-12 * np.abs(e1).max()
    assert np.abs(c1 - c2).max() <= 1e-11 * np.abs(c1).max()


def test_fastmode(ctx):
    nsnp, nind = 4000, 600
    P = synth.pack(synth.genotypes(7, nsnp, nind, missing=0.03, npops=6, pop_delta=np.linspace(0.2, 0.5, 6)))
    ctx.upload_packed(P, nind); ctx.set_rows(None)
    ev1, v1 = _with(ctx, 1, lambda: ctx.fpca(4, 8, 3, seed=11))
    ev2, v2 = _with(ctx, 2, lambda: ctx.fpca(4, 8, 3, seed=11))
    assert (np.abs(ev1 - ev2) / ev1).max() <= 1e-10
    for j in range(4):
        assert abs(abs(float(v1[:, j] @ v2[:, j])) - 1.0) <= 1e-9


def test_shrinkmode(ctx):
    nsnp, nind = 1500, 300
    P = synth.pack(synth.genotypes(9, nsnp, nind, missing=0.05, npops=3, delta=0.3))
    ctx.upload_packed(P, nind); ctx.set_rows(None)
    ctx.grm(want_snp=False)
    for new in (False, True):
        a1, l1, o1 = _with(ctx, 1, lambda: ctx.shrink_coords(3, newshrink=new))
        a2, l2, o2 = _with(ctx, 2, lambda: ctx.shrink_coords(3, newshrink=new))
        assert o1.all() and o2.all() and np.abs(l1 - l2).max() <= 1e-12 * l1[0]
        sg = np.sign((a1 * a2).sum(1))
        assert np.abs(a1 - a2 * sg[:, None]).max() <= 1e-8 * np.abs(a1).max()
