"""CPU restatement of the arithmetic behind the integer tensor-core kernels (grm_i8.cu, pg_i8.cu), checked with exact integers:
the two-basis identity x_i x_j = h_i T1[k_j] + v_i T2[k_j], the 7-bit digit decomposition of the table entries, and the claim that the
digit products reproduce the FP64 GRM of the reference's own normalised columns (oracle port) to the precision of the tables."""
from fractions import Fraction

import numpy as np

from eig_b200 import synth
from oracle import bindings as ob

NSL = 8


def _digits(t, E, nsl=NSL):
    """signed 7-bit digits (most significant first) of t on the fixed-point scale 2^(E - 7 nsl), rounded to nearest"""
    q = int(round(abs(t) * 2.0 ** (7 * nsl - E)))
    q = min(q, (1 << (7 * nsl)) - 1)
    d = [(q >> (7 * (nsl - 1 - k))) & 127 for k in range(nsl)]
    return [-x for x in d] if t < 0 else d


def test_two_basis_identity_is_exact():
    rs = np.random.RandomState(0)
    for _ in range(200):
        a, b = Fraction(float(rs.randn())), Fraction(float(rs.rand() + 0.1))
        for ki in (0, 1, 2, 3):
            for kj in (0, 1, 2, 3):
                vi, vj = int(ki != 3), int(kj != 3)
                xi = vi * (a + b * ki) if vi else 0
                xj = vj * (a + b * kj) if vj else 0
                hi = ki * vi
                T1 = (a * b + b * b * kj) if vj else 0
                T2 = (a * a + a * b * kj) if vj else 0
                assert xi * xj == hi * T1 + vi * T2


def test_digits_reconstruct_the_table_entries():
    rs = np.random.RandomState(1)
    vals = np.concatenate([rs.randn(500) * 10.0 ** rs.uniform(-3, 3, 500), [0.0, 1.0, -1.0, 127.0 / 128.0]])
    E = int(np.frexp(np.abs(vals).max())[1])
    for t in vals:
        d = _digits(float(t), E)
        assert all(-127 <= x <= 127 for x in d)
        rec = sum(Fraction(x) * Fraction(2) ** (E - 7 * (k + 1)) for k, x in enumerate(d))
        assert abs(rec - Fraction(float(t))) <= Fraction(2) ** (E - 7 * NSL - 1) + Fraction(2) ** (E - 7 * NSL)   # half a unit (+ the clamp)


def test_integer_formulation_matches_the_port():
    """small matrix, exact integer accumulation of the digit products in Python ints, FP64 only in the final digit sum"""
    nsnp, nind = 60, 23
    g = synth.genotypes(3, nsnp, nind, missing=0.15, npops=2, delta=0.3)
    g[5, :] = np.where(g[5, :] < 0, 1, g[5, :])                      # a SNP without missing genotypes (rank-one path)
    P = synth.pack(g)
    o = ob.port_grm(P, nind)
    used = o["used"].astype(bool)
    xm, xf = o["xmean"], o["xfancy"]
    a = -xm                                                         # table row: (k - mean) * scale = k * xf - xm
    b = xf
    tabs = []
    for s in range(nsnp):
        if not used[s]:
            tabs.append(None); continue
        classF = (g[s] >= 0).all()
        bb, ab, aa = b[s] * b[s], a[s] * b[s], a[s] * a[s]
        T1 = [0.0, bb, 2 * bb] if classF else [ab, ab + bb, ab + 2 * bb]
        T2 = [0.0, 0.0, 0.0] if classF else [aa, aa + ab, aa + 2 * ab]
        tabs.append((classF, T1, T2))
    E = int(np.frexp(max(max(abs(x) for x in t[1] + t[2]) for t in tabs if t))[1])
    C = [[[0] * nind for _ in range(nind)] for _ in range(NSL)]
    r = np.zeros(nind); cc = 0.0
    for s in range(nsnp):
        if tabs[s] is None:
            continue
        classF, T1, T2 = tabs[s]
        d1 = [_digits(T1[k], E) for k in range(3)]
        d2 = [_digits(T2[k], E) for k in range(3)]
        if classF:
            r += a[s] * b[s] * g[s]; cc += a[s] * a[s]
        for i in range(nind):
            if g[s, i] < 0:
                continue
            h = int(g[s, i])
            for j in range(nind):
                if g[s, j] < 0:
                    continue
                kj = int(g[s, j])
                for k in range(NSL):
                    C[k][i][j] += h * d1[kj][k] + d2[kj][k]
    X = np.zeros((nind, nind))
    for k in range(NSL - 1, -1, -1):
        X += np.array(C[k], dtype=np.float64) * 2.0 ** (E - 7 * (k + 1))
    X += r[:, None] + r[None, :] + cc
    assert np.array_equal(X, X.T) or np.abs(X - X.T).max() <= 1e-13 * np.abs(X).max()
    assert np.abs(X - o["XTX"]).max() <= 1e-13 * np.abs(o["XTX"]).max()


def test_skinny_operand_digits_and_the_packed_product():
    """pg_i8.cu: out[s][l] = a_s sum_q 2^(e_l-7(q+1)) (V D^T)[s][(l,q)] + b_s sum_q ... (H D^T)[s][(l,q)] with D = signed 7-bit digits of
    the FP64 operand on one power-of-two scale per column; integer sums in Python ints."""
    rs = np.random.RandomState(5)
    nsnp, nind, L = 40, 37, 3
    g = synth.genotypes(8, nsnp, nind, missing=0.2)
    a = rs.randn(nsnp); b = rs.rand(nsnp) + 0.2
    B = rs.randn(nind, L) * 10.0 ** rs.uniform(-4, 4, size=(1, L))
    X = np.where(g >= 0, a[:, None] + b[:, None] * g, 0.0)
    want = X @ B
    got = np.zeros((nsnp, L))
    for l in range(L):
        e = int(np.frexp(np.abs(B[:, l]).max())[1])
        dig = [_digits(float(B[i, l]) / 2.0 ** e, 0) for i in range(nind)]         # |value| < 1 on the column's scale
        for s in range(nsnp):
            cv = [0] * NSL; ch = [0] * NSL
            for i in range(nind):
                if g[s, i] < 0:
                    continue
                for q in range(NSL):
                    cv[q] += dig[i][q]
                    ch[q] += int(g[s, i]) * dig[i][q]
            sv = sum(cv[q] * 2.0 ** (-7 * (q + 1)) for q in range(NSL))
            sh = sum(ch[q] * 2.0 ** (-7 * (q + 1)) for q in range(NSL))
            got[s, l] = (a[s] * sv + b[s] * sh) * 2.0 ** e
    scale = np.abs(want).max(axis=0, keepdims=True)
    assert (np.abs(got - want) / scale).max() <= 1e-13
