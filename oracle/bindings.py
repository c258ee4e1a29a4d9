"""ctypes bindings for the CHECKERS (test infrastructure only).

`port()`  -> oracle/liboracle.so   (plain-C restatement, eig_oracle.c)
`ref()`   -> oracle/_ref/libeigref.so (the unmodified reference behind ref_harness.c)

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) import this module.
The product package eig_b200 never does.
"""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


def build(ref=True):
    """(Re)build the checkers; the reference part is skipped when /root/reference is absent."""
    subprocess.run(["make", "-s", "-C", HERE, "port"] + (["ref"] if ref else []), check=True)


_port = None
_ref = None


def port():
    global _port
    if _port is None:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build(ref=False)
        _port = C.CDLL(path)
        _port.orc_mt_first.restype = C.c_uint32
        _port.orc_mt_first.argtypes = [C.c_ulong]
    return _port


def ref():
    """The compiled reference; None when it was never built (no /root/reference and no prebuilt copy)."""
    global _ref
    if _ref is None:
        path = os.path.join(HERE, "_ref", "libeigref.so")
        if not os.path.exists(path):
            if os.path.exists("/root/reference/src/eigensrc/smartpca.c"):
                build(ref=True)
            else:
                return None
        _ref = C.CDLL(path)
        _ref.refh_mt_first.restype = C.c_ulong
        _ref.refh_mt_first.argtypes = [C.c_ulong]
    return _ref


def ref_smartpca_binary():
    p = os.path.join(HERE, "_ref", "smartpca")
    return p if os.path.exists(p) else None


def _xi(xindex, nind):
    return np.ascontiguousarray(np.arange(nind) if xindex is None else xindex, dtype=np.int32)


# ----------------------------------------------------------------------------- port
def port_snp_counts(packed, xindex=None, numindivs=None):
    nsnp, rlen = packed.shape
    xi = _xi(xindex, numindivs)
    c0 = np.empty(nsnp, np.int32); c1 = np.empty(nsnp, np.int32); nm = np.empty(nsnp, np.int32)
    port().orc_snp_counts(packed.ctypes.data_as(C.c_void_p), C.c_long(nsnp), C.c_long(rlen),
                          xi.ctypes.data_as(C.c_void_p), C.c_int(len(xi)),
                          c0.ctypes.data_as(C.c_void_p), c1.ctypes.data_as(C.c_void_p), nm.ctypes.data_as(C.c_void_p))
    return c0, c1, nm


def port_indiv_valid_counts(packed, numindivs, snp_keep=None):
    nsnp, rlen = packed.shape
    out = np.empty(numindivs, np.int32)
    keep = None if snp_keep is None else np.ascontiguousarray(snp_keep, np.uint8)
    port().orc_indiv_valid_counts(packed.ctypes.data_as(C.c_void_p), C.c_long(nsnp), C.c_long(rlen), C.c_int(numindivs),
                                  keep.ctypes.data_as(C.c_void_p) if keep is not None else None,
                                  out.ctypes.data_as(C.c_void_p))
    return out


def _grm_call(fn, packed, numindivs, xindex, fancynorm, altnormstyle, minallelecnt, maxmissing, weights, extra):
    nsnp, rlen = packed.shape
    xi = _xi(xindex, numindivs)
    n = len(xi)
    r = dict(c0=np.empty(nsnp, np.int32), c1=np.empty(nsnp, np.int32), nmiss=np.empty(nsnp, np.int32),
             used=np.empty(nsnp, np.uint8), xmean=np.zeros(nsnp), xfancy=np.zeros(nsnp),
             XTX=np.zeros((n, n)), y=np.zeros(1))
    w = None if weights is None else np.ascontiguousarray(weights, np.float64)
    fn(packed, xi, n, fancynorm, altnormstyle, minallelecnt, maxmissing, w, r, extra)
    r["y"] = float(r["y"][0])
    return r


def port_grm(packed, numindivs, xindex=None, fancynorm=1, altnormstyle=1, minallelecnt=1, maxmissing=9999999, weights=None):
    def fn(packed, xi, n, fn_, an, mac, mm, w, r, _):
        nsnp, rlen = packed.shape
        port().orc_grm(packed.ctypes.data_as(C.c_void_p), C.c_long(nsnp), C.c_long(rlen), xi.ctypes.data_as(C.c_void_p), C.c_int(n),
                       C.c_int(fn_), C.c_int(an), C.c_int(mac), C.c_int(mm), w.ctypes.data_as(C.c_void_p) if w is not None else None,
                       *[r[k].ctypes.data_as(C.c_void_p) for k in ("c0", "c1", "nmiss", "used", "xmean", "xfancy", "XTX", "y")])
    return _grm_call(fn, packed, numindivs, xindex, fancynorm, altnormstyle, minallelecnt, maxmissing, weights, None)


def port_eigvecs(mat, want_vectors=True):
    mat = np.ascontiguousarray(mat, np.float64); n = mat.shape[0]
    ev = np.empty(n); vec = np.empty((n, n)) if want_vectors else None
    rc = port().orc_eigvecs(mat.ctypes.data_as(C.c_void_p), ev.ctypes.data_as(C.c_void_p),
                            vec.ctypes.data_as(C.c_void_p) if want_vectors else None, C.c_int(n))
    assert rc == 0
    return ev, vec


def port_ridoutlier(evecs, neigs, thresh=6.0, mode=0):
    evecs = np.ascontiguousarray(evecs, np.float64); n = evecs.shape[1]
    bad = np.empty(n, np.int32); vecno = np.empty(n, np.int32); score = np.zeros(n)
    nb = port().orc_ridoutlier(evecs.ctypes.data_as(C.c_void_p), C.c_int(n), C.c_int(neigs), C.c_double(thresh), C.c_int(mode),
                               bad.ctypes.data_as(C.c_void_p), vecno.ctypes.data_as(C.c_void_p), score.ctypes.data_as(C.c_void_p))
    return bad[:nb].copy(), vecno, score


def port_gauss(seed, n, L):
    out = np.empty((n, L))
    port().orc_gauss_matrix(C.c_long(seed), C.c_long(n), C.c_long(L), out.ctypes.data_as(C.c_void_p))
    return out


def port_gtable(packed, numindivs, xindex=None, fancynorm=1, altnormstyle=1):
    nsnp, rlen = packed.shape; xi = _xi(xindex, numindivs)
    gt = np.empty((nsnp, 4)); mono = np.empty(nsnp, np.uint8)
    port().orc_gtable(packed.ctypes.data_as(C.c_void_p), C.c_long(nsnp), C.c_long(rlen), xi.ctypes.data_as(C.c_void_p), C.c_int(len(xi)),
                      C.c_int(fancynorm), C.c_int(altnormstyle), gt.ctypes.data_as(C.c_void_p), mono.ctypes.data_as(C.c_void_p))
    return gt, mono


def port_fpca(packed, numindivs, K, L, I, seed, xindex=None, fancynorm=1, altnormstyle=1):
    nsnp, rlen = packed.shape; xi = _xi(xindex, numindivs); n = len(xi)
    ev = np.empty(K); vec = np.empty((n, K))
    rc = port().orc_fpca(packed.ctypes.data_as(C.c_void_p), C.c_long(nsnp), C.c_long(rlen), xi.ctypes.data_as(C.c_void_p), C.c_long(n),
                         C.c_int(fancynorm), C.c_int(altnormstyle), C.c_long(K), C.c_long(L), C.c_long(I), C.c_long(seed),
                         ev.ctypes.data_as(C.c_void_p), vec.ctypes.data_as(C.c_void_p))
    assert rc == 0
    return ev, vec


def port_project(packed, numindivs, used, xmean, xfancy, evecs, xindex=None):
    nsnp, rlen = packed.shape; xi = _xi(xindex, numindivs); n = len(xi)
    evecs = np.ascontiguousarray(evecs, np.float64); k = evecs.shape[0]
    ff = np.empty((k, nsnp)); fx = np.empty((k, n)); sc = np.empty(k)
    port().orc_project(packed.ctypes.data_as(C.c_void_p), C.c_long(nsnp), C.c_long(rlen), xi.ctypes.data_as(C.c_void_p), C.c_int(n),
                       np.ascontiguousarray(used, np.uint8).ctypes.data_as(C.c_void_p),
                       np.ascontiguousarray(xmean).ctypes.data_as(C.c_void_p), np.ascontiguousarray(xfancy).ctypes.data_as(C.c_void_p),
                       evecs.ctypes.data_as(C.c_void_p), C.c_int(k),
                       ff.ctypes.data_as(C.c_void_p), fx.ctypes.data_as(C.c_void_p), sc.ctypes.data_as(C.c_void_p))
    return ff, fx, sc


# ----------------------------------------------------------------------------- reference (oracle/_ref)
def ref_grm(packed, numindivs, xindex=None, fancynorm=1, altnormstyle=1, minallelecnt=1, maxmissing=9999999,
            weights=None, nthreads=None):
    secs = np.zeros(2)

    def fn(packed, xi, n, fn_, an, mac, mm, w, r, nthr):
        nsnp, rlen = packed.shape
        ref().refh_grm(packed.ctypes.data_as(C.c_void_p), C.c_long(nsnp), C.c_long(rlen), C.c_int(numindivs),
                       xi.ctypes.data_as(C.c_void_p), C.c_int(n), C.c_int(fn_), C.c_int(an), C.c_int(mac), C.c_int(mm),
                       w.ctypes.data_as(C.c_void_p) if w is not None else None, C.c_int(nthr),
                       *[r[k].ctypes.data_as(C.c_void_p) for k in ("c0", "c1", "nmiss", "used", "xmean", "xfancy", "XTX", "y")],
                       secs.ctypes.data_as(C.c_void_p))
    r = _grm_call(fn, packed, numindivs, xindex, fancynorm, altnormstyle, minallelecnt, maxmissing, weights,
                  nthreads or os.cpu_count())
    r["secs_loop"], r["secs_lookup"] = float(secs[0]), float(secs[1])
    return r


def ref_grm_loop(packed, numindivs, nthreads=None, tri=None):
    """Timing arm: the reference's per-SNP loop (smartpca.c:1116-1221) into a packed lower triangle, without symit2 / trace.
    Returns used flags, the loop's wall seconds and the triangle (reused between calls when given)."""
    nsnp, rlen = packed.shape; n = numindivs
    xi = np.arange(n, dtype=np.int32)
    if tri is None:
        tri = np.zeros(n * (n + 1) // 2)
    r = dict(c0=np.empty(nsnp, np.int32), c1=np.empty(nsnp, np.int32), nmiss=np.empty(nsnp, np.int32),
             used=np.empty(nsnp, np.uint8), xmean=np.zeros(nsnp), xfancy=np.zeros(nsnp))
    secs = np.zeros(2)
    ref().refh_grm_loop(packed.ctypes.data_as(C.c_void_p), C.c_long(nsnp), C.c_long(rlen), C.c_int(n), xi.ctypes.data_as(C.c_void_p),
                        C.c_int(n), C.c_int(1), C.c_int(1), C.c_int(1), C.c_int(9999999), C.c_int(nthreads or os.cpu_count()),
                        *[r[k].ctypes.data_as(C.c_void_p) for k in ("c0", "c1", "nmiss", "used", "xmean", "xfancy")],
                        tri.ctypes.data_as(C.c_void_p), secs.ctypes.data_as(C.c_void_p))
    r["secs_loop"], r["secs_lookup"], r["tri"] = float(secs[0]), float(secs[1]), tri
    return r


def ref_eigvecs(mat):
    mat = np.ascontiguousarray(mat, np.float64).copy(); n = mat.shape[0]
    ev = np.empty(n); vec = np.empty((n, n))
    ref().refh_eigvecs(mat.ctypes.data_as(C.c_void_p), ev.ctypes.data_as(C.c_void_p), vec.ctypes.data_as(C.c_void_p), C.c_int(n))
    return ev, vec


def ref_ridoutlier(evecs, neigs, thresh=6.0, mode=0):
    evecs = np.ascontiguousarray(evecs, np.float64).copy(); n = evecs.shape[1]
    bad = np.empty(n, np.int32); vecno = np.empty(n, np.int32); score = np.zeros(n)
    nb = ref().refh_ridoutlier(evecs.ctypes.data_as(C.c_void_p), C.c_int(n), C.c_int(neigs), C.c_double(thresh), C.c_int(mode),
                               bad.ctypes.data_as(C.c_void_p), vecno.ctypes.data_as(C.c_void_p), score.ctypes.data_as(C.c_void_p))
    return bad[:nb].copy(), vecno, score


def ref_gauss(seed, n, L):
    out = np.empty((n, L))
    ref().refh_gauss(C.c_long(seed), C.c_long(n), C.c_long(L), out.ctypes.data_as(C.c_void_p))
    return out


def ref_fpca(packed, numindivs, K, L, I, seed, xindex=None, fancynorm=1, altnormstyle=1):
    nsnp, rlen = packed.shape; xi = _xi(xindex, numindivs); n = len(xi)
    ev = np.empty(K); vec = np.empty((n, K)); secs = np.zeros(1)
    ref().refh_fpca(packed.ctypes.data_as(C.c_void_p), C.c_long(nsnp), C.c_long(rlen), C.c_int(numindivs),
                    xi.ctypes.data_as(C.c_void_p), C.c_int(n), C.c_int(fancynorm), C.c_int(altnormstyle),
                    C.c_long(K), C.c_long(L), C.c_long(I), C.c_long(seed),
                    ev.ctypes.data_as(C.c_void_p), vec.ctypes.data_as(C.c_void_p), secs.ctypes.data_as(C.c_void_p))
    return ev, vec, float(secs[0])


def port_lsqproj(packed, indiv, used, xmean, xfancy, ffvecs, fxscal):
    """oracle port of lsqproj (smartpca.c:4606-4757): acoeffs, bcoeffs [k][nlist], nvalid, ok"""
    nsnp, rlen = packed.shape
    indiv = np.ascontiguousarray(indiv, np.int32); nl = len(indiv)
    ffvecs = np.ascontiguousarray(ffvecs, np.float64); k = ffvecs.shape[0]
    a = np.empty((k, nl)); b = np.empty((k, nl)); nv = np.empty(nl, np.int32); ok = np.empty(nl, np.uint8)
    port().orc_lsqproj(packed.ctypes.data_as(C.c_void_p), C.c_long(nsnp), C.c_long(rlen), indiv.ctypes.data_as(C.c_void_p), C.c_int(nl),
                       np.ascontiguousarray(used, np.uint8).ctypes.data_as(C.c_void_p), np.ascontiguousarray(xmean).ctypes.data_as(C.c_void_p),
                       np.ascontiguousarray(xfancy).ctypes.data_as(C.c_void_p), ffvecs.ctypes.data_as(C.c_void_p),
                       np.ascontiguousarray(fxscal).ctypes.data_as(C.c_void_p), C.c_int(k),
                       a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), nv.ctypes.data_as(C.c_void_p), ok.ctypes.data_as(C.c_void_p))
    return a, b, nv, ok


def port_seteigscale(acoeffs, bcoeffs, rowpos):
    acoeffs = np.ascontiguousarray(acoeffs); bcoeffs = np.ascontiguousarray(bcoeffs); k, nl = acoeffs.shape
    rowpos = np.ascontiguousarray(rowpos, np.int32); out = np.empty(k)
    port().orc_seteigscale(acoeffs.ctypes.data_as(C.c_void_p), bcoeffs.ctypes.data_as(C.c_void_p), C.c_int(nl),
                           rowpos.ctypes.data_as(C.c_void_p), C.c_int(len(rowpos)), C.c_int(k), out.ctypes.data_as(C.c_void_p))
    return out


def port_evec_coords(packed, numindivs, used, xmean, xfancy, evecs, xindex=None, indiv_ignore=None):
    """the whole .evec value pipeline with the port: project -> lsqproj over all non-ignored individuals -> eigscale.
    Returns coords [k][numindivs] (zero rows for ignored / insufficient individuals), eigscale, ok[numindivs]"""
    xi = _xi(xindex, numindivs)
    ff, fx, sc = port_project(packed, numindivs, used, xmean, xfancy, evecs, xindex=xi)
    keep = np.ones(numindivs, bool) if indiv_ignore is None else ~np.asarray(indiv_ignore, bool)
    lst = np.flatnonzero(keep).astype(np.int32)
    a, b, nv, ok = port_lsqproj(packed, lst, used, xmean, xfancy, ff, sc)
    pos = -np.ones(numindivs, np.int64); pos[lst] = np.arange(len(lst))
    es = port_seteigscale(a, b, pos[xi])
    k = a.shape[0]
    coords = np.zeros((k, numindivs)); coords[:, lst] = a * es[:, None]
    okf = np.zeros(numindivs, np.uint8); okf[lst] = ok
    return coords, es, okf, ff, sc


def ref_evec_coords(packed, numindivs, used, xmean, xfancy, evecs, xindex=None, indiv_ignore=None, fancynorm=1, altnormstyle=1):
    """the unmodified reference's own post-eigen sequence (smartpca.c:1440-1564), run in a forked child"""
    nsnp, rlen = packed.shape; xi = _xi(xindex, numindivs); n = len(xi)
    evecs = np.ascontiguousarray(evecs, np.float64); k = evecs.shape[0]
    a = np.zeros((k, numindivs)); b = np.zeros((k, numindivs)); es = np.zeros(k); ff = np.zeros((k, nsnp)); sc = np.zeros(k)
    ign = np.zeros(numindivs, np.uint8)
    ig_in = None if indiv_ignore is None else np.ascontiguousarray(indiv_ignore, np.uint8)
    rc = ref().refh_evec_coords(packed.ctypes.data_as(C.c_void_p), C.c_long(nsnp), C.c_long(rlen), C.c_int(numindivs),
                                xi.ctypes.data_as(C.c_void_p), C.c_int(n), C.c_int(fancynorm), C.c_int(altnormstyle),
                                np.ascontiguousarray(used, np.uint8).ctypes.data_as(C.c_void_p),
                                np.ascontiguousarray(xmean).ctypes.data_as(C.c_void_p), np.ascontiguousarray(xfancy).ctypes.data_as(C.c_void_p),
                                evecs.ctypes.data_as(C.c_void_p), C.c_int(k), None if ig_in is None else ig_in.ctypes.data_as(C.c_void_p),
                                a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), es.ctypes.data_as(C.c_void_p),
                                ff.ctypes.data_as(C.c_void_p), sc.ctypes.data_as(C.c_void_p), ign.ctypes.data_as(C.c_void_p))
    assert rc == 0, rc
    return dict(coords=a, bcoeffs=b, eigscale=es, ffvecs=ff, fxscal=sc, ignored=ign)


def port_dense_grm(tblock):
    tblock = np.ascontiguousarray(tblock, np.float64); ncols, nrows = tblock.shape
    out = np.empty((nrows, nrows))
    port().orc_dense_grm(tblock.ctypes.data_as(C.c_void_p), C.c_long(ncols), C.c_int(nrows), out.ctypes.data_as(C.c_void_p))
    return out


def ref_dense_grm(tblock, blocksize=1024, nthreads=4):
    tblock = np.ascontiguousarray(tblock, np.float64); ncols, nrows = tblock.shape
    out = np.zeros((nrows, nrows))
    rc = ref().refh_dense_grm(tblock.ctypes.data_as(C.c_void_p), C.c_long(ncols), C.c_int(nrows), C.c_int(blocksize), C.c_int(nthreads),
                              out.ctypes.data_as(C.c_void_p))
    assert rc == 0
    return out


def ref_popfill_cols(packed, numindivs, xtypes, npops, xindex=None, fancynorm=1, altnormstyle=1):
    """usepopsformissing: the reference's getcolxz for every SNP -> columns [nsnp][nrows], c0, c1, nmiss (after the fill), xmean, xfancy"""
    nsnp, rlen = packed.shape; xi = _xi(xindex, numindivs); n = len(xi)
    xt = np.ascontiguousarray(xtypes, np.int32)
    r = dict(c0=np.empty(nsnp, np.int32), c1=np.empty(nsnp, np.int32), nmiss=np.empty(nsnp, np.int32), xmean=np.zeros(nsnp), xfancy=np.zeros(nsnp),
             cols=np.zeros((nsnp, n)))
    rc = ref().refh_popfill_cols(packed.ctypes.data_as(C.c_void_p), C.c_long(nsnp), C.c_long(rlen), C.c_int(numindivs), xi.ctypes.data_as(C.c_void_p),
                                 xt.ctypes.data_as(C.c_void_p), C.c_int(n), C.c_int(npops), C.c_int(fancynorm), C.c_int(altnormstyle),
                                 *[r[k].ctypes.data_as(C.c_void_p) for k in ("c0", "c1", "nmiss", "xmean", "xfancy", "cols")])
    assert rc == 0
    return r


def port_pop_counts(packed, numindivs, xtypes, npops, xindex=None):
    nsnp, rlen = packed.shape; xi = _xi(xindex, numindivs); xt = np.ascontiguousarray(xtypes, np.int32)
    out = np.empty((nsnp, npops, 3), np.int32)
    port().orc_pop_counts(packed.ctypes.data_as(C.c_void_p), C.c_long(nsnp), C.c_long(rlen), xi.ctypes.data_as(C.c_void_p),
                          xt.ctypes.data_as(C.c_void_p), C.c_int(len(xi)), C.c_int(npops), out.ctypes.data_as(C.c_void_p))
    return out


def port_fstcol(counts):
    """Fst numerator / denominator matrices per SNP from class counts [nsnp][numeg][3] (qpsubs.c:1303-1340)"""
    counts = np.ascontiguousarray(counts, np.int32); nsnp, numeg, _ = counts.shape
    en = np.empty((nsnp, numeg, numeg)); ed = np.empty((nsnp, numeg, numeg))
    for s in range(nsnp):
        port().orc_fstcol(counts[s].ctypes.data_as(C.c_void_p), C.c_int(numeg), en[s].ctypes.data_as(C.c_void_p), ed[s].ctypes.data_as(C.c_void_p))
    return en, ed


def ref_fstcol(packed, numindivs, xtypes, numeg, xindex=None):
    nsnp, rlen = packed.shape; xi = _xi(xindex, numindivs); xt = np.ascontiguousarray(xtypes, np.int32)
    en = np.empty((nsnp, numeg, numeg)); ed = np.empty((nsnp, numeg, numeg))
    rc = ref().refh_fstcol(packed.ctypes.data_as(C.c_void_p), C.c_long(nsnp), C.c_long(rlen), C.c_int(numindivs),
                           xi.ctypes.data_as(C.c_void_p), xt.ctypes.data_as(C.c_void_p), C.c_int(len(xi)), C.c_int(numeg),
                           en.ctypes.data_as(C.c_void_p), ed.ctypes.data_as(C.c_void_p))
    assert rc == 0
    return en, ed


def ref_shrink(packed, numindivs, used, xmean, xfancy, XTXn, k, xindex=None, newshrink=False, fancynorm=1, altnormstyle=1):
    """the unmodified reference's doshrinkp / doshrinkp2 (smartpca.c:4223-4419 / 4022-4220), run in a forked child;
    returns the shrinkmode .evec values [k][numindivs]"""
    nsnp, rlen = packed.shape; xi = _xi(xindex, numindivs); n = len(xi)
    X = np.ascontiguousarray(XTXn, np.float64); assert X.shape == (n, n)
    out = np.zeros((k, numindivs))
    rc = ref().refh_shrink(packed.ctypes.data_as(C.c_void_p), C.c_long(nsnp), C.c_long(rlen), C.c_int(numindivs),
                           xi.ctypes.data_as(C.c_void_p), C.c_int(n), C.c_int(fancynorm), C.c_int(altnormstyle),
                           np.ascontiguousarray(used, np.uint8).ctypes.data_as(C.c_void_p),
                           np.ascontiguousarray(xmean).ctypes.data_as(C.c_void_p), np.ascontiguousarray(xfancy).ctypes.data_as(C.c_void_p),
                           X.ctypes.data_as(C.c_void_p), C.c_int(k), C.c_int(1 if newshrink else 0), out.ctypes.data_as(C.c_void_p))
    assert rc == 0, rc
    return out


def port_shrink(packed, numindivs, used, xmean, xfancy, XTXn, k, xindex=None, newshrink=False):
    """numpy restatement of doshrinkp (smartpca.c:4223-4419) and doshrinkp2 (4022-4220) -- small cases only (O(k m^3 + k m^2 n)).
    Follows the reference step by step: trace normalisation 4285-4290, eigvecs 4292, norme + loadings 4305-4311, old-style
    projection doproj 4313-4318 (3986-4019: least squares on the observed SNPs, x = g*xfancy - xmean via fixxrow qpsubs.c:338),
    then per (eigenvector i, sample a): dd / ww / delta 4332-4339, first-order perturbation 4340-4347, ymul 4348-4357
    (old: lam > delta; new: lam > -delta, 4143), zero / centre / norme 4359-4364, loadings 4366-4368, re-projection
    4369-4374 (old: only row i replaced; new 4165-4169: all rows replaced, one regression per sample), printevecs 3849-3866."""
    from eig_b200 import synth          # unpack helper only (2-bit layout, admutils.c:718-735)
    nsnp, rlen = packed.shape; xi = _xi(xindex, numindivs); m = len(xi)
    used = np.asarray(used).astype(bool)
    g = synth.unpack(packed, numindivs).astype(np.float64).T        # [numindivs][nsnp], -1 missing
    # mmat = getcolxf columns (smartpca.c:3564-3597): (g - mean) * yfancy over the PCA rows, 0 where missing / ignored SNP
    gp = g[xi]
    valid = gp >= 0
    cnt = np.maximum(valid.sum(0), 1)
    ymean = np.where(valid, gp, 0).sum(0) / cnt
    mmat = np.where(valid, gp - ymean, 0.0) * np.asarray(xfancy)[None, :]
    mmat[:, ~used] = 0.0
    n = nsnp
    xmat = np.array(XTXn, np.float64)
    xmat = xmat * (1.0 / (np.trace(xmat) / (m - 1)))
    lam, vec = np.linalg.eigh(xmat)
    lam = lam[::-1].copy(); evecs = vec[:, ::-1].T.copy()            # row i = eigenvector i, descending (eigsubs.c:39-55)

    def norme(v):
        v = v - v.sum() / len(v)
        return v / np.sqrt((v * v).sum())

    def doproj(isample, fx):
        row = g[isample]
        ok = (row >= 0) & used
        x = row * xfancy - xmean
        e = fx[:, ok].T
        co = e.T @ e; rr = e.T @ x[ok]
        return np.linalg.solve(co, rr)

    ffvecs = np.zeros((k, n))
    for i in range(k):
        evecs[i] = norme(evecs[i])
        ff = evecs[i] @ mmat
        ffvecs[i] = ff / np.sqrt((ff * ff).sum() / n)
    ss = np.zeros((k, numindivs))
    for i in range(numindivs):
        ss[:, i] = doproj(i, ffvecs)
    snew = np.zeros((k, m))

    def enew_of(i, a):
        evec = evecs[i]; l = lam[i]
        ww = -xmat[a] * evec[a]
        ww[a] = -(xmat[a] @ evec)
        delta = ww @ evec
        eco = evecs @ ww
        den = l - lam
        den[i] = 1.0
        eco = eco / den
        eco[i] = 0.0
        ediff = eco @ evecs
        good = (l > -delta) if newshrink else (l > delta)
        ymul = l / (l + delta) if good else 1.0
        en = evec + ediff
        en[a] = 0
        en = en - en.sum() / (m - 1)
        en[a] = 0
        en = norme(en)
        ff = en @ mmat
        return ff / np.sqrt((ff * ff).sum() / n), ymul

    if newshrink:
        for a in range(m):
            fx = ffvecs.copy(); ym = np.ones(k)
            for i in range(k):
                fx[i], ym[i] = enew_of(i, a)
            snew[:, a] = doproj(xi[a], fx) * ym
    else:
        for i in range(k):
            for a in range(m):
                fx = ffvecs.copy()
                fx[i], ym = enew_of(i, a)
                snew[i, a] = doproj(xi[a], fx)[i] * ym
    ss[:, xi] = snew
    out = 10.0 * ss
    out /= np.sqrt((out * out).sum(1))[:, None]
    return out, lam[:k]
