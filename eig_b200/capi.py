"""ctypes mirror of include/eigb200.h (the C-ABI of libeigb200.so).  No compute happens in Python.

The library is loaded lazily; if it is missing it is built with nvcc (eig_b200/build.py).  Every entry point
raises EigB200Error when the library reports an error -- in particular when no B200 is visible: there is no
CPU fallback.
"""
import ctypes as C
import os
import numpy as np

from . import build as _build

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libeigb200.so")


class EigB200Error(RuntimeError):
    pass


class GrmOpts(C.Structure):
    _fields_ = [("fancynorm", C.c_int), ("altnormstyle", C.c_int), ("minallelecnt", C.c_int), ("maxmissing", C.c_int),
                ("snp_ignore", C.c_void_p), ("snp_weight", C.c_void_p)]


class PcaOpts(C.Structure):
    _fields_ = [("grm", GrmOpts), ("numeigs", C.c_int), ("numoutliter", C.c_int), ("numoutleigs", C.c_int),
                ("outlthresh", C.c_double), ("outliermode", C.c_int)]


class PcaResult(C.Structure):
    _fields_ = [("nrows_final", C.c_int), ("niter", C.c_int), ("nremoved", C.c_int), ("nused", C.c_int64), ("y", C.c_double),
                ("secs_grm", C.c_double), ("secs_eig", C.c_double), ("secs_total", C.c_double)]


class Timings(C.Structure):
    _fields_ = [("gather_ms", C.c_float), ("stats_ms", C.c_float), ("grm_ms", C.c_float), ("finalize_ms", C.c_float),
                ("tridiag_ms", C.c_float), ("bisect_ms", C.c_float), ("vectors_ms", C.c_float),
                ("grm_launches", C.c_int), ("nsplit", C.c_int), ("eig_method", C.c_int), ("chfsi_iters", C.c_int),
                ("chfsi_matvecs", C.c_int), ("grm_sm_mhz", C.c_float), ("grm_cta_min_ms", C.c_float), ("grm_cta_max_ms", C.c_float),
                ("grm_span_ms", C.c_float), ("grm_sms", C.c_int), ("chfsi_converged", C.c_int), ("chfsi_resid", C.c_float),
                ("exchange_wait_ms", C.c_float), ("band_ms", C.c_float), ("chase_ms", C.c_float),
                ("grm_method", C.c_int), ("i8_slices", C.c_int), ("i8_segments", C.c_int), ("i8_flag_blocks", C.c_int),
                ("i8_tera_ops", C.c_float), ("i8_slab_rows", C.c_int), ("i8_gemm_ms", C.c_float)]


ALLGATHER_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64)
BARRIER_CB = C.CFUNCTYPE(C.c_int, C.c_void_p)


class Comm(C.Structure):
    """eb_comm: host-side plumbing of the SNP-sharded path (include/eigb200.h)."""
    _fields_ = [("rank", C.c_int), ("world", C.c_int), ("allgather_host", ALLGATHER_CB), ("barrier", BARRIER_CB), ("user", C.c_void_p)]


EXPORTS = [
    "eb_last_error", "eb_version", "eb_create", "eb_destroy", "eb_device_count", "eb_stream", "eb_sync", "eb_launch_count",
    "eb_reset_launch_count", "eb_upload_packed", "eb_upload_packed_rows", "eb_adopt_packed_device", "eb_synth_packed_device",
    "eb_set_rows", "eb_snp_counts", "eb_indiv_valid_counts", "eb_grm", "eb_grm_partial", "eb_grm_device_ptr", "eb_grm_finish",
    "eb_eig", "eb_eigvecs", "eb_ridoutlier", "eb_pca_full", "eb_fpca", "eb_gauss_matrix", "eb_project", "eb_get_timings",
    "eb_microbench_fp64", "eb_set_option", "eb_debug_tridiag", "eb_lsqproj", "eb_evec_coords", "eb_pop_counts", "eb_hash_ids", "eb_packed_file_header", "eb_upload_packed_file",
    "eb_download_packed", "eb_write_eval", "eb_write_evec", "eb_write_grm", "eb_grm_dense_begin", "eb_grm_dense_add", "eb_grm_dense_end", "eigvecs", "eigvals",
    "eb_set_comm", "eb_peer_allreduce_test", "eb_snp_used_count", "eb_shrink_coords", "eb_debug_gemm", "eb_local_comm_create", "eb_local_comm_get", "eb_local_comm_destroy", "eb_numgtz", "eb_tw_stats", "eb_tw_tail",
    "eb_setgval_packed", "eb_unsetgval", "kjg_fpca", "eb_write_grm_bin", "eb_grm_popfill",
]

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH) or (os.path.exists("/usr/local/cuda/bin/nvcc") and _build.needs_build()):
            _build.build()
        L = C.CDLL(LIB_PATH)
        L.eb_last_error.restype = C.c_char_p
        L.eb_create.restype = C.c_void_p
        L.eb_create.argtypes = [C.c_int]
        L.eb_destroy.argtypes = [C.c_void_p]
        L.eb_stream.restype = C.c_void_p
        L.eb_stream.argtypes = [C.c_void_p]
        L.eb_launch_count.restype = C.c_int64
        L.eb_launch_count.argtypes = [C.c_void_p]
        L.eb_grm_device_ptr.restype = C.c_void_p
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _chk(rc):
    if rc != 0:
        raise EigB200Error("libeigb200: %s (code %d)" % (lib().eb_last_error().decode(), rc))


def gauss_matrix(seed, n, L):
    out = np.empty((n, L))
    lib().eb_gauss_matrix(C.c_long(seed), C.c_size_t(n), C.c_size_t(L), _p(out))
    return out


def ridoutlier(evecs, neigs, thresh=6.0, outliermode=0):
    evecs = np.ascontiguousarray(evecs, np.float64); n = evecs.shape[1]
    bad = np.empty(n, np.int32); vecno = np.empty(n, np.int32); score = np.zeros(n)
    nb = lib().eb_ridoutlier(_p(evecs), C.c_int(n), C.c_int(neigs), C.c_double(thresh), C.c_int(outliermode), _p(bad), _p(vecno), _p(score))
    return bad[:nb].copy(), vecno, score


def _strarr(strings):
    arr = (C.c_char_p * len(strings))(*[s.encode() for s in strings])
    return arr


def hash_ids(ids):
    return lib().eb_hash_ids(_strarr(list(ids)), C.c_int(len(ids)))


def packed_file_header(path):
    ni = C.c_int(0); ns = C.c_int(0); ih = C.c_int(0); sh = C.c_int(0); rl = C.c_int64(0); fb = C.c_int64(0)
    _chk(lib().eb_packed_file_header(path.encode(), C.byref(ni), C.byref(ns), C.byref(ih), C.byref(sh), C.byref(rl), C.byref(fb)))
    return dict(nind=ni.value, nsnp=ns.value, ihash=ih.value, shash=sh.value, rlen=rl.value, file_bytes=fb.value)


def tw_stats(lam, znval=-1.0, minm=10):
    """Tracy-Widom statistic and effective n per eigenvalue (smartpca.c:1336-1366, dotwcalc statsubs.c:1680-1725)"""
    lam = np.ascontiguousarray(lam, np.float64)
    m = lib().eb_numgtz(_p(lam), C.c_int(len(lam)))
    tw = np.empty(m); zn = np.empty(m)
    _chk(lib().eb_tw_stats(_p(lam), C.c_int(m), C.c_double(znval), C.c_int(minm), _p(tw), _p(zn)))
    return tw, zn


def tw_tail(tw, table):
    """Tracy-Widom right tail ("p-value" column) from a POPGEN/twtable-style table [n][3] = x, tail, density"""
    t = np.ascontiguousarray(table, np.float64)
    x = np.ascontiguousarray(t[:, 0]); tl = np.ascontiguousarray(t[:, 1]); pd = np.ascontiguousarray(t[:, 2])
    lib().eb_tw_tail.restype = C.c_double
    return np.array([lib().eb_tw_tail(C.c_double(v), _p(x), _p(tl), _p(pd), C.c_int(len(x))) for v in np.atleast_1d(tw)])


def write_eval(path, lam):
    lam = np.ascontiguousarray(lam, np.float64)
    _chk(lib().eb_write_eval(path.encode(), _p(lam), C.c_int(len(lam))))


def write_evec(path, lam, ids, groups, coords, hiprec=False):
    coords = np.ascontiguousarray(coords, np.float64); k, n = coords.shape
    lam = np.ascontiguousarray(lam, np.float64)
    _chk(lib().eb_write_evec(path.encode(), _p(lam), C.c_int(k), _strarr(list(ids)), _strarr(list(groups)), _p(coords), C.c_int(n),
                             C.c_int(1 if hiprec else 0)))


def write_grm(path, xtx, numsnps):
    xtx = np.ascontiguousarray(xtx, np.float64)
    _chk(lib().eb_write_grm(path.encode(), _p(xtx), C.c_int(xtx.shape[0]), C.c_int(numsnps)))


def write_grm_bin(prefix, xtx, numsnps):
    xtx = np.ascontiguousarray(xtx, np.float64)
    _chk(lib().eb_write_grm_bin(prefix.encode(), _p(xtx), C.c_int(xtx.shape[0]), C.c_int(numsnps)))


class LocalComm:
    """eb_local_comm: in-process communicator for `world` host threads with one Context each (include/eigb200.h)."""

    def __init__(self, world):
        lib().eb_local_comm_create.restype = C.c_void_p
        self.h = lib().eb_local_comm_create(C.c_int(world))
        if not self.h:
            raise EigB200Error("libeigb200: %s" % lib().eb_last_error().decode())
        self.h = C.c_void_p(self.h); self.world = world

    def rank_comm(self, rank):
        st = Comm()
        lib().eb_local_comm_get.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        _chk(lib().eb_local_comm_get(self.h, C.c_int(rank), C.byref(st)))
        return st

    def close(self):
        if self.h:
            lib().eb_local_comm_destroy.argtypes = [C.c_void_p]
            lib().eb_local_comm_destroy(self.h); self.h = None


class Context:
    """One eb_ctx = one GPU."""

    def __init__(self, device=-1):
        self.h = lib().eb_create(device)
        if not self.h:
            raise EigB200Error("libeigb200: %s" % lib().eb_last_error().decode())
        self.h = C.c_void_p(self.h)
        self.nsnp = self.numindivs = self.nrows = 0
        self._keep = []

    def close(self):
        if self.h:
            lib().eb_destroy(self.h); self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- multi-GPU
    def set_comm(self, comm):
        """comm: an object with .rank, .world, .allgather_host(src_ptr, dst_ptr, nbytes) and .barrier() (parallel.TorchComm),
        or None for single-GPU operation."""
        if comm is None:
            _chk(lib().eb_set_comm(self.h, None)); self._comm = None
            return

        def _ag(user, src, dst, nbytes):
            try:
                comm.allgather_host(src, dst, nbytes); return 0
            except Exception as ex:  # surfaced by the library as EB_ERR_STATE
                import sys
                print("eb_comm.allgather_host: %r" % (ex,), file=sys.stderr); return 1

        def _bar(user):
            try:
                comm.barrier(); return 0
            except Exception as ex:
                import sys
                print("eb_comm.barrier: %r" % (ex,), file=sys.stderr); return 1

        st = Comm(comm.rank, comm.world, ALLGATHER_CB(_ag), BARRIER_CB(_bar), None)
        self._comm = (st, comm)          # keep the callbacks alive as long as the context uses them
        _chk(lib().eb_set_comm(self.h, C.byref(st)))

    def set_comm_struct(self, comm_struct, keep=None):
        """comm_struct: a filled capi.Comm (e.g. from LocalComm.rank_comm) whose callbacks are C functions"""
        self._comm = (comm_struct, keep)
        _chk(lib().eb_set_comm(self.h, C.byref(comm_struct)))

    def peer_allreduce_test(self, vec):
        v = np.ascontiguousarray(vec, np.float64).copy()
        _chk(lib().eb_peer_allreduce_test(self.h, _p(v), C.c_int64(v.size)))
        return v

    # ---- store
    def upload_packed(self, packed, numindivs):
        packed = np.ascontiguousarray(packed, np.uint8)
        self.nsnp, rlen = packed.shape
        self.numindivs = numindivs
        _chk(lib().eb_upload_packed(self.h, _p(packed), C.c_int64(self.nsnp), C.c_int64(rlen), C.c_int(numindivs)))
        _chk(lib().eb_sync(self.h))

    def upload_packed_file(self, path, numindivs, nsnp, ihash=None, shash=None, snp_is_x=None, indiv_is_male=None):
        chk = ihash is not None and shash is not None
        sx = None if snp_is_x is None else np.ascontiguousarray(snp_is_x, np.uint8)
        im = None if indiv_is_male is None else np.ascontiguousarray(indiv_is_male, np.uint8)
        _chk(lib().eb_upload_packed_file(self.h, path.encode(), C.c_int(numindivs), C.c_int64(nsnp), C.c_int(1 if chk else 0),
                                         C.c_int(ihash if chk else 0), C.c_int(shash if chk else 0), _p(sx), _p(im)))
        self.nsnp, self.numindivs = nsnp, numindivs

    def download_packed(self, rlen):
        out = np.empty((self.nsnp, rlen), np.uint8)
        _chk(lib().eb_download_packed(self.h, _p(out)))
        return out

    def adopt_packed_device(self, dev_ptr, nsnp, pitch, numindivs):
        self.nsnp, self.numindivs = nsnp, numindivs
        _chk(lib().eb_adopt_packed_device(self.h, C.c_void_p(dev_ptr), C.c_int64(nsnp), C.c_int64(pitch), C.c_int(numindivs)))

    def synth_packed_device(self, dev_ptr, nsnp, pitch, numindivs, seed, s0=0, missing=0.0, npops=1, delta=0.0):
        _chk(lib().eb_synth_packed_device(self.h, C.c_void_p(dev_ptr), C.c_int64(nsnp), C.c_int64(pitch), C.c_int(numindivs),
                                          C.c_uint64(seed), C.c_int64(s0), C.c_double(missing), C.c_int(npops), C.c_double(delta)))

    def set_rows(self, xindex=None):
        if xindex is None:
            self.nrows = self.numindivs
            _chk(lib().eb_set_rows(self.h, None, C.c_int(self.numindivs)))
        else:
            xi = np.ascontiguousarray(xindex, np.int32); self.nrows = len(xi)
            _chk(lib().eb_set_rows(self.h, _p(xi), C.c_int(len(xi))))

    # ---- integer reductions
    def snp_counts(self):
        c0 = np.empty(self.nsnp, np.int32); c1 = np.empty(self.nsnp, np.int32); nm = np.empty(self.nsnp, np.int32)
        _chk(lib().eb_snp_counts(self.h, _p(c0), _p(c1), _p(nm)))
        return c0, c1, nm

    def pop_counts(self, xtypes, npops):
        xt = np.ascontiguousarray(xtypes, np.int32)
        assert len(xt) == self.nrows
        out = np.empty((self.nsnp, npops, 3), np.int32)
        _chk(lib().eb_pop_counts(self.h, _p(xt), C.c_int(npops), _p(out)))
        return out

    def indiv_valid_counts(self, snp_keep=None):
        out = np.empty(self.numindivs, np.int32)
        keep = None if snp_keep is None else np.ascontiguousarray(snp_keep, np.uint8)
        _chk(lib().eb_indiv_valid_counts(self.h, _p(keep), _p(out)))
        return out

    # ---- GRM
    def _opts(self, fancynorm, altnormstyle, minallelecnt, maxmissing, snp_ignore, snp_weight):
        ig = None if snp_ignore is None else np.ascontiguousarray(snp_ignore, np.uint8)
        w = None if snp_weight is None else np.ascontiguousarray(snp_weight, np.float64)
        self._keep = [ig, w]
        return GrmOpts(fancynorm, altnormstyle, minallelecnt, maxmissing,
                       None if ig is None else ig.ctypes.data, None if w is None else w.ctypes.data)

    def grm(self, fancynorm=1, altnormstyle=1, minallelecnt=1, maxmissing=9999999, snp_ignore=None, snp_weight=None,
            want_xtx=False, want_snp=True, partial=False):
        o = self._opts(fancynorm, altnormstyle, minallelecnt, maxmissing, snp_ignore, snp_weight)
        m = self.nsnp
        r = {}
        if want_snp:
            r = dict(c0=np.empty(m, np.int32), c1=np.empty(m, np.int32), nmiss=np.empty(m, np.int32), used=np.empty(m, np.uint8),
                     xmean=np.empty(m), xfancy=np.empty(m))
        y = C.c_double(0); nused = C.c_int64(0)
        xtx = np.empty((self.nrows, self.nrows)) if want_xtx else None
        args = [_p(r.get(k)) for k in ("c0", "c1", "nmiss", "used", "xmean", "xfancy")]
        if partial:
            _chk(lib().eb_grm_partial(self.h, C.byref(o), *args, C.byref(nused)))
        else:
            _chk(lib().eb_grm(self.h, C.byref(o), *args, C.byref(y), C.byref(nused), _p(xtx)))
            r["y"] = y.value
            if want_xtx:
                r["XTX"] = xtx
        r["nused"] = nused.value
        return r

    def grm_popfill(self, xtypes, npops, fancynorm=1, altnormstyle=1, minallelecnt=1, maxmissing=9999999, snp_ignore=None, snp_weight=None,
                    want_xtx=False):
        """usepopsformissing: YES -- eb_grm_popfill (getcolxz with the population fill, dense path, all on the device)"""
        o = self._opts(fancynorm, altnormstyle, minallelecnt, maxmissing, snp_ignore, snp_weight)
        xt = np.ascontiguousarray(xtypes, np.int32)
        assert len(xt) == self.nrows
        m = self.nsnp
        r = dict(c0=np.empty(m, np.int32), c1=np.empty(m, np.int32), nmiss=np.empty(m, np.int32), used=np.empty(m, np.uint8),
                 xmean=np.empty(m), xfancy=np.empty(m))
        y = C.c_double(0); nused = C.c_int64(0)
        xtx = np.empty((self.nrows, self.nrows)) if want_xtx else None
        _chk(lib().eb_grm_popfill(self.h, C.byref(o), _p(xt), C.c_int(npops), *[_p(r[k]) for k in ("c0", "c1", "nmiss", "used", "xmean", "xfancy")],
                                  C.byref(y), C.byref(nused), _p(xtx)))
        r["y"] = y.value; r["nused"] = nused.value
        if want_xtx:
            r["XTX"] = xtx
        return r

    def snp_used_count(self):
        """SNPs of THIS context's shard that entered XTX in the last GRM pass"""
        lib().eb_snp_used_count.restype = C.c_int64
        return lib().eb_snp_used_count(self.h)

    def grm_device_ptr(self):
        ld = C.c_int64(0); n = C.c_int64(0)
        lib().eb_grm_device_ptr.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        p = lib().eb_grm_device_ptr(self.h, C.byref(ld), C.byref(n))
        return p, ld.value, n.value

    def grm_finish(self, want_xtx=False):
        y = C.c_double(0)
        xtx = np.empty((self.nrows, self.nrows)) if want_xtx else None
        _chk(lib().eb_grm_finish(self.h, C.byref(y), _p(xtx)))
        return y.value, xtx

    def grm_dense(self, tblocks, nrows, want_xtx=False):
        """dense path: tblocks = iterable of [nblock][nrows] FP64 blocks of normalised columns"""
        _chk(lib().eb_grm_dense_begin(self.h, C.c_int(nrows)))
        self.nrows = nrows
        for tb in tblocks:
            tb = np.ascontiguousarray(tb, np.float64)
            assert tb.shape[1] == nrows
            _chk(lib().eb_grm_dense_add(self.h, _p(tb), C.c_int(tb.shape[0])))
        y = C.c_double(0)
        xtx = np.empty((nrows, nrows)) if want_xtx else None
        _chk(lib().eb_grm_dense_end(self.h, C.byref(y), _p(xtx)))
        return y.value, xtx

    # ---- eigen
    def set_option(self, key, value):
        _chk(lib().eb_set_option(self.h, key.encode(), C.c_int(value)))

    def debug_tridiag(self, mat, want_band=True):
        mat = np.ascontiguousarray(mat, np.float64); n = mat.shape[0]
        d = np.empty(n); e = np.empty(n); band = np.empty((n, 128)) if want_band else None
        _chk(lib().eb_debug_tridiag(self.h, _p(mat), C.c_int(n), _p(d), _p(e), _p(band)))
        return d, e[:n - 1], band

    def eig(self, nvec, want_lambda=True):
        lam = np.empty(self.nrows) if want_lambda else None
        vec = np.empty((max(nvec, 1), self.nrows))
        _chk(lib().eb_eig(self.h, C.c_int(nvec), _p(lam), _p(vec)))
        return lam, vec[:nvec]

    def eigvecs(self, mat, nvec=None):
        mat = np.ascontiguousarray(mat, np.float64); n = mat.shape[0]
        nvec = n if nvec is None else nvec
        lam = np.empty(n); vec = np.empty((max(nvec, 1), n))
        _chk(lib().eb_eigvecs(self.h, _p(mat), _p(lam), _p(vec), C.c_int(n), C.c_int(nvec)))
        return lam, vec[:nvec]

    def pca_full(self, xindex=None, numeigs=10, numoutliter=5, numoutleigs=10, outlthresh=6.0, outliermode=0,
                 fancynorm=1, altnormstyle=1, minallelecnt=1, maxmissing=9999999, snp_ignore=None, snp_weight=None):
        xi = np.ascontiguousarray(np.arange(self.numindivs) if xindex is None else xindex, np.int32).copy()
        n0 = len(xi)
        o = PcaOpts(self._opts(fancynorm, altnormstyle, minallelecnt, maxmissing, snp_ignore, snp_weight),
                    numeigs, numoutliter, numoutleigs, outlthresh, outliermode)
        lam = np.empty(n0); vec = np.empty((max(numeigs, 1), n0))
        used = np.empty(self.nsnp, np.uint8); xmean = np.empty(self.nsnp); xfancy = np.empty(self.nsnp)
        ri = np.empty(n0, np.int32); rit = np.empty(n0, np.int32); rv = np.empty(n0, np.int32); rs = np.empty(n0)
        res = PcaResult()
        _chk(lib().eb_pca_full(self.h, C.byref(o), _p(xi), C.c_int(n0), _p(lam), _p(vec), _p(used), _p(xmean), _p(xfancy),
                               _p(ri), _p(rit), _p(rv), _p(rs), C.byref(res)))
        n = res.nrows_final; k = min(numeigs, n)
        self.nrows = n
        return dict(xindex=xi[:n].copy(), lambda_=lam[:n].copy(), evecs=vec.reshape(-1)[:k * n].reshape(k, n).copy(),
                    used=used, xmean=xmean, xfancy=xfancy, removed_index=ri[:res.nremoved].copy(),
                    removed_iter=rit[:res.nremoved].copy(), removed_vecno=rv[:res.nremoved].copy(),
                    removed_score=rs[:res.nremoved].copy(), niter=res.niter, nused=res.nused, y=res.y,
                    secs_grm=res.secs_grm, secs_eig=res.secs_eig, secs_total=res.secs_total)

    # ---- fastmode / projections
    def fpca(self, K, L, I, seed, fancynorm=1, altnormstyle=1):
        ev = np.empty(K); vec = np.empty((self.nrows, K))
        _chk(lib().eb_fpca(self.h, C.c_int(fancynorm), C.c_int(altnormstyle), C.c_size_t(K), C.c_size_t(L), C.c_size_t(I),
                           C.c_long(seed), _p(ev), _p(vec)))
        return ev, vec

    def project(self, evecs):
        evecs = np.ascontiguousarray(evecs, np.float64); k = evecs.shape[0]
        ff = np.empty((k, self.nsnp)); fx = np.empty((k, self.nrows)); sc = np.empty(k)
        _chk(lib().eb_project(self.h, _p(evecs), C.c_int(k), _p(ff), _p(fx), _p(sc)))
        return ff, fx, sc

    def lsqproj(self, ffvecs, fxscal, indiv=None):
        ffvecs = np.ascontiguousarray(ffvecs, np.float64); k = ffvecs.shape[0]
        fxscal = np.ascontiguousarray(fxscal, np.float64)
        lst = np.arange(self.numindivs, dtype=np.int32) if indiv is None else np.ascontiguousarray(indiv, np.int32)
        nl = len(lst)
        a = np.empty((k, nl)); b = np.empty((k, nl)); nv = np.empty(nl, np.int32); ok = np.empty(nl, np.uint8)
        _chk(lib().eb_lsqproj(self.h, _p(lst), C.c_int(nl), _p(ffvecs), _p(fxscal), C.c_int(k), _p(a), _p(b), _p(nv), _p(ok)))
        return a, b, nv, ok

    def evec_coords(self, evecs, indiv=None):
        evecs = np.ascontiguousarray(evecs, np.float64); k = evecs.shape[0]
        lst = np.arange(self.numindivs, dtype=np.int32) if indiv is None else np.ascontiguousarray(indiv, np.int32)
        nl = len(lst)
        co = np.empty((k, nl)); es = np.empty(k); ok = np.empty(nl, np.uint8)
        _chk(lib().eb_evec_coords(self.h, _p(evecs), C.c_int(k), _p(lst), C.c_int(nl), _p(co), _p(es), _p(ok)))
        return co, es, ok

    def shrink_coords(self, numeigs, newshrink=False):
        """shrinkmode coordinates [numeigs][numindivs] (as printevecs writes them), header eigenvalues, ok flags"""
        co = np.empty((numeigs, self.numindivs)); lam = np.empty(numeigs); ok = np.empty(self.numindivs, np.uint8)
        _chk(lib().eb_shrink_coords(self.h, C.c_int(numeigs), C.c_int(1 if newshrink else 0), _p(co), _p(lam), _p(ok)))
        return co, lam, ok

    def debug_gemm(self, A, B, a_km=False, b_kn=False):
        A = np.ascontiguousarray(A, np.float64); B = np.ascontiguousarray(B, np.float64)
        K, M = A.shape if a_km else A.shape[::-1]
        N = B.shape[1] if b_kn else B.shape[0]
        assert (B.shape[0] if b_kn else B.shape[1]) == K
        Cm = np.empty((M, N))
        _chk(lib().eb_debug_gemm(self.h, C.c_int(int(a_km)), C.c_int(int(b_kn)), _p(A), _p(B), _p(Cm), C.c_int(M), C.c_int(N), C.c_int(K)))
        return Cm

    # ---- measurement
    def timings(self):
        t = Timings()
        _chk(lib().eb_get_timings(self.h, C.byref(t)))
        return {k: getattr(t, k) for k, _ in Timings._fields_}

    def launch_count(self):
        return lib().eb_launch_count(self.h)

    def reset_launch_count(self):
        lib().eb_reset_launch_count(self.h)

    def stream(self):
        return lib().eb_stream(self.h)

    def sync(self):
        _chk(lib().eb_sync(self.h))

    def microbench_fp64(self):
        a = C.c_double(0); b = C.c_double(0)
        _chk(lib().eb_microbench_fp64(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value
