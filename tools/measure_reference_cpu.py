"""CPU probe (run on the GPU box's host): the UNMODIFIED reference (oracle/_ref) timed stage by stage at sizes it finishes
in seconds, next to the same stage on the B200 through the C-ABI -- the 'reference CPU path timed beside it' for the stages
other than the GRM (which bench.py times).  Test/measurement infrastructure only.
usage: python tools/measure_reference_cpu.py [--out gpurun_out/reference_cpu.json]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from eig_b200 import capi, synth
from oracle import bindings as ob


def t(fn):
    t0 = time.perf_counter(); r = fn(); return time.perf_counter() - t0, r


def main():
    out = {"host_cpus": os.cpu_count(), "note": "reference = oracle/_ref (unmodified EIGENSOFT sources, OpenBLAS 0.3.15), all host threads; b200 = libeigb200 through the C-ABI, wall clock incl. transfers"}
    ob.ref().refh_openblas_threads(os.cpu_count())
    ctx = capi.Context(0)
    # eigvecs (dspev) vs eb_eigvecs (all eigenvalues + all vectors, as dspev returns them)
    rows = []
    for n in (1000, 2000, 3000):
        rs = np.random.RandomState(n); X = rs.randn(n, 2 * n); A = X @ X.T / (2 * n)
        tr, (lr, vr) = t(lambda: ob.ref_eigvecs(A))
        ctx.eigvecs(A[:64, :64])
        tg, (lg, vg) = t(lambda: ctx.eigvecs(A))
        tl, _ = t(lambda: ctx.eigvecs(A, nvec=10))
        rows.append(dict(n=n, reference_dspev_s=tr, b200_all_vectors_s=tg, b200_spectrum_plus_10_vectors_s=tl,
                         max_rel_eval_diff=float(np.abs(lr - lg).max() / lr[0])))
    out["eigvecs"] = rows
    # fastmode
    rows = []
    for nsnp, nind in ((20000, 2000), (60000, 4000)):
        P = synth.packed_genotypes(3, nsnp, nind, missing=0.02, npops=4, delta=0.1)
        tr, (er, ur, _) = t(lambda: ob.ref_fpca(P, nind, K=10, L=20, I=10, seed=7))
        ctx.upload_packed(P, nind); ctx.set_rows(None)
        tg, (eg, ug) = t(lambda: ctx.fpca(10, 20, 10, seed=7))
        rows.append(dict(nsnp=nsnp, nind=nind, reference_kjg_fpca_s=tr, b200_s=tg, max_rel_eval_diff=float(np.abs(er - eg).max() / er[0])))
    out["fastmode"] = rows
    # .evec coordinates (loadings + projections + lsqproj) and shrinkmode
    nsnp, nind = 20000, 1000
    P = synth.packed_genotypes(5, nsnp, nind, missing=0.1, npops=4, delta=0.15)
    ctx.upload_packed(P, nind); ctx.set_rows(None)
    r = ctx.grm(want_xtx=True); lam, vec = ctx.eig(10)
    tr, rr = t(lambda: ob.ref_evec_coords(P, nind, r["used"], r["xmean"], r["xfancy"], vec))
    tg, (co, es, ok) = t(lambda: ctx.evec_coords(vec))
    sg = np.sign((rr["coords"] * co).sum(1))
    out["evec_coords"] = dict(nsnp=nsnp, nind=nind, k=10, reference_s=tr, b200_s=tg, max_abs_diff=float(np.abs(rr["coords"] - co * sg[:, None]).max()))
    nsnp, nind = 6000, 400
    P = synth.packed_genotypes(6, nsnp, nind, missing=0.1, npops=4, delta=0.2)
    ctx.upload_packed(P, nind); ctx.set_rows(None)
    r = ctx.grm(want_xtx=True)
    tr, want = t(lambda: ob.ref_shrink(P, nind, r["used"], r["xmean"], r["xfancy"], r["XTX"], 3))
    tg, (got, sl, sok) = t(lambda: ctx.shrink_coords(3))
    sg = np.sign((got * want).sum(1))
    out["shrinkmode"] = dict(nsnp=nsnp, nind=nind, k=3, reference_doshrinkp_s=tr, b200_s=tg, max_abs_diff=float(np.abs(got * sg[:, None] - want).max()),
                             note="reference cost grows as k m (m^2 + m n): x16 per doubling of m at fixed n/m")
    path = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else "gpurun_out/reference_cpu.json"
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
