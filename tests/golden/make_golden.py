"""Generate tests/golden/ref_vectors.npz by running the UNMODIFIED reference (oracle/_ref/libeigref.so, built from
/root/reference by oracle/Makefile) on seeded synthetic input.  Run in the build container:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from eig_b200 import synth
from oracle import bindings as ob

seed, nsnp, nind, missing, alt, gseed = 17, 600, 90, 0.12, 0, 424242
P = synth.packed_genotypes(seed, nsnp, nind, missing=missing, npops=3, delta=0.3)
xi = np.sort(np.random.RandomState(0).choice(nind, 75, replace=False)).astype(np.int32)
r = ob.ref_grm(P, nind, xindex=xi, altnormstyle=alt, nthreads=2)
lam, vec = ob.ref_eigvecs(r["XTX"] / r["y"])
ev, u, _ = ob.ref_fpca(P, nind, K=3, L=6, I=2, seed=gseed, xindex=xi, altnormstyle=alt)
np.savez_compressed(os.path.join(os.path.dirname(__file__), "ref_vectors.npz"), seed=seed, nsnp=nsnp, nind=nind, missing=missing,
                    altnormstyle=alt, xindex=xi, c0=r["c0"], c1=r["c1"], nmiss=r["nmiss"], used=r["used"], xmean=r["xmean"],
                    xfancy=r["xfancy"], y=r["y"], lambda_=lam, evecs=vec[:4], gseed=gseed, gauss=ob.ref_gauss(gseed, 64, 6),
                    fpca_eval=ev, fpca_evec=u)
print("written")
