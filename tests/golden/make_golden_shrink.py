"""Generate tests/golden/ref_shrink.npz by running the UNMODIFIED reference's doshrinkp / doshrinkp2 (through
oracle/_ref/libeigref.so: refh_shrink) on seeded synthetic input.  Run in the build container:
  python tests/golden/make_golden_shrink.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from eig_b200 import synth
from oracle import bindings as ob

seed, nsnp, nind, missing, k = 29, 900, 70, 0.15, 3
P = synth.packed_genotypes(seed, nsnp, nind, missing=missing, npops=4, delta=0.35)
xi = np.arange(3, nind - 4, dtype=np.int32)
r = ob.ref_grm(P, nind, xindex=xi, nthreads=2)
X = r["XTX"] / r["y"]
old = ob.ref_shrink(P, nind, r["used"], r["xmean"], r["xfancy"], X, k, xindex=xi, newshrink=False)
new = ob.ref_shrink(P, nind, r["used"], r["xmean"], r["xfancy"], X, k, xindex=xi, newshrink=True)
np.savez_compressed(os.path.join(os.path.dirname(__file__), "ref_shrink.npz"), seed=seed, nsnp=nsnp, nind=nind, missing=missing, k=k,
                    xindex=xi, shrink_old=old, shrink_new=new)
print("written", old.shape, np.abs(old - new).max())
