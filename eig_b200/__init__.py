"""eig_b200 -- B200-native hot path for EIGENSOFT smartpca (GRM accumulation, eigendecomposition, fastmode PCA).

The product is the C-ABI shared library `eig_b200/libeigb200.so` (CUDA, sm_100a), declared in
`include/eigb200.h`.  This package holds its sources (`csrc/`), the build recipe (`build.py`), a thin
ctypes mirror of the C-ABI (`capi.py`, used by tests and bench), the torch.distributed plumbing for the `eb_comm`
callbacks (`parallel.py`) and the synthetic genotype generator (`synth.py`).  The smartpca-shaped host side is the reference's
own smartpca.c with `integration/smartpca_b200.patch` applied.  There is no CPU fallback: every compute entry
point raises if the CUDA library is missing.
"""
__version__ = "0.1.0"
