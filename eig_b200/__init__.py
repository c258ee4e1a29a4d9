"""eig_b200 -- B200-native hot path for EIGENSOFT smartpca (GRM accumulation, eigendecomposition, fastmode PCA).

The product is the C-ABI shared library `eig_b200/libeigb200.so` (CUDA, sm_100a), declared in
`include/eigb200.h`.  This package holds its sources (`csrc/`), the build recipe (`build.py`), a thin
ctypes mirror of the C-ABI (`capi.py`), the smartpca-shaped host driver (`smartpca.py`) and the synthetic
genotype generator used by tests and bench (`synth.py`).  There is no CPU fallback: every compute entry
point raises if the CUDA library is missing.
"""
__version__ = "0.1.0"
