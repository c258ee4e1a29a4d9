import sys
import numpy as np
sys.path.insert(0, ".")
from eig_b200 import capi
c = capi.Context(0)
n = int(sys.argv[1]); tail = int(sys.argv[2]) if len(sys.argv) > 2 else n // 10
rs = np.random.RandomState(n)
m = n - tail
X = rs.randn(n, m) + 0.5 * rs.randn(3, m)[rs.randint(0, 3, n)]
A = X @ X.T / m
c.set_option("eig_method", 1)
lam, vec = c.eigvecs(A)
w = np.linalg.eigvalsh(A)[::-1]
G = np.abs(vec @ vec.T - np.eye(n))
i, j = np.unravel_index(np.argmax(G), G.shape)
print("max orth err", G.max(), "at", i, j, "lam", lam[i], lam[j], "tn", w[0])
R = np.abs(vec @ A - lam[:, None] * vec).max(axis=1)
print("worst residual", R.max(), "at", R.argmax(), "lam there", lam[R.argmax()])
rowmax = G.max(axis=1)
bad = np.flatnonzero(rowmax > 1e-9)
print("rows with orth err > 1e-9:", len(bad), bad[:10], bad[-10:])
gaps = -np.diff(lam)
print("min gap outside tail", gaps[:n - tail - 1].min() / w[0], "eig err", np.abs(lam - w).max() / w[0])
