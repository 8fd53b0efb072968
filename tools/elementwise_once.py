"""Runs the Gram, likelihood and predictive-distribution kernels once each at roofline-relevant sizes (ncu target)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, ctypes as C
from probit_b200 import _lib, linalg, kernels as PK, utilities as PU
lib = _lib.load()
n = 32768
X = torch.rand(n, 4, dtype=torch.float64, device="cuda")
spec = (1.0 * PK.Matern12().stretch(1.0)).lower()
for _ in range(2):
    K = linalg.gram(spec, X)
del K
cut = torch.tensor([-np.inf, -0.9, -0.2, 0.3, 1.0, np.inf], dtype=torch.float64)
nn, batch = 65536, 256
y = torch.randint(0, 5, (nn,), device="cuda")
f = torch.randn(batch * nn, dtype=torch.float64, device="cuda")
for _ in range(2):
    out = PU.evaluate_likelihood(_lib.PB_LIK_ORDINAL_PROBIT, f, y, (0.63, cut), ("ll", "g", "h"))
m = torch.randn(4_000_000, dtype=torch.float64, device="cuda"); v = torch.rand(4_000_000, dtype=torch.float64, device="cuda") + 0.1
for _ in range(2):
    P = PU.probit_predictive_distributions((0.63, cut), m, v)
torch.cuda.synchronize()
print("done")
