"""One SYRK-shaped GEMM (n x n x k, lower) — target for ncu captures."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from probit_b200 import linalg
n = int(sys.argv[1]); k = int(sys.argv[2]); reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
A = linalg.empty_matrix(n, k); A.normal_()
Cm = linalg.empty_matrix(n, n); Cm.zero_()
for r in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); linalg.gemm_nt(A, A, Cm, alpha=-1.0, beta=1.0, lower_only=True); e1.record(); e1.synchronize()
    print("syrk", n, k, "ms", e0.elapsed_time(e1), "TF", n * (n + 128) * k / e0.elapsed_time(e1) * 1e-9)
