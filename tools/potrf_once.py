"""Run one potrf of size n (argv[1]) — target for ncu launch lists."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from probit_b200 import linalg
n = int(sys.argv[1]); reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
A = linalg.empty_matrix(n, n)
for r in range(reps):
    A.zero_(); A.diagonal().fill_(float(n)); A[:, 0].fill_(1.0); A[0, 0] = float(n)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fac = linalg.potrf_(A, check=False); e1.record(); e1.synchronize()
    print("potrf", n, "ms", e0.elapsed_time(e1), "TF", n**3/3/e0.elapsed_time(e1)*1e-9, "info", int(fac.info.item()))
