import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from probit_b200 import linalg
def run(M, N, K, lower, beta=1.0):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = linalg.empty_matrix(M, K); A.copy_(torch.randn(M, K, dtype=torch.float64, device="cuda", generator=g))
    C0 = linalg.empty_matrix(M, N); C0.copy_(torch.randn(M, N, dtype=torch.float64, device="cuda", generator=g))
    ref = beta * C0 - A[:N] @ A[:N].T if M == N else beta * C0 - A @ A[:N].T
    linalg.gemm_nt(A, A[:N], C0, alpha=-1.0, beta=beta, lower_only=lower)
    torch.cuda.synchronize()
    d = C0 - ref
    if lower: d = torch.tril(d)
    nbad = int((d.abs() > 1e-9).sum().item())
    print(f"dbg={os.environ.get('PB_GEMM_DBG','0')} M={M} N={N} K={K} lower={lower} beta={beta} nbad={nbad}", flush=True)
run(3584, 3584, 512, True); run(8192, 512, 512, False); run(3584, 3584, 512, True, 0.0); run(3584,3584,512,True,0.0)
