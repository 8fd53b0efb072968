import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from probit_b200 import linalg
def run(M, N, K, lower, ldpad=0, beta=1.0):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    def mat(r, c):
        buf = torch.empty((r, linalg.padded_ld(c) + ldpad), dtype=torch.float64, device="cuda")
        v = buf[:, :c]; v.copy_(torch.randn(r, c, dtype=torch.float64, device="cuda", generator=g)); return v
    A = mat(M, K); B = A if lower else mat(N, K); C0 = mat(M, N)
    ref = beta * C0 - A @ B.T
    out = C0.clone() if False else C0
    Cc = torch.empty_like(C0.contiguous()); 
    linalg.gemm_nt(A, B, out, alpha=-1.0, beta=beta, lower_only=lower)
    torch.cuda.synchronize()
    d = (out - ref)
    if lower: d = torch.tril(d)
    bad = (d.abs() > 1e-9).nonzero()
    print(f"M={M} N={N} K={K} lower={lower} ldC={out.stride(0)} beta={beta} maxerr={d.abs().max().item():.3e} nbad={bad.shape[0]}", 
          (bad[:3].tolist(), bad[-3:].tolist()) if bad.shape[0] else "", flush=True)
for args in [(4096,4096,512,False), (4096,4096,512,True), (3584,3584,512,True), (3584,3584,512,True,16), (2488,2488,512,True),
             (4096,4096,512,True,16), (4096,4096,64,True), (8192,8192,512,True), (8192, 512, 512, False), (16384, 256, 256, False), (4096,4096,512,True,0,0.0)]:
    run(*args)
