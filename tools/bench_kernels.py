"""Micro-benchmarks of the individual CUDA kernels (CUDA events, warm-up, best of reps).  Not the driver bench."""
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from probit_b200 import linalg, kernels as PK, _lib
import ctypes as C


def timeit(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


out = {}
which = sys.argv[1:] or ["gemm", "potrf", "gram", "blas2"]
if "gemm" in which:
    for n, k in [(8192, 8192), (8192, 512), (16384, 512), (32768, 512), (16384, 1024), (16384, 256)]:
        A = linalg.empty_matrix(n, k); A.normal_()
        Cm = linalg.empty_matrix(n, n); Cm.zero_()
        ms = timeit(lambda: linalg.gemm_nt(A, A, Cm, alpha=-1.0, beta=1.0))
        out[f"gemm_nt_{n}x{n}x{k}_tflops"] = 2.0 * n * n * k / ms * 1e-9
        ms = timeit(lambda: linalg.gemm_nt(A, A, Cm, alpha=-1.0, beta=1.0, lower_only=True))
        out[f"syrk_lower_{n}x{k}_tflops"] = 1.0 * n * (n + 128) * k / ms * 1e-9
        del A, Cm
if "potrf" in which:
    for n in [4096, 8192, 16384, 32768]:
        A = linalg.empty_matrix(n, n)
        def reset():
            A.zero_(); A.diagonal().fill_(float(n)); A[:, 0].fill_(1.0); A[0, 0] = float(n)
        lib = _lib.load()
        wsb = lib.pb_potrf_workspace_bytes(n)
        ws = torch.empty(wsb // 8, dtype=torch.float64, device="cuda")
        info = torch.zeros(1, dtype=torch.int32, device="cuda")
        best = 1e30
        for r in range(3):
            reset(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            lib.pb_potrf(C.c_void_p(torch.cuda.current_stream().cuda_stream), C.c_void_p(A.data_ptr()), n, A.stride(0),
                         C.c_void_p(ws.data_ptr()), wsb, C.c_void_p(info.data_ptr()), None)
            e1.record(); e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        assert int(info.item()) == 0
        out[f"potrf_{n}_ms"] = best
        out[f"potrf_{n}_tflops"] = n ** 3 / 3.0 / best * 1e-9
        del A, ws
if "gram" in which:
    for n, D, fam in [(16384, 8, "eq"), (32768, 4, "matern12")]:
        X = torch.rand(n, D, dtype=torch.float64, device="cuda")
        k = (1.0 * (PK.EQ() if fam == "eq" else PK.Matern12()).stretch(1.0)).lower()
        ms = timeit(lambda: linalg.gram(k, X))
        out[f"gram_{fam}_{n}_D{D}_ms"] = ms
        out[f"gram_{fam}_{n}_D{D}_GBs"] = 8.0 * n * n / ms * 1e-6
if "blas2" in which or "blas2_64k" in which:
    n = 65536 if "blas2_64k" in which else 32768
    A = linalg.empty_matrix(n, n); A.normal_()
    x = torch.randn(n, dtype=torch.float64, device="cuda")
    ms = timeit(lambda: linalg.symv(A, x))
    out[f"symv_{n}_ms"] = ms
    out[f"symv_{n}_GBs"] = 8.0 * n * n / ms * 1e-6
    A.zero_(); A.diagonal().fill_(2.0)
    fac = linalg.potrf_(A)
    ms = timeit(lambda: linalg.trsv(fac, x))
    out[f"trsv_fwd_{n}_ms"] = ms
    ms = timeit(lambda: linalg.trsv(fac, x, True))
    out[f"trsv_bwd_{n}_ms"] = ms
print(json.dumps(out, indent=1))
