"""INT8-sliced (Ozaki) trailing update vs the FP64 DMMA kernel: SYRK rate, and potrf with potrf_ozaki on / off (time,
TFLOP/s-equivalent, factor difference).  One GPU.   usage: ozaki_bench.py [N ...]"""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from probit_b200 import _lib, linalg
lib = _lib.load()
sizes = [int(a) for a in sys.argv[1:]] or [16384, 32768]
out = {}

def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

for n, k in [(8192, 512), (16384, 512), (16384, 1024), (32768, 1024)]:
    A = linalg.empty_matrix(n, k); A.normal_()
    Cm = linalg.empty_matrix(n, n); Cm.zero_()
    nbytes = lib.pb_ozaki_scratch_bytes(n, n, k)
    scratch = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    oz = lambda: lib.pb_ozaki_gemm_nt(st, n, n, k, -1.0, C.c_void_p(A.data_ptr()), A.stride(0), C.c_void_p(A.data_ptr()), A.stride(0),
                                      C.c_void_p(Cm.data_ptr()), Cm.stride(0), 1, C.c_void_p(scratch.data_ptr()), nbytes)
    dm = lambda: linalg.gemm_nt(A, A, Cm, alpha=-1.0, beta=1.0, lower_only=True)
    t_oz, t_dm = timeit(oz), timeit(dm)
    fl = 1.0 * n * (n + 128) * k
    out[f"syrk_{n}x{k}"] = {"ozaki_ms": t_oz, "dmma_ms": t_dm, "ozaki_tflops": fl / t_oz * 1e-9, "dmma_tflops": fl / t_dm * 1e-9}
    print(f"syrk {n} x {k}: ozaki {t_oz:.2f} ms = {fl / t_oz * 1e-9:.1f} TF-equivalent, dmma {t_dm:.2f} ms = {fl / t_dm * 1e-9:.1f} TF", flush=True)
    del A, Cm, scratch
    torch.cuda.empty_cache()

for n in sizes:
    g = torch.Generator(device="cuda"); g.manual_seed(n)
    X = torch.rand(n, 4, generator=g, device="cuda", dtype=torch.float64)
    from probit_b200 import kernels as PK
    K0 = linalg.gram((1.0 * PK.Matern12().stretch(1.0)).lower(), X, diag_add=2.5)      # B-like: I + s s^T o K with s^2 ~ 2.5
    res = {}
    L = {}
    for mode in (0, 1):
        opt = _lib.default_options(potrf_ozaki=mode)
        A = linalg.empty_matrix(n, n)
        wsb = lib.pb_potrf_workspace_bytes(n); ws = torch.empty(wsb // 8, dtype=torch.float64, device="cuda"); info = torch.zeros(1, dtype=torch.int32, device="cuda")
        best = 1e30
        for r in range(3):
            A.copy_(K0); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            lib.pb_potrf(C.c_void_p(torch.cuda.current_stream().cuda_stream), C.c_void_p(A.data_ptr()), n, A.stride(0), C.c_void_p(ws.data_ptr()), wsb, C.c_void_p(info.data_ptr()), C.byref(opt))
            e1.record(); e1.synchronize()
            if r >= 1: best = min(best, e0.elapsed_time(e1))
        assert int(info.item()) == 0
        res[mode] = best
        L[mode] = torch.tril(A)
        del A, ws
    diff = ((L[1] - L[0]).norm() / L[0].norm()).item()
    mx = ((L[1] - L[0]).abs().max()).item()
    out[f"potrf_{n}"] = {"dmma_ms": res[0], "ozaki_ms": res[1], "dmma_tflops": n ** 3 / 3 / res[0] * 1e-9, "ozaki_tflops": n ** 3 / 3 / res[1] * 1e-9,
                         "factor_rel_diff": diff, "factor_max_abs_diff": mx}
    print(f"potrf {n}: dmma {res[0]:.1f} ms ({n ** 3 / 3 / res[0] * 1e-9:.1f} TF)  ozaki {res[1]:.1f} ms ({n ** 3 / 3 / res[1] * 1e-9:.1f} TF-eq)  factor rel diff {diff:.2e} max abs {mx:.2e}", flush=True)
    del L, K0
    torch.cuda.empty_cache()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/r02_ozaki_bench.json", "w"), indent=1)
