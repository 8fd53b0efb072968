"""Timeline of ONE CTA of oz_gemm_kernel (clock64 stations, csrc/ozaki.cu oz_dbg):  PB_OZ_TIMING=<block> python tools/oz_timeline.py n K"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from probit_b200 import _lib, linalg
lib = C.CDLL(_lib.LIB_PATH)
L = _lib.load()
n, k = int(sys.argv[1]), int(sys.argv[2])
A = linalg.empty_matrix(n, k); A.normal_()
Cm = linalg.empty_matrix(n, n); Cm.zero_()
nbytes = L.pb_ozaki_scratch_bytes(n, n, k)
scratch = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(3):
    L.pb_ozaki_gemm_nt(st, n, n, k, -1.0, C.c_void_p(A.data_ptr()), A.stride(0), C.c_void_p(A.data_ptr()), A.stride(0),
                       C.c_void_p(Cm.data_ptr()), Cm.stride(0), 1, C.c_void_p(scratch.data_ptr()), nbytes)
torch.cuda.synchronize()
buf = (C.c_ulonglong * 16)()
lib.pb_debug_oz_times(buf)
names = ["entry", "setup", "first_full", "last_mma_issued", "tfull", "drained", "c_done", "exit", "first_tma"]
t0 = buf[0]
print(f"n={n} K={k} block={os.environ.get('PB_OZ_TIMING')} red={os.environ.get('PB_OZ_RED')}: " +
      ", ".join(f"{nm}={buf[i] - t0}" for i, nm in enumerate(names)))
