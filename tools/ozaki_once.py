"""One INT8-sliced SYRK (lower, n x K) for ncu:  ncu --set full -k regex:oz_gemm_kernel -c 1 python tools/ozaki_once.py [n K]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from probit_b200 import _lib, linalg
lib = _lib.load()
n, k = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (8192, 1024)
A = linalg.empty_matrix(n, k); A.normal_()
Cm = linalg.empty_matrix(n, n); Cm.zero_()
nbytes = lib.pb_ozaki_scratch_bytes(n, n, k)
scratch = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(2):
    lib.pb_ozaki_gemm_nt(st, n, n, k, -1.0, C.c_void_p(A.data_ptr()), A.stride(0), C.c_void_p(A.data_ptr()), A.stride(0),
                         C.c_void_p(Cm.data_ptr()), Cm.stride(0), 1, C.c_void_p(scratch.data_ptr()), nbytes)
torch.cuda.synchronize()
