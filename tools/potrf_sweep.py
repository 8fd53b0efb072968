"""potrf time vs panel width: python tools/potrf_sweep.py N[,N...] NB [NB ...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from probit_b200 import _lib, linalg
for nb in sys.argv[2:]:
    opt = _lib.default_options(potrf_block=int(nb))
    for n in [int(a) for a in sys.argv[1].split(",")]:
        A = linalg.empty_matrix(n, n)
        best = 1e30
        for r in range(2):
            A.zero_(); A.diagonal().fill_(float(n)); A[:, 0].fill_(1.0); A[0, 0] = float(n)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fac = linalg.potrf_(A, check=False, options=opt); e1.record(); e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print(f"NB {nb} potrf {n} ms {best:.2f} TF {n**3/3/best*1e-9:.2f} info {int(fac.info.item())}", flush=True)
        del A
        torch.cuda.empty_cache()
