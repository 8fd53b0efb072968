import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from probit_b200 import linalg, kernels as PK
for n in [int(a) for a in sys.argv[1:]]:
    torch.manual_seed(n)
    X = torch.rand(n, 4, dtype=torch.float64, device="cuda")
    K = linalg.gram((1.0 * PK.Matern12().stretch(1.0)).lower(), X, diag_add=1e-6)
    ref = torch.linalg.cholesky(K.contiguous())
    A = K.clone()
    fac = linalg.potrf_(A, check=False)
    L = torch.tril(A)
    err = (L - ref).abs().max().item()
    print(n, "info", int(fac.info.item()), "max abs err vs cusolver", err, "min diag ref", ref.diagonal().min().item(), flush=True)
