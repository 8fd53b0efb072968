"""One launch of each HBM-bound kernel north_star names, sized past L2 but small enough for ncu's kernel replay
(it saves/restores the memory a kernel writes once per pass).  Run under

    ncu --set full --clock-control none --import-source on \
        -k regex:'likelihood_kernel|gram_sym_kernel|gram_matvec_kernel|predictive_kernel' -o gpurun_out/hbm python tools/ncu_hbm_kernels.py

Without ncu it prints CUDA-event timings of the same launches (warm, best of 3)."""
import ctypes as C
import json
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from probit_b200 import _lib, linalg, kernels as PK, utilities as PU, approximators as PA

lib = _lib.load()
timed = "--time" in sys.argv
out = {}


def run(name, fn, alg_bytes):
    if not timed:
        fn()
        torch.cuda.synchronize()
        return
    fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    out[name] = {"ms": best, "algorithmic_bytes": alg_bytes, "GBs": alg_bytes / best * 1e-6}


cut = torch.tensor([-np.inf, -0.9, -0.2, 0.3, 1.0, np.inf], dtype=torch.float64)
J = 5

# likelihood: 2^25 data (256 restarts x N=131072 shape; f 256 MiB in, ll/g/h 768 MiB out)
n, batch = 65536, 512
y = torch.randint(0, J, (n,), device="cuda")
f = torch.randn(batch * n, dtype=torch.float64, device="cuda")
spec, keep = PU.make_likelihood_spec(_lib.PB_LIK_ORDINAL_PROBIT, (0.63, cut))
ll, g, h = (torch.empty_like(f) for _ in range(3))
run("likelihood_ordinal_2^25", lambda: lib.pb_likelihood(linalg._stream(), C.byref(spec), linalg._ptr(f), linalg._ptr(y), n, batch,
                                                         linalg._ptr(ll), linalg._ptr(g), linalg._ptr(h), None), 32.0 * batch * n)
del f, ll, g, h
torch.cuda.empty_cache()

# predictive distributions: 10^7 test points, J=5
nt = 10_000_000
m = torch.randn(nt, dtype=torch.float64, device="cuda")
v = torch.rand(nt, dtype=torch.float64, device="cuda") + 0.1
run("predictive_1e7_J5", lambda: PU.probit_predictive_distributions((0.63, cut), m, v), (16.0 + 8 * J) * nt)
del m, v
torch.cuda.empty_cache()

# Gram: N=16384 (2 GiB written), Matern12 D=4 and EQ D=8
for nn, D, fam in [(16384, 4, "matern12"), (16384, 8, "eq")]:
    X = torch.rand(nn, D, dtype=torch.float64, device="cuda")
    kspec = (1.0 * (PK.EQ() if fam == "eq" else PK.Matern12()).stretch(1.0)).lower()
    Z = linalg.features(kspec, X)
    K = linalg.empty_matrix(nn, nn)
    run(f"gram_sym_{fam}_N{nn}_D{D}", lambda: lib.pb_gram_sym(linalg._stream(), C.byref(kspec), linalg._ptr(Z), nn, Z.shape[0], nn,
                                                              linalg._ptr(K), K.stride(0), None, 0.0), 8.0 * nn * nn)
    del K, X, Z
    torch.cuda.empty_cache()

# predict mean (gram_matvec_kernel): 65536 test points against N=65536, cross-covariance generated in registers
n, nt, D = 65536, 65536, 4
rng = np.random.default_rng(0)
X = rng.uniform(size=(n, D)); yv = rng.integers(0, J, size=n)
gp = PA.LaplaceGP((X, yv), lambda l: 1.0 * PK.Matern12().stretch(l), PU.log_probit_likelihood, predict_chunk=65536)
params = (1.0, (0.63, cut.numpy()))
w = torch.randn(n, dtype=torch.float64, device="cuda") * 0.01
Xs = torch.rand(nt, D, dtype=torch.float64, device="cuda") * 2 - 0.5
run("predict_mean_N65536_Ntest65536", lambda: gp.predict(Xs, params, w, None, variance=False), 8.0 * (n * D + nt * D + n + nt))
if timed:
    out["predict_mean_N65536_Ntest65536"]["kernel_evals_per_s"] = float(n) * nt / out["predict_mean_N65536_Ntest65536"]["ms"] * 1e3
    print(json.dumps(out, indent=1))
