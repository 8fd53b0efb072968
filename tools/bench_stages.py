"""Per-stage roofline measurements on one B200: CUDA events, warm-up, best of 3.
Writes one JSON object; peaks: HBM from MEASURED_PEAKS.json, FP64 DMMA from profiles/r01_fp64_peaks.json."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from probit_b200 import _lib, linalg, kernels as PK, utilities as PU, approximators as PA

HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6545.3
DMMA = json.load(open(os.path.join(ROOT, "profiles", "r01_fp64_peaks.json")))["dmma_m16n8k8_tflops_w32"]


def best_ms(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


out = {"peaks": {"hbm_gbs": HBM, "fp64_dmma_tflops": DMMA}}
which = sys.argv[1:] or ["gram", "symv", "likelihood", "predictive", "predict", "c3"]

if "gram" in which:
    # write-only reference: torch fill_ over the same 32 GiB (what an ideal Gram kernel is bounded by)
    buf = torch.empty(65536 * 65536, dtype=torch.float64, device="cuda")
    ms = best_ms(lambda: buf.fill_(1.0))
    out["write_only_fill_32GiB"] = {"ms": ms, "GBs": buf.numel() * 8 / ms * 1e-6, "frac_hbm": buf.numel() * 8 / ms * 1e-6 / HBM}
    del buf
    torch.cuda.empty_cache()
    for n, D, fam in [(16384, 8, "eq"), (65536, 4, "matern12")]:
        X = torch.rand(n, D, dtype=torch.float64, device="cuda")
        spec = (1.0 * (PK.EQ() if fam == "eq" else PK.Matern12()).stretch(1.0)).lower()
        Z = linalg.features(spec, X)
        K = linalg.empty_matrix(n, n)
        lib = _lib.load()
        import ctypes as C
        f = lambda: lib.pb_gram_sym(linalg._stream(), C.byref(spec), linalg._ptr(Z), n, Z.shape[0], n, linalg._ptr(K), K.stride(0), None, 0.0)
        ms = best_ms(f)
        gbs = 8.0 * n * n / ms * 1e-6
        out[f"gram_{fam}_N{n}_D{D}"] = {"ms": ms, "GBs": gbs, "frac_hbm": gbs / HBM, "algorithmic_bytes": 8.0 * n * n}
        del K, X, Z
        torch.cuda.empty_cache()

if "symv" in which:
    n = 65536
    K = linalg.empty_matrix(n, n)
    K.normal_()
    x = torch.randn(n, dtype=torch.float64, device="cuda")
    y = torch.empty_like(x)
    lib = _lib.load()
    import ctypes as C
    ms = best_ms(lambda: lib.pb_symv(linalg._stream(), linalg._ptr(K), n, K.stride(0), linalg._ptr(x), linalg._ptr(y)))
    out["symv_full_65536"] = {"ms": ms, "GBs": n * n * 8 / ms * 1e-6, "frac_hbm": n * n * 8 / ms * 1e-6 / HBM}
    nbytes = lib.pb_symv_lower_scratch_bytes(n)
    scratch = torch.empty(nbytes // 8, dtype=torch.float64, device="cuda")
    ms = best_ms(lambda: lib.pb_symv_lower(linalg._stream(), linalg._ptr(K), n, K.stride(0), linalg._ptr(x), linalg._ptr(y), linalg._ptr(scratch), nbytes))
    alg = n * (n + 64) / 2 * 8
    out["symv_lower_65536"] = {"ms": ms, "algorithmic_bytes": alg, "GBs": alg / ms * 1e-6, "frac_hbm": alg / ms * 1e-6 / HBM,
                               "note": "algorithmic bytes = the tiles on and below the diagonal, read once"}
    del K, scratch
    torch.cuda.empty_cache()

if "likelihood" in which:
    n, batch, J = 65536, 1024, 5           # restarts x N = 2^26 elements
    cut = torch.tensor([-np.inf, -0.9, -0.2, 0.3, 1.0, np.inf], dtype=torch.float64)
    y = torch.randint(0, J, (n,), device="cuda")
    f = torch.randn(batch * n, dtype=torch.float64, device="cuda")
    lib = _lib.load()
    import ctypes as C
    spec, keep = PU.make_likelihood_spec(_lib.PB_LIK_ORDINAL_PROBIT, (0.63, cut))
    ll, g, h = (torch.empty_like(f) for _ in range(3))
    fn = lambda: lib.pb_likelihood(linalg._stream(), C.byref(spec), linalg._ptr(f), linalg._ptr(y), n, batch, linalg._ptr(ll), linalg._ptr(g), linalg._ptr(h), None)
    ms = best_ms(fn)
    byts = (8 + 24) * batch * n + 8 * n          # f in, ll/g/h out; y (int64) is read once per datum but re-used across the batch from L2
    out["likelihood_ordinal_2^26"] = {"ms": ms, "GBs": byts / ms * 1e-6, "frac_hbm": byts / ms * 1e-6 / HBM, "algorithmic_bytes": byts,
                                      "elements": batch * n}
    spec_g, _ = PU.make_likelihood_spec(_lib.PB_LIK_GAUSSIAN, (0.3,))
    yg = torch.randn(n, dtype=torch.float64, device="cuda")
    fn = lambda: lib.pb_likelihood(linalg._stream(), C.byref(spec_g), linalg._ptr(f), linalg._ptr(yg), n, batch, linalg._ptr(ll), linalg._ptr(g), linalg._ptr(h), None)
    ms = best_ms(fn)
    out["likelihood_gaussian_2^26"] = {"ms": ms, "GBs": byts / ms * 1e-6, "frac_hbm": byts / ms * 1e-6 / HBM}
    del f, ll, g, h
    torch.cuda.empty_cache()

if "predictive" in which:
    nt, J = 10_000_000, 5
    cut = torch.tensor([-np.inf, -0.9, -0.2, 0.3, 1.0, np.inf], dtype=torch.float64)
    m = torch.randn(nt, dtype=torch.float64, device="cuda"); v = torch.rand(nt, dtype=torch.float64, device="cuda") + 0.1
    ms = best_ms(lambda: PU.probit_predictive_distributions((0.63, cut), m, v))
    byts = (16 + 8 * J) * nt
    out["predictive_distributions_1e7xJ5"] = {"ms": ms, "GBs": byts / ms * 1e-6, "frac_hbm": byts / ms * 1e-6 / HBM, "algorithmic_bytes": byts}
    del m, v
    torch.cuda.empty_cache()

if "predict" in which:
    # predict at N=32768 (factor 8 GiB): mean+variance throughput; variance = N^2 * N_test flops (TRSM as DMMA GEMMs)
    n, nt, D = 32768, 8192, 4
    rng = np.random.default_rng(0)
    X = rng.uniform(size=(n, D)); y = rng.integers(0, 5, size=n)
    cut = np.array([-np.inf, -0.9, -0.2, 0.3, 1.0, np.inf])
    gp = PA.LaplaceGP((X, y), lambda l: 1.0 * PK.Matern12().stretch(l), PU.log_probit_likelihood)
    params = (1.0, (0.63, cut))
    w = torch.randn(n, dtype=torch.float64, device="cuda") * 0.01
    p = torch.rand(n, dtype=torch.float64, device="cuda") + 0.5
    Xs = torch.rand(nt, D, dtype=torch.float64, device="cuda")
    gp.predict(Xs[:64], params, w, p)          # prepare (Gram + potrf) once; cached afterwards
    ms = best_ms(lambda: gp.predict(Xs, params, w, p), reps=2)
    out["predict_N32768_Ntest8192"] = {"ms": ms, "variance_tflops": float(n) * n * nt / ms * 1e-9,
                                       "frac_dmma": float(n) * n * nt / ms * 1e-9 / DMMA, "test_points_per_s": nt / ms * 1e3}
    del gp
    torch.cuda.empty_cache()

if "predict_mean" in which:
    # BASELINE configs[4] (mean part): 10^7 test points against N=65536 training points, cross-covariance generated on
    # the fly inside the fused matvec kernel (6.6e11 kernel evaluations; nothing but X*, w and the mean touches HBM)
    n, nt, D = 65536, 10_000_000, 4
    rng = np.random.default_rng(0)
    X = rng.uniform(size=(n, D)); y = rng.integers(0, 5, size=n)
    cut = np.array([-np.inf, -0.9, -0.2, 0.3, 1.0, np.inf])
    gp = PA.LaplaceGP((X, y), lambda l: 1.0 * PK.Matern12().stretch(l), PU.log_probit_likelihood, predict_chunk=65536)
    params = (1.0, (0.63, cut))
    w = torch.randn(n, dtype=torch.float64, device="cuda") * 0.01
    Xs = torch.rand(nt, D, dtype=torch.float64, device="cuda") * 2 - 0.5
    ms = best_ms(lambda: gp.predict(Xs, params, w, None, variance=False), reps=2)
    out["predict_mean_N65536_Ntest1e7"] = {"ms": ms, "kernel_evals_per_s": float(n) * nt / ms * 1e3,
                                           "test_points_per_s": nt / ms * 1e3}
    del gp, Xs
    torch.cuda.empty_cache()

if "c3" in which:
    # BASELINE configs[2]: GP regression N=16384, D=8, EQ, FP64: Gram + Cholesky + evidence (LaplaceGP, Gaussian likelihood)
    n, D = 16384, 8
    rng = np.random.default_rng(0)
    X = rng.uniform(size=(n, D)); yv = np.sin(X.sum(1)) + 0.2 * rng.standard_normal(n)
    gp = PA.LaplaceGP((X, yv), lambda th: th[1] * PK.EQ().stretch(th[0]), PU.log_gaussian_likelihood)
    obj = gp.objective()
    params = ((1.0, 1.0), (0.2,))
    ms = best_ms(lambda: obj(params), reps=3)
    r = gp.last_result
    out["c3_regression_N16384_D8_EQ_objective"] = {"ms": ms, "newton_iterations": r.iterations, "factorizations": r.factorizations,
                                                   "cholesky_tflops_lower_bound": r.factorizations * n ** 3 / 3.0 / ms * 1e-9}
print(json.dumps(out, indent=1))
