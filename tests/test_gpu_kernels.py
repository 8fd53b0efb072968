"""GPU parity tests of the individual CUDA kernels, all called through the C ABI (ctypes)."""
import numpy as np
import pytest
import scipy.linalg as sla

pytestmark = pytest.mark.gpu

from helpers import relerr  # noqa: E402
from oracle import kernels as OK, utilities as OU  # noqa: E402


def _torch():
    import torch
    return torch


@pytest.mark.parametrize("M,N,K", [(128, 128, 16), (256, 128, 64), (300, 200, 100), (64, 64, 64), (1, 1, 1),
                                   (1000, 40, 40), (513, 257, 129), (2048, 1024, 512)])
def test_gemm_nt_matches_fp64_matmul(M, N, K):
    torch = _torch()
    from probit_b200 import linalg
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    A = linalg.empty_matrix(M, K); A.copy_(torch.randn(M, K, dtype=torch.float64, device="cuda", generator=g))
    B = linalg.empty_matrix(N, K); B.copy_(torch.randn(N, K, dtype=torch.float64, device="cuda", generator=g))
    C0 = linalg.empty_matrix(M, N); C0.copy_(torch.randn(M, N, dtype=torch.float64, device="cuda", generator=g))
    ref = 0.5 * (A @ B.T) - 2.0 * C0
    out = linalg.gemm_nt(A, B, C0.clone() if False else C0, alpha=0.5, beta=-2.0)
    torch.cuda.synchronize()
    err = (out - ref).abs().max().item() / max(ref.abs().max().item(), 1e-300)
    assert err < 1e-13, err


def test_gemm_nt_lower_only_leaves_upper_untouched():
    torch = _torch()
    from probit_b200 import linalg
    n, k = 700, 96
    g = torch.Generator(device="cuda").manual_seed(1)
    A = linalg.empty_matrix(n, k); A.copy_(torch.randn(n, k, dtype=torch.float64, device="cuda", generator=g))
    C = linalg.empty_matrix(n, n); C.fill_(7.0)
    linalg.gemm_nt(A, A, C, alpha=-1.0, beta=1.0, lower_only=True)
    ref = 7.0 - A @ A.T
    assert torch.allclose(torch.tril(C), torch.tril(ref), rtol=1e-13, atol=1e-12)
    assert torch.all(torch.triu(C, 1) == torch.triu(torch.full_like(C, 7.0), 1))


@pytest.mark.parametrize("n", [1, 5, 31, 32, 33, 64, 65, 100, 128, 200, 513, 1000, 2500])
def test_potrf_matches_lapack(n):
    torch = _torch()
    from probit_b200 import linalg
    rng = np.random.default_rng(n)
    G = rng.standard_normal((n, n + 5))
    A = G @ G.T + n * 0.01 * np.eye(n)
    Ad = linalg.empty_matrix(n, n); Ad.copy_(torch.as_tensor(A, device="cuda"))
    Ad_upper_before = torch.triu(Ad, 1).clone()
    fac = linalg.potrf_(Ad)
    L = np.tril(Ad.cpu().numpy())
    Lref = np.linalg.cholesky(A)
    assert relerr(L, Lref) < 1e-12
    assert torch.equal(torch.triu(Ad, 1), Ad_upper_before)       # strict upper never written
    # solves through the leaf inverses
    b = rng.standard_normal(n)
    x = linalg.cholesky_solve(fac, torch.as_tensor(b, device="cuda")).cpu().numpy()
    assert relerr(x, np.linalg.solve(A, b)) < 1e-9
    ld = linalg.logdet_chol(fac).item()
    assert abs(ld - np.sum(np.log(np.diag(Lref)))) < 1e-10 * max(1.0, abs(ld))
    # many-RHS right solve X L^-T
    Xr = rng.standard_normal((37, n))
    Xd = linalg.empty_matrix(37, n); Xd.copy_(torch.as_tensor(Xr, device="cuda"))
    linalg.trsm_right_lt_(fac, Xd)
    ref = sla.solve_triangular(Lref, Xr.T, lower=True).T
    assert relerr(Xd.cpu().numpy(), ref) < 1e-10


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (128, 64, 128), (200, 150, 192), (1000, 333, 512), (4096, 4096, 1024)])
def test_ozaki_int8_gemm_matches_fp64_product(M, N, K):
    """csrc/ozaki.cu: C += alpha A B^T through 7 int8 digit planes on tcgen05 (UTCIMMA).  Entries span 12 orders of
    magnitude across rows (each row has its own exponent) and 3 within a row; the error bound is 2^-49 of the row scales
    times sqrt(K) (asserted with a factor-of-32 margin for the worst entry; observed 8 x at K = 512)."""
    torch = _torch()
    from probit_b200 import linalg
    g = torch.Generator(device="cuda"); g.manual_seed(M + N + K)
    A = linalg.empty_matrix(M, K); B = linalg.empty_matrix(N, K); Cm = linalg.empty_matrix(M, N)
    A.copy_(torch.randn(M, K, generator=g, device="cuda", dtype=torch.float64) * torch.exp(3 * torch.randn(M, K, generator=g, device="cuda", dtype=torch.float64).clamp(-1, 1))
            * (10.0 ** torch.randint(-6, 7, (M, 1), generator=g, device="cuda").double()))
    B.copy_(torch.randn(N, K, generator=g, device="cuda", dtype=torch.float64) * (10.0 ** torch.randint(-3, 4, (N, 1), generator=g, device="cuda").double()))
    A[M // 2].zero_()                                        # an all-zero row
    C0 = torch.randn(M, N, generator=g, device="cuda", dtype=torch.float64)
    Cm.copy_(C0)
    linalg.ozaki_gemm_nt(A, B, Cm, alpha=-0.75)
    ref = C0 - 0.75 * (A @ B.T)
    scale = A.abs().amax(dim=1, keepdim=True) * B.abs().amax(dim=1, keepdim=True).T * (K ** 0.5) + C0.abs()
    err = ((Cm - ref).abs() / scale).max().item()
    assert err < 32 * 2.0 ** -49, err
    # SYRK form: only tiles on / below the diagonal are touched
    if M == N:
        Cs = linalg.empty_matrix(M, M); Cs.copy_(C0)
        linalg.ozaki_gemm_nt(A, A, Cs, alpha=1.0, lower_only=True)
        ref = C0 + A @ A.T
        sc = A.abs().amax(dim=1, keepdim=True) * A.abs().amax(dim=1, keepdim=True).T * (K ** 0.5) + C0.abs()
        low = torch.tril(torch.ones(M, M, device="cuda", dtype=torch.bool))
        assert (((Cs - ref).abs() / sc)[low]).max().item() < 32 * 2.0 ** -49
        i = torch.arange(M, device="cuda")
        untouched = (i[None, :] // 64) > (i[:, None] // 128) * 2 + 1          # column tiles right of the row tile's diagonal
        assert torch.equal(Cs[untouched], C0[untouched])


@pytest.mark.parametrize("n,nb,tile", [(6000, 0, 0), (9000, 1024, 0), (5200, 768, 0), (6000, 0, 1), (9000, 1024, 2), (5200, 768, 2)])
def test_potrf_with_int8_trailing_updates_matches_lapack(n, nb, tile):
    """potrf_ozaki = 1: every panel is sliced once into int8 digit planes and both the look-ahead column update and the
    trailing SYRK multiply those planes on tcgen05 (ozaki.cu, potrf.cu); blocks smaller than 2048 rows stay on DMMA.
    Held to LAPACK like the FP64 path; strict upper triangle untouched; ragged n (not a multiple of the panel width)."""
    torch = _torch()
    from probit_b200 import linalg, _lib
    rng = np.random.default_rng(n)
    G = rng.standard_normal((n, 64))
    A = G @ G.T / 64 + np.diag(rng.uniform(1.0, 3.0, n))            # B-like: identity + low-rank-ish PSD part
    Ad = linalg.empty_matrix(n, n); Ad.copy_(torch.as_tensor(A, device="cuda"))
    upper = torch.triu(Ad, 1).clone()
    # tile: kernel variant (pb_options.ozaki_tile) — 0 = 128x64 one pass, 1 = 128x128 two passes, 2 = 2-CTA clusters with multicast
    fac = linalg.potrf_(Ad, options=_lib.default_options(potrf_ozaki=1, potrf_block=nb, ozaki_tile=tile))
    L = np.tril(Ad.cpu().numpy())
    Lref = np.linalg.cholesky(A)
    assert relerr(L, Lref) < 1e-12
    assert torch.equal(torch.triu(Ad, 1), upper)
    b = rng.standard_normal(n)
    x = linalg.cholesky_solve(fac, torch.as_tensor(b, device="cuda")).cpu().numpy()
    assert relerr(x, np.linalg.solve(A, b)) < 1e-10


def test_int8_right_trsm_and_default_policy_n8192():
    """Default options switch the K >= 1024 contractions of potrf and of the many-RHS right solve to the INT8 path from
    n = 8192 on (potrf_ozaki = -1).  Factor and X L^-T against LAPACK, and against the FP64-only path (potrf_ozaki = 0)."""
    torch = _torch()
    from probit_b200 import linalg, _lib
    n = 8192
    rng = np.random.default_rng(8)
    G = rng.standard_normal((n, 96))
    A = G @ G.T / 96 + np.diag(rng.uniform(1.0, 2.0, n))
    Ad = linalg.empty_matrix(n, n); Ad.copy_(torch.as_tensor(A, device="cuda"))
    fac = linalg.potrf_(Ad)                                           # library defaults
    Lref = np.linalg.cholesky(A)
    assert relerr(np.tril(Ad.cpu().numpy()), Lref) < 1e-12
    Xr = rng.standard_normal((600, n)) * 10.0 ** rng.integers(-3, 4, size=(600, 1))
    Xd = linalg.empty_matrix(600, n); Xd.copy_(torch.as_tensor(Xr, device="cuda"))
    linalg.trsm_right_lt_(fac, Xd)
    ref = sla.solve_triangular(Lref, Xr.T, lower=True).T
    err = np.abs(Xd.cpu().numpy() - ref).max(axis=1) / np.abs(ref).max(axis=1)
    assert err.max() < 1e-11, err.max()
    Bd = linalg.empty_matrix(n, n); Bd.copy_(torch.as_tensor(A, device="cuda"))
    linalg.potrf_(Bd, options=_lib.default_options(potrf_ozaki=0))
    assert relerr(np.tril(Bd.cpu().numpy()), np.tril(Ad.cpu().numpy())) < 1e-12


@pytest.mark.parametrize("n", [700, 3000, 5000])
def test_potrf_graph_replay_is_bitwise_identical_to_eager(n):
    """pb_options.potrf_graph: the first call with a given set of buffers runs eagerly, the second captures the two-stream
    launch DAG into a CUDA graph, later calls replay it.  All of them must produce the same bits as the eager path
    (same kernels, same order), report failures through `info` the same way, and count the same launches."""
    torch = _torch()
    import ctypes as C
    from probit_b200 import linalg, _lib
    lib = _lib.load()
    rng = np.random.default_rng(n)
    G = rng.standard_normal((n, 40))
    A = torch.as_tensor(G @ G.T + np.eye(n), device="cuda")
    Ad = linalg.empty_matrix(n, n)
    wsb = lib.pb_potrf_workspace_bytes(n)
    ws = torch.empty(wsb // 8, dtype=torch.float64, device="cuda")
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def run(options):
        Ad.copy_(A)
        ws.zero_()
        n0 = lib.pb_launch_count()
        _lib.check(lib.pb_potrf(st, C.c_void_p(Ad.data_ptr()), n, Ad.stride(0), C.c_void_p(ws.data_ptr()), wsb,
                                C.c_void_p(info.data_ptr()), C.byref(options)))
        torch.cuda.synchronize()
        return torch.tril(Ad).clone(), ws.clone(), int(info.item()), lib.pb_launch_count() - n0

    eager = run(_lib.default_options(potrf_graph=0))
    graph_opt = _lib.default_options(potrf_graph=1)
    for call in range(4):                                   # eager (first sighting), capture + replay, replay, replay
        L, w, i, launches = run(graph_opt)
        assert i == 0 and torch.equal(L, eager[0]) and torch.equal(w, eager[1]), call
        assert launches == eager[3], (call, launches, eager[3])
    assert relerr(eager[0].cpu().numpy(), np.linalg.cholesky(A.cpu().numpy())) < 1e-12
    # a replayed graph still reports a non-SPD matrix
    A[n // 2, n // 2] = -1.0
    L, w, i, _ = run(graph_opt)
    assert i == n // 2 + 1


def test_potrf_reports_first_bad_pivot():
    torch = _torch()
    from probit_b200 import linalg, _lib
    n = 300
    A = np.eye(n); A[150, 150] = -1.0
    Ad = linalg.empty_matrix(n, n); Ad.copy_(torch.as_tensor(A, device="cuda"))
    fac = linalg.potrf_(Ad, check=False)
    assert int(fac.info.item()) == 151
    Ad.copy_(torch.as_tensor(A, device="cuda"))
    with pytest.raises(_lib.NumericError):
        linalg.potrf_(Ad)


@pytest.mark.parametrize("family,D", [("eq", 1), ("eq", 8), ("matern12", 4), ("eq_periodic", 1), ("eq_periodic", 3)])
@pytest.mark.parametrize("n", [1, 63, 64, 130, 257])
def test_gram_matches_oracle(family, D, n):
    torch = _torch()
    from probit_b200 import kernels as PK
    from helpers import make_prior
    rng = np.random.default_rng(n + D)
    X = rng.uniform(0, 1, size=(n, D))
    Y = rng.uniform(-0.5, 1.5, size=(45, D))
    th = 0.7 if family in ("eq", "matern12") else (0.7, 1.3)
    ko, kp = make_prior(OK, family)(th), make_prior(PK, family)(th)
    Kref = ko(X)
    K = kp(X).cpu().numpy()
    assert np.abs(K - Kref).max() < 1e-14 * max(1.0, np.abs(Kref).max())
    assert np.array_equal(K, K.T)                                  # mirror store is exact
    Kc = kp(X, Y).cpu().numpy()
    assert np.abs(Kc - ko(X, Y)).max() < 1e-14 * max(1.0, np.abs(Kref).max())


def test_gram_diag_add():
    torch = _torch()
    from probit_b200 import kernels as PK, linalg
    rng = np.random.default_rng(0)
    X = rng.uniform(0, 1, size=(150, 2))
    dv = rng.uniform(0.1, 1, size=150)
    k = 2.0 * PK.EQ().stretch(0.5)
    K = linalg.gram(k.lower(), X, diag_add=1e-3, diag_vec=dv).cpu().numpy()
    ref = (2.0 * OK.EQ().stretch(0.5))(X) + 1e-3 * np.eye(150) + np.diag(dv)
    assert np.abs(K - ref).max() < 1e-14


def _ordinal_inputs(n, J, seed):
    rng = np.random.default_rng(seed)
    cut = np.concatenate([[-np.inf], np.sort(rng.normal(0, 1, J - 1)), [np.inf]])
    y = rng.integers(0, J, size=n)
    f = rng.normal(0, 2.5, size=n)
    return f, y, (0.63, cut)


@pytest.mark.parametrize("J", [2, 3, 5, 11])
def test_ordinal_likelihood_matches_oracle(J):
    from probit_b200 import utilities as PU, _lib
    f, y, lp = _ordinal_inputs(4099, J, J)
    out = PU.evaluate_likelihood(_lib.PB_LIK_ORDINAL_PROBIT, f, y, lp, ("ll", "g", "h", "d3"))
    ref = {"ll": OU.log_probit_likelihood(f, y, lp), "g": OU.grad_log_probit_likelihood_autodiff(f, y, lp),
           "h": OU.hessian_log_probit_likelihood_autodiff(f, y, lp),
           "d3": OU.third_log_probit_likelihood_autodiff(f, y, lp)}
    # CUDA erf and SciPy erf differ by ~1 ulp, i.e. ~2e-16 ABSOLUTE on Z; every output divides by
    # u = Z + eps (up to the third power for d3) and h, d3 are differences of large terms, so the
    # admissible error is a few ulp / u relative to the magnitude T of the terms being combined
    # (SURVEY.md §7.2(d)).  Where Z is O(1) this is rounding level.
    s_, cut = lp
    z1 = np.where(np.isfinite(cut[y]), (np.where(np.isfinite(cut[y]), cut[y], 0) - f) / s_, 0.0)
    z2 = np.where(np.isfinite(cut[y + 1]), (np.where(np.isfinite(cut[y + 1]), cut[y + 1], 0) - f) / s_, 0.0)
    p1 = np.where(np.isfinite(cut[y]), OU.norm_z_pdf(z1), 0.0)
    p2 = np.where(np.isfinite(cut[y + 1]), OU.norm_z_pdf(z2), 0.0)
    u = OU.probit_likelihood(f, y, lp) + 1e-10
    g_ = np.abs(ref["g"])
    h_ = (np.abs(z1) * p1 + np.abs(z2) * p2) / (s_ ** 2 * u) + g_ ** 2
    d_ = ((z1 ** 2 + 1) * p1 + (z2 ** 2 + 1) * p2) / (s_ ** 3 * u) + 3 * g_ * h_ + g_ ** 3
    T = {"ll": 1.0 + np.abs(ref["ll"]), "g": 1.0 + g_, "h": 1.0 + h_, "d3": 1.0 + d_}
    for k in ref:
        got = out[k].cpu().numpy()
        assert np.all(np.isfinite(got))
        bound = (16 * 2.3e-16 / u + 1e-13) * T[k]
        assert np.all(np.abs(got - ref[k]) <= bound), (k, np.max(np.abs(got - ref[k]) / bound))
    ok = u > 1e-2          # rounding-level agreement away from the tails
    for k in ref:
        assert np.max(np.abs(out[k].cpu().numpy()[ok] - ref[k][ok]) / T[k][ok]) < 1e-12, k


def test_gaussian_and_safe_likelihood_match_oracle():
    from probit_b200 import utilities as PU, _lib
    rng = np.random.default_rng(3)
    f, yv = rng.normal(size=1000), rng.normal(size=1000)
    out = PU.evaluate_likelihood(_lib.PB_LIK_GAUSSIAN, f, yv, (0.3,), ("ll", "g", "h", "d3"))
    assert np.allclose(out["ll"].cpu().numpy(), OU.log_gaussian_likelihood(f, yv, (0.3,)), rtol=1e-14, atol=1e-14)
    assert np.allclose(out["g"].cpu().numpy(), OU.grad_log_gaussian_likelihood(f, yv, (0.3,)), rtol=1e-14)
    assert np.allclose(out["h"].cpu().numpy(), OU.hessian_log_gaussian_likelihood(f, yv, (0.3,)), rtol=1e-14)
    assert np.all(out["d3"].cpu().numpy() == 0)
    for single in (True, False):
        f, y, lp = _ordinal_inputs(5000, 5, 9)
        g = PU.grad_log_probit_likelihood(f, y, lp, single).cpu().numpy()
        h = PU.hessian_log_probit_likelihood(f, y, lp, single).cpu().numpy()
        gr, hr = OU.grad_log_probit_likelihood(f, y, lp, single), OU.hessian_log_probit_likelihood(f, y, lp, single)
        assert np.max(np.abs(g - gr) / (1 + np.abs(gr))) < 1e-10
        assert np.max(np.abs(h - hr) / (1 + np.abs(hr))) < 1e-10


def test_predictive_distributions_match_oracle():
    from probit_b200 import utilities as PU
    rng = np.random.default_rng(5)
    m, v = rng.normal(size=3001), rng.uniform(0.01, 2, size=3001)
    cut = np.array([-np.inf, -0.5, 0.1, 0.9, np.inf])
    P = PU.probit_predictive_distributions((0.6, cut), m, v).cpu().numpy()
    ref = OU.probit_predictive_distributions((0.6, cut), m, v)
    assert np.abs(P - ref).max() < 1e-14
    assert np.abs(P.sum(1) - 1).max() < 1e-14


@pytest.mark.parametrize("n", [1, 2, 63, 64, 65, 127, 128, 129, 1000, 4097])
def test_symv_lower_matches_dense_product_ragged_sizes(n):
    """The half-traffic symmetric matvec (lower tiles only) on ragged sizes; padding columns hold NaN on purpose."""
    torch = _torch()
    from probit_b200 import linalg
    rng = np.random.default_rng(n)
    A = rng.standard_normal((n, n)); A = A + A.T
    ld = (n + 63) // 64 * 64
    buf = torch.full((n, ld), float("nan"), dtype=torch.float64, device="cuda")
    buf[:, :n] = torch.as_tensor(A, device="cuda")
    x = rng.standard_normal(n)
    y = linalg.symv_lower(buf[:, :n], torch.as_tensor(x, device="cuda")).cpu().numpy()
    assert np.all(np.isfinite(y))
    assert relerr(y, A @ x) < 1e-13
    y2 = linalg.symv_lower(buf[:, :n], torch.as_tensor(x, device="cuda")).cpu().numpy()
    assert np.array_equal(y, y2)                      # deterministic: fixed summation order, no atomics


def test_symv_and_trsv_large_ragged():
    torch = _torch()
    from probit_b200 import linalg
    n = 1237
    rng = np.random.default_rng(2)
    A = rng.standard_normal((n, n)); A = A + A.T
    Ad = linalg.empty_matrix(n, n); Ad.copy_(torch.as_tensor(A, device="cuda"))
    x = rng.standard_normal(n)
    y = linalg.symv(Ad, torch.as_tensor(x, device="cuda")).cpu().numpy()
    assert relerr(y, A @ x) < 1e-13
