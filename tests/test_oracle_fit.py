"""Pins the oracle's solver / approximator restatement — CPU only (closed forms, algebraic identities, fixtures)."""
import glob
import os

import numpy as np
import pytest
import scipy.linalg as sla

from helpers import make_prior, ordinal_problem, regression_problem, relerr
from oracle import approximators as OA, kernels as OK, solvers as OS, utilities as OU

GOLDEN = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))
                if not os.path.basename(p).startswith("ref_"))


def test_fwd_solver_follows_jaxopt_stopping_rule():
    # x <- 0.5 x + 1 converges to 2; error_k = |x_{k+1} - x_k| = 2^-k... stop at the first error <= tol
    trace = []
    z = OS.fwd_solver(lambda x: 0.5 * x + 1.0, np.zeros(1), 1e-3, trace=trace)
    assert trace[-1] <= 1e-3 < trace[-2]
    assert len(trace) == 11 and abs(z[0] - (2 - 2.0**-10)) < 1e-15
    # maxiter caps the loop and the last iterate is returned
    trace = []
    z = OS.fwd_solver(lambda x: x + 1.0, np.zeros(1), 1e-3, maxiter=7, trace=trace)
    assert len(trace) == 7 and z[0] == 7.0


def test_newton_solver_on_a_linear_map_converges_in_one_step_plus_confirmation():
    A = np.array([[0.2, 0.1], [0.0, 0.3]])
    c = np.array([1.0, -2.0])
    trace = []
    z = OS.newton_solver(lambda x: A @ x + c, lambda x: A, np.zeros(2), 1e-8, trace=trace)
    assert np.allclose(z, np.linalg.solve(np.eye(2) - A, c)) and len(trace) == 2


def test_gaussian_laplace_is_exact_gp_regression():
    X, y, params, family = regression_problem(0, 20)
    gp = OA.LaplaceGP((X, y), make_prior(OK, family), OU.log_gaussian_likelihood)
    w, p = gp.approximate_posterior(params)
    K = make_prior(OK, family)(params[0])(X)
    s2 = params[1][0] ** 2
    wc = np.linalg.solve(K + s2 * np.eye(20), y)
    assert relerr(w, wc) < 1e-12 and np.allclose(p, 1 / s2) and len(gp.trace) == 2
    nlml = 0.5 * y @ wc + 0.5 * np.linalg.slogdet(K + s2 * np.eye(20))[1] + 10 * np.log(2 * np.pi)
    assert abs(gp.objective()(params) - nlml) < 1e-9
    Xs = np.linspace(-0.5, 1.5, 50)[:, None]
    m, v = gp.predict(Xs, params, w, p)
    Ks = make_prior(OK, family)(params[0])(X, Xs)
    assert np.allclose(m, Ks.T @ wc) and np.allclose(v, params[0][1] - np.einsum("ij,ij->j", Ks, np.linalg.solve(K + s2 * np.eye(20), Ks)))


# seed, N, D, J, family, sigma, reproducible?   (shared with tests/test_gpu_fit.py)
#   True      : iteration count and weights survive a re-ordering of the linear algebra AND a 1-ulp change of Phi
#   "weights" : every variant reaches the same fixed point (weights to 1e-9) but the iteration count moves
#   False     : knife-edge stopping test or chaotic wandering between several fixed points: no single reference answer
NEGATIVE_CURVATURE_CASES = [
    (17, 257, 1, 3, "eq", 0.08, True), (21, 400, 2, 4, "matern12", 0.12, True), (4, 600, 3, 4, "eq", 0.18, True),
    # 31 iterations, 32 with Phi = erfc(-z / sqrt 2) / 2 (same fixed point on the CPU); the CUDA arithmetic — every piece
    # within 2 ulp of the reference's — takes 40 steps or lands on ANOTHER fixed point of the same map (12 % away),
    # depending on the build: several fixed points, reached through ~30 wandering steps
    (26, 400, 2, 4, "matern12", 0.12, False),
    (27, 300, 1, 3, "eq", 0.1, False),        # 51 / 74 / 90 iterations for three 1-ulp-equivalent Phi, two different fixed points
    (9, 500, 1, 5, "eq", 0.1, False),         # last step 9.5e-6 vs tol 1e-5: the iteration count is a coin toss (7 or 8)
    (3, 350, 1, 3, "matern12", 0.1, False)]   # wanders for ~35 steps and lands on DIFFERENT fixed points (rel. diff 0.3)


@pytest.mark.parametrize("seed,N,D,J,family,sigma,reproducible", NEGATIVE_CURVATURE_CASES)
def test_signed_block_form_and_how_reproducible_the_reference_is_with_negative_curvature(seed, N, D, J, family, sigma,
                                                                                         reproducible, monkeypatch):
    """With a small noise std, log(Z + 1e-10) has positive second derivative where Z <~ 1e-10 and the reference's LU step
    (solvers.py:24) walks through indefinite Jacobians.  The signed block elimination (what fit.cu does) is the same
    step algebraically, and the CUDA likelihood's Phi is the reference's up to ~1 ulp.  Where Z ~ 1e-10 the Hessian is
    defined to ~1e-6 relative only (1e-16 of rounding in a Phi difference divided by Z + 1e-10), so an iteration that
    wanders for tens of steps amplifies exactly those differences.  Three variants of the reference's own arithmetic are
    run here — the literal LU form, the signed block form, and the LU form with Phi(z) = erfc(-z / sqrt 2) / 2 in place
    of (1 + erf(z / sqrt 2)) / 2 (utilities.py:18-19), a 1-ulp-equivalent expression — and the case is classified by
    what survives: where all agree the reference's answer is REPRODUCIBLE and the CUDA path is held to it at 1e-8;
    otherwise there is no single reference answer for that quantity and the CUDA path is held to what every variant
    satisfies (convergence to a fixed point of the reference's map)."""
    from scipy.special import erfc
    X, y, params, family = ordinal_problem(seed, N, D, J, family)
    prm = (params[0], (sigma, params[1][1]))
    a = OA.LaplaceGP((X, y), make_prior(OK, family), OU.log_probit_likelihood)
    b = OA.LaplaceGP((X, y), make_prior(OK, family), OU.log_probit_likelihood, newton_form="signed_block")
    wa, _ = a.approximate_posterior(prm)
    wb, _ = b.approximate_posterior(prm)
    K = a._K(prm[0])
    with monkeypatch.context() as mp_:
        mp_.setattr(OU, "ndtr", lambda z: 0.5 * erfc(-z / OU.sqrt_2))
        c = OA.LaplaceGP((X, y), make_prior(OK, family), OU.log_probit_likelihood)
        wc, _ = c.approximate_posterior(prm)
    assert max(b.negative_curvature) > 0                      # the indefinite step is exercised
    counts = {len(a.trace), len(b.trace), len(c.trace)}
    assert max(counts) < 100                                  # every variant converges
    same_point = relerr(wb, wa) < 5e-9 and relerr(wc, wa) < 5e-9
    if reproducible is True:
        assert len(counts) == 1 and same_point
    elif reproducible == "weights":
        assert len(counts) > 1 and same_point
    else:
        assert len(counts) > 1
    # either way the result is a fixed point of the reference's map: one more literal Newton step moves it by <= tol
    for w in (wa, wb, wc):
        fm = K @ w
        g, h = OU.grad_log_probit_likelihood_autodiff(fm, y, prm[1]), OU.hessian_log_probit_likelihood_autodiff(fm, y, prm[1])
        step = np.linalg.solve(h[:, None] * K - np.eye(N), g - w)
        assert np.linalg.norm(step) < 1e-5


@pytest.mark.parametrize("N,D,J,family", [(60, 1, 3, "eq"), (300, 4, 5, "matern12")])
def test_lu_jacobian_and_cholesky_forms_give_the_same_iterates(N, D, J, family):
    X, y, params, _ = ordinal_problem(N, N, D, J, family)
    a = OA.LaplaceGP((X, y), make_prior(OK, family), OU.log_probit_likelihood, newton_form="lu_jacobian")
    b = OA.LaplaceGP((X, y), make_prior(OK, family), OU.log_probit_likelihood, newton_form="cholesky_B")
    wa, pa = a.approximate_posterior(params)
    wb, pb = b.approximate_posterior(params)
    assert len(a.trace) == len(b.trace)
    assert relerr(wb, wa) < 1e-11 and relerr(pb, pa) < 1e-11
    assert np.allclose(a.trace[:-1], b.trace[:-1], rtol=1e-8)
    # fixed point: g(K w) = w
    K = make_prior(OK, family)(params[0])(X)
    assert np.linalg.norm(OU.grad_log_probit_likelihood_autodiff(K @ wa, y, params[1]) - wa) < 1e-8


def test_objective_LA_equals_B_form():
    X, y, params, family = ordinal_problem(3, 150, 2, 3, "eq")
    gp = OA.LaplaceGP((X, y), make_prior(OK, family), OU.log_probit_likelihood)
    w, p = gp.approximate_posterior(params)
    K = make_prior(OK, family)(params[0])(X)
    f = K @ w
    s = np.sqrt(p)
    Bm = np.eye(150) + s[:, None] * (K + 1e-12 * np.eye(150)) * s[None, :]
    b_form = -np.sum(OU.log_probit_likelihood(f, y, params[1])) + 0.5 * f @ w + np.sum(np.log(np.diag(np.linalg.cholesky(Bm))))
    assert abs(gp.objective()(params) - b_form) < 1e-10 * abs(b_form)


def test_vb_objective_closed_form_equals_literal():
    X, y, params, family = ordinal_problem(5, 90, 2, 3, "eq")
    gp = OA.VBGP((X, y), make_prior(OK, family), OU.log_probit_likelihood)
    w, p = gp.approximate_posterior(params)
    assert gp.trace[-1] <= 1e-5 and np.allclose(p, 1 / params[1][0] ** 2)
    K = make_prior(OK, family)(params[0])(X)
    s = params[1][0]
    L = np.linalg.cholesky(s * s * np.eye(90) + K)
    f = K @ w
    closed = 0.5 * f @ w - 90 * np.log(s) + np.sum(np.log(np.diag(L))) - np.sum(OU.log_probit_likelihood(f, y, params[1]))
    assert abs(gp.objective()(params) - closed) < 1e-9 * abs(closed)
    # VB fixed point: (s^2 I + K) w = K w + s g(K w)
    g = OU.grad_log_probit_likelihood_autodiff(f, y, params[1])
    assert np.linalg.norm((s * s * np.eye(90) + K) @ w - (f + s * g)) < 1e-4


def test_kernel_shim_semantics():
    rng = np.random.default_rng(0)
    x, y = rng.uniform(size=(7, 3)), rng.uniform(size=(5, 3))
    d2 = ((x[:, None, :] - y[None, :, :]) ** 2).sum(-1)
    assert np.allclose(OK.EQ()(x, y), np.exp(-0.5 * d2), rtol=1e-15)
    assert np.allclose(OK.Matern12()(x, y), np.exp(-np.sqrt(d2)), rtol=1e-15)
    assert np.allclose((2.5 * OK.EQ().stretch(0.7))(x, y), 2.5 * np.exp(-0.5 * d2 / 0.49), rtol=1e-14)
    x1, y1 = rng.uniform(size=6), rng.uniform(size=4)
    per = (1.3 * OK.EQ().stretch(0.8).periodic(0.5))(x1, y1)
    assert np.allclose(per, 1.3 * np.exp(-2 * np.sin(np.pi * (x1[:, None] - y1[None, :]) / 0.5) ** 2 / 0.64), rtol=1e-12)
    assert (1.3 * OK.EQ().stretch(0.8).periodic(0.5)).elwise(x1, x1).shape == (6, 1)
    assert np.allclose((1.3 * OK.EQ()).elwise(x, x), 1.3)


def test_expansion_form_distance_is_the_references_noise_floor():
    """lab's pw_dists2 expansion puts O(1e-8) noise on diag(K) for Matern12 with D>1."""
    rng = np.random.default_rng(1)
    x = rng.uniform(size=(200, 4))
    Kd = OK.Matern12()(x)
    Ke = OK.Matern12()(x, dist_mode="expand")
    assert np.all(np.diag(Kd) == 1.0)
    noise = np.abs(np.diag(Ke) - 1.0).max()
    assert 1e-10 < noise < 1e-6
    assert np.abs(OK.EQ()(x) - OK.EQ()(x, dist_mode="expand")).max() < 1e-14


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_reproduces_committed_fixtures(path):
    g = np.load(path)
    family, gaussian, cls = str(g["family"]), bool(g["gaussian"]), str(g["cls"])
    theta = tuple(g["theta"]) if g["theta"].ndim else float(g["theta"])
    lik = (float(g["sigma"]),) if gaussian else (float(g["sigma"]), g["cutpoints"])
    params = (theta, lik)
    extra = dict(grad_log_likelihood=OU.grad_log_probit_likelihood,
                 hessian_log_likelihood=OU.hessian_log_probit_likelihood) if ("safe" in g.files and bool(g["safe"])) else {}
    gp = getattr(OA, cls)((g["X"], g["y"]), make_prior(OK, family),
                          OU.log_gaussian_likelihood if gaussian else OU.log_probit_likelihood, **extra)
    w, p = gp.approximate_posterior(params)
    assert len(gp.trace) == int(g["iterations"])
    assert relerr(w, g["weight"]) < 1e-11 and relerr(p, g["precision"]) < 1e-11
    m, v = gp.predict(g["Xs"], params, w, p)
    assert relerr(m, g["mean"]) < 1e-10 and relerr(v, g["variance"]) < 1e-10
    assert abs(gp.objective()(params) - float(g["objective"])) < 1e-10 * abs(float(g["objective"]))
    assert len(GOLDEN) >= 4
