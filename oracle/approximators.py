"""Oracle restatement of probit/approximators.py, probit/implicit/Laplace.py and probit/implicit/VB.py.

Same class and method names and argument meaning as the reference so parity tests read like
the reference's own usage (examples/regression.py:128-151, examples/classification.py:402-425).
`prior(prior_parameters)` must return an `oracle.kernels.Kernel`; `log_likelihood` must be
`oracle.utilities.log_probit_likelihood` or `log_gaussian_likelihood`.

newton_form:
  "lu_jacobian" (default) — the literal reference operation sequence: dense Jacobian diag(h)K - I
                 (what jax.jacobian of f_root materialises, solvers.py:23-24) + LU solve.
  "cholesky_B"  — the algebraically identical SPD form the CUDA path uses:
                 w+ = b - s*B^{-1}(s*K b), b = W f + g, s = sqrt(W), B = I + s s^T o K.
  "signed_block" — the same step where some W < 0 (log(Z + 1e-10) is not log-concave for Z <~ 1e-10):
                 w+ = b - S (D + S K S)^{-1} S K b, S = |W|^1/2, D = sign(W), solved by block elimination
                 (Cholesky of the W >= 0 block, LU of the Schur complement) as fit.cu indefinite_newton_solve does.
                 Algebraically identical to the LU step; the two are compared to MEASURE how reproducible the
                 reference's own iterates are in that regime (tests/test_oracle_fit.py, tests/test_gpu_fit.py).
"""
import numpy as np
import scipy.linalg as sla

from . import utilities as U
from .solvers import fwd_solver, newton_solver

CHOLESKY_JITTER = 1e-12   # lab's B.epsilon added by B.cholesky(matrix.Dense) at Laplace.py:24


def _likelihood_family(log_likelihood, grad_log_likelihood, hessian_log_likelihood):
    if log_likelihood is U.log_probit_likelihood:
        g = grad_log_likelihood or U.grad_log_probit_likelihood_autodiff
        h = hessian_log_likelihood or U.hessian_log_probit_likelihood_autodiff
    elif log_likelihood is U.log_gaussian_likelihood:
        g = grad_log_likelihood or U.grad_log_gaussian_likelihood
        h = hessian_log_likelihood or U.hessian_log_gaussian_likelihood
    else:
        raise NotImplementedError("oracle supports the reference's two likelihoods only")
    return g, h


class Approximator:
    """approximators.py:16-210."""

    def __init__(self, data, prior, log_likelihood, grad_log_likelihood=None,
                 hessian_log_likelihood=None, tolerance=1e-5, newton_form="lu_jacobian",
                 dist_mode=None, maxiter=100):
        self.tolerance = tolerance                                     # approximators.py:90
        self.prior = prior
        self.log_likelihood = log_likelihood
        self.grad_log_likelihood, self.hessian_log_likelihood = _likelihood_family(
            log_likelihood, grad_log_likelihood, hessian_log_likelihood)
        X_train, _ = data
        X_train = np.asarray(X_train, dtype=np.float64)
        if X_train.ndim == 1:
            X_train = X_train[:, None]
        self.N, self.D = X_train.shape                                 # approximators.py:108
        self.data = (X_train, np.asarray(data[1]))
        self.newton_form = newton_form
        self.dist_mode = dist_mode
        self.maxiter = maxiter
        self.trace = []

    def _K(self, prior_parameters):
        return self.prior(prior_parameters)(self.data[0], dist_mode=self.dist_mode)

    def predict(self, X_test, parameters, weight, precision):
        """approximators.py:154-180 — literal: dense K_fs, LU solve against every test column."""
        kernel = self.prior(parameters[0])
        Kss = kernel.elwise(X_test, X_test).reshape(-1)                # :172
        Kfs = kernel(self.data[0], X_test, dist_mode=self.dist_mode)   # :173
        Kff = self._K(parameters[0])                                   # :174
        K = Kff + np.diag(1.0 / np.asarray(precision))                 # :175
        var = Kss - np.einsum("ij,ij->j", Kfs, np.linalg.solve(K, Kfs))   # :176-178
        mean = (Kfs.T @ weight).reshape(-1)                            # :179
        return mean, var

    def predict_covariance(self, X_test, parameters, weight, precision):
        """approximators.py:182-197."""
        kernel = self.prior(parameters[0])
        Kss = kernel(X_test, X_test, dist_mode=self.dist_mode)
        Kfs = kernel(self.data[0], X_test, dist_mode=self.dist_mode)
        K = self._K(parameters[0]) + np.diag(1.0 / np.asarray(precision))
        return Kss - Kfs.T @ np.linalg.solve(K, Kfs)

    def approximate_posterior(self, parameters):
        """approximators.py:204-210."""
        w = self.weight(parameters)
        p, _ = self.precision(w, parameters)
        return w, p


class LaplaceGP(Approximator):
    """approximators.py:213-277 with implicit/Laplace.py."""

    def __repr__(self):
        return "LaplaceGP"

    def construct(self):
        """approximators.py:238-246 -> f_LA (Laplace.py:4-9)."""
        def f(parameters, weight):
            K = self._K(parameters[0])
            posterior_mean = K @ weight
            return self.grad_log_likelihood(posterior_mean, self.data[1], parameters[1])
        return f

    def weight(self, parameters):
        """approximators.py:265-269."""
        K = self._K(parameters[0])   # loop-invariant; the reference rebuilds it per evaluation (same values)
        y, lik = self.data[1], parameters[1]
        self.trace = []
        self.negative_curvature = []
        z0 = np.zeros(self.N)
        if self.newton_form == "lu_jacobian":
            f = lambda z: self.grad_log_likelihood(K @ z, y, lik)
            jac = lambda z: self.hessian_log_likelihood(K @ z, y, lik)[:, None] * K
            return newton_solver(f, jac, z0, self.tolerance, self.maxiter, self.trace)

        def signed_step(w, fm, g, W):
            self.negative_curvature.append(int((W < 0).sum()))
            S = np.sqrt(np.abs(W))
            order = np.concatenate([np.flatnonzero(W >= 0), np.flatnonzero(W < 0)])   # stable partition
            p = int((W >= 0).sum())
            Sp, Kp = S[order], K[np.ix_(order, order)]
            M = (Sp[:, None] * Kp) * Sp[None, :]
            M[np.diag_indices(self.N)] += np.where(np.arange(self.N) < p, 1.0, -1.0)
            b = W * fm + g
            c = (S * (K @ b))[order]
            L = np.linalg.cholesky(M[:p, :p]) if p else np.zeros((0, 0))
            Yt = sla.solve_triangular(L, M[p:, :p].T, lower=True).T if p else np.zeros((self.N - p, 0))
            z1 = sla.solve_triangular(L, c[:p], lower=True) if p else np.zeros(0)
            x = np.empty(self.N)
            x2 = np.linalg.solve(M[p:, p:] - Yt @ Yt.T, c[p:] - Yt @ z1) if p < self.N else np.zeros(0)
            x[order[p:]] = x2
            if p:
                x[order[:p]] = sla.solve_triangular(L.T, z1 - Yt.T @ x2, lower=False)
            return b - S * x

        def step(w):
            fm = K @ w
            g = self.grad_log_likelihood(fm, y, lik)
            W = -self.hessian_log_likelihood(fm, y, lik)
            if self.newton_form == "signed_block":
                return signed_step(w, fm, g, W)
            s = np.sqrt(W)
            Bm = np.eye(self.N) + (s[:, None] * K) * s[None, :]
            L = np.linalg.cholesky(Bm)
            b = W * fm + g
            c = sla.solve_triangular(L, s * (K @ b), lower=True)
            return b - s * sla.solve_triangular(L.T, c, lower=False)
        return fwd_solver(step, z0, self.tolerance, self.maxiter, self.trace)

    def precision(self, weight, parameters):
        """approximators.py:271-277."""
        K = self._K(parameters[0])
        posterior_mean = K @ weight
        return -self.hessian_log_likelihood(posterior_mean, self.data[1], parameters[1]), posterior_mean

    def objective(self, jitter=CHOLESKY_JITTER):
        """approximators.py:248-263 -> objective_LA (Laplace.py:12-30)."""
        def obj(parameters):
            weight = self.weight(parameters)                            # fixed_point_layer forward
            K = self._K(parameters[0])
            posterior_mean = K @ weight
            precision = -self.hessian_log_likelihood(posterior_mean, self.data[1], parameters[1])
            L_cov = np.linalg.cholesky(K + np.diag(1.0 / precision) + jitter * np.eye(self.N))
            return (-np.sum(self.log_likelihood(posterior_mean, self.data[1], parameters[1]))
                    + 0.5 * posterior_mean @ weight
                    + np.sum(np.log(np.diag(L_cov)))
                    + 0.5 * np.sum(np.log(precision)))
        return obj


class VBGP(Approximator):
    """approximators.py:280-339 with implicit/VB.py."""

    def __repr__(self):
        return "VBGP"

    def construct(self):
        """approximators.py:306-314 -> f_VB (VB.py:4-16)."""
        def f(parameters, weight):
            K = self._K(parameters[0])
            N = self.N
            posterior_mean = K @ weight
            L_cov = np.linalg.cholesky(parameters[1][0] ** 2 * np.eye(N) + K)
            rhs = posterior_mean + parameters[1][0] * self.grad_log_likelihood(
                posterior_mean, self.data[1], parameters[1])
            return sla.cho_solve((L_cov, True), rhs)
        return f

    def weight(self, parameters):
        """approximators.py:332-334 — the factor is loop-invariant, so it is hoisted (same values)."""
        K = self._K(parameters[0])
        s = float(parameters[1][0])
        L_cov = np.linalg.cholesky(s**2 * np.eye(self.N) + K)
        y, lik = self.data[1], parameters[1]
        self.trace = []

        def f(w):
            m = K @ w
            return sla.cho_solve((L_cov, True), m + s * self.grad_log_likelihood(m, y, lik))
        return fwd_solver(f, np.zeros(self.N), self.tolerance, self.maxiter, self.trace)

    def precision(self, weight, parameters):
        """approximators.py:336-339."""
        K = self._K(parameters[0])
        return 1.0 / float(parameters[1][0]) ** 2 * np.ones(weight.shape[0]), K @ weight

    def objective(self):
        """approximators.py:316-330 -> objective_VB (VB.py:19-40), literal (explicit inverse)."""
        def obj(parameters):
            weight = self.weight(parameters)
            K = self._K(parameters[0])
            s = float(parameters[1][0])
            N = self.N
            posterior_mean = K @ weight
            L_cov = np.linalg.cholesky(s**2 * np.eye(N) + K)
            L_covT_inv = sla.solve_triangular(L_cov, np.eye(N), lower=True)
            cov = sla.solve_triangular(L_cov.T, L_covT_inv, lower=False)
            log_det_cov = -2 * np.sum(np.log(np.diag(L_cov)))
            trace_cov = np.sum(np.diag(cov))
            trace_posterior_cov_div_var = np.einsum("ij,ij->", K, cov)
            trace_K_inv_posterior_cov = s**2 * trace_cov
            return (0.5 * trace_posterior_cov_div_var + 0.5 * trace_K_inv_posterior_cov
                    + 0.5 * posterior_mean @ weight - N * np.log(s) - 0.5 * log_det_cov - 0.5 * N
                    - np.sum(self.log_likelihood(posterior_mean, self.data[1], parameters[1])))
        return obj
