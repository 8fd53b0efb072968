"""Shared test helpers: seeded problem builders that feed the SAME inputs to the oracle and the CUDA path."""
import numpy as np

from oracle import kernels as OK, utilities as OU, approximators as OA
from probit_b200.datasets import generate_ordinal_data, generate_regression_data


def numpy_latent_sampler(kernel, jitter):
    """f = chol(K + jitter I) z with the ORACLE kernels (test-side input generation only)."""
    def sample(X, z):
        K = kernel(X) + jitter * np.eye(len(X))
        return np.linalg.cholesky(K) @ z
    return sample


def make_prior(ns, family):
    """The reference examples' priors, parameterised by the namespace providing EQ/Matern12 (oracle or product)."""
    if family == "matern12":          # examples/classification.py:375
        return lambda l: 1.0 * ns.Matern12().stretch(l)
    if family == "eq":                # examples/classification.py:389-391
        return lambda l: 1.0 * ns.EQ().stretch(l)
    if family == "eq_periodic":       # examples/regression.py:120-123
        return lambda th: th[1] * ns.EQ().stretch(th[0]).periodic(0.5)
    if family == "eq_scaled":
        return lambda th: th[1] * ns.EQ().stretch(th[0])
    raise KeyError(family)


def ordinal_problem(seed, N, D, J, family="matern12", lengthscale=1.0, noise_variance=0.4):
    gen = 1.0 * OK.Matern12().stretch(1.0)
    X, g, y, cut = generate_ordinal_data(seed, N, D, J, noise_variance, numpy_latent_sampler(gen, 1e-6))
    params = (lengthscale, (float(np.sqrt(noise_variance)), cut))
    return X, y, params, family


def regression_problem(seed, N, D=1, family="eq_periodic", theta=(0.3, 0.8), noise_std=0.2):
    gen = make_prior(OK, family)((1.0, 1.0))
    X, y, f = generate_regression_data(seed, N, D, noise_std, numpy_latent_sampler(gen, 1e-10))
    params = (theta, (0.25,))
    return X, y, params, family


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
