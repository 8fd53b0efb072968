"""Oracle restatement of probit/utilities.py (likelihoods and their derivatives), vectorised over N.

Every function cites the reference lines it follows (paths relative to /root/reference).
The reference takes derivatives by JAX autodiff of log_probit_likelihood
(probit/approximators.py:92-95); here they are the closed forms autodiff produces,
pinned against 50-digit mpmath differentiation in tests/test_oracle_likelihood.py.
"""
import numpy as np
from scipy.special import erf

over_sqrt_2_pi = 1.0 / np.sqrt(2 * np.pi)            # utilities.py:10
log_over_sqrt_2_pi = np.log(over_sqrt_2_pi)          # utilities.py:11
sqrt_2 = np.sqrt(2.0)                                # utilities.py:12
BOUNDS = {"single": [1.3, 1.8, 2.3], "double": [2.3, 3.6, 4.8]}   # utilities.py:15
LIKELIHOOD_EPS = 1e-10                               # utilities.py:57


def ndtr(z):
    """utilities.py:18-19."""
    return 0.5 * (1 + erf(z / sqrt_2))


def norm_z_pdf(z):
    """utilities.py:22-23."""
    return over_sqrt_2_pi * np.exp(-0.5 * z**2)


def norm_pdf(x, loc=0.0, scale=1.0):
    """utilities.py:26-28."""
    z = (x - loc) / scale
    return norm_z_pdf(z) / scale


def norm_cdf(x):
    """utilities.py:31-34."""
    x = np.asarray(x, dtype=np.float64)
    _x = np.where(np.isinf(x), 1.0, x)
    extrema = np.where(x == np.inf, 1.0, 0.0)
    return np.where(np.isinf(x), extrema, ndtr(_x))


def h(x):
    """utilities.py:37-44 — series polynomial (the reference's only unit-tested function)."""
    x = np.asarray(x, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        return -1 * x**-2.0 + 5 / 2 * x**-4.0 - 37 / 3 * x**-6.0


def grad_h(x):
    """d/dx of utilities.py:44 (what jax.grad(h) returns; test_implicit.py:20-22)."""
    x = np.asarray(x, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        return 2 * x**-3.0 - 10 * x**-5.0 + 74 * x**-7.0


def probit(noise_std, cutpoints_y, cutpoints_yplus1, f):
    """utilities.py:195-229 — Z = Phi((b_{y+1}-f)/s) - Phi((b_y-f)/s) with +-inf guards."""
    cutpoints_y = np.asarray(cutpoints_y, dtype=np.float64)
    cutpoints_yplus1 = np.asarray(cutpoints_yplus1, dtype=np.float64)
    with np.errstate(invalid="ignore"):
        safe_z1s = np.where(cutpoints_y == -np.inf, 0.0, cutpoints_y - f)
        safe_z2s = np.where(cutpoints_yplus1 == np.inf, 0.0, cutpoints_yplus1 - f)
    norm_cdf_z1s = np.where(cutpoints_y == -np.inf, 0.0, norm_cdf(safe_z1s / noise_std))
    norm_cdf_z2s = np.where(cutpoints_yplus1 == np.inf, 1.0, norm_cdf(safe_z2s / noise_std))
    return norm_cdf_z2s - norm_cdf_z1s


def _split(likelihood_parameters):
    noise_std = float(likelihood_parameters[0])
    cutpoints = np.asarray(likelihood_parameters[1], dtype=np.float64)
    return noise_std, cutpoints


def probit_likelihood(f, y, likelihood_parameters):
    """utilities.py:47-53."""
    noise_std, cutpoints = _split(likelihood_parameters)
    y = np.asarray(y, dtype=np.int64)
    return probit(noise_std, cutpoints[y], cutpoints[y + 1], f)


def log_probit_likelihood(f, y, likelihood_parameters, eps=LIKELIHOOD_EPS):
    """utilities.py:56-57."""
    return np.log(probit_likelihood(f, y, likelihood_parameters) + eps)


def _probit_terms(f, y, likelihood_parameters, eps):
    noise_std, cutpoints = _split(likelihood_parameters)
    y = np.asarray(y, dtype=np.int64)
    f = np.asarray(f, dtype=np.float64)
    b1, b2 = cutpoints[y], cutpoints[y + 1]
    fin1, fin2 = b1 != -np.inf, b2 != np.inf
    with np.errstate(invalid="ignore"):
        z1 = np.where(fin1, (np.where(fin1, b1, 0.0) - f) / noise_std, 0.0)
        z2 = np.where(fin2, (np.where(fin2, b2, 0.0) - f) / noise_std, 0.0)
    p1 = np.where(fin1, norm_z_pdf(z1), 0.0)   # the where-guards stop gradient flow at infinite cutpoints
    p2 = np.where(fin2, norm_z_pdf(z2), 0.0)
    u = probit(noise_std, b1, b2, f) + eps
    return noise_std, z1, z2, p1, p2, u


def grad_log_probit_likelihood_autodiff(f, y, likelihood_parameters, eps=LIKELIHOOD_EPS):
    """d/df of utilities.py:56-57 — what `grad(log_likelihood)` (approximators.py:92-93) evaluates."""
    s, z1, z2, p1, p2, u = _probit_terms(f, y, likelihood_parameters, eps)
    return (p1 - p2) / (s * u)


def hessian_log_probit_likelihood_autodiff(f, y, likelihood_parameters, eps=LIKELIHOOD_EPS):
    """d2/df2 of utilities.py:56-57 (approximators.py:94-95)."""
    s, z1, z2, p1, p2, u = _probit_terms(f, y, likelihood_parameters, eps)
    g = (p1 - p2) / (s * u)
    return (z1 * p1 - z2 * p2) / (s * s * u) - g * g


def third_log_probit_likelihood_autodiff(f, y, likelihood_parameters, eps=LIKELIHOOD_EPS):
    """d3/df3 of utilities.py:56-57 — needed by the implicit gradient (solvers.py:52-64 differentiates f_LA)."""
    s, z1, z2, p1, p2, u = _probit_terms(f, y, likelihood_parameters, eps)
    g = (p1 - p2) / (s * u)
    hh = (z1 * p1 - z2 * p2) / (s * s * u) - g * g
    return ((z1 * z1 - 1) * p1 - (z2 * z2 - 1) * p2) / (s**3 * u) - 3 * g * hh - g**3


def norm_z_logpdf(x):
    """utilities.py:64-65."""
    return log_over_sqrt_2_pi - x**2 / 2.0


def norm_logpdf(x, loc=0.0, scale=1.0):
    """utilities.py:68-70."""
    z = (x - loc) / scale
    return norm_z_logpdf(z) - np.log(scale)


def log_gaussian_likelihood(f, y, likelihood_parameters):
    """utilities.py:60-61."""
    return norm_logpdf(np.asarray(f, dtype=np.float64), loc=np.asarray(y, dtype=np.float64),
                       scale=float(likelihood_parameters[0]))


def grad_log_gaussian_likelihood(f, y, likelihood_parameters):
    s = float(likelihood_parameters[0])
    return (np.asarray(y, dtype=np.float64) - f) / (s * s)


def hessian_log_gaussian_likelihood(f, y, likelihood_parameters):
    s = float(likelihood_parameters[0])
    return -np.ones_like(np.asarray(f, dtype=np.float64)) / (s * s)


# ---- the optional "safe" path (utilities.py:73-192); no caller inside the reference repo ----

def _Z_far_tails(z):
    """utilities.py:83-85."""
    with np.errstate(all="ignore"):
        return over_sqrt_2_pi / z * np.exp(-0.5 * z**2 + h(z))


def _Z_tails(z1, z2):
    """utilities.py:73-80."""
    return _Z_far_tails(z1) - _Z_far_tails(z2)


def _safe_Z(f, y, likelihood_parameters, upper_bound=np.inf, upper_bound2=np.inf, upper_bound3=np.inf):
    """utilities.py:88-148, statement by statement."""
    noise_std, cutpoints = _split(likelihood_parameters)
    y = np.asarray(y, dtype=np.int64)
    f = np.asarray(f, dtype=np.float64)
    cutpoints_tplus1 = cutpoints[y + 1]
    cutpoints_t = cutpoints[y]
    _b = np.where(cutpoints_tplus1 == np.inf, 0.0, cutpoints_tplus1)
    _a = np.where(cutpoints_t == -np.inf, 0.0, cutpoints_t)
    z2s = np.where(cutpoints_tplus1 == np.inf, np.inf, (_b - f) / noise_std)
    z1s = np.where(cutpoints_t == -np.inf, -np.inf, (_a - f) / noise_std)
    SAFE = 1.0
    Z = norm_cdf(z2s) - norm_cdf(z1s)
    _z1s = np.where((upper_bound < z1s) & (z1s <= upper_bound2), z1s, SAFE)
    __z2s = np.where(upper_bound < z1s, z2s, SAFE)
    _z2s = np.where((-upper_bound2 <= z2s) & (z2s < -upper_bound), z2s, SAFE)
    __z1s = np.where(-upper_bound > z2s, z1s, SAFE)
    Z = np.where(z1s > upper_bound, _Z_tails(_z1s, __z2s), Z)
    Z = np.where(z2s < -upper_bound, _Z_tails(__z1s, _z2s), Z)
    _z1s = np.where((upper_bound2 < np.abs(z1s)) & (np.abs(z1s) < upper_bound3), z1s, SAFE)
    _z2s = np.where((upper_bound2 < np.abs(z2s)) & (np.abs(z2s) < upper_bound3), z2s, SAFE)
    Z = np.where(z1s > upper_bound2, _Z_far_tails(_z1s), Z)
    Z = np.where(z2s < -upper_bound2, _Z_far_tails(-_z2s), Z)
    Z = np.where(z1s >= upper_bound3, SAFE, Z)
    Z = np.where(z2s <= -upper_bound3, SAFE, Z)
    return Z, z1s, z2s


def grad_log_probit_likelihood(f, y, likelihood_parameters, single_precision=True):
    """utilities.py:151-169."""
    ub, ub2, ub3 = BOUNDS["single" if single_precision else "double"]
    noise_std = float(likelihood_parameters[0])
    Z, z1s, z2s = _safe_Z(f, y, likelihood_parameters, ub, ub2, ub3)
    with np.errstate(all="ignore"):
        E = (norm_pdf(z1s) - norm_pdf(z2s)) / Z
    E = np.where(z1s > ub3, z1s, E)
    E = np.where(z2s < -ub3, z2s, E)
    return E / noise_std


def hessian_log_probit_likelihood(f, y, likelihood_parameters, single_precision=True):
    """utilities.py:172-192."""
    ub, ub2, ub3 = BOUNDS["single" if single_precision else "double"]
    noise_std = float(likelihood_parameters[0])
    Z, z1s, z2s = _safe_Z(f, y, likelihood_parameters, ub, ub2, ub3)
    p1, p2 = norm_pdf(z1s), norm_pdf(z2s)
    w = grad_log_probit_likelihood(f, y, likelihood_parameters, single_precision)
    _z1s = np.where(np.isinf(z1s), 0.0, z1s)
    _z2s = np.where(np.isinf(z2s), 0.0, z2s)
    with np.errstate(all="ignore"):
        V = -(w**2) + (_z1s * p1 - _z2s * p2) / Z / noise_std**2
    V = np.where(z1s > ub3, -(noise_std**-2.0), V)
    V = np.where(z2s < -ub3, -(noise_std**-2.0), V)
    return V


def probit_predictive_distributions(likelihood_parameters, posterior_mean, posterior_variance):
    """utilities.py:232-249 — (N_test, J) table of class probabilities."""
    noise_std, cutpoints = _split(likelihood_parameters)
    posterior_mean = np.asarray(posterior_mean, dtype=np.float64)
    J = cutpoints.size - 1
    out = np.ones((posterior_mean.shape[0], J))
    posterior_pred_std = np.sqrt(np.asarray(posterior_variance, dtype=np.float64) + noise_std**2)
    for j in range(J):
        out[:, j] = probit(posterior_pred_std, np.full_like(posterior_mean, cutpoints[j]),
                           np.full_like(posterior_mean, cutpoints[j + 1]), posterior_mean)
    return out
