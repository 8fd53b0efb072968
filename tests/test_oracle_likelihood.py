"""Pins the oracle's likelihood restatement (oracle/utilities.py) — CPU only.

The reference supplies no golden vectors for the hot path; these tests pin the oracle against
  * the reference's own h(x) assertions (probit/test/test_implicit.py:11-22),
  * 50-digit mpmath evaluation / differentiation of the literal reference expression
    log(Phi(z2) - Phi(z1) + 1e-10) (probit/utilities.py:56-57,195-229),
  * structural identities of the ordinal-probit model.
"""
import math

import mpmath as mp
import numpy as np
import pytest

from oracle import utilities as OU

mp.mp.dps = 50


def test_reference_h_assertions():
    """probit/test/test_implicit.py:11-22, statement by statement."""
    assert np.isnan(OU.h(0.0))
    assert OU.h(np.inf) == 0.0
    assert OU.h(-np.inf) == 0.0
    assert np.isclose(OU.h(1.0), -65 / 6, rtol=1.0001)
    assert np.isclose(OU.h(1.0), -1 + 5 / 2 - 37 / 3, rtol=1e-15)      # the exact value the loose rtol hides
    assert np.isnan(OU.grad_h(0.0))
    assert OU.grad_h(np.inf) == 0.0
    assert OU.grad_h(-np.inf) == 0.0


def test_product_h_matches_reference_assertions():
    from probit_b200.utilities import h
    assert math.isnan(h(0.0)) and h(math.inf) == 0.0 and h(-math.inf) == 0.0
    assert abs(h(1.0) - (-65 / 6)) < 1e-14


def _mp_ll(f, b1, b2, sigma, eps):
    """The literal reference expression at 50 digits."""
    def Phi(z):
        return (1 + mp.erf(z / mp.sqrt(2))) / 2
    c2 = mp.mpf(1) if b2 == math.inf else Phi((mp.mpf(b2) - f) / sigma)
    c1 = mp.mpf(0) if b1 == -math.inf else Phi((mp.mpf(b1) - f) / sigma)
    return mp.log(c2 - c1 + eps)


CASES = [(-math.inf, -0.4, 0.3), (-0.4, 0.7, 0.1), (0.7, math.inf, 1.5), (-0.4, 0.7, 2.5), (-math.inf, -0.4, 1.0),
         (0.7, math.inf, -1.2), (-1.0, -0.9, 0.0), (0.1, 0.2, -3.0)]


@pytest.mark.parametrize("b1,b2,f", CASES)
def test_ordinal_value_and_derivatives_match_mpmath(b1, b2, f):
    sigma, eps = 0.63, 1e-10
    cut = np.array([-np.inf, b1, b2, np.inf]) if np.isfinite(b1) and np.isfinite(b2) else (
        np.array([-np.inf, b2, np.inf]) if not np.isfinite(b1) else np.array([-np.inf, b1, np.inf]))
    y = np.array([1 if np.isfinite(b1) else 0])
    if np.isfinite(b1) and not np.isfinite(b2):
        y = np.array([1])
    lp = (sigma, cut)
    fa = np.array([f])
    fun = lambda x: _mp_ll(x, b1, b2, mp.mpf(sigma), mp.mpf(eps))
    ref = [float(mp.diff(fun, mp.mpf(f), n)) for n in range(4)]
    got = [OU.log_probit_likelihood(fa, y, lp)[0], OU.grad_log_probit_likelihood_autodiff(fa, y, lp)[0],
           OU.hessian_log_probit_likelihood_autodiff(fa, y, lp)[0], OU.third_log_probit_likelihood_autodiff(fa, y, lp)[0]]
    # float64 Phi-difference carries ~1e-16 absolute error, amplified by 1/(Z+eps)
    u = float(mp.e ** fun(mp.mpf(f)))
    for n, (g, r) in enumerate(zip(got, ref)):
        assert abs(g - r) <= (1e-13 + 4e-16 / u * 4 ** n) * (1 + abs(r)) * 10, (n, g, r)


def test_class_probabilities_sum_to_one_and_reflect():
    rng = np.random.default_rng(0)
    cut = np.array([-np.inf, -0.8, -0.1, 0.4, 1.1, np.inf])
    f = rng.normal(size=200)
    tot = sum(OU.probit_likelihood(f, np.full(200, j), (0.7, cut)) for j in range(5))
    assert np.allclose(tot, 1.0, atol=1e-15)
    # symmetry (f, b) -> (-f, -b) with class order reversed
    y = rng.integers(0, 5, size=200)
    a = OU.log_probit_likelihood(f, y, (0.7, cut))
    b = OU.log_probit_likelihood(-f, 4 - y, (0.7, -cut[::-1]))
    # Phi(z2)-Phi(z1) carries ~1e-16 absolute error either way round; log() amplifies it by 1/Z
    Z = OU.probit_likelihood(f, y, (0.7, cut))
    assert np.all(np.abs(a - b) <= 1e-15 / Z + 1e-14)
    ga = OU.grad_log_probit_likelihood_autodiff(f, y, (0.7, cut))
    gb = OU.grad_log_probit_likelihood_autodiff(-f, 4 - y, (0.7, -cut[::-1]))
    assert np.all(np.abs(ga + gb) <= (1e-15 / Z + 1e-13) * (1 + np.abs(ga)) * 10)


def test_derivatives_by_finite_differences():
    rng = np.random.default_rng(1)
    cut = np.array([-np.inf, -0.5, 0.5, np.inf])
    f = rng.normal(size=300) * 0.7      # stay out of the far tails: this is a coarse cross-check, mpmath pins the rest
    y = rng.integers(0, 3, size=300)
    lp = (0.63, cut)
    e = 1e-5
    ll = lambda x: OU.log_probit_likelihood(x, y, lp)
    g = OU.grad_log_probit_likelihood_autodiff(f, y, lp)
    hh = OU.hessian_log_probit_likelihood_autodiff(f, y, lp)
    d3 = OU.third_log_probit_likelihood_autodiff(f, y, lp)
    assert np.allclose((ll(f + e) - ll(f - e)) / (2 * e), g, rtol=1e-4, atol=1e-5)
    gfun = lambda x: OU.grad_log_probit_likelihood_autodiff(x, y, lp)
    assert np.allclose((gfun(f + e) - gfun(f - e)) / (2 * e), hh, rtol=1e-4, atol=1e-4)
    hfun = lambda x: OU.hessian_log_probit_likelihood_autodiff(x, y, lp)
    assert np.allclose((hfun(f + e) - hfun(f - e)) / (2 * e), d3, rtol=1e-3, atol=1e-3)


def test_gaussian_likelihood_closed_forms():
    f, y, s = np.array([0.3, -1.0]), np.array([0.1, 0.4]), 0.25
    ll = OU.log_gaussian_likelihood(f, y, (s,))
    assert np.allclose(ll, -0.5 * np.log(2 * np.pi) - np.log(s) - 0.5 * ((f - y) / s) ** 2, rtol=1e-15)
    assert np.allclose(OU.grad_log_gaussian_likelihood(f, y, (s,)), (y - f) / s**2)
    assert np.allclose(OU.hessian_log_gaussian_likelihood(f, y, (s,)), -1 / s**2)


def test_safe_mode_agrees_with_exact_where_the_reference_says_so():
    """utilities.py:39-41,77: the series path is 'accurate to three decimal places'; in the central region
    (|z| below the first bound) it IS the exact expression (without the +1e-10)."""
    cut = np.array([-np.inf, -0.5, 0.5, np.inf])
    f = np.linspace(-0.3, 0.3, 50)
    y = np.ones(50, dtype=np.int64)
    lp = (1.0, cut)          # |z| <= 0.8 < 1.3: central branch
    g_safe = OU.grad_log_probit_likelihood(f, y, lp)
    h_safe = OU.hessian_log_probit_likelihood(f, y, lp)
    g = OU.grad_log_probit_likelihood_autodiff(f, y, lp, eps=0.0)
    hh = OU.hessian_log_probit_likelihood_autodiff(f, y, lp, eps=0.0)
    assert np.allclose(g_safe, g, rtol=1e-13, atol=1e-15) and np.allclose(h_safe, hh, rtol=1e-12)
    # far tail: linear / constant approximations (utilities.py:166-167,190-191)
    ff = np.array([-10.0, 10.0])
    yy = np.array([2, 0])
    gs = OU.grad_log_probit_likelihood(ff, yy, lp)
    hs = OU.hessian_log_probit_likelihood(ff, yy, lp)
    assert np.allclose(gs, [10.5, -10.5]) and np.allclose(hs, [-1.0, -1.0])


def test_predictive_distributions_rows_sum_to_one():
    rng = np.random.default_rng(2)
    m, v = rng.normal(size=100), rng.uniform(0.1, 2.0, size=100)
    cut = np.array([-np.inf, -0.5, 0.2, 0.9, np.inf])
    P = OU.probit_predictive_distributions((0.6, cut), m, v)
    assert P.shape == (100, 4) and np.allclose(P.sum(1), 1.0, atol=1e-15) and np.all(P >= 0)
    s = np.sqrt(v + 0.36)
    from scipy.stats import norm
    assert np.allclose(P[:, 1], norm.cdf((0.2 - m) / s) - norm.cdf((-0.5 - m) / s), atol=1e-15)
