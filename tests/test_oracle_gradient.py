"""Pins the oracle's closed-form evidence gradient (oracle/gradients.py) against central finite differences of
the oracle objective — the same numerical check the reference plots (examples/classification.py:104-111). CPU only."""
import numpy as np
import pytest

from helpers import make_prior, ordinal_problem, regression_problem
from oracle import approximators as OA, gradients as OG, kernels as OK, utilities as OU


def _fd(fun, x, h=1e-5):
    return (fun(x + h) - fun(x - h)) / (2 * h)


@pytest.mark.parametrize("family,base", [("eq_scaled", "eq"), ("matern_scaled", "exp")])
def test_ordinal_gradient_matches_finite_differences(family, base):
    X, y, params, _ = ordinal_problem(3, 150, 2, 3, "eq")
    lik = params[1]
    prior = make_prior(OK, "eq_scaled") if base == "eq" else (lambda th: th[1] * OK.Matern12().stretch(th[0]))

    def obj(l, c):
        gp = OA.LaplaceGP((X, y), prior, OU.log_probit_likelihood, tolerance=1e-10)
        return gp.objective(jitter=0.0)(((l, c), lik))

    l, c = 0.9, 1.3
    gp = OA.LaplaceGP((X, y), prior, OU.log_probit_likelihood, tolerance=1e-10)
    w = gp.weight(((l, c), lik))
    G = OG.laplace_gradient(prior((l, c))(X), X, y, w, lik,
                            dict(base=base, periodic=0, scale=c, stretch_in=1.0, period=1.0, stretch_out=l), False)
    assert abs(G["stretch_out"] - _fd(lambda t: obj(t, c), l)) < 1e-6 * max(1, abs(G["stretch_out"]))
    assert abs(G["scale"] - _fd(lambda t: obj(l, t), c)) < 1e-6 * max(1, abs(G["scale"]))


def test_gaussian_periodic_gradient_matches_finite_differences():
    X, y, _, family = regression_problem(0, 20)
    prior = make_prior(OK, family)

    def obj(l, c, s):
        gp = OA.LaplaceGP((X, y), prior, OU.log_gaussian_likelihood, tolerance=1e-10)
        return gp.objective(jitter=0.0)(((l, c), (s,)))

    l, c, s = 0.3, 0.8, 0.25
    gp = OA.LaplaceGP((X, y), prior, OU.log_gaussian_likelihood, tolerance=1e-10)
    w = gp.weight(((l, c), (s,)))
    G = OG.laplace_gradient(prior((l, c))(X), X, y, w, (s,),
                            dict(base="eq", periodic=1, scale=c, stretch_in=1.0, period=0.5, stretch_out=l), True)
    assert abs(G["stretch_out"] - _fd(lambda t: obj(t, c, s), l)) < 1e-6 * abs(G["stretch_out"])
    assert abs(G["scale"] - _fd(lambda t: obj(l, t, s), c)) < 1e-6 * abs(G["scale"])
    assert abs(G["sigma"] - _fd(lambda t: obj(l, c, t), s)) < 1e-6 * abs(G["sigma"])


def test_ordinal_likelihood_parameter_gradients_match_finite_differences():
    X, y, params, _ = ordinal_problem(3, 150, 2, 4, "eq")
    sig, cut = params[1]
    prior = make_prior(OK, "eq_scaled")
    th = (0.9, 1.3)

    def obj(s, c):
        gp = OA.LaplaceGP((X, y), prior, OU.log_probit_likelihood, tolerance=1e-10)
        return gp.objective(jitter=0.0)((th, (s, c)))

    gp = OA.LaplaceGP((X, y), prior, OU.log_probit_likelihood, tolerance=1e-10)
    w = gp.weight((th, (sig, cut)))
    G = OG.laplace_gradient(prior(th)(X), X, y, w, (sig, cut),
                            dict(base="eq", periodic=0, scale=1.3, stretch_in=1.0, period=1.0, stretch_out=0.9), False)
    assert abs(G["sigma"] - _fd(lambda t: obj(t, cut), sig)) < 1e-6 * abs(G["sigma"])
    for j in range(1, 4):
        def shifted(t, j=j):
            c = cut.copy()
            c[j] = t
            return obj(sig, c)
        assert abs(G["cutpoints"][j] - _fd(shifted, cut[j])) < 1e-6 * max(1.0, abs(G["cutpoints"][j]))
    assert G["cutpoints"][0] == 0 and G["cutpoints"][-1] == 0


def test_vb_gradient_matches_finite_differences():
    """oracle/gradients.py::vb_gradient (closed form of the implicit-function gradient of objective_VB) against
    central differences of the oracle's VB objective: kernel parameters, noise std and every finite cutpoint."""
    X, y, params, _ = ordinal_problem(3, 120, 2, 4, "eq")
    sig, cut = params[1]
    prior = make_prior(OK, "eq_scaled")
    th = (0.9, 1.3)

    def obj(t, s, c):
        gp = OA.VBGP((X, y), prior, OU.log_probit_likelihood, tolerance=1e-12, maxiter=5000)
        return gp.objective()((t, (s, c)))

    gp = OA.VBGP((X, y), prior, OU.log_probit_likelihood, tolerance=1e-12, maxiter=5000)
    w = gp.weight((th, (sig, cut)))
    G = OG.vb_gradient(prior(th)(X), X, y, w, (sig, cut),
                       dict(base="eq", periodic=0, scale=th[1], stretch_in=1.0, period=1.0, stretch_out=th[0]), False)
    assert abs(G["stretch_out"] - _fd(lambda t: obj((t, th[1]), sig, cut), th[0])) < 1e-6 * max(1, abs(G["stretch_out"]))
    assert abs(G["scale"] - _fd(lambda t: obj((th[0], t), sig, cut), th[1])) < 1e-6 * max(1, abs(G["scale"]))
    assert abs(G["sigma"] - _fd(lambda t: obj(th, t, cut), sig)) < 1e-6 * max(1, abs(G["sigma"]))
    for j in range(1, len(cut) - 1):
        def oc(t):
            c = cut.copy()
            c[j] = t
            return obj(th, sig, c)
        assert abs(G["cutpoints"][j] - _fd(oc, cut[j])) < 1e-6 * max(1, abs(G["cutpoints"][j]))
    assert G["cutpoints"][0] == 0 and G["cutpoints"][-1] == 0


def test_vb_gaussian_gradient_matches_finite_differences():
    Xr, yr, _, fam = regression_problem(3, 40)
    prior = make_prior(OK, fam)

    def obj(l, c, s):
        gp = OA.VBGP((Xr, yr), prior, OU.log_gaussian_likelihood, tolerance=1e-13, maxiter=5000)
        return gp.objective()(((l, c), (s,)))

    l, c, s = 0.3, 0.8, 0.8            # f_VB only contracts for sigma > 1/2 with this likelihood
    gp = OA.VBGP((Xr, yr), prior, OU.log_gaussian_likelihood, tolerance=1e-13, maxiter=5000)
    w = gp.weight(((l, c), (s,)))
    G = OG.vb_gradient(prior((l, c))(Xr), Xr, yr, w, (s,),
                       dict(base="eq", periodic=1, scale=c, stretch_in=1.0, period=0.5, stretch_out=l), True)
    assert abs(G["stretch_out"] - _fd(lambda t: obj(t, c, s), l)) < 1e-6 * abs(G["stretch_out"])
    assert abs(G["scale"] - _fd(lambda t: obj(l, t, s), c)) < 1e-6 * abs(G["scale"])
    assert abs(G["sigma"] - _fd(lambda t: obj(l, c, t), s)) < 1e-6 * abs(G["sigma"])
