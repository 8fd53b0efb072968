"""Pins the oracle against tests/golden/ref_*.npz — outputs of the REFERENCE'S OWN SOURCE FILES executed in the
build container over the torch-backed dependency shim (oracle/make_reference_golden.py, oracle/refshim/README.md).
CPU only.  The GPU-side comparison against the same fixtures lives in tests/test_gpu_fit.py."""
import os

import numpy as np
import pytest

from helpers import make_prior, relerr
from oracle import approximators as OA, gradients as OG, kernels as OK, utilities as OU

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["c1_regression_n20", "c2_ordinal_j3_n30", "c4_small_ordinal_j5_n250", "vb_ordinal_j3_n120",
         "binary_j2_n80", "vb_regression_n40", "safe_ordinal_j3_n30"]


def _load(name):
    fx = np.load(os.path.join(GOLDEN, name + ".npz"))
    ref = np.load(os.path.join(GOLDEN, "ref_" + name + ".npz"))
    gaussian = bool(fx["gaussian"])
    if gaussian:
        params = (tuple(float(v) for v in fx["theta"]), (float(fx["sigma"]),))
    else:
        params = (float(fx["theta"]), (float(fx["sigma"]), fx["cutpoints"]))
    return fx, ref, params, gaussian


def _oracle(fx, params, gaussian, dist_mode):
    y = fx["y"] if gaussian else fx["y"].astype(np.int64)
    safe = "safe" in fx.files and bool(fx["safe"])
    extra = dict(grad_log_likelihood=OU.grad_log_probit_likelihood,
                 hessian_log_likelihood=OU.hessian_log_probit_likelihood) if safe else {}
    gp = getattr(OA, str(fx["cls"]))((fx["X"], y), make_prior(OK, str(fx["family"])),
                                     OU.log_gaussian_likelihood if gaussian else OU.log_probit_likelihood,
                                     dist_mode=dist_mode, **extra)
    return gp, y


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_source(name):
    """Same pairwise-distance form as lab (expansion for D > 1): agreement at rounding level."""
    fx, ref, params, gaussian = _load(name)
    gp, y = _oracle(fx, params, gaussian, "expand")
    w, p = gp.approximate_posterior(params)
    m, v = gp.predict(fx["Xs"], params, w, p)
    cov = gp.predict_covariance(fx["Xs"], params, w, p)
    obj = gp.objective()(params)
    tol = 1e-11
    assert relerr(w, ref["weight"]) < tol
    assert relerr(p, ref["precision"]) < tol
    assert relerr(m, ref["mean"]) < tol
    assert relerr(v, ref["variance"]) < tol
    assert relerr(cov, ref["covariance"]) < tol
    assert abs(obj - ref["objective"]) < tol * abs(ref["objective"])
    assert abs(float(ref["vg_value"]) - float(ref["objective"])) < 1e-12 * abs(ref["objective"]) if "vg_value" in ref else True


@pytest.mark.parametrize("name", CASES)
def test_oracle_direct_distance_mode_is_within_the_parity_tolerance(name):
    """The product computes ||a-b||^2 by direct differences; north_star's 1e-8 covers the reference's expansion."""
    fx, ref, params, gaussian = _load(name)
    gp, y = _oracle(fx, params, gaussian, "direct")
    w, p = gp.approximate_posterior(params)
    m, v = gp.predict(fx["Xs"], params, w, p)
    assert relerr(w, ref["weight"]) < 2e-8
    assert relerr(gp.predict_covariance(fx["Xs"], params, w, p), ref["covariance"]) < 2e-8
    assert relerr(m, ref["mean"]) < 1e-8
    assert relerr(v, ref["variance"]) < 1e-8
    assert abs(gp.objective()(params) - ref["objective"]) < 1e-8 * abs(ref["objective"])


@pytest.mark.parametrize("name", ["c2_ordinal_j3_n30", "c4_small_ordinal_j5_n250", "vb_ordinal_j3_n120", "binary_j2_n80"])
def test_likelihood_derivatives_match_reference_autodiff(name):
    """The oracle's hand-derived g, h against jax.grad-style autodiff of the reference's log-likelihood."""
    fx, ref, params, gaussian = _load(name)
    f, y = ref["posterior_mean"], fx["y"].astype(np.int64)
    lp = params[1]
    assert np.allclose(OU.log_probit_likelihood(f, y, lp), ref["ll"], rtol=1e-13, atol=1e-13)
    u = OU.probit_likelihood(f, y, lp) + 1e-10
    scale = 1e-13 + 4e-16 / u            # erf implementations differ by an ulp; every output divides by u
    assert np.all(np.abs(OU.grad_log_probit_likelihood_autodiff(f, y, lp) - ref["grad_ll"]) <= scale * (1 + np.abs(ref["grad_ll"])) * 10)
    assert np.all(np.abs(OU.hessian_log_probit_likelihood_autodiff(f, y, lp) - ref["hess_ll"]) <= scale * (1 + np.abs(ref["hess_ll"])) * 100)
    for single, tag in ((True, "single"), (False, "double")):
        assert np.allclose(OU.grad_log_probit_likelihood(f, y, lp, single), ref["safe_grad_" + tag], rtol=1e-11, atol=1e-12)
        assert np.allclose(OU.hessian_log_probit_likelihood(f, y, lp, single), ref["safe_hess_" + tag], rtol=1e-10, atol=1e-11)
    assert np.allclose(OU.probit_predictive_distributions(lp, ref["mean"], ref["variance"]), ref["predictive"],
                       rtol=0, atol=1e-15)


SPEC = {"c1_regression_n20": lambda th: dict(base="eq", periodic=1, scale=th[1], stretch_in=1.0, period=0.5, stretch_out=th[0]),
        "c2_ordinal_j3_n30": lambda l: dict(base="eq", periodic=0, scale=1.0, stretch_in=1.0, period=1.0, stretch_out=l),
        "c4_small_ordinal_j5_n250": lambda l: dict(base="exp", periodic=0, scale=1.0, stretch_in=1.0, period=1.0, stretch_out=l),
        "binary_j2_n80": lambda l: dict(base="exp", periodic=0, scale=1.0, stretch_in=1.0, period=1.0, stretch_out=l),
        "vb_ordinal_j3_n120": lambda l: dict(base="eq", periodic=0, scale=1.0, stretch_in=1.0, period=1.0, stretch_out=l),
        "vb_regression_n40": lambda th: dict(base="eq", periodic=1, scale=th[1], stretch_in=1.0, period=0.5, stretch_out=th[0])}


@pytest.mark.parametrize("name", sorted(SPEC))
def test_closed_form_gradient_matches_reference_implicit_differentiation(name):
    """oracle/gradients.py (closed form at the fixed point) against the reference's value_and_grad, which
    differentiates objective_LA through its custom-VJP fixed-point layer (solvers.py:28-64).  The reference's
    adjoint is itself a fixed-point solve to tolerance 1e-5, so agreement is to ~1e-6, not to rounding."""
    fx, ref, params, gaussian = _load(name)
    y = fx["y"] if gaussian else fx["y"].astype(np.int64)
    vb = str(fx["cls"]) == "VBGP"
    gp = getattr(OA, str(fx["cls"]))((fx["X"], y), make_prior(OK, str(fx["family"])),
                                     OU.log_gaussian_likelihood if gaussian else OU.log_probit_likelihood,
                                     tolerance=1e-12 if vb else 1e-10, maxiter=5000)
    w = gp.weight(params)
    K = make_prior(OK, str(fx["family"]))(params[0])(fx["X"])
    G = (OG.vb_gradient if vb else OG.laplace_gradient)(K, fx["X"], y, w, params[1], SPEC[name](params[0]), gaussian)
    vg_theta = np.atleast_1d(ref["vg_theta"])
    assert abs(G["stretch_out"] - vg_theta[0]) < 2e-5 * max(1.0, abs(vg_theta[0]))
    if gaussian:
        assert abs(G["scale"] - vg_theta[1]) < 2e-5 * max(1.0, abs(vg_theta[1]))
    assert abs(G["sigma"] - float(ref["vg_sigma"])) < 2e-5 * max(1.0, abs(float(ref["vg_sigma"])))
    if not gaussian:
        inner = slice(1, -1)
        assert np.allclose(G["cutpoints"][inner], ref["vg_cutpoints"][inner], rtol=2e-5, atol=2e-5)
