"""Generates tests/golden/ref_*.npz by EXECUTING THE REFERENCE'S OWN SOURCE (TEST INFRASTRUCTURE).

/root/reference/probit/{approximators,utilities,implicit/*}.py are imported unmodified; their four third-party
dependencies (jax, jaxopt, lab, mlkernels — not installable here) resolve to the torch-backed API shim under
oracle/refshim/ (see its README for exactly what is substituted and what is restated from memory).  The inputs
are the ones already frozen in tests/golden/{c1,c2,c4_small,vb}_*.npz, so reference-on-shim, oracle and CUDA
outputs are compared on identical data.

Runs only in the build container (needs /root/reference); the fixtures it writes travel with the repo.
Re-run:  python oracle/make_reference_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = os.environ.get("PROBIT_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(ROOT, "oracle", "refshim"))
sys.path.insert(0, REFERENCE)

import torch  # noqa: E402
import jax  # noqa: E402  (the shim)
import jax.numpy as jnp  # noqa: E402
from mlkernels import EQ, Matern12  # noqa: E402  (the shim)
from _refshim_core import Array1D  # noqa: E402
import probit.approximators as RA  # noqa: E402  (the reference)
import probit.utilities as RU  # noqa: E402  (the reference)

assert os.path.realpath(RA.__file__).startswith(os.path.realpath(REFERENCE)), RA.__file__
GOLDEN = os.path.join(ROOT, "tests", "golden")


def np64(x):
    return np.asarray(x.detach() if isinstance(x, torch.Tensor) else x, dtype=np.float64)


def prior_for(family):
    # the reference examples' priors: classification.py:375, :389-391, regression.py:120-123
    if family == "matern12":
        return lambda l: 1.0 * Matern12().stretch(l)
    if family == "eq":
        return lambda l: 1.0 * EQ().stretch(l)
    if family == "eq_periodic":
        return lambda th: th[1] * EQ().stretch(th[0]).periodic(0.5)
    raise KeyError(family)


def run_case(name):
    fx = np.load(os.path.join(GOLDEN, name + ".npz"))
    family, gaussian, cls = str(fx["family"]), bool(fx["gaussian"]), str(fx["cls"])
    X, Xs = torch.from_numpy(fx["X"]), torch.from_numpy(fx["Xs"])
    if gaussian:
        y = torch.from_numpy(fx["y"])
        params = ((torch.tensor(float(fx["theta"][0])), torch.tensor(float(fx["theta"][1]))),
                  (torch.tensor(float(fx["sigma"])),))
        ll = RU.log_gaussian_likelihood
    else:
        y = torch.from_numpy(fx["y"].astype(np.int64))
        params = (torch.tensor(float(fx["theta"])), (torch.tensor(float(fx["sigma"])), Array1D(torch.from_numpy(fx["cutpoints"]))))
        ll = RU.log_probit_likelihood
    safe = "safe" in fx.files and bool(fx["safe"])
    extra = dict(grad_log_likelihood=RU.grad_log_probit_likelihood,
                 hessian_log_likelihood=RU.hessian_log_probit_likelihood) if safe else {}
    gp = getattr(RA, cls)(data=(X, y), prior=prior_for(family), log_likelihood=ll, **extra)
    weight, precision = gp.approximate_posterior(params)
    mean, variance = gp.predict(Xs, params, weight, precision)
    cov = gp.predict_covariance(Xs, params, weight, precision)
    objective = gp.objective()(params)
    out = dict(weight=np64(weight), precision=np64(precision), mean=np64(mean), variance=np64(variance),
               covariance=np64(cov), objective=np64(objective))
    # per-datum likelihood derivatives exactly as the approximator evaluates them (autodiff of the log-likelihood)
    f = gp.precision(weight, params)[1]
    out["posterior_mean"] = np64(f)
    out["ll"] = np64(gp.log_likelihood(f, y, params[1]))
    out["grad_ll"] = np64(gp.grad_log_likelihood(f, y, params[1]))
    out["hess_ll"] = np64(gp.hessian_log_likelihood(f, y, params[1]))
    if not safe:
        # implicit-function gradient through the reference's custom VJP (solvers.py:28-64)
        value, grads = gp.value_and_grad()(params)
        out["vg_value"] = np64(value)
        if gaussian:
            out["vg_theta"] = np.array([float(grads[0][0]), float(grads[0][1])])
            out["vg_sigma"] = np64(grads[1][0])
        else:
            out["vg_theta"] = np64(grads[0])
            out["vg_sigma"] = np64(grads[1][0])
            out["vg_cutpoints"] = np64(grads[1][1].vec)
    if not gaussian:
        out["predictive"] = np64(RU.probit_predictive_distributions(params[1], mean, variance))
        for single in (True, False):
            tag = "single" if single else "double"
            out["safe_grad_" + tag] = np64(RU.grad_log_probit_likelihood(f, y, params[1], single))
            out["safe_hess_" + tag] = np64(RU.hessian_log_probit_likelihood(f, y, params[1], single))
    return out


def reference_unit_checks():
    """The reference's only test (probit/test/test_implicit.py:11-22), run through the shim."""
    import probit.test.test_implicit as T
    T.test_values_and_gradient_of_series_expansion()
    c = RU.check_cutpoints([-0.5, 0.5], 3)
    assert float(c[0]) == -np.inf and float(c[-1]) == np.inf and len(c) == 4


if __name__ == "__main__":
    import warnings
    warnings.simplefilter("ignore")
    reference_unit_checks()
    print("reference test_implicit.py passes on the shim")
    for name in ["c1_regression_n20", "c2_ordinal_j3_n30", "c4_small_ordinal_j5_n250", "vb_ordinal_j3_n120",
                 "binary_j2_n80", "vb_regression_n40", "safe_ordinal_j3_n30"]:
        out = run_case(name)
        np.savez(os.path.join(GOLDEN, "ref_" + name + ".npz"), **out)
        fx = np.load(os.path.join(GOLDEN, name + ".npz"))
        rel = lambda a, b: float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300))
        print(name, "| vs oracle fixture: weight %.2e precision %.2e mean %.2e variance %.2e objective %.2e" % (
            rel(out["weight"], fx["weight"]), rel(out["precision"], fx["precision"]), rel(out["mean"], fx["mean"]),
            rel(out["variance"], fx["variance"]), rel(out["objective"], fx["objective"])))
