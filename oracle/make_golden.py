"""Generates tests/golden/<case>.npz from the oracle (TEST INFRASTRUCTURE).

These fixtures hold the INPUTS of each case (X, y, test points, parameters) and the oracle's outputs in its
LITERAL mode (dense Jacobian + LU Newton, LU predict); they freeze the oracle so that later edits to either the
oracle or the CUDA path are caught.  The reference's own outputs on the same inputs are produced separately by
oracle/make_reference_golden.py (reference source executed over the dependency shim) as tests/golden/ref_<case>.npz.
Existing files are left untouched unless --all is given.  Re-run:  python oracle/make_golden.py [--all]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import make_prior, ordinal_problem, regression_problem  # noqa: E402
from oracle import approximators as OA, kernels as OK, utilities as OU  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

CASES = {
    # BASELINE configs[0]: examples/regression.py — LaplaceGP, Gaussian likelihood, EQ().stretch(l).periodic(0.5), N=20
    "c1_regression_n20": dict(kind="regression", seed=0, N=20, cls="LaplaceGP"),
    # BASELINE configs[1]: examples/classification.py — ordinal LaplaceGP J=3, 10 train per class, EQ prior, l=1.2
    "c2_ordinal_j3_n30": dict(kind="ordinal", seed=1, N=30, D=1, J=3, family="eq", lengthscale=1.2, cls="LaplaceGP"),
    # scaled-down BASELINE configs[3]: ordinal J=5, D=4, Matern12
    "c4_small_ordinal_j5_n250": dict(kind="ordinal", seed=4, N=250, D=4, J=5, family="matern12", lengthscale=1.0,
                                     cls="LaplaceGP"),
    "vb_ordinal_j3_n120": dict(kind="ordinal", seed=7, N=120, D=2, J=3, family="eq", lengthscale=0.8, cls="VBGP"),
    # binary classification (J=2: both cutpoint intervals are half-infinite), D=2
    "binary_j2_n80": dict(kind="ordinal", seed=12, N=80, D=2, J=2, family="matern12", lengthscale=0.7, cls="LaplaceGP"),
    # VBGP with the Gaussian likelihood (regression through the variational path).  f_VB's iteration matrix is
    # (1 - 1/sigma) M^-1 K for this likelihood, so it only converges for sigma > 1/2: sigma = 0.8 here
    "vb_regression_n40": dict(kind="regression", seed=3, N=40, cls="VBGP", sigma=0.8),
    # the reference's optional "safe" derivative functions passed explicitly (examples/classification.py:406-407)
    "safe_ordinal_j3_n30": dict(kind="ordinal", seed=1, N=30, D=1, J=3, family="eq", lengthscale=1.2, cls="LaplaceGP",
                                safe=True),
}


def build(case):
    if case["kind"] == "regression":
        X, y, params, family = regression_problem(case["seed"], case["N"])
        if "sigma" in case:
            params = (params[0], (case["sigma"],))
        gaussian = True
        n_test, D = 100, 1
    else:
        X, y, params, family = ordinal_problem(case["seed"], case["N"], case["D"], case["J"], case["family"],
                                               case["lengthscale"])
        gaussian = False
        n_test, D = 64, case["D"]
    Xs = np.random.default_rng(case["seed"] + 100).uniform(-0.5, 1.5, size=(n_test, D))
    extra = dict(grad_log_likelihood=OU.grad_log_probit_likelihood,
                 hessian_log_likelihood=OU.hessian_log_probit_likelihood) if case.get("safe") else {}
    gp = getattr(OA, case["cls"])((X, y), make_prior(OK, family),
                                  OU.log_gaussian_likelihood if gaussian else OU.log_probit_likelihood, **extra)
    w, p = gp.approximate_posterior(params)
    iters = len(gp.trace)
    m, v = gp.predict(Xs, params, w, p)
    obj = gp.objective()(params)
    out = dict(X=X, y=y, Xs=Xs, weight=w, precision=p, mean=m, variance=v, objective=np.array(obj),
               iterations=np.array(iters), family=np.array(family), gaussian=np.array(gaussian),
               cls=np.array(case["cls"]), sigma=np.array(params[1][0]), safe=np.array(bool(case.get("safe"))))
    if gaussian:
        out["theta"] = np.array(params[0])
    else:
        out["theta"] = np.array(params[0])
        out["cutpoints"] = np.asarray(params[1][1])
        out["predictive"] = OU.probit_predictive_distributions(params[1], m, v)
    return out


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    for name, case in CASES.items():
        if os.path.exists(os.path.join(OUT, name + ".npz")) and "--all" not in sys.argv:
            continue                       # frozen fixtures are not rewritten unless asked
        np.savez(os.path.join(OUT, name + ".npz"), **build(case))
        print("wrote", name)
