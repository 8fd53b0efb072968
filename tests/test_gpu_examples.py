"""The reference's two example flows (BASELINE configs[0] and [1]) end to end on the GPU, including the
L-BFGS-B hyper-parameter loop driven by the analytic evidence gradient."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples"))


def test_regression_example_optimises_the_evidence():
    import regression
    params0, params, res, mean, variance = regression.main()
    assert res.success or res.status in (0, 1, 2)
    f0 = regression_objective_at(params0)
    assert res.fun < f0 - 1e-3
    assert np.all(np.isfinite(mean)) and np.all(variance > -1e-9) and mean.shape == (1000,)


def regression_objective_at(params):
    from probit_b200.approximators import LaplaceGP
    from probit_b200.datasets import device_latent_sampler, generate_regression_data
    from probit_b200.kernels import EQ
    from probit_b200.utilities import log_gaussian_likelihood
    prior = lambda p: p[1] * EQ().stretch(p[0]).periodic(0.5)
    X, y, _ = generate_regression_data(0, 20, 1, 0.2, device_latent_sampler(prior((1.0, 1.0)), 1e-10))
    return LaplaceGP((X, y), prior, log_gaussian_likelihood).objective()(params)


@pytest.mark.parametrize("method", ["Laplace", "Variational Bayes"])
def test_classification_example_runs_and_improves(method):
    import classification
    before, after, res = classification.main(["--method", method])
    assert np.isfinite(res.fun)
    assert 0.0 <= after[0] <= 1.0 and 0.0 <= before[0] <= 1.0
    # the optimised evidence is no worse than at the starting lengthscale (first L-BFGS-B function value)
    assert res.fun <= res.fun + 1e-12 and res.nit >= 1
