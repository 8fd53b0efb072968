"""GP regression — the reference's examples/regression.py on probit_b200 (LaplaceGP with a Gaussian likelihood).

Prior: signal_variance * EQ().stretch(lengthscale).periodic(0.5) (regression.py:120-123); N_train = 20; the three
hyper-parameters (lengthscale, signal variance, noise std) are optimised with L-BFGS-B on their logs using the
analytic evidence gradient (the reference: varz.minimise_l_bfgs_b with JAX autodiff, regression.py:157).
"""
import os
import sys

import numpy as np
from scipy.optimize import minimize

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from probit_b200.approximators import LaplaceGP as GP  # noqa: E402
from probit_b200.datasets import device_latent_sampler, generate_regression_data  # noqa: E402
from probit_b200.kernels import EQ  # noqa: E402
from probit_b200.utilities import log_gaussian_likelihood  # noqa: E402


def main(n_train=20, n_show=1000, seed=0):
    def prior(prior_parameters):
        lengthscale, signal_variance = prior_parameters
        return signal_variance * EQ().stretch(lengthscale).periodic(0.5)

    noise_std = 0.2
    X, y, _ = generate_regression_data(seed, n_train, 1, noise_std, device_latent_sampler(prior((1.0, 1.0)), 1e-10))
    X_show = np.linspace(-0.5, 1.5, n_show)[:, None]
    gp = GP(data=(X, y), prior=prior, log_likelihood=log_gaussian_likelihood)
    vg = gp.value_and_grad()

    def fun(phi):
        l, s2, sn = np.exp(phi)
        value, ((gl, gs2), (gsn,)) = vg(((float(l), float(s2)), (float(sn),)))
        return value, np.array([gl * l, gs2 * s2, gsn * sn])

    phi0 = np.log([0.10536897, 0.2787192, 0.6866876])            # the reference README's "before" parameters
    params0 = ((float(np.exp(phi0[0])), float(np.exp(phi0[1]))), (float(np.exp(phi0[2])),))
    print("Before optimization, params=", params0, "objective=", fun(phi0)[0])
    res = minimize(fun, phi0, jac=True, method="L-BFGS-B")
    l, s2, sn = (float(v) for v in np.exp(res.x))
    params = ((l, s2), (sn,))
    print("After optimization, params=", params, "objective=", res.fun)
    weight, precision = gp.approximate_posterior(params)
    mean, variance = gp.predict(X_show, params, weight, precision)
    return params0, params, res, mean.cpu().numpy(), variance.cpu().numpy()


if __name__ == "__main__":
    main()
