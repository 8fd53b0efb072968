"""Ordinal GP regression via approximate inference — the reference's examples/classification.py on probit_b200.

Same flow as the reference script (generate data -> build classifier -> approximate_posterior -> predict ->
predictive distributions -> metrics -> L-BFGS-B on the lengthscale -> repeat), minus the plotting.  Differences
forced by the environment: mlkernels -> probit_b200.kernels, varz.minimise_l_bfgs_b -> scipy L-BFGS-B on
log(lengthscale) driven by LaplaceGP.value_and_grad (the analytic evidence gradient), numpy RNG.
"""
import argparse
import os
import sys

import numpy as np
from scipy.optimize import minimize

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from probit_b200.approximators import LaplaceGP, VBGP  # noqa: E402
from probit_b200.datasets import device_latent_sampler, generate_ordinal_data  # noqa: E402
from probit_b200.kernels import EQ, Matern12  # noqa: E402
from probit_b200.utilities import check_cutpoints, log_probit_likelihood, probit_predictive_distributions  # noqa: E402


def calculate_metrics(y_test, predictive_distributions):
    """examples/classification.py:325-341 with the log-probability taken per test point."""
    y_pred = np.argmax(predictive_distributions, axis=1)
    mae = np.mean(np.abs(y_pred - y_test))
    zero_one = np.mean(y_pred != y_test)
    logp = np.sum(np.log(predictive_distributions[np.arange(len(y_test)), y_test]))
    print(f"incorrect={np.sum(y_pred != y_test)} correct={np.sum(y_pred == y_test)} mean_absolute_error={mae:.2f} "
          f"log_pred_probability={logp:.2f} mean_zero_one_error={zero_one:.2f}")
    return zero_one, mae, logp


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--method", default="Laplace", choices=["Laplace", "Variational Bayes"])
    ap.add_argument("--train-per-class", type=int, default=10)
    ap.add_argument("--test-per-class", type=int, default=100)
    args = ap.parse_args(argv)
    Approximator = LaplaceGP if args.method == "Laplace" else VBGP

    J, noise_variance, signal_variance, lengthscale = 3, 0.4, 1.0, 1.0
    kernel = signal_variance * Matern12().stretch(lengthscale)                 # classification.py:375
    n_total = J * (args.train_per_class + args.test_per_class)
    X_all, g_all, y_all, cutpoints = generate_ordinal_data(1, n_total, 1, J, noise_variance,
                                                           device_latent_sampler(kernel, 1e-6))
    rng = np.random.default_rng(2)
    train = np.concatenate([rng.permutation(np.flatnonzero(y_all == j))[: args.train_per_class] for j in range(J)])
    test = np.setdiff1d(np.arange(n_total), train)
    X, y, X_test, y_test = X_all[train], y_all[train], X_all[test], y_all[test]

    def prior(prior_parameters):                                                # classification.py:389-391
        return signal_variance * EQ().stretch(prior_parameters)

    cutpoints = check_cutpoints(cutpoints, J)
    print(f"cutpoints={cutpoints.tolist()}")
    classifier = Approximator(data=(X, y), prior=prior, log_likelihood=log_probit_likelihood, tolerance=1e-5)
    noise_std = float(np.sqrt(noise_variance))

    def evaluate(ls, tag):
        parameters = (ls, (noise_std, cutpoints))
        weight, precision = classifier.approximate_posterior(parameters)
        mean, variance = classifier.predict(X_test, parameters, weight, precision)
        dist = probit_predictive_distributions(parameters[1], mean, variance).cpu().numpy()
        print(f"\n{tag}: lengthscale={ls:.6f} objective={classifier.objective()(parameters):.4f}")
        return calculate_metrics(y_test, dist)

    before = evaluate(1.2, "Before optimization")                               # classification.py:415
    vg = classifier.value_and_grad()       # closed-form implicit gradient on the GPU (both approximators)

    def fun(phi):                                                               # varz optimises log(lengthscale)
        ls = float(np.exp(phi[0]))
        value, (g_prior, _) = vg((ls, (noise_std, cutpoints)))
        return value, np.array([g_prior * ls])

    res = minimize(fun, np.log([1.2]), jac=True, method="L-BFGS-B")
    ls_opt = float(np.exp(res.x[0]))
    after = evaluate(ls_opt, "After optimization")
    return before, after, res


if __name__ == "__main__":
    main()
