"""Synthetic input generators restating the reference examples' `generate_data` recipes.

Reference: examples/regression.py:28-80 (GP draw + Gaussian noise) and
examples/classification.py:181-322 (GP draw + noise, sort, split into J equal bins, cutpoints at
the mid-points between adjacent bins, b_0=-inf, b_J=+inf).  JAX's threefry stream is unavailable,
so draws come from numpy.random.default_rng(seed); the reference's
`random.shuffle` calls discard their result (a no-op) and are therefore omitted.

The latent draw f = chol(K + jitter I) z is delegated to a `latent_sampler(X, z) -> f` callable so
that the same recipe serves the no-GPU tests (a NumPy sampler supplied by the tests) and the
GPU bench (`device_latent_sampler`, which runs the product's own CUDA Gram + Cholesky).
"""
import numpy as np


def device_latent_sampler(kernel, jitter):
    """f = chol(K(X,X) + jitter I) z on the GPU through the product's Gram and potrf kernels."""
    def sample(X, z):
        import torch
        from . import linalg
        Xd = torch.as_tensor(X, dtype=torch.float64, device="cuda")
        A = linalg.gram(kernel.lower(), Xd, diag_add=jitter)
        linalg.potrf_(A)
        zd = torch.as_tensor(z, dtype=torch.float64, device="cuda")
        f = linalg.trmv_lower(A, zd)
        return f.cpu().numpy()
    return sample


def generate_regression_data(seed, N_train, D, noise_std, latent_sampler):
    """examples/regression.py:28-80 without the plotting grid: returns (X, y, f)."""
    rng = np.random.default_rng(seed)
    X = rng.uniform(0.0, 1.0, size=(N_train, D))
    z = rng.standard_normal(N_train)
    f = np.asarray(latent_sampler(X, z))
    y = f + noise_std * rng.standard_normal(N_train)
    return X, y, f


def generate_ordinal_data(seed, N, D, J, noise_variance, latent_sampler):
    """examples/classification.py:181-322 restricted to the training split.

    Returns (X, g, y, cutpoints) with y int64 in {0..J-1} and cutpoints of length J+1.
    The reference builds J classes of exactly N/J points; when J does not divide N (N=65536, J=5 in
    BASELINE configs[3]) the first N mod J classes get one extra point.
    """
    if N < J:
        raise ValueError("need at least one point per class")
    rng = np.random.default_rng(seed)
    X = rng.uniform(0.0, 1.0, size=(N, D))
    z = rng.standard_normal(N)
    f = np.asarray(latent_sampler(X, z))
    g = f + np.sqrt(noise_variance) * rng.standard_normal(N)       # classification.py:241-243
    idx = np.argsort(g, kind="stable")                              # :247
    g, X = g[idx], X[idx]
    counts = np.full(J, N // J, dtype=np.int64)
    counts[: N % J] += 1
    starts = np.concatenate([[0], np.cumsum(counts)])
    cutpoints = np.empty(J + 1)
    for j in range(1, J):                                           # :262-268
        cutpoints[j] = 0.5 * (g[starts[j]] + g[starts[j] - 1])
    cutpoints[0], cutpoints[-1] = -np.inf, np.inf                   # :269-270
    y = np.repeat(np.arange(J, dtype=np.int64), counts)
    perm = rng.permutation(N)     # de-sort the rows so that class order carries no structure
    return X[perm], g[perm], y[perm], cutpoints
