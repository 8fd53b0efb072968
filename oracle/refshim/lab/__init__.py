"""Minimal `lab` (PyPI `backends`) API over torch (see ../README.md).  Semantics restated from the published
package: epsilon = 1e-12; cholesky regularises a matrix.Dense argument with epsilon."""
import math

import torch

from _refshim_core import Dense, t

pi = math.pi
epsilon = 1e-12


def sqrt(x):
    return math.sqrt(x) if isinstance(x, (int, float)) else torch.sqrt(t(x))


def log(x):
    return torch.log(t(x))


def dense(x):
    return t(x)


def shape(x):
    return tuple(t(x).shape)


def size(x):
    return t(x).numel()


def sum(x, axis=None):
    return torch.sum(t(x)) if axis is None else torch.sum(t(x), dim=axis)


def diag(x):
    x = t(x)
    return torch.diag(x)           # vector -> matrix, matrix -> vector


def eye(n):
    return torch.eye(n, dtype=torch.float64)


def ones(*shape):
    return torch.ones(*shape, dtype=torch.float64)


def flatten(x):
    return t(x).reshape(-1)


def einsum(eq, *xs):
    return torch.einsum(eq.replace(" ", ""), *[t(x) for x in xs])


def cholesky(a):
    if isinstance(a, Dense):       # matrix.Dense path: B.cholesky(B.reg(mat)) — Laplace.py:24 reaches this one
        m = a.mat
        return torch.linalg.cholesky(m + epsilon * torch.eye(m.shape[0], dtype=m.dtype))
    return torch.linalg.cholesky(t(a))


def triangular_solve(a, b, lower_a=True):
    b = t(b)
    vec = b.ndim == 1
    x = torch.linalg.solve_triangular(t(a), b[:, None] if vec else b, upper=not lower_a)
    return x[:, 0] if vec else x


def cholesky_solve(L, b):
    return triangular_solve(t(L).T, triangular_solve(L, b, lower_a=True), lower_a=False)


def solve(a, b):
    return torch.linalg.solve(t(a), t(b))
