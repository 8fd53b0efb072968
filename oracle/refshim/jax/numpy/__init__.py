"""Minimal `jax.numpy` over torch float64 tensors (see ../../README.md)."""
import math

import torch

from _refshim_core import Array1D, t
from . import linalg  # noqa: F401

inf = math.inf
NINF = -math.inf            # removed from current JAX; the reference (jax ~0.4.1) still uses it
int32 = torch.int32


def asarray(x, dtype=None):
    return x if isinstance(x, Array1D) else t(x)


array = asarray


def shape(x):
    return tuple(t(x).shape)


def zeros(n):
    return torch.zeros(n, dtype=torch.float64)


def ones(n):
    return torch.ones(n, dtype=torch.float64)


def where(c, a, b):
    c = t(c)
    if not isinstance(a, torch.Tensor) and not isinstance(b, torch.Tensor):
        a = t(a)
    return torch.where(c, a, b)


def isinf(x):
    return torch.isinf(t(x))


def isnan(x):
    return torch.isnan(t(x))


def abs(x):
    return torch.abs(t(x))


def exp(x):
    return torch.exp(t(x))


def log(x):
    return torch.log(t(x))


def sqrt(x):
    return torch.sqrt(t(x))


def isclose(a, b, rtol=1e-5, atol=1e-8):
    return torch.isclose(t(a), t(b), rtol=rtol, atol=atol)


def append(a, v):
    return torch.cat([t(a).reshape(-1), t(v).reshape(-1).to(t(a).dtype)])


def insert(a, idx, v):
    a = t(a)
    return torch.cat([a[:idx], t(v).reshape(-1).to(a.dtype), a[idx:]])
