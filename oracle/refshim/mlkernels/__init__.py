"""`mlkernels` (>=0.3.6) subset restated from the published package, over torch so that the
reference's autodiff paths can differentiate through the prior.  Pairwise calls return a `Dense`."""
import math

import torch

from _refshim_core import Dense, t


def _uprank(x):
    x = t(x).to(torch.float64)
    return x[:, None] if x.ndim == 1 else x


def _pw_dists2(a, b):
    if a.shape[-1] == 1:                                   # lab: (a - b^T)^2 for one-dimensional inputs
        return (a - b.T) ** 2
    na, nb = (a * a).sum(-1)[:, None], (b * b).sum(-1)[None, :]
    return na + nb - 2.0 * (a @ b.T)


class Kernel:
    def __call__(self, x, y=None):
        x = _uprank(x)
        y = x if y is None else _uprank(y)
        return Dense(self._pairwise(x, y))

    def elwise(self, x, y=None):
        x = _uprank(x)
        y = x if y is None else _uprank(y)
        return self._elwise(x, y)

    def stretch(self, s):
        return _Stretched(self, s)

    def periodic(self, p):
        return _Periodic(self, p)

    def __mul__(self, c):
        return _Scaled(self, c)

    __rmul__ = __mul__


class EQ(Kernel):
    def _pairwise(self, x, y):
        return torch.exp(-0.5 * _pw_dists2(x, y))

    def _elwise(self, x, y):
        return torch.exp(-0.5 * ((x - y) ** 2).sum(-1, keepdim=True))


class Exp(Kernel):
    def _pairwise(self, x, y):
        return torch.exp(-torch.sqrt(torch.clamp(_pw_dists2(x, y), min=1e-30)))

    def _elwise(self, x, y):
        return torch.exp(-torch.sqrt(torch.clamp(((x - y) ** 2).sum(-1, keepdim=True), min=1e-30)))


Matern12 = Exp


class _Scaled(Kernel):
    def __init__(self, k, c):
        self.k, self.c = k, c

    def _pairwise(self, x, y):
        return t(self.c) * self.k._pairwise(x, y)

    def _elwise(self, x, y):
        return t(self.c) * self.k._elwise(x, y)


class _Stretched(Kernel):
    def __init__(self, k, s):
        self.k, self.s = k, s

    def _pairwise(self, x, y):
        return self.k._pairwise(x / t(self.s), y / t(self.s))

    def _elwise(self, x, y):
        return self.k._elwise(x / t(self.s), y / t(self.s))


class _Periodic(Kernel):
    def __init__(self, k, p):
        self.k, self.p = k, p

    def _feat(self, x):
        a = 2.0 * math.pi * x / t(self.p)
        return torch.cat([torch.sin(a), torch.cos(a)], dim=-1)

    def _pairwise(self, x, y):
        return self.k._pairwise(self._feat(x), self._feat(y))

    def _elwise(self, x, y):
        return self.k._elwise(self._feat(x), self._feat(y))
