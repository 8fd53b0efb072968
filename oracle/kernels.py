"""Oracle restatement of the mlkernels subset the reference's examples use.

Reference call sites: `signal_variance * EQ().stretch(l).periodic(0.5)` (examples/regression.py:120-123),
`signal_variance * EQ().stretch(l)` (examples/classification.py:389-391),
`signal_variance * Matern12().stretch(l)` (examples/classification.py:375).
mlkernels (>=0.3.6, absent here) published semantics restated:
  EQ       k(x,y) = exp(-0.5 * ||x-y||^2)
  Exp      k(x,y) = exp(-||x-y||)            (Matern12 is an alias of Exp)
  stretch  k(x/l, y/l)
  periodic k(u(x), u(y)),  u(x) = [sin(2 pi x / p), cos(2 pi x / p)]  (features concatenated)
  c * k    scalar multiple of the matrix
  k(x)     = k(x, x);  1-D inputs are up-ranked to (N, 1);  k.elwise(x, y) -> (N, 1)
`dist_mode`:
  "direct" : ||a-b||^2 = sum_d (a_d - b_d)^2
  "expand" : lab's B.pw_dists2 for D>1: ||a||^2 + ||b||^2 - 2 a.b ; B.pw_dists = sqrt(max(.,1e-30))
             (FROM MEMORY) — kept so the reference's own noise floor can be measured.
"""
import numpy as np

DIST_MODE = "direct"


def _uprank(x):
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 0:
        x = x[None, None]
    elif x.ndim == 1:
        x = x[:, None]
    return x


def pw_dists2(a, b, mode=None):
    mode = mode or DIST_MODE
    if mode == "direct" or a.shape[1] == 1:
        d2 = np.zeros((a.shape[0], b.shape[0]))
        for d in range(a.shape[1]):
            diff = a[:, d][:, None] - b[:, d][None, :]
            d2 += diff * diff
        return d2
    na = np.sum(a * a, axis=1)[:, None]
    nb = np.sum(b * b, axis=1)[None, :]
    return na + nb - 2.0 * (a @ b.T)


def ew_dists2(a, b):
    return np.sum((a - b) ** 2, axis=1)[:, None]


class Kernel:
    """Fluent kernel-expression node (mirrors the mlkernels API subset)."""

    def stretch(self, l):
        return Stretched(self, float(l))

    def periodic(self, p=1.0):
        return Periodic(self, float(p))

    def __rmul__(self, c):
        return Scaled(self, float(c))

    def __mul__(self, c):
        return Scaled(self, float(c))

    def __call__(self, x, y=None, dist_mode=None):
        x = _uprank(x)
        y = x if y is None else _uprank(y)
        return self._pairwise(x, y, dist_mode)

    def elwise(self, x, y=None):
        x = _uprank(x)
        y = x if y is None else _uprank(y)
        return self._elwise(x, y)


class EQ(Kernel):
    def _pairwise(self, x, y, mode):
        return np.exp(-0.5 * pw_dists2(x, y, mode))

    def _elwise(self, x, y):
        return np.exp(-0.5 * ew_dists2(x, y))


class Exp(Kernel):
    def _pairwise(self, x, y, mode):
        d2 = pw_dists2(x, y, mode)
        if (mode or DIST_MODE) == "expand" and x.shape[1] > 1:
            return np.exp(-np.sqrt(np.maximum(d2, 1e-30)))
        return np.exp(-np.sqrt(d2))

    def _elwise(self, x, y):
        return np.exp(-np.sqrt(ew_dists2(x, y)))


Matern12 = Exp


class Stretched(Kernel):
    def __init__(self, k, l):
        self.k, self.l = k, l

    def _pairwise(self, x, y, mode):
        return self.k._pairwise(x / self.l, y / self.l, mode)

    def _elwise(self, x, y):
        return self.k._elwise(x / self.l, y / self.l)


def _periodic_features(x, p):
    a = 2.0 * np.pi * x / p
    return np.concatenate([np.sin(a), np.cos(a)], axis=1)


class Periodic(Kernel):
    def __init__(self, k, p):
        self.k, self.p = k, p

    def _pairwise(self, x, y, mode):
        return self.k._pairwise(_periodic_features(x, self.p), _periodic_features(y, self.p), mode)

    def _elwise(self, x, y):
        return self.k._elwise(_periodic_features(x, self.p), _periodic_features(y, self.p))


class Scaled(Kernel):
    def __init__(self, k, c):
        self.k, self.c = k, c

    def _pairwise(self, x, y, mode):
        return self.c * self.k._pairwise(x, y, mode)

    def _elwise(self, x, y):
        return self.c * self.k._elwise(x, y)
