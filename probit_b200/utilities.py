"""Likelihood callables and helpers with the names of probit/utilities.py.

`log_probit_likelihood` and `log_gaussian_likelihood` are the callables a user passes as
`log_likelihood=` (examples/classification.py:405, examples/regression.py:129).  The approximators
recognise them by identity and run the fused CUDA likelihood kernel; called directly they evaluate
on the GPU too (vectorised over f, y).  `grad_log_probit_likelihood` /
`hessian_log_probit_likelihood` select the optional series-expansion ("safe") mode
(probit/utilities.py:151-192).  Arbitrary Python likelihood callables cannot be compiled to CUDA and
raise NotImplementedError in the approximators — there is no CPU fallback.
"""
import ctypes as C
import math

import torch

from . import _lib
from .linalg import _dev, _ptr, _stream

LIKELIHOOD_EPS = 1e-10          # probit/utilities.py:57
BOUNDS = {"single": [1.3, 1.8, 2.3], "double": [2.3, 3.6, 4.8]}   # probit/utilities.py:15


class CutpointValueError(Exception):
    """probit/utilities.py:323-342."""

    def __init__(self, cutpoint):
        super().__init__(f"The cutpoint list or array must be in ascending order,  {cutpoint} was given.")


class InvalidKernel(Exception):
    """probit/utilities.py:345-361."""

    def __init__(self, kernel):
        super().__init__(f"{kernel} is not an instance of a supported kernel specification (probit_b200.kernels.Kernel)")


def h(x):
    """Series polynomial of probit/utilities.py:37-44 (host scalar helper; the device copy lives in likelihood.cuh)."""
    x = float(x)
    if x == 0.0:
        return math.nan
    if math.isinf(x):
        return 0.0
    return -1 * x**-2 + 5 / 2 * x**-4 - 37 / 3 * x**-6


def make_likelihood_spec(kind, likelihood_parameters, eps=LIKELIHOOD_EPS, single_precision=True):
    """Build the pb_likelihood_spec; returns (spec, keepalive) — keepalive holds the device cutpoints."""
    sigma = float(likelihood_parameters[0])
    if kind == _lib.PB_LIK_GAUSSIAN:
        return _lib.LikelihoodSpec(kind, 0, sigma, float(eps), None, 0, 0), None
    cut = _dev(likelihood_parameters[1]).reshape(-1)
    J = cut.numel() - 1
    return _lib.LikelihoodSpec(kind, J, sigma, float(eps), cut.data_ptr(), int(bool(single_precision)), 0), cut


def _labels(kind, y):
    return _dev(y, torch.float64 if kind == _lib.PB_LIK_GAUSSIAN else torch.int64).reshape(-1)


def evaluate_likelihood(kind, f, y, likelihood_parameters, want=("ll", "g", "h"), eps=LIKELIHOOD_EPS,
                        single_precision=True):
    """Run the fused likelihood kernel; returns a dict of the requested (N,) tensors."""
    lib = _lib.load()
    f = _dev(f).reshape(-1)
    y = _labels(kind, y)
    spec, keep = make_likelihood_spec(kind, likelihood_parameters, eps, single_precision)
    n = y.numel()
    batch = f.numel() // max(n, 1)
    outs = {k: (torch.empty_like(f) if k in want else None) for k in ("ll", "g", "h", "d3")}
    _lib.check(lib.pb_likelihood(_stream(), C.byref(spec), _ptr(f), _ptr(y), n, batch, _ptr(outs["ll"]),
                                 _ptr(outs["g"]), _ptr(outs["h"]), _ptr(outs["d3"])))
    del keep
    return {k: v for k, v in outs.items() if v is not None}


def log_probit_likelihood(f, y, likelihood_parameters):
    """probit/utilities.py:56-57, vectorised over (f, y)."""
    return evaluate_likelihood(_lib.PB_LIK_ORDINAL_PROBIT, f, y, likelihood_parameters, ("ll",))["ll"]


def log_gaussian_likelihood(f, y, likelihood_parameters):
    """probit/utilities.py:60-61, vectorised over (f, y)."""
    return evaluate_likelihood(_lib.PB_LIK_GAUSSIAN, f, y, likelihood_parameters, ("ll",))["ll"]


def grad_log_probit_likelihood(f, y, likelihood_parameters, single_precision=True):
    """probit/utilities.py:151-169 (series-expansion mode)."""
    return evaluate_likelihood(_lib.PB_LIK_ORDINAL_PROBIT_SAFE, f, y, likelihood_parameters, ("g",),
                               single_precision=single_precision)["g"]


def hessian_log_probit_likelihood(f, y, likelihood_parameters, single_precision=True):
    """probit/utilities.py:172-192 (series-expansion mode)."""
    return evaluate_likelihood(_lib.PB_LIK_ORDINAL_PROBIT_SAFE, f, y, likelihood_parameters, ("h",),
                               single_precision=single_precision)["h"]


def probit_predictive_distributions(likelihood_parameters, posterior_mean, posterior_variance):
    """probit/utilities.py:232-249 — (N_test, J) class probabilities, one fused CUDA pass."""
    lib = _lib.load()
    mean = _dev(posterior_mean).reshape(-1)
    var = _dev(posterior_variance).reshape(-1)
    spec, keep = make_likelihood_spec(_lib.PB_LIK_ORDINAL_PROBIT, likelihood_parameters)
    out = torch.empty((mean.numel(), spec.J), dtype=torch.float64, device="cuda")
    _lib.check(lib.pb_predictive_distributions(_stream(), C.byref(spec), _ptr(mean), _ptr(var), mean.numel(),
                                               _ptr(out)))
    del keep
    return out


def check_cutpoints(cutpoints, J):
    """probit/utilities.py:263-320 — host-side validation; same accepted shapes and error types.

    (J-1,) -> both infinities added; (J,) -> the missing infinity added; (J+1,) -> checked.
    ValueError for a wrong shape or missing infinity, CutpointValueError if not ascending.
    """
    inf = math.inf
    c = torch.as_tensor(cutpoints, dtype=torch.float64).reshape(-1).cpu()
    lo, hi = torch.tensor([-inf], dtype=torch.float64), torch.tensor([inf], dtype=torch.float64)
    if c.numel() == J - 1:
        c = torch.cat([lo, c, hi])
    elif c.numel() == J:
        if c[-1] != inf:
            if c[0] != -inf:
                raise ValueError(
                    "Either the largest cutpoint parameter b_J is not positive infinity, or the smallest "
                    "cutpoint parameter must b_0 is not negative infinity."
                    "(got {}, expected {})".format([float(c[0]), float(c[-1])], [inf, -inf]))
            c = torch.cat([c, hi])
        else:
            c = torch.cat([lo, c])
    elif c.numel() == J + 1:
        if c[0] != -inf:
            raise ValueError("The smallest cutpoint parameter b_0 must be negative infinity "
                             "(got {}, expected {})".format(float(c[0]), -inf))
        if c[-1] != inf:
            raise ValueError("The largest cutpoint parameter b_J must be positive infinity "
                             "(got {}, expected {})".format(float(c[-1]), inf))
    else:
        raise ValueError("Could not recognise cutpoints shape. (shape was {})".format(tuple(c.shape)))
    assert c[0] == -inf and c[-1] == inf and c.numel() == J + 1
    if not all(bool(c[i] <= c[i + 1]) for i in range(J)):
        raise CutpointValueError(cutpoints)
    return c
