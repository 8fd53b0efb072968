"""`jaxopt.FixedPointIteration` (>=0.5.5) restated from its published algorithm:
state.error starts at inf; each update sets params <- f(params) and error <- ||f(params) - params||_2; the loop
runs while error > tol and iter_num < maxiter (default 100)."""
import math
from collections import namedtuple

import torch

State = namedtuple("FixedPointState", "iter_num error")


class FixedPointIteration:
    def __init__(self, fixed_point_fun, maxiter=100, tol=1e-5, **unused):
        self.f, self.maxiter, self.tol = fixed_point_fun, maxiter, tol

    def run(self, init_params, *args, **kwargs):
        params, error, it = init_params, math.inf, 0
        while error > self.tol and it < self.maxiter:
            nxt = self.f(params, *args, **kwargs)
            error = float(torch.linalg.vector_norm(nxt - params))
            params, it = nxt, it + 1
        return params, State(it, error)
