"""Shared helpers for the shim packages (TEST INFRASTRUCTURE): everything is a float64 CPU torch tensor."""
import numpy as np
import torch
from torch.utils import _pytree as pytree

torch.set_default_dtype(torch.float64)


class Dense:
    """Stand-in for `matrix.Dense`, the type mlkernels' pairwise calls return:
    Dense + array stays Dense, and lab's cholesky regularises a Dense (but not a raw array)."""

    def __init__(self, mat):
        self.mat = mat

    def __add__(self, other):
        return Dense(self.mat + t(other))

    __radd__ = __add__

    def __mul__(self, other):
        return Dense(self.mat * t(other))

    __rmul__ = __mul__

    def __matmul__(self, other):
        return self.mat @ t(other)

    def __rmatmul__(self, other):
        return t(other) @ self.mat

    @property
    def T(self):
        return Dense(self.mat.T)

    @property
    def shape(self):
        return self.mat.shape


class Array1D:
    """A vector that can be indexed by a traced integer: `cutpoints[y]` inside vmap (utilities.py:50-51,105-106).
    torch turns a 0-d integer tensor index into `.item()`, which vmap forbids; JAX gathers.  This wrapper gathers."""

    def __init__(self, vec):
        self.vec = vec

    def __getitem__(self, idx):
        if isinstance(idx, torch.Tensor):
            return self.vec.index_select(0, idx.reshape(-1)).reshape(idx.shape)
        return self.vec[idx]

    def __len__(self):
        return self.vec.shape[0]

    @property
    def shape(self):
        return self.vec.shape


pytree.register_pytree_node(Array1D, lambda a: ([a.vec], None), lambda leaves, ctx: Array1D(leaves[0]))


def t(x):
    """Anything array-like -> torch tensor (float64 for floats, int64 for ints); Dense / Array1D unwrap."""
    if isinstance(x, Dense):
        return x.mat
    if isinstance(x, Array1D):
        return x.vec
    if isinstance(x, torch.Tensor):
        return x
    if isinstance(x, np.ndarray):
        if x.dtype.kind == "f":
            return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64))
        return torch.from_numpy(np.ascontiguousarray(x))
    if isinstance(x, (list, tuple)):
        if any(isinstance(e, torch.Tensor) for e in x):
            return torch.stack([t(e).to(torch.float64) for e in x])
        return t(np.asarray(x))
    if isinstance(x, bool):
        return torch.tensor(x)
    if isinstance(x, int):
        return torch.tensor(x, dtype=torch.int64)
    return torch.tensor(float(x), dtype=torch.float64)
