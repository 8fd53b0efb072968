"""Minimal `jax` API over torch.func — only what /root/reference/probit uses (see ../README.md)."""
import torch
from torch.utils import _pytree as pytree

from _refshim_core import t
from . import lax, numpy  # noqa: F401


class _Config:
    def update(self, key, value):       # jax.config.update("jax_enable_x64", True): the shim is always float64
        pass


config = _Config()


def _arrays(args):
    """jit / grad trace Python scalars as arrays; mirror that for top-level scalar arguments."""
    return tuple(t(a) if isinstance(a, (int, float)) and not isinstance(a, bool) else a for a in args)


def jit(fun, *a, **k):
    return lambda *args, **kwargs: fun(*_arrays(args), **kwargs)


def grad(fun, argnums=0):
    g = torch.func.grad(fun, argnums=argnums)
    return lambda *args, **kwargs: g(*_arrays(args), **kwargs)


def vmap(fun, in_axes=0, out_axes=0):
    return torch.func.vmap(fun, in_dims=in_axes, out_dims=out_axes)


def jacobian(fun, argnums=0):           # jax.jacobian is jacrev
    return torch.func.jacrev(fun, argnums=argnums)


def vjp(fun, *primals):
    return torch.func.vjp(fun, *primals)


def value_and_grad(fun, argnums=0):
    """Reverse mode through ordinary torch autograd (outermost level), so that `custom_vjp` below can be an
    old-style autograd.Function whose forward is free to run torch.func transforms."""
    def wrapped(*args):
        leaves, spec = pytree.tree_flatten(args[argnums])
        leaves = [t(x).detach().clone().to(torch.float64).requires_grad_(True) for x in leaves]
        new_args = list(args)
        new_args[argnums] = pytree.tree_unflatten(leaves, spec)
        with torch.enable_grad():
            value = fun(*new_args)
            grads = torch.autograd.grad(value, leaves, allow_unused=True)
        grads = [g if g is not None else torch.zeros_like(x) for g, x in zip(grads, leaves)]
        return value.detach(), pytree.tree_unflatten(grads, spec)
    return wrapped


class custom_vjp:
    """jax.custom_vjp(fun, nondiff_argnums) with .defvjp(fwd, bwd); bwd(*nondiff, residuals, cotangent)."""

    def __init__(self, fun, nondiff_argnums=()):
        self.fun, self.nondiff = fun, tuple(nondiff_argnums)
        self.fwd = self.bwd = None

    def defvjp(self, fwd, bwd):
        self.fwd, self.bwd = fwd, bwd

    def __call__(self, *args):
        diff_idx = [i for i in range(len(args)) if i not in self.nondiff]
        leaves, spec = pytree.tree_flatten([args[i] for i in diff_idx])
        is_tensor = [isinstance(x, torch.Tensor) for x in leaves]
        tensors = [x for x in leaves if isinstance(x, torch.Tensor)]
        if not (torch.is_grad_enabled() and any(x.requires_grad for x in tensors)):
            return self.fun(*args)
        outer = self

        class Fn(torch.autograd.Function):
            @staticmethod
            def forward(ctx, *tens):
                it = iter(tens)
                full = [next(it) if flag else x for flag, x in zip(is_tensor, leaves)]
                diff_args = pytree.tree_unflatten(full, spec)
                call = list(args)
                for i, a in zip(diff_idx, diff_args):
                    call[i] = a
                out, res = outer.fwd(*call)
                ctx.res = res
                return out

            @staticmethod
            def backward(ctx, ct):
                nd = [args[i] for i in outer.nondiff]
                cts = outer.bwd(*nd, ctx.res, ct)          # one entry per differentiable argument
                flat = []
                for a, c in zip([args[i] for i in diff_idx], cts):
                    a_leaves = pytree.tree_flatten(a)[0]
                    c_leaves = pytree.tree_flatten(c)[0] if c is not None else [None] * len(a_leaves)
                    flat += c_leaves
                return tuple(c for c, flag in zip(flat, is_tensor) if flag)

        return Fn.apply(*tensors)
