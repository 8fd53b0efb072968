"""Prior (kernel) specifications with the mlkernels fluent API subset the reference uses.

The reference's `prior(prior_parameters)` returns an `mlkernels.Kernel`
(examples/regression.py:120-123: `signal_variance * EQ().stretch(lengthscale).periodic(0.5)`;
examples/classification.py:375,389-391: `signal_variance * Matern12().stretch(l)`,
`signal_variance * EQ().stretch(l)`).  mlkernels is not importable in this environment and an
arbitrary Python kernel cannot be compiled to CUDA, so these classes are *specifications*: they
record the expression tree and `lower()` it to the `pb_kernel_spec` struct the CUDA Gram kernels
consume.  Anything outside {EQ, Matern12/Exp, stretch, periodic, scalar scale} raises
NotImplementedError — there is no CPU fallback.

Calling a kernel, `k(X)` / `k(X, Y)` / `k.elwise(X, Y)`, evaluates it on the GPU through the same
CUDA path (returns a torch.float64 CUDA tensor), mirroring how mlkernels kernels are called.
"""
from . import _lib


class Kernel:
    def stretch(self, lengthscale):
        return Stretched(self, lengthscale)

    def periodic(self, period=1.0):
        return Periodic(self, period)

    def __rmul__(self, c):
        return Scaled(self, c)

    def __mul__(self, c):
        if isinstance(c, Kernel):
            raise NotImplementedError("products of kernels are outside the probit_b200 hot path")
        return Scaled(self, c)

    def __add__(self, other):
        raise NotImplementedError("sums of kernels are outside the probit_b200 hot path")

    __radd__ = __add__

    def lower(self, distance_form=0):
        """Flatten the expression tree into a `pb_kernel_spec` (see include/probit_b200.h).

        distance_form=1 selects lab's expansion-form squared distance (TEST-ONLY, see the header)."""
        scale, outer, inner, period = 1.0, 1.0, 1.0, None
        node = self
        while True:
            if isinstance(node, Scaled):
                scale *= float(node.c)
                node = node.k
            elif isinstance(node, Stretched):
                if period is None:
                    outer *= float(node.l)
                else:
                    inner *= float(node.l)
                node = node.k
            elif isinstance(node, Periodic):
                if period is not None:
                    raise NotImplementedError("nested periodic kernels are not supported")
                period = float(node.p)
                node = node.k
            elif isinstance(node, (EQ, Exp)):
                base = _lib.PB_BASE_EQ if isinstance(node, EQ) else _lib.PB_BASE_EXP
                break
            else:
                raise NotImplementedError(f"kernel node {type(node).__name__} is not supported")
        if period is None:
            return _lib.KernelSpec(base, 0, scale, 1.0, 1.0, outer, int(distance_form), 0)
        return _lib.KernelSpec(base, 1, scale, outer, period, inner, int(distance_form), 0)

    def __call__(self, x, y=None):
        from . import linalg
        return linalg.gram(self.lower(), x, y)

    def elwise(self, x, y=None):
        from . import linalg
        return linalg.gram_elwise(self.lower(), x, y)


class EQ(Kernel):
    """exp(-0.5 ||x - y||^2)."""


class Exp(Kernel):
    """exp(-||x - y||)."""


Matern12 = Exp


class Stretched(Kernel):
    def __init__(self, k, l):
        self.k, self.l = k, l


class Periodic(Kernel):
    def __init__(self, k, p):
        self.k, self.p = k, p


class Scaled(Kernel):
    def __init__(self, k, c):
        self.k, self.c = k, c


def from_mlkernels(kernel):
    """Translate an `mlkernels` expression by class name, if mlkernels is what the user's prior returns."""
    name = type(kernel).__name__
    if name == "EQ":
        return EQ()
    if name in ("Exp", "Matern12"):
        return Exp()
    if name == "StretchedKernel":
        return Stretched(from_mlkernels(kernel[0]), float(kernel.stretches[0]))
    if name == "PeriodicKernel":
        return Periodic(from_mlkernels(kernel[0]), float(kernel.period))
    if name == "ScaledKernel":
        return Scaled(from_mlkernels(kernel[0]), float(kernel.scale))
    raise NotImplementedError(f"mlkernels node {name} is outside the probit_b200 hot path")


def as_kernel(obj):
    if isinstance(obj, Kernel):
        return obj
    return from_mlkernels(obj)
