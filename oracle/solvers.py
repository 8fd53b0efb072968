"""Oracle restatement of probit/implicit/solvers.py plus the jaxopt loop it delegates to.

jaxopt.FixedPointIteration (>=0.5.5, absent here; published semantics):
maxiter=100; state.error starts at +inf; update: next = f(params), error = ||next - params||_2;
loop while error > tol and iter_num < maxiter; run() returns the last params.
"""
import numpy as np

MAXITER = 100


def fwd_solver(f, z_init, tolerance, maxiter=MAXITER, trace=None):
    """solvers.py:7-15 — plain fixed-point iteration with jaxopt's stopping rule."""
    z = np.asarray(z_init, dtype=np.float64)
    error = np.inf
    it = 0
    while error > tolerance and it < maxiter:
        z_next = f(z)
        error = float(np.linalg.norm(z_next - z))
        z = z_next
        it += 1
        if trace is not None:
            trace.append(error)
    return z


def newton_solver(f, jac_f, z_init, tolerance, maxiter=MAXITER, trace=None):
    """solvers.py:18-25 — Newton on f_root(z) = f(z) - z with the dense Jacobian and an LU solve.

    `jac_f(z)` returns the dense Jacobian of f at z (the reference gets it from jax.jacobian);
    jacobian(f_root) = jac_f - I.
    """
    n = np.asarray(z_init).shape[0]
    eye = np.eye(n)

    def g(z):
        f_root = f(z) - z
        return z - np.linalg.solve(jac_f(z) - eye, f_root)

    return fwd_solver(g, z_init, tolerance, maxiter, trace)
