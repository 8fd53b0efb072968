"""Sanity of the dependency shim that lets the reference's own source run here (oracle/refshim/, TEST
INFRASTRUCTURE), and a regeneration check of the committed reference fixtures.  CPU only.

The shim packages shadow `jax`, `lab`, `jaxopt`, `mlkernels`, so everything runs in a subprocess with its own
sys.path; the regeneration check needs /root/reference and is skipped where that tree is absent (the GPU box)."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "oracle", "refshim")


def _run(code):
    env = dict(os.environ, PYTHONPATH=SHIM + os.pathsep + ROOT)
    out = subprocess.run([sys.executable, "-c", textwrap.dedent(code)], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    return out.stdout


def test_shim_transforms_and_linear_algebra():
    _run("""
        import math, numpy as np, torch
        import jax, jax.numpy as jnp, lab as B, jaxopt, mlkernels
        from jax import grad, vmap, jacobian, value_and_grad, custom_vjp, vjp
        from _refshim_core import Array1D, Dense
        from oracle import kernels as OK

        # grad / vmap / jacobian against closed forms
        f = lambda x, a: jnp.exp(-a * x ** 2)
        x = torch.linspace(-1, 1, 7, dtype=torch.float64)
        g = vmap(grad(f), in_axes=(0, None))(x, 0.7)
        assert torch.allclose(g, -1.4 * x * torch.exp(-0.7 * x ** 2), atol=1e-15)
        J = jacobian(lambda z: torch.sin(z) * z.sum())(x)
        Jref = torch.diag(torch.cos(x) * x.sum()) + torch.sin(x)[:, None] * torch.ones(7)[None, :]
        assert torch.allclose(J, Jref, atol=1e-14)
        # gather-indexing of the cutpoint vector inside vmap (utilities.py:50-51)
        cut = Array1D(torch.tensor([-math.inf, -0.5, 0.5, math.inf], dtype=torch.float64))
        y = torch.tensor([0, 2, 1])
        assert torch.equal(vmap(lambda yy, c: c[yy + 1], in_axes=(0, None))(y, cut), torch.tensor([-0.5, math.inf, 0.5]))
        # custom_vjp drives the user's bwd rule; nondiff arguments are passed first
        def fun(scale, p):
            return (scale * p ** 2).sum()
        cv = custom_vjp(fun, nondiff_argnums=(0,))
        cv.defvjp(lambda scale, p: (fun(scale, p), p), lambda scale, res, ct: (ct * 2 * scale * res + 1.0,))   # "+1" marks the rule
        val, gr = value_and_grad(lambda p: cv(3.0, p))(torch.tensor([1.0, 2.0]))
        assert float(val) == 15.0 and torch.equal(gr, torch.tensor([7.0, 13.0]))
        # jaxopt loop semantics (SURVEY.md §9.4)
        calls = []
        fp = jaxopt.FixedPointIteration(lambda z: (calls.append(1), 0.5 * z + 1.0)[1], tol=1e-3)
        z, state = fp.run(torch.zeros(1))
        assert state.iter_num == len(calls) == 11 and state.error <= 1e-3
        assert jaxopt.FixedPointIteration(lambda z: z + 1.0, tol=1e-3).run(torch.zeros(1))[1].iter_num == 100
        # lab: cholesky regularises a Dense only; diag both ways; triangular and cholesky solves
        A = torch.tensor([[4.0, 1.0], [1.0, 3.0]])
        assert torch.allclose(B.cholesky(A), torch.linalg.cholesky(A), atol=0)
        assert torch.allclose(B.cholesky(Dense(A)), torch.linalg.cholesky(A + 1e-12 * torch.eye(2)), atol=0)
        assert B.diag(torch.tensor([1.0, 2.0])).shape == (2, 2) and torch.equal(B.diag(A), torch.tensor([4.0, 3.0]))
        L = B.cholesky(A); b = torch.tensor([1.0, 2.0])
        assert torch.allclose(B.cholesky_solve(L, b), torch.linalg.solve(A, b), atol=1e-15)
        assert isinstance(Dense(A) + B.diag(b), Dense)
        # mlkernels subset against the oracle's NumPy kernels (both distance forms agree to rounding here)
        X = np.random.default_rng(0).uniform(size=(9, 3)); Y = np.random.default_rng(1).uniform(size=(4, 3))
        for shim_k, ora_k in ((1.3 * mlkernels.EQ().stretch(0.7), 1.3 * OK.EQ().stretch(0.7)),
                              (0.8 * mlkernels.Matern12().stretch(1.1), 0.8 * OK.Matern12().stretch(1.1)),
                              (2.0 * mlkernels.EQ().stretch(0.3).periodic(0.5), 2.0 * OK.EQ().stretch(0.3).periodic(0.5))):
            assert np.allclose(B.dense(shim_k(X, Y)).numpy(), ora_k(X, Y, dist_mode="expand"), rtol=0, atol=1e-13)
            assert np.allclose(B.dense(shim_k(X)).numpy(), ora_k(X, dist_mode="expand"), rtol=0, atol=1e-13)
            assert shim_k.elwise(X, X).shape == (9, 1)
        print("ok")
    """)


@pytest.mark.skipif(not os.path.isdir("/root/reference/probit"), reason="needs the reference tree (build container only)")
def test_committed_reference_fixtures_regenerate():
    """Re-run the reference source over the shim for two cases and compare with the committed ref_*.npz."""
    out = _run(f"""
        import sys, numpy as np
        sys.argv = ["make_reference_golden"]
        sys.path.insert(0, {os.path.join(ROOT, "oracle")!r})
        import make_reference_golden as M
        M.reference_unit_checks()
        for name in ("c2_ordinal_j3_n30", "vb_ordinal_j3_n120"):
            new = M.run_case(name)
            old = np.load({os.path.join(ROOT, "tests", "golden")!r} + "/ref_" + name + ".npz")
            assert set(new) == set(old.files), (sorted(new), sorted(old.files))
            for k in new:
                assert np.allclose(new[k], old[k], rtol=1e-12, atol=1e-13, equal_nan=True), (name, k)
        print("ok")
    """)
    assert out.strip().endswith("ok")
