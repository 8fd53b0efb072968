"""CPU oracle for the probit GP inference hot path — TEST INFRASTRUCTURE ONLY.

This package is a NumPy/SciPy float64 restatement of the reference's algorithm
(bb515/probit, /root/reference) for the path SURVEY.md §8 scopes:
Gram -> Laplace/VB fixed point -> evidence -> predict.

PARITY STATUS: pinned to the REFERENCE SOURCE, not to the JAX stack.  The reference ships no golden
vectors (its only test pins the scalar series h(x), probit/test/test_implicit.py:11-22) and JAX / lab /
mlkernels / jaxopt are neither vendored nor installable here.  oracle/make_reference_golden.py therefore
imports the reference's own modules unmodified over a torch-backed shim of those four packages
(oracle/refshim/) and commits what they compute as tests/golden/ref_*.npz; tests/test_reference_golden.py
holds this oracle to those outputs at 1e-11 (observed <= 4e-15; gradients against the reference's implicit
differentiation 1e-10 .. 1e-7).  Further pins (tests/test_oracle_*.py):
  * the reference's own h(x) assertions,
  * 50-digit mpmath evaluation and differentiation of the literal reference expression
    log(Phi(z2) - Phi(z1) + 1e-10) (probit/utilities.py:56-57,195-229),
  * the closed form of exact GP regression for the Gaussian likelihood,
  * finite differences of its own objective for the gradient rows,
  * frozen fixtures under tests/golden/ produced by oracle/make_golden.py.
Third-party behaviour restated from the published algorithms:
mlkernels>=0.3.6 (EQ, Exp/Matern12, stretch, periodic, scale), backends(lab)>=1.4.32
(B.cholesky on matrix.Dense adds B.epsilon=1e-12; pw_dists2 expansion form),
jaxopt>=0.5.5 FixedPointIteration (maxiter=100, l2 stopping rule), jax>=0.4.1.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package.  The product (probit_b200/) never does.
"""
