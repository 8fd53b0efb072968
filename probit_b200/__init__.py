"""probit_b200 — B200-native (sm_100a) GP inference hot path of bb515/probit.

Drop-in surface (same names as the reference's `probit` package):
    probit_b200.approximators.LaplaceGP / VBGP
    probit_b200.utilities.log_probit_likelihood / log_gaussian_likelihood / probit_predictive_distributions
    probit_b200.kernels.EQ / Matern12 / Exp  (mlkernels fluent API subset)
Everything computes through `libprobit_b200.so` (hand-written CUDA behind the C ABI in
include/probit_b200.h).  Importing the package does not need a GPU; calling it does.
"""
__all__ = ["approximators", "kernels", "utilities", "linalg", "datasets"]
