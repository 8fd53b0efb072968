"""Oracle for the evidence gradient — TEST INFRASTRUCTURE.

The reference obtains d objective / d parameters by JAX reverse mode through `fixed_point_layer`'s
custom VJP (probit/implicit/solvers.py:28-64, probit/approximators.py:132-134): explicit partials at the
fixed point plus the implicit-function term.  JAX is unavailable, so the oracle states the same quantity
in closed form (Rasmussen & Williams 2006, Alg. 5.1 and §5.5.1) with dense NumPy algebra, and
tests/test_oracle_gradient.py pins it against central finite differences of the oracle objective
(the check the reference itself plots, examples/classification.py:104-111).

Objective: Psi = -sum ll(f) + 1/2 f^T w + 1/2 log|B|,  B = I + W^1/2 K W^1/2,  f = K w  (Laplace.py:12-30).
With R = W^1/2 B^-1 W^1/2, V = diag((K^-1 + W)^-1) = (1 - diag(B^-1)) / W, d3 = d^3 ll/df^3, g = d ll/df:
  s2 = +1/2 V o d3          (R&W print -1/2 with the opposite sign convention for the third derivative)
  d(-Psi)/d theta_j = 1/2 w^T C_j w - 1/2 tr(R C_j) + s2^T (b_j - K R b_j),   C_j = dK/d theta_j, b_j = C_j g
  Gaussian noise std: d(-Psi)/d sigma = sum d ll/d sigma - 1/2 sum V dW/d sigma      (d3 = 0)
"""
import numpy as np

from . import utilities as U


def kernel_derivatives(spec_fields, X, K):
    """dK/d scale and dK/d stretch_out for the supported families, from K itself and the pairwise distances.
    spec_fields = dict(base, periodic, scale, stretch_in, period, stretch_out)."""
    base, periodic = spec_fields["base"], spec_fields["periodic"]
    c, l = spec_fields["scale"], spec_fields["stretch_out"]
    Z = np.asarray(X, dtype=np.float64)
    if Z.ndim == 1:
        Z = Z[:, None]
    Z = Z / spec_fields["stretch_in"]
    if periodic:
        a = 2 * np.pi * Z / spec_fields["period"]
        Z = np.concatenate([np.sin(a), np.cos(a)], axis=1)
    Z = Z / l
    r2 = np.zeros_like(K)
    for d in range(Z.shape[1]):
        diff = Z[:, d][:, None] - Z[:, d][None, :]
        r2 += diff * diff
    dK_dscale = K / c
    dK_dl = K * (r2 / l) if base == "eq" else K * (np.sqrt(r2) / l)
    return dK_dscale, dK_dl


def ordinal_parameter_partials(f, y, lik, eps=U.LIKELIHOOD_EPS):
    """Per-datum partial derivatives of ll, g = dll/df and h = d2ll/df2 of the ordinal-probit likelihood
    (utilities.py:56-57) with respect to the noise std and to the datum's lower / upper cutpoint.

    With L(z1, z2) = log(Phi(z2) - Phi(z1) + eps), A = phi(z1)/u, B = phi(z2)/u:
      L_1 = -A, L_2 = B, L_11 = z1 A - A^2, L_12 = A B, L_22 = -z2 B - B^2,
      L_111 = A(1 - z1^2) + 3 z1 A^2 - 2 A^3,  L_112 = -z1 A B + 2 A^2 B,
      L_122 = -2 A B^2 - z2 A B,               L_222 = B(z2^2 - 1) + 3 z2 B^2 + 2 B^3,
    and z_k = (b_k - f)/sigma: dz/df = -1/s, dz/db_k = 1/s, dz/dsigma = -z/s."""
    s, z1, z2, p1, p2, u = U._probit_terms(f, y, lik, eps)
    A, B = p1 / u, p2 / u
    L1, L2 = -A, B
    L11, L12, L22 = z1 * A - A * A, A * B, -z2 * B - B * B
    L111 = A * (1 - z1 * z1) + 3 * z1 * A * A - 2 * A**3
    L112 = -z1 * A * B + 2 * A * A * B
    L122 = -2 * A * B * B - z2 * A * B
    L222 = B * (z2 * z2 - 1) + 3 * z2 * B * B + 2 * B**3
    S = L11 + 2 * L12 + L22
    out = {
        "sigma": dict(ll=-(z1 * L1 + z2 * L2) / s,
                      g=((L1 + L2) + z1 * (L11 + L12) + z2 * (L12 + L22)) / s**2,
                      h=-(2 * S + z1 * (L111 + 2 * L112 + L122) + z2 * (L112 + 2 * L122 + L222)) / s**3),
        "lower": dict(ll=L1 / s, g=-(L11 + L12) / s**2, h=(L111 + 2 * L112 + L122) / s**3),
        "upper": dict(ll=L2 / s, g=-(L12 + L22) / s**2, h=(L112 + 2 * L122 + L222) / s**3),
    }
    return out


def laplace_gradient(K, X, y, w, lik, spec_fields, gaussian):
    """Returns dict(scale=, stretch_out=, sigma=) of d Psi / d (spec field) at the converged weight w."""
    n = K.shape[0]
    f = K @ w
    if gaussian:
        g = U.grad_log_gaussian_likelihood(f, y, lik)
        W = -U.hessian_log_gaussian_likelihood(f, y, lik)
        d3 = np.zeros(n)
    else:
        g = U.grad_log_probit_likelihood_autodiff(f, y, lik)
        W = -U.hessian_log_probit_likelihood_autodiff(f, y, lik)
        d3 = U.third_log_probit_likelihood_autodiff(f, y, lik)
    s = np.sqrt(W)
    B = np.eye(n) + s[:, None] * K * s[None, :]
    Binv = np.linalg.inv(B)
    R = s[:, None] * Binv * s[None, :]
    V = (1.0 - np.diag(Binv)) / W
    s2 = 0.5 * V * d3          # d(-Psi)/df_i = -1/2 V_i dW_ii/df_i = +1/2 V_i d3_i  (W = -d2 ll)
    out = {}
    for name, C in zip(("scale", "stretch_out"), kernel_derivatives(spec_fields, X, K)):
        b = C @ g
        s3 = b - K @ (R @ b)
        dZ = 0.5 * w @ C @ w - 0.5 * np.sum(R * C) + s2 @ s3
        out[name] = -dZ
    if gaussian:
        sigma = float(lik[0])
        dll = np.sum(-1.0 / sigma + (np.asarray(y) - f) ** 2 / sigma**3)
        dW = -2.0 / sigma**3
        out["sigma"] = -(dll - 0.5 * np.sum(V * dW))
        return out
    # ordinal likelihood parameters (R&W §5.5.1): explicit ll and log|B| terms + implicit term through f-hat,
    #   d(-Psi)/dphi = sum dll/dphi - 1/2 sum V dW/dphi + uvec . dg/dphi,  uvec = (K^-1 + W)^-1 s2 = (K - K R K) s2
    uvec = K @ s2 - K @ (R @ (K @ s2))
    parts = ordinal_parameter_partials(f, y, lik)
    y = np.asarray(y, dtype=np.int64)
    cut = np.asarray(lik[1], dtype=np.float64)

    def total(d):
        return d["ll"] - 0.5 * V * (-d["h"]) + uvec * d["g"]

    out["sigma"] = -np.sum(total(parts["sigma"]))
    gc = np.zeros(cut.size)
    np.add.at(gc, y, total(parts["lower"]))          # b[y] is the datum's lower cutpoint
    np.add.at(gc, y + 1, total(parts["upper"]))      # b[y+1] its upper one
    gc[~np.isfinite(cut)] = 0.0
    out["cutpoints"] = -gc
    return out
