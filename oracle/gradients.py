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


def vb_gradient(K, X, y, w, lik, spec_fields, gaussian):
    """d objective_VB / d (scale, stretch_out, sigma, cutpoints) at the fixed point w of f_VB (VB.py:4-40), in
    closed form.  The reference gets it by reverse mode through `fixed_point_layer` (solvers.py:28-64).

    With M = sigma^2 I + K, f = K w, g = dll/df, W = -d2ll/df2, S = W^1/2 and the identities of DESIGN.md §3.4
        F(theta, w) = 1/2 w^T K w - N log sigma + 1/2 log|M| - sum ll(K w)          (the two traces sum to N)
        T(theta, w) = M^-1 (K w + sigma g(K w)),   fixed point: sigma w = g(K w)
        dF/dw       = K (w - g) = (1 - sigma) f
        I - dT/dw   = sigma M^-1 (sigma I + K W)
    the total derivative is dF/dphi = F_phi + u^T (M T_phi) with
        u = (sigma I + K W)^-1 F_w / sigma = (1 - sigma)/sigma^2 (f - K S A^-1 S f),
        using (sigma I + K W)^-1 = 1/sigma (I - K S A^-1 S) with the SPD matrix A = sigma I + S K S,
    and, using T = w at the fixed point,
        kernel parameter (C = dK/dphi):  M T_phi = -sigma W C w,       F_phi = (1/2 - sigma) w^T C w + 1/2 tr(M^-1 C)
        sigma:                           M T_phi = sigma (g_sigma - w), F_phi = -N/sigma + sigma tr(M^-1) - sum ll_sigma
        cutpoint b:                      M T_phi = sigma g_b,           F_phi = -sum ll_b."""
    n = K.shape[0]
    sigma = float(lik[0])
    f = K @ w
    if gaussian:
        W = -U.hessian_log_gaussian_likelihood(f, y, lik) * np.ones(n)
    else:
        W = -U.hessian_log_probit_likelihood_autodiff(f, y, lik)
    s = np.sqrt(W)
    A = sigma * np.eye(n) + s[:, None] * K * s[None, :]
    Fw = (1.0 - sigma) * f
    u = (Fw - K @ (s * np.linalg.solve(A, s * Fw))) / sigma**2
    Minv = np.linalg.inv(sigma**2 * np.eye(n) + K)
    out = {}
    for name, C in zip(("scale", "stretch_out"), kernel_derivatives(spec_fields, X, K)):
        Cw = C @ w
        out[name] = (0.5 - sigma) * (w @ Cw) + 0.5 * np.sum(Minv * C) - sigma * ((W * u) @ Cw)
    if gaussian:
        r = np.asarray(y, dtype=np.float64) - f
        ll_s = -1.0 / sigma + r * r / sigma**3
        g_s = -2.0 * r / sigma**3
        out["sigma"] = -n / sigma + sigma * np.trace(Minv) - np.sum(ll_s) + sigma * (u @ (g_s - w))
        return out
    parts = ordinal_parameter_partials(f, y, lik)
    y = np.asarray(y, dtype=np.int64)
    cut = np.asarray(lik[1], dtype=np.float64)
    out["sigma"] = (-n / sigma + sigma * np.trace(Minv) - np.sum(parts["sigma"]["ll"])
                    + sigma * (u @ (parts["sigma"]["g"] - w)))
    gc = np.zeros(cut.size)
    np.add.at(gc, y, -parts["lower"]["ll"] + sigma * u * parts["lower"]["g"])
    np.add.at(gc, y + 1, -parts["upper"]["ll"] + sigma * u * parts["upper"]["g"])
    gc[~np.isfinite(cut)] = 0.0
    out["cutpoints"] = gc
    return out
