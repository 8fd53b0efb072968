"""LaplaceGP / VBGP with the reference's class API, running entirely on the B200 CUDA path.

Mirrors probit/approximators.py: same constructor arguments, `approximate_posterior(parameters)`,
`predict(X_test, parameters, weight, precision)`, `objective()`, `weight`, `precision`,
`construct`.  `parameters = (prior_parameters, likelihood_parameters)` exactly as in the examples
(examples/regression.py:139-143, examples/classification.py:414-417).  Arrays come back as
torch.float64 CUDA tensors.

What is recognised: `prior(prior_parameters)` must return a
`probit_b200.kernels.Kernel` (or an mlkernels expression made of the same nodes) and
`log_likelihood` must be `probit_b200.utilities.log_probit_likelihood` or
`log_gaussian_likelihood`.  Anything else raises NotImplementedError: there is no CPU fallback.
"""
import ctypes as C
import math
from abc import ABC, abstractmethod

import torch

from . import _lib, kernels as _kernels, utilities as _util
from .linalg import _dev, _mat2, _ptr, _stream


def _is_float32(x):
    dt = getattr(x, "dtype", None)
    return dt is not None and str(dt).endswith("float32")


def _likelihood_kind(log_likelihood, grad_log_likelihood, hessian_log_likelihood):
    if log_likelihood is _util.log_gaussian_likelihood:
        if grad_log_likelihood is not None or hessian_log_likelihood is not None:
            raise NotImplementedError("custom derivatives of the Gaussian likelihood are not supported")
        return _lib.PB_LIK_GAUSSIAN
    if log_likelihood is _util.log_probit_likelihood:
        safe_g = grad_log_likelihood is _util.grad_log_probit_likelihood
        safe_h = hessian_log_likelihood is _util.hessian_log_probit_likelihood
        if grad_log_likelihood is None and hessian_log_likelihood is None:
            return _lib.PB_LIK_ORDINAL_PROBIT            # autodiff derivatives (approximators.py:92-95)
        if safe_g and safe_h:
            return _lib.PB_LIK_ORDINAL_PROBIT_SAFE        # examples/classification.py:406-407 (commented out there)
        raise NotImplementedError("only the library's own grad/hessian_log_probit_likelihood pair is supported")
    raise NotImplementedError(
        "log_likelihood must be probit_b200.utilities.log_probit_likelihood or log_gaussian_likelihood; "
        "arbitrary Python likelihoods cannot run on the CUDA path and there is no CPU fallback")


class Approximator(ABC):
    """probit/approximators.py:16-210."""

    @abstractmethod
    def __repr__(self):
        ...

    def __init__(self, data, prior, log_likelihood, grad_log_likelihood=None, hessian_log_likelihood=None,
                 tolerance=1e-5, maxiter=100, jitter=1e-12, likelihood_eps=_util.LIKELIHOOD_EPS,
                 predict_chunk=None, options=None, distance_form="direct"):
        self.tolerance = tolerance                      # approximators.py:90
        self.maxiter = maxiter                          # jaxopt FixedPointIteration default
        self.jitter = jitter                            # lab's B.epsilon in B.cholesky(Dense) (Laplace.py:24)
        self.likelihood_eps = likelihood_eps
        self.prior = prior
        self.log_likelihood = log_likelihood
        self._kind = _likelihood_kind(log_likelihood, grad_log_likelihood, hessian_log_likelihood)
        self.lib = _lib.load()
        # driver tunables (pb_options), passed with every call: `options` is a dict of field overrides
        self.options = _lib.default_options(**(options or {}))
        if distance_form not in ("direct", "expand"):
            raise ValueError("distance_form must be 'direct' (the product) or 'expand' (test-only: lab's pw_dists2 form)")
        self._distance_form = 1 if distance_form == "expand" else 0
        X_train, y_train = data
        # float32 inputs (the reference without jax_enable_x64, BASELINE configs[0]) are promoted to float64 on the
        # device — the FP64 path is the product — and results are handed back in the input precision.
        self.out_dtype = torch.float32 if _is_float32(X_train) else torch.float64
        self.X = _mat2(X_train)
        (self.N, self.D) = self.X.shape                 # approximators.py:108
        self.y = _util._labels(self._kind, y_train)
        if self.y.numel() != self.N:
            raise ValueError("X_train and y_train disagree on N")
        self.data = (self.X, self.y)
        self.predict_chunk = predict_chunk
        self._ws = None
        self._ws_bytes = 0
        self._gram_key = None        # kernel spec bytes the workspace's K was built for
        self._factor_key = None      # (spec bytes, precision clone) the workspace's factor was built for
        self.last_result = None

    # -- plumbing ---------------------------------------------------------------------------------
    def _workspace(self):
        if self._ws is None:
            self._ws_bytes = self.lib.pb_fit_workspace_bytes(self.N, self.D)
            self._ws = torch.empty(self._ws_bytes, dtype=torch.uint8, device="cuda")
        return self._ws

    def _spec(self, prior_parameters):
        kernel = _kernels.as_kernel(self.prior(prior_parameters))
        return kernel.lower(self._distance_form)

    def _problem(self, parameters):
        prior_parameters, likelihood_parameters = parameters
        spec = self._spec(prior_parameters)
        lik, keep = _util.make_likelihood_spec(self._kind, likelihood_parameters, self.likelihood_eps)
        prob = _lib.Problem(self.X.data_ptr(), self.y.data_ptr(), self.N, self.D, 0, spec, lik)
        return prob, keep

    @staticmethod
    def _spec_key(spec):
        return bytes(spec)

    def _ensure_gram(self, prob):
        """Features + K(theta) in the workspace (for helpers such as `precision` / `posterior_mean`)."""
        key = self._spec_key(prob.kernel)
        ws = self._workspace()
        if self._gram_key != key:
            self._gram_key, self._factor_key = None, None
            _lib.check(self.lib.pb_build_gram(_stream(), C.byref(prob), _ptr(ws), self._ws_bytes))
            self._gram_key = key
        return ws

    def _K_view(self):
        """(N, N) view of the Gram matrix inside the workspace."""
        ws = self._workspace()
        Kp, ld = C.c_void_p(0), C.c_int64(0)
        _lib.check(self.lib.pb_workspace_gram(_ptr(ws), self.N, self.D, C.byref(Kp), C.byref(ld)))
        off = Kp.value - ws.data_ptr()
        return ws[off: off + self.N * ld.value * 8].view(torch.float64).view(self.N, ld.value)[:, : self.N]

    # -- reference API ----------------------------------------------------------------------------
    @abstractmethod
    def construct(self):
        ...

    @abstractmethod
    def objective(self):
        ...

    @abstractmethod
    def weight(self, parameters):
        ...

    @abstractmethod
    def precision(self, weight, parameters):
        ...

    def posterior_mean(self, weight, parameters):
        """K @ weight (approximators.py:199-202; the reference's version forgets its `parameters` argument)."""
        prob, keep = self._problem(parameters)
        self._ensure_gram(prob)
        from . import linalg
        return linalg.symv(self._K_view(), _dev(weight))

    def predict(self, X_test, parameters, weight, precision, variance=True):
        """approximators.py:154-180: (mean, variance) of the latent GP at X_test, both (N_test,).

        `variance=False` (extension) skips the factorisation and the N^2 * N_test variance solve and returns
        (mean, None): the mean-only sweep over very large test sets (BASELINE configs[4])."""
        if not variance:
            return self._predict_mean(X_test, parameters, weight), None
        prob, keep = self._problem(parameters)
        ws = self._workspace()
        weight = _dev(weight).reshape(-1)
        precision = _dev(precision).reshape(-1)
        X_test = _mat2(X_test)
        if X_test.shape[1] != self.D:
            raise ValueError("X_test has the wrong input dimension")
        ws = self._prepare_factor(prob, precision)
        n_test = X_test.shape[0]
        chunk = self.predict_chunk or max(1, min(n_test, max(256, (1 << 31) // (8 * max(self.N, 1)))))
        scratch_bytes = self.lib.pb_predict_scratch_bytes(self.N, self.D, chunk)
        scratch = torch.empty(scratch_bytes, dtype=torch.uint8, device="cuda")
        mean = torch.empty(n_test, dtype=torch.float64, device="cuda")
        var = torch.empty(n_test, dtype=torch.float64, device="cuda")
        _lib.check(self.lib.pb_predict(_stream(), C.byref(prob), _ptr(ws), _ptr(weight), _ptr(X_test), n_test, chunk,
                                       _ptr(scratch), scratch_bytes, _ptr(mean), _ptr(var)))
        del keep
        return mean.to(self.out_dtype), var.to(self.out_dtype)

    def _prepare_factor(self, prob, precision):
        """Factor of B = I + P^1/2 K P^1/2 for (theta, precision) in the workspace (cached across calls)."""
        ws = self._workspace()
        key = self._spec_key(prob.kernel)
        reuse_factor = (self._factor_key is not None and self._factor_key[0] == key
                        and torch.equal(self._factor_key[1], precision))
        if not reuse_factor:
            info = C.c_int32(0)
            reuse_gram = int(self._gram_key == key)
            self._gram_key, self._factor_key = None, None
            _lib.check(self.lib.pb_predict_prepare(_stream(), C.byref(prob), _ptr(precision), reuse_gram, _ptr(ws),
                                                   self._ws_bytes, C.byref(info), C.byref(self.options)))
            self._gram_key, self._factor_key = key, (key, precision.clone())
        return ws

    def predict_covariance(self, X_test, parameters, weight, precision):
        """approximators.py:182-197: full (N_test, N_test) posterior predictive covariance."""
        from . import linalg
        prob, keep = self._problem(parameters)
        precision = _dev(precision).reshape(-1)
        X_test = _mat2(X_test)
        ws = self._prepare_factor(prob, precision)
        n_test = X_test.shape[0]
        sbytes = self.lib.pb_predict_scratch_bytes(self.N, self.D, n_test)
        scratch = torch.empty(sbytes, dtype=torch.uint8, device="cuda")
        cov = linalg.empty_matrix(n_test, n_test)
        _lib.check(self.lib.pb_predict_covariance(_stream(), C.byref(prob), _ptr(ws), _ptr(X_test), n_test, _ptr(scratch),
                                                  sbytes, _ptr(cov), cov.stride(0)))
        del keep
        return cov.to(self.out_dtype)

    def _predict_mean(self, X_test, parameters, weight):
        prob, keep = self._problem(parameters)
        ws = self._workspace()
        if self._gram_key != self._spec_key(prob.kernel):    # only the training features are needed, not K
            self._gram_key, self._factor_key = None, None
            _lib.check(self.lib.pb_build_features(_stream(), C.byref(prob), _ptr(ws), self._ws_bytes))
        weight = _dev(weight).reshape(-1)
        X_test = _mat2(X_test)
        if X_test.shape[1] != self.D:
            raise ValueError("X_test has the wrong input dimension")
        n_test = X_test.shape[0]
        chunk = self.predict_chunk or max(1, min(n_test, max(256, (1 << 31) // (8 * max(self.N, 1)))))
        scratch_bytes = self.lib.pb_predict_scratch_bytes(self.N, self.D, chunk)
        scratch = torch.empty(scratch_bytes, dtype=torch.uint8, device="cuda")
        mean = torch.empty(n_test, dtype=torch.float64, device="cuda")
        _lib.check(self.lib.pb_predict(_stream(), C.byref(prob), _ptr(ws), _ptr(weight), _ptr(X_test), n_test, chunk,
                                       _ptr(scratch), scratch_bytes, _ptr(mean), None))
        del keep
        return mean.to(self.out_dtype)

    def approximate_posterior(self, parameters):
        """approximators.py:204-210."""
        w, p, _ = self._fit(parameters, final_factor=False)
        return w.to(self.out_dtype), p.to(self.out_dtype)

    def value_and_grad(self):
        """approximators.py:132-134: callable(parameters) -> (objective, gradient with the structure of `parameters`)."""
        raise NotImplementedError(f"{self!r}: value_and_grad is implemented for LaplaceGP only")


class LaplaceGP(Approximator):
    """probit/approximators.py:213-277 + probit/implicit/Laplace.py."""

    def __repr__(self):
        return "LaplaceGP"

    def _fit(self, parameters, final_factor):
        prob, keep = self._problem(parameters)
        ws = self._workspace()
        w = torch.empty(self.N, dtype=torch.float64, device="cuda")
        p = torch.empty_like(w)
        f = torch.empty_like(w)
        res = _lib.FitResult()
        self._gram_key, self._factor_key = None, None
        status = self.lib.pb_laplace_fit(_stream(), C.byref(prob), float(self.tolerance), int(self.maxiter),
                                         float(self.jitter), int(final_factor), _ptr(ws), self._ws_bytes, _ptr(w),
                                         _ptr(p), _ptr(f), C.byref(res), C.byref(self.options))
        self.last_result = res
        _lib.check(status)
        self._gram_key = self._spec_key(prob.kernel)
        del keep
        return w, p, f

    def construct(self):
        """approximators.py:238-246: (parameters, weight) -> f_LA = grad_ll(K w) (Laplace.py:4-9)."""
        def f(parameters, weight):
            m = self.posterior_mean(weight, parameters)
            return _util.evaluate_likelihood(self._kind, m, self.y, parameters[1], ("g",), self.likelihood_eps)["g"]
        return f

    def weight(self, parameters):
        """approximators.py:265-269."""
        return self._fit(parameters, final_factor=False)[0]

    def precision(self, weight, parameters):
        """approximators.py:271-277: (-hessian_ll(K w), K w)."""
        m = self.posterior_mean(weight, parameters)
        h = _util.evaluate_likelihood(self._kind, m, self.y, parameters[1], ("h",), self.likelihood_eps)["h"]
        return -h, m

    def objective(self):
        """approximators.py:248-263 -> objective_LA (Laplace.py:12-30): negative Laplace evidence.

        -sum ll(f) + 0.5 f^T w + sum log diag chol(K + P^-1 + jitter I) + 0.5 sum log p, the last two
        terms evaluated as sum log diag chol(I + P^1/2 (K + jitter I) P^1/2).
        """
        def obj(parameters):
            self._fit(parameters, final_factor=True)
            r = self.last_result
            return -r.sum_ll + 0.5 * r.ftw + r.logdet
        return obj


def _flatten_scalars(tree):
    """Flatten a scalar / nested tuple of scalars; returns (list_of_floats, rebuild(list) -> same structure)."""
    if isinstance(tree, (tuple, list)):
        parts = [_flatten_scalars(t) for t in tree]
        sizes = [len(p[0]) for p in parts]
        flat = [x for p in parts for x in p[0]]

        def rebuild(vals):
            out, i = [], 0
            for (_, rb), k in zip(parts, sizes):
                out.append(rb(vals[i:i + k]))
                i += k
            return type(tree)(out) if isinstance(tree, tuple) else out
        return flat, rebuild
    return [float(tree)], (lambda vals: vals[0])


def _closed_form_value_and_grad(self):
    """value_and_grad(): the objective (negative Laplace evidence / negative ELBO) and its gradient.

    Replaces jit(value_and_grad(objective)) (approximators.py:132-134), i.e. JAX's reverse pass through
    `fixed_point_layer` (implicit/solvers.py:28-64), by the closed form evaluated on the GPU
    (pb_laplace_gradient / pb_vb_gradient; derivations in oracle/gradients.py, checked there against the
    reference's own implicit differentiation).  The gradient w.r.t. the kernel's scale and stretch is mapped back
    to the user's `prior_parameters` through the Jacobian of the (host-side, cheap) kernel-spec lowering, taken by
    central differences.  Returned structure mirrors `parameters`: ((d prior parameters), (d noise_std[, d cutpoints]));
    the infinite end cutpoints get 0.  (Only the series-expansion "safe" likelihood mode has no parameter
    gradient: its sigma / cutpoint entries follow the autodiff expression.)"""
    laplace = isinstance(self, LaplaceGP)

    def vg(parameters):
        prior_parameters, likelihood_parameters = parameters
        w, p, _ = self._fit(parameters, final_factor=True)
        r = self.last_result
        if laplace:
            value = -r.sum_ll + 0.5 * r.ftw + r.logdet
        else:
            value = 0.5 * r.ftw - self.N * math.log(float(likelihood_parameters[0])) + r.logdet - r.sum_ll
        prob, keep = self._problem(parameters)
        ws = self._workspace()
        sbytes = self.lib.pb_gradient_scratch_bytes(self.N)
        scratch = torch.empty(sbytes, dtype=torch.uint8, device="cuda")
        glen = 3 + (prob.lik.J + 1 if self._kind != _lib.PB_LIK_GAUSSIAN else 0)
        g3 = (C.c_double * glen)()
        self._factor_key = None
        if laplace:
            _lib.check(self.lib.pb_laplace_gradient(_stream(), C.byref(prob), _ptr(ws), self._ws_bytes, _ptr(w), _ptr(p),
                                                    _ptr(scratch), sbytes, g3, glen, C.byref(self.options)))
        else:
            _lib.check(self.lib.pb_vb_gradient(_stream(), C.byref(prob), _ptr(ws), self._ws_bytes, _ptr(w),
                                               _ptr(scratch), sbytes, g3, glen, C.byref(self.options)))
        del keep, scratch
        d_scale, d_stretch, d_sigma = g3[0], g3[1], g3[2]
        flat, rebuild = _flatten_scalars(prior_parameters)
        base = self._spec(prior_parameters)
        grads = []
        objective = None
        for i, t in enumerate(flat):
            h = 1e-6 * max(1.0, abs(t))
            up, dn = list(flat), list(flat)
            up[i], dn[i] = t + h, t - h
            su, sd = self._spec(rebuild(up)), self._spec(rebuild(dn))
            if su.base != base.base or su.periodic != base.periodic:
                raise NotImplementedError("a prior parameter that switches the kernel family has no gradient")
            if su.stretch_in != sd.stretch_in or su.period != sd.period:
                # The period and the stretch applied BEFORE the periodic map (EQ().periodic(p).stretch(l)) have no closed
                # form on the GPU (dK/dtheta is not a function of the feature distance alone).  The derivative is taken by
                # central differences of the objective itself: two more fits + factorisations on the GPU per such
                # parameter (the reference's examples do not optimise these; both of theirs use the closed form above).
                if objective is None:
                    objective = self.objective()
                hh = 2e-5 * max(1.0, abs(t))
                up[i], dn[i] = t + hh, t - hh
                fu = objective((rebuild(up), likelihood_parameters))
                fd = objective((rebuild(dn), likelihood_parameters))
                grads.append((fu - fd) / (2 * hh))
                continue
            grads.append(d_scale * (su.scale - sd.scale) / (2 * h) + d_stretch * (su.stretch_out - sd.stretch_out) / (2 * h))
        if objective is not None:
            self._factor_key = None            # the workspace now holds the factor of the last perturbed fit
        g_prior = rebuild(grads)
        if self._kind == _lib.PB_LIK_GAUSSIAN:
            g_lik = (d_sigma,)
        else:
            g_lik = (d_sigma, torch.tensor([g3[3 + j] for j in range(prob.lik.J + 1)], dtype=torch.float64))
        return value, (g_prior, g_lik)
    return vg


LaplaceGP.value_and_grad = _closed_form_value_and_grad


class VBGP(Approximator):
    """probit/approximators.py:280-339 + probit/implicit/VB.py."""

    def __repr__(self):
        return "VBGP"

    def _fit(self, parameters, final_factor):
        prob, keep = self._problem(parameters)
        ws = self._workspace()
        w = torch.empty(self.N, dtype=torch.float64, device="cuda")
        p = torch.empty_like(w)
        f = torch.empty_like(w)
        res = _lib.FitResult()
        self._gram_key, self._factor_key = None, None
        status = self.lib.pb_vb_fit(_stream(), C.byref(prob), float(self.tolerance), int(self.maxiter), _ptr(ws),
                                    self._ws_bytes, _ptr(w), _ptr(p), _ptr(f), C.byref(res), C.byref(self.options))
        self.last_result = res
        _lib.check(status)
        self._gram_key = self._spec_key(prob.kernel)
        del keep
        return w, p, f

    def construct(self):
        """approximators.py:306-314: (parameters, weight) -> f_VB (VB.py:4-16)."""
        def f(parameters, weight):
            from . import linalg
            prob, keep = self._problem(parameters)
            self._ensure_gram(prob)
            sigma = float(parameters[1][0])
            K = self._K_view()
            m = linalg.symv(K, _dev(weight))
            g = _util.evaluate_likelihood(self._kind, m, self.y, parameters[1], ("g",), self.likelihood_eps)["g"]
            A = linalg.empty_matrix(self.N, self.N)
            _lib.check(self.lib.pb_copy_lower_add_diag(_stream(), _ptr(K), self.N, K.stride(0), sigma * sigma,
                                                       _ptr(A), A.stride(0)))
            return linalg.cholesky_solve(linalg.potrf_(A), m + sigma * g)
        return f

    def weight(self, parameters):
        """approximators.py:332-334."""
        return self._fit(parameters, final_factor=False)[0]

    def precision(self, weight, parameters):
        """approximators.py:336-339."""
        m = self.posterior_mean(weight, parameters)
        sigma = float(parameters[1][0])
        return torch.full_like(m, 1.0 / sigma**2), m

    value_and_grad = _closed_form_value_and_grad

    def objective(self):
        """approximators.py:316-330 -> objective_VB (VB.py:19-40): negative ELBO.

        The reference forms C = (sigma^2 I + K)^-1 explicitly; since K C = I - sigma^2 C the two trace
        terms sum to N/2 and cancel the -N/2, leaving
        0.5 f^T w - N log sigma + sum log diag L - sum ll(f)   (SURVEY.md §8a row A9).
        """
        def obj(parameters):
            self._fit(parameters, final_factor=True)
            r = self.last_result
            sigma = float(parameters[1][0])
            return 0.5 * r.ftw - self.N * math.log(sigma) + r.logdet - r.sum_ll
        return obj
