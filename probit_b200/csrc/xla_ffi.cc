// jax.ffi veneer over the C ABI (include/probit_b200.h): one XLA custom call per reference-facing operation.
//
// BASELINE.json north_star: "Python calls CUDA through thin jax.ffi custom-call C-ABI bindings".
// The XLA FFI headers ship with jaxlib (`jax.ffi.include_dir()`); JAX is NOT installable in this build
// environment, so this file compiles to nothing unless <xla/ffi/api/ffi.h> is on the
// include path, and it has NOT been compiled or run here.  The tested boundary is the C ABI underneath
// (ctypes, probit_b200/_lib.py); INTEGRATION.md §3 shows the jax.ffi.ffi_call side of every handler.
// Build where JAX exists:
//   g++ -shared -fPIC -std=c++17 -I$(python -c 'import jax; print(jax.ffi.include_dir())') \
//       -I/usr/local/cuda/include xla_ffi.cc -L.. -lprobit_b200 -o ../libprobit_b200_xla.so
//
// Handlers (reference call site each one replaces):
//   probit_b200_gram                       prior(theta)(X)                      Laplace.py:7,21,24, VB.py:7,22
//   probit_b200_likelihood                 jit(vmap(ll / grad / hessian))       approximators.py:92-104
//   probit_b200_potrf                      B.cholesky                           Laplace.py:24, VB.py:10,25
//   probit_b200_laplace_fit                LaplaceGP.approximate_posterior      approximators.py:204-210,265-277
//   probit_b200_vb_fit                     VBGP.approximate_posterior           approximators.py:332-339
//   probit_b200_predict                    Approximator.predict                 approximators.py:154-180
//   probit_b200_laplace_gradient           grad of objective_LA (custom VJP)    solvers.py:28-64, approximators.py:132-134
//   probit_b200_vb_gradient                grad of objective_VB                 approximators.py:132-134,316-330
//   probit_b200_predictive_distributions   probit_predictive_distributions      utilities.py:232-249
// Kernel and likelihood are passed as scalar attributes (the lowered specification of probit_b200/kernels.py); the
// cutpoints are an operand.  Workspaces come from XLA's scratch allocator, so a handler holds no state between calls:
// predict and the gradients rebuild what they need from (X, weight, precision).
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define PROBIT_B200_HAVE_XLA_FFI 1
#endif
#endif

#ifdef PROBIT_B200_HAVE_XLA_FFI
#include <cuda_runtime.h>
#include <cmath>
#include <vector>
#include "xla/ffi/api/ffi.h"
#include "../../include/probit_b200.h"

namespace ffi = xla::ffi;
using F64 = ffi::Buffer<ffi::F64>;
using F64Out = ffi::ResultBuffer<ffi::F64>;

static ffi::Error status_to_error(int status) {
    if (status == PB_OK) return ffi::Error::Success();
    return ffi::Error(status == PB_ERR_NUMERIC ? ffi::ErrorCode::kFailedPrecondition : ffi::ErrorCode::kInternal, pb_last_error());
}
static ffi::Error no_memory() { return ffi::Error(ffi::ErrorCode::kResourceExhausted, "probit_b200: workspace allocation failed"); }
static pb_stream_t S(cudaStream_t s) { return reinterpret_cast<pb_stream_t>(s); }

static pb_kernel_spec make_kernel(int32_t base, int32_t periodic, double scale, double stretch_in, double period,
                                  double stretch_out) {
    pb_kernel_spec k{};
    k.base = base; k.periodic = periodic; k.scale = scale;
    k.stretch_in = stretch_in; k.period = period; k.stretch_out = stretch_out;
    return k;
}

// y is an ffi::AnyBuffer: s64 labels for the ordinal likelihoods, f64 targets for the Gaussian one
static pb_problem make_problem(const F64& X, const ffi::AnyBuffer& y, const F64& cutpoints, const pb_kernel_spec& kernel,
                               int32_t lik_kind, double sigma, double eps, int32_t safe_single) {
    pb_problem prob{};
    prob.X = X.typed_data();
    prob.y = y.untyped_data();
    prob.n = X.dimensions()[0];
    prob.D = static_cast<int32_t>(X.dimensions()[1]);
    prob.kernel = kernel;
    prob.lik.kind = lik_kind;
    prob.lik.J = lik_kind == PB_LIK_GAUSSIAN ? 0 : static_cast<int32_t>(cutpoints.dimensions()[0]) - 1;
    prob.lik.sigma = sigma;
    prob.lik.eps = eps;
    prob.lik.safe_single_precision = safe_single;
    prob.lik.cutpoints = lik_kind == PB_LIK_GAUSSIAN ? nullptr : cutpoints.typed_data();
    return prob;
}

// every handler that takes a problem shares these attributes
#define PB_PROBLEM_ATTRS                                                                                              \
    .Attr<int32_t>("base").Attr<int32_t>("periodic").Attr<double>("scale").Attr<double>("stretch_in")                 \
    .Attr<double>("period").Attr<double>("stretch_out").Attr<int32_t>("lik_kind").Attr<double>("sigma")               \
    .Attr<double>("eps").Attr<int32_t>("safe_single")
#define PB_PROBLEM_PARAMS                                                                                             \
    int32_t base, int32_t periodic, double scale, double stretch_in, double period, double stretch_out,              \
    int32_t lik_kind, double sigma, double eps, int32_t safe_single
#define PB_MAKE_PROBLEM make_problem(X, y, cutpoints, make_kernel(base, periodic, scale, stretch_in, period, stretch_out), \
                                     lik_kind, sigma, eps, safe_single)

static void publish_stats(cudaStream_t stream, F64Out& stats, const pb_fit_result& r) {
    const double host[8] = {double(r.iterations), double(r.info), r.error, r.sum_ll, r.ftw, r.logdet,
                            double(r.factorizations), double(r.pcg_iterations)};
    cudaMemcpyAsync(stats->typed_data(), host, sizeof(host), cudaMemcpyHostToDevice, stream);
    cudaStreamSynchronize(stream);
}

// ---- K(X, X) + diag ----------------------------------------------------------------------------------------------
//   operands : X (n, D);  results : K (n, n) row-major;  attrs : kernel spec, jitter (added to the diagonal)
static ffi::Error GramImpl(cudaStream_t stream, ffi::ScratchAllocator scratch, F64 X, F64Out K, int32_t base, int32_t periodic,
                           double scale, double stretch_in, double period, double stretch_out, double jitter) {
    const int64_t n = X.dimensions()[0];
    const int32_t D = static_cast<int32_t>(X.dimensions()[1]);
    if ((n & 1) != 0) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "probit_b200_gram: n must be even (16-byte rows); pad X");
    const pb_kernel_spec k = make_kernel(base, periodic, scale, stretch_in, period, stretch_out);
    const int Df = pb_feature_dim(&k, D);
    auto Z = scratch.Allocate(sizeof(double) * static_cast<size_t>(Df) * n, 256);
    if (!Z.has_value()) return no_memory();
    int st = pb_features(S(stream), &k, X.typed_data(), n, D, D, static_cast<double*>(*Z), n);
    if (st == PB_OK)
        st = pb_gram_sym(S(stream), &k, static_cast<const double*>(*Z), n, Df, n, K->typed_data(), n, nullptr, jitter);
    return status_to_error(st);
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(probit_b200_gram, GramImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Ctx<ffi::ScratchAllocator>()
                                  .Arg<F64>().Ret<F64>()
                                  .Attr<int32_t>("base").Attr<int32_t>("periodic").Attr<double>("scale")
                                  .Attr<double>("stretch_in").Attr<double>("period").Attr<double>("stretch_out")
                                  .Attr<double>("jitter"));

// ---- fused likelihood ---------------------------------------------------------------------------------------------
//   operands : f (batch * n,), y (n,), cutpoints;  results : ll, g, h, d3 (each batch * n)
static ffi::Error LikelihoodImpl(cudaStream_t stream, F64 f, ffi::AnyBuffer y, F64 cutpoints, F64Out ll, F64Out g, F64Out h,
                                 F64Out d3, int32_t lik_kind, double sigma, double eps, int32_t safe_single) {
    pb_likelihood_spec lik{};
    lik.kind = lik_kind;
    lik.J = lik_kind == PB_LIK_GAUSSIAN ? 0 : static_cast<int32_t>(cutpoints.dimensions()[0]) - 1;
    lik.sigma = sigma; lik.eps = eps; lik.safe_single_precision = safe_single;
    lik.cutpoints = lik_kind == PB_LIK_GAUSSIAN ? nullptr : cutpoints.typed_data();
    const int64_t n = y.dimensions()[0];
    const int64_t batch = n ? static_cast<int64_t>(f.element_count()) / n : 1;
    return status_to_error(pb_likelihood(S(stream), &lik, f.typed_data(), y.untyped_data(), n, batch, ll->typed_data(),
                                         g->typed_data(), h->typed_data(), d3->typed_data()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(probit_b200_likelihood, LikelihoodImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<F64>().Arg<ffi::AnyBuffer>().Arg<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>()
                                  .Attr<int32_t>("lik_kind").Attr<double>("sigma").Attr<double>("eps").Attr<int32_t>("safe_single"));

// ---- in-place lower Cholesky --------------------------------------------------------------------------------------
//   operand A (n, n) f64 aliased to the result (input_output_aliases={0: 0}); info (1,) s32 as LAPACK
static ffi::Error PotrfImpl(cudaStream_t stream, ffi::ScratchAllocator scratch, F64 a, F64Out out,
                            ffi::ResultBuffer<ffi::S32> info) {
    const int64_t n = a.dimensions()[0];
    if ((n & 1) != 0) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "probit_b200_potrf: n must be even (16-byte rows)");
    if (out->typed_data() != a.typed_data())
        cudaMemcpyAsync(out->typed_data(), a.typed_data(), sizeof(double) * n * n, cudaMemcpyDeviceToDevice, stream);
    const int64_t bytes = pb_potrf_workspace_bytes(n);
    auto ws = scratch.Allocate(static_cast<size_t>(bytes), 256);
    if (!ws.has_value()) return no_memory();
    return status_to_error(pb_potrf(S(stream), out->typed_data(), n, n, *ws, bytes, info->typed_data(), nullptr));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(probit_b200_potrf, PotrfImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Ctx<ffi::ScratchAllocator>()
                                  .Arg<F64>().Ret<F64>().Ret<ffi::Buffer<ffi::S32>>());

// ---- LaplaceGP.approximate_posterior / VBGP.approximate_posterior ---------------------------------------------------
//   operands : X (n, D), y (n,), cutpoints (J + 1,) (ignored for the Gaussian likelihood)
//   results  : weight (n,), precision (n,), posterior_mean (n,),
//              stats (8,) = [iterations, info, err, sum_ll, f.w, logdet, factorizations, pcg_iterations]
//   Laplace: objective_LA = -sum_ll + f.w / 2 + logdet when final_factor != 0 (Laplace.py:12-30)
static ffi::Error LaplaceFitImpl(cudaStream_t stream, ffi::ScratchAllocator scratch, F64 X, ffi::AnyBuffer y, F64 cutpoints,
                                 F64Out weight, F64Out precision, F64Out mean, F64Out stats, PB_PROBLEM_PARAMS,
                                 double tolerance, int32_t maxiter, double jitter, int32_t final_factor) {
    const pb_problem prob = PB_MAKE_PROBLEM;
    const int64_t bytes = pb_fit_workspace_bytes(prob.n, prob.D);
    auto ws = scratch.Allocate(static_cast<size_t>(bytes), 256);
    if (!ws.has_value()) return no_memory();
    pb_fit_result res{};
    const int status = pb_laplace_fit(S(stream), &prob, tolerance, maxiter, jitter, final_factor, *ws, bytes,
                                      weight->typed_data(), precision->typed_data(), mean->typed_data(), &res, nullptr);
    publish_stats(stream, stats, res);
    return status_to_error(status);
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(probit_b200_laplace_fit, LaplaceFitImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Ctx<ffi::ScratchAllocator>()
                                  .Arg<F64>().Arg<ffi::AnyBuffer>().Arg<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>()
                                  PB_PROBLEM_ATTRS
                                  .Attr<double>("tolerance").Attr<int32_t>("maxiter").Attr<double>("jitter")
                                  .Attr<int32_t>("final_factor"));

static ffi::Error VbFitImpl(cudaStream_t stream, ffi::ScratchAllocator scratch, F64 X, ffi::AnyBuffer y, F64 cutpoints,
                            F64Out weight, F64Out precision, F64Out mean, F64Out stats, PB_PROBLEM_PARAMS, double tolerance,
                            int32_t maxiter) {
    const pb_problem prob = PB_MAKE_PROBLEM;
    const int64_t bytes = pb_fit_workspace_bytes(prob.n, prob.D);
    auto ws = scratch.Allocate(static_cast<size_t>(bytes), 256);
    if (!ws.has_value()) return no_memory();
    pb_fit_result res{};
    const int status = pb_vb_fit(S(stream), &prob, tolerance, maxiter, *ws, bytes, weight->typed_data(),
                                 precision->typed_data(), mean->typed_data(), &res, nullptr);
    publish_stats(stream, stats, res);
    return status_to_error(status);
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(probit_b200_vb_fit, VbFitImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Ctx<ffi::ScratchAllocator>()
                                  .Arg<F64>().Arg<ffi::AnyBuffer>().Arg<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>()
                                  PB_PROBLEM_ATTRS
                                  .Attr<double>("tolerance").Attr<int32_t>("maxiter"));

// ---- Approximator.predict -------------------------------------------------------------------------------------------
//   operands : X (n, D), y (n,) (unused, kept for a uniform problem signature), cutpoints, weight (n,), precision (n,),
//              X_test (n_test, D)
//   results  : mean (n_test,), variance (n_test,)
//   attrs    : problem, chunk (test rows per pass, 0 = 4096), want_variance
static ffi::Error PredictImpl(cudaStream_t stream, ffi::ScratchAllocator scratch, F64 X, ffi::AnyBuffer y, F64 cutpoints,
                              F64 weight, F64 precision, F64 X_test, F64Out mean, F64Out variance, PB_PROBLEM_PARAMS,
                              int64_t chunk, int32_t want_variance) {
    const pb_problem prob = PB_MAKE_PROBLEM;
    const int64_t n_test = X_test.dimensions()[0];
    if (chunk <= 0) chunk = 4096;
    if (chunk > n_test && n_test > 0) chunk = n_test;
    const int64_t bytes = pb_fit_workspace_bytes(prob.n, prob.D);
    auto ws = scratch.Allocate(static_cast<size_t>(bytes), 256);
    if (!ws.has_value()) return no_memory();
    int status;
    if (want_variance) {
        int32_t info = 0;
        status = pb_predict_prepare(S(stream), &prob, precision.typed_data(), 0, *ws, bytes, &info, nullptr);
    } else {
        status = pb_build_features(S(stream), &prob, *ws, bytes);
    }
    if (status != PB_OK) return status_to_error(status);
    const int64_t sbytes = pb_predict_scratch_bytes(prob.n, prob.D, chunk);
    auto sc = scratch.Allocate(static_cast<size_t>(sbytes), 256);
    if (!sc.has_value()) return no_memory();
    status = pb_predict(S(stream), &prob, *ws, weight.typed_data(), X_test.typed_data(), n_test, chunk, *sc, sbytes,
                        mean->typed_data(), want_variance ? variance->typed_data() : nullptr);
    return status_to_error(status);
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(probit_b200_predict, PredictImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Ctx<ffi::ScratchAllocator>()
                                  .Arg<F64>().Arg<ffi::AnyBuffer>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>()
                                  .Ret<F64>().Ret<F64>()
                                  PB_PROBLEM_ATTRS
                                  .Attr<int64_t>("chunk").Attr<int32_t>("want_variance"));

// ---- gradients of the objectives (what the custom VJP of fixed_point_layer returns) ---------------------------------
//   operands : X, y, cutpoints;  result : value_and_grad (4 + J + 1,) =
//              [objective, d/dscale, d/dstretch_out, d/dsigma, d/db_0 .. d/db_J]
//   The handler runs the fit itself (final_factor = 1), so value and gradient come from one call, as
//   jax.value_and_grad(objective) does at approximators.py:132-134.
static ffi::Error LaplaceGradientImpl(cudaStream_t stream, ffi::ScratchAllocator scratch, F64 X, ffi::AnyBuffer y,
                                      F64 cutpoints, F64Out value_and_grad, PB_PROBLEM_PARAMS, double tolerance,
                                      int32_t maxiter, double jitter) {
    const pb_problem prob = PB_MAKE_PROBLEM;
    const int64_t n = prob.n;
    const int64_t bytes = pb_fit_workspace_bytes(n, prob.D);
    auto ws = scratch.Allocate(static_cast<size_t>(bytes), 256);
    auto vecs = scratch.Allocate(sizeof(double) * 3 * static_cast<size_t>(n), 256);
    const int64_t gbytes = pb_gradient_scratch_bytes(n);
    auto gs = scratch.Allocate(static_cast<size_t>(gbytes), 256);
    if (!ws.has_value() || !vecs.has_value() || !gs.has_value()) return no_memory();
    double* w = static_cast<double*>(*vecs);
    double* p = w + n;
    double* f = p + n;
    pb_fit_result res{};
    int status = pb_laplace_fit(S(stream), &prob, tolerance, maxiter, jitter, 1, *ws, bytes, w, p, f, &res, nullptr);
    if (status != PB_OK) return status_to_error(status);
    const int32_t len = static_cast<int32_t>(value_and_grad->element_count()) - 1;
    std::vector<double> host(static_cast<size_t>(len) + 1, 0.0);
    host[0] = -res.sum_ll + 0.5 * res.ftw + res.logdet;                 // objective_LA (Laplace.py:26-30)
    status = pb_laplace_gradient(S(stream), &prob, *ws, bytes, w, p, *gs, gbytes, host.data() + 1, len, nullptr);
    if (status != PB_OK) return status_to_error(status);
    cudaMemcpyAsync(value_and_grad->typed_data(), host.data(), sizeof(double) * host.size(), cudaMemcpyHostToDevice, stream);
    cudaStreamSynchronize(stream);
    return ffi::Error::Success();
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(probit_b200_laplace_gradient, LaplaceGradientImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Ctx<ffi::ScratchAllocator>()
                                  .Arg<F64>().Arg<ffi::AnyBuffer>().Arg<F64>().Ret<F64>()
                                  PB_PROBLEM_ATTRS
                                  .Attr<double>("tolerance").Attr<int32_t>("maxiter").Attr<double>("jitter"));

//   result : value_and_grad as above with objective_VB = f.w / 2 - N log sigma + sum log diag chol(sigma^2 I + K) - sum_ll
//            (closed form of VB.py:26-39); the fit leaves that log-determinant in stats.logdet
static ffi::Error VbGradientImpl(cudaStream_t stream, ffi::ScratchAllocator scratch, F64 X, ffi::AnyBuffer y, F64 cutpoints,
                                 F64Out value_and_grad, PB_PROBLEM_PARAMS, double tolerance, int32_t maxiter) {
    const pb_problem prob = PB_MAKE_PROBLEM;
    const int64_t n = prob.n;
    const int64_t bytes = pb_fit_workspace_bytes(n, prob.D);
    auto ws = scratch.Allocate(static_cast<size_t>(bytes), 256);
    auto vecs = scratch.Allocate(sizeof(double) * 3 * static_cast<size_t>(n), 256);
    const int64_t gbytes = pb_gradient_scratch_bytes(n);
    auto gs = scratch.Allocate(static_cast<size_t>(gbytes), 256);
    if (!ws.has_value() || !vecs.has_value() || !gs.has_value()) return no_memory();
    double* w = static_cast<double*>(*vecs);
    double* p = w + n;
    double* f = p + n;
    pb_fit_result res{};
    int status = pb_vb_fit(S(stream), &prob, tolerance, maxiter, *ws, bytes, w, p, f, &res, nullptr);
    if (status != PB_OK) return status_to_error(status);
    const int32_t len = static_cast<int32_t>(value_and_grad->element_count()) - 1;
    std::vector<double> host(static_cast<size_t>(len) + 1, 0.0);
    host[0] = 0.5 * res.ftw - double(n) * std::log(sigma) + res.logdet - res.sum_ll;
    status = pb_vb_gradient(S(stream), &prob, *ws, bytes, w, *gs, gbytes, host.data() + 1, len, nullptr);
    if (status != PB_OK) return status_to_error(status);
    cudaMemcpyAsync(value_and_grad->typed_data(), host.data(), sizeof(double) * host.size(), cudaMemcpyHostToDevice, stream);
    cudaStreamSynchronize(stream);
    return ffi::Error::Success();
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(probit_b200_vb_gradient, VbGradientImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Ctx<ffi::ScratchAllocator>()
                                  .Arg<F64>().Arg<ffi::AnyBuffer>().Arg<F64>().Ret<F64>()
                                  PB_PROBLEM_ATTRS
                                  .Attr<double>("tolerance").Attr<int32_t>("maxiter"));

// ---- probit_predictive_distributions --------------------------------------------------------------------------------
//   operands : mean (n_test,), variance (n_test,), cutpoints (J + 1,);  result : probabilities (n_test, J)
static ffi::Error PredictiveImpl(cudaStream_t stream, F64 mean, F64 variance, F64 cutpoints, F64Out out, double sigma) {
    pb_likelihood_spec lik{};
    lik.kind = PB_LIK_ORDINAL_PROBIT;
    lik.J = static_cast<int32_t>(cutpoints.dimensions()[0]) - 1;
    lik.sigma = sigma; lik.eps = 0.0; lik.cutpoints = cutpoints.typed_data();
    return status_to_error(pb_predictive_distributions(S(stream), &lik, mean.typed_data(), variance.typed_data(),
                                                       static_cast<int64_t>(mean.element_count()), out->typed_data()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(probit_b200_predictive_distributions, PredictiveImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<F64>().Arg<F64>().Arg<F64>().Ret<F64>().Attr<double>("sigma"));
#endif  // PROBIT_B200_HAVE_XLA_FFI
