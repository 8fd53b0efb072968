// jax.ffi veneer over the C ABI (include/probit_b200.h).
//
// BASELINE.json north_star: "Python calls CUDA through thin jax.ffi custom-call C-ABI bindings".
// The XLA FFI headers ship with jaxlib (`jax.ffi.include_dir()`); JAX is NOT installable in this build
// environment, so this file compiles to nothing unless <xla/ffi/api/ffi.h> is on the
// include path, and it has NOT been compiled or run here.  The tested boundary is the C ABI underneath.
// Build where JAX exists:
//   g++ -shared -fPIC -std=c++17 -I$(python -c 'import jax; print(jax.ffi.include_dir())') \
//       -I/usr/local/cuda/include xla_ffi.cc -L.. -lprobit_b200 -o ../libprobit_b200_xla.so
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define PROBIT_B200_HAVE_XLA_FFI 1
#endif
#endif

#ifdef PROBIT_B200_HAVE_XLA_FFI
#include <cuda_runtime.h>
#include "xla/ffi/api/ffi.h"
#include "../../include/probit_b200.h"

namespace ffi = xla::ffi;

static ffi::Error status_to_error(int status) {
    if (status == PB_OK) return ffi::Error::Success();
    return ffi::Error(ffi::ErrorCode::kInternal, pb_last_error());
}

static pb_kernel_spec make_kernel(int32_t base, int32_t periodic, double scale, double stretch_in, double period,
                                  double stretch_out) {
    pb_kernel_spec k;
    k.base = base; k.periodic = periodic; k.scale = scale;
    k.stretch_in = stretch_in; k.period = period; k.stretch_out = stretch_out;
    return k;
}

// LaplaceGP.approximate_posterior (probit/approximators.py:204-210) as one custom call.
//   operands : X (n, D) f64, y (n,) s64, cutpoints (J+1,) f64
//   results  : weight (n,), precision (n,), posterior_mean (n,), stats (6,) f64 = [iters, info, err, sum_ll, f.w, logdet]
//   attrs    : kernel spec fields, sigma, eps, tolerance, maxiter, jitter, final_factor
static ffi::Error LaplaceFitImpl(cudaStream_t stream, ffi::ScratchAllocator scratch, ffi::Buffer<ffi::F64> X,
                                 ffi::Buffer<ffi::S64> y, ffi::Buffer<ffi::F64> cutpoints,
                                 ffi::ResultBuffer<ffi::F64> weight, ffi::ResultBuffer<ffi::F64> precision,
                                 ffi::ResultBuffer<ffi::F64> mean, ffi::ResultBuffer<ffi::F64> stats, int32_t base,
                                 int32_t periodic, double scale, double stretch_in, double period, double stretch_out,
                                 double sigma, double eps, double tolerance, int32_t maxiter, double jitter,
                                 int32_t final_factor) {
    const int64_t n = X.dimensions()[0];
    const int32_t D = static_cast<int32_t>(X.dimensions()[1]);
    pb_problem prob{};
    prob.X = X.typed_data(); prob.y = y.typed_data(); prob.n = n; prob.D = D;
    prob.kernel = make_kernel(base, periodic, scale, stretch_in, period, stretch_out);
    prob.lik.kind = PB_LIK_ORDINAL_PROBIT;
    prob.lik.J = static_cast<int32_t>(cutpoints.dimensions()[0]) - 1;
    prob.lik.sigma = sigma; prob.lik.eps = eps; prob.lik.cutpoints = cutpoints.typed_data();
    const int64_t bytes = pb_fit_workspace_bytes(n, D);
    auto ws = scratch.Allocate(static_cast<size_t>(bytes), 256);
    if (!ws.has_value()) return ffi::Error(ffi::ErrorCode::kResourceExhausted, "probit_b200: workspace allocation failed");
    pb_fit_result res{};
    int status = pb_laplace_fit(reinterpret_cast<pb_stream_t>(stream), &prob, tolerance, maxiter, jitter, final_factor,
                                *ws, bytes, weight->typed_data(), precision->typed_data(), mean->typed_data(), &res);
    const double host_stats[6] = {double(res.iterations), double(res.info), res.error, res.sum_ll, res.ftw, res.logdet};
    cudaMemcpyAsync(stats->typed_data(), host_stats, sizeof(host_stats), cudaMemcpyHostToDevice, stream);
    cudaStreamSynchronize(stream);
    return status_to_error(status);
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(probit_b200_laplace_fit, LaplaceFitImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Ctx<ffi::ScratchAllocator>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::S64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Attr<int32_t>("base")
                                  .Attr<int32_t>("periodic")
                                  .Attr<double>("scale")
                                  .Attr<double>("stretch_in")
                                  .Attr<double>("period")
                                  .Attr<double>("stretch_out")
                                  .Attr<double>("sigma")
                                  .Attr<double>("eps")
                                  .Attr<double>("tolerance")
                                  .Attr<int32_t>("maxiter")
                                  .Attr<double>("jitter")
                                  .Attr<int32_t>("final_factor"));

// Fused likelihood (probit/approximators.py:96-104): operands f (n,), y (n,) s64, cutpoints; results ll, g, h.
static ffi::Error LikelihoodImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> f, ffi::Buffer<ffi::S64> y,
                                 ffi::Buffer<ffi::F64> cutpoints, ffi::ResultBuffer<ffi::F64> ll,
                                 ffi::ResultBuffer<ffi::F64> g, ffi::ResultBuffer<ffi::F64> h, double sigma, double eps) {
    pb_likelihood_spec lik{};
    lik.kind = PB_LIK_ORDINAL_PROBIT;
    lik.J = static_cast<int32_t>(cutpoints.dimensions()[0]) - 1;
    lik.sigma = sigma; lik.eps = eps; lik.cutpoints = cutpoints.typed_data();
    const int64_t n = y.dimensions()[0];
    const int64_t batch = n ? static_cast<int64_t>(f.element_count()) / n : 1;
    return status_to_error(pb_likelihood(reinterpret_cast<pb_stream_t>(stream), &lik, f.typed_data(), y.typed_data(), n,
                                         batch, ll->typed_data(), g->typed_data(), h->typed_data(), nullptr));
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(probit_b200_likelihood, LikelihoodImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::S64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Attr<double>("sigma")
                                  .Attr<double>("eps"));

// In-place lower Cholesky (Laplace.py:24): operand A (n, n) f64 aliased to the result (input_output_aliases={0: 0}).
static ffi::Error PotrfImpl(cudaStream_t stream, ffi::ScratchAllocator scratch, ffi::Buffer<ffi::F64> a,
                            ffi::ResultBuffer<ffi::F64> out, ffi::ResultBuffer<ffi::S32> info) {
    const int64_t n = a.dimensions()[0];
    if (out->typed_data() != a.typed_data())
        cudaMemcpyAsync(out->typed_data(), a.typed_data(), sizeof(double) * n * n, cudaMemcpyDeviceToDevice, stream);
    const int64_t bytes = pb_potrf_workspace_bytes(n);
    auto ws = scratch.Allocate(static_cast<size_t>(bytes), 256);
    if (!ws.has_value()) return ffi::Error(ffi::ErrorCode::kResourceExhausted, "probit_b200: workspace allocation failed");
    return status_to_error(pb_potrf(reinterpret_cast<pb_stream_t>(stream), out->typed_data(), n, n, *ws, bytes,
                                    info->typed_data()));
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(probit_b200_potrf, PotrfImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Ctx<ffi::ScratchAllocator>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::S32>>());
#endif  // PROBIT_B200_HAVE_XLA_FFI
