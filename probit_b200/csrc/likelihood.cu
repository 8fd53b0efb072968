// Fused per-datum likelihood kernel: ll, d ll/df, d2 ll/df2, d3 ll/df3 in one pass over f.
//
// Replaces jit(vmap(log_likelihood / grad / hessian)) at probit/approximators.py:92-104 applied to
// log_probit_likelihood (probit/utilities.py:56-57 -> probit :195-229 -> norm_cdf :31-34 -> ndtr
// :18-19) and log_gaussian_likelihood (:60-61, :64-70).  The reference obtains g and h by JAX
// autodiff of log(Z + 1e-10); the closed forms in likelihood.cuh are what that autodiff evaluates
// (the +eps stays inside every denominator; infinite cutpoints carry zero gradient because of
// the jnp.where guards).  PB_LIK_ORDINAL_PROBIT_SAFE follows utilities.py:73-192 instead.
// Only the two cutpoints b[y], b[y+1] are gathered per datum.  HBM traffic: 16 B in, 8 B per
// requested output.
#include "likelihood.cuh"

namespace pb {

namespace {

__global__ void __launch_bounds__(256)
likelihood_kernel(lik::Params p, const double* __restrict__ f, const void* __restrict__ yv, int64_t n, int64_t total,
                  const double* __restrict__ cut, double* __restrict__ ll, double* __restrict__ g,
                  double* __restrict__ h, double* __restrict__ d3) {
    __shared__ double sc[lik::SMEM_DOUBLES];
    lik::stage_cutpoints(p, cut, sc);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t d = (total == n) ? i : i % n;
        const lik::Out o = lik::eval(p, f[i], yv, d, sc);
        if (ll) ll[i] = o.ll;
        if (g) g[i] = o.g;
        if (h) h[i] = o.h;
        if (d3) d3[i] = o.d3;
    }
}

// probit_predictive_distributions (probit/utilities.py:232-249): out[i][j] = probit(s_i, b_j, b_{j+1}, m_i),
// s_i = sqrt(var_i + sigma^2); one thread per test point writes its J probabilities.
__global__ void __launch_bounds__(256)
predictive_kernel(const double* __restrict__ mean, const double* __restrict__ var, int64_t n,
                  const double* __restrict__ cut, int J, double sigma, double* __restrict__ out) {
    __shared__ double sc[lik::SMEM_DOUBLES];
    for (int i = threadIdx.x; i <= J; i += blockDim.x) sc[i] = cut[i];
    for (int i = threadIdx.x; i < lik::NCDF_DOUBLES; i += blockDim.x) sc[lik::MAX_CUT + 1 + i] = lik::NCDF_TABLE[i];
    __syncthreads();
    const double* tbl = sc + lik::MAX_CUT + 1;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double m = mean[i];
        const double s = sqrt(var[i] + sigma * sigma);
        double lo = (sc[0] == -INFINITY) ? 0.0 : lik::norm_cdf((sc[0] - m) / s, tbl);     // utilities.py:219-221
        for (int j = 0; j < J; ++j) {
            const double b2 = sc[j + 1];
            const double hi = (b2 == INFINITY) ? 1.0 : lik::norm_cdf((b2 - m) / s, tbl);  // utilities.py:222-224
            out[i * J + j] = hi - lo;
            lo = (b2 == -INFINITY) ? 0.0 : hi;
        }
    }
}

}  // namespace

int likelihood(cudaStream_t stream, const pb_likelihood_spec& spec, const double* f, const void* y, int64_t n,
               int64_t batch, double* ll, double* g, double* h, double* d3) {
    PB_CHECK(n >= 0 && batch >= 1, PB_ERR_INVALID, "likelihood: bad n/batch");
    const int64_t total = n * batch;
    if (total == 0) return PB_OK;
    lik::Params p;
    PB_TRY(lik::make_params(spec, p));
    const int64_t want = ceil_div<int64_t>(total, 256);
    const int64_t cap = (int64_t)num_sms() * 8;
    likelihood_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, stream>>>(p, f, y, n, total, spec.cutpoints, ll, g,
                                                                               h, d3); pb::note_launch();
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

}  // namespace pb

extern "C" int pb_predictive_distributions(pb_stream_t stream, const pb_likelihood_spec* lik, const double* mean,
                                           const double* variance, int64_t n_test, double* out) {
    PB_CHECK(lik != nullptr && lik->cutpoints != nullptr, PB_ERR_INVALID, "predictive_distributions: null spec");
    PB_CHECK(lik->J >= 1 && lik->J <= pb::lik::MAX_CUT, PB_ERR_INVALID, "predictive_distributions: bad J");
    if (n_test == 0) return PB_OK;
    const int64_t want = pb::ceil_div<int64_t>(n_test, 256);
    const int64_t cap = (int64_t)pb::num_sms() * 8;
    pb::predictive_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        mean, variance, n_test, lik->cutpoints, lik->J, lik->sigma, out); pb::note_launch();
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

extern "C" int pb_likelihood(pb_stream_t stream, const pb_likelihood_spec* lik, const double* f, const void* y,
                             int64_t n, int64_t batch, double* ll, double* g, double* h, double* d3) {
    PB_CHECK(lik != nullptr, PB_ERR_INVALID, "likelihood: null spec");
    return pb::likelihood(reinterpret_cast<cudaStream_t>(stream), *lik, f, y, n, batch, ll, g, h, d3);
}
