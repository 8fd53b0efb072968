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

// One datum per thread, the batch of latent vectors in the inner loop: the datum's label, its two cutpoints and their
// finiteness are resolved once and 1 / sigma comes in through the parameters, so an element costs one 8-byte load, the
// arithmetic and one 8-byte store per requested output (no index division, no gather).  blockIdx.y strides over the
// batch; the next element's load is issued before the current one is evaluated.  KIND is a compile-time constant so
// the ordinal path carries none of the safe-mode / Gaussian code (registers, instruction cache).
template <int KIND>
__global__ void __launch_bounds__(256)
likelihood_kernel(lik::Params p, const double* __restrict__ f, const void* __restrict__ yv, int64_t n, int64_t batch,
                  const double* __restrict__ cut, double* __restrict__ ll, double* __restrict__ g,
                  double* __restrict__ h, double* __restrict__ d3) {
    PB_LIK_SMEM(sc);
    lik::stage_cutpoints(p, cut, sc);
    const int64_t d = blockIdx.x * 256ll + threadIdx.x;
    if (d >= n) return;
    const double* tbl = sc + lik::TBL_OFF;
    double b1 = 0.0, b2 = 0.0, yg = 0.0;
    if (KIND == PB_LIK_GAUSSIAN) {
        yg = reinterpret_cast<const double*>(yv)[d];
    } else {
        long long yi = reinterpret_cast<const long long*>(yv)[d];
        yi = yi < 0 ? 0 : (yi >= p.J ? p.J - 1 : yi);     // JAX clamps out-of-range gather indices
        b1 = sc[yi];
        b2 = sc[yi + 1];
    }
    const bool fin1 = b1 != -INFINITY, fin2 = b2 != INFINITY;
    const double is = 1.0 / p.sigma;
    const int64_t step = (int64_t)gridDim.y * n;
    int64_t i = (int64_t)blockIdx.y * n + d;
    const int64_t total = batch * n;
    double fi = i < total ? f[i] : 0.0;
    for (; i < total; i += step) {
        const double fnext = i + step < total ? f[i + step] : 0.0;
        lik::Out o;
        if (KIND == PB_LIK_GAUSSIAN) o = lik::gaussian(fi, yg, p.sigma);
        else if (KIND == PB_LIK_ORDINAL_PROBIT) {
            if (d3) o = lik::ordinal_core<true, true>(fi, b1, b2, fin1, fin2, is, p.eps, tbl);
            else if (ll) o = lik::ordinal_core<true, false>(fi, b1, b2, fin1, fin2, is, p.eps, tbl);
            else o = lik::ordinal_core<false, false>(fi, b1, b2, fin1, fin2, is, p.eps, tbl);
        } else o = lik::ordinal_safe(fi, b1, b2, p.sigma, p.eps, p.ub, p.ub2, p.ub3, tbl);
        if (ll) ll[i] = o.ll;
        if (g) g[i] = o.g;
        if (h) h[i] = o.h;
        if (d3) d3[i] = o.d3;
        fi = fnext;
    }
}

// probit_predictive_distributions (probit/utilities.py:232-249): out[i][j] = probit(s_i, b_j, b_{j+1}, m_i),
// s_i = sqrt(var_i + sigma^2).  One thread per test point: one reciprocal square root, then Phi at each finite cutpoint
// (branch-free table evaluation, likelihood.cuh) — the J + 1 divisions and the sqrt of the literal expression were half
// of the instructions.  STAGED: the CTA's 256 x J probabilities are one contiguous block of `out`, so they go through
// shared memory and leave as full 128-byte lines instead of J strided 8-byte stores per thread.
template <bool STAGED>
__global__ void __launch_bounds__(256)
predictive_kernel(const double* __restrict__ mean, const double* __restrict__ var, int64_t n,
                  const double* __restrict__ cut, int J, double sigma, double* __restrict__ out) {
    PB_LIK_SMEM(sc);
    extern __shared__ __align__(16) double stage[];               // STAGED: 256 x J
    for (int i = threadIdx.x; i <= J; i += blockDim.x) sc[i] = cut[i];
    lik::stage_tables(sc);
    __syncthreads();
    const double* tbl = sc + lik::TBL_OFF;
    const double s2 = sigma * sigma;
    for (int64_t base = blockIdx.x * 256ll; base < n; base += (int64_t)gridDim.x * 256) {
        const int64_t i = base + threadIdx.x;
        if (i < n) {
            const double m = mean[i];
            const double rs = rsqrt(var[i] + s2);
            double* dst = STAGED ? stage + threadIdx.x * J : out + i * J;
            // infinite cutpoints are 0 / 1 whatever the moments (the jnp.where guards of utilities.py:219-224); the
            // test is on the cutpoints, so it is uniform across the warp
            double lo = sc[0] == -INFINITY ? 0.0 : lik::norm_cdf((sc[0] - m) * rs, tbl);
            for (int j = 0; j < J; ++j) {
                const double b2 = sc[j + 1];
                const double hi = b2 == INFINITY ? 1.0 : lik::norm_cdf((b2 - m) * rs, tbl);
                dst[j] = hi - lo;
                lo = b2 == -INFINITY ? 0.0 : hi;
            }
        }
        if (STAGED) {
            __syncthreads();
            const int rows = n - base < 256 ? (int)(n - base) : 256;
            const int count = rows * J;
            double* gout = out + base * J;                       // 2048-byte aligned: base is a multiple of 256
            if ((count & 1) == 0) {                              // 16-byte copies (always for a full block)
                const int pairs = count >> 1;
                const double2* src = reinterpret_cast<const double2*>(stage);
                double2* dst2 = reinterpret_cast<double2*>(gout);
                for (int e = threadIdx.x; e < pairs; e += 256) dst2[e] = src[e];
            } else {
                for (int e = threadIdx.x; e < count; e += 256) gout[e] = stage[e];
            }
            __syncthreads();
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
    const int64_t gx = ceil_div<int64_t>(n, 256);
    PB_CHECK(gx < (1ll << 31), PB_ERR_INVALID, "likelihood: n too large");
    const int64_t cap = (int64_t)num_sms() * 8;
    int64_t gy = ceil_div<int64_t>(cap, gx);
    gy = gy < 1 ? 1 : (gy > batch ? batch : (gy > 65535 ? 65535 : gy));
    const dim3 grid((unsigned)gx, (unsigned)gy, 1);
    if (p.kind == PB_LIK_GAUSSIAN)
        likelihood_kernel<PB_LIK_GAUSSIAN><<<grid, 256, 0, stream>>>(p, f, y, n, batch, spec.cutpoints, ll, g, h, d3);
    else if (p.kind == PB_LIK_ORDINAL_PROBIT)
        likelihood_kernel<PB_LIK_ORDINAL_PROBIT><<<grid, 256, 0, stream>>>(p, f, y, n, batch, spec.cutpoints, ll, g, h, d3);
    else
        likelihood_kernel<PB_LIK_ORDINAL_PROBIT_SAFE><<<grid, 256, 0, stream>>>(p, f, y, n, batch, spec.cutpoints, ll, g, h, d3);
    pb::note_launch();
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
    const unsigned grid = (unsigned)(want < cap ? want : cap);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (lik->J <= 16)
        pb::predictive_kernel<true><<<grid, 256, 256 * lik->J * sizeof(double), st>>>(mean, variance, n_test, lik->cutpoints,
                                                                                      lik->J, lik->sigma, out);
    else
        pb::predictive_kernel<false><<<grid, 256, 0, st>>>(mean, variance, n_test, lik->cutpoints, lik->J, lik->sigma, out);
    pb::note_launch();
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

extern "C" int pb_likelihood(pb_stream_t stream, const pb_likelihood_spec* lik, const double* f, const void* y,
                             int64_t n, int64_t batch, double* ll, double* g, double* h, double* d3) {
    PB_CHECK(lik != nullptr, PB_ERR_INVALID, "likelihood: null spec");
    return pb::likelihood(reinterpret_cast<cudaStream_t>(stream), *lik, f, y, n, batch, ll, g, h, d3);
}
