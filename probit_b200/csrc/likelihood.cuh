// Device functions of the per-datum likelihoods, shared by likelihood.cu and the fused
// Newton-step kernels in fit.cu.  Reference: probit/utilities.py (line cites inline).
#pragma once
#include "common.cuh"
#include "ncdf_table.cuh"
#include "log_table.cuh"

namespace pb {
namespace lik {

constexpr double OVER_SQRT_2PI = 0.3989422804014327;    // utilities.py:10
constexpr double LOG_OVER_SQRT_2PI = -0.9189385332046727; // utilities.py:11
constexpr double SQRT2 = 1.4142135623730951;             // utilities.py:12

constexpr int MAX_CUT = 256;
constexpr int NCDF_DOUBLES = NCDF_INTERVALS * (NCDF_DEGREE + 1);
constexpr int LOG_DOUBLES = 2 * LOG_ENTRIES;
constexpr int TBL_OFF = MAX_CUT + 2;                          // even: the log table behind it is read as 16-byte pairs
constexpr int SMEM_DOUBLES = TBL_OFF + NCDF_DOUBLES + LOG_DOUBLES;   // [cutpoints | normal-CDF table | log table] per CTA
static_assert(NCDF_DOUBLES % 2 == 0, "log table must stay 16-byte aligned");
#define PB_LIK_SMEM(name) __shared__ __align__(16) double name[pb::lik::SMEM_DOUBLES]


// Phi(z) - 1/2 = erf(z / sqrt 2) / 2 from the piecewise degree-12 polynomials of ncdf_table.cuh staged in shared
// memory (tools/gen_ncdf_table.py: within ~0.5 ulp(1/2) of the exact value, tighter than a 1-ulp erf).  13 FMAs and 13 shared
// loads instead of the ~100-instruction branchy library erf: the likelihood kernels are FP64-issue bound.
__device__ __forceinline__ double ncdf_half(double z, const double* tbl) {
    // branch-free: |z| >= 8.5 (also +-inf) is evaluated at the top of the last interval, where Q = 1/2 - P ~ 1e-17
    // rounds away and P = 1/2 exactly, the value the reference's erf saturates to; NaN is passed through at the end
    const double a = fmin(fabs(z), 8.499999999999998);
    int k = (int)(a * 4.0);
    const double t = a - (k + 0.5) * 0.25;
    double p = tbl[NCDF_DEGREE * NCDF_INTERVALS + k];
#pragma unroll
    for (int j = NCDF_DEGREE - 1; j >= 0; --j) p = fma(p, t, tbl[j * NCDF_INTERVALS + k]);
    p = k < NCDF_DIRECT ? p : 0.5 - p;                           // outer intervals tabulate Q = 1/2 - P
    p = z != z ? z : p;
    return copysign(p, z);
}
// utilities.py:18-19,31-34: ndtr(z) = 0.5 (1 + erf(z / sqrt 2)), with +-inf mapped to 1 / 0
__device__ __forceinline__ double norm_cdf(double x, const double* tbl) { return 0.5 + ncdf_half(x, tbl); }
// utilities.py:22-23; exp(-inf) = 0 for infinite z
__device__ __forceinline__ double norm_z_pdf(double z) { return OVER_SQRT_2PI * exp_neg<true>(0.5 * z * z); }
__device__ __forceinline__ double series_h(double z) {   // utilities.py:37-44
    const double q = 1.0 / (z * z);
    return -q + 2.5 * q * q - (37.0 / 3.0) * q * q * q;
}
__device__ __forceinline__ double z_far_tails(double z) {   // utilities.py:83-85
    return OVER_SQRT_2PI / z * exp(-0.5 * z * z + series_h(z));
}
__device__ __forceinline__ double z_tails(double z1, double z2) { return z_far_tails(z1) - z_far_tails(z2); }  // :73-80

// log(u) for the likelihood's u = Z + eps (utilities.py:57).  u = m' 2^e with m' in (sqrt 1/2, sqrt 2], J = round(128 m'),
// r = m' fl(128 / J) - 1 (one fma, |r| <= 1/181), log u = e ln2 - log fl(128 / J) + log1p(r) with a degree-7 Taylor
// polynomial: 2 shared loads and ~25 instructions instead of the library's ~60 (these kernels are issue bound).
// <= 1.4 ulp (tools/gen_log_table.py, tests/test_ncdf_table.py).  Anything that is not a positive normal number
// (u <= 0 with eps = 0, NaN, denormals) goes to the library routine.  `ltab` = LOG_TABLE staged in shared memory.
__constant__ double LOG1P_C[6] = {1.0 / 7, -1.0 / 6, 1.0 / 5, -1.0 / 4, 1.0 / 3, -1.0 / 2};
__device__ __forceinline__ double log_pos(double u, const double* ltab) {
    if (__builtin_expect(!(u >= 2.2250738585072014e-308 && u <= 1.7976931348623157e308), 0)) return log(u);
    const int hi = __double2hiint(u);
    const int hm = hi & 0xfffff;
    const int wrap = hm > 0x6a09e ? 1 : 0;                                 // m > sqrt 2: use m / 2, e + 1
    const int J = (128 >> wrap) + ((hm + (0x1000 << wrap)) >> (13 + wrap));
    const double ed = (double)((hi >> 20) - 1023 + wrap);
    const double m = __hiloint2double((hm | 0x3ff00000) - (wrap << 20), __double2loint(u));
    const double2 tc = *reinterpret_cast<const double2*>(ltab + 2 * (J - LOG_J0));
    const double r = fma(m, tc.x, -1.0);
    double q = LOG1P_C[0];
#pragma unroll
    for (int j = 1; j < 6; ++j) q = fma(q, r, LOG1P_C[j]);
    const double t = fma(ed, LOG_LN2_LO, fma(r * r, q, r));
    return fma(ed, LOG_LN2_HI, tc.y) + t;
}

struct Out { double ll, g, h, d3; };

// utilities.py:56-57 and its first three derivatives in f, with everything that depends only on the datum (its two
// cutpoints, which of them are finite) and on the likelihood parameters (1 / sigma) already resolved by the caller:
// the batched kernel evaluates one datum for many latent vectors and pays for those once.
// (two reciprocals — 1/sigma and 1/u — replace the five divisions of the literal expression: <= 2 ulp)
template <bool WANT_LL, bool WANT_D3>
__device__ __forceinline__ Out ordinal_core(double f, double b1, double b2, bool fin1, bool fin2, double is, double eps,
                                            const double* tbl) {
    // An infinite cutpoint gives z = -+inf here, for which norm_cdf returns exactly 0 / 1 and norm_z_pdf exactly 0 — the
    // values of the jnp.where guards at utilities.py:219-224 — so no data-dependent branch is needed (every warp holds a
    // mix of classes and would execute both sides anyway); the z that multiplies the densities is the guarded 0.
    const double zc1 = (b1 - f) * is, zc2 = (b2 - f) * is;
    const double z1 = fin1 ? zc1 : 0.0;                 // utilities.py:217,219-221
    const double z2 = fin2 ? zc2 : 0.0;                 // utilities.py:218,222-224
    const double cdf1 = norm_cdf(zc1, tbl);
    const double cdf2 = norm_cdf(zc2, tbl);
    const double p1 = norm_z_pdf(zc1);
    const double p2 = norm_z_pdf(zc2);
    const double u = (cdf2 - cdf1) + eps;               // utilities.py:225, :57
    const double ru = 1.0 / u, r1 = is * ru, r2 = is * r1;
    Out o;
    o.ll = WANT_LL ? log_pos(u, tbl + NCDF_DOUBLES) : 0.0;
    o.g = (p1 - p2) * r1;
    o.h = (z1 * p1 - z2 * p2) * r2 - o.g * o.g;
    o.d3 = WANT_D3 ? ((z1 * z1 - 1.0) * p1 - (z2 * z2 - 1.0) * p2) * (is * r2) - 3.0 * o.g * o.h - o.g * o.g * o.g : 0.0;
    return o;
}

__device__ __forceinline__ Out ordinal_autodiff(double f, double b1, double b2, double sigma, double eps,
                                                const double* tbl) {
    return ordinal_core<true, true>(f, b1, b2, b1 != -INFINITY, b2 != INFINITY, 1.0 / sigma, eps, tbl);
}

// utilities.py:88-148
__device__ __forceinline__ void safe_Z(double f, double bt, double btp1, double sigma, double ub, double ub2,
                                       double ub3, double& Z, double& z1s, double& z2s, const double* tbl) {
    const double SAFE = 1.0;
    const double _b = (btp1 == INFINITY) ? 0.0 : btp1;
    const double _a = (bt == -INFINITY) ? 0.0 : bt;
    z2s = (btp1 == INFINITY) ? INFINITY : (_b - f) / sigma;
    z1s = (bt == -INFINITY) ? -INFINITY : (_a - f) / sigma;
    Z = norm_cdf(z2s, tbl) - norm_cdf(z1s, tbl);
    double _z1s = (ub < z1s && z1s <= ub2) ? z1s : SAFE;
    const double __z2s = (ub < z1s) ? z2s : SAFE;
    double _z2s = (-ub2 <= z2s && z2s < -ub) ? z2s : SAFE;
    const double __z1s = (-ub > z2s) ? z1s : SAFE;
    Z = (z1s > ub) ? z_tails(_z1s, __z2s) : Z;
    Z = (z2s < -ub) ? z_tails(__z1s, _z2s) : Z;
    _z1s = (ub2 < fabs(z1s) && fabs(z1s) < ub3) ? z1s : SAFE;
    _z2s = (ub2 < fabs(z2s) && fabs(z2s) < ub3) ? z2s : SAFE;
    Z = (z1s > ub2) ? z_far_tails(_z1s) : Z;
    Z = (z2s < -ub2) ? z_far_tails(-_z2s) : Z;
    Z = (z1s >= ub3) ? SAFE : Z;
    Z = (z2s <= -ub3) ? SAFE : Z;
}

// utilities.py:151-192; ll and d3 stay the autodiff expressions (the reference defines no safe ll)
__device__ __forceinline__ Out ordinal_safe(double f, double b1, double b2, double sigma, double eps, double ub,
                                            double ub2, double ub3, const double* tbl) {
    Out o = ordinal_autodiff(f, b1, b2, sigma, eps, tbl);
    double Z, z1s, z2s;
    safe_Z(f, b1, b2, sigma, ub, ub2, ub3, Z, z1s, z2s, tbl);
    const double p1 = norm_z_pdf(z1s), p2 = norm_z_pdf(z2s);
    double E = (p1 - p2) / Z;
    E = (z1s > ub3) ? z1s : E;
    E = (z2s < -ub3) ? z2s : E;
    const double w = E / sigma;
    const double _z1 = isinf(z1s) ? 0.0 : z1s, _z2 = isinf(z2s) ? 0.0 : z2s;
    double V = -(w * w) + (_z1 * p1 - _z2 * p2) / Z / (sigma * sigma);
    V = (z1s > ub3) ? -1.0 / (sigma * sigma) : V;
    V = (z2s < -ub3) ? -1.0 / (sigma * sigma) : V;
    o.g = w;
    o.h = V;
    return o;
}

// utilities.py:60-70.  Written with 1 / sigma so that everything but two multiplications per output is loop invariant in
// the batched kernel (the literal divisions differ from this by <= 2 ulp).
__device__ __forceinline__ Out gaussian(double f, double y, double sigma) {
    const double is = 1.0 / sigma, is2 = is * is;
    const double z = (f - y) * is;                        // utilities.py:68-70
    Out o;
    o.ll = (LOG_OVER_SQRT_2PI - log(sigma)) - 0.5 * z * z;  // utilities.py:64-65,70
    o.g = (y - f) * is2;
    o.h = -is2;
    o.d3 = 0.0;
    return o;
}


struct Params {
    int kind;
    int J;
    double sigma, eps, ub, ub2, ub3;
};

inline int make_params(const pb_likelihood_spec& l, Params& p) {
    p.kind = l.kind; p.J = l.J; p.sigma = l.sigma; p.eps = l.eps;
    p.ub = p.ub2 = p.ub3 = 0.0;
    PB_CHECK(l.sigma > 0, PB_ERR_INVALID, "likelihood: sigma must be positive");
    if (l.kind == PB_LIK_GAUSSIAN) return PB_OK;
    PB_CHECK(l.kind == PB_LIK_ORDINAL_PROBIT || l.kind == PB_LIK_ORDINAL_PROBIT_SAFE, PB_ERR_UNSUPPORTED,
             "likelihood: unknown kind %d", l.kind);
    PB_CHECK(l.J >= 1 && l.J <= MAX_CUT, PB_ERR_INVALID, "likelihood: J must be in [1, %d]", MAX_CUT);
    PB_CHECK(l.cutpoints != nullptr, PB_ERR_INVALID, "likelihood: cutpoints missing");
    if (l.kind == PB_LIK_ORDINAL_PROBIT_SAFE) {
        if (l.safe_single_precision) { p.ub = 1.3; p.ub2 = 1.8; p.ub3 = 2.3; }   // utilities.py:15
        else { p.ub = 2.3; p.ub2 = 3.6; p.ub3 = 4.8; }
    }
    return PB_OK;
}

// Evaluate datum d: `sc` = [cutpoints | normal-CDF table] staged in shared memory by stage_cutpoints (SMEM_DOUBLES).
__device__ __forceinline__ Out eval(const Params& p, double f, const void* __restrict__ yv, int64_t d,
                                    const double* sc) {
    if (p.kind == PB_LIK_GAUSSIAN) return gaussian(f, reinterpret_cast<const double*>(yv)[d], p.sigma);
    long long yi = reinterpret_cast<const long long*>(yv)[d];
    yi = yi < 0 ? 0 : (yi >= p.J ? p.J - 1 : yi);     // JAX clamps out-of-range gather indices
    const double b1 = sc[yi], b2 = sc[yi + 1];
    const double* tbl = sc + TBL_OFF;
    return p.kind == PB_LIK_ORDINAL_PROBIT ? ordinal_autodiff(f, b1, b2, p.sigma, p.eps, tbl)
                                           : ordinal_safe(f, b1, b2, p.sigma, p.eps, p.ub, p.ub2, p.ub3, tbl);
}

// normal-CDF and log tables into sc[TBL_OFF ..] (no barrier)
__device__ __forceinline__ void stage_tables(double* sc) {
    for (int i = threadIdx.x; i < NCDF_DOUBLES; i += blockDim.x) sc[TBL_OFF + i] = NCDF_TABLE[i];
    for (int i = threadIdx.x; i < LOG_DOUBLES; i += blockDim.x) sc[TBL_OFF + NCDF_DOUBLES + i] = LOG_TABLE[i];
}

__device__ __forceinline__ void stage_cutpoints(const Params& p, const double* __restrict__ cut, double* sc) {
    if (p.kind != PB_LIK_GAUSSIAN) {
        for (int i = threadIdx.x; i <= p.J; i += blockDim.x) sc[i] = cut[i];
        stage_tables(sc);
    }
    __syncthreads();
}

}  // namespace lik
}  // namespace pb
