// Shared helpers for the probit_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <cmath>
#include <atomic>
#include "../../include/probit_b200.h"

namespace pb {

// thread-local message behind pb_last_error()
void set_error(const char* fmt, ...);

// profiling hooks (capi.cu): every kernel launch of the library is counted; when profiling is on the
// trailing-update GEMM launches are bracketed by CUDA events on their own stream.
void note_launch();
void note_launches(long long n);
long long launch_count();
bool profiling_enabled();
void profile_gemm(cudaEvent_t e0, cudaEvent_t e1, double flops, long long launches = 1, int kind = 0);   // kind 1 = INT8-sliced
// While a factorisation is being captured into a CUDA graph the GEMM launcher cannot time its launches with events;
// it adds their algorithmic flops to the capturing thread's tally instead (potrf.cu attributes them to the replay).
struct CaptureTally { double flops = 0; long long launches = 0; };
void set_capture_tally(CaptureTally* t);
CaptureTally* capture_tally();

// Tunables (capi.cu).  There is no mutable global configuration: an extern "C" driver installs the caller's
// pb_options for the duration of its call on the calling thread (OptScope); everything below it reads them here.
const pb_options& opts();
struct OptScope {
    const pb_options* prev;
    explicit OptScope(const pb_options* o);
    ~OptScope();
};
inline long long opt_pcg_min_n() { return opts().laplace_pcg_min_n; }
inline long long opt_nystrom_rank() { return opts().laplace_nystrom_rank; }
inline double opt_cg_tol() { return opts().laplace_cg_tol; }
inline int opt_potrf_nb() { return opts().potrf_block; }
inline bool opt_lookahead() { return opts().potrf_lookahead != 0; }

// true exactly once per (call site, device): guards cudaFuncSetAttribute, which is a per-device setting
struct PerDeviceOnce {
    std::atomic<unsigned long long> done{0};
    bool first() {
        int dev = 0;
        cudaGetDevice(&dev);
        const unsigned long long bit = 1ull << (dev & 63);
        return (done.fetch_or(bit, std::memory_order_acq_rel) & bit) == 0;
    }
};

#define PB_CUDA(expr)                                                                         \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            pb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return PB_ERR_CUDA;                                                               \
        }                                                                                     \
    } while (0)

#define PB_CHECK(cond, code, ...)                                                             \
    do {                                                                                      \
        if (!(cond)) {                                                                        \
            pb::set_error(__VA_ARGS__);                                                       \
            return (code);                                                                    \
        }                                                                                     \
    } while (0)

#define PB_TRY(expr)                                                                          \
    do {                                                                                      \
        int _s = (expr);                                                                      \
        if (_s != PB_OK) return _s;                                                           \
    } while (0)

inline int num_sms() {
    static std::atomic<int> cache[64];
    int dev = 0;
    cudaGetDevice(&dev);
    int n = cache[dev & 63].load(std::memory_order_relaxed);
    if (!n) {
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
        cache[dev & 63].store(n, std::memory_order_relaxed);
    }
    return n;
}

template <typename T>
__host__ __device__ inline T ceil_div(T a, T b) { return (a + b - 1) / b; }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum; result valid in thread 0 (and broadcast to all through smem)
template <int THREADS>
__device__ __forceinline__ double block_sum(double v) {
    __shared__ double red[THREADS / 32];
    __shared__ double total;
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        double x = threadIdx.x < THREADS / 32 ? red[threadIdx.x] : 0.0;
        x = warp_sum(x);
        if (threadIdx.x == 0) total = x;
    }
    __syncthreads();
    return total;
}

// exp(-x) for x >= 0 without the special-case handling of the library exp (the Gram and likelihood
// kernels are issue bound on FP64 transcendentals: ncu sm__throughput 74-76 %).
// Cody-Waite reduction with fdlibm's ln2 split, degree-13 Taylor polynomial on |r| <= ln2/2 (truncation
// 4e-18), exponent added directly to the high word; returns 0 for x > 707 (e^-707 = 8e-308).  <= 2 ulp.
// Two coefficient sources.  As FP64 immediates (default) ptxas re-materialises each coefficient with two UMOVs in front
// of its DFMA; from the constant bank (CONST_BANK) the instruction count drops by a quarter but every evaluation waits
// on LDC latency and holds more registers.  Measured on B200 (ncu, profiles/r02_hbm_kernels_ncu.md): the Gram kernel
// (16 independent evaluations per thread, occupancy limited by registers) is 18 % FASTER with immediates, the likelihood
// kernel (2 evaluations inside a long dependent chain) 5 % faster with the constant bank.
__constant__ double EXP_NEG_C[11] = {
    1.6059043836821613e-10,   // 1/13!
    2.08767569878681e-09,     // 1/12!
    2.505210838544172e-08,    // 1/11!
    2.755731922398589e-07,    // 1/10!
    2.7557319223985893e-06,   // 1/9!
    2.48015873015873e-05,     // 1/8!
    1.984126984126984e-04,    // 1/7!
    1.388888888888889e-03,    // 1/6!
    8.333333333333333e-03,    // 1/5!
    4.1666666666666664e-02,   // 1/4!
    1.6666666666666666e-01};  // 1/3!
template <bool CONST_BANK = false>
__device__ __forceinline__ double exp_neg(double x) {
    const double MAGIC = 6755399441055744.0;                  // 2^52 + 2^51: rounds to nearest integer
    const double t = fma(-x, 1.4426950408889634, MAGIC);
    const int n = __double2loint(t);                          // n = round(-x log2 e) <= 0
    const double nf = t - MAGIC;
    double r = fma(nf, -6.93147180369123816490e-01, -x);      // -x - n ln2_hi (exact product)
    r = fma(nf, -1.90821492927058770002e-10, r);              //      - n ln2_lo
    double p;
    if (CONST_BANK) {
        p = EXP_NEG_C[0];
#pragma unroll
        for (int j = 1; j <= 10; ++j) p = fma(p, r, EXP_NEG_C[j]);
    } else {
        p = 1.6059043836821613e-10;                           // 1/13!
        p = fma(p, r, 2.08767569878681e-09);                  // 1/12!
        p = fma(p, r, 2.505210838544172e-08);                 // 1/11!
        p = fma(p, r, 2.755731922398589e-07);                 // 1/10!
        p = fma(p, r, 2.7557319223985893e-06);                // 1/9!
        p = fma(p, r, 2.48015873015873e-05);                  // 1/8!
        p = fma(p, r, 1.984126984126984e-04);                 // 1/7!
        p = fma(p, r, 1.388888888888889e-03);                 // 1/6!
        p = fma(p, r, 8.333333333333333e-03);                 // 1/5!
        p = fma(p, r, 4.1666666666666664e-02);                // 1/4!
        p = fma(p, r, 1.6666666666666666e-01);                // 1/3!
    }
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const double res = __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
    return x > 707.0 ? 0.0 : res;                              // also x = +inf (n is meaningless there); NaN stays NaN
}

// ---- internal entry points shared between translation units (all enqueue on `stream`) ----

// C[M x N] = alpha * A[M x K] * B[N x K]^T + beta * C, all row-major (K contiguous for A and B).
// lower_only: skip tiles strictly above the diagonal and mask the upper part of diagonal tiles
// (M == N required).  Pointers 16-byte aligned, leading dimensions even.
int gemm_nt(cudaStream_t stream, int64_t M, int64_t N, int64_t K, double alpha, const double* A, int64_t lda,
            const double* B, int64_t ldb, double beta, double* C, int64_t ldc, bool lower_only);
int gemm_nt_mode(cudaStream_t stream, int64_t M, int64_t N, int64_t K, double alpha, const double* A, int64_t lda,
                 const double* B, int64_t ldb, double beta, double* C, int64_t ldc, int lower_only);

int gemm_nt_groups(cudaStream_t stream, const double* P, int64_t ldp, int64_t R, int64_t K, double alpha, double beta,
                   double* c0, int64_t ldc, int64_t cstep, int count, int64_t a0, int64_t astep, int64_t nb);

// ozaki.cu: FP64 contractions on the INT8 tensor cores (tcgen05 kind::i8) by error-free slicing
bool ozaki_supported(int64_t K);
int64_t ozaki_scratch_bytes(int64_t rows, int64_t K);
int ozaki_syrk_lower(cudaStream_t st, int64_t n, int64_t K, double alpha, const double* P, int64_t ldp, double* C,
                     int64_t ldc, void* scratch, int64_t scratch_bytes);
int ozaki_slice(cudaStream_t st, const double* P, int64_t rows, int64_t K, int64_t ldp, void* scratch, int64_t scratch_bytes);
int ozaki_apply(cudaStream_t st, int64_t K, const void* a_scratch, int64_t a_rows, int64_t a_off, int64_t M,
                const void* b_scratch, int64_t b_rows, int64_t b_off, int64_t N, double alpha, double* C, int64_t ldc,
                bool lower_only);
int ozaki_gemm_nt(cudaStream_t st, int64_t M, int64_t N, int64_t K, double alpha, const double* A, int64_t lda,
                  const double* B, int64_t ldb, double* C, int64_t ldc, void* scratch, int64_t scratch_bytes);

// potrf.cu
bool ozaki_enabled(int64_t n);      // pb_options.potrf_ozaki resolved for a matrix of order n
int potrf(cudaStream_t stream, double* A, int64_t n, int64_t lda, void* workspace, int64_t workspace_bytes,
          int32_t* info);
int trsm_right_lt(cudaStream_t stream, const double* L, int64_t n, int64_t ldl, const void* potrf_workspace,
                  double* X, int64_t m, int64_t ldx);

// gram.cu
int feature_dim(const pb_kernel_spec& spec, int D);
int features(cudaStream_t stream, const pb_kernel_spec& spec, const double* X, int64_t n, int D, int64_t ldx,
             double* Z, int64_t ldz);
int gram_sym(cudaStream_t stream, const pb_kernel_spec& spec, const double* Z, int64_t n, int Df, int64_t ldz,
             double* K, int64_t ldk, const double* diag_vec, double diag_scalar);
int gram_cross(cudaStream_t stream, const pb_kernel_spec& spec, const double* Z1, int64_t n1, const double* Z2,
               int64_t n2, int Df, int64_t ldz1, int64_t ldz2, double* K, int64_t ldk, const double* col_scale);
int gram_block(cudaStream_t stream, const pb_kernel_spec& spec, const double* Z, int64_t ldz, int Df, int64_t row0,
               int64_t rows, int64_t col0, int64_t cols, const double* s, double a, double jitter, double* out,
               int64_t ldo);
int gram_matvec_splits(int64_t n1);
int gram_matvec(cudaStream_t stream, const pb_kernel_spec& spec, const double* Z1, int64_t n1, const double* Z2,
                int64_t n2, int Df, int64_t ldz1, int64_t ldz2, const double* v, double* partial, double* y);
int gram_deriv_matvec(cudaStream_t stream, const pb_kernel_spec& spec, const double* Z, int64_t n, int Df, int64_t ldz,
                      const double* K, int64_t ldk, const double* v, double* y);
int64_t gram_deriv_partial_doubles(int64_t n);
int gram_deriv_dots(cudaStream_t stream, const pb_kernel_spec& spec, const double* Z, int64_t n, int Df, int64_t ldz,
                    const double* K, int64_t ldk, const double* Binv, int64_t ldb, const double* a, const double* s,
                    double* partial, double* out);
int set_identity(cudaStream_t stream, double* A, int64_t n, int64_t ld);
int transform_block(cudaStream_t stream, const double* K, int64_t ldk, const double* s, double a, double jitter,
                    int64_t row0, int64_t col0, int64_t rows, int64_t cols, double* out, int64_t ldo);
int sym_transform(cudaStream_t stream, const double* K, int64_t n, int64_t ldk, const double* s, double a,
                  double jitter, double* B, int64_t ldb);

// blas2.cu
int gemv(cudaStream_t stream, const double* A, int64_t rows, int64_t cols, int64_t lda, const double* x, double* y);
int gemv_t_splits(int64_t rows);
int gemv_t_partial(cudaStream_t stream, const double* A, int64_t rows, int64_t cols, int64_t lda, const double* x,
                   double* partial, int64_t ldp);
int64_t symv_lower_scratch_doubles(int64_t n);
int symv_lower(cudaStream_t stream, const double* K, int64_t n, int64_t ld, const double* x, double* y, double* scratch);
int trsv(cudaStream_t stream, const double* L, int64_t n, int64_t ldl, const double* dinv, bool trans, double* rhs,
         double* x);
int logdet_chol(cudaStream_t stream, const double* L, int64_t n, int64_t ldl, double* out);
int build_block_inverses(cudaStream_t stream, const double* L, int64_t n, int64_t ldl, double* dinv);

// likelihood.cu
int likelihood(cudaStream_t stream, const pb_likelihood_spec& spec, const double* f, const void* y, int64_t n,
               int64_t batch, double* ll, double* g, double* h, double* d3);

}  // namespace pb
