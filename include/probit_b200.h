/* probit_b200 — C ABI of the B200-native (sm_100a) GP inference hot path of bb515/probit.
 *
 * The reference has no FFI of its own: its boundary is the Python class API
 * (probit/approximators.py:55-63,154-180,204-210,248-263).  These entry points are what a
 * `jax.ffi` custom call (or ctypes / DLPack shim) for that path binds; INTEGRATION.md shows the
 * reference-side binding.  Each function cites the reference code it replaces.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless the name ends in `_host`;
 *  - matrices are row-major with an explicit leading dimension (elements, must be even and the
 *    base pointer 16-byte aligned: the TMA tensor maps require it);
 *  - all work is enqueued on the caller's stream; functions that must read a convergence scalar
 *    synchronise that stream themselves and say so;
 *  - the caller owns every buffer, including scratch (`*_workspace_bytes` tells the size);
 *  - return value: PB_OK or a negative PB_ERR_* code; `pb_last_error()` holds a thread-local
 *    message.  Numerical failure (non-positive Cholesky pivot) is reported through the device
 *    `info` word (LAPACK convention: 1-based index of the failing column), never by aborting.
 *  - there is no CPU fallback anywhere behind this interface.
 */
#ifndef PROBIT_B200_H
#define PROBIT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* pb_stream_t; /* == cudaStream_t */

enum {
    PB_OK = 0,
    PB_ERR_INVALID = -1,  /* bad argument */
    PB_ERR_CUDA = -2,     /* CUDA runtime/driver error */
    PB_ERR_UNSUPPORTED = -3,
    PB_ERR_NUMERIC = -4   /* non-SPD matrix / NaN detected by a driver that reads `info` */
};

/* ---- prior (kernel) specification ------------------------------------------------------------
 * Replaces the mlkernels expression returned by the user's `prior(prior_parameters)`
 * (examples/regression.py:120-123, examples/classification.py:375,389-391):
 *   k(x,y) = scale * base( phi(x), phi(y) ),  phi(x) = T(x / stretch_in) / stretch_out,
 *   T = identity, or the periodic feature map u -> [sin(2 pi u / period), cos(2 pi u / period)].
 *   base EQ : exp(-0.5 r^2);  base EXP (Matern12): exp(-r);  r = ||phi(x) - phi(y)||_2.          */
enum { PB_BASE_EQ = 0, PB_BASE_EXP = 1 };

typedef struct pb_kernel_spec {
    int32_t base;       /* PB_BASE_* */
    int32_t periodic;   /* 0 / 1 */
    double scale;       /* c in c * k */
    double stretch_in;  /* stretch applied before the periodic map (1 if none) */
    double period;      /* period p (ignored unless periodic) */
    double stretch_out; /* stretch applied to the features handed to the base kernel */
    int32_t distance_form; /* 0: r^2 = sum_d (a_d - b_d)^2 (the product); 1: lab's B.pw_dists2 expansion
                              ||a||^2 + ||b||^2 - 2 a.b with sqrt(max(., 1e-30)) (TEST-ONLY: reproduces the
                              reference's O(1e-8) diagonal noise for Matern12, D > 1) */
    int32_t _pad;
} pb_kernel_spec;

/* ---- likelihood specification ---------------------------------------------------------------
 * Replaces the per-datum callables wrapped at probit/approximators.py:92-104:
 *  PB_LIK_ORDINAL_PROBIT  log_probit_likelihood (probit/utilities.py:56-57,195-229) and its
 *                         autodiff gradient/Hessian: ll = log(Z + eps),
 *                         Z = Phi((b[y+1]-f)/sigma) - Phi((b[y]-f)/sigma), Phi = 0.5(1+erf(z/sqrt2)).
 *  PB_LIK_GAUSSIAN        log_gaussian_likelihood (probit/utilities.py:60-61).
 *  PB_LIK_ORDINAL_PROBIT_SAFE  the optional series-expansion gradient/Hessian
 *                         (probit/utilities.py:88-192); ll is still utilities.py:56-57.           */
enum { PB_LIK_ORDINAL_PROBIT = 0, PB_LIK_GAUSSIAN = 1, PB_LIK_ORDINAL_PROBIT_SAFE = 2 };

typedef struct pb_likelihood_spec {
    int32_t kind;            /* PB_LIK_* */
    int32_t J;               /* number of classes; cutpoints has J+1 entries (ordinal only) */
    double sigma;            /* likelihood_parameters[0] (noise std) */
    double eps;              /* the +1e-10 inside log() at probit/utilities.py:57 */
    const double* cutpoints; /* device, J+1 doubles, b[0] = -inf, b[J] = +inf (ordinal only) */
    int32_t safe_single_precision; /* SAFE mode: 1 -> BOUNDS["single"], 0 -> BOUNDS["double"] (utilities.py:15) */
    int32_t _pad;
} pb_likelihood_spec;

int pb_version(void);
const char* pb_last_error(void);

/* Tunables of the drivers.  The library keeps NO mutable global configuration: every driver that has a policy to
 * choose takes a `const pb_options*` (NULL = the defaults below), so concurrent callers on different host threads
 * (XLA runs FFI handlers from several) cannot disturb one another.  Fill with pb_options_default() first. */
typedef struct pb_options {
    int64_t laplace_pcg_min_n;    /* smallest N at which Newton steps are solved by preconditioned CG instead of a
                                     fresh factorisation; default 24576, 0 = always, huge = never */
    int64_t laplace_nystrom_rank; /* landmarks of the Nystrom CG preconditioner: -1 = auto = n/16 clamped to
                                     [256, 4096]; 0 = off (factor once, reuse the stale factor as preconditioner) */
    double laplace_cg_tol;        /* eta: a CG Newton solve stops once the error it leaves in the step is
                                     <= eta * tolerance; default 1e-2 (floor: 1e-15 relative residual) */
    double negative_curvature_tol; /* W = -h is negative where Z <~ eps (utilities.py:57 is not log-concave there); values in
                                     [-tol / sigma^2, 0) are treated as 0; anything below sends that Newton step through the
                                     signed Cholesky M = L diag(I, -I) L^T of D + S K S (single GPU), which follows the
                                     reference's LU step whenever K^-1 + W is positive definite and reports
                                     PB_ERR_NUMERIC otherwise; default 1e-6 */
    int32_t potrf_block;          /* Cholesky panel width, 0 = auto */
    int32_t potrf_lookahead;      /* 0/1, default 1 */
    int32_t potrf_graph;          /* 1: replay the factorisation's launch DAG from a cached CUDA graph (second and later
                                     calls with the same buffers).  Default 0: measured SLOWER than eager issue on B200
                                     (N=16384: 64.8 vs 59.7 ms) — the look-ahead stream's priority does not survive
                                     capture, and eager issue is not host-bound */
    int32_t dist_block;           /* panel width of the multi-GPU block-cyclic factorisation, 0 = auto */
    int32_t potrf_ozaki;          /* trailing updates of the Cholesky factorisation on the INT8 tensor cores (tcgen05
                                     kind::i8 through error-free slicing, pb_ozaki_gemm_nt): -1 = auto (on for n >= 8192),
                                     0 = FP64 DMMA only, 1 = wherever the shapes allow */
    int32_t ozaki_tile;           /* variant of the INT8-sliced contraction kernel: 0 = 128x64 tiles, all 7 levels in one pass
                                     (default: the fastest measured); 1 = 128x128 tiles, levels 2-5 and 6-8 in two passes
                                     (TMEM holds 512 columns); 2 = 128x64 with clusters of two CTAs sharing the A tile by TMA
                                     multicast.  All three give identical results (tests/test_gpu_kernels.py). */
} pb_options;
int pb_options_default(pb_options* options);

/* Measurement hooks used by bench.py: total kernel launches issued by this library so far, and
 * CUDA-event timing (on the launching stream) of the trailing-update GEMM kernel between
 * pb_profile_begin() and pb_profile_end() (call the latter after a device synchronise).        */
long long pb_launch_count(void);
int pb_profile_begin(void);
int pb_profile_end(long long* gemm_launches, double* gemm_ms, double* gemm_flops);
/* INT8-sliced contractions (oz_gemm_kernel) inside the region closed by the last pb_profile_end: launches, summed
 * CUDA-event duration, summed FP64-equivalent flops (2 M N K; each launch is 28 exact int8 GEMMs of that shape). */
int pb_profile_int8(long long* launches, double* ms, double* fp64_equivalent_flops);
/* FP64 tensor-core (DMMA) peak of the current device in TFLOP/s, from a register-resident mma.sync loop run on the
 * default stream (synchronises it): the denominator of the Cholesky roofline fractions (MEASURED_PEAKS.json has no
 * FP64 entry).  ~50 ms. */
int pb_measure_fp64_tensor_peak(double* tflops_host);

/* ---- K4: fused per-datum likelihood (value, gradient, Hessian, third derivative) --------------
 * probit/approximators.py:96-104.  y is int64 class labels (ordinal) or f64 targets (Gaussian).
 * Any of ll/g/h/d3 may be NULL.  `batch` independent f-vectors of length n share y (the restart
 * batch of BASELINE config 5); f and outputs are (batch, n) row-major contiguous.               */
int pb_likelihood(pb_stream_t stream, const pb_likelihood_spec* lik, const double* f, const void* y,
                  int64_t n, int64_t batch, double* ll, double* g, double* h, double* d3);

/* ---- K1/K2: Gram assembly ---------------------------------------------------------------------
 * pb_features: Z[Df x n] = phi(X[n x D]) stored FEATURE-MAJOR (Z[d * ldz + i], ldz >= n);
 *              Df = pb_feature_dim(spec, D).
 * pb_gram_sym: K = k(X,X) (+ diag_scalar I) (+ diag(diag_vec)); both triangles written (mirror
 *              store); replaces `prior(theta)(X)` at Laplace.py:7,21,24, VB.py:7,22.
 * pb_gram_cross: K[n1 x n2] = k(X1, X2); replaces `kernel(X_train, X_test)` approximators.py:173.
 * Z buffers come from pb_features.                                                              */
int pb_feature_dim(const pb_kernel_spec* spec, int D);
int pb_features(pb_stream_t stream, const pb_kernel_spec* spec, const double* X, int64_t n, int D,
                int64_t ldx, double* Z, int64_t ldz);
int pb_gram_sym(pb_stream_t stream, const pb_kernel_spec* spec, const double* Z, int64_t n, int Df,
                int64_t ldz, double* K, int64_t ldk, const double* diag_vec, double diag_scalar);
int pb_gram_cross(pb_stream_t stream, const pb_kernel_spec* spec, const double* Z1, int64_t n1,
                  const double* Z2, int64_t n2, int Df, int64_t ldz1, int64_t ldz2, double* K, int64_t ldk);

/* B = I + s_i (K_ij + jitter*delta_ij) s_j on the lower triangle (upper left untouched):
 * the SPD Newton matrix equivalent to `K + diag(1/precision)` (Laplace.py:24, approximators.py:175). */
int pb_scale_sym_plus_identity(pb_stream_t stream, const double* K, int64_t n, int64_t ldk,
                               const double* s, double jitter, double* B, int64_t ldb);
/* A = K + diag_scalar*I on the lower triangle (VB.py:10,25: sigma^2 I + K). */
int pb_copy_lower_add_diag(pb_stream_t stream, const double* K, int64_t n, int64_t ldk,
                           double diag_scalar, double* A, int64_t lda);

/* ---- K7: Cholesky ----------------------------------------------------------------------------
 * In-place lower Cholesky A = L L^T of the row-major lower triangle (strict upper never read or
 * written).  Replaces B.cholesky at Laplace.py:24, VB.py:10,25 and the LU at solvers.py:24.
 * Blocked right-looking with one-panel look-ahead.  Trailing updates: for n >= 8192 (pb_options.potrf_ozaki) exact
 * int8 digit-plane GEMMs on the tcgen05 tensor cores (FP64 by error-free slicing, csrc/ozaki.cu), otherwise FP64
 * tensor-core (DMMA) tiles fed by TMA.  The workspace holds the 64x64 leaf inverses and 256x256 block inverses the
 * solves need and, for n >= 4096, two int8 slicing buffers of one 1024-wide panel (7 n KiB + 4 n bytes each).
 * `info` (device int32): 0, or 1-based column of the first non-positive pivot.                   */
int64_t pb_potrf_workspace_bytes(int64_t n);
int pb_potrf(pb_stream_t stream, double* A, int64_t n, int64_t lda, void* workspace,
             int64_t workspace_bytes, int32_t* info, const pb_options* options);

/* Fill the solve workspace (64x64 leaf inverses + 256x256 diagonal-block inverses used by pb_trsv /
 * pb_trsm_right_lt) for a factor L that was produced elsewhere (multi-GPU block-cyclic Cholesky). */
int pb_rebuild_solve_workspace(pb_stream_t stream, const double* L, int64_t n, int64_t ldl, void* workspace,
                               int64_t workspace_bytes);

/* out[r][c] = s_i K_ij s_j (+ a + s_i^2 jitter where i == j), i = row0 + r, j = col0 + c: one rectangular
 * block of a I + s s^T o (K + jitter I) (s may be NULL = ones), e.g. a block column of a block-cyclic layout. */
int pb_transform_block(pb_stream_t stream, const double* K, int64_t ldk, const double* s, double a,
                       double jitter, int64_t row0, int64_t col0, int64_t rows, int64_t cols, double* out,
                       int64_t ldo);

/* C[M x N] = alpha * A[M x K] * B[N x K]^T + beta * C (all row-major). Exposed for tests/bench. */
int pb_gemm_nt(pb_stream_t stream, int64_t M, int64_t N, int64_t K, double alpha, const double* A,
               int64_t lda, const double* B, int64_t ldb, double beta, double* C, int64_t ldc,
               int32_t lower_only);

/* The same contraction, C += alpha * A * B^T (no beta), on the INT8 tensor cores (tcgen05.mma.kind::i8, TMEM) by
 * error-free slicing of the FP64 operands into 7 base-128 digit planes with one exponent per row (Ozaki scheme,
 * csrc/ozaki.cu): 28 exact int8 x int8 -> int32 products per FP64 product, about 3x the FP64 DMMA rate.  K must be a
 * multiple of 64.  lower_only != 0 is the SYRK form (A == B, M == N, only tiles on/below the diagonal are updated).
 * The result differs from the FP64 product by ~2^-49 of (row scale of A) x (row scale of B) x sqrt(K).
 * scratch: pb_ozaki_scratch_bytes(M, N, K), 256-byte aligned. */
int64_t pb_ozaki_scratch_bytes(int64_t M, int64_t N, int64_t K);
int pb_ozaki_gemm_nt(pb_stream_t stream, int64_t M, int64_t N, int64_t K, double alpha, const double* A, int64_t lda,
                     const double* B, int64_t ldb, double* C, int64_t ldc, int32_t lower_only, void* scratch,
                     int64_t scratch_bytes);

/* ---- K3/K8: level-2 kernels -------------------------------------------------------------------
 * pb_symv : y = K x for a full-storage symmetric K (Laplace.py:8,22; VB.py:9,23).
 * pb_trsv : solve with the lower factor; trans=0: L x = rhs, trans=1: L^T x = rhs
 *           (B.cholesky_solve at VB.py:11).  `rhs` is destroyed, `x` must not alias it;
 *           `potrf_workspace` is the workspace pb_potrf filled (64x64 leaf inverses).
 * pb_logdet_chol: out[0] = sum_i log L_ii (Laplace.py:28, VB.py:28).                              */
int pb_symv(pb_stream_t stream, const double* K, int64_t n, int64_t ldk, const double* x, double* y);
/* The same product reading only the tiles on and below the diagonal (half the HBM traffic); needs ldk >= n rounded
 * up to 64 and pb_symv_lower_scratch_bytes(n) = ~n^2/8 bytes of scratch.  Deterministic (no atomics). */
int64_t pb_symv_lower_scratch_bytes(int64_t n);
/* y = A x for a general row-major A (rows x cols, leading dimension lda): one row block of the product above. */
int pb_gemv(pb_stream_t stream, const double* A, int64_t rows, int64_t cols, int64_t lda, const double* x, double* y);
int pb_symv_lower(pb_stream_t stream, const double* K, int64_t n, int64_t ldk, const double* x, double* y,
                  void* scratch, int64_t scratch_bytes);
int pb_trsv(pb_stream_t stream, const double* L, int64_t n, int64_t ldl, const void* potrf_workspace,
            int32_t trans, double* rhs, double* x);
int pb_logdet_chol(pb_stream_t stream, const double* L, int64_t n, int64_t ldl, double* out);

/* X[m x n] <- X L^{-T} (right solve with the lower factor), the many-RHS form used by predict
 * (replaces B.solve(K, Kfs) at approximators.py:177).  Uses the block inverses potrf left in
 * `workspace`.                                                                                   */
int pb_trsm_right_lt(pb_stream_t stream, const double* L, int64_t n, int64_t ldl, const void* potrf_workspace,
                     double* X, int64_t m, int64_t ldx);

/* ---- A4-A9: fused fit drivers -----------------------------------------------------------------
 * pb_laplace_fit: LaplaceGP.weight + precision (approximators.py:204-210,265-277) = Newton on
 *   g(Kw) - w = 0 with jaxopt's stopping rule (solvers.py:7-25): w0 = 0; repeat w+ = Newton(w);
 *   err = ||w+ - w||_2; until err <= tol or iters == maxiter.  Synchronises `stream` once per
 *   iteration (8-byte readback of err).  Each step solves with B = I + W^1/2 K W^1/2.  Below
 *   "laplace_pcg_min_n" B is factored every step.  Above it the solve is CG until the error left in the
 *   step is <= 1e-2 * tolerance, preconditioned by a rank-r Nystrom approximation of K rebuilt for the step's W (no N^3 work);
 *   should that stall, B is factored once and later steps run PCG on the stale factor (refactoring if
 *   that stalls too).  pb_fit_result.factorizations / pcg_iterations report what happened.  Outputs: weight w (n), precision p = -h(Kw) (n),
 *   posterior mean f = K w (n); when `final_factor` != 0 the workspace additionally ends holding
 *   the Cholesky factor of B(w*) used by pb_laplace_objective / pb_predict.
 * pb_vb_fit: VBGP.weight + precision (approximators.py:332-339; VB.py:4-16).                     */
typedef struct pb_fit_result {
    int32_t iterations;
    int32_t info;       /* potrf info of the last factorisation (0 = ok) */
    double error;       /* last ||w+ - w||_2 */
    double sum_ll;      /* sum_i ll(f_i) at the returned w */
    double ftw;         /* f^T w at the returned w */
    double logdet;      /* sum log diag chol(...) of the final factor (if computed) */
    int32_t factorizations; /* Cholesky factorisations performed */
    int32_t pcg_iterations; /* total PCG iterations of the Newton steps that reused a stale factor */
} pb_fit_result;

typedef struct pb_problem {
    const double* X;       /* (n, D) row-major training inputs (device) */
    const void* y;         /* int64 labels or f64 targets (device) */
    int64_t n;
    int32_t D;
    int32_t _pad;
    pb_kernel_spec kernel;
    pb_likelihood_spec lik;
} pb_problem;

/* Workspace layout is private; size it with pb_fit_workspace_bytes(n, D).  It holds K (n x ld),
 * the factor buffer (n x ld), features and O(n) vectors. */
int64_t pb_fit_workspace_bytes(int64_t n, int D);
/* Build features + K(theta) into the workspace (what every fit does first); and locate K inside it. */
int pb_build_gram(pb_stream_t stream, const pb_problem* prob, void* workspace, int64_t workspace_bytes);
/* Features of the training inputs only (all that a mean-only pb_predict with variance == NULL needs). */
int pb_build_features(pb_stream_t stream, const pb_problem* prob, void* workspace, int64_t workspace_bytes);
int pb_workspace_gram(void* workspace, int64_t n, int D, double** K, int64_t* ldk);
int pb_laplace_fit(pb_stream_t stream, const pb_problem* prob, double tolerance, int32_t maxiter,
                   double jitter, int32_t final_factor, void* workspace, int64_t workspace_bytes,
                   double* weight, double* precision, double* posterior_mean, pb_fit_result* result_host,
                   const pb_options* options);
int pb_vb_fit(pb_stream_t stream, const pb_problem* prob, double tolerance, int32_t maxiter,
              void* workspace, int64_t workspace_bytes, double* weight, double* precision,
              double* posterior_mean, pb_fit_result* result_host, const pb_options* options);

/* ---- A11: evidence gradient ("next" row 1) ------------------------------------------------------
 * d objective_LA / d (kernel scale, kernel stretch_out, noise std) at the converged weight: the closed form
 * (Rasmussen & Williams Alg. 5.1 / §5.5.1) of what JAX's reverse pass through fixed_point_layer returns
 * (probit/implicit/solvers.py:28-64, probit/approximators.py:132-134).  Call right after
 * pb_laplace_fit(final_factor = 1): the workspace must hold K and the factor of B(w*); the factor is
 * consumed.  grad_host[0..2] = d/dscale, d/dstretch_out, d/dsigma; for the ordinal likelihood, when
 * grad_len >= 3 + J + 1, grad_host[3 + j] = d/db_j (0 for the infinite end cutpoints), otherwise the
 * ordinal sigma slot is NaN.
 * Cost: one N x N triangular inverse + U U^T (both DMMA GEMMs) + two passes over K.                  */
int64_t pb_gradient_scratch_bytes(int64_t n);
int pb_laplace_gradient(pb_stream_t stream, const pb_problem* prob, void* workspace, int64_t workspace_bytes,
                        const double* weight, const double* precision, void* scratch, int64_t scratch_bytes,
                        double* grad_host, int32_t grad_len, const pb_options* options);
/* The same for objective_VB at the fixed point of f_VB (VBGP.value_and_grad, approximators.py:132-134,316-330;
 * VB.py:4-40): closed form of the implicit-function gradient (oracle/gradients.py::vb_gradient).  Factors
 * sigma^2 I + K and sigma I + W^1/2 K W^1/2 in the workspace (any factor held there is overwritten).
 * grad_host layout as above; scratch as pb_gradient_scratch_bytes(n). */
int pb_vb_gradient(pb_stream_t stream, const pb_problem* prob, void* workspace, int64_t workspace_bytes,
                   const double* weight, void* scratch, int64_t scratch_bytes, double* grad_host, int32_t grad_len,
                   const pb_options* options);

/* ---- A10: predict -----------------------------------------------------------------------------
 * Approximator.predict (approximators.py:154-180): mean = K_*f w,
 * var = k_** - diag(K_*f (K + P^-1)^-1 K_f*) = k_** - || L_B^-1 (s o k_*) ||^2, s = sqrt(P).
 * pb_predict_prepare builds K (skipped when reuse_gram != 0 and the workspace still holds K for
 * these prior parameters) and factors B = I + s s^T o K into the workspace (one potrf);
 * pb_predict then streams test points in chunks of `chunk` rows through a caller scratch buffer
 * of pb_predict_scratch_bytes(n, chunk) (cross-covariance tiles are generated on the fly and
 * never exist for more than one chunk).                                                         */
int pb_predict_prepare(pb_stream_t stream, const pb_problem* prob, const double* precision,
                       int32_t reuse_gram, void* workspace, int64_t workspace_bytes, int32_t* info_host,
                       const pb_options* options);
int64_t pb_predict_scratch_bytes(int64_t n, int D, int64_t chunk);
int pb_predict(pb_stream_t stream, const pb_problem* prob, const void* workspace, const double* weight,
               const double* X_test, int64_t n_test, int64_t chunk, void* scratch, int64_t scratch_bytes,
               double* mean, double* variance);

/* Approximator.predict_covariance (approximators.py:182-197): full n_test x n_test posterior covariance
 * K_** - V V^T, V = (s o K_*f) L_B^-T; same preconditions as pb_predict; scratch sized by
 * pb_predict_scratch_bytes(n, D, n_test). */
int pb_predict_covariance(pb_stream_t stream, const pb_problem* prob, const void* workspace,
                          const double* X_test, int64_t n_test, void* scratch, int64_t scratch_bytes,
                          double* cov, int64_t ldc);

/* ---- ONE fit + predict partitioned over the GPUs of a box -----------------------------------
 * One process per GPU.  The reference has no distributed code; these entry points partition
 * exactly the path above: `LaplaceGP.approximate_posterior` (approximators.py:204-210 over solvers.py:18-25,
 * Laplace.py:4-9) and `Approximator.predict` (approximators.py:154-180) for ONE training set of n points.
 *
 * pb_comm wraps an NCCL communicator (libnccl.so.2 is reached through dlopen, the single-GPU entry points do not
 * need it) plus the library's communication / look-ahead streams.  Rank 0 calls pb_comm_unique_id, the 128 bytes
 * travel to the other ranks by any means (torch.distributed, MPI, a file), every rank calls pb_comm_create on its
 * own device.  world == 1 needs neither NCCL nor an id.  Every collective is enqueued on a CUDA stream from inside
 * the library; every rank must make the same sequence of pb_dist_* calls.
 *
 * pb_dist_laplace_fit: the Newton loop of pb_laplace_fit with the rows of K sharded (each rank generates and keeps
 *   K[lo:hi, :] only), y = K x as a local row-block product + one in-place all-gather of 8n bytes, and the Nystrom
 *   preconditioner split by columns (one r x r all-reduce per Newton step).  Outputs (weight, precision, posterior
 *   mean: n doubles each) are complete and identical on every rank.
 * pb_dist_predict: factors B = I + P^1/2 (K + jitter I) P^1/2 in a 1-D block-column-cyclic layout (every rank fills
 *   and keeps only its own block columns; panels travel by ncclBroadcast, overlapped with the trailing DMMA
 *   updates) and, in the same sweep over the panels, solves for this rank's `n_test` LOCAL test points (test points
 *   are sharded by the caller; there is no collective on the data path).  `chunk` rows of s o k(X*, X) are in flight at
 *   a time (scratch: pb_dist_predict_scratch_bytes(n, D, min(chunk, n_test))); further chunks re-stream the stored
 *   panels.  reuse_factor != 0: the workspace still holds the factor of a previous call with the same (theta,
 *   precision) — skip the factorisation.  n_test may be 0 (factor + log-determinant only: the objective).
 *   logdet_host (may be NULL) receives sum log diag chol(B), the last two terms of objective_LA (Laplace.py:24-30).
 *   variance == NULL: means only (no factorisation unless logdet_host is given).                                  */
typedef struct pb_comm pb_comm;
int pb_comm_unique_id(void* id_host_128_bytes);
int pb_comm_create(const void* id_host_128_bytes, int32_t rank, int32_t world, pb_comm** comm);
int pb_comm_destroy(pb_comm* comm);
int pb_comm_rank(const pb_comm* comm);
int pb_comm_size(const pb_comm* comm);
int64_t pb_dist_workspace_bytes(int64_t n, int D, int32_t world, int32_t rank, const pb_options* options);
int pb_dist_laplace_fit(pb_stream_t stream, pb_comm* comm, const pb_problem* prob, double tolerance, int32_t maxiter,
                        void* workspace, int64_t workspace_bytes, double* weight, double* precision,
                        double* posterior_mean, pb_fit_result* result_host, const pb_options* options);
int64_t pb_dist_predict_scratch_bytes(int64_t n, int D, int64_t rows);
int pb_dist_predict(pb_stream_t stream, pb_comm* comm, const pb_problem* prob, const double* precision,
                    const double* weight, void* workspace, int64_t workspace_bytes, double jitter, int32_t reuse_factor,
                    const double* X_test, int64_t n_test, int64_t chunk, void* scratch, int64_t scratch_bytes,
                    double* mean, double* variance, double* logdet_host, int32_t* info_host, const pb_options* options);

/* ---- K15: ordinal predictive distributions (probit/utilities.py:232-249) ----------------------
 * out[n_test x J] = Phi((b[j+1]-m)/s) - Phi((b[j]-m)/s), s = sqrt(var + sigma^2).                */
int pb_predictive_distributions(pb_stream_t stream, const pb_likelihood_spec* lik, const double* mean,
                                const double* variance, int64_t n_test, double* out);

#ifdef __cplusplus
}
#endif
#endif /* PROBIT_B200_H */
