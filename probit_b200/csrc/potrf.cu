// Blocked right-looking FP64 Cholesky for sm_100a (row-major, lower).
//
// Replaces B.cholesky at probit/implicit/Laplace.py:24 and VB.py:10,25, and (through the SPD
// Newton form) the LU solve at probit/implicit/solvers.py:24.
//
//   for each block column k (width NB):
//     1. diagonal block: recursive potrf down to 64x64 leaves; a leaf is one CTA that factors the
//        block with register-resident rows and warp shuffles (32x32 at a time) and also emits the
//        inverse of the leaf (used by every TRSM as a GEMM operand);
//     2. panel TRSM  P <- P L_kk^{-T}: recursive, every flop is a DMMA GEMM (gemm_dmma.cu);
//     3. trailing SYRK  A22 -= P P^T on lower tiles only (DMMA GEMM, K = NB).
//   All launches are asynchronous on one stream; failure (non-positive pivot) lands in *info.
#include "common.cuh"

namespace pb {

namespace {

constexpr int LEAF = 64;
constexpr int LDS = LEAF + 1;   // padded smem leading dimension

// Factor a 32x32 SPD block held one row per lane (a[k] = A[lane][k], lower part valid).
// On exit a[k] = L[lane][k] for k <= lane.  Returns false if a pivot was not positive.
__device__ __forceinline__ bool warp_potrf32(double (&a)[32], int lane, int& bad_col) {
    bool ok = true;
    bad_col = -1;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const double d = __shfl_sync(0xffffffffu, a[j], j);
        if (!(d > 0.0) && ok) { ok = false; bad_col = j; }
        const double r = 1.0 / sqrt(d);
        const double l = (lane == j) ? sqrt(d) : a[j] * r;
        a[j] = l;
#pragma unroll
        for (int k = j + 1; k < 32; ++k) {
            const double lk = __shfl_sync(0xffffffffu, l, k);
            a[k] = fma(-l, lk, a[k]);
        }
    }
    return ok;
}

// One CTA (256 threads): factor the nv x nv (nv <= 64) diagonal block at A (lda) in place and
// write the 64x64 inverse of the (identity-padded) factor to Dinv (row-major, ld 64).
__global__ void __launch_bounds__(256, 1)
potrf_leaf_kernel(double* __restrict__ A, int64_t lda, int nv, double* __restrict__ Dinv, int32_t* info, int col0) {
    extern __shared__ double leaf_smem[];
    double* S = leaf_smem;                  // working block -> L
    double* V = leaf_smem + LEAF * LDS;     // inverse
    double* rd = V + LEAF * LDS;            // reciprocal diagonal of L
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    for (int e = tid; e < LEAF * LEAF; e += 256) {
        const int i = e >> 6, k = e & 63;
        double v = 0.0;
        if (i < nv && k <= i) v = A[(int64_t)i * lda + k];
        else if (i == k) v = 1.0;
        S[i * LDS + k] = v;
        V[i * LDS + k] = 0.0;
    }
    __syncthreads();

    // --- L11 = chol(A11) : warp 0, rows in registers, pivots/columns exchanged by shuffle ---
    if (warp == 0) {
        double a[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) a[k] = S[lane * LDS + k];
        int bad;
        if (!warp_potrf32(a, lane, bad) && lane == 0) atomicCAS(info, 0, col0 + bad + 1);
#pragma unroll
        for (int k = 0; k < 32; ++k) S[lane * LDS + k] = (k <= lane) ? a[k] : 0.0;
        __syncwarp();
        rd[lane] = 1.0 / S[lane * LDS + lane];
    }
    __syncthreads();

    // --- A21 <- A21 L11^{-T} : warp 1, one row per lane, column-oriented substitution ---
    if (warp == 1) {
        double b[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) b[k] = S[(32 + lane) * LDS + k];
#pragma unroll
        for (int c = 0; c < 32; ++c) {
            const double x = b[c] * rd[c];
            b[c] = x;
#pragma unroll
            for (int k = c + 1; k < 32; ++k) b[k] = fma(-x, S[k * LDS + c], b[k]);
        }
#pragma unroll
        for (int k = 0; k < 32; ++k) S[(32 + lane) * LDS + k] = b[k];
    }
    __syncthreads();

    // --- A22 -= A21 A21^T (lower) : all threads, 4 elements each ---
    for (int e = tid; e < 32 * 32; e += 256) {
        const int i = e >> 5, j = e & 31;
        if (j <= i) {
            double s = 0.0;
#pragma unroll 8
            for (int k = 0; k < 32; ++k) s = fma(S[(32 + i) * LDS + k], S[(32 + j) * LDS + k], s);
            S[(32 + i) * LDS + 32 + j] -= s;
        }
    }
    __syncthreads();

    // --- L22 = chol(A22) : warp 0 ---
    if (warp == 0) {
        double a[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) a[k] = S[(32 + lane) * LDS + 32 + k];
        int bad;
        if (!warp_potrf32(a, lane, bad) && lane == 0) atomicCAS(info, 0, col0 + 32 + bad + 1);
#pragma unroll
        for (int k = 0; k < 32; ++k) S[(32 + lane) * LDS + 32 + k] = (k <= lane) ? a[k] : 0.0;
        __syncwarp();
        rd[32 + lane] = 1.0 / S[(32 + lane) * LDS + 32 + lane];
    }
    __syncthreads();

    // --- inverses of the two diagonal 32x32 factors: warp 0 -> inv(L11), warp 1 -> inv(L22);
    //     lane = column j of the inverse, column-oriented forward substitution on e_j ---
    if (warp < 2) {
        const int o = warp * 32;
        double b[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) b[k] = (k == lane) ? 1.0 : 0.0;
#pragma unroll
        for (int c = 0; c < 32; ++c) {
            const double x = b[c] * rd[o + c];
            b[c] = x;
#pragma unroll
            for (int k = c + 1; k < 32; ++k) b[k] = fma(-x, S[(o + k) * LDS + o + c], b[k]);
        }
#pragma unroll
        for (int k = 0; k < 32; ++k) V[(o + k) * LDS + o + lane] = b[k];
    }
    __syncthreads();

    // --- inv21 = -inv22 * (L21 * inv11): two 32^3 products through a temporary in the (unused)
    //     upper-right quadrant of V ---
    for (int e = tid; e < 32 * 32; e += 256) {
        const int i = e >> 5, j = e & 31;
        double s = 0.0;
#pragma unroll 8
        for (int k = 0; k < 32; ++k) s = fma(S[(32 + i) * LDS + k], V[k * LDS + j], s);
        V[i * LDS + 32 + j] = s;     // T = L21 * inv11 (parked in the upper-right quadrant)
    }
    __syncthreads();
    for (int e = tid; e < 32 * 32; e += 256) {
        const int i = e >> 5, j = e & 31;
        double s = 0.0;
#pragma unroll 8
        for (int k = 0; k < 32; ++k) s = fma(V[(32 + i) * LDS + 32 + k], V[k * LDS + 32 + j], s);
        V[(32 + i) * LDS + j] = -s;
    }
    __syncthreads();

    for (int e = tid; e < LEAF * LEAF; e += 256) {
        const int i = e >> 6, k = e & 63;
        if (i < nv && k <= i) A[(int64_t)i * lda + k] = S[i * LDS + k];
        Dinv[e] = (k <= i) ? V[i * LDS + k] : 0.0;
    }
}

constexpr int LEAF_SMEM = (2 * LEAF * LDS + LEAF) * 8;

struct Ctx {
    cudaStream_t stream;
    int64_t lda;
    double* dinv;     // (n/64) leaf inverses, 64*64 doubles each, indexed by global column / 64
    int32_t* info;
};

inline int64_t split(int64_t n) { return ((n / LEAF + 1) / 2) * LEAF; }   // first-half size, multiple of 64

// B[m x n] <- B * L^{-T}, L the n x n lower factor whose first column is global column col0.
int trsm_rec(const Ctx& c, double* B, int64_t ldb, int64_t m, const double* L, int64_t n, int64_t col0) {
    if (m <= 0 || n <= 0) return PB_OK;
    if (n <= LEAF) {
        const double* inv = c.dinv + (col0 / LEAF) * LEAF * LEAF;
        return gemm_nt(c.stream, m, n, n, 1.0, B, ldb, inv, LEAF, 0.0, B, ldb, false);   // in place: one tile column
    }
    const int64_t n1 = split(n), n2 = n - n1;
    PB_TRY(trsm_rec(c, B, ldb, m, L, n1, col0));
    PB_TRY(gemm_nt(c.stream, m, n2, n1, -1.0, B, ldb, L + n1 * c.lda, c.lda, 1.0, B + n1, ldb, false));
    return trsm_rec(c, B + n1, ldb, m, L + n1 * c.lda + n1, n2, col0 + n1);
}

int potrf_rec(const Ctx& c, double* A, int64_t n, int64_t col0) {
    if (n <= LEAF) {
        static bool configured = false;
        if (!configured) {
            PB_CUDA(cudaFuncSetAttribute(potrf_leaf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LEAF_SMEM));
            configured = true;
        }
        potrf_leaf_kernel<<<1, 256, LEAF_SMEM, c.stream>>>(A, c.lda, (int)n, c.dinv + (col0 / LEAF) * LEAF * LEAF,
                                                           c.info, (int)col0);
        PB_CUDA(cudaGetLastError());
        return PB_OK;
    }
    const int64_t n1 = split(n), n2 = n - n1;
    PB_TRY(potrf_rec(c, A, n1, col0));
    double* A21 = A + n1 * c.lda;
    PB_TRY(trsm_rec(c, A21, c.lda, n2, A, n1, col0));
    PB_TRY(gemm_nt(c.stream, n2, n2, n1, -1.0, A21, c.lda, A21, c.lda, 1.0, A21 + n1, c.lda, true));
    return potrf_rec(c, A21 + n1, n2, col0 + n1);
}

}  // namespace

int potrf_block_size(int64_t n) {
    if (n <= 2048) return 256;
    return 512;
}

int potrf(cudaStream_t stream, double* A, int64_t n, int64_t lda, void* workspace, int64_t workspace_bytes,
          int32_t* info) {
    PB_CHECK(n >= 0 && lda >= n, PB_ERR_INVALID, "potrf: bad n/lda");
    PB_CHECK(workspace_bytes >= pb_potrf_workspace_bytes(n), PB_ERR_INVALID, "potrf: workspace too small");
    PB_CUDA(cudaMemsetAsync(info, 0, sizeof(int32_t), stream));
    if (n == 0) return PB_OK;
    Ctx c{stream, lda, reinterpret_cast<double*>(workspace), info};
    const int64_t NB = potrf_block_size(n);
    for (int64_t k0 = 0; k0 < n; k0 += NB) {
        const int64_t nb = n - k0 < NB ? n - k0 : NB;
        double* Akk = A + k0 * lda + k0;
        PB_TRY(potrf_rec(c, Akk, nb, k0));
        const int64_t m = n - k0 - nb;
        if (m > 0) {
            double* P = A + (k0 + nb) * lda + k0;
            PB_TRY(trsm_rec(c, P, lda, m, Akk, nb, k0));
            PB_TRY(gemm_nt(stream, m, m, nb, -1.0, P, lda, P, lda, 1.0, P + nb, lda, true));
        }
    }
    return PB_OK;
}

int trsm_right_lt(cudaStream_t stream, const double* L, int64_t n, int64_t ldl, const void* potrf_workspace,
                  double* X, int64_t m, int64_t ldx) {
    Ctx c{stream, ldl, const_cast<double*>(reinterpret_cast<const double*>(potrf_workspace)), nullptr};
    return trsm_rec(c, X, ldx, m, L, n, 0);
}

}  // namespace pb

extern "C" int64_t pb_potrf_workspace_bytes(int64_t n) {
    const int64_t leaves = (n + pb::LEAF - 1) / pb::LEAF;
    return (leaves > 0 ? leaves : 1) * pb::LEAF * pb::LEAF * (int64_t)sizeof(double);
}

extern "C" int pb_potrf(pb_stream_t stream, double* A, int64_t n, int64_t lda, void* workspace,
                        int64_t workspace_bytes, int32_t* info) {
    return pb::potrf(reinterpret_cast<cudaStream_t>(stream), A, n, lda, workspace, workspace_bytes, info);
}

extern "C" int pb_trsm_right_lt(pb_stream_t stream, const double* L, int64_t n, int64_t ldl,
                                const void* potrf_workspace, double* X, int64_t m, int64_t ldx) {
    return pb::trsm_right_lt(reinterpret_cast<cudaStream_t>(stream), L, n, ldl, potrf_workspace, X, m, ldx);
}
