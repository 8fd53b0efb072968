// Level-2 kernels of the Newton / fixed-point iterations (all HBM-bound):
//   symv   y = K x           probit/implicit/Laplace.py:8,22, VB.py:9,23, approximators.py:273,338
//   trsv   L x = b, L^T x = b   B.cholesky_solve at VB.py:11 and the SPD Newton step replacing
//                               jnp.linalg.solve at solvers.py:24
//   logdet sum_i log L_ii    Laplace.py:28, VB.py:28
// The triangular solves walk the factor in 64-wide block steps; the 64x64 diagonal solves reuse
// the leaf inverses potrf left in its workspace, recomputed redundantly by every CTA of a step so
// that one launch per step suffices.
#include "common.cuh"

namespace pb {

namespace {

constexpr int LEAF = 64;

// y[r] = sum_c A[r][c] x[c]; 4 rows per CTA, 256 threads stride the columns with 16-byte loads.
__global__ void __launch_bounds__(256)
gemv_kernel(const double* __restrict__ A, int64_t rows, int64_t cols, int64_t lda, const double* __restrict__ x,
            double* __restrict__ y) {
    const int64_t r0 = (int64_t)blockIdx.x * 4;
    double acc[4] = {0, 0, 0, 0};
    const bool vec = ((lda & 1) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    if (vec) {
        const int64_t c2 = cols >> 1;
        for (int64_t j = threadIdx.x; j < c2; j += 256) {
            const double2 xv = reinterpret_cast<const double2*>(x)[j];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                if (r0 + r < rows) {
                    const double2 a = __ldcs(reinterpret_cast<const double2*>(A + (r0 + r) * lda) + j);
                    acc[r] = fma(a.x, xv.x, acc[r]);
                    acc[r] = fma(a.y, xv.y, acc[r]);
                }
            }
        }
        if ((cols & 1) && threadIdx.x == 0) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
                if (r0 + r < rows) acc[r] = fma(A[(r0 + r) * lda + cols - 1], x[cols - 1], acc[r]);
        }
    } else {
        for (int64_t j = threadIdx.x; j < cols; j += 256) {
            const double xv = x[j];
#pragma unroll
            for (int r = 0; r < 4; ++r)
                if (r0 + r < rows) acc[r] = fma(A[(r0 + r) * lda + j], xv, acc[r]);
        }
    }
    __shared__ double red[4][8];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const double v = warp_sum(acc[r]);
        if ((threadIdx.x & 31) == 0) red[r][threadIdx.x >> 5] = v;
    }
    __syncthreads();
    if (threadIdx.x < 4 && r0 + threadIdx.x < rows) {
        double s = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[threadIdx.x][w];
        y[r0 + threadIdx.x] = s;
    }
}

// One forward step of L x = b for the 64-wide block starting at j0:
//   x_j = Dinv_j * b_j (every CTA recomputes it; CTA 0 publishes it), then
//   b[r] -= L[r, j0:j0+64] . x_j for the rows r > j0+63 owned by this CTA (256 rows per CTA).
__global__ void __launch_bounds__(256)
trsv_fwd_step_kernel(const double* __restrict__ L, int64_t n, int64_t ldl, const double* __restrict__ dinv, int64_t j0,
                     double* __restrict__ b, double* __restrict__ x) {
    __shared__ double bj[LEAF], xj[LEAF];
    const int nv = (int)(n - j0 < LEAF ? n - j0 : LEAF);
    if (threadIdx.x < LEAF) bj[threadIdx.x] = threadIdx.x < nv ? b[j0 + threadIdx.x] : 0.0;
    __syncthreads();
    {   // 64x64 lower-triangular matvec: 4 threads per row
        const int r = threadIdx.x >> 2, q = threadIdx.x & 3;
        double s = 0;
        for (int c = q; c <= r; c += 4) s = fma(dinv[r * LEAF + c], bj[c], s);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (q == 0) xj[r] = s;
    }
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x < nv) x[j0 + threadIdx.x] = xj[threadIdx.x];
    // trailing update: 8 warps, each warp takes rows; lanes stride the 64 columns (2 per lane)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row_base = j0 + LEAF + (int64_t)blockIdx.x * 256;
    const double x0 = xj[2 * lane], x1 = xj[2 * lane + 1];
    for (int rr = warp; rr < 256; rr += 8) {
        const int64_t r = row_base + rr;
        if (r >= n) break;
        const double* lp = L + r * ldl + j0 + 2 * lane;
        double s = fma(lp[0], x0, lp[1] * x1);
        s = warp_sum(s);
        if (lane == 0) b[r] -= s;
    }
}

// One backward step of L^T x = y for the block starting at j0:
//   x_j = Dinv_j^T * y_j, then y[c] -= sum_{r in block} L[j0+r][c] x_j[r] for the columns c < j0
//   owned by this CTA (256 columns per CTA, one per thread: coalesced row reads).
__global__ void __launch_bounds__(256)
trsv_bwd_step_kernel(const double* __restrict__ L, int64_t n, int64_t ldl, const double* __restrict__ dinv, int64_t j0,
                     double* __restrict__ y, double* __restrict__ x) {
    __shared__ double yj[LEAF], xj[LEAF];
    const int nv = (int)(n - j0 < LEAF ? n - j0 : LEAF);
    if (threadIdx.x < LEAF) yj[threadIdx.x] = threadIdx.x < nv ? y[j0 + threadIdx.x] : 0.0;
    __syncthreads();
    {   // x[c] = sum_{r >= c} Dinv[r][c] y[r]: 4 threads per column
        const int c = threadIdx.x >> 2, q = threadIdx.x & 3;
        double s = 0;
        for (int r = c + q; r < LEAF; r += 4) s = fma(dinv[r * LEAF + c], yj[r], s);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (q == 0) xj[c] = s;
    }
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x < nv) x[j0 + threadIdx.x] = xj[threadIdx.x];
    const int64_t c = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (c < j0) {
        double s = 0;
#pragma unroll 8
        for (int r = 0; r < nv; ++r) s = fma(L[(j0 + r) * ldl + c], xj[r], s);
        y[c] -= s;
    }
}

__global__ void __launch_bounds__(1024)
logdet_kernel(const double* __restrict__ L, int64_t n, int64_t ldl, double* __restrict__ out) {
    double s = 0;
    for (int64_t i = threadIdx.x; i < n; i += 1024) s += log(L[i * ldl + i]);
    s = block_sum<1024>(s);
    if (threadIdx.x == 0) out[0] = s;
}

}  // namespace

int gemv(cudaStream_t stream, const double* A, int64_t rows, int64_t cols, int64_t lda, const double* x, double* y) {
    if (rows == 0) return PB_OK;
    gemv_kernel<<<(unsigned)ceil_div<int64_t>(rows, 4), 256, 0, stream>>>(A, rows, cols, lda, x, y); pb::note_launch();
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

// Solves with the lower factor; `rhs` is destroyed, the solution goes to `x` (may not alias rhs).
int trsv(cudaStream_t stream, const double* L, int64_t n, int64_t ldl, const double* dinv, bool trans, double* rhs,
         double* x) {
    if (n == 0) return PB_OK;
    const int64_t nblk = ceil_div<int64_t>(n, LEAF);
    if (!trans) {
        for (int64_t jb = 0; jb < nblk; ++jb) {
            const int64_t j0 = jb * LEAF;
            const int64_t below = n - j0 - LEAF;
            const unsigned grid = below > 0 ? (unsigned)ceil_div<int64_t>(below, 256) : 1u;
            trsv_fwd_step_kernel<<<grid, 256, 0, stream>>>(L, n, ldl, dinv + jb * LEAF * LEAF, j0, rhs, x); pb::note_launch();
        }
    } else {
        for (int64_t jb = nblk - 1; jb >= 0; --jb) {
            const int64_t j0 = jb * LEAF;
            const unsigned grid = j0 > 0 ? (unsigned)ceil_div<int64_t>(j0, 256) : 1u;
            trsv_bwd_step_kernel<<<grid, 256, 0, stream>>>(L, n, ldl, dinv + jb * LEAF * LEAF, j0, rhs, x); pb::note_launch();
        }
    }
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

int logdet_chol(cudaStream_t stream, const double* L, int64_t n, int64_t ldl, double* out) {
    logdet_kernel<<<1, 1024, 0, stream>>>(L, n, ldl, out); pb::note_launch();
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

}  // namespace pb

extern "C" int pb_symv(pb_stream_t stream, const double* K, int64_t n, int64_t ldk, const double* x, double* y) {
    return pb::gemv(reinterpret_cast<cudaStream_t>(stream), K, n, n, ldk, x, y);
}

extern "C" int pb_trsv(pb_stream_t stream, const double* L, int64_t n, int64_t ldl, const void* potrf_workspace,
                       int32_t trans, double* rhs, double* x) {
    return pb::trsv(reinterpret_cast<cudaStream_t>(stream), L, n, ldl,
                    reinterpret_cast<const double*>(potrf_workspace), trans != 0, rhs, x);
}

extern "C" int pb_logdet_chol(pb_stream_t stream, const double* L, int64_t n, int64_t ldl, double* out) {
    return pb::logdet_chol(reinterpret_cast<cudaStream_t>(stream), L, n, ldl, out);
}
