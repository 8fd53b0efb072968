// Level-2 kernels of the Newton / fixed-point iterations (all HBM-bound):
//   symv   y = K x           probit/implicit/Laplace.py:8,22, VB.py:9,23, approximators.py:273,338
//   trsv   L x = b, L^T x = b   B.cholesky_solve at VB.py:11 and the SPD Newton step replacing
//                               jnp.linalg.solve at solvers.py:24
//   logdet sum_i log L_ii    Laplace.py:28, VB.py:28
// The triangular solves retire 256 unknowns per launch: all CTAs stream the 256-wide panel of the factor
// once (coalesced 16-byte loads) and one CTA additionally solves the next 256x256 diagonal block with
// the 64x64 leaf inverses potrf left in its workspace, so a solve is n/256 launches and one pass over L.
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

constexpr int TB = 256;        // rows/columns retired per triangular-solve step (4 leaves)

// Solve the TB x TB diagonal block that starts at row/col j0 against the vector v (length nv <= TB,
// staged in shared memory as sv) using the 64x64 leaf inverses potrf left behind:
//   forward  (trans = false):  x_q = Dinv_q (v_q - sum_{p<q} L_qp x_p),  q = 0..3
//   backward (trans = true):   x_q = Dinv_q^T (v_q - sum_{p>q} L_pq^T x_p),  q = 3..0
// 256 threads: 4 threads per row of the current 64-row block.  Result left in sx (and sv is clobbered).
__device__ void diag_block_solve(const double* __restrict__ L, int64_t ldl, const double* __restrict__ dinv, int64_t j0,
                                 int nv, bool trans, double* sv, double* sx, double* st) {
    const int r = threadIdx.x >> 2, q4 = threadIdx.x & 3;
    const int nq = (nv + LEAF - 1) / LEAF;
    for (int step = 0; step < nq; ++step) {
        const int q = trans ? nq - 1 - step : step;
        // t = v_q - sum_p (L_qp x_p)  |  v_q - sum_p (L_pq^T x_p)
        double acc = 0.0;
        if (!trans) {
            for (int p = 0; p < q; ++p) {
                const double* blk = L + (j0 + q * LEAF + r) * ldl + j0 + p * LEAF;     // row r of L_qp
                if (q * LEAF + r < nv)
                    for (int c = q4; c < LEAF; c += 4) acc = fma(blk[c], sx[p * LEAF + c], acc);
            }
        } else {
            for (int p = q + 1; p < nq; ++p) {
                // (L_pq^T x_p)[r] = sum_c L[j0 + p*64 + c][j0 + q*64 + r] x_p[c]
                const double* blk = L + (j0 + p * LEAF) * ldl + j0 + q * LEAF + r;
                for (int c = q4; c < LEAF; c += 4)
                    if (p * LEAF + c < nv) acc = fma(blk[c * ldl], sx[p * LEAF + c], acc);
            }
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        if (q4 == 0) st[r] = sv[q * LEAF + r] - acc;
        __syncthreads();
        // x_q = Dinv_q t   |  Dinv_q^T t
        const double* di = dinv + (j0 / LEAF + q) * LEAF * LEAF;
        double xs = 0.0;
        if (!trans) {
            for (int c = q4; c <= r; c += 4) xs = fma(di[r * LEAF + c], st[c], xs);
        } else {
            for (int c = r + q4; c < LEAF; c += 4) xs = fma(di[c * LEAF + r], st[c], xs);
        }
        xs += __shfl_xor_sync(0xffffffffu, xs, 1);
        xs += __shfl_xor_sync(0xffffffffu, xs, 2);
        if (q4 == 0) sx[q * LEAF + r] = (q * LEAF + r < nv) ? xs : 0.0;
        __syncthreads();
    }
}

// Forward substitution, one launch per TB-wide block column (look-ahead fused in):
//   every CTA applies  b[r] -= L[r, j0:j0+TB] . x_j  to its rows (x_j was produced by the previous launch);
//   CTA 0 owns the TB rows of the NEXT diagonal block and, once they are final, solves that block and
//   publishes x_{j+1}, so the next launch can start immediately.  `j0 < 0` runs only the initial solve.
__global__ void __launch_bounds__(256)
trsv_fwd_kernel(const double* __restrict__ L, int64_t n, int64_t ldl, const double* __restrict__ dinv, int64_t j0,
                double* __restrict__ b, double* __restrict__ x) {
    __shared__ double sx[TB], sv[TB], st[LEAF];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (j0 < 0) {                                   // bootstrap: x_0 = L_00^{-1} b_0
        const int nv = (int)(n < TB ? n : TB);
        sv[threadIdx.x] = threadIdx.x < nv ? b[threadIdx.x] : 0.0;
        __syncthreads();
        diag_block_solve(L, ldl, dinv, 0, nv, false, sv, sx, st);
        if (threadIdx.x < nv) x[threadIdx.x] = sx[threadIdx.x];
        return;
    }
    sx[threadIdx.x] = x[j0 + threadIdx.x];          // block j is full width whenever rows remain below it
    __syncthreads();
    const int64_t next0 = j0 + TB;
    // CTA 0: the TB rows of the next diagonal block (32 per warp); CTA c >= 1: 64 rows (8 per warp)
    const int64_t row_begin = blockIdx.x == 0 ? next0 : next0 + TB + (int64_t)(blockIdx.x - 1) * 64;
    const int rows_here = blockIdx.x == 0 ? TB : 64;
    const int per_warp = rows_here / 8;
    double xr[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        xr[2 * k] = sx[2 * lane + 64 * k];
        xr[2 * k + 1] = sx[2 * lane + 64 * k + 1];
    }
    for (int rr = 0; rr < per_warp; rr += 4) {
        double acc[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t r = row_begin + warp * per_warp + rr + u;
            acc[u] = 0.0;
            if (r < n) {
                const double2* lp = reinterpret_cast<const double2*>(L + r * ldl + j0) + lane;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const double2 v = __ldcs(lp + 32 * k);
                    acc[u] = fma(v.x, xr[2 * k], acc[u]);
                    acc[u] = fma(v.y, xr[2 * k + 1], acc[u]);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const double sum = warp_sum(acc[u]);
            const int64_t r = row_begin + warp * per_warp + rr + u;
            if (lane == 0 && r < n) b[r] -= sum;
        }
    }
    if (blockIdx.x != 0 || next0 >= n) return;
    __threadfence_block();
    __syncthreads();
    const int nv = (int)(n - next0 < TB ? n - next0 : TB);
    sv[threadIdx.x] = threadIdx.x < nv ? b[next0 + threadIdx.x] : 0.0;
    __syncthreads();
    diag_block_solve(L, ldl, dinv, next0, nv, false, sv, sx, st);
    if (threadIdx.x < nv) x[next0 + threadIdx.x] = sx[threadIdx.x];
}

// Backward substitution L^T x = y, blocks from the bottom up, same one-launch-per-step scheme:
//   every CTA applies  y[c] -= sum_{r in block j} L[j0 + r][c] x_j[r]  to its 256 columns c < j0
//   (two 128-row halves per CTA, 16-byte coalesced row reads);
//   the LAST CTA owns the columns of the next (upper) diagonal block, solves it and publishes x_{j-1}.
__global__ void __launch_bounds__(256)
trsv_bwd_kernel(const double* __restrict__ L, int64_t n, int64_t ldl, const double* __restrict__ dinv, int64_t j0,
                int bootstrap, double* __restrict__ y, double* __restrict__ x) {
    __shared__ double sx[TB], sv[TB], st[LEAF];
    __shared__ double2 part[128];
    const int nvj = (int)(n - j0 < TB ? n - j0 : TB);
    if (bootstrap) {                                // x_last = L_last^{-T} y_last
        sv[threadIdx.x] = threadIdx.x < nvj ? y[j0 + threadIdx.x] : 0.0;
        __syncthreads();
        diag_block_solve(L, ldl, dinv, j0, nvj, true, sv, sx, st);
        if (threadIdx.x < nvj) x[j0 + threadIdx.x] = sx[threadIdx.x];
        return;
    }
    sx[threadIdx.x] = threadIdx.x < nvj ? x[j0 + threadIdx.x] : 0.0;
    __syncthreads();
    const int half = threadIdx.x >> 7, cp = threadIdx.x & 127;
    const int64_t c = (int64_t)blockIdx.x * 256 + 2 * cp;          // this thread's column pair (c, c+1) < j0
    double2 acc = make_double2(0.0, 0.0);
    if (c < j0) {
        const int r_lo = half * 128, r_hi = min(nvj, r_lo + 128);
        const double* lp = L + (j0 + r_lo) * ldl + c;
#pragma unroll 8
        for (int r = r_lo; r < r_hi; ++r) {
            const double2 v = __ldcs(reinterpret_cast<const double2*>(lp));
            acc.x = fma(v.x, sx[r], acc.x);
            acc.y = fma(v.y, sx[r], acc.y);
            lp += ldl;
        }
    }
    if (half == 1) part[cp] = acc;
    __syncthreads();
    if (half == 0 && c < j0) {
        const double2 o = part[cp];
        double2* yp = reinterpret_cast<double2*>(y + c);
        double2 cur = *yp;
        cur.x -= acc.x + o.x;
        cur.y -= acc.y + o.y;
        *yp = cur;
    }
    if (blockIdx.x != gridDim.x - 1) return;
    __threadfence_block();
    __syncthreads();
    const int64_t p0 = j0 - TB;                      // previous (upper) block, always full width
    sv[threadIdx.x] = y[p0 + threadIdx.x];
    __syncthreads();
    diag_block_solve(L, ldl, dinv, p0, TB, true, sv, sx, st);
    x[p0 + threadIdx.x] = sx[threadIdx.x];
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
// n/256 launches per solve; the factor is streamed once (8 * n^2 / 2 bytes).
int trsv(cudaStream_t stream, const double* L, int64_t n, int64_t ldl, const double* dinv, bool trans, double* rhs,
         double* x) {
    if (n == 0) return PB_OK;
    PB_CHECK((ldl & 1) == 0 && (reinterpret_cast<uintptr_t>(L) & 15) == 0, PB_ERR_INVALID,
             "trsv: factor must be 16-byte aligned with an even leading dimension");
    const int64_t nblk = ceil_div<int64_t>(n, TB);
    if (!trans) {
        trsv_fwd_kernel<<<1, 256, 0, stream>>>(L, n, ldl, dinv, -1, rhs, x); pb::note_launch();
        for (int64_t jb = 0; jb + 1 < nblk; ++jb) {
            const int64_t j0 = jb * TB;
            const int64_t below_next = n - j0 - 2 * TB;      // rows below the next diagonal block
            const unsigned grid = 1u + (below_next > 0 ? (unsigned)ceil_div<int64_t>(below_next, 64) : 0u);
            trsv_fwd_kernel<<<grid, 256, 0, stream>>>(L, n, ldl, dinv, j0, rhs, x); pb::note_launch();
        }
    } else {
        const int64_t jlast = (nblk - 1) * TB;
        trsv_bwd_kernel<<<1, 256, 0, stream>>>(L, n, ldl, dinv, jlast, 1, rhs, x); pb::note_launch();
        for (int64_t jb = nblk - 1; jb >= 1; --jb) {
            const int64_t j0 = jb * TB;
            trsv_bwd_kernel<<<(unsigned)(j0 / 256), 256, 0, stream>>>(L, n, ldl, dinv, j0, 0, rhs, x); pb::note_launch();
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
