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

// Inverses of the TB x TB diagonal blocks of the factor, X_j = L_jj^{-1} (row-major, lower), assembled
// once per factorisation from the 64x64 leaf inverses:  X_qq = Dinv_q,
// X_qp = -Dinv_q * sum_{r=p}^{q-1} L_qr X_rp  (q > p).  One CTA per diagonal block; 64^3 products on
// register tiles (4x4 per thread) with operands staged in shared memory.  With these, every
// triangular-solve step is a single 256x256 matvec instead of a chain of dependent 64-wide substitutions.
constexpr int TLD = LEAF + 1;
constexpr int TBINV_SMEM = 5 * LEAF * TLD * 8;

__device__ __forceinline__ void mm64_acc(const double* A, const double* B, int ti, int tk, double (&acc)[4][4]) {
    for (int m = 0; m < LEAF; ++m) {
        double a[4], b[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) a[u] = A[(ti + 16 * u) * TLD + m];
#pragma unroll
        for (int v = 0; v < 4; ++v) b[v] = B[m * TLD + tk + 16 * v];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int v = 0; v < 4; ++v) acc[u][v] = fma(a[u], b[v], acc[u][v]);
    }
}

__global__ void __launch_bounds__(256)
tb_inverse_kernel(const double* __restrict__ L, int64_t n, int64_t ldl, const double* __restrict__ dinv,
                  double* __restrict__ tinv, double* __restrict__ tinv_t) {
    extern __shared__ double tsm[];
    double* Xc = tsm;                          // X_rp for r = p, p+1, p+2   [3][64][65]
    double* A = tsm + 3 * LEAF * TLD;          // staged L_qr / Dinv_q        [64][65]
    double* T = A + LEAF * TLD;                // sum_r L_qr X_rp              [64][65]
    const int64_t j0 = (int64_t)blockIdx.x * TB;
    const int nv = (int)(n - j0 < TB ? n - j0 : TB);
    const int nq = (nv + LEAF - 1) / LEAF;
    double* out = tinv + (int64_t)blockIdx.x * TB * TB;
    double* out_t = tinv_t + (int64_t)blockIdx.x * TB * TB;      // X^T (upper), so the backward solve reads rows too
    const int tid = threadIdx.x, ti = tid >> 4, tk = tid & 15;
    for (int e = tid; e < TB * TB; e += 256) { out[e] = 0.0; out_t[e] = 0.0; }
    __syncthreads();
    for (int p = 0; p < nq; ++p) {
        const double* dp = dinv + (j0 / LEAF + p) * LEAF * LEAF;
        for (int e = tid; e < LEAF * LEAF; e += 256) {
            const int i = e >> 6, k = e & 63;
            const double v = dp[e];
            Xc[i * TLD + k] = v;
            out[(p * LEAF + i) * TB + p * LEAF + k] = v;
            out_t[(p * LEAF + k) * TB + p * LEAF + i] = v;
        }
        __syncthreads();
        for (int q = p + 1; q < nq; ++q) {
            double acc[4][4] = {};
            for (int r = p; r < q; ++r) {
                for (int e = tid; e < LEAF * LEAF; e += 256) {
                    const int i = e >> 6, k = e & 63;
                    const int64_t row = j0 + q * LEAF + i;
                    A[i * TLD + k] = (row < n) ? L[row * ldl + j0 + r * LEAF + k] : 0.0;
                }
                __syncthreads();
                mm64_acc(A, Xc + (r - p) * LEAF * TLD, ti, tk, acc);
                __syncthreads();
            }
            const double* dq = dinv + (j0 / LEAF + q) * LEAF * LEAF;
            for (int e = tid; e < LEAF * LEAF; e += 256) A[(e >> 6) * TLD + (e & 63)] = dq[e];
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int v = 0; v < 4; ++v) T[(ti + 16 * u) * TLD + tk + 16 * v] = acc[u][v];
            __syncthreads();
            double x[4][4] = {};
            mm64_acc(A, T, ti, tk, x);
            __syncthreads();
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const int i = ti + 16 * u, k = tk + 16 * v;
                    out[(q * LEAF + i) * TB + p * LEAF + k] = -x[u][v];
                    out_t[(p * LEAF + k) * TB + q * LEAF + i] = -x[u][v];
                    if (q - p <= 2) Xc[(q - p) * LEAF * TLD + i * TLD + k] = -x[u][v];
                }
            __syncthreads();
        }
    }
}

constexpr int TSV_THREADS = 1024;     // 32 warps per CTA in the triangular-solve step kernels

// x = X v (forward, X = L_jj^{-1}, lower, row-major) for one TB-block: one warp per row (coalesced row
// reads + shuffle reduction), 8 rows per warp.  v staged in shared memory (sv), result to sx.
__device__ __forceinline__ void tinv_matvec_fwd(const double* __restrict__ X, const double* sv, double* sx) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int u = 0; u < TB / 32; ++u) {
        const int r = warp + 32 * u;
        const double* row = X + (int64_t)r * TB;
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < TB / 32; ++k) {
            const int c = lane + 32 * k;
            if (c <= r) acc = fma(row[c], sv[c], acc);
        }
        acc = warp_sum(acc);
        if (lane == 0) sx[r] = acc;
    }
    __syncthreads();
}

// x = X^T v (backward) using the transposed inverse XT (upper, row-major): the same warp-per-row dot product.
__device__ __forceinline__ void tinv_matvec_bwd(const double* __restrict__ XT, const double* sv, double* sx) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int u = 0; u < TB / 32; ++u) {
        const int r = warp + 32 * u;
        const double* row = XT + (int64_t)r * TB;
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < TB / 32; ++k) {
            const int c = lane + 32 * k;
            if (c >= r) acc = fma(row[c], sv[c], acc);
        }
        acc = warp_sum(acc);
        if (lane == 0) sx[r] = acc;
    }
    __syncthreads();
}

// Forward substitution, one launch per TB-wide block column (look-ahead fused in):
//   every CTA applies  b[r] -= L[r, j0:j0+TB] . x_j  to its 256 rows, 8 per warp (x_j was produced by the
//   previous launch);  CTA 0 owns the rows of the NEXT diagonal block and, once they are final, multiplies
//   them by that block's inverse and publishes x_{j+1}, so the next launch can start immediately.
//   `j0 < 0` runs only the initial solve.
__global__ void __launch_bounds__(TSV_THREADS)
trsv_fwd_kernel(const double* __restrict__ L, int64_t n, int64_t ldl, const double* __restrict__ tinv, int64_t j0,
                double* __restrict__ b, double* __restrict__ x) {
    __shared__ double sx[TB], sv[TB];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (j0 < 0) {                                   // bootstrap: x_0 = L_00^{-1} b_0
        const int nv = (int)(n < TB ? n : TB);
        if (threadIdx.x < TB) sv[threadIdx.x] = threadIdx.x < nv ? b[threadIdx.x] : 0.0;
        __syncthreads();
        tinv_matvec_fwd(tinv, sv, sx);
        if (threadIdx.x < nv) x[threadIdx.x] = sx[threadIdx.x];
        return;
    }
    if (threadIdx.x < TB) sx[threadIdx.x] = x[j0 + threadIdx.x];     // block j is full width whenever rows remain below it
    __syncthreads();
    const int64_t next0 = j0 + TB;
    const int64_t row_begin = next0 + (int64_t)blockIdx.x * TB + warp * 8;
    double xr[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        xr[2 * k] = sx[2 * lane + 64 * k];
        xr[2 * k + 1] = sx[2 * lane + 64 * k + 1];
    }
#pragma unroll
    for (int rr = 0; rr < 8; rr += 4) {
        double acc[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t r = row_begin + rr + u;
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
            const int64_t r = row_begin + rr + u;
            if (lane == 0 && r < n) b[r] -= sum;
        }
    }
    if (blockIdx.x != 0 || next0 >= n) return;
    __syncthreads();                                 // CTA 0: its 256 rows (the next diagonal block) are final
    const int nv = (int)(n - next0 < TB ? n - next0 : TB);
    if (threadIdx.x < TB) sv[threadIdx.x] = threadIdx.x < nv ? b[next0 + threadIdx.x] : 0.0;
    __syncthreads();
    tinv_matvec_fwd(tinv + (next0 / TB) * TB * TB, sv, sx);
    if (threadIdx.x < nv) x[next0 + threadIdx.x] = sx[threadIdx.x];
}

// Backward substitution L^T x = y, blocks from the bottom up, same one-launch-per-step scheme:
//   every CTA applies  y[c] -= sum_{r in block j} L[j0 + r][c] x_j[r]  to its 256 columns c < j0
//   (eight 32-row groups per CTA, 16-byte coalesced row reads, partials combined in shared memory);
//   the LAST CTA owns the columns of the next (upper) diagonal block and publishes x_{j-1}.
__global__ void __launch_bounds__(TSV_THREADS)
trsv_bwd_kernel(const double* __restrict__ L, int64_t n, int64_t ldl, const double* __restrict__ tinv, int64_t j0,
                int bootstrap, double* __restrict__ y, double* __restrict__ x) {
    __shared__ double sx[TB], sv[TB];
    __shared__ double red[8 * TB];
    const int nvj = (int)(n - j0 < TB ? n - j0 : TB);
    if (bootstrap) {                                // x_last = L_last^{-T} y_last
        if (threadIdx.x < TB) sv[threadIdx.x] = threadIdx.x < nvj ? y[j0 + threadIdx.x] : 0.0;
        __syncthreads();
        tinv_matvec_bwd(tinv + (j0 / TB) * TB * TB, sv, sx);
        if (threadIdx.x < nvj) x[j0 + threadIdx.x] = sx[threadIdx.x];
        return;
    }
    if (threadIdx.x < TB) sx[threadIdx.x] = threadIdx.x < nvj ? x[j0 + threadIdx.x] : 0.0;
    __syncthreads();
    const int grp = threadIdx.x >> 7, cp = threadIdx.x & 127;       // 8 row groups x 128 column pairs
    const int64_t c = (int64_t)blockIdx.x * TB + 2 * cp;            // columns (c, c+1) < j0
    double2 acc = make_double2(0.0, 0.0);
    {
        const int r_lo = grp * 32, r_hi = min(nvj, r_lo + 32);
        const double* lp = L + (j0 + r_lo) * ldl + c;
#pragma unroll 16
        for (int r = r_lo; r < r_hi; ++r) {
            const double2 v = __ldcs(reinterpret_cast<const double2*>(lp));
            acc.x = fma(v.x, sx[r], acc.x);
            acc.y = fma(v.y, sx[r], acc.y);
            lp += ldl;
        }
    }
    red[grp * TB + 2 * cp] = acc.x;
    red[grp * TB + 2 * cp + 1] = acc.y;
    __syncthreads();
    if (threadIdx.x < TB) {
        double ssum = 0.0;
#pragma unroll
        for (int g = 0; g < 8; ++g) ssum += red[g * TB + threadIdx.x];
        y[(int64_t)blockIdx.x * TB + threadIdx.x] -= ssum;
    }
    if (blockIdx.x != gridDim.x - 1) return;
    __syncthreads();
    const int64_t p0 = j0 - TB;                      // previous (upper) block, always full width
    if (threadIdx.x < TB) sv[threadIdx.x] = y[p0 + threadIdx.x];
    __syncthreads();
    tinv_matvec_bwd(tinv + (p0 / TB) * TB * TB, sv, sx);
    if (threadIdx.x < TB) x[p0 + threadIdx.x] = sx[threadIdx.x];
}

__global__ void __launch_bounds__(1024)
logdet_kernel(const double* __restrict__ L, int64_t n, int64_t ldl, double* __restrict__ out) {
    double s = 0;
    for (int64_t i = threadIdx.x; i < n; i += 1024) s += log(L[i * ldl + i]);
    s = block_sum<1024>(s);
    if (threadIdx.x == 0) out[0] = s;
}


// ---- transposed matvec: partial[split][j] = sum_{i in split} A[i][j] x[i], A rows x cols row-major ----
// Column-parallel (lane = 2 adjacent columns, 16-byte loads), rows cut into gridDim.y chunks so that a short, wide
// A (the r x N Nystrom factor) still fills the machine; the caller sums the chunks in fixed order.
constexpr int GT_ROWS = 512;        // rows per chunk staged in shared memory

__global__ void __launch_bounds__(256)
gemv_t_kernel(const double* __restrict__ A, int64_t rows, int64_t cols, int64_t lda, const double* __restrict__ x,
              double* __restrict__ partial, int64_t ldp) {
    __shared__ double xs[GT_ROWS];
    const int64_t i0 = (int64_t)blockIdx.y * GT_ROWS;
    const int cnt = (int)(rows - i0 < GT_ROWS ? rows - i0 : GT_ROWS);
    for (int i = threadIdx.x; i < cnt; i += 256) xs[i] = x[i0 + i];
    __syncthreads();
    const int64_t j = (blockIdx.x * 256ll + threadIdx.x) * 2;
    if (j >= cols) return;
    const double* a = A + i0 * lda + j;                 // j + 1 < lda always: lda is even and >= cols
    double s0 = 0, s1 = 0, t0 = 0, t1 = 0;
    int i = 0;
    for (; i + 8 <= cnt; i += 8) {
        double2 v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = __ldcs(reinterpret_cast<const double2*>(a + (int64_t)(i + k) * lda));
#pragma unroll
        for (int k = 0; k < 8; k += 2) {
            s0 = fma(v[k].x, xs[i + k], s0);
            s1 = fma(v[k].y, xs[i + k], s1);
            t0 = fma(v[k + 1].x, xs[i + k + 1], t0);
            t1 = fma(v[k + 1].y, xs[i + k + 1], t1);
        }
    }
    for (; i < cnt; ++i) {
        const double2 v = __ldcs(reinterpret_cast<const double2*>(a + (int64_t)i * lda));
        s0 = fma(v.x, xs[i], s0);
        s1 = fma(v.y, xs[i], s1);
    }
    partial[blockIdx.y * ldp + j] = s0 + t0;
    if (j + 1 < cols) partial[blockIdx.y * ldp + j + 1] = s1 + t1;
}

// ---- symv that reads only the lower triangle's tiles: half the HBM traffic of the row-wise gemv ----
// K is stored in full, but y = K x only needs each off-diagonal 64x64 tile once: the tile (I, J), J < I, gives
// y_I += T x_J (kept in registers along the CTA's row strip) and y_J += T^T x_I (summed over the CTA's 64 rows and
// written to P[I][J*64 ..]).  CTA I walks J = 0..I (longest strips first); lane l of every warp owns columns
// 2l, 2l+1 of the tile and the warp owns 8 rows, so a tile is 8 coalesced 512-byte row segments per warp and
// 5 instructions per 16 bytes loaded.  The next tile's loads are issued before the current one is consumed.
// symv_reduce_kernel then forms y[j] = rowpart[j] + sum_{I > j/64} P[I][j] in a fixed order (deterministic, no
// atomics).  P costs n^2/8 bytes of scratch and 2 x n^2/16 bytes of extra traffic (1.6 % each way).
constexpr int SYT = 64;

__global__ void __launch_bounds__(256, 2)
symv_lower_kernel(const double* __restrict__ K, int64_t n, int64_t ld, const double* __restrict__ x,
                  double* __restrict__ rowpart, double* __restrict__ P, int64_t ldp) {
    const int64_t I = (int64_t)gridDim.x - 1 - blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t r0 = I * SYT + warp * 8;
    __shared__ double2 colred[2][8][32];
    double xi[8], racc[8];
    // rows past the end are redirected to the last valid row (memory-safe) and get x_i = 0 (no column contribution)
    const int64_t rb = r0 < n ? r0 : n - 1;
    const int kmax = (int)(n - rb < 8 ? n - rb : 8);
    const double* base = K + rb * ld + 2 * lane;
#define SYMV_ROW(k) (base + (int64_t)((k) < kmax ? (k) : kmax - 1) * ld)
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        xi[k] = r0 + k < n ? x[r0 + k] : 0.0;
        racc[k] = 0.0;
    }
    double2 a[8], b[8];
    if (I > 0) {
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = __ldcs(reinterpret_cast<const double2*>(SYMV_ROW(k)));
    }
    for (int64_t J = 0; J < I; ++J) {
        const double2 xj = *reinterpret_cast<const double2*>(x + J * SYT + 2 * lane);
        if (J + 1 < I) {
#pragma unroll
            for (int k = 0; k < 8; ++k) b[k] = __ldcs(reinterpret_cast<const double2*>(SYMV_ROW(k) + (J + 1) * SYT));
        }
        double c0 = 0.0, c1 = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            racc[k] = fma(a[k].x, xj.x, racc[k]);
            racc[k] = fma(a[k].y, xj.y, racc[k]);
            c0 = fma(a[k].x, xi[k], c0);
            c1 = fma(a[k].y, xi[k], c1);
        }
        const int buf = (int)(J & 1);
        colred[buf][warp][lane] = make_double2(c0, c1);
        __syncthreads();                                   // one barrier per tile: colred is double-buffered
        if (threadIdx.x < SYT) {
            const double* cr = reinterpret_cast<const double*>(&colred[buf][0][0]) + threadIdx.x;
            double sum = 0.0;
#pragma unroll
            for (int w = 0; w < 8; ++w) sum += cr[w * SYT];
            P[I * ldp + J * SYT + threadIdx.x] = sum;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = b[k];
    }
    // diagonal tile: row part only (the tile holds both triangles), with column guards for a ragged last tile
    {
        const int64_t c = I * SYT + 2 * lane;
        const bool v0 = c < n, v1 = c + 1 < n;
        const double x0 = v0 ? x[c] : 0.0, x1 = v1 ? x[c + 1] : 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const double2 d = __ldcs(reinterpret_cast<const double2*>(SYMV_ROW(k) + I * SYT));   // inside the padded row
            racc[k] = fma(v0 ? d.x : 0.0, x0, racc[k]);
            racc[k] = fma(v1 ? d.y : 0.0, x1, racc[k]);
        }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const double v = warp_sum(racc[k]);
        if (lane == 0 && r0 + k < n) rowpart[r0 + k] = v;
    }
#undef SYMV_ROW
}

__global__ void __launch_bounds__(256)
symv_reduce_kernel(const double* __restrict__ rowpart, const double* __restrict__ P, int64_t ldp, int64_t n, int64_t nb,
                   double* __restrict__ y) {
    const int64_t j = blockIdx.x * 256ll + threadIdx.x;
    if (j >= n) return;
    double s0 = rowpart[j], s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int64_t I = j / SYT + 1;
    for (; I + 15 < nb; I += 16) {                 // 16 independent loads in flight per thread: rows of P are 8 ldp apart
        double v[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = P[(I + k) * ldp + j];
#pragma unroll
        for (int k = 0; k < 16; k += 4) {
            s0 += v[k];
            s1 += v[k + 1];
            s2 += v[k + 2];
            s3 += v[k + 3];
        }
    }
    for (; I < nb; ++I) s0 += P[I * ldp + j];
    y[j] = (s0 + s1) + (s2 + s3);
}

}  // namespace

int gemv(cudaStream_t stream, const double* A, int64_t rows, int64_t cols, int64_t lda, const double* x, double* y) {
    if (rows == 0) return PB_OK;
    gemv_kernel<<<(unsigned)ceil_div<int64_t>(rows, 4), 256, 0, stream>>>(A, rows, cols, lda, x, y); pb::note_launch();
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

int gemv_t_splits(int64_t rows) { return (int)ceil_div<int64_t>(rows, GT_ROWS); }

// partial[k * ldp + j], k < gemv_t_splits(rows): the caller adds the chunks.  A 16-byte aligned, lda even.
int gemv_t_partial(cudaStream_t stream, const double* A, int64_t rows, int64_t cols, int64_t lda, const double* x,
                   double* partial, int64_t ldp) {
    if (rows == 0 || cols == 0) return PB_OK;
    PB_CHECK((lda & 1) == 0 && lda >= cols && (reinterpret_cast<uintptr_t>(A) & 15) == 0, PB_ERR_INVALID,
             "gemv_t: A must be 16-byte aligned with an even leading dimension");
    dim3 grid((unsigned)ceil_div<int64_t>(ceil_div<int64_t>(cols, 2), 256), (unsigned)gemv_t_splits(rows));
    gemv_t_kernel<<<grid, 256, 0, stream>>>(A, rows, cols, lda, x, partial, ldp); pb::note_launch();
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

int64_t symv_lower_scratch_doubles(int64_t n) {
    const int64_t nb = ceil_div<int64_t>(n, SYT);
    return nb * (nb * SYT) + nb * SYT;
}

// y = K x for a symmetric K stored in full, reading only the tiles on and below the diagonal.
// K 16-byte aligned, ld even and >= n rounded up to 64 (padded rows are read, never used); x, y 16-byte aligned.
int symv_lower(cudaStream_t stream, const double* K, int64_t n, int64_t ld, const double* x, double* y, double* scratch) {
    if (n == 0) return PB_OK;
    const int64_t nb = ceil_div<int64_t>(n, SYT), ldp = nb * SYT;
    PB_CHECK((ld & 1) == 0 && ld >= ldp && (reinterpret_cast<uintptr_t>(K) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(x) & 15) == 0,
             PB_ERR_INVALID, "symv_lower: K and x must be 16-byte aligned and ld >= n rounded up to 64");
    double* P = scratch;
    double* rowpart = scratch + nb * ldp;
    symv_lower_kernel<<<(unsigned)nb, 256, 0, stream>>>(K, n, ld, x, rowpart, P, ldp); pb::note_launch();
    symv_reduce_kernel<<<(unsigned)ceil_div<int64_t>(n, 256), 256, 0, stream>>>(rowpart, P, ldp, n, nb, y); pb::note_launch();
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
    const double* tinv = dinv + ceil_div<int64_t>(n, LEAF) * LEAF * LEAF;     // TB-block inverses follow the leaf inverses
    if (trans) tinv += nblk * TB * TB;                                        // ... and their transposes follow those
    if (!trans) {
        trsv_fwd_kernel<<<1, TSV_THREADS, 0, stream>>>(L, n, ldl, tinv, -1, rhs, x); pb::note_launch();
        for (int64_t jb = 0; jb + 1 < nblk; ++jb) {
            const int64_t j0 = jb * TB;
            const unsigned grid = (unsigned)ceil_div<int64_t>(n - j0 - TB, TB);      // 256 rows per CTA below block j
            trsv_fwd_kernel<<<grid, TSV_THREADS, 0, stream>>>(L, n, ldl, tinv, j0, rhs, x); pb::note_launch();
        }
    } else {
        const int64_t jlast = (nblk - 1) * TB;
        trsv_bwd_kernel<<<1, TSV_THREADS, 0, stream>>>(L, n, ldl, tinv, jlast, 1, rhs, x); pb::note_launch();
        for (int64_t jb = nblk - 1; jb >= 1; --jb) {
            const int64_t j0 = jb * TB;
            trsv_bwd_kernel<<<(unsigned)(j0 / TB), TSV_THREADS, 0, stream>>>(L, n, ldl, tinv, j0, 0, rhs, x); pb::note_launch();
        }
    }
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

// Builds the TB-block inverses behind the leaf inverses in the potrf workspace (called at the end of potrf).
int build_block_inverses(cudaStream_t stream, const double* L, int64_t n, int64_t ldl, double* dinv) {
    if (n == 0) return PB_OK;
    static PerDeviceOnce configured;
    if (configured.first()) {
        PB_CUDA(cudaFuncSetAttribute(tb_inverse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TBINV_SMEM));
    }
    double* tinv = dinv + ceil_div<int64_t>(n, LEAF) * LEAF * LEAF;
    const int64_t nblk = ceil_div<int64_t>(n, TB);
    tb_inverse_kernel<<<(unsigned)nblk, 256, TBINV_SMEM, stream>>>(L, n, ldl, dinv, tinv, tinv + nblk * TB * TB); pb::note_launch();
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

extern "C" int pb_gemv(pb_stream_t stream, const double* A, int64_t rows, int64_t cols, int64_t lda, const double* x,
                       double* y) {
    PB_CHECK(A && x && y && lda >= cols, PB_ERR_INVALID, "gemv: bad arguments");
    return pb::gemv(reinterpret_cast<cudaStream_t>(stream), A, rows, cols, lda, x, y);
}

extern "C" int64_t pb_symv_lower_scratch_bytes(int64_t n) { return pb::symv_lower_scratch_doubles(n) * 8; }

extern "C" int pb_symv_lower(pb_stream_t stream, const double* K, int64_t n, int64_t ldk, const double* x, double* y,
                             void* scratch, int64_t scratch_bytes) {
    PB_CHECK(scratch != nullptr && scratch_bytes >= pb_symv_lower_scratch_bytes(n), PB_ERR_INVALID,
             "symv_lower: scratch too small");
    return pb::symv_lower(reinterpret_cast<cudaStream_t>(stream), K, n, ldk, x, y, reinterpret_cast<double*>(scratch));
}

extern "C" int pb_trsv(pb_stream_t stream, const double* L, int64_t n, int64_t ldl, const void* potrf_workspace,
                       int32_t trans, double* rhs, double* x) {
    return pb::trsv(reinterpret_cast<cudaStream_t>(stream), L, n, ldl,
                    reinterpret_cast<const double*>(potrf_workspace), trans != 0, rhs, x);
}

extern "C" int pb_logdet_chol(pb_stream_t stream, const double* L, int64_t n, int64_t ldl, double* out) {
    return pb::logdet_chol(reinterpret_cast<cudaStream_t>(stream), L, n, ldl, out);
}
