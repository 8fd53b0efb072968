// Fused Gram-matrix assembly for the mlkernels priors the reference examples use
// (EQ, Matern12/Exp, stretch, periodic, scalar scale).
//
// Replaces `prior(theta)(X)` at probit/implicit/Laplace.py:7,21,24, VB.py:7,22,
// probit/approximators.py:174,272,337 and `kernel(X_train, X_test)` at approximators.py:173.
//
//  1. pb_features maps inputs once: Z = T(X / stretch_in) / stretch_out (T = periodic sin/cos
//     feature map or identity), stored feature-major (Df x n), so the N^2 loop contains no trigonometry.
//  2. gram kernels evaluate r^2 = sum_d (z_id - z_jd)^2 with exact differences (no GEMM expansion:
//     SURVEY.md §7.2(b) — the expansion form puts O(1e-8) noise on diag(K) for Matern12) in 64x64
//     tiles, 4x4 outputs per thread, features staged in shared memory, 256-byte coalesced row
//     stores.  The symmetric variant computes lower tiles only and mirror-stores the transposed
//     tile through shared memory, so it is bound by the 8*N^2-byte HBM write, not by FP64 exp.
#include "common.cuh"

namespace pb {

namespace {

constexpr int TILE = 64;
constexpr int MAX_DF = 64;
constexpr double TWO_PI = 6.283185307179586;

__global__ void __launch_bounds__(256)
features_kernel(const double* __restrict__ X, int64_t n, int D, int64_t ldx, double* __restrict__ Z, int64_t ldz,
                int periodic, double stretch_in, double period, double stretch_out) {
    const int64_t total = n * D;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = e / D;
        const int d = (int)(e % D);
        double u = X[i * ldx + d];
        if (stretch_in != 1.0) u = u / stretch_in;
        if (periodic) {
            const double a = TWO_PI * u / period;
            double s, c;
            sincos(a, &s, &c);
            Z[(int64_t)d * ldz + i] = s / stretch_out;
            Z[(int64_t)(D + d) * ldz + i] = c / stretch_out;
        } else {
            Z[(int64_t)d * ldz + i] = u / stretch_out;
        }
    }
}

// sqrt(r2) for r2 >= 0 without the library's special-case branch (5 of the Matern tile's instructions per element were
// BSSY / BRA / BSYNC around rsqrt's slow path): x = MUFU.RSQ64H seed (2^-22), t = r2 x, e = 1 - r2 x^2,
// sqrt = t (1 + e/2 + 3 e^2 / 8) with remainder 2.5 e^3 ~ 2^-65; ~1.5 ulp.  Exactly 0 below 1e-300 (the seed flushes
// subnormals), so k(x, x) = scale exactly.
__device__ __forceinline__ double sqrt_pos(double r2) {
    double x;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(r2));
    const double t = r2 * x;
    const double e = fma(-t, x, 1.0);
    const double s = fma(t * e, fma(e, 0.375, 0.5), t);
    return r2 > 1e-300 ? s : 0.0;
}

template <int BASE>
__device__ __forceinline__ double base_eval_t(double scale, double r2) {
    if (BASE == PB_BASE_EQ) return scale * exp_neg(0.5 * r2);
    return scale * exp_neg(sqrt_pos(r2));
}

__device__ __forceinline__ double base_eval(int base, double scale, double r2) {
    if (base == PB_BASE_EQ) return scale * exp_neg(0.5 * r2);
    return scale * exp_neg(sqrt_pos(r2));
}

// Loads the features of 64 points starting at row0 into s[d*64 + r] (zero beyond n).  Features are stored
// feature-major, Z[d * ldz + i], so this is a straight coalesced copy with shift/mask index math only.
__device__ __forceinline__ void stage_features(double* s, const double* __restrict__ Z, int64_t ldz, int64_t row0,
                                               int64_t n, int Df) {
    for (int e = threadIdx.x; e < TILE * Df; e += 256) {
        const int d = e >> 6, r = e & (TILE - 1);
        const int64_t i = row0 + r;
        s[e] = i < n ? Z[(int64_t)d * ldz + i] : 0.0;
    }
}

// thread (ty, tx): rows ty + 16*r, cols 2*tx + 32*c + e
__device__ __forceinline__ void tile_distances(const double* si, const double* sj, int Df, int ty, int tx,
                                               double (&acc)[4][4]) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = 0.0;
    for (int d = 0; d < Df; ++d) {
        double zi[4], zj[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) zi[r] = si[d * TILE + ty + 16 * r];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const double2 v = *reinterpret_cast<const double2*>(&sj[d * TILE + 2 * tx + 32 * c]);
            zj[2 * c] = v.x;
            zj[2 * c + 1] = v.y;
        }
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const double diff = zi[r] - zj[c];
                acc[r][c] = fma(diff, diff, acc[r][c]);
            }
    }
}

// Lower-triangular tile enumeration: block b -> (ti, tj), tj <= ti.
__device__ __forceinline__ void tri_tile(int64_t b, int& ti, int& tj) {
    int64_t r = (int64_t)((sqrt(8.0 * (double)b + 1.0) - 1.0) * 0.5);
    while ((r + 1) * (r + 2) / 2 <= b) ++r;
    while (r * (r + 1) / 2 > b) --r;
    ti = (int)r;
    tj = (int)(b - r * (r + 1) / 2);
}

// BASE: kernel family at compile time (no per-element branch); FULL: n is a multiple of 64 and ldk is even,
// so every tile is interior and no bounds check is compiled in.  One tile per CTA: a persistent variant with
// register-prefetched features was measured 15-20 % SLOWER (two barriers per tile serialise the three
// resident CTAs; the hardware overlaps independent one-tile CTAs better).  A variant that walks the tile one group of
// 16 rows at a time (4 accumulators, 40 registers, 15 % fewer instructions) re-reads the column features from shared
// memory for every group and ran into the L1TEX limit instead (ncu l1tex throughput 70 - 86 %): no faster for Matern12
// D = 4, 15 % slower for EQ D = 8 (profiles/r02_hbm_kernels_ncu.md).
template <int BASE, bool FULL>
__global__ void __launch_bounds__(256)
gram_sym_kernel(const double* __restrict__ Z, int64_t n, int Df, int64_t ldz, double* __restrict__ K, int64_t ldk,
                double scale, const double* __restrict__ diag_vec, double diag_scalar) {
    extern __shared__ __align__(16) double sm[];
    double* si = sm;                   // [Df][64]
    double* sj = sm + Df * TILE;       // [Df][64]
    double* T = sj + Df * TILE;        // [64][65] transposed staging
    int ti, tj;
    tri_tile(blockIdx.x, ti, tj);
    const int64_t i0 = (int64_t)ti * TILE, j0 = (int64_t)tj * TILE;
    stage_features(si, Z, ldz, i0, n, Df);
    stage_features(sj, Z, ldz, j0, n, Df);
    __syncthreads();
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    double acc[4][4];
    tile_distances(si, sj, Df, ty, tx, acc);
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = base_eval_t<BASE>(scale, acc[r][c]);

    const bool diag_tile = ti == tj;
    if (diag_tile) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int row = ty + 16 * r, col = 2 * tx + 32 * (c >> 1) + (c & 1);
                if (row == col && (FULL || i0 + row < n))
                    acc[r][c] += diag_scalar + (diag_vec ? diag_vec[i0 + row] : 0.0);
            }
    }
    const bool full_cols = FULL || ((j0 + TILE <= n) && ((ldk & 1) == 0));
    double* kbase = K + (i0 + ty) * ldk + j0 + 2 * tx;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        if (!FULL && i0 + ty + 16 * r >= n) continue;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            double* p = kbase + (int64_t)(16 * r) * ldk + 32 * c;
            if (full_cols) {
                *reinterpret_cast<double2*>(p) = make_double2(acc[r][2 * c], acc[r][2 * c + 1]);
            } else {
                const int64_t col = j0 + 2 * tx + 32 * c;
                if (col < n) p[0] = acc[r][2 * c];
                if (col + 1 < n) p[1] = acc[r][2 * c + 1];
            }
        }
    }
    if (diag_tile) return;
    // mirror: K[j0 + col][i0 + row] = acc(row, col), transposed through smem for coalesced rows
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) T[(2 * tx + 32 * (c >> 1) + (c & 1)) * (TILE + 1) + ty + 16 * r] = acc[r][c];
    __syncthreads();
    const bool full_rows = FULL || ((i0 + TILE <= n) && ((ldk & 1) == 0));
    double* mbase = K + (j0 + ty) * ldk + i0 + 2 * tx;
#pragma unroll
    for (int r = 0; r < 4; ++r) {            // rows j0 + ty + 16 r are always < n because tj < ti
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int lc = 2 * tx + 32 * c;
            const double v0 = T[(ty + 16 * r) * (TILE + 1) + lc], v1 = T[(ty + 16 * r) * (TILE + 1) + lc + 1];
            double* p = mbase + (int64_t)(16 * r) * ldk + 32 * c;
            if (full_rows) {
                *reinterpret_cast<double2*>(p) = make_double2(v0, v1);
            } else {
                const int64_t col = i0 + lc;
                if (col < n) p[0] = v0;
                if (col + 1 < n) p[1] = v1;
            }
        }
    }
}

// One rectangular block of a I + s s^T o (K + jitter I) generated straight from the features (no resident K):
//   out[r][c] = s_i k(z_i, z_j) s_j  (+ a + s_i^2 jitter where i == j),  i = row0 + r, j = col0 + c   (s null -> ones)
// with exactly the arithmetic of sym_transform_kernel, so a block-cyclic layout holds bit-identical entries.
// Used by the multi-GPU Cholesky: every rank fills only the block columns it owns.
__global__ void __launch_bounds__(256)
gram_block_kernel(const double* __restrict__ Z, int64_t ldz, int Df, int64_t row0, int64_t rows, int64_t col0,
                  int64_t cols, int base, double scale, const double* __restrict__ s, double a, double jitter,
                  double* __restrict__ out, int64_t ldo) {
    extern __shared__ __align__(16) double sm[];
    double* si = sm;
    double* sj = sm + Df * TILE;
    const int64_t i0 = (int64_t)blockIdx.y * TILE, j0 = (int64_t)blockIdx.x * TILE;
    stage_features(si, Z + row0, ldz, i0, rows, Df);
    stage_features(sj, Z + col0, ldz, j0, cols, Df);
    __syncthreads();
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    double acc[4][4];
    tile_distances(si, sj, Df, ty, tx, acc);
    double cs[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const int64_t col = j0 + 2 * tx + 32 * (c >> 1) + (c & 1);
        cs[c] = (s && col < cols) ? s[col0 + col] : 1.0;
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int64_t row = i0 + ty + 16 * r;
        if (row >= rows) continue;
        const double sr = s ? s[row0 + row] : 1.0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int64_t col = j0 + 2 * tx + 32 * (c >> 1) + (c & 1);
            if (col >= cols) continue;
            double v = base_eval(base, scale, acc[r][c]);
            v = (row0 + row == col0 + col) ? a + sr * (v + jitter) * cs[c] : sr * v * cs[c];
            out[row * ldo + col] = v;
        }
    }
}

// K[n1 x n2] = k(Z1, Z2) * (col_scale ? col_scale[j] : 1)
__global__ void __launch_bounds__(256)
gram_cross_kernel(const double* __restrict__ Z1, int64_t n1, const double* __restrict__ Z2, int64_t n2, int Df,
                  int64_t ldz1, int64_t ldz2, double* __restrict__ K, int64_t ldk, int base, double scale,
                  const double* __restrict__ col_scale) {
    extern __shared__ __align__(16) double sm[];
    double* si = sm;
    double* sj = sm + Df * TILE;
    const int64_t i0 = (int64_t)blockIdx.y * TILE, j0 = (int64_t)blockIdx.x * TILE;
    stage_features(si, Z1, ldz1, i0, n1, Df);
    stage_features(sj, Z2, ldz2, j0, n2, Df);
    __syncthreads();
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    double acc[4][4];
    tile_distances(si, sj, Df, ty, tx, acc);
    double cs[4] = {1.0, 1.0, 1.0, 1.0};
    if (col_scale) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int64_t col = j0 + 2 * tx + 32 * (c >> 1) + (c & 1);
            cs[c] = col < n2 ? col_scale[col] : 0.0;
        }
    }
    const bool full_cols = (j0 + TILE <= n2) && ((ldk & 1) == 0);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int64_t row = i0 + ty + 16 * r;
        if (row >= n1) continue;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int64_t col = j0 + 2 * tx + 32 * c;
            const double v0 = base_eval(base, scale, acc[r][2 * c]) * cs[2 * c];
            const double v1 = base_eval(base, scale, acc[r][2 * c + 1]) * cs[2 * c + 1];
            double* p = K + row * ldk + col;
            if (full_cols) {
                *reinterpret_cast<double2*>(p) = make_double2(v0, v1);
            } else {
                if (col < n2) p[0] = v0;
                if (col + 1 < n2) p[1] = v1;
            }
        }
    }
}

// B_ij = a*delta_ij + s_i (K_ij + jitter*delta_ij) s_j for j <= i (s == nullptr -> s = 1).
__global__ void __launch_bounds__(256)
sym_transform_kernel(const double* __restrict__ K, int64_t n, int64_t ldk, const double* __restrict__ s, double a,
                     double jitter, double* __restrict__ B, int64_t ldb) {
    int ti, tj;
    tri_tile(blockIdx.x, ti, tj);
    const int64_t i0 = (int64_t)ti * TILE, j0 = (int64_t)tj * TILE;
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    double sj[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const int64_t col = j0 + 2 * tx + 32 * (c >> 1) + (c & 1);
        sj[c] = (s && col < n) ? s[col] : 1.0;
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int64_t row = i0 + ty + 16 * r;
        if (row >= n) continue;
        const double si = s ? s[row] : 1.0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int64_t col = j0 + 2 * tx + 32 * (c >> 1) + (c & 1);
            if (col > row) continue;
            double v = K[row * ldk + col];
            if (col == row) v = a + si * (v + jitter) * sj[c];
            else v = si * v * sj[c];
            B[row * ldb + col] = v;
        }
    }
}

// out[r][c] = s_i K_ij s_j (+ a + s_i^2 jitter on the global diagonal) for the rectangle
// i = row0 + r, j = col0 + c — the block-cyclic layouts of the multi-GPU Cholesky take their block columns
// of B = a I + s s^T o (K + jitter I) straight from the resident K with this.
__global__ void __launch_bounds__(256)
transform_block_kernel(const double* __restrict__ K, int64_t ldk, const double* __restrict__ s, double a,
                       double jitter, int64_t row0, int64_t col0, int64_t rows, int64_t cols,
                       double* __restrict__ out, int64_t ldo) {
    const int64_t c = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (c >= cols) return;
    const double sj = s ? s[col0 + c] : 1.0;
    for (int64_t r = (int64_t)blockIdx.y * 64; r < rows && r < (int64_t)(blockIdx.y + 1) * 64; ++r) {
        const int64_t i = row0 + r, j = col0 + c;
        const double si = s ? s[i] : 1.0;
        double v = K[i * ldk + j];
        v = (i == j) ? a + si * (v + jitter) * sj : si * v * sj;
        out[r * ldo + c] = v;
    }
}

// Fused predictive mean (probit/approximators.py:173,179: Kfs.T @ weight): partial[split][i] =
// sum_{j in split} k(z1_i, z2_j) v_j with the cross-covariance tile generated in registers and never
// written anywhere.  grid = (row tiles of 64 test points, column splits); a second tiny kernel adds the
// splits in a fixed order (deterministic).
template <int BASE>
__global__ void __launch_bounds__(256)
gram_matvec_kernel(const double* __restrict__ Z1, int64_t n1, const double* __restrict__ Z2, int64_t n2, int Df,
                   int64_t ldz1, int64_t ldz2, double scale, const double* __restrict__ v, int64_t cols_per_split,
                   double* __restrict__ partial) {
    extern __shared__ __align__(16) double sm[];
    double* si = sm;
    double* sj = sm + Df * TILE;
    double* sv = sj + Df * TILE;
    const int64_t i0 = (int64_t)blockIdx.x * TILE;
    const int64_t jbeg = (int64_t)blockIdx.y * cols_per_split;
    const int64_t jend = jbeg + cols_per_split < n2 ? jbeg + cols_per_split : n2;
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    stage_features(si, Z1, ldz1, i0, n1, Df);
    double rowacc[4] = {0, 0, 0, 0};
    for (int64_t j0 = jbeg; j0 < jend; j0 += TILE) {
        __syncthreads();
        stage_features(sj, Z2, ldz2, j0, jend, Df);
        if (threadIdx.x < TILE) sv[threadIdx.x] = (j0 + threadIdx.x < jend) ? v[j0 + threadIdx.x] : 0.0;
        __syncthreads();
        double acc[4][4];
        tile_distances(si, sj, Df, ty, tx, acc);
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int lc = 2 * tx + 32 * (c >> 1) + (c & 1);
                rowacc[r] = fma(base_eval_t<BASE>(scale, acc[r][c]), sv[lc], rowacc[r]);   // sv = 0 beyond the split
            }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        double a = rowacc[r];
        a += __shfl_xor_sync(0xffffffffu, a, 1);
        a += __shfl_xor_sync(0xffffffffu, a, 2);
        a += __shfl_xor_sync(0xffffffffu, a, 4);
        a += __shfl_xor_sync(0xffffffffu, a, 8);
        const int64_t row = i0 + ty + 16 * r;
        if (tx == 0 && row < n1) partial[(int64_t)blockIdx.y * n1 + row] = a;
    }
}

__global__ void __launch_bounds__(256)
sum_splits_kernel(const double* __restrict__ partial, int splits, int64_t n, double* __restrict__ y) {
    for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        double a = 0.0;
        for (int s = 0; s < splits; ++s) a += partial[(int64_t)s * n + i];
        y[i] = a;
    }
}

// ---- kernel-derivative reductions for the evidence gradient (probit/implicit/solvers.py:52-64 replaced by
// the closed form of Rasmussen & Williams Alg. 5.1; oracle/gradients.py) ---------------------------------
// For a stationary kernel K = c * base(rho), rho^2 = ||z_i - z_j||^2 with z = features / l:
//   dK/dc = K / c,   dK/dl = K * rho^2 / l (EQ)  |  K * rho / l (Exp).
// The kernels below stream K (and B^{-1}) once and regenerate rho from the staged features.
__device__ __forceinline__ double rho_term(int base, double r2) { return base == PB_BASE_EQ ? r2 : sqrt(r2); }

// y[i] = sum_j K_ij * rho_term_ij * v_j  (the l-derivative matvec, up to the 1/l factor); one 64-row tile per CTA
__global__ void __launch_bounds__(256)
gram_deriv_matvec_kernel(const double* __restrict__ Z, int64_t n, int Df, int64_t ldz, const double* __restrict__ K,
                         int64_t ldk, int base, const double* __restrict__ v, double* __restrict__ y) {
    extern __shared__ __align__(16) double sm[];
    double* si = sm;
    double* sj = sm + Df * TILE;
    double* sv = sj + Df * TILE;           // v for the current column tile [64]
    const int64_t i0 = (int64_t)blockIdx.x * TILE;
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    stage_features(si, Z, ldz, i0, n, Df);
    double rowacc[4] = {0, 0, 0, 0};
    for (int64_t j0 = 0; j0 < n; j0 += TILE) {
        __syncthreads();
        stage_features(sj, Z, ldz, j0, n, Df);
        if (threadIdx.x < TILE) sv[threadIdx.x] = (j0 + threadIdx.x < n) ? v[j0 + threadIdx.x] : 0.0;
        __syncthreads();
        double acc[4][4];
        tile_distances(si, sj, Df, ty, tx, acc);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int64_t row = i0 + ty + 16 * r;
            if (row >= n) continue;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int lc = 2 * tx + 32 * (c >> 1) + (c & 1);
                if (j0 + lc < n) rowacc[r] = fma(K[row * ldk + j0 + lc] * rho_term(base, acc[r][c]), sv[lc], rowacc[r]);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        double a = rowacc[r];
        a += __shfl_xor_sync(0xffffffffu, a, 1);
        a += __shfl_xor_sync(0xffffffffu, a, 2);
        a += __shfl_xor_sync(0xffffffffu, a, 4);
        a += __shfl_xor_sync(0xffffffffu, a, 8);
        const int64_t row = i0 + ty + 16 * r;
        if (tx == 0 && row < n) y[row] = a;
    }
}

// Over lower tiles (each off-diagonal pair counted twice):
//   partial[3b+0] = sum a_i a_j K_ij rho_ij           (-> w^T dK/dl w * l)
//   partial[3b+1] = sum s_i s_j Binv_ij K_ij rho_ij   (-> tr(R dK/dl) * l)
//   partial[3b+2] = sum s_i s_j Binv_ij K_ij          (-> tr(R dK/dc) * c)
__global__ void __launch_bounds__(256)
gram_deriv_dots_kernel(const double* __restrict__ Z, int64_t n, int Df, int64_t ldz, const double* __restrict__ K,
                       int64_t ldk, const double* __restrict__ Binv, int64_t ldb, int base,
                       const double* __restrict__ a, const double* __restrict__ s, double* __restrict__ partial) {
    extern __shared__ __align__(16) double sm[];
    double* si = sm;
    double* sj = sm + Df * TILE;
    int ti, tj;
    tri_tile(blockIdx.x, ti, tj);
    const int64_t i0 = (int64_t)ti * TILE, j0 = (int64_t)tj * TILE;
    stage_features(si, Z, ldz, i0, n, Df);
    stage_features(sj, Z, ldz, j0, n, Df);
    __syncthreads();
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    double acc[4][4];
    tile_distances(si, sj, Df, ty, tx, acc);
    double p0 = 0, p1 = 0, p2 = 0;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int64_t row = i0 + ty + 16 * r;
        if (row >= n) continue;
        const double ai = a[row], sr = s[row];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int64_t col = j0 + 2 * tx + 32 * (c >> 1) + (c & 1);
            if (col > row || col >= n) continue;
            const double wgt = (col == row) ? 1.0 : 2.0;
            const double k = K[row * ldk + col];
            const double kr = k * rho_term(base, acc[r][c]);
            const double sb = wgt * sr * s[col] * Binv[row * ldb + col];
            p0 = fma(wgt * ai * a[col], kr, p0);
            p1 = fma(sb, kr, p1);
            p2 = fma(sb, k, p2);
        }
    }
    p0 = block_sum<256>(p0);
    p1 = block_sum<256>(p1);
    p2 = block_sum<256>(p2);
    if (threadIdx.x == 0) {
        partial[3 * (int64_t)blockIdx.x] = p0;
        partial[3 * (int64_t)blockIdx.x + 1] = p1;
        partial[3 * (int64_t)blockIdx.x + 2] = p2;
    }
}

__global__ void __launch_bounds__(1024)
sum3_kernel(const double* __restrict__ partial, int64_t nblk, double* __restrict__ out) {
    double a = 0, b = 0, c = 0;
    for (int64_t i = threadIdx.x; i < nblk; i += 1024) {
        a += partial[3 * i];
        b += partial[3 * i + 1];
        c += partial[3 * i + 2];
    }
    a = block_sum<1024>(a);
    b = block_sum<1024>(b);
    c = block_sum<1024>(c);
    if (threadIdx.x == 0) { out[0] = a; out[1] = b; out[2] = c; }
}

__global__ void __launch_bounds__(256)
identity_kernel(double* __restrict__ A, int64_t n, int64_t ld) {
    for (int64_t row = blockIdx.y; row < n; row += gridDim.y)
        for (int64_t c = blockIdx.x * 256ll + threadIdx.x; c < n; c += (int64_t)gridDim.x * 256)
            A[row * ld + c] = (c == row) ? 1.0 : 0.0;
}

// ---- TEST-ONLY distance form (pb_kernel_spec.distance_form = 1) ------------------------------------------------
// lab's B.pw_dists2 for more than one feature: ||a||^2 + ||b||^2 - 2 a.b, and B.pw_dists = sqrt(max(., 1e-30))
// (SURVEY.md §9.1).  On the diagonal the three terms cancel to a rounding residue instead of 0, which a Matern12
// kernel turns into diag(K) = 1 - O(1e-8): the reference's own noise floor at the 1e-8 tolerance.  The product never
// uses this form; the parity tests switch it on to show that the residual difference to the reference-source
// fixtures is this effect and nothing else.  One thread per element, no tiling: correctness only.
__device__ __forceinline__ double expand_eval(const double* __restrict__ Z1, int64_t ldz1, int64_t i,
                                              const double* __restrict__ Z2, int64_t ldz2, int64_t j, int Df,
                                              int base, double scale) {
    double na = 0.0, nb = 0.0, dot = 0.0;
    for (int d = 0; d < Df; ++d) {
        const double a = Z1[(int64_t)d * ldz1 + i], b = Z2[(int64_t)d * ldz2 + j];
        na = fma(a, a, na);
        nb = fma(b, b, nb);
        dot = fma(a, b, dot);
    }
    const double r2 = (na + nb) - 2.0 * dot;
    if (base == PB_BASE_EQ) return scale * exp(-0.5 * r2);
    return scale * exp(-sqrt(fmax(r2, 1e-30)));
}

__global__ void __launch_bounds__(256)
gram_expand_kernel(const double* __restrict__ Z1, int64_t n1, const double* __restrict__ Z2, int64_t n2, int Df,
                   int64_t ldz1, int64_t ldz2, double* __restrict__ K, int64_t ldk, int base, double scale,
                   const double* __restrict__ col_scale, const double* __restrict__ diag_vec, double diag_scalar,
                   int add_diag) {
    const int64_t j = blockIdx.x * 256ll + threadIdx.x, i = blockIdx.y;
    if (j >= n2) return;
    double v = expand_eval(Z1, ldz1, i, Z2, ldz2, j, Df, base, scale);
    if (col_scale) v *= col_scale[j];
    if (add_diag && i == j) v += diag_scalar + (diag_vec ? diag_vec[i] : 0.0);
    K[i * ldk + j] = v;
}

// y[i] = sum_j k(z1_i, z2_j) v_j, one warp per row
__global__ void __launch_bounds__(256)
gram_expand_matvec_kernel(const double* __restrict__ Z1, int64_t n1, const double* __restrict__ Z2, int64_t n2, int Df,
                          int64_t ldz1, int64_t ldz2, int base, double scale, const double* __restrict__ v,
                          double* __restrict__ y) {
    const int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= n1) return;
    double acc = 0.0;
    for (int64_t j = threadIdx.x & 31; j < n2; j += 32)
        acc = fma(expand_eval(Z1, ldz1, i, Z2, ldz2, j, Df, base, scale), v[j], acc);
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) y[i] = acc;
}

inline bool use_expand(const pb_kernel_spec& spec, int Df) { return spec.distance_form == 1 && Df > 1; }

int gram_expand(cudaStream_t stream, const pb_kernel_spec& spec, const double* Z1, int64_t n1, const double* Z2,
                int64_t n2, int Df, int64_t ldz1, int64_t ldz2, double* K, int64_t ldk, const double* col_scale,
                const double* diag_vec, double diag_scalar, bool add_diag) {
    dim3 grid((unsigned)ceil_div<int64_t>(n2, 256), (unsigned)n1);
    PB_CHECK(n1 < 65536, PB_ERR_UNSUPPORTED, "distance_form = 1 is a test-only mode (n < 65536)");
    gram_expand_kernel<<<grid, 256, 0, stream>>>(Z1, n1, Z2, n2, Df, ldz1, ldz2, K, ldk, spec.base, spec.scale, col_scale,
                                                 diag_vec, diag_scalar, add_diag ? 1 : 0); pb::note_launch();
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

inline int64_t tri_tiles(int64_t n) {
    const int64_t T = ceil_div<int64_t>(n, TILE);
    return T * (T + 1) / 2;
}

constexpr int GRAM_SMEM_SYM_MAX = (2 * MAX_DF * TILE + TILE * (TILE + 1)) * 8;
constexpr int GRAM_SMEM_CROSS_MAX = (2 * MAX_DF * TILE) * 8;
inline int gram_smem_sym(int Df) { return (2 * Df * TILE + TILE * (TILE + 1)) * 8; }
inline int gram_smem_cross(int Df) { return (2 * Df * TILE) * 8; }

}  // namespace

int feature_dim(const pb_kernel_spec& spec, int D) { return spec.periodic ? 2 * D : D; }

int check_spec(const pb_kernel_spec& spec) {
    PB_CHECK(spec.base == PB_BASE_EQ || spec.base == PB_BASE_EXP, PB_ERR_UNSUPPORTED, "kernel base %d unsupported",
             spec.base);
    PB_CHECK(spec.stretch_in > 0 && spec.stretch_out > 0, PB_ERR_INVALID, "kernel stretch must be positive");
    PB_CHECK(!spec.periodic || spec.period > 0, PB_ERR_INVALID, "kernel period must be positive");
    PB_CHECK(spec.distance_form == 0 || spec.distance_form == 1, PB_ERR_INVALID, "kernel distance_form must be 0 or 1");
    return PB_OK;
}

int features(cudaStream_t stream, const pb_kernel_spec& spec, const double* X, int64_t n, int D, int64_t ldx,
             double* Z, int64_t ldz) {
    PB_TRY(check_spec(spec));
    PB_CHECK(D >= 1 && feature_dim(spec, D) <= MAX_DF, PB_ERR_UNSUPPORTED, "feature dimension %d exceeds %d",
             feature_dim(spec, D), MAX_DF);
    if (n == 0) return PB_OK;
    const int64_t want = ceil_div<int64_t>(n * D, 256);
    const int64_t cap = (int64_t)num_sms() * 8;
    features_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, stream>>>(X, n, D, ldx, Z, ldz, spec.periodic,
                                                                             spec.stretch_in, spec.period,
                                                                             spec.stretch_out); pb::note_launch();
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

int gram_sym(cudaStream_t stream, const pb_kernel_spec& spec, const double* Z, int64_t n, int Df, int64_t ldz,
             double* K, int64_t ldk, const double* diag_vec, double diag_scalar) {
    PB_TRY(check_spec(spec));
    PB_CHECK(Df >= 1 && Df <= MAX_DF, PB_ERR_UNSUPPORTED, "feature dimension %d exceeds %d", Df, MAX_DF);
    if (n == 0) return PB_OK;
    if (use_expand(spec, Df)) return gram_expand(stream, spec, Z, n, Z, n, Df, ldz, ldz, K, ldk, nullptr, diag_vec, diag_scalar, true);
    const bool full = (n % TILE == 0) && ((ldk & 1) == 0) && ((reinterpret_cast<uintptr_t>(K) & 15) == 0);
    const unsigned grid = (unsigned)tri_tiles(n);
    const int smem = gram_smem_sym(Df);
#define PB_LAUNCH_GRAM_SYM(BASE, FULL)                                                                              \
    do {                                                                                                           \
        static PerDeviceOnce configured;                                                                            \
        if (configured.first()) {                                                                                         \
            PB_CUDA(cudaFuncSetAttribute(gram_sym_kernel<BASE, FULL>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                         GRAM_SMEM_SYM_MAX));                                                      \
        }                                                                                                          \
        gram_sym_kernel<BASE, FULL><<<grid, 256, smem, stream>>>(Z, n, Df, ldz, K, ldk, spec.scale, diag_vec,       \
                                                                 diag_scalar);                                     \
    } while (0)
    if (spec.base == PB_BASE_EQ) {
        if (full) PB_LAUNCH_GRAM_SYM(PB_BASE_EQ, true); else PB_LAUNCH_GRAM_SYM(PB_BASE_EQ, false);
    } else {
        if (full) PB_LAUNCH_GRAM_SYM(PB_BASE_EXP, true); else PB_LAUNCH_GRAM_SYM(PB_BASE_EXP, false);
    }
#undef PB_LAUNCH_GRAM_SYM
    pb::note_launch();
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

int gram_cross(cudaStream_t stream, const pb_kernel_spec& spec, const double* Z1, int64_t n1, const double* Z2,
               int64_t n2, int Df, int64_t ldz1, int64_t ldz2, double* K, int64_t ldk, const double* col_scale) {
    PB_TRY(check_spec(spec));
    PB_CHECK(Df >= 1 && Df <= MAX_DF, PB_ERR_UNSUPPORTED, "feature dimension %d exceeds %d", Df, MAX_DF);
    if (n1 == 0 || n2 == 0) return PB_OK;
    if (use_expand(spec, Df)) return gram_expand(stream, spec, Z1, n1, Z2, n2, Df, ldz1, ldz2, K, ldk, col_scale, nullptr, 0.0, false);
    static PerDeviceOnce configured;
    if (configured.first()) {
        PB_CUDA(cudaFuncSetAttribute(gram_cross_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GRAM_SMEM_CROSS_MAX));
    }
    dim3 grid((unsigned)ceil_div<int64_t>(n2, TILE), (unsigned)ceil_div<int64_t>(n1, TILE));
    PB_CHECK(grid.y < 65536, PB_ERR_INVALID, "gram_cross: too many row tiles (chunk the rows)");
    gram_cross_kernel<<<grid, 256, gram_smem_cross(Df), stream>>>(Z1, n1, Z2, n2, Df, ldz1, ldz2, K, ldk, spec.base,
                                                             spec.scale, col_scale); pb::note_launch();
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

int gram_block(cudaStream_t stream, const pb_kernel_spec& spec, const double* Z, int64_t ldz, int Df, int64_t row0,
               int64_t rows, int64_t col0, int64_t cols, const double* s, double a, double jitter, double* out,
               int64_t ldo) {
    PB_TRY(check_spec(spec));
    PB_CHECK(Df >= 1 && Df <= MAX_DF, PB_ERR_UNSUPPORTED, "feature dimension %d exceeds %d", Df, MAX_DF);
    PB_CHECK(spec.distance_form == 0, PB_ERR_UNSUPPORTED, "the test-only distance_form = 1 has no block-cyclic variant");
    if (rows <= 0 || cols <= 0) return PB_OK;
    static PerDeviceOnce configured;
    if (configured.first()) {
        PB_CUDA(cudaFuncSetAttribute(gram_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GRAM_SMEM_CROSS_MAX));
    }
    dim3 grid((unsigned)ceil_div<int64_t>(cols, TILE), (unsigned)ceil_div<int64_t>(rows, TILE));
    PB_CHECK(grid.y < 65536, PB_ERR_INVALID, "gram_block: too many row tiles");
    gram_block_kernel<<<grid, 256, gram_smem_cross(Df), stream>>>(Z, ldz, Df, row0, rows, col0, cols, spec.base, spec.scale,
                                                             s, a, jitter, out, ldo); pb::note_launch();
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

int sym_transform(cudaStream_t stream, const double* K, int64_t n, int64_t ldk, const double* s, double a,
                  double jitter, double* B, int64_t ldb) {
    if (n == 0) return PB_OK;
    sym_transform_kernel<<<(unsigned)tri_tiles(n), 256, 0, stream>>>(K, n, ldk, s, a, jitter, B, ldb); pb::note_launch();
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

// y[i] = sum_j k(z1_i, z2_j) v_j without materialising the cross Gram; `partial` needs splits * n1 doubles
int gram_matvec_splits(int64_t n1) {
    const int64_t row_tiles = ceil_div<int64_t>(n1 > 0 ? n1 : 1, TILE);
    int64_t s = ceil_div<int64_t>(4 * (int64_t)num_sms(), row_tiles);
    return (int)(s < 1 ? 1 : (s > 64 ? 64 : s));
}

int gram_matvec(cudaStream_t stream, const pb_kernel_spec& spec, const double* Z1, int64_t n1, const double* Z2,
                int64_t n2, int Df, int64_t ldz1, int64_t ldz2, const double* v, double* partial, double* y) {
    PB_TRY(check_spec(spec));
    if (n1 == 0) return PB_OK;
    if (use_expand(spec, Df)) {
        gram_expand_matvec_kernel<<<(unsigned)ceil_div<int64_t>(n1, 8), 256, 0, stream>>>(Z1, n1, Z2, n2, Df, ldz1, ldz2,
                                                                                       spec.base, spec.scale, v, y); pb::note_launch();
        PB_CUDA(cudaGetLastError());
        return PB_OK;
    }
    const int splits = gram_matvec_splits(n1);
    const int64_t cols = ceil_div<int64_t>(ceil_div<int64_t>(n2, splits), TILE) * TILE;
    const int smem = (2 * Df * TILE + TILE) * 8;
    dim3 grid((unsigned)ceil_div<int64_t>(n1, TILE), (unsigned)splits);
    static PerDeviceOnce configured;
    if (configured.first()) {
        PB_CUDA(cudaFuncSetAttribute(gram_matvec_kernel<PB_BASE_EQ>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (2 * MAX_DF * TILE + TILE) * 8));
        PB_CUDA(cudaFuncSetAttribute(gram_matvec_kernel<PB_BASE_EXP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (2 * MAX_DF * TILE + TILE) * 8));
    }
    if (spec.base == PB_BASE_EQ)
        gram_matvec_kernel<PB_BASE_EQ><<<grid, 256, smem, stream>>>(Z1, n1, Z2, n2, Df, ldz1, ldz2, spec.scale, v, cols, partial);
    else
        gram_matvec_kernel<PB_BASE_EXP><<<grid, 256, smem, stream>>>(Z1, n1, Z2, n2, Df, ldz1, ldz2, spec.scale, v, cols, partial);
    pb::note_launch();
    const int64_t want = ceil_div<int64_t>(n1, 256);
    sum_splits_kernel<<<(unsigned)(want < 1024 ? want : 1024), 256, 0, stream>>>(partial, splits, n1, y); pb::note_launch();
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

int gram_deriv_matvec(cudaStream_t stream, const pb_kernel_spec& spec, const double* Z, int64_t n, int Df, int64_t ldz,
                      const double* K, int64_t ldk, const double* v, double* y) {
    if (n == 0) return PB_OK;
    PB_CHECK(spec.distance_form == 0, PB_ERR_UNSUPPORTED, "gradients are not available in the test-only distance_form = 1");
    const int smem = (2 * Df * TILE + TILE) * 8;
    static PerDeviceOnce configured;
    if (configured.first()) {
        PB_CUDA(cudaFuncSetAttribute(gram_deriv_matvec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (2 * MAX_DF * TILE + TILE) * 8));
        PB_CUDA(cudaFuncSetAttribute(gram_deriv_dots_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     GRAM_SMEM_CROSS_MAX));
    }
    gram_deriv_matvec_kernel<<<(unsigned)ceil_div<int64_t>(n, TILE), 256, smem, stream>>>(Z, n, Df, ldz, K, ldk, spec.base, v, y); pb::note_launch();
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

// out[0..2] (device) = the three lower-triangle sums of gram_deriv_dots_kernel; `partial` needs 3*tri_tiles(n) doubles
int64_t gram_deriv_partial_doubles(int64_t n) { return 3 * tri_tiles(n); }

int gram_deriv_dots(cudaStream_t stream, const pb_kernel_spec& spec, const double* Z, int64_t n, int Df, int64_t ldz,
                    const double* K, int64_t ldk, const double* Binv, int64_t ldb, const double* a, const double* s,
                    double* partial, double* out) {
    if (n == 0) return PB_OK;
    static PerDeviceOnce configured;
    if (configured.first()) {
        PB_CUDA(cudaFuncSetAttribute(gram_deriv_dots_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     GRAM_SMEM_CROSS_MAX));
    }
    const int64_t tiles = tri_tiles(n);
    gram_deriv_dots_kernel<<<(unsigned)tiles, 256, gram_smem_cross(Df), stream>>>(Z, n, Df, ldz, K, ldk, Binv, ldb, spec.base, a, s, partial); pb::note_launch();
    sum3_kernel<<<1, 1024, 0, stream>>>(partial, tiles, out); pb::note_launch();
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

int set_identity(cudaStream_t stream, double* A, int64_t n, int64_t ld) {
    if (n == 0) return PB_OK;
    dim3 grid((unsigned)(ceil_div<int64_t>(n, 256) < 64 ? ceil_div<int64_t>(n, 256) : 64), (unsigned)(n < 32768 ? n : 32768));
    identity_kernel<<<grid, 256, 0, stream>>>(A, n, ld); pb::note_launch();
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

int transform_block(cudaStream_t stream, const double* K, int64_t ldk, const double* s, double a, double jitter,
                    int64_t row0, int64_t col0, int64_t rows, int64_t cols, double* out, int64_t ldo) {
    if (rows <= 0 || cols <= 0) return PB_OK;
    dim3 grid((unsigned)ceil_div<int64_t>(cols, 256), (unsigned)ceil_div<int64_t>(rows, 64));
    PB_CHECK(grid.y < 65536, PB_ERR_INVALID, "transform_block: too many rows");
    transform_block_kernel<<<grid, 256, 0, stream>>>(K, ldk, s, a, jitter, row0, col0, rows, cols, out, ldo); pb::note_launch();
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

}  // namespace pb

extern "C" int pb_transform_block(pb_stream_t stream, const double* K, int64_t ldk, const double* s, double a,
                                  double jitter, int64_t row0, int64_t col0, int64_t rows, int64_t cols, double* out,
                                  int64_t ldo) {
    return pb::transform_block(reinterpret_cast<cudaStream_t>(stream), K, ldk, s, a, jitter, row0, col0, rows, cols,
                               out, ldo);
}

extern "C" int pb_feature_dim(const pb_kernel_spec* spec, int D) { return pb::feature_dim(*spec, D); }

extern "C" int pb_features(pb_stream_t stream, const pb_kernel_spec* spec, const double* X, int64_t n, int D,
                           int64_t ldx, double* Z, int64_t ldz) {
    return pb::features(reinterpret_cast<cudaStream_t>(stream), *spec, X, n, D, ldx, Z, ldz);
}

extern "C" int pb_gram_sym(pb_stream_t stream, const pb_kernel_spec* spec, const double* Z, int64_t n, int Df,
                           int64_t ldz, double* K, int64_t ldk, const double* diag_vec, double diag_scalar) {
    return pb::gram_sym(reinterpret_cast<cudaStream_t>(stream), *spec, Z, n, Df, ldz, K, ldk, diag_vec, diag_scalar);
}

extern "C" int pb_gram_cross(pb_stream_t stream, const pb_kernel_spec* spec, const double* Z1, int64_t n1,
                             const double* Z2, int64_t n2, int Df, int64_t ldz1, int64_t ldz2, double* K,
                             int64_t ldk) {
    return pb::gram_cross(reinterpret_cast<cudaStream_t>(stream), *spec, Z1, n1, Z2, n2, Df, ldz1, ldz2, K, ldk,
                          nullptr);
}

extern "C" int pb_scale_sym_plus_identity(pb_stream_t stream, const double* K, int64_t n, int64_t ldk,
                                          const double* s, double jitter, double* B, int64_t ldb) {
    return pb::sym_transform(reinterpret_cast<cudaStream_t>(stream), K, n, ldk, s, 1.0, jitter, B, ldb);
}

extern "C" int pb_copy_lower_add_diag(pb_stream_t stream, const double* K, int64_t n, int64_t ldk,
                                      double diag_scalar, double* A, int64_t lda) {
    return pb::sym_transform(reinterpret_cast<cudaStream_t>(stream), K, n, ldk, nullptr, diag_scalar, 0.0, A, lda);
}
