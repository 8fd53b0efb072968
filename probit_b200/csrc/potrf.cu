// Blocked right-looking FP64 Cholesky for sm_100a (row-major, lower).
//
// Replaces B.cholesky at probit/implicit/Laplace.py:24 and VB.py:10,25, and (through the SPD
// Newton form) the LU solve at probit/implicit/solvers.py:24.
//
//   for each block column k (width NB):
//     1. diagonal block: recursive potrf down to 64x64 leaves; a leaf is one CTA that factors the
//        block held in registers (4x4 per thread), exchanging one pivot column per step through
//        shared memory, and also emits the inverse of the leaf (the TRSMs use it as a GEMM operand);
//     2. panel TRSM  P <- P L_kk^{-T}: recursive, every flop is a DMMA GEMM (gemm_dmma.cu);
//     3. trailing SYRK  A22 -= P P^T on lower tiles only (DMMA GEMM, K = NB).
//   All launches are asynchronous on one stream; failure (non-positive pivot) lands in *info.
#include "common.cuh"
#include <mutex>
#include <vector>
#include <cstdlib>

namespace pb {

namespace {

constexpr int LEAF = 64;
constexpr int LDS = LEAF + 1;   // padded smem leading dimension

// One CTA (256 threads): factor the nv x nv (nv <= 64) diagonal block at A (lda) in place and write
// the 64x64 inverse of the (identity-padded) factor to Dinv (row-major, ld 64).
//
// Right-looking, register-resident: thread (ti, tk) = (tid/16, tid%16) owns the 4x4 elements
// (ti + 16a, tk + 16b) of the (symmetrised) block in registers for the whole factorisation.  Each of
// the 64 steps publishes one column through a double-buffered shared-memory vector (one
// __syncthreads per step), forms the pivot's reciprocal square root once per thread, and applies the
// rank-1 update to the register tile.  All loops are rolled: the kernel is ~1.5k instructions, so it
// stays inside the instruction cache (a fully unrolled warp-shuffle variant was 27k instructions and
// ran 94 us per leaf, instruction-fetch bound).
// The inverse is then built row by row (left-looking forward substitution, 4 threads per column).
__global__ void __launch_bounds__(256, 1)
potrf_leaf_kernel(double* __restrict__ A, int64_t lda, int nv, double* __restrict__ Dinv, int32_t* info, int col0) {
    extern __shared__ double leaf_smem[];
    double* Ls = leaf_smem;                                   // the factor L          [64][65]
    double* Vs = Ls + LEAF * LDS;                             // its inverse           [64][65]
    double (*col)[LEAF] = reinterpret_cast<double (*)[LEAF]>(Vs + LEAF * LDS);   // published pivot column [2][64]
    double* rdiag = Vs + LEAF * LDS + 2 * LEAF;               // 1 / L_jj              [64]
    const int tid = threadIdx.x, ti = tid >> 4, tk = tid & 15;

    double reg[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int i = ti + 16 * a, k = tk + 16 * b;
            const int r = i > k ? i : k, c = i > k ? k : i;        // symmetrise from the stored lower triangle
            double v = (i == k) ? 1.0 : 0.0;                      // identity padding beyond nv
            if (r < nv) v = A[(int64_t)r * lda + c];
            reg[a][b] = v;
        }
    for (int e = tid; e < LEAF * LDS; e += 256) Ls[e] = 0.0;
    if (tk == 0) {
#pragma unroll
        for (int a = 0; a < 4; ++a) col[0][ti + 16 * a] = reg[a][0];
    }
    __syncthreads();

    for (int j = 0; j < LEAF; ++j) {
        const double* cb = col[j & 1];
        const double d = cb[j];
        if (tid == 0 && !(d > 0.0)) atomicCAS(info, 0, col0 + j + 1);
        const double rinv = rsqrt(d);
        const double rd = rinv * rinv;
        double ci[4], ck[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) ci[a] = cb[ti + 16 * a];
#pragma unroll
        for (int b = 0; b < 4; ++b) ck[b] = cb[tk + 16 * b];
        if (tk == (j & 15)) {                                      // owners of column j emit L[:, j]
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const int i = ti + 16 * a;
                if (i > j) Ls[i * LDS + j] = ci[a] * rinv;
                else if (i == j) { Ls[i * LDS + j] = d * rinv; rdiag[j] = rinv; }
            }
        }
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const double t = ci[a] * rd;
#pragma unroll
            for (int b = 0; b < 4; ++b) reg[a][b] = fma(-t, ck[b], reg[a][b]);
        }
        if (j + 1 < LEAF && tk == ((j + 1) & 15)) {                // publish column j+1
            const int bsel = (j + 1) >> 4;
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                double v = reg[a][0];
#pragma unroll
                for (int b = 1; b < 4; ++b) v = (bsel == b) ? reg[a][b] : v;
                col[(j + 1) & 1][ti + 16 * a] = v;
            }
        }
        __syncthreads();
    }

    // write the factor back (lower triangle of the valid part only)
    for (int e = tid; e < LEAF * LEAF; e += 256) {
        const int i = e >> 6, k = e & 63;
        if (i < nv && k <= i) A[(int64_t)i * lda + k] = Ls[i * LDS + k];
    }

    // inverse V = L^{-1}: column c handled by 4 threads (q = k mod 4 slices of the dot product);
    // V is accumulated in place of the register tile's smem image: Vs[i][c].
    {
        const int c = tid >> 2, q = tid & 3;
        for (int i = 0; i < LEAF; ++i) {
            double s = 0.0;
            for (int k = c + q; k < i; k += 4) s = fma(Ls[i * LDS + k], Vs[k * LDS + c], s);   // V[k][c] = 0 for k < c
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            if (q == 0) Vs[i * LDS + c] = (i < c) ? 0.0 : ((i == c ? 1.0 : 0.0) - s) * rdiag[i];
            __syncwarp();          // the 4 threads of a column live in one warp; columns are independent
        }
    }
    __syncthreads();
    for (int e = tid; e < LEAF * LEAF; e += 256) {
        const int i = e >> 6, k = e & 63;
        Dinv[e] = (k <= i) ? Vs[i * LDS + k] : 0.0;
    }
}


constexpr int LEAF_SMEM = (2 * LEAF * LDS + 3 * LEAF) * 8;

// Inverse of every 64x64 diagonal leaf of an ALREADY factored matrix (one CTA per leaf): the same row-by-row
// substitution as in potrf_leaf_kernel.  Used when the factor was produced elsewhere (multi-GPU Cholesky).
__global__ void __launch_bounds__(256, 1)
leaf_inverse_kernel(const double* __restrict__ L, int64_t n, int64_t ldl, double* __restrict__ dinv) {
    extern __shared__ double leaf_smem[];
    double* Ls = leaf_smem;
    double* Vs = Ls + LEAF * LDS;
    double* rdiag = Vs + LEAF * LDS;
    const int tid = threadIdx.x;
    const int64_t j0 = (int64_t)blockIdx.x * LEAF;
    const int nv = (int)(n - j0 < LEAF ? n - j0 : LEAF);
    for (int e = tid; e < LEAF * LEAF; e += 256) {
        const int i = e >> 6, k = e & 63;
        double v = (i == k) ? 1.0 : 0.0;
        if (i < nv && k <= i) v = L[(j0 + i) * ldl + j0 + k];
        else if (k > i) v = 0.0;
        Ls[i * LDS + k] = v;
    }
    __syncthreads();
    if (tid < LEAF) rdiag[tid] = 1.0 / Ls[tid * LDS + tid];
    __syncthreads();
    {
        const int c = tid >> 2, q = tid & 3;
        for (int i = 0; i < LEAF; ++i) {
            double s = 0.0;
            for (int k = c + q; k < i; k += 4) s = fma(Ls[i * LDS + k], Vs[k * LDS + c], s);
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            if (q == 0) Vs[i * LDS + c] = (i < c) ? 0.0 : ((i == c ? 1.0 : 0.0) - s) * rdiag[i];
            __syncwarp();
        }
    }
    __syncthreads();
    double* out = dinv + (int64_t)blockIdx.x * LEAF * LEAF;
    for (int e = tid; e < LEAF * LEAF; e += 256) {
        const int i = e >> 6, k = e & 63;
        out[e] = (k <= i) ? Vs[i * LDS + k] : 0.0;
    }
}

struct Ctx {
    cudaStream_t stream;
    int64_t lda;
    double* dinv;     // (n/64) leaf inverses, 64*64 doubles each, indexed by global column / 64
    int32_t* info;
    void* oz_scratch = nullptr;   // non-null: the K >= 1024 GEMMs of the solve run on the INT8 tensor cores (ozaki.cu)
    int64_t oz_bytes = 0;
};
constexpr int64_t OZ_TRSM_K = 1024;

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
    int64_t k = 0;
    if (c.oz_scratch && n1 >= OZ_TRSM_K && m >= 256 && n2 >= 256 &&
        ozaki_scratch_bytes(m, OZ_TRSM_K) + ozaki_scratch_bytes(n2, OZ_TRSM_K) <= c.oz_bytes) {
        // B2 -= B1 L21^T in K-blocks of 1024: each block slices its m x 1024 and n2 x 1024 operands once (every entry of L
        // takes part in exactly one GEMM of the recursion, so the factor is sliced once per solve)
        for (; k + OZ_TRSM_K <= n1; k += OZ_TRSM_K)
            PB_TRY(ozaki_gemm_nt(c.stream, m, n2, OZ_TRSM_K, -1.0, B + k, ldb, L + n1 * c.lda + k, c.lda, B + n1, ldb, c.oz_scratch,
                                 c.oz_bytes));
    }
    if (k < n1)
        PB_TRY(gemm_nt(c.stream, m, n2, n1 - k, -1.0, B + k, ldb, L + n1 * c.lda + k, c.lda, 1.0, B + n1, ldb, false));
    return trsm_rec(c, B + n1, ldb, m, L + n1 * c.lda + n1, n2, col0 + n1);
}

int potrf_rec(const Ctx& c, double* A, int64_t n, int64_t col0) {
    if (n <= LEAF) {
        static PerDeviceOnce configured;
        if (configured.first()) {
            PB_CUDA(cudaFuncSetAttribute(potrf_leaf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LEAF_SMEM));
        }
        potrf_leaf_kernel<<<1, 256, LEAF_SMEM, c.stream>>>(A, c.lda, (int)n, c.dinv + (col0 / LEAF) * LEAF * LEAF,
                                                           c.info, (int)col0); pb::note_launch();
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

// workspace = [leaf inverses | 256-block inverses and their transposes] (the solves read these) | slicing scratch of the
// INT8 trailing update (n >= OZ_MIN_N only)
constexpr int64_t OZ_MIN_N = 4096;
int64_t potrf_inverse_bytes(int64_t n) {
    const int64_t leaves = (n + LEAF - 1) / LEAF;
    const int64_t blocks = (n + 255) / 256;
    return ((leaves > 0 ? leaves : 1) * LEAF * LEAF + 2 * (blocks > 0 ? blocks : 1) * 256 * 256) * (int64_t)sizeof(double);
}
int potrf_block_size(int64_t n);
// two slicing buffers: the look-ahead slices panel k + 1 while the trailing update of step k still reads panel k's planes
int64_t potrf_ozaki_bytes(int64_t n) { return n >= OZ_MIN_N ? 2 * ozaki_scratch_bytes(n, 1024) : 0; }
bool ozaki_enabled(int64_t n) {
    const int mode = opts().potrf_ozaki;
    return n >= OZ_MIN_N && (mode == 1 || (mode < 0 && n >= 8192));
}

int potrf_block_size(int64_t n) {
    const int forced = opt_potrf_nb();
    if (forced > 0) return forced / 64 * 64 > 0 ? forced / 64 * 64 : 64;
    // measured on B200 (tools/potrf_sweep.py): N=65536: 512 -> 34.2, 768 -> 34.7, 1024 -> 34.9 TFLOP/s;
    // deeper panels raise the SYRK's K (fewer C round trips) and the look-ahead hides the longer panel chain
    if (n <= 2048) return 256;
    if (n < 24576) return 512;
    if (n < 49152) return 768;
    return 1024;
}

namespace {

// Highest-priority non-blocking side stream (one per device) for the look-ahead panel work.
int side_stream(cudaStream_t* out) {
    static std::mutex mu;
    static cudaStream_t streams[64] = {};
    int dev = 0;
    PB_CUDA(cudaGetDevice(&dev));
    PB_CHECK(dev >= 0 && dev < 64, PB_ERR_INVALID, "device index out of range");
    std::lock_guard<std::mutex> lock(mu);
    if (!streams[dev]) {
        int lo = 0, hi = 0;
        PB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        PB_CUDA(cudaStreamCreateWithPriority(&streams[dev], cudaStreamNonBlocking, hi));
    }
    *out = streams[dev];
    return PB_OK;
}

struct EventPool {
    std::vector<cudaEvent_t> evs;
    ~EventPool() { for (cudaEvent_t e : evs) cudaEventDestroy(e); }
    int get(cudaEvent_t* e) {
        PB_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
        evs.push_back(*e);
        return PB_OK;
    }
};

bool lookahead_enabled() { return opt_lookahead(); }

}  // namespace

// Right-looking blocked factorisation with one-panel look-ahead.
//   main stream : trailing SYRK of step k restricted to the columns right of panel k+1
//   side stream : (a) update of panel k+1's block column with panel k, (b) factorisation of panel k+1
// so the latency-bound panel chain (leaf kernels, small GEMMs) runs underneath the big SYRK.
static int potrf_eager(cudaStream_t stream, cudaStream_t side, double* A, int64_t n, int64_t lda, void* workspace,
                       int32_t* info) {
    PB_CUDA(cudaMemsetAsync(info, 0, sizeof(int32_t), stream));
    if (n == 0) return PB_OK;
    const int64_t NB = potrf_block_size(n);
    Ctx cm{stream, lda, reinterpret_cast<double*>(workspace), info};

    if (!lookahead_enabled() || n <= 2 * NB) {
        for (int64_t k0 = 0; k0 < n; k0 += NB) {
            const int64_t nb = n - k0 < NB ? n - k0 : NB;
            double* Akk = A + k0 * lda + k0;
            PB_TRY(potrf_rec(cm, Akk, nb, k0));
            const int64_t m = n - k0 - nb;
            if (m > 0) {
                double* P = A + (k0 + nb) * lda + k0;
                PB_TRY(trsm_rec(cm, P, lda, m, Akk, nb, k0));
                PB_TRY(gemm_nt(stream, m, m, nb, -1.0, P, lda, P, lda, 1.0, P + nb, lda, true));
            }
        }
        return build_block_inverses(stream, A, n, lda, cm.dinv);
    }

    Ctx cs{side, lda, reinterpret_cast<double*>(workspace), info};
    // Trailing updates on the INT8 tensor cores (ozaki.cu).  Panel k (all its rows below the diagonal block) is sliced
    // ONCE, right after it is factored, into buffer k % 2 behind the inverses in the workspace; the side stream's update
    // of block column k + 1 and the main stream's SYRK both multiply those planes.  Buffer k % 2 is rewritten at step
    // k + 2 on the side stream, which by then has waited for the SYRK of step k (ev_trail) in step k + 1.
    const bool oz = ozaki_enabled(n) && ozaki_supported(NB) && NB <= 1024;
    const int64_t oz_half = potrf_ozaki_bytes(n) / 2;
    uint8_t* oz_base = reinterpret_cast<uint8_t*>(workspace) + potrf_inverse_bytes(n);
    constexpr int64_t OZ_MIN_ROWS = 2048;            // smaller trailing blocks stay on the DMMA kernel
    EventPool pool;
    cudaEvent_t ev_panel, ev_trail = nullptr;
    int step = 0;
    auto slice_panel = [&](cudaStream_t st, int64_t k0, int64_t k1, int buf) -> int {   // rows k1.. of columns [k0, k1)
        return ozaki_slice(st, A + k1 * lda + k0, n - k1, k1 - k0, lda, oz_base + buf * oz_half, oz_half);
    };

    // panel 0 on the main stream
    {
        const int64_t nb = n < NB ? n : NB;
        PB_TRY(potrf_rec(cm, A, nb, 0));
        PB_TRY(trsm_rec(cm, A + nb * lda, lda, n - nb, A, nb, 0));
        if (oz && n - nb >= OZ_MIN_ROWS) PB_TRY(slice_panel(stream, 0, nb, 0));
        PB_TRY(pool.get(&ev_panel));
        PB_CUDA(cudaEventRecord(ev_panel, stream));
    }
    for (int64_t k0 = 0; k0 + NB < n; k0 += NB, ++step) {
        const int64_t nb = NB;                       // panel k is full width here (there are rows below it)
        const int64_t k1 = k0 + nb;                  // first row/col of panel k+1
        const int64_t nb1 = n - k1 < NB ? n - k1 : NB;
        const int64_t k2 = k1 + nb1;                 // first row/col right of panel k+1
        const int64_t m2 = n - k2;
        double* P1 = A + k1 * lda + k0;              // rows of panel k belonging to block row k+1 (nb1 x nb)
        double* P2 = A + k2 * lda + k0;              // rows of panel k below that (m2 x nb)
        const bool oz_k = oz && n - k1 >= OZ_MIN_ROWS;                     // panel k was sliced (buffer step % 2)
        const bool oz_next = oz && nb1 == NB && n - k2 >= OZ_MIN_ROWS;     // panel k+1 will be
        const void* planes = oz_base + (step % 2) * oz_half;

        // ---- side: bring block column k+1 up to date with panel k, then factor it ----
        PB_CUDA(cudaStreamWaitEvent(side, ev_panel, 0));
        if (ev_trail) PB_CUDA(cudaStreamWaitEvent(side, ev_trail, 0));
        double* A11 = A + k1 * lda + k1;
        PB_TRY(gemm_nt(side, nb1, nb1, nb, -1.0, P1, lda, P1, lda, 1.0, A11, lda, true));
        if (m2 > 0) {
            if (oz_k && m2 >= OZ_MIN_ROWS)
                PB_TRY(ozaki_apply(side, nb, planes, n - k1, nb1, m2, planes, n - k1, 0, nb1, -1.0, A + k2 * lda + k1, lda, false));
            else
                PB_TRY(gemm_nt(side, m2, nb1, nb, -1.0, P2, lda, P1, lda, 1.0, A + k2 * lda + k1, lda, false));
        }
        PB_TRY(potrf_rec(cs, A11, nb1, k1));
        if (m2 > 0) PB_TRY(trsm_rec(cs, A + k2 * lda + k1, lda, m2, A11, nb1, k1));
        if (oz_next) PB_TRY(slice_panel(side, k1, k2, (step + 1) % 2));
        cudaEvent_t ev_next;
        PB_TRY(pool.get(&ev_next));
        PB_CUDA(cudaEventRecord(ev_next, side));

        // ---- main: the rest of the trailing update with panel k ----
        PB_CUDA(cudaStreamWaitEvent(stream, ev_panel, 0));
        if (m2 > 0) {
            if (oz_k && m2 >= OZ_MIN_ROWS)
                PB_TRY(ozaki_apply(stream, nb, planes, n - k1, nb1, m2, planes, n - k1, nb1, m2, -1.0, A + k2 * lda + k2, lda, true));
            else
                PB_TRY(gemm_nt(stream, m2, m2, nb, -1.0, P2, lda, P2, lda, 1.0, A + k2 * lda + k2, lda, true));
            PB_TRY(pool.get(&ev_trail));
            PB_CUDA(cudaEventRecord(ev_trail, stream));
        }
        ev_panel = ev_next;
    }
    PB_CUDA(cudaStreamWaitEvent(stream, ev_panel, 0));
    return build_block_inverses(stream, A, n, lda, cm.dinv);
}

// ---- CUDA-graph replay of the factorisation's launch DAG ----
// One factorisation is 10^2 .. 10^4 small dependent launches on two streams (N = 16384: ~4 k, each with two
// cuTensorMapEncodeTiled calls on the host); issued eagerly the host falls behind the GPU for N <~ 32768 and the
// panel chain stalls on launch latency.  The DAG depends only on (A, n, lda, workspace, info, block size,
// look-ahead), and the drivers call potrf with the SAME buffers step after step (caller-owned workspaces), so the
// second call with a given key captures the two streams into a graph (capture on an internal stream: the caller's
// may be the legacy default stream, which cannot be captured) and every later call is ONE cudaGraphLaunch.
namespace {

struct GraphKey {
    int dev;
    const void *A, *ws, *info;
    int64_t n, lda;
    int nb, lookahead;
    bool operator==(const GraphKey& o) const {
        return dev == o.dev && A == o.A && ws == o.ws && info == o.info && n == o.n && lda == o.lda && nb == o.nb &&
               lookahead == o.lookahead;
    }
};

struct GraphEntry {
    GraphKey key{};
    cudaGraphExec_t exec = nullptr;
    double gemm_flops = 0;            // algorithmic flops of the main-config GEMM launches inside (profiling hook)
    long long gemm_launches = 0, launches = 0;
    unsigned long long stamp = 0;
    bool used = false;
};

constexpr int GRAPH_SLOTS = 16;
std::mutex g_graph_mu;
GraphEntry g_graphs[GRAPH_SLOTS];
unsigned long long g_graph_clock = 0;
constexpr int64_t GRAPH_MIN_N = 512;     // below this a factorisation is a handful of launches

int capture_stream(cudaStream_t* out) {
    static std::mutex mu;
    static cudaStream_t streams[64] = {};
    int dev = 0;
    PB_CUDA(cudaGetDevice(&dev));
    PB_CHECK(dev >= 0 && dev < 64, PB_ERR_INVALID, "device index out of range");
    std::lock_guard<std::mutex> lock(mu);
    if (!streams[dev]) PB_CUDA(cudaStreamCreateWithFlags(&streams[dev], cudaStreamNonBlocking));
    *out = streams[dev];
    return PB_OK;
}

}  // namespace

int potrf(cudaStream_t stream, double* A, int64_t n, int64_t lda, void* workspace, int64_t workspace_bytes,
          int32_t* info) {
    PB_CHECK(n >= 0 && lda >= n, PB_ERR_INVALID, "potrf: bad n/lda");
    PB_CHECK(workspace_bytes >= pb_potrf_workspace_bytes(n), PB_ERR_INVALID, "potrf: workspace too small");
    cudaStream_t side;
    PB_TRY(side_stream(&side));
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (stream != nullptr && stream != cudaStreamLegacy && stream != cudaStreamPerThread)
        cudaStreamIsCapturing(stream, &cap);                 // the caller is capturing us into ITS graph: just issue
    if (!opts().potrf_graph || n < GRAPH_MIN_N || cap != cudaStreamCaptureStatusNone)
        return potrf_eager(stream, side, A, n, lda, workspace, info);

    GraphKey key{};
    PB_CUDA(cudaGetDevice(&key.dev));
    key.A = A; key.ws = workspace; key.info = info; key.n = n; key.lda = lda;
    key.nb = potrf_block_size(n); key.lookahead = lookahead_enabled() ? 1 : 0;

    std::unique_lock<std::mutex> lock(g_graph_mu);
    GraphEntry* hit = nullptr;
    GraphEntry* victim = &g_graphs[0];
    for (GraphEntry& e : g_graphs) {
        if (e.used && e.key == key) { hit = &e; break; }
        if (!e.used) { if (victim->used) victim = &e; }
        else if (victim->used && e.stamp < victim->stamp) victim = &e;
    }
    if (!hit) {
        // first sighting of this key: remember it and run eagerly (one-shot callers never pay for an instantiation)
        if (victim->exec) { cudaGraphExecDestroy(victim->exec); }
        *victim = GraphEntry{};
        victim->key = key;
        victim->used = true;
        victim->stamp = ++g_graph_clock;
        lock.unlock();
        return potrf_eager(stream, side, A, n, lda, workspace, info);
    }
    hit->stamp = ++g_graph_clock;
    if (!hit->exec) {
        cudaStream_t cs;
        PB_TRY(capture_stream(&cs));
        // every >48 KB kernel attribute must be set before capture starts: one eager factorisation already ran (above)
        CaptureTally tally;
        PB_CUDA(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
        set_capture_tally(&tally);
        const long long launches0 = launch_count();
        const int rc = potrf_eager(cs, side, A, n, lda, workspace, info);
        const long long launched = launch_count() - launches0;
        set_capture_tally(nullptr);
        cudaGraph_t graph = nullptr;
        const cudaError_t ce = cudaStreamEndCapture(cs, &graph);
        if (rc != PB_OK || ce != cudaSuccess || !graph) {
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            hit->used = false;                               // do not try again with this key
            lock.unlock();
            if (rc != PB_OK) return rc;
            return potrf_eager(stream, side, A, n, lda, workspace, info);
        }
        cudaGraphExec_t exec = nullptr;
        const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ie != cudaSuccess) {
            cudaGetLastError();
            hit->used = false;
            lock.unlock();
            return potrf_eager(stream, side, A, n, lda, workspace, info);
        }
        hit->exec = exec;
        hit->gemm_flops = tally.flops;
        hit->gemm_launches = tally.launches;
        hit->launches = launched;
        note_launches(-launched);                            // capture issued nothing; replays are counted below
    }
    cudaGraphExec_t exec = hit->exec;
    const double flops = hit->gemm_flops;
    const long long gl = hit->gemm_launches, nl = hit->launches;
    lock.unlock();
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    const bool prof = profiling_enabled() && gl > 0;
    if (prof) {
        PB_CUDA(cudaEventCreate(&e0));
        PB_CUDA(cudaEventCreate(&e1));
        PB_CUDA(cudaEventRecord(e0, stream));
    }
    PB_CUDA(cudaGraphLaunch(exec, stream));
    note_launches(nl);
    if (prof) {
        PB_CUDA(cudaEventRecord(e1, stream));
        profile_gemm(e0, e1, flops, gl);                     // whole replay: GEMMs + leaves + gaps (a lower bound on the GEMM rate)
    }
    return PB_OK;
}

int rebuild_solve_workspace(cudaStream_t stream, const double* L, int64_t n, int64_t ldl, void* workspace,
                            int64_t workspace_bytes) {
    PB_CHECK(workspace_bytes >= pb_potrf_workspace_bytes(n), PB_ERR_INVALID, "rebuild_solve_workspace: workspace too small");
    if (n == 0) return PB_OK;
    static PerDeviceOnce configured;
    if (configured.first()) {
        PB_CUDA(cudaFuncSetAttribute(leaf_inverse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LEAF_SMEM));
    }
    double* dinv = reinterpret_cast<double*>(workspace);
    leaf_inverse_kernel<<<(unsigned)((n + LEAF - 1) / LEAF), 256, LEAF_SMEM, stream>>>(L, n, ldl, dinv); pb::note_launch();
    PB_CUDA(cudaGetLastError());
    return build_block_inverses(stream, L, n, ldl, dinv);
}

int trsm_right_lt(cudaStream_t stream, const double* L, int64_t n, int64_t ldl, const void* potrf_workspace,
                  double* X, int64_t m, int64_t ldx) {
    Ctx c{stream, ldl, const_cast<double*>(reinterpret_cast<const double*>(potrf_workspace)), nullptr};
    if (ozaki_enabled(n)) {        // the workspace is the one potrf(n) used: the slicing scratch sits behind the inverses
        c.oz_scratch = const_cast<uint8_t*>(reinterpret_cast<const uint8_t*>(potrf_workspace)) + potrf_inverse_bytes(n);
        c.oz_bytes = potrf_ozaki_bytes(n);
    }
    return trsm_rec(c, X, ldx, m, L, n, 0);
}

}  // namespace pb

extern "C" int64_t pb_potrf_workspace_bytes(int64_t n) {
    // [ceil(n/64) leaf inverses, 64x64] followed by [ceil(n/256) diagonal-block inverses, 256x256] and their
    // transposes (blas2.cu)
    return pb::potrf_inverse_bytes(n) + pb::potrf_ozaki_bytes(n);
}

extern "C" int pb_potrf(pb_stream_t stream, double* A, int64_t n, int64_t lda, void* workspace,
                        int64_t workspace_bytes, int32_t* info, const pb_options* options) {
    pb::OptScope opt_scope(options);
    return pb::potrf(reinterpret_cast<cudaStream_t>(stream), A, n, lda, workspace, workspace_bytes, info);
}

extern "C" int pb_rebuild_solve_workspace(pb_stream_t stream, const double* L, int64_t n, int64_t ldl, void* workspace,
                                          int64_t workspace_bytes) {
    return pb::rebuild_solve_workspace(reinterpret_cast<cudaStream_t>(stream), L, n, ldl, workspace, workspace_bytes);
}

extern "C" int pb_trsm_right_lt(pb_stream_t stream, const double* L, int64_t n, int64_t ldl,
                                const void* potrf_workspace, double* X, int64_t m, int64_t ldx) {
    return pb::trsm_right_lt(reinterpret_cast<cudaStream_t>(stream), L, n, ldl, potrf_workspace, X, m, ldx);
}
