// FP64 tensor-core GEMM for sm_100a:  C = alpha * A * B^T + beta * C   (row-major, "NT").
//
// This is the contraction behind the Cholesky trailing update (SYRK), the panel TRSMs and the
// predict variance solve — i.e. >95% of the flops of the hot path.
// tcgen05.mma has no f64 kind, so FP64 tensor work on Blackwell is mma.sync -> SASS DMMA.8x8x4
// (measured peak 37.0 TFLOP/s on B200, profiles/r01_fp64_peaks.json).
//
// Design
//  * CTA tile 128x128, K step 16 doubles (= one 128-byte TMA swizzle span).
//  * One producer warp: a single elected lane drives a STAGES-deep ring of TMA box loads
//    (cp.async.bulk.tensor.2d, SWIZZLE_128B) signalled through mbarriers (full/empty pairs).
//  * Eight consumer warps (2 x 4), warp tile 64x32 -> 32 m8n8k4 accumulators (64 doubles/lane).
//  * Bank-conflict-free fragment loads under the 128B swizzle: the 8 rows of an m8 fragment are
//    taken in the order perm(g) = 2*(g&3) + (g>>2), so the 16 lanes of each LDS.64 phase touch
//    16 distinct 8-byte bank pairs.  The permutation is undone when C is addressed.
//  * TMA zero-fills out-of-bounds boxes, so ragged M/N/K need no special main-loop code; the
//    epilogue masks stores.
//  * `lower_only` enumerates only tiles with tile_n <= tile_m in L2-friendly column strips.
#include "common.cuh"

namespace pb {

namespace {

constexpr int BM = 128, BN = 128, BK = 16;
constexpr int STAGES = 6;
constexpr int CONSUMER_WARPS = 8;
constexpr int THREADS = (CONSUMER_WARPS + 1) * 32;
constexpr int A_STAGE_BYTES = BM * BK * 8;  // 16 KB
constexpr int B_STAGE_BYTES = BN * BK * 8;  // 16 KB
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 2 * STAGES * 8;
constexpr int STRIP_W = 12;  // tiles per column strip in the lower-only rasterisation

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
__device__ __forceinline__ double lds64(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ int perm8(int g) { return 2 * (g & 3) + (g >> 2); }

// Linear block id -> (tile_m, tile_n) over the lower triangle of a T x T tile grid, enumerated in
// column strips of STRIP_W tiles (row-major inside a strip) so that a wave of CTAs shares few
// operand panels.
__device__ __forceinline__ void lower_tile(int bid, int T, int& tm, int& tn) {
    int s = 0;
    while (true) {
        int c0 = s * STRIP_W;
        int w = min(STRIP_W, T - c0);
        int head = w * (w + 1) / 2;               // triangular head: rows c0 .. c0+w-1
        int cnt = head + (T - c0 - w) * w;        // + full rows below
        if (bid < cnt) {
            if (bid < head) {
                int r = (int)((sqrtf(8.0f * bid + 1.0f) - 1.0f) * 0.5f);
                while ((r + 1) * (r + 2) / 2 <= bid) ++r;
                while (r * (r + 1) / 2 > bid) --r;
                tm = c0 + r;
                tn = c0 + bid - r * (r + 1) / 2;
            } else {
                int q = bid - head;
                tm = c0 + w + q / w;
                tn = c0 + q % w;
            }
            return;
        }
        bid -= cnt;
        ++s;
    }
}

__global__ void __launch_bounds__(THREADS, 1)
gemm_nt_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
               double* __restrict__ C, int64_t ldc, int M, int N, int K, double alpha, double beta,
               int lower_only, int tiles_n) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + STAGES * STAGE_BYTES;   // full[s] at +8s, empty[s] at +8(STAGES+s)

    int tile_m, tile_n;
    if (lower_only) {
        lower_tile(blockIdx.x, tiles_n, tile_m, tile_n);
    } else {
        tile_m = blockIdx.x / tiles_n;
        tile_n = blockIdx.x % tiles_n;
    }
    const int m0 = tile_m * BM, n0 = tile_n * BN;
    const int kblocks = (K + BK - 1) / BK;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bar_base + 8 * s, 1);
            mbar_init(bar_base + 8 * (STAGES + s), CONSUMER_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == CONSUMER_WARPS) {
        // ===== TMA producer =====
        if (lane == 0) {
            for (int kb = 0; kb < kblocks; ++kb) {
                const int s = kb % STAGES;
                const uint32_t round = kb / STAGES;
                if (kb >= STAGES) mbar_wait(bar_base + 8 * (STAGES + s), (round - 1) & 1);
                const uint32_t full = bar_base + 8 * s;
                mbar_expect_tx(full, STAGE_BYTES);
                const uint32_t dstA = smem_base + s * STAGE_BYTES;
                tma_load_2d(dstA, &mapA, kb * BK, m0, full);
                tma_load_2d(dstA + A_STAGE_BYTES, &mapB, kb * BK, n0, full);
            }
        }
        return;
    }

    // ===== consumers =====
    const int wm = warp >> 2, wn = warp & 3;      // 2 x 4 warps, warp tile 64 x 32
    const int g = lane >> 2, t = lane & 3;
    const int pg = perm8(g);

    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    // byte offset inside a tile of (row = base + 8*i + pg, k = 4*ks + t) under SWIZZLE_128B
    uint32_t koff[4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) koff[ks] = ((((2 * ks) | (t >> 1)) ^ pg) << 4) | ((t & 1) << 3);
    const uint32_t a_row = (wm * 64 + pg) * 128;
    const uint32_t b_row = A_STAGE_BYTES + (wn * 32 + pg) * 128;

    for (int kb = 0; kb < kblocks; ++kb) {
        const int s = kb % STAGES;
        mbar_wait(bar_base + 8 * s, (kb / STAGES) & 1);
        const uint32_t st = smem_base + s * STAGE_BYTES;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            double a[8], b[4];
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = lds64(st + a_row + i * 1024 + koff[ks]);
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = lds64(st + b_row + j * 1024 + koff[ks]);
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_base + 8 * (STAGES + s));
    }

    // ===== epilogue: C = alpha * acc + beta * C (direct global access, permutation undone) =====
    const int c0 = perm8(2 * t), c1 = perm8(2 * t + 1);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int row = m0 + wm * 64 + 8 * i + pg;
        if (row >= M) continue;
        double* crow = C + (int64_t)row * ldc;
        const int col_lim = lower_only ? min(N, row + 1) : N;
        double old[4][2];
        if (beta != 0.0) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int cb = n0 + wn * 32 + 8 * j;
                old[j][0] = (cb + c0 < col_lim) ? crow[cb + c0] : 0.0;
                old[j][1] = (cb + c1 < col_lim) ? crow[cb + c1] : 0.0;
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int cb = n0 + wn * 32 + 8 * j;
            double v0 = alpha * acc[i][j][0], v1 = alpha * acc[i][j][1];
            if (beta != 0.0) { v0 += beta * old[j][0]; v1 += beta * old[j][1]; }
            if (cb + c0 < col_lim) crow[cb + c0] = v0;
            if (cb + c1 < col_lim) crow[cb + c1] = v1;
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 2-D f64 tensor map over a row-major (rows x cols) matrix, box = (BK cols) x (box_rows rows).
int make_map(CUtensorMap* map, const double* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
    EncodeTiledFn enc = get_encode();
    PB_CHECK(enc != nullptr, PB_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
    PB_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, PB_ERR_INVALID, "gemm operand not 16-byte aligned");
    PB_CHECK((ld & 1) == 0, PB_ERR_INVALID, "gemm leading dimension must be even (got %lld)", (long long)ld);
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 8};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PB_CHECK(r == CUDA_SUCCESS, PB_ERR_CUDA, "cuTensorMapEncodeTiled failed with %d (rows=%lld cols=%lld ld=%lld)",
             (int)r, (long long)rows, (long long)cols, (long long)ld);
    return PB_OK;
}

}  // namespace

int gemm_nt(cudaStream_t stream, int64_t M, int64_t N, int64_t K, double alpha, const double* A, int64_t lda,
            const double* B, int64_t ldb, double beta, double* C, int64_t ldc, bool lower_only) {
    if (M <= 0 || N <= 0) return PB_OK;
    PB_CHECK(K > 0, PB_ERR_INVALID, "gemm_nt: K must be positive");
    PB_CHECK(!lower_only || M == N, PB_ERR_INVALID, "gemm_nt: lower_only needs a square C");
    PB_CHECK(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), PB_ERR_INVALID, "gemm_nt: dimension too large");
    static bool configured = false;
    if (!configured) {
        PB_CUDA(cudaFuncSetAttribute(gemm_nt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        configured = true;
    }
    CUtensorMap mapA, mapB;
    PB_TRY(make_map(&mapA, A, M, K, lda, BM));
    PB_TRY(make_map(&mapB, B, N, K, ldb, BN));
    const int tm = (int)ceil_div<int64_t>(M, BM), tn = (int)ceil_div<int64_t>(N, BN);
    int64_t blocks;
    if (lower_only) {
        blocks = 0;
        for (int c0 = 0; c0 < tn; c0 += STRIP_W) {
            int w = tn - c0 < STRIP_W ? tn - c0 : STRIP_W;
            blocks += (int64_t)w * (w + 1) / 2 + (int64_t)(tn - c0 - w) * w;
        }
    } else {
        blocks = (int64_t)tm * tn;
    }
    PB_CHECK(blocks < (1ll << 31), PB_ERR_INVALID, "gemm_nt: too many tiles");
    gemm_nt_kernel<<<(unsigned)blocks, THREADS, SMEM_BYTES, stream>>>(mapA, mapB, C, ldc, (int)M, (int)N, (int)K, alpha,
                                                                      beta, lower_only ? 1 : 0, tn);
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

}  // namespace pb

extern "C" int pb_gemm_nt(pb_stream_t stream, int64_t M, int64_t N, int64_t K, double alpha, const double* A,
                          int64_t lda, const double* B, int64_t ldb, double beta, double* C, int64_t ldc,
                          int32_t lower_only) {
    return pb::gemm_nt(reinterpret_cast<cudaStream_t>(stream), M, N, K, alpha, A, lda, B, ldb, beta, C, ldc,
                       lower_only != 0);
}
