// FP64 tensor-core GEMM for sm_100a:  C = alpha * A * B^T + beta * C   (row-major, "NT").
//
// This is the contraction behind the Cholesky trailing update (SYRK), the panel TRSMs and the
// predict variance solve — i.e. >95% of the flops of the hot path.
// tcgen05.mma has no f64 kind, so FP64 tensor work on Blackwell is mma.sync -> SASS DMMA.8x8x4
// (measured peak 37.0 TFLOP/s on B200, profiles/r01_fp64_peaks.json).
//
// Design
//  * CTA tile 128x64 (two CTAs per SM), K step 16 doubles (= one 128-byte TMA swizzle span).
//  * Thread 0 drives a STAGES-deep ring of TMA box loads (cp.async.bulk.tensor.2d, SWIZZLE_128B)
//    signalled through mbarriers (full/empty pairs), keeping up to STAGES loads in flight.
//  * Eight warps (4 x 2), warp tile 32x32 -> 16 m8n8k4 accumulators (32 doubles/lane).
//  * Bank-conflict-free fragment loads under the 128B swizzle: the 8 rows of an m8 fragment are
//    taken in the order perm(g) = 2*(g&3) + (g>>2), so the 16 lanes of each LDS.64 phase touch
//    16 distinct 8-byte bank pairs.  The permutation is undone when C is addressed.
//  * TMA zero-fills out-of-bounds boxes, so ragged M/N/K need no special main-loop code; the
//    epilogue masks stores.
//  * `lower_only` enumerates only tiles with tile_n <= tile_m in L2-friendly column strips.
#include "common.cuh"
#include <cstdlib>
#include <algorithm>

namespace pb {

namespace {

constexpr int BK = 16;
constexpr int STRIP_W = 12;  // tiles per column strip in the lower-only rasterisation

// Tile configuration: CTA tile BM x BN, WM x WN warps.
template <int BM_, int BN_, int WM_, int WN_, int STAGES_, int MIN_CTAS_>
struct Cfg {
    static constexpr int BM = BM_, BN = BN_, WM = WM_, WN = WN_, STAGES = STAGES_, MIN_CTAS = MIN_CTAS_;
    static constexpr int CONSUMER_WARPS = WM * WN;
    static constexpr int THREADS = CONSUMER_WARPS * 32;
    static constexpr int WARP_M = BM / WM, WARP_N = BN / WN;     // warp tile
    static constexpr int MI = WARP_M / 8, NJ = WARP_N / 8;       // m8n8 accumulator tiles per warp
    static constexpr int A_STAGE_BYTES = BM * BK * 8;
    static constexpr int B_STAGE_BYTES = BN * BK * 8;
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 2 * STAGES * 8;
};
// Measured on B200 (SYRK 16384 x K, TFLOP/s of the 37.0 DMMA peak), profiles/r01_gemm_configs.md:
//   128x128, 8 warps of 64x32, 1 CTA/SM : K=512 32.3   K=8192 34.0   (prologue/epilogue exposed)
//   128x64,  4 warps of 64x32, 2 CTA/SM : K=512 32.6   (one warp per sub-partition cannot saturate DMMA)
//   128x64,  4 warps of 64x32, 3 CTA/SM : K=512 34.4
//   128x64,  8 warps of 32x32, 2 CTA/SM : K=512 34.7   K=1024 35.4   K=8192 35.1   <- CfgMain
// Two co-resident CTAs hide each other's prologue (C prefetch, pipeline fill) and epilogue, and the
// other CTA still has two warps per sub-partition, which is what it takes to keep the DMMA pipe full.
using CfgMain = Cfg<128, 64, 4, 2, 4, 2>;
using CfgSmall = Cfg<64, 64, 2, 2, 4, 3>;    // 4 warps of 32x32, 3 CTAs/SM: panel / leaf-sized problems

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
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
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

// Same for BM = 2*BN tiles (row-major over tile rows): tile row tm needs column tiles 0 .. 2*tm+1.
__device__ __forceinline__ void lower_tile_2to1(int bid, int& tm, int& tn) {
    int r = (int)((sqrtf(4.0f * bid + 1.0f) - 1.0f) * 0.5f);
    while ((r + 1) * (r + 2) <= bid) ++r;
    while (r * (r + 1) > bid) --r;
    tm = r;
    tn = bid - r * (r + 1);
}

// lower_only == 3: "grouped" mode for the block-column-cyclic trailing update of the multi-GPU Cholesky (dist.cu).
// A and B are the SAME panel (R rows); group q updates one owned block column:
//     C_q[M_q x N_q] += alpha * P[a0 + q astep :, :] * P[a0 + q astep : +N_q, :]^T,   M_q = R - (a0 + q astep), N_q = min(nb, M_q),
// C_q = c0 + q cstep (leading dimension ldc).  One launch covers every group, so a panel step is one grid of
// thousands of tiles instead of a launch per block column (each with its own ramp and tail).
struct GemmGroups {
    int count = 0;          // groups
    int R = 0;              // rows of the panel
    int a0 = 0, astep = 0;  // panel row of group 0 / increment per group
    int nb = 0;             // block column width
    int64_t cstep = 0;      // elements between the C origins of consecutive groups
};

template <class CF>
__global__ void __launch_bounds__(CF::THREADS, CF::MIN_CTAS)
gemm_nt_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
               double* __restrict__ C_arg, int64_t ldc, int M_arg, int N_arg, int K, double alpha, double beta,
               int lower_only, int tiles_n, const GemmGroups grp) {
    constexpr int BM = CF::BM, BN = CF::BN, STAGES = CF::STAGES, CONSUMER_WARPS = CF::CONSUMER_WARPS;
    constexpr int A_STAGE_BYTES = CF::A_STAGE_BYTES, STAGE_BYTES = CF::STAGE_BYTES;
    constexpr int MI = CF::MI, NJ = CF::NJ, WARP_M = CF::WARP_M, WARP_N = CF::WARP_N;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + STAGES * STAGE_BYTES;   // full[s] at +8s, empty[s] at +8(STAGES+s)

    int tile_m, tile_n;
    int M = M_arg, N = N_arg, a_row0 = 0;         // a_row0: panel row of this tile's group (grouped mode)
    double* C = C_arg;
    if (lower_only == 3) {
        int b = (int)blockIdx.x, q = 0;
        for (;; ++q) {
            const int Mq = grp.R - (grp.a0 + q * grp.astep);
            const int t = ((Mq + BM - 1) / BM) * tiles_n;
            if (b < t || q + 1 >= grp.count) break;
            b -= t;
        }
        a_row0 = grp.a0 + q * grp.astep;
        M = grp.R - a_row0;
        N = M < grp.nb ? M : grp.nb;
        C = C_arg + (int64_t)q * grp.cstep;
        tile_m = b / tiles_n;
        tile_n = b % tiles_n;
        lower_only = 0;
    } else if (lower_only) {
        if (CF::BM == CF::BN) lower_tile(blockIdx.x, tiles_n, tile_m, tile_n);
        else lower_tile_2to1(blockIdx.x, tile_m, tile_n);
    } else {
        tile_m = blockIdx.x / tiles_n;
        tile_n = blockIdx.x % tiles_n;
    }
    const int m0 = tile_m * BM, n0 = tile_n * BN;
    if (m0 >= M || n0 >= N) return;               // whole CTA leaves together (2:1 lower rasterisation overshoot)
    // lower_only == 2: A and B are upper triangular (U U^T): every k < m0 contributes zero to this tile
    const int k_begin = (lower_only == 2) ? (m0 / BK) * BK : 0;
    const int kblocks = (K - k_begin + BK - 1) / BK;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bar_base + 8 * s, 1);
            mbar_init(bar_base + 8 * (STAGES + s), CONSUMER_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // ===== TMA producer duty: thread 0 keeps up to STAGES loads in flight from inside the consumer loop.
    // (A dedicated 9th producer warp would put 3 warps on one SM sub-partition and cap every thread
    // at 168 registers — the register file is 16K per sub-partition — which spills the 128-register
    // accumulator tile; with 8 warps the cap is 255.)
    auto issue_stage = [&](int kb) {
        const int s = kb % STAGES;
        const uint32_t full = bar_base + 8 * s;
        mbar_expect_tx(full, STAGE_BYTES);
        const uint32_t dstA = smem_base + s * STAGE_BYTES;
        tma_load_2d(dstA, &mapA, k_begin + kb * BK, a_row0 + m0, full);
        tma_load_2d(dstA + A_STAGE_BYTES, &mapB, k_begin + kb * BK, a_row0 + n0, full);
    };
    int next_kb = kblocks < STAGES ? kblocks : STAGES;             // first k-block not yet requested (thread 0)
    if (threadIdx.x == 0) {
        for (int kb = 0; kb < next_kb; ++kb) issue_stage(kb);      // every stage starts out free
    }

    // ===== consumers =====
    const int wm = warp / CF::WN, wn = warp % CF::WN;
    const int g = lane >> 2, t = lane & 3;
    const int pg = perm8(g);

    // C is folded into the accumulators up front (acc = (beta/alpha) C): the loads overlap the TMA
    // pipeline fill and the epilogue becomes store-only (a read-modify-write epilogue cost ~8 serial
    // DRAM round trips per tile, 14% of a K=512 tile).
    const int c0 = perm8(2 * t), c1 = perm8(2 * t + 1);
    double acc[MI][NJ][2];
    if (beta != 0.0) {
        // All MI*NJ*2 loads are issued back to back (volatile asm keeps them ahead of the scaling, which
        // depends on a value produced after the last load), so the tile costs one DRAM round trip
        // instead of one per row group.  Masked elements read C[0] (always valid) and are zeroed.
#pragma unroll
        for (int i = 0; i < MI; ++i) {
            const int row = m0 + wm * WARP_M + 8 * i + pg;
            const double* crow = C + (int64_t)row * ldc;
            const int col_lim = row < M ? (lower_only ? min(N, row + 1) : N) : 0;
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const int cb = n0 + wn * WARP_N + 8 * j;
                const double* p0 = (cb + c0 < col_lim) ? crow + cb + c0 : C;
                const double* p1 = (cb + c1 < col_lim) ? crow + cb + c1 : C;
                asm volatile("ld.global.f64 %0, [%1];" : "=d"(acc[i][j][0]) : "l"(p0));
                asm volatile("ld.global.f64 %0, [%1];" : "=d"(acc[i][j][1]) : "l"(p1));
            }
        }
        double cs;
        asm volatile("mov.f64 %0, %1;" : "=d"(cs) : "d"(beta / alpha));
#pragma unroll
        for (int i = 0; i < MI; ++i) {
            const int row = m0 + wm * WARP_M + 8 * i + pg;
            const int col_lim = row < M ? (lower_only ? min(N, row + 1) : N) : 0;
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const int cb = n0 + wn * WARP_N + 8 * j;
                acc[i][j][0] = (cb + c0 < col_lim) ? cs * acc[i][j][0] : 0.0;
                acc[i][j][1] = (cb + c1 < col_lim) ? cs * acc[i][j][1] : 0.0;
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    }

    // byte offset inside a tile of (row = base + 8*i + pg, k = 4*ks + t) under SWIZZLE_128B
    uint32_t koff[4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) koff[ks] = ((((2 * ks) | (t >> 1)) ^ pg) << 4) | ((t & 1) << 3);
    const uint32_t a_row = (wm * WARP_M + pg) * 128;
    const uint32_t b_row = A_STAGE_BYTES + (wn * WARP_N + pg) * 128;

    for (int kb = 0; kb < kblocks; ++kb) {
        const int s = kb % STAGES;
        mbar_wait(bar_base + 8 * s, (kb / STAGES) & 1);
        // Release the stage consumed in iteration kb-1 only now, behind the wait loop above.  The wait
        // is real control flow, so every DMMA of iteration kb-1 (and therefore every LDS feeding it)
        // has issued before this arrive.  Releasing at the end of iteration kb-1 is NOT safe: ptxas
        // hoists the arrive above the last DMMAs, right behind the last LDS *issue*, and the TMA
        // refill then overwrites shared memory an in-flight LDS has not read yet (observed as sporadic
        // 32-byte fragment corruption on multi-wave grids).
        if (kb >= 1) {
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_base + 8 * (STAGES + ((kb - 1) % STAGES)));
            if (threadIdx.x == 0) {
                // Refill every stage whose release has completed.  Do not stall this warp on a slow
                // sibling unless the k-block is needed by the very next iteration.
                while (next_kb < kblocks && next_kb <= kb - 1 + STAGES) {
                    const uint32_t eb = bar_base + 8 * (STAGES + (next_kb % STAGES));
                    const uint32_t par = ((next_kb / STAGES) - 1) & 1;
                    if (!mbar_test(eb, par)) {
                        if (next_kb > kb + 1) break;
                        mbar_wait(eb, par);
                    }
                    issue_stage(next_kb);
                    ++next_kb;
                }
            }
            __syncwarp();
        }
        const uint32_t st = smem_base + s * STAGE_BYTES;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            double a[MI], b[NJ];
#pragma unroll
            for (int i = 0; i < MI; ++i) a[i] = lds64(st + a_row + i * 1024 + koff[ks]);
#pragma unroll
            for (int j = 0; j < NJ; ++j) b[j] = lds64(st + b_row + j * 1024 + koff[ks]);
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NJ; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }

    // ===== epilogue: C = alpha * acc (store-only; the row/column permutation is undone here) =====
#pragma unroll
    for (int i = 0; i < MI; ++i) {
        const int row = m0 + wm * WARP_M + 8 * i + pg;
        if (row >= M) continue;
        double* crow = C + (int64_t)row * ldc;
        const int col_lim = lower_only ? min(N, row + 1) : N;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int cb = n0 + wn * WARP_N + 8 * j;
            if (cb + c0 < col_lim) crow[cb + c0] = alpha * acc[i][j][0];
            if (cb + c1 < col_lim) crow[cb + c1] = alpha * acc[i][j][1];
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

template <class CF>
int launch(cudaStream_t stream, int64_t M, int64_t N, int64_t K, double alpha, const double* A, int64_t lda,
           const double* B, int64_t ldb, double beta, double* C, int64_t ldc, int lower_only) {
    static PerDeviceOnce configured;
    if (configured.first()) {
        PB_CUDA(cudaFuncSetAttribute(gemm_nt_kernel<CF>, cudaFuncAttributeMaxDynamicSharedMemorySize, CF::SMEM_BYTES));
        PB_CUDA(cudaFuncSetAttribute(gemm_nt_kernel<CF>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                     cudaSharedmemCarveoutMaxShared));
    }
    CUtensorMap mapA, mapB;
    PB_TRY(make_map(&mapA, A, M, K, lda, CF::BM));
    PB_TRY(make_map(&mapB, B, N, K, ldb, CF::BN));
    const int tm = (int)ceil_div<int64_t>(M, CF::BM), tn = (int)ceil_div<int64_t>(N, CF::BN);
    int64_t blocks;
    static_assert(CF::BM == CF::BN || CF::BM == 2 * CF::BN, "lower-only rasterisation supports 1:1 and 2:1 tiles");
    if (lower_only && CF::BM != CF::BN) {
        blocks = (int64_t)tm * (tm + 1);          // may include column tiles past N in the last row: they exit early
    } else if (lower_only) {
        blocks = 0;
        for (int c0 = 0; c0 < tn; c0 += STRIP_W) {
            int w = tn - c0 < STRIP_W ? tn - c0 : STRIP_W;
            blocks += (int64_t)w * (w + 1) / 2 + (int64_t)(tn - c0 - w) * w;
        }
    } else {
        blocks = (int64_t)tm * tn;
    }
    PB_CHECK(blocks < (1ll << 31), PB_ERR_INVALID, "gemm_nt: too many tiles");
    bool prof = profiling_enabled() && CF::BM == 128;
    const double algorithmic = lower_only ? (double)N * (double)(N + 1) * (double)K : 2.0 * M * (double)N * (double)K;
    if (CaptureTally* tally = capture_tally()) {      // being captured into a graph: no events, tally the flops
        if (CF::BM == 128) { tally->flops += algorithmic; tally->launches += 1; }
        prof = false;
    }
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (prof) {
        PB_CUDA(cudaEventCreate(&e0));
        PB_CUDA(cudaEventCreate(&e1));
        PB_CUDA(cudaEventRecord(e0, stream));
    }
    gemm_nt_kernel<CF><<<(unsigned)blocks, CF::THREADS, CF::SMEM_BYTES, stream>>>(
        mapA, mapB, C, ldc, (int)M, (int)N, (int)K, alpha, beta, lower_only, tn, GemmGroups{}); pb::note_launch();
    if (prof) {
        PB_CUDA(cudaEventRecord(e1, stream));
        // algorithmic flops: 2MNK, or the lower triangle N(N+1)K for the SYRK form
        profile_gemm(e0, e1, algorithmic);
    }
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

template <class CF>
int launch_groups(cudaStream_t stream, const double* P, int64_t ldp, int64_t K, double alpha, double beta, double* c0,
                  int64_t ldc, const GemmGroups& g, double flops) {
    static PerDeviceOnce configured;
    if (configured.first()) {
        PB_CUDA(cudaFuncSetAttribute(gemm_nt_kernel<CF>, cudaFuncAttributeMaxDynamicSharedMemorySize, CF::SMEM_BYTES));
        PB_CUDA(cudaFuncSetAttribute(gemm_nt_kernel<CF>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                     cudaSharedmemCarveoutMaxShared));
    }
    CUtensorMap mapA, mapB;
    PB_TRY(make_map(&mapA, P, g.R, K, ldp, CF::BM));
    PB_TRY(make_map(&mapB, P, g.R, K, ldp, CF::BN));
    const int tn = (int)ceil_div<int64_t>(g.nb, CF::BN);
    int64_t blocks = 0;
    for (int q = 0; q < g.count; ++q) blocks += ceil_div<int64_t>(g.R - (g.a0 + (int64_t)q * g.astep), CF::BM) * tn;
    PB_CHECK(blocks > 0 && blocks < (1ll << 31), PB_ERR_INVALID, "gemm_nt_groups: bad tile count");
    const bool prof = profiling_enabled() && CF::BM == 128;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (prof) {
        PB_CUDA(cudaEventCreate(&e0));
        PB_CUDA(cudaEventCreate(&e1));
        PB_CUDA(cudaEventRecord(e0, stream));
    }
    gemm_nt_kernel<CF><<<(unsigned)blocks, CF::THREADS, CF::SMEM_BYTES, stream>>>(
        mapA, mapB, c0, ldc, 0, 0, (int)K, alpha, beta, 3, tn, g); pb::note_launch();
    if (prof) {
        PB_CUDA(cudaEventRecord(e1, stream));
        profile_gemm(e0, e1, flops);
    }
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

}  // namespace

// Grouped trailing update (see GemmGroups): for q < count,
//   C_q[M_q x N_q] = alpha * P[a0 + q astep :, :K] P[a0 + q astep : +N_q, :K]^T + beta * C_q,  C_q = c0 + q cstep.
int gemm_nt_groups(cudaStream_t stream, const double* P, int64_t ldp, int64_t R, int64_t K, double alpha, double beta,
                   double* c0, int64_t ldc, int64_t cstep, int count, int64_t a0, int64_t astep, int64_t nb) {
    if (count <= 0) return PB_OK;
    PB_CHECK(K > 0 && alpha != 0.0 && nb > 0 && R < (1ll << 31) && a0 + (count - 1) * astep < R, PB_ERR_INVALID,
             "gemm_nt_groups: bad arguments");
    GemmGroups g;
    g.count = count; g.R = (int)R; g.a0 = (int)a0; g.astep = (int)astep; g.nb = (int)nb; g.cstep = cstep;
    double flops = 0;
    int64_t tiles = 0;
    for (int q = 0; q < count; ++q) {
        const int64_t Mq = R - (a0 + q * astep), Nq = std::min<int64_t>(nb, Mq);
        flops += 2.0 * Mq * (double)Nq * (double)K;
        tiles += ceil_div<int64_t>(Mq, 128) * ceil_div<int64_t>(nb, 64);
    }
    if (tiles < 2 * num_sms()) return launch_groups<CfgSmall>(stream, P, ldp, K, alpha, beta, c0, ldc, g, flops);
    return launch_groups<CfgMain>(stream, P, ldp, K, alpha, beta, c0, ldc, g, flops);
}

int gemm_nt(cudaStream_t stream, int64_t M, int64_t N, int64_t K, double alpha, const double* A, int64_t lda,
            const double* B, int64_t ldb, double beta, double* C, int64_t ldc, bool lower_only) {
    return gemm_nt_mode(stream, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, lower_only ? 1 : 0);
}

// mode 0: full; 1: lower tiles only; 2: lower tiles only AND both operands upper triangular (C = U U^T:
// k-blocks left of the tile's first row are skipped, N^3/3 instead of N^3 flops).
int gemm_nt_mode(cudaStream_t stream, int64_t M, int64_t N, int64_t K, double alpha, const double* A, int64_t lda,
                 const double* B, int64_t ldb, double beta, double* C, int64_t ldc, int lower_only) {
    if (M <= 0 || N <= 0) return PB_OK;
    PB_CHECK(K > 0, PB_ERR_INVALID, "gemm_nt: K must be positive");
    PB_CHECK(alpha != 0.0, PB_ERR_INVALID, "gemm_nt: alpha must be non-zero");
    PB_CHECK(!lower_only || M == N, PB_ERR_INVALID, "gemm_nt: lower_only needs a square C");
    PB_CHECK(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), PB_ERR_INVALID, "gemm_nt: dimension too large");
    // Small problems (panel factorisation, leaf TRSMs) would occupy only a handful of SMs with
    // 128x64 tiles: give them 64x64 tiles so that twice as many CTAs share the work.
    const int64_t main_tiles = ceil_div<int64_t>(M, 128) * ceil_div<int64_t>(N, 64) / (lower_only ? 2 : 1);
    if (N <= 64 || M <= 64 || main_tiles < 2 * num_sms())
        return launch<CfgSmall>(stream, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, lower_only);
    return launch<CfgMain>(stream, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, lower_only);
}

}  // namespace pb

extern "C" int pb_gemm_nt(pb_stream_t stream, int64_t M, int64_t N, int64_t K, double alpha, const double* A,
                          int64_t lda, const double* B, int64_t ldb, double beta, double* C, int64_t ldc,
                          int32_t lower_only) {
    return pb::gemm_nt(reinterpret_cast<cudaStream_t>(stream), M, N, K, alpha, A, lda, B, ldb, beta, C, ldc,
                       lower_only != 0);
}
