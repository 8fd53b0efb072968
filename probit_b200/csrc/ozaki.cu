// FP64 contraction on the 5th-generation INT8 tensor cores (tcgen05.mma.kind::i8, TMEM accumulators, TMA-fed):
//     C[M x N] += alpha * A[M x K] * B[N x K]^T        (row-major "NT", optionally only the lower tiles of a square C)
// by error-free slicing (Ozaki scheme).  This is the trailing update of the Cholesky factorisation
// (probit/implicit/Laplace.py:24 -> potrf.cu) past the 37 TFLOP/s wall of the FP64 DMMA pipe: tcgen05 has no f64 kind,
// but the int8 kind runs at 4.5 POPS and its int32 accumulation is exact.
//
//   slicing   every row of an operand gets one exponent e (max |x| 2^-e in [1/4, 1/2)); the scaled row is cut into S = 7
//             signed base-128 digits d_1 .. d_S in [-64, 64] by repeated round-to-nearest (each step exact in FP64):
//             x = 2^e (sum_p d_p 128^-p + r 128^-S), |r| <= 1/2, i.e. 49 bits below the row's leading bit (2^-49 .. 2^-48 of
//             its largest entry; CPU model of the whole scheme: tests/test_int8_slicing_model.py).
//   products  A B^T = 2^(ea_i + eb_j) sum_{p,q} 128^-(p+q) (A_p B_q^T); every A_p B_q^T is an exact int8 x int8 -> int32
//             GEMM.  Pairs are grouped by level t = p + q; levels t <= S + 1 are kept (28 products for S = 7, the
//             dropped ones are below 128^-(S+2)), and the products of one level share one int32 TMEM accumulator.
//   kernel    128 x 64 tiles of C, runs of two tiles per CTA: warp 0 = TMA producer (3-D box: 64 K-bytes x rows x 7 planes,
//             SWIZZLE_64B, two 86 KB stages), warp 1 = one elected lane issuing 56 UTCIMMA per stage into 7 accumulators
//             (448 TMEM columns), warps 2..5 = epilogue (tcgen05.ld, exact int32 -> f64, level scaling by exact powers of
//             two, row / column exponents, C += by asynchronous 256-byte bulk reductions).  Variants behind
//             pb_options.ozaki_tile: 128 x 128 tiles in two passes over the levels, and clusters of two CTAs sharing the A
//             tile by TMA multicast — both correct, neither faster.
// Accuracy: the slicing error is at most 2^-48 of each ROW's largest entry and the int32 sums are exact; a K = 1024 update of
// O(1) entries is perturbed by ~1e-14 (DMMA: ~3e-15).  tests/test_gpu_kernels.py pins it against FP64 products and the
// factorisation built on it against LAPACK.
#include "common.cuh"
#include <algorithm>
#include <cstdlib>

namespace pb {

namespace {

constexpr int OZ_S = 7;                                  // slices per value; levels t = 2 .. S + 1
constexpr int OZ_BM = 128, OZ_BN = 64, OZ_KC = 64;       // CTA tile; K bytes per stage (= one SWIZZLE_64B span)
constexpr int OZ_STAGES = 2;
constexpr int OZ_A_PLANE = OZ_BM * OZ_KC, OZ_B_PLANE = OZ_BN * OZ_KC;
constexpr int OZ_A_STAGE = OZ_S * OZ_A_PLANE, OZ_B_STAGE = OZ_S * OZ_B_PLANE;
constexpr int OZ_STAGE = OZ_A_STAGE + OZ_B_STAGE;        // 86016 B
constexpr int OZ_TP = OZ_BN / 2 + 2;                     // pitch of the epilogue's staging half-tile (doubles): rows 16-byte aligned
                                                         // for the bulk reduction, 4-way bank conflicts on the row-wise stores
constexpr int OZ_T_BYTES = (OZ_BM * OZ_TP * 8 + 127) / 128 * 128;
constexpr int OZ_SMEM = OZ_STAGES * OZ_STAGE + OZ_T_BYTES + 1024 /* align slack */ + 64 /* barriers */;
static_assert(OZ_SMEM <= 227 * 1024, "shared memory budget");
constexpr int OZ_THREADS = 192;                          // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue
constexpr int OZ_TMEM_COLS = 512;
constexpr int OZ_TILES_PER_CTA = 2;
// instruction descriptor (cute::UMMA::InstrDescriptor): D = S32 (2 << 4), A = B = signed int8 (1 << 7, 1 << 10), both
// K-major, N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t OZ_IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(OZ_BN >> 3) << 17) | ((uint32_t)(OZ_BM >> 4) << 24);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "OZ_WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra OZ_WAIT_DONE;\n"
        "bra OZ_WAIT_LOOP;\n"
        "OZ_WAIT_DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
        : "memory");
}
// K-major operand tile, rows at a 64-byte pitch, SWIZZLE_64B (cute::UMMA::SmemDescriptor): start address >> 4,
// leading byte offset 1 (unused for swizzled K-major), stride byte offset = 8 rows x 64 B = 512 B >> 4, version 1,
// layout type SWIZZLE_64B = 4.
// The two operand descriptors differ only in their low words (start address >> 4 | LBO << 16), and every plane / K-step
// offset stays inside the 14-bit address field, so the issuing thread works on 32-bit low words with immediate offsets:
// 3-4 instructions per UTCIMMA.  (The first version rebuilt both 64-bit descriptors in rolled loops: ~170 cycles of
// dependent scalar code per MMA on one thread, five times the MMA's own 32 cycles — the tensor pipe sat idle.)
__device__ __forceinline__ uint32_t oz_desc_lo(uint32_t addr) { return ((addr >> 4) & 0x3fffu) | (1u << 16); }
constexpr uint32_t OZ_DESC_HI = 32u | (1u << 14) | (4u << 29);   // SBO = 512 B >> 4, version 1, SWIZZLE_64B
__device__ __forceinline__ void oz_mma(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .b64 da, db;\n"
        "mov.b64 da, {%1, %3};\n"
        "mov.b64 db, {%2, %3};\n"
        "setp.ne.b32 p, %5, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], da, db, %4, p;\n"
        "}\n" ::"r"(tmem_d),
        "r"(a_lo), "r"(b_lo), "r"(OZ_DESC_HI), "r"(OZ_IDESC), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// as tma_load_3d, delivered at the same shared-memory offset (and completing the same-offset barrier) in every CTA of `mask`
__device__ __forceinline__ void tma_load_3d_multicast(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar,
                                                      uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, %3, %4}], [%5], %6;"
        ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void oz_commit_multicast(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
                 : "memory");
}
__device__ __forceinline__ bool oz_elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void oz_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ double pow2i(int e) {          // 2^e for -1022 <= e <= 1023; anything else -> NaN (bad row)
    if (e < -1022 || e > 1023) return __longlong_as_double(0x7ff8000000000000ll);
    return __hiloint2double((e + 1023) << 20, 0);
}
__device__ __forceinline__ void lower_tile_2to1(int bid, int& tm, int& tn) {   // tile row tm has column tiles 0 .. 2 tm + 1
    int r = (int)((sqrtf(4.0f * bid + 1.0f) - 1.0f) * 0.5f);
    while ((r + 1) * (r + 2) <= bid) ++r;
    while (r * (r + 1) > bid) --r;
    tm = r;
    tn = bid - r * (r + 1);
}

// One warp per row: exponent + OZ_S digit planes.  planes[p][row][k] (K contiguous), expo[row].
__global__ void __launch_bounds__(256)
oz_slice_kernel(const double* __restrict__ P, int64_t rows, int K, int64_t ld, int8_t* __restrict__ planes,
                int64_t plane_stride, int32_t* __restrict__ expo) {
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const double* p = P + row * ld;
    double mx = 0.0;
    bool bad = false;
    for (int k = lane * 2; k < K; k += 64) {
        const double2 v = *reinterpret_cast<const double2*>(p + k);
        mx = fmax(mx, fmax(fabs(v.x), fabs(v.y)));
        bad |= !(fabs(v.x) <= 1.7976931348623157e308) || !(fabs(v.y) <= 1.7976931348623157e308);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    bad = __any_sync(0xffffffffu, bad);
    int e = 0;
    if (mx > 0.0) {
        int q;
        frexp(mx, &q);           // mx = m 2^q, m in [1/2, 1)
        e = q + 1;               // |x| 2^-e < 1/2
    }
    // x 2^(7 - e) in (-64, 64); e is clamped so that the scale is a normal number (rows of denormals lose nothing that matters)
    e = max(e, -900);
    const double scale = pow2i(7 - e);
    int8_t* out = planes + row * K;
    for (int k = lane * 4; k < K; k += 128) {
        const double2 v0 = *reinterpret_cast<const double2*>(p + k);
        const double2 v1 = *reinterpret_cast<const double2*>(p + k + 2);
        double t[4] = {v0.x * scale, v0.y * scale, v1.x * scale, v1.y * scale};
#pragma unroll
        for (int s = 0; s < OZ_S; ++s) {
            uint32_t word = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const double d = rint(t[i]);
                t[i] = (t[i] - d) * 128.0;                    // exact: |t - d| <= 1/2
                word |= ((uint32_t)(__double2int_rn(d)) & 0xffu) << (8 * i);
            }
            *reinterpret_cast<uint32_t*>(out + (int64_t)s * plane_stride + k) = word;
        }
    }
    if (lane == 0) expo[row] = bad ? 0x7fffffff : e;         // NaN / Inf in the row: poison the outputs it touches
}

// Debug timeline (PB_OZ_TIMING=<block index>): clock64 of one CTA at the stations below, read back by pb_debug_oz_times.
__device__ unsigned long long oz_dbg[16];
enum { DBG_ENTRY = 0, DBG_SETUP, DBG_FIRST_FULL, DBG_LAST_MMA, DBG_TFULL, DBG_DRAINED, DBG_C_DONE, DBG_EXIT, DBG_FIRST_TMA };

// Each CTA walks a run of up to OZ_TILES_PER_CTA (2) consecutive tiles (consecutive tiles share their row tile, i.e. the A
// planes in L2), then retires: long enough to amortise the prologue, short enough that the SM is handed back every
// ~100 us — the Cholesky look-ahead runs its panel work on a high-priority stream UNDER this kernel and needs SMs to
// free up (a fully persistent grid starved it).  The three roles run as independent pipelines across tile boundaries: the
// producer keeps the TMA ring full
// into the next tile while the epilogue is still draining the previous one, the MMA lane starts the next tile as soon as
// the epilogue has read the accumulators out of TMEM (tmem_empty), and the read-modify-write of C — the part that waits
// on global memory — overlaps the next tile's MMAs.  (One tile per CTA paid ~17 k cycles of prologue, pipeline fill and
// serial epilogue per tile: 25 % of a K = 1024 tile, 40 % of a K = 512 one.)
// CL: clusters of two CTAs work on tiles (tm, 2i) and (tm, 2i + 1), which share their 128 rows of A.  Each CTA fetches
// its own B tile and HALF of the A tile (64 rows, plane by plane) and multicasts that half into both CTAs' shared memory,
// so a chunk costs 57 KB of L2 -> SM traffic per CTA instead of 86 KB.  A stage may be refilled only when BOTH CTAs have
// consumed it: the MMA lane's commit arrives on the empty barrier of both CTAs.  MEASURED (profiles/r02_oz_cta_timeline.txt):
// the mainloop stays at 2790 cycles per 56-MMA chunk with or without the multicast — the L2 -> SM feed was NOT what set
// the pace (the M128 N64 K32 int8 MMA itself takes ~50 cycles) — so this variant is opt-in (pb_options.ozaki_tile = 2).
template <bool CL>
__global__ void __launch_bounds__(OZ_THREADS, 1)
oz_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
               const __grid_constant__ CUtensorMap mapA1, const int32_t* __restrict__ ea, const int32_t* __restrict__ eb,
               double* __restrict__ C, int64_t ldc, int M, int N, int K, double alpha, int lower_only, int tiles, int tn_count,
               int tiles_per_cta, int red) {
    extern __shared__ uint8_t oz_raw[];
    const uint32_t base = (smem_u32(oz_raw) + 1023u) & ~1023u;
    const uint32_t stage_t = base + OZ_STAGES * OZ_STAGE;      // epilogue staging: T[128][OZ_TP] doubles (half a tile)
    const uint32_t bars = stage_t + OZ_T_BYTES;                // full[0..1], empty[0..1], tmem_full, tmem_empty
    const uint32_t bar_tfull = bars + 8 * 2 * OZ_STAGES, bar_tempty = bar_tfull + 8;
    __shared__ uint32_t tmem_slot;
    __shared__ double colscale[2][OZ_BN];                     // by tile parity: the bulk-reduction path has no barrier between a
                                                             // tile's last read of its scales and the next tile's write
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nk = K / OZ_KC;
    // slots: tiles of this CTA (CL: slot i of cluster c is tile 2 i + rank; `tiles` is even and no tile is skipped)
    const uint32_t crank = CL ? cluster_ctarank() : 0u;
    const int slots = CL ? tiles / 2 : tiles;
    const int owner = CL ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int t_begin = owner * tiles_per_cta, t_end = min(slots, t_begin + tiles_per_cta);
    const bool dbg = (red >> 8) == (int)blockIdx.x + 1;
    red &= 1;
    if (dbg && threadIdx.x == 0) oz_dbg[DBG_ENTRY] = clock64();

    auto decode = [&](int slot, int& m0, int& n0) -> bool {    // false: the tile lies outside the matrix (ragged last row tile)
        const int t = CL ? 2 * slot + (int)crank : slot;
        int tm, tn;
        if (lower_only) lower_tile_2to1(t, tm, tn);
        else { tm = t / tn_count; tn = t - tm * tn_count; }
        m0 = tm * OZ_BM;
        n0 = tn * OZ_BN;
        return CL || (m0 < M && n0 < N);                        // CL: both CTAs of a pair always walk in lockstep; TMA zero-fills
    };

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < OZ_STAGES; ++s) {
            mbar_init(bars + 8 * s, 1);
            mbar_init(bars + 8 * (OZ_STAGES + s), CL ? 2 : 1);
        }
        mbar_init(bar_tfull, 1);
        mbar_init(bar_tempty, 128);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(OZ_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CL) cluster_sync();                                    // the peer's barriers exist before anything is multicast to them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (dbg && threadIdx.x == 0) oz_dbg[DBG_SETUP] = clock64();

    if (warp == 0) {
        if (lane == 0) {
            int c = 0;                                          // chunk counter across tiles: stage = c % 2
            for (int t = t_begin; t < t_end; ++t) {
                int m0, n0;
                if (!decode(t, m0, n0)) continue;
                for (int kc = 0; kc < nk; ++kc, ++c) {
                    const int s = c % OZ_STAGES;
                    if (c >= OZ_STAGES) mbar_wait(bars + 8 * (OZ_STAGES + s), ((c / OZ_STAGES) - 1) & 1);
                    const uint32_t full = bars + 8 * s;
                    mbar_expect_tx(full, OZ_STAGE);
                    if (CL) {
                        // my half of the A tile, one plane per instruction (the stage keeps [plane][128 rows][64 B]), to both CTAs
#pragma unroll
                        for (int pl = 0; pl < OZ_S; ++pl)
                            tma_load_3d_multicast(base + s * OZ_STAGE + pl * OZ_A_PLANE + crank * (OZ_A_PLANE / 2), &mapA1, kc * OZ_KC,
                                                  m0 + 64 * (int)crank, pl, full, (uint16_t)3);
                    } else {
                        tma_load_3d(base + s * OZ_STAGE, &mapA, kc * OZ_KC, m0, 0, full);
                    }
                    tma_load_3d(base + s * OZ_STAGE + OZ_A_STAGE, &mapB, kc * OZ_KC, n0, 0, full);
                    if (dbg && c == 0) oz_dbg[DBG_FIRST_TMA] = clock64();
                }
            }
        }
    } else if (warp == 1) {
        // one ELECTED lane issues: with `if (lane == 0)` around it ptxas wraps every UTCIMMA in its own elect-and-branch
        // loop (9 instructions and a branch per MMA)
        if (oz_elect_one()) {
            int c = 0, j = 0;                                   // chunk counter, tile counter of this CTA
            for (int t = t_begin; t < t_end; ++t) {
                int m0, n0;
                if (!decode(t, m0, n0)) continue;
                if (j > 0) {                                    // the epilogue must have read tile j - 1 out of TMEM
                    mbar_wait(bar_tempty, (j - 1) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                for (int kc = 0; kc < nk; ++kc, ++c) {
                    const int s = c % OZ_STAGES;
                    mbar_wait(bars + 8 * s, (c / OZ_STAGES) & 1);
                    if (dbg && c == 0) oz_dbg[DBG_FIRST_FULL] = clock64();
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_lo = oz_desc_lo(base + s * OZ_STAGE), b_lo = oz_desc_lo(base + s * OZ_STAGE + OZ_A_STAGE);
                    const uint32_t first = kc > 0 ? 1u : 0u;    // the very first MMA of a level overwrites its accumulator
                    // Round-robin over the levels: consecutive MMAs go to DIFFERENT accumulators (level lt has 2 (lt - 1)
                    // MMAs per stage; issued level by level, up to 14 back-to-back MMAs would chain on one accumulator).
#pragma unroll
                    for (int i = 0; i < 2 * OZ_S; ++i) {
#pragma unroll
                        for (int lt = 2; lt <= OZ_S + 1; ++lt) {    // level lt = p + q -> accumulator lt - 2
                            if (i >= 2 * (lt - 1)) continue;
                            const int p = i / 2 + 1, q = lt - p, ks = i & 1;   // one UTCIMMA = 32 bytes of K: advance the start address
                            oz_mma(tmem + (uint32_t)(lt - 2) * OZ_BN, a_lo + (uint32_t)((p - 1) * (OZ_A_PLANE >> 4) + 2 * ks),
                                   b_lo + (uint32_t)((q - 1) * (OZ_B_PLANE >> 4) + 2 * ks), i == 0 ? first : 1u);
                        }
                    }
                    // frees the stage once these MMAs have read it (CL: in both CTAs of the pair, whose loads land in both)
                    if (CL) oz_commit_multicast(bars + 8 * (OZ_STAGES + s), (uint16_t)3);
                    else oz_commit(bars + 8 * (OZ_STAGES + s));
                }
                oz_commit(bar_tfull);                           // accumulators of this tile complete
                if (dbg && j == 0) oz_dbg[DBG_LAST_MMA] = clock64();
                ++j;
            }
        }
    } else {
        const int quarter = warp & 3;                           // TMEM lanes this warp may touch: 32 (warp % 4) ..
        const int trow = quarter * 32 + lane;                   // row of the tile this thread reads out of TMEM
        const int et = threadIdx.x - 64;                        // 0 .. 127
        double* T = reinterpret_cast<double*>(oz_raw + (stage_t - smem_u32(oz_raw)));
        const double MAGIC = 4503601774854144.0;                // 2^52 + 2^31
        auto prefetch_tile = [&](int m0, int n0) {              // this thread's 512 bytes of C towards L2 while the MMAs run
            const int64_t row = m0 + trow;
            if (row < M) {
                const double* crow = C + row * ldc + n0;
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (n0 + 16 * i < N) asm volatile("prefetch.global.L2 [%0];" ::"l"(crow + 16 * i));
            }
        };
        int j = 0;
        {
            int m0, n0;
            for (int t = t_begin; t < t_end; ++t)
                if (decode(t, m0, n0)) { prefetch_tile(m0, n0); break; }
        }
        for (int t = t_begin; t < t_end; ++t) {
            int m0, n0;
            if (!decode(t, m0, n0)) continue;
            const int64_t row = m0 + trow;
            const double rs = row < M ? alpha * pow2i(ea[row]) : 0.0;
            double* cscale = colscale[j & 1];
            if (et < OZ_BN) cscale[et] = (n0 + et < N) ? pow2i(eb[n0 + et]) : 0.0;
            mbar_wait(bar_tfull, j & 1);
            if (dbg && j == 0 && et == 0) oz_dbg[DBG_TFULL] = clock64();
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            asm volatile("bar.sync 1, 128;" ::: "memory");      // colscale visible; T free (previous tile's stores done)
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                double acc[32];
#pragma unroll
                for (int jj = 0; jj < 32; ++jj) acc[jj] = 0.0;
                // int32 -> f64 without the conversion unit (I2F.F64 issues at a fraction of the DFMA rate): the bits
                // 0x43300000'(v ^ 0x80000000) are the double 2^52 + 2^31 + v, and subtracting that constant is exact.
                // Two levels per TMEM round trip, least significant first.
                auto load32 = [&](int lvl, uint32_t (&v)[32]) {
                    const uint32_t addr = tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(lvl * OZ_BN + half * 32);
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                        : "r"(addr)
                        : "memory");
                };
                auto fold32 = [&](int lvl, const uint32_t (&v)[32]) {
                    const double sc = pow2i(-7 * (lvl + 2));
#pragma unroll
                    for (int jj = 0; jj < 32; ++jj)
                        acc[jj] = fma(__hiloint2double(0x43300000, (int)(v[jj] ^ 0x80000000u)) - MAGIC, sc, acc[jj]);
                };
#pragma unroll
                for (int lvl = OZ_S - 1; lvl >= 0; lvl -= 2) {
                    uint32_t va[32], vb[32];
                    load32(lvl, va);
                    if (lvl >= 1) load32(lvl - 1, vb);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    fold32(lvl, va);
                    if (lvl >= 1) fold32(lvl - 1, vb);
                }
                if (half == 1) {                                // every accumulator has been read: the MMA lane may start the next tile
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    mbar_arrive(bar_tempty);
                    if (dbg && j == 0 && et == 0) oz_dbg[DBG_DRAINED] = clock64();
                }
                // the bulk reduction that last read this thread's row of T must have finished with it
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
#pragma unroll
                for (int jj = 0; jj < 32; ++jj) T[trow * OZ_TP + jj] = rs * cscale[half * 32 + jj] * acc[jj];
                // Interior segment (all 128 rows inside the matrix, all 32 columns inside it and, for the SYRK form, on or
                // below the diagonal for every row): C += T by ONE asynchronous bulk reduction per row, issued by the thread
                // that owns the row — the TMA unit performs the 256-byte add in L2, nobody waits for a load, no barrier.
                // (Timeline of one CTA, K = 512: the load-add-store version spent 5.8 k cycles per half here, RED.F64 4.5 k,
                // against 1.75 k for draining TMEM and 22 k for the whole mainloop.)
                const int seg0 = n0 + half * 32;
                const bool interior = m0 + OZ_BM <= M && seg0 + 32 <= N && (!lower_only || seg0 + 31 <= m0);
                if (interior && !red) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    double* dst = C + (int64_t)(m0 + trow) * ldc + seg0;
                    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], 256;"
                                 ::"l"(dst), "r"(stage_t + (uint32_t)(trow * OZ_TP * 8)) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    continue;
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                // C += T with full 256-byte row segments: thread = (row within a group of 8, column pair), 8 rows in flight
                const int cp = et & 15, r0 = et >> 4;
                const int col = n0 + half * 32 + 2 * cp;
                if (red) {
                    // fire-and-forget: one reduction per element and launch (deterministic), no round trip to wait for
#pragma unroll 4
                    for (int tr = r0; tr < OZ_BM; tr += 8) {
                        const int64_t grow = m0 + tr;
                        const int lim = grow < M ? (lower_only ? (int)min((int64_t)N, grow + 1) : N) : 0;
                        double* pc = C + grow * ldc + col;
                        if (col < lim) atomicAdd(pc, T[tr * OZ_TP + 2 * cp]);
                        if (col + 1 < lim) atomicAdd(pc + 1, T[tr * OZ_TP + 2 * cp + 1]);
                    }
                } else
#pragma unroll 1
                for (int rb = 0; rb < OZ_BM; rb += 64) {
                    double2 cc[8];
                    int valid[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int64_t grow = m0 + rb + 8 * i + r0;
                        const int lim = grow < M ? (lower_only ? (int)min((int64_t)N, grow + 1) : N) : 0;   // columns < lim are written
                        valid[i] = col + 1 < lim ? 2 : (col < lim ? 1 : 0);
                        const double* pc = C + grow * ldc + col;
                        cc[i] = make_double2(0.0, 0.0);
                        if (valid[i] == 2) cc[i] = *reinterpret_cast<const double2*>(pc);
                        else if (valid[i] == 1) cc[i].x = *pc;
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int tr = rb + 8 * i + r0;
                        double* pc = C + (int64_t)(m0 + tr) * ldc + col;
                        cc[i].x += T[tr * OZ_TP + 2 * cp];
                        cc[i].y += T[tr * OZ_TP + 2 * cp + 1];
                        if (valid[i] == 2) *reinterpret_cast<double2*>(pc) = cc[i];
                        else if (valid[i] == 1) *pc = cc[i].x;
                    }
                }
                if (half == 0) asm volatile("bar.sync 1, 128;" ::: "memory");     // T is rewritten by the second half
            }
            if (dbg && j == 0 && et == 0) oz_dbg[DBG_C_DONE] = clock64();
            ++j;
            {                                                   // C of this CTA's next tile towards L2
                int m1, n1;
                for (int t2 = t + 1; t2 < t_end; ++t2)
                    if (decode(t2, m1, n1)) { prefetch_tile(m1, n1); break; }
            }
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // this thread's reductions have left shared memory and landed
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CL) cluster_sync();                                    // no multicast write or remote arrive may target a CTA that has left
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(OZ_TMEM_COLS) : "memory");
        if (dbg && lane == 0) oz_dbg[DBG_EXIT] = clock64();
    }
}

// ------------------------------------------------------------------------------------------------ 128 x 128, two passes
// The 128 x 64 kernel above is bound by the shared-memory operand feed of the tensor core (ncu: tensor-core shared
// wavefronts 64 % with the tensor pipe 46 % active): an M128 N64 K32 int8 MMA reads 6 KB for 32 cycles of math.  With
// N = 128 the same bytes feed twice the math, but 7 level accumulators x 128 columns exceed TMEM's 512.  So the levels
// are split over TWO launches that both accumulate into C:
//     pass LO : levels 2 .. 5  = products (p, q), p + q <= 5   -> planes 1 .. 4 only, 10 products, 4 accumulators (512 cols)
//     pass HI : levels 6 .. 8  = products with 6 <= p + q <= 8 -> planes 1 .. 7,      18 products, 3 accumulators (384 cols)
// K is walked 32 bytes at a time (SWIZZLE_32B rows, one UTCIMMA per product and step), 4 / 3 stages of 32 / 56 KB.
// Same warp roles, tile runs and staged epilogue (32 columns at a time) as above.
constexpr int OZ2_BM = 128, OZ2_BN = 128, OZ2_KC = 32;
constexpr int OZ2_PLANE = 128 * OZ2_KC;                   // 4096 B, A and B alike
constexpr int OZ2_T_BYTES = OZ_T_BYTES;                   // T[128][33] doubles
constexpr uint32_t OZ2_IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(OZ2_BN >> 3) << 17) | ((uint32_t)(OZ2_BM >> 4) << 24);
constexpr uint32_t OZ2_DESC_HI = 16u | (1u << 14) | (6u << 29);     // SBO = 8 rows x 32 B = 256 B >> 4, version 1, SWIZZLE_32B
template <int NPL> struct Oz2Cfg {
    static constexpr int STAGES = NPL <= 4 ? 4 : 3;
    static constexpr int STAGE = 2 * NPL * OZ2_PLANE;
    static constexpr int SMEM = STAGES * STAGE + OZ2_T_BYTES + 1024 + 128;
};
__device__ __forceinline__ void oz2_mma(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .b64 da, db;\n"
        "mov.b64 da, {%1, %3};\n"
        "mov.b64 db, {%2, %3};\n"
        "setp.ne.b32 p, %5, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], da, db, %4, p;\n"
        "}\n" ::"r"(tmem_d),
        "r"(a_lo), "r"(b_lo), "r"(OZ2_DESC_HI), "r"(OZ2_IDESC), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tri_tile_128(int t, int& tm, int& tn) {      // t -> (tm, tn), tn <= tm
    int r = (int)((sqrtf(8.0f * t + 1.0f) - 1.0f) * 0.5f);
    while ((r + 1) * (r + 2) / 2 <= t) ++r;
    while (r * (r + 1) / 2 > t) --r;
    tm = r;
    tn = t - r * (r + 1) / 2;
}

template <int NPL, int T0, int NLEV>
__global__ void __launch_bounds__(OZ_THREADS, 1)
oz2_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                const int32_t* __restrict__ ea, const int32_t* __restrict__ eb, double* __restrict__ C, int64_t ldc, int M,
                int N, int K, double alpha, int lower_only, int tiles, int tn_count, int tiles_per_cta) {
    using CF = Oz2Cfg<NPL>;
    extern __shared__ uint8_t oz_raw[];
    const uint32_t base = (smem_u32(oz_raw) + 1023u) & ~1023u;
    const uint32_t stage_t = base + CF::STAGES * CF::STAGE;
    const uint32_t bars = stage_t + OZ2_T_BYTES;               // full[S], empty[S], tmem_full, tmem_empty
    const uint32_t bar_tfull = bars + 8 * 2 * CF::STAGES, bar_tempty = bar_tfull + 8;
    __shared__ uint32_t tmem_slot;
    __shared__ double colscale[OZ2_BN];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nk = K / OZ2_KC;
    const int t_begin = blockIdx.x * tiles_per_cta, t_end = min(tiles, t_begin + tiles_per_cta);

    auto decode = [&](int t, int& m0, int& n0) {
        int tm, tn;
        if (lower_only) tri_tile_128(t, tm, tn);
        else { tm = t / tn_count; tn = t - tm * tn_count; }
        m0 = tm * OZ2_BM;
        n0 = tn * OZ2_BN;
    };

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < CF::STAGES; ++s) {
            mbar_init(bars + 8 * s, 1);
            mbar_init(bars + 8 * (CF::STAGES + s), 1);
        }
        mbar_init(bar_tfull, 1);
        mbar_init(bar_tempty, 128);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(OZ_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int c = 0;
            for (int t = t_begin; t < t_end; ++t) {
                int m0, n0;
                decode(t, m0, n0);
                for (int kc = 0; kc < nk; ++kc, ++c) {
                    const int s = c % CF::STAGES;
                    if (c >= CF::STAGES) mbar_wait(bars + 8 * (CF::STAGES + s), ((c / CF::STAGES) - 1) & 1);
                    const uint32_t full = bars + 8 * s;
                    mbar_expect_tx(full, CF::STAGE);
                    tma_load_3d(base + s * CF::STAGE, &mapA, kc * OZ2_KC, m0, 0, full);
                    tma_load_3d(base + s * CF::STAGE + NPL * OZ2_PLANE, &mapB, kc * OZ2_KC, n0, 0, full);
                }
            }
        }
    } else if (warp == 1) {
        if (oz_elect_one()) {
            int c = 0, j = 0;
            for (int t = t_begin; t < t_end; ++t, ++j) {
                if (j > 0) {
                    mbar_wait(bar_tempty, (j - 1) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                for (int kc = 0; kc < nk; ++kc, ++c) {
                    const int s = c % CF::STAGES;
                    mbar_wait(bars + 8 * s, (c / CF::STAGES) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_lo = oz_desc_lo(base + s * CF::STAGE), b_lo = oz_desc_lo(base + s * CF::STAGE + NPL * OZ2_PLANE);
                    const uint32_t first = kc > 0 ? 1u : 0u;
#pragma unroll
                    for (int lt = T0; lt < T0 + NLEV; ++lt) {
                        const uint32_t d = tmem + (uint32_t)(lt - T0) * OZ2_BN;
                        bool fresh = true;                      // first product of this level in this step
#pragma unroll
                        for (int p = 1; p <= NPL; ++p) {
                            const int q = lt - p;
                            if (q < 1 || q > NPL) continue;
                            oz2_mma(d, a_lo + (uint32_t)((p - 1) * (OZ2_PLANE >> 4)), b_lo + (uint32_t)((q - 1) * (OZ2_PLANE >> 4)),
                                    fresh ? first : 1u);
                            fresh = false;
                        }
                    }
                    oz_commit(bars + 8 * (CF::STAGES + s));
                }
                oz_commit(bar_tfull);
            }
        }
    } else {
        const int quarter = warp & 3;
        const int trow = quarter * 32 + lane;
        const int et = threadIdx.x - 64;
        double* T = reinterpret_cast<double*>(oz_raw + (stage_t - smem_u32(oz_raw)));
        const double MAGIC = 4503601774854144.0;                // 2^52 + 2^31
        auto prefetch_tile = [&](int m0, int n0) {
            const int64_t row = m0 + trow;
            if (row < M) {
                const double* crow = C + row * ldc + n0;
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (n0 + 16 * i < N) asm volatile("prefetch.global.L2 [%0];" ::"l"(crow + 16 * i));
            }
        };
        int j = 0;
        if (t_begin < t_end) { int m0, n0; decode(t_begin, m0, n0); prefetch_tile(m0, n0); }
        for (int t = t_begin; t < t_end; ++t, ++j) {
            int m0, n0;
            decode(t, m0, n0);
            const int64_t row = m0 + trow;
            const double rs = row < M ? alpha * pow2i(ea[row]) : 0.0;
            colscale[et] = (n0 + et < N) ? pow2i(eb[n0 + et]) : 0.0;
            mbar_wait(bar_tfull, j & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll 1
            for (int qd = 0; qd < 4; ++qd) {                    // 32 columns at a time
                double acc[32];
#pragma unroll
                for (int jj = 0; jj < 32; ++jj) acc[jj] = 0.0;
                auto load32 = [&](int lvl, uint32_t (&v)[32]) {
                    const uint32_t addr = tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(lvl * OZ2_BN + qd * 32);
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                        : "r"(addr)
                        : "memory");
                };
                auto fold32 = [&](int lvl, const uint32_t (&v)[32]) {
                    const double sc = pow2i(-7 * (lvl + T0));
#pragma unroll
                    for (int jj = 0; jj < 32; ++jj)
                        acc[jj] = fma(__hiloint2double(0x43300000, (int)(v[jj] ^ 0x80000000u)) - MAGIC, sc, acc[jj]);
                };
#pragma unroll
                for (int lvl = NLEV - 1; lvl >= 0; lvl -= 2) {  // least significant level first
                    uint32_t va[32], vb[32];
                    load32(lvl, va);
                    if (lvl >= 1) load32(lvl - 1, vb);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    fold32(lvl, va);
                    if (lvl >= 1) fold32(lvl - 1, vb);
                }
                if (qd == 3) {
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    mbar_arrive(bar_tempty);
                }
#pragma unroll
                for (int jj = 0; jj < 32; ++jj) T[trow * OZ_TP + jj] = rs * colscale[qd * 32 + jj] * acc[jj];
                asm volatile("bar.sync 1, 128;" ::: "memory");
                const int cp = et & 15, r0 = et >> 4;
                const int col = n0 + qd * 32 + 2 * cp;
#pragma unroll 1
                for (int rb = 0; rb < OZ2_BM; rb += 64) {
                    double2 cc[8];
                    int valid[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int64_t grow = m0 + rb + 8 * i + r0;
                        const int lim = grow < M ? (lower_only ? (int)min((int64_t)N, grow + 1) : N) : 0;
                        valid[i] = col + 1 < lim ? 2 : (col < lim ? 1 : 0);
                        const double* pc = C + grow * ldc + col;
                        cc[i] = make_double2(0.0, 0.0);
                        if (valid[i] == 2) cc[i] = *reinterpret_cast<const double2*>(pc);
                        else if (valid[i] == 1) cc[i].x = *pc;
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int tr = rb + 8 * i + r0;
                        double* pc = C + (int64_t)(m0 + tr) * ldc + col;
                        cc[i].x += T[tr * OZ_TP + 2 * cp];
                        cc[i].y += T[tr * OZ_TP + 2 * cp + 1];
                        if (valid[i] == 2) *reinterpret_cast<double2*>(pc) = cc[i];
                        else if (valid[i] == 1) *pc = cc[i].x;
                    }
                }
                if (qd < 3) asm volatile("bar.sync 1, 128;" ::: "memory");
            }
            if (t + 1 < t_end) { int m1, n1; decode(t + 1, m1, n1); prefetch_tile(m1, n1); }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(OZ_TMEM_COLS) : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn oz_encode() {
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

// 3-D uint8 map over planes[S][rows][K]: box = 64 K-bytes x box_rows x S planes, SWIZZLE_64B, out-of-range rows read 0
int oz_map(CUtensorMap* map, const int8_t* planes, int64_t rows, int64_t K, int64_t plane_stride, int box_rows, int box_planes) {
    EncodeTiledFn enc = oz_encode();
    PB_CHECK(enc != nullptr, PB_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
    cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)OZ_S};
    cuuint64_t strides[2] = {(cuuint64_t)K, (cuuint64_t)plane_stride};
    cuuint32_t box[3] = {(cuuint32_t)OZ_KC, (cuuint32_t)box_rows, (cuuint32_t)box_planes};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<int8_t*>(planes), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PB_CHECK(r == CUDA_SUCCESS, PB_ERR_CUDA, "cuTensorMapEncodeTiled (int8 planes) failed with %d (rows=%lld K=%lld)", (int)r,
             (long long)rows, (long long)K);
    return PB_OK;
}

// 3-D uint8 map over planes[S][rows][K] for the 128 x 128 kernel: box = 32 K-bytes x 128 rows x npl planes, SWIZZLE_32B
int oz2_map(CUtensorMap* map, const int8_t* planes, int64_t rows, int64_t K, int64_t plane_stride, int npl) {
    EncodeTiledFn enc = oz_encode();
    PB_CHECK(enc != nullptr, PB_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
    cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)OZ_S};
    cuuint64_t strides[2] = {(cuuint64_t)K, (cuuint64_t)plane_stride};
    cuuint32_t box[3] = {(cuuint32_t)OZ2_KC, 128u, (cuuint32_t)npl};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<int8_t*>(planes), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PB_CHECK(r == CUDA_SUCCESS, PB_ERR_CUDA, "cuTensorMapEncodeTiled (int8 planes, 32B) failed with %d (rows=%lld K=%lld)", (int)r,
             (long long)rows, (long long)K);
    return PB_OK;
}

template <int NPL, int T0, int NLEV>
int oz2_pass(cudaStream_t st, int64_t M, int64_t N, int64_t K, double alpha, const int8_t* Ap, const int32_t* ea, int64_t a_rows,
             const int8_t* Bp, const int32_t* eb, int64_t b_rows, double* C, int64_t ldc, bool lower_only, int64_t tiles,
             int64_t tn, int64_t tpc) {
    using CF = Oz2Cfg<NPL>;
    static PerDeviceOnce configured;
    if (configured.first())
        PB_CUDA(cudaFuncSetAttribute(oz2_gemm_kernel<NPL, T0, NLEV>, cudaFuncAttributeMaxDynamicSharedMemorySize, CF::SMEM));
    CUtensorMap mapA, mapB;
    PB_TRY(oz2_map(&mapA, Ap, M, K, a_rows * K, NPL));
    PB_TRY(oz2_map(&mapB, Bp, N, K, b_rows * K, NPL));
    const dim3 grid((unsigned)ceil_div<int64_t>(tiles, tpc), 1, 1);
    oz2_gemm_kernel<NPL, T0, NLEV><<<grid, OZ_THREADS, CF::SMEM, st>>>(mapA, mapB, ea, eb, C, ldc, (int)M, (int)N, (int)K, alpha,
                                                                       lower_only ? 1 : 0, (int)tiles, (int)tn, (int)tpc);
    pb::note_launch();
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

// C += alpha A B^T with 128 x 128 tiles in two passes over the levels (see oz2_gemm_kernel)
int oz2_launch(cudaStream_t st, int64_t M, int64_t N, int64_t K, double alpha, const int8_t* Ap, const int32_t* ea,
               int64_t a_rows, const int8_t* Bp, const int32_t* eb, int64_t b_rows, double* C, int64_t ldc, bool lower_only) {
    const int64_t tm = ceil_div<int64_t>(M, OZ2_BM), tn = ceil_div<int64_t>(N, OZ2_BN);
    const int64_t tiles = lower_only ? tm * (tm + 1) / 2 : tm * tn;
    PB_CHECK(tiles < (1ll << 31), PB_ERR_INVALID, "ozaki: too many tiles");
    const int64_t tpc = K >= 1024 ? std::max<int64_t>(1, std::min<int64_t>(OZ_TILES_PER_CTA, tiles / num_sms())) : 1;
    const bool prof = profiling_enabled();
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (prof) {
        PB_CUDA(cudaEventCreate(&e0));
        PB_CUDA(cudaEventCreate(&e1));
        PB_CUDA(cudaEventRecord(e0, st));
    }
    PB_TRY((oz2_pass<4, 2, 4>(st, M, N, K, alpha, Ap, ea, a_rows, Bp, eb, b_rows, C, ldc, lower_only, tiles, tn, tpc)));
    PB_TRY((oz2_pass<7, 6, 3>(st, M, N, K, alpha, Ap, ea, a_rows, Bp, eb, b_rows, C, ldc, lower_only, tiles, tn, tpc)));
    if (prof) {
        PB_CUDA(cudaEventRecord(e1, st));
        profile_gemm(e0, e1, lower_only ? (double)N * (double)(N + 1) * (double)K : 2.0 * M * (double)N * (double)K, 2, 1);
    }
    return PB_OK;
}

// pb_options.ozaki_tile: 0 = 128 x 64 tiles, one pass over all 7 levels; 1 = 128 x 128 tiles, two passes.  The
// environment variable PB_OZ_TILE (0 / 1) overrides it for the entry points that take no options (tools, tests).
static bool oz_use_tile128() {
    static const int env = [] { const char* e = getenv("PB_OZ_TILE"); return e && *e ? atoi(e) : -1; }();
    return (env >= 0 ? env : opts().ozaki_tile) == 1;
}

static int oz_timing_block() {      // PB_OZ_TIMING=<block index>: that CTA records its timeline (0 = off, value is index + 1)
    static const int env = [] { const char* e = getenv("PB_OZ_TIMING"); return e && *e ? atoi(e) + 1 : 0; }();
    return env;
}
// PB_OZ_CLUSTER=0/1 overrides pb_options.ozaki_tile == 2 (clusters of two CTAs sharing the A tile by TMA multicast)
static bool oz_use_cluster() {
    static const int env = [] { const char* e = getenv("PB_OZ_CLUSTER"); return e && *e ? atoi(e) : -1; }();
    return (env >= 0 ? env : (opts().ozaki_tile == 2 ? 1 : 0)) == 1;
}
static bool oz_use_red() {
    static const int env = [] { const char* e = getenv("PB_OZ_RED"); return e && *e ? atoi(e) : 0; }();
    return env == 1;
}

int oz_launch(cudaStream_t st, int64_t M, int64_t N, int64_t K, double alpha, const int8_t* Ap, const int32_t* ea,
              int64_t a_rows, const int8_t* Bp, const int32_t* eb, int64_t b_rows, double* C, int64_t ldc, bool lower_only) {
    if (oz_use_tile128()) return oz2_launch(st, M, N, K, alpha, Ap, ea, a_rows, Bp, eb, b_rows, C, ldc, lower_only);
    static PerDeviceOnce configured;
    if (configured.first()) {
        PB_CUDA(cudaFuncSetAttribute(oz_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, OZ_SMEM));
        PB_CUDA(cudaFuncSetAttribute(oz_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, OZ_SMEM));
    }
    const int64_t tm = ceil_div<int64_t>(M, OZ_BM);
    int64_t tn = ceil_div<int64_t>(N, OZ_BN);
    // pairs of CTAs (clusters of 2) on column tiles 2 i, 2 i + 1 of a row tile: the lower enumeration has 2 r + 2 tiles in
    // row tile r (always even); a rectangular product pads the column tiles to an even count (the phantom tile is all masks)
    const bool cluster = oz_use_cluster() && tm * (lower_only ? tm + 1 : tn) >= 2;
    if (cluster && !lower_only) tn += tn & 1;
    CUtensorMap mapA, mapB, mapA1;
    PB_TRY(oz_map(&mapA, Ap, M, K, a_rows * K, OZ_BM, OZ_S));
    PB_TRY(oz_map(&mapB, Bp, N, K, b_rows * K, OZ_BN, OZ_S));
    PB_TRY(oz_map(&mapA1, Ap, M, K, a_rows * K, OZ_BM / 2, 1));          // half an A tile, one plane: the multicast unit
    const int64_t tiles = lower_only ? tm * (tm + 1) : tm * tn;   // lower: row tile r has column tiles 0 .. 2 r + 1 (those past N are skipped)
    PB_CHECK(tiles < (1ll << 31), PB_ERR_INVALID, "ozaki: too many tiles");
    // runs of 2 tiles: the second tile's MMAs hide the first tile's C update and half of the prologue; longer runs delay
    // the hand-back of the SM to the look-ahead stream (8-GPU trace: the panel chain is the limiter there)
    const int64_t tpc = std::max<int64_t>(1, std::min<int64_t>(OZ_TILES_PER_CTA, tiles / num_sms()));
    const bool prof = profiling_enabled();
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (prof) {
        PB_CUDA(cudaEventCreate(&e0));
        PB_CUDA(cudaEventCreate(&e1));
        PB_CUDA(cudaEventRecord(e0, st));
    }
    const int flags = (oz_use_red() ? 1 : 0) | (oz_timing_block() << 8);
    if (cluster) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(2 * ceil_div<int64_t>(tiles / 2, tpc)), 1, 1);
        cfg.blockDim = dim3(OZ_THREADS, 1, 1);
        cfg.dynamicSmemBytes = OZ_SMEM;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        PB_CUDA(cudaLaunchKernelEx(&cfg, oz_gemm_kernel<true>, mapA, mapB, mapA1, ea, eb, C, ldc, (int)M, (int)N, (int)K, alpha,
                                   lower_only ? 1 : 0, (int)tiles, (int)tn, (int)tpc, flags));
    } else {
        const dim3 grid((unsigned)ceil_div<int64_t>(tiles, tpc), 1, 1);
        oz_gemm_kernel<false><<<grid, OZ_THREADS, OZ_SMEM, st>>>(mapA, mapB, mapA1, ea, eb, C, ldc, (int)M, (int)N, (int)K, alpha,
                                                                lower_only ? 1 : 0, (int)tiles, (int)tn, (int)tpc, flags);
    }
    pb::note_launch();
    if (prof) {
        PB_CUDA(cudaEventRecord(e1, st));
        profile_gemm(e0, e1, lower_only ? (double)N * (double)(N + 1) * (double)K : 2.0 * M * (double)N * (double)K, 1, 1);
    }
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

int oz_slice(cudaStream_t st, const double* P, int64_t rows, int64_t K, int64_t ld, int8_t* planes, int32_t* expo) {
    PB_CHECK((ld & 1) == 0 && (reinterpret_cast<uintptr_t>(P) & 15) == 0, PB_ERR_INVALID, "ozaki: operand must be 16-byte aligned, ld even");
    oz_slice_kernel<<<(unsigned)ceil_div<int64_t>(rows, 8), 256, 0, st>>>(P, rows, (int)K, ld, planes, rows * K, expo); pb::note_launch();
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

inline int64_t oz_align(int64_t b) { return (b + 255) / 256 * 256; }

}  // namespace

bool ozaki_supported(int64_t K) { return K >= OZ_KC && K % OZ_KC == 0 && K <= 65536; }

int64_t ozaki_scratch_bytes(int64_t rows, int64_t K) { return oz_align(OZ_S * rows * K) + oz_align(4 * rows); }

// Slice once, multiply many times: `scratch` receives the digit planes and exponents of P (rows x K); ozaki_apply then
// forms  C[M x N] += alpha * A[a_off : a_off + M] B[b_off : b_off + N]^T  from row ranges of two sliced operands (which
// may be the same one).
// The Cholesky look-ahead uses this to share one slicing of panel k between the update of block column k + 1 (side
// stream) and the trailing SYRK (main stream).
int ozaki_slice(cudaStream_t st, const double* P, int64_t rows, int64_t K, int64_t ldp, void* scratch, int64_t scratch_bytes) {
    if (rows <= 0) return PB_OK;
    PB_CHECK(ozaki_supported(K), PB_ERR_INVALID, "ozaki: K must be a multiple of %d", OZ_KC);
    PB_CHECK(scratch && scratch_bytes >= ozaki_scratch_bytes(rows, K), PB_ERR_INVALID, "ozaki: scratch too small");
    int8_t* planes = reinterpret_cast<int8_t*>(scratch);
    int32_t* expo = reinterpret_cast<int32_t*>(reinterpret_cast<uint8_t*>(scratch) + oz_align(OZ_S * rows * K));
    return oz_slice(st, P, rows, K, ldp, planes, expo);
}

int ozaki_apply(cudaStream_t st, int64_t K, const void* a_scratch, int64_t a_rows, int64_t a_off, int64_t M,
                const void* b_scratch, int64_t b_rows, int64_t b_off, int64_t N, double alpha, double* C, int64_t ldc,
                bool lower_only) {
    if (M <= 0 || N <= 0) return PB_OK;
    PB_CHECK(a_off >= 0 && b_off >= 0 && a_off + M <= a_rows && b_off + N <= b_rows, PB_ERR_INVALID, "ozaki_apply: row range");
    PB_CHECK(!lower_only || (M == N && a_off == b_off && a_scratch == b_scratch), PB_ERR_INVALID,
             "ozaki_apply: lower_only is the SYRK form");
    PB_CHECK((ldc & 1) == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0, PB_ERR_INVALID, "ozaki: C must be 16-byte aligned, ld even");
    const int8_t* ap = reinterpret_cast<const int8_t*>(a_scratch);
    const int32_t* ae = reinterpret_cast<const int32_t*>(reinterpret_cast<const uint8_t*>(a_scratch) + oz_align(OZ_S * a_rows * K));
    const int8_t* bp = reinterpret_cast<const int8_t*>(b_scratch);
    const int32_t* be = reinterpret_cast<const int32_t*>(reinterpret_cast<const uint8_t*>(b_scratch) + oz_align(OZ_S * b_rows * K));
    return oz_launch(st, M, N, K, alpha, ap + a_off * K, ae + a_off, a_rows, bp + b_off * K, be + b_off, b_rows, C, ldc, lower_only);
}

// C (lower tiles of the n x n matrix) += alpha * P P^T, P n x K (ld ldp)
int ozaki_syrk_lower(cudaStream_t st, int64_t n, int64_t K, double alpha, const double* P, int64_t ldp, double* C,
                     int64_t ldc, void* scratch, int64_t scratch_bytes) {
    if (n <= 0) return PB_OK;
    PB_CHECK(ozaki_supported(K), PB_ERR_INVALID, "ozaki: K must be a multiple of %d", OZ_KC);
    PB_CHECK(scratch && scratch_bytes >= ozaki_scratch_bytes(n, K), PB_ERR_INVALID, "ozaki: scratch too small");
    PB_CHECK((ldc & 1) == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0, PB_ERR_INVALID, "ozaki: C must be 16-byte aligned, ld even");
    int8_t* planes = reinterpret_cast<int8_t*>(scratch);
    int32_t* expo = reinterpret_cast<int32_t*>(reinterpret_cast<uint8_t*>(scratch) + oz_align(OZ_S * n * K));
    PB_TRY(oz_slice(st, P, n, K, ldp, planes, expo));
    return oz_launch(st, n, n, K, alpha, planes, expo, n, planes, expo, n, C, ldc, true);
}

// C[M x N] += alpha * A B^T, A M x K (lda), B N x K (ldb)
int ozaki_gemm_nt(cudaStream_t st, int64_t M, int64_t N, int64_t K, double alpha, const double* A, int64_t lda,
                  const double* B, int64_t ldb, double* C, int64_t ldc, void* scratch, int64_t scratch_bytes) {
    if (M <= 0 || N <= 0) return PB_OK;
    PB_CHECK(ozaki_supported(K), PB_ERR_INVALID, "ozaki: K must be a multiple of %d", OZ_KC);
    const int64_t need = ozaki_scratch_bytes(M, K) + ozaki_scratch_bytes(N, K);
    PB_CHECK(scratch && scratch_bytes >= need, PB_ERR_INVALID, "ozaki: scratch too small");
    PB_CHECK((ldc & 1) == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0, PB_ERR_INVALID, "ozaki: C must be 16-byte aligned, ld even");
    uint8_t* s = reinterpret_cast<uint8_t*>(scratch);
    int8_t* Ap = reinterpret_cast<int8_t*>(s);
    int32_t* ea = reinterpret_cast<int32_t*>(s + oz_align(OZ_S * M * K));
    uint8_t* s2 = s + ozaki_scratch_bytes(M, K);
    int8_t* Bp = reinterpret_cast<int8_t*>(s2);
    int32_t* eb = reinterpret_cast<int32_t*>(s2 + oz_align(OZ_S * N * K));
    PB_TRY(oz_slice(st, A, M, K, lda, Ap, ea));
    PB_TRY(oz_slice(st, B, N, K, ldb, Bp, eb));
    return oz_launch(st, M, N, K, alpha, Ap, ea, M, Bp, eb, N, C, ldc, false);
}

}  // namespace pb

extern "C" int64_t pb_ozaki_scratch_bytes(int64_t M, int64_t N, int64_t K) {
    return pb::ozaki_scratch_bytes(M, K) + pb::ozaki_scratch_bytes(N, K);
}

extern "C" int pb_ozaki_gemm_nt(pb_stream_t stream, int64_t M, int64_t N, int64_t K, double alpha, const double* A,
                                int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc, int32_t lower_only,
                                void* scratch, int64_t scratch_bytes) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (lower_only) {
        PB_CHECK(M == N && A == B && lda == ldb, PB_ERR_INVALID, "ozaki: lower_only is the SYRK form (A == B, M == N)");
        return pb::ozaki_syrk_lower(st, M, K, alpha, A, lda, C, ldc, scratch, scratch_bytes);
    }
    return pb::ozaki_gemm_nt(st, M, N, K, alpha, A, lda, B, ldb, C, ldc, scratch, scratch_bytes);
}

// Debug: the timeline recorded by the CTA named in PB_OZ_TIMING during the LAST oz_gemm_kernel launch (clock64 values:
// entry, setup done, first full barrier, last MMA issued, accumulators complete, TMEM drained, C updated, exit, first TMA).
extern "C" int pb_debug_oz_times(unsigned long long* out16) {
    PB_CUDA(cudaDeviceSynchronize());
    PB_CUDA(cudaMemcpyFromSymbol(out16, pb::oz_dbg, sizeof(unsigned long long) * 16));
    return PB_OK;
}
