// Private layout of the caller-provided workspaces (fit.cu, dist.cu).
//
//  single GPU : K (n x ld) | factor buffer B | features Z | permuted features | potrf workspace | O(n) vectors | partials | scalars | info
//  multi GPU  : the same slots, but `K` holds only this rank's rows [lo, hi) of the Gram matrix (nloc_max x ld) and
//               `B` is the rank's share of the block-column-cyclic factor (n x ld_loc); three panel buffers, the
//               leaf inverses of the owned panels and a panel-sized potrf workspace follow.  During the Newton
//               iterations (no factor yet) the factor region holds the Nystrom preconditioner, as on one GPU.
#pragma once
#include "common.cuh"
#include "dist.cuh"
#include <algorithm>

namespace pb {

constexpr int VEC_BLOCKS_MAX = 1024;

inline int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }

enum VecSlot { V_W = 0, V_WN, V_F, V_S, V_B, V_T, V_C, V_X, V_G, V_SF, V_E, V_R, V_Z, V_P, V_Q, V_Y, V_U, V_COUNT };
enum Scalar { S_ERR2 = 0, S_SUMLL, S_FTW, S_LOGDET, S_BAD, S_PBP, S_RZ0, S_RZ1, S_RR, S_R0, S_G0, S_G1, S_G2, S_GD0, S_GD1, S_GSIG, S_COUNT = 16 };

struct Layout {
    int64_t n = 0, ld = 0;
    int Dfmax = 0;
    // byte offsets
    int64_t K = 0, B = 0, Z = 0, Zp = 0, potrf_ws = 0, vec = 0, partial = 0, scalars = 0, info = 0, total = 0;
    int64_t vec_stride = 0;   // doubles per vector slot
    int64_t B_doubles = 0;    // capacity of the factor region
    // multi-GPU only
    int world = 1, rank = 0;
    int64_t nb = 0, nblk = 0, owned_max = 0, ld_loc = 0, nloc_max = 0;
    int64_t panel_doubles = 0, panels = 0, dinv_store = 0;
    int64_t oz_slices = 0, oz_slice_bytes = 0;   // three int8 slicing buffers, one per panel buffer (0 bytes: FP64 DMMA only)
};

inline Layout make_layout(int64_t n, int D) {
    Layout L;
    L.n = n;
    L.ld = round_up(n > 0 ? n : 1, 16);
    L.Dfmax = 2 * D;
    int64_t off = 0;
    auto take = [&](int64_t bytes) { int64_t o = off; off += round_up(bytes, 256); return o; };
    L.K = take(n * L.ld * 8);
    // the factor region also takes the (n + 1) x ld(n + 1) signed matrix of an indefinite Newton step (fit.cu)
    L.B_doubles = (n + 1) * round_up(n + 1, 16);
    L.B = take(L.B_doubles * 8);
    L.Z = take(n * (int64_t)L.Dfmax * 8);
    L.Zp = take((n + 1) * (int64_t)L.Dfmax * 8);
    L.potrf_ws = take(pb_potrf_workspace_bytes(n + 1));
    L.vec_stride = round_up(n + 1, 32);
    L.vec = take(L.vec_stride * V_COUNT * 8);
    L.partial = take(VEC_BLOCKS_MAX * 2 * 8);
    L.scalars = take(S_COUNT * 8);
    L.info = take(256);
    L.total = off;
    return L;
}

// Panel width of the block-cyclic factorisation: multiple of 64 (leaf size); NB = 512 keeps every trailing update a
// K = 512 DMMA GEMM (34 TFLOP/s) while the panel chain (N^2 NB flops on one GPU at a time) stays short.
inline int64_t dist_block(int64_t n, int world) {
    int64_t nb = opts().dist_block > 0 ? opts().dist_block / 64 * 64 : 0;
    if (nb <= 0) nb = n >= 16384 ? 512 : (n >= 4096 ? 256 : 128);
    return nb < 64 ? 64 : nb;
}

// Doubles the Nystrom preconditioner of the sharded Newton iterations needs inside the factor region.
int64_t dist_nystrom_doubles(int64_t n, int64_t nloc_max, int Dfmax);

inline Layout make_dist_layout(int64_t n, int D, int world, int rank) {
    Layout L;
    L.n = n;
    L.ld = round_up(n > 0 ? n : 1, 16);
    L.Dfmax = 2 * D;
    L.world = world;
    L.rank = rank;
    L.nb = dist_block(n, world);
    L.nblk = ceil_div<int64_t>(n, L.nb);
    L.owned_max = ceil_div<int64_t>(L.nblk, world);
    L.ld_loc = L.owned_max * L.nb;
    L.nloc_max = dist_rows_per_rank(n, world);
    int64_t off = 0;
    auto take = [&](int64_t bytes) { int64_t o = off; off += round_up(bytes, 256); return o; };
    L.K = take(L.nloc_max * L.ld * 8);
    L.B_doubles = std::max<int64_t>(n * L.ld_loc, dist_nystrom_doubles(n, L.nloc_max, L.Dfmax));
    L.B = take(L.B_doubles * 8);
    L.Z = take(n * (int64_t)L.Dfmax * 8);
    L.potrf_ws = take(pb_potrf_workspace_bytes(L.nb));
    L.panel_doubles = L.nb * 64 + n * L.nb;                   // [leaf inverses of the diagonal block | panel rows]
    L.panels = take(3 * L.panel_doubles * 8);
    L.dinv_store = take(L.owned_max * L.nb * 64 * 8);
    // trailing updates on the INT8 tensor cores: every arriving panel is sliced once into digit planes (ozaki.cu)
    L.oz_slice_bytes = (ozaki_enabled(n) && ozaki_supported(L.nb) && L.nb <= 1024) ? ozaki_scratch_bytes(n, L.nb) : 0;
    L.oz_slices = take(3 * L.oz_slice_bytes);
    L.vec_stride = round_up(world * L.nloc_max, 32);
    L.vec = take(L.vec_stride * V_COUNT * 8);
    L.partial = take(VEC_BLOCKS_MAX * 2 * 8);
    L.scalars = take(S_COUNT * 8);
    L.info = take(256);
    L.total = off;
    return L;
}

struct Ws {
    Layout L;
    uint8_t* base = nullptr;
    const DistCtx* dist = nullptr;      // non-null: K() holds rows [dist->lo, dist->hi) only
    double* K() const { return reinterpret_cast<double*>(base + L.K); }
    double* B() const { return reinterpret_cast<double*>(base + L.B); }
    double* Z() const { return reinterpret_cast<double*>(base + L.Z); }
    void* potrf_ws() const { return base + L.potrf_ws; }
    double* Zp() const { return reinterpret_cast<double*>(base + L.Zp); }
    double* dinv() const { return reinterpret_cast<double*>(base + L.potrf_ws); }
    double* vec(int slot) const { return reinterpret_cast<double*>(base + L.vec) + slot * L.vec_stride; }
    double* partial() const { return reinterpret_cast<double*>(base + L.partial); }
    double* scalars() const { return reinterpret_cast<double*>(base + L.scalars); }
    int32_t* info() const { return reinterpret_cast<int32_t*>(base + L.info); }
    double* panel(int64_t k) const { return reinterpret_cast<double*>(base + L.panels) + (k % 3) * L.panel_doubles; }
    double* dinv_store() const { return reinterpret_cast<double*>(base + L.dinv_store); }
    void* oz_slice(int64_t k) const { return L.oz_slice_bytes ? base + L.oz_slices + (k % 3) * L.oz_slice_bytes : nullptr; }
};

inline unsigned vec_blocks(int64_t n) {
    int64_t b = ceil_div<int64_t>(n, 256);
    return (unsigned)(b < 1 ? 1 : (b > VEC_BLOCKS_MAX ? VEC_BLOCKS_MAX : b));
}

int precision_sqrt(cudaStream_t st, const Ws& ws, const pb_problem* prob, const double* precision);
int row_sumsq(cudaStream_t st, const double* V, int64_t rows, int64_t cols, int64_t ld, double kss, double* var);

// fit.cu: the sharded Laplace fit (same Newton loop as pb_laplace_fit, rows of K and the Nystrom build split over the ranks)
int laplace_fit_impl(cudaStream_t st, const pb_problem* prob, double tolerance, int32_t maxiter, double jitter,
                     int32_t final_factor, Ws& ws, double* weight, double* precision, double* posterior_mean,
                     pb_fit_result* result_host);

}  // namespace pb
