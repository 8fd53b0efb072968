// Multi-GPU partitioning of the hot path: ONE LaplaceGP fit + predict over G GPUs of a box, one
// process per GPU, every collective enqueued from C++ on CUDA streams (no host callbacks on the data path).
//
//   Newton iterations (probit/implicit/solvers.py:18-25, Laplace.py:4-9)   -> fit.cu with a DistCtx:
//       rows of K sharded (each rank generates K[lo:hi, :] from the features), y = K x = local gemv + in-place
//       all-gather of 8N bytes; Nystrom preconditioner split by columns, A completed by one r x r all-reduce.
//   Cholesky of B = I + s s^T o K (Laplace.py:24, approximators.py:175)      -> bc_factor below:
//       1-D block-column-cyclic, right-looking.  Every rank FILLS only its own block columns straight from the
//       features (no replicated K, no replicated factor: memory per GPU is N^2/G), the owner of panel k+1 brings it
//       up to date and factors it on a high-priority side stream (look-ahead), the panel travels by ncclBroadcast
//       on a dedicated communication stream into one of three panel buffers while the previous trailing update
//       (DMMA GEMMs on the main stream) is still running.
//   predict (approximators.py:154-180)                                       -> pb_dist_predict:
//       test points sharded, no collective on the data path.  The rows V = s o k(X*, X) of a rank's test shard
//       ride along with the factorisation as extra rows of the matrix: each arriving panel applies
//       V_k <- V_k L_kk^-T, V[:, k+1:] -= V_k L[k+1:, k]^T, so var = k** - ||V||^2_rows needs no gathered factor.
//       Further chunks (or later predict calls) re-stream the stored panels (bc_stream).
#include "workspace.cuh"
#include <dlfcn.h>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>

namespace pb {

// ------------------------------------------------------------------------------------------------ NCCL via dlopen
namespace {

struct NcclId { char internal[128]; };

struct NcclApi {
    int (*GetUniqueId)(NcclId*) = nullptr;
    int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    int (*GetVersion)(int*) = nullptr;
    bool ok = false;
    char why[256] = "";
};

constexpr int NCCL_INT64 = 4, NCCL_FLOAT64 = 8, NCCL_SUM = 0, NCCL_MAX = 2;

NcclApi& nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        // torch.distributed (the plumbing) has normally loaded its bundled libnccl.so.2 already: reuse that image
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        if (!h) {
            const char* env = getenv("PB_NCCL_LIB");
            if (env && *env) h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
        }
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) {
            snprintf(api.why, sizeof(api.why), "libnccl.so.2 not found (%s); set PB_NCCL_LIB", dlerror());
            return;
        }
        auto sym = [&](const char* name) -> void* {
            void* p = dlsym(h, name);
            if (!p && !api.why[0]) snprintf(api.why, sizeof(api.why), "symbol %s missing from libnccl", name);
            return p;
        };
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
        api.Broadcast = reinterpret_cast<decltype(api.Broadcast)>(sym("ncclBroadcast"));
        api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
        api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
        api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(sym("ncclGetVersion"));
        api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.Broadcast && api.AllGather &&
                 api.AllReduce && api.GetErrorString;
    });
    return api;
}

#define PB_NCCL(expr)                                                                              \
    do {                                                                                           \
        int _r = (expr);                                                                           \
        if (_r != 0) {                                                                             \
            pb::set_error("%s failed: %s (%s:%d)", #expr, nccl().GetErrorString(_r), __FILE__, __LINE__); \
            return PB_ERR_CUDA;                                                                    \
        }                                                                                          \
    } while (0)

}  // namespace

int comm_allgather(Comm* c, cudaStream_t st, double* buf, int64_t per_rank) {
    if (!c || c->world == 1) return PB_OK;
    PB_NCCL(nccl().AllGather(buf + (int64_t)c->rank * per_rank, buf, (size_t)per_rank, NCCL_FLOAT64, c->nccl, st));
    return PB_OK;
}

int comm_allreduce_sum(Comm* c, cudaStream_t st, double* buf, int64_t count) {
    if (!c || c->world == 1) return PB_OK;
    PB_NCCL(nccl().AllReduce(buf, buf, (size_t)count, NCCL_FLOAT64, NCCL_SUM, c->nccl, st));
    return PB_OK;
}

int comm_allreduce_max_i64(Comm* c, cudaStream_t st, long long* dev, int count) {
    if (!c || c->world == 1) return PB_OK;
    PB_NCCL(nccl().AllReduce(dev, dev, (size_t)count, NCCL_INT64, NCCL_MAX, c->nccl, st));
    return PB_OK;
}

int comm_broadcast(Comm* c, cudaStream_t st, double* buf, int64_t count, int root) {
    if (!c || c->world == 1) return PB_OK;
    PB_NCCL(nccl().Broadcast(buf, buf, (size_t)count, NCCL_FLOAT64, root, c->nccl, st));
    return PB_OK;
}

int comm_event(Comm* c, size_t index, cudaEvent_t* out) {
    while (c->events.size() <= index) {
        cudaEvent_t e;
        PB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        c->events.push_back(e);
    }
    *out = c->events[index];
    return PB_OK;
}

namespace {

// ------------------------------------------------------------------------------------------------ small kernels
// *acc += sum_i log P[i * ld + i], i < w   (sum log diag L of one panel, Laplace.py:28)
__global__ void __launch_bounds__(256)
logdet_add_kernel(const double* __restrict__ P, int64_t ld, int w, double* __restrict__ acc) {
    double s = 0.0;
    for (int i = threadIdx.x; i < w; i += 256) s += log(P[(int64_t)i * ld + i]);
    s = block_sum<256>(s);
    if (threadIdx.x == 0) *acc += s;
}

// LAPACK-style info of the whole matrix from the info of one panel factorisation
__global__ void fold_info_kernel(const int32_t* __restrict__ panel_info, int32_t col0, int32_t* __restrict__ info) {
    if (*panel_info != 0 && *info == 0) *info = *panel_info + col0;
}

// info words <-> a max-reducible key (smallest failing column wins; 0 = none)
__global__ void info_encode_kernel(const int32_t* __restrict__ info, long long* __restrict__ key) {
    *key = *info != 0 ? (long long)(0x7fffffff - *info) : -1ll;
}
__global__ void info_decode_kernel(const long long* __restrict__ key, int32_t* __restrict__ info) {
    *info = *key >= 0 ? (int32_t)(0x7fffffff - *key) : 0;
}

constexpr int64_t OZ_DIST_MIN_ROWS = 2048;     // smaller updates stay on the DMMA kernels

struct Bc {                      // geometry of the block-column-cyclic layout on this rank
    int64_t n, nb, nblk, ld_loc;
    int world, me;
    double* Aloc;
    int64_t width(int64_t k) const { return std::min(nb, n - k * nb); }
    int owner(int64_t k) const { return (int)(k % world); }
    double* col(int64_t j) const { return Aloc + (j * nb) * ld_loc + (j / world) * nb; }    // (row j0, first column) of block column j
    int64_t count(int64_t k) const { return nb * 64 + (n - k * nb) * nb; }                    // doubles in panel message k
    // INT8 path: the rows of panel k BELOW its diagonal block are sliced once (ship_panel); block row j > k starts at
    // sliced row (j - k - 1) nb.  Only full-width panels with enough rows below them are sliced.
    int64_t oz_bytes = 0;
    int64_t sliced_rows(int64_t k) const { return n - (k + 1) * nb; }
    bool sliced(int64_t k) const { return oz_bytes > 0 && width(k) == nb && sliced_rows(k) >= OZ_DIST_MIN_ROWS; }
};

Bc make_bc(const Ws& ws) {
    Bc b;
    b.n = ws.L.n; b.nb = ws.L.nb; b.nblk = ws.L.nblk; b.ld_loc = ws.L.ld_loc;
    b.world = ws.L.world; b.me = ws.L.rank;
    b.Aloc = ws.B();
    b.oz_bytes = ws.L.oz_slice_bytes;
    return b;
}

// hook(k, k0, w, dinv, P) runs on the main stream once panel k (leaf inverses `dinv` of its diagonal block, rows
// k0.. of the factor's columns [k0, k0 + w) in `P`, leading dimension nb) is on this rank.
// `slices` (may be null) are the digit planes of the panel's rows below its diagonal block (Bc::sliced).
using PanelHook = std::function<int(int64_t, int64_t, int64_t, const double*, const double*, const void*)>;

enum { EV_RECV = 0, EV_PACKED, EV_TRAIL, EV_COLREADY, EV_COPIED, EV_KINDS };

struct Events {
    Comm* c;
    int get(int kind, int64_t k, cudaEvent_t* e) const { return comm_event(c, (size_t)(k * EV_KINDS + kind) + 1, e); }
};

// Optional timeline of one factorisation (debugging aid): PB_DIST_TRACE=<path prefix> makes every rank write
// <prefix>.<rank>.csv with, per panel step, the start / end times (ms from the first event) of the trailing update on
// the main stream, of the look-ahead panel work on the side stream and of the panel broadcast on the communication
// stream.  Costs a device synchronise at the end of the factorisation; off by default.
struct Trace {
    struct Rec { int kind; int64_t k; cudaEvent_t e; };
    std::vector<Rec> recs;
    const char* prefix = nullptr;
    Trace() { const char* p = getenv("PB_DIST_TRACE"); if (p && *p) prefix = p; }
    bool on() const { return prefix != nullptr; }
    void mark(int kind, int64_t k, cudaStream_t s) {
        if (!on()) return;
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return;
        cudaEventRecord(e, s);
        recs.push_back({kind, k, e});
    }
    void dump(int rank, int64_t nblk) {
        if (!on() || recs.empty()) return;
        cudaDeviceSynchronize();
        char path[512];
        snprintf(path, sizeof(path), "%s.%d.csv", prefix, rank);
        FILE* f = fopen(path, "w");
        if (f) {
            std::vector<float> t(6 * nblk, -1.f);
            for (const Rec& r : recs) {
                float ms = 0;
                if (cudaEventElapsedTime(&ms, recs[0].e, r.e) == cudaSuccess && r.k >= 0 && r.k < nblk) t[r.k * 6 + r.kind] = ms;
            }
            fprintf(f, "k,main_start,main_end,side_start,side_end,bcast_start,bcast_end\n");
            for (int64_t k = 0; k < nblk; ++k)
                fprintf(f, "%lld,%.3f,%.3f,%.3f,%.3f,%.3f,%.3f\n", (long long)k, t[k * 6], t[k * 6 + 1], t[k * 6 + 2], t[k * 6 + 3],
                        t[k * 6 + 4], t[k * 6 + 5]);
            fclose(f);
        }
        for (const Rec& r : recs) cudaEventDestroy(r.e);
        recs.clear();
    }
};
enum { TR_MAIN0 = 0, TR_MAIN1, TR_SIDE0, TR_SIDE1, TR_BC0, TR_BC1 };

// Owner: copy factored block column k (rows k0.., strided) + its leaf inverses into the contiguous panel message.
int pack_panel(cudaStream_t s, const Ws& ws, const Bc& b, int64_t k, const double* leaf_inv) {
    const int64_t w = b.width(k), rows = b.n - k * b.nb;
    double* msg = ws.panel(k);
    PB_CUDA(cudaMemcpyAsync(msg, leaf_inv, b.nb * 64 * sizeof(double), cudaMemcpyDeviceToDevice, s));
    PB_CUDA(cudaMemcpy2DAsync(msg + b.nb * 64, b.nb * sizeof(double), b.col(k), b.ld_loc * sizeof(double),
                              w * sizeof(double), rows, cudaMemcpyDeviceToDevice, s));
    return PB_OK;
}

// Column j (owned) -= P_k[rows >= j0] * P_k[rows of block j]^T
int update_column(cudaStream_t s, const Ws& ws, const Bc& b, int64_t j, int64_t k) {
    const double* P = ws.panel(k) + b.nb * 64;
    const int64_t off = (j - k) * b.nb;
    if (b.sliced(k) && b.n - j * b.nb >= OZ_DIST_MIN_ROWS) {
        const int64_t so = off - b.nb;             // block row j inside the sliced rows
        return ozaki_apply(s, b.nb, ws.oz_slice(k), b.sliced_rows(k), so, b.n - j * b.nb, ws.oz_slice(k), b.sliced_rows(k), so,
                           b.width(j), -1.0, b.col(j), b.ld_loc, false);
    }
    return gemm_nt(s, b.n - j * b.nb, b.width(j), b.width(k), -1.0, P + off * b.nb, b.nb, P + off * b.nb, b.nb, 1.0, b.col(j),
                   b.ld_loc, false);
}

// Factor block column k on stream s (diagonal block + the rows below), keep its leaf inverses, and leave the packed
// panel message in ws.panel(k).
// Fast path (w a multiple of 256): the rows below the diagonal block are solved with the 256 x 256 diagonal-block
// inverses potrf leaves in its workspace — X_b <- (X_b - sum_{a<b} Y_a L_ba^T) L_bb^-T, one or two large GEMMs per
// 256 columns instead of the 15 dependent launches of the recursive TRSM — and the results go STRAIGHT into the
// message (out of place), so the 2-D pack copy of the whole column disappears from the panel's critical path.  The
// local copy of the factor (needed by bc_stream later) is refreshed from the message afterwards (*copy_back).
int factor_column(cudaStream_t s, const Ws& ws, const Bc& b, int64_t k, int32_t* info_panel, bool* copy_back) {
    const int64_t w = b.width(k), k0 = k * b.nb, below = b.n - k0 - w;
    double* A = b.col(k);
    *copy_back = false;
    PB_TRY(potrf(s, A, w, b.ld_loc, ws.potrf_ws(), pb_potrf_workspace_bytes(b.nb), info_panel));
    fold_info_kernel<<<1, 1, 0, s>>>(info_panel, (int32_t)k0, ws.info()); pb::note_launch();
    double* keep = ws.dinv_store() + (k / b.world) * b.nb * 64;
    PB_CUDA(cudaMemsetAsync(keep, 0, b.nb * 64 * sizeof(double), s));
    PB_CUDA(cudaMemcpyAsync(keep, ws.dinv(), ceil_div<int64_t>(w, 64) * 64 * 64 * sizeof(double), cudaMemcpyDeviceToDevice, s));
    if (below > 0 && w % 256 == 0 && w == b.nb) {
        double* msg = ws.panel(k);
        double* Y = msg + b.nb * 64 + w * b.nb;                      // rows below the diagonal block inside the message
        double* X = A + w * b.ld_loc;
        const double* tinv = ws.dinv() + ceil_div<int64_t>(w, 64) * 64 * 64;    // L_bb^-1, row-major, 256 x 256 each
        for (int64_t c = 0; c < w; c += 256) {
            if (c > 0) PB_TRY(gemm_nt(s, below, 256, c, -1.0, Y, b.nb, A + c * b.ld_loc, b.ld_loc, 1.0, X + c, b.ld_loc, false));
            PB_TRY(gemm_nt(s, below, 256, 256, 1.0, X + c, b.ld_loc, tinv + (c / 256) * 256 * 256, 256, 0.0, Y + c, b.nb, false));
        }
        PB_CUDA(cudaMemcpyAsync(msg, keep, b.nb * 64 * sizeof(double), cudaMemcpyDeviceToDevice, s));
        PB_CUDA(cudaMemcpy2DAsync(msg + b.nb * 64, b.nb * sizeof(double), A, b.ld_loc * sizeof(double), w * sizeof(double), w,
                                  cudaMemcpyDeviceToDevice, s));
        *copy_back = true;
        return PB_OK;
    }
    if (below > 0) PB_TRY(trsm_right_lt(s, A, w, b.ld_loc, ws.potrf_ws(), A + w * b.ld_loc, below, b.ld_loc));
    return pack_panel(s, ws, b, k, keep);
}

// Owner, after the panel is on its way: refresh the local block column below the diagonal from the message.
int unpack_column(cudaStream_t s, const Ws& ws, const Bc& b, int64_t k) {
    const int64_t w = b.width(k), below = b.n - k * b.nb - w;
    if (below <= 0) return PB_OK;
    PB_CUDA(cudaMemcpy2DAsync(b.col(k) + w * b.ld_loc, b.ld_loc * sizeof(double), ws.panel(k) + b.nb * 64 + w * b.nb,
                              b.nb * sizeof(double), w * sizeof(double), below, cudaMemcpyDeviceToDevice, s));
    return PB_OK;
}

// Enqueue the broadcast of panel k (root = its owner) on the communication stream and record EV_RECV[k].
// `free_after` (may be null) is the event after which this rank's target buffer is no longer read.
int ship_panel(Comm* c, const Ws& ws, const Bc& b, const Events& ev, int64_t k, cudaEvent_t free_after, Trace* tr = nullptr) {
    cudaEvent_t packed, recv;
    PB_TRY(ev.get(EV_PACKED, k, &packed));
    PB_TRY(ev.get(EV_RECV, k, &recv));
    cudaStream_t cs = c->comm_stream;
    if (free_after) PB_CUDA(cudaStreamWaitEvent(cs, free_after, 0));
    if (b.owner(k) == b.me) PB_CUDA(cudaStreamWaitEvent(cs, packed, 0));
    if (tr) tr->mark(TR_BC0, k, cs);
    PB_TRY(comm_broadcast(c, cs, ws.panel(k), b.count(k), b.owner(k)));
    if (b.sliced(k))       // digit planes of the rows below the diagonal block, into the slicing buffer that goes with this panel buffer
        PB_TRY(ozaki_slice(cs, ws.panel(k) + b.nb * 64 + b.nb * b.nb, b.sliced_rows(k), b.nb, b.nb, ws.oz_slice(k), b.oz_bytes));
    if (tr) tr->mark(TR_BC1, k, cs);
    PB_CUDA(cudaEventRecord(recv, cs));
    return PB_OK;
}

// Right-looking block-cyclic Cholesky of a I + s s^T o (K + jitter I), generated from the features in ws.Z().
// Streams: `st` trailing updates + hook, c->side_stream look-ahead panel work, c->comm_stream broadcasts.
int bc_factor(Comm* c, cudaStream_t st, const Ws& ws, const pb_problem* prob, const double* s, double a, double jitter,
              const PanelHook& hook) {
    const Bc b = make_bc(ws);
    const Events ev{c};
    const int Df = feature_dim(prob->kernel, prob->D);
    cudaStream_t side = c->side_stream;
    int32_t* info_panel = ws.info() + 2;
    Trace trace;
    Trace* tr = trace.on() ? &trace : nullptr;
    if (tr) tr->mark(TR_MAIN0, -1, st);
    PB_CUDA(cudaMemsetAsync(ws.info(), 0, 256, st));
    // every rank fills the block columns it owns (rows j0.. only): 8 N^2 / (2G) bytes, no communication
    for (int64_t j = b.me; j < b.nblk; j += b.world)
        PB_TRY(gram_block(st, prob->kernel, ws.Z(), b.n, Df, j * b.nb, b.n - j * b.nb, j * b.nb, b.width(j), s, a, jitter,
                          b.col(j), b.ld_loc));
    cudaEvent_t filled;
    PB_TRY(comm_event(c, 0, &filled));
    PB_CUDA(cudaEventRecord(filled, st));
    PB_CUDA(cudaStreamWaitEvent(side, filled, 0));
    PB_CUDA(cudaStreamWaitEvent(c->comm_stream, filled, 0));
    if (b.owner(0) == b.me) {
        cudaEvent_t packed;
        PB_TRY(ev.get(EV_PACKED, 0, &packed));
        if (tr) tr->mark(TR_SIDE0, 0, side);
        bool copy_back = false;
        PB_TRY(factor_column(side, ws, b, 0, info_panel, &copy_back));
        if (tr) tr->mark(TR_SIDE1, 0, side);
        PB_CUDA(cudaEventRecord(packed, side));
        cudaEvent_t copied;
        PB_TRY(ev.get(EV_COPIED, 0, &copied));
        if (copy_back) PB_TRY(unpack_column(side, ws, b, 0));
        PB_CUDA(cudaEventRecord(copied, side));
    }
    PB_TRY(ship_panel(c, ws, b, ev, 0, nullptr, tr));
    for (int64_t k = 0; k < b.nblk; ++k) {
        cudaEvent_t recv_k, trail_k, trail_km2 = nullptr;
        PB_TRY(ev.get(EV_RECV, k, &recv_k));
        PB_TRY(ev.get(EV_TRAIL, k, &trail_k));
        if (k >= 2) PB_TRY(ev.get(EV_TRAIL, k - 2, &trail_km2));
        const int64_t nx = k + 1;
        if (nx < b.nblk) {
            if (b.owner(nx) == b.me) {
                // look-ahead: column k+1 has every update through panel k-1 (main stream, EV_COLREADY); apply panel
                // k, factor, pack into buffer (k+1) % 3, which the main stream last read during step k-2
                cudaEvent_t ready, packed;
                PB_CUDA(cudaStreamWaitEvent(side, recv_k, 0));
                if (nx >= 2) {
                    PB_TRY(ev.get(EV_COLREADY, nx, &ready));
                    PB_CUDA(cudaStreamWaitEvent(side, ready, 0));
                }
                if (trail_km2) PB_CUDA(cudaStreamWaitEvent(side, trail_km2, 0));
                if (tr) tr->mark(TR_SIDE0, nx, side);
                PB_TRY(update_column(side, ws, b, nx, k));
                bool copy_back = false;
                PB_TRY(factor_column(side, ws, b, nx, info_panel, &copy_back));
                if (tr) tr->mark(TR_SIDE1, nx, side);
                PB_TRY(ev.get(EV_PACKED, nx, &packed));
                PB_CUDA(cudaEventRecord(packed, side));
                cudaEvent_t copied;
                PB_TRY(ev.get(EV_COPIED, nx, &copied));
                if (copy_back) PB_TRY(unpack_column(side, ws, b, nx));     // off the critical path: the broadcast is already released
                PB_CUDA(cudaEventRecord(copied, side));
            }
            PB_TRY(ship_panel(c, ws, b, ev, nx, trail_km2, tr));
        }
        PB_CUDA(cudaStreamWaitEvent(st, recv_k, 0));
        if (tr) tr->mark(TR_MAIN0, k, st);
        // trailing update of the owned block columns right of the look-ahead column, nearest first: column k+2 is
        // the next one the panel chain needs, so it is released (EV_COLREADY) before the rest of the update
        int64_t first = k + 2;
        first += ((b.me - first) % b.world + b.world) % b.world;            // first owned column >= k + 2
        int64_t j = first;
        if (j < b.nblk && j == k + 2) {
            PB_TRY(update_column(st, ws, b, j, k));
            cudaEvent_t ready;
            PB_TRY(ev.get(EV_COLREADY, j, &ready));
            PB_CUDA(cudaEventRecord(ready, st));
            j += b.world;
        }
        if (b.sliced(k))                     // tall columns on the INT8 tensor cores, one launch each
            for (; j < b.nblk && b.n - j * b.nb >= OZ_DIST_MIN_ROWS; j += b.world) PB_TRY(update_column(st, ws, b, j, k));
        if (j < b.nblk) {
            const int64_t count = (b.nblk - 1 - j) / b.world + 1;
            if (count >= 2 && b.nb % 128 == 0 && b.width(k) == b.nb) {
                // every remaining owned block column in ONE grid (gemm_dmma.cu GemmGroups): group q is column j + q world
                PB_TRY(gemm_nt_groups(st, ws.panel(k) + b.nb * 64, b.nb, b.n - k * b.nb, b.nb, -1.0, 1.0, b.col(j), b.ld_loc,
                                      (int64_t)b.world * b.nb * b.ld_loc + b.nb, (int)count, (j - k) * b.nb,
                                      (int64_t)b.world * b.nb, b.nb));
            } else {
                for (; j < b.nblk; j += b.world) PB_TRY(update_column(st, ws, b, j, k));
            }
        }
        if (hook) PB_TRY(hook(k, k * b.nb, b.width(k), ws.panel(k), ws.panel(k) + b.nb * 64, b.sliced(k) ? ws.oz_slice(k) : nullptr));
        if (b.owner(k) == b.me) {            // the local copy of column k is refreshed from buffer k % 3: keep the buffer until then
            cudaEvent_t copied;
            PB_TRY(ev.get(EV_COPIED, k, &copied));
            PB_CUDA(cudaStreamWaitEvent(st, copied, 0));
        }
        if (tr) tr->mark(TR_MAIN1, k, st);
        PB_CUDA(cudaEventRecord(trail_k, st));
    }
    if (tr) tr->dump(b.me, b.nblk);
    // smallest failing column over the ranks -> ws.info() on every rank
    info_encode_kernel<<<1, 1, 0, st>>>(ws.info(), c->dev_i64); pb::note_launch();
    PB_TRY(comm_allreduce_max_i64(c, st, c->dev_i64, 1));
    info_decode_kernel<<<1, 1, 0, st>>>(c->dev_i64, ws.info()); pb::note_launch();
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

// Re-broadcast the stored factor panel by panel (ascending) and run the hook on each: used by later test chunks
// and later predict calls.  Three panel buffers, so packing / broadcasting run ahead of the hook.
int bc_stream(Comm* c, cudaStream_t st, const Ws& ws, const PanelHook& hook) {
    const Bc b = make_bc(ws);
    const Events ev{c};
    cudaStream_t side = c->side_stream;
    cudaEvent_t start;
    PB_TRY(comm_event(c, 0, &start));
    PB_CUDA(cudaEventRecord(start, st));
    PB_CUDA(cudaStreamWaitEvent(side, start, 0));
    PB_CUDA(cudaStreamWaitEvent(c->comm_stream, start, 0));
    for (int64_t k = 0; k < b.nblk; ++k) {
        cudaEvent_t free_after = nullptr, packed, recv_k, trail_k;
        if (k >= 3) PB_TRY(ev.get(EV_TRAIL, k - 3, &free_after));
        if (b.owner(k) == b.me) {
            if (free_after) PB_CUDA(cudaStreamWaitEvent(side, free_after, 0));
            PB_TRY(pack_panel(side, ws, b, k, ws.dinv_store() + (k / b.world) * b.nb * 64));
            PB_TRY(ev.get(EV_PACKED, k, &packed));
            PB_CUDA(cudaEventRecord(packed, side));
        }
        PB_TRY(ship_panel(c, ws, b, ev, k, free_after));
        PB_TRY(ev.get(EV_RECV, k, &recv_k));
        PB_TRY(ev.get(EV_TRAIL, k, &trail_k));
        PB_CUDA(cudaStreamWaitEvent(st, recv_k, 0));
        PB_TRY(hook(k, k * b.nb, b.width(k), ws.panel(k), ws.panel(k) + b.nb * 64, b.sliced(k) ? ws.oz_slice(k) : nullptr));
        PB_CUDA(cudaEventRecord(trail_k, st));
    }
    return PB_OK;
}

int bind_dist(const pb_problem* prob, Comm* comm, void* workspace, int64_t workspace_bytes, Ws& ws, DistCtx& ctx) {
    PB_CHECK(prob != nullptr && prob->n >= 1 && prob->D >= 1 && prob->X && prob->y, PB_ERR_INVALID, "dist: bad problem");
    PB_CHECK(comm != nullptr, PB_ERR_INVALID, "dist: null communicator");
    ws.L = make_dist_layout(prob->n, prob->D, comm->world, comm->rank);
    PB_CHECK(workspace != nullptr && workspace_bytes >= ws.L.total, PB_ERR_INVALID,
             "dist workspace too small: need %lld bytes, got %lld", (long long)ws.L.total, (long long)workspace_bytes);
    PB_CHECK((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, PB_ERR_INVALID, "workspace must be 256-byte aligned");
    ws.base = reinterpret_cast<uint8_t*>(workspace);
    ctx.comm = comm;
    ctx.nloc_max = ws.L.nloc_max;
    ctx.lo = std::min<int64_t>(prob->n, (int64_t)comm->rank * ctx.nloc_max);
    ctx.hi = std::min<int64_t>(prob->n, ctx.lo + ctx.nloc_max);
    ws.dist = &ctx;
    return PB_OK;
}

}  // namespace
}  // namespace pb

using namespace pb;

extern "C" int pb_comm_unique_id(void* id_host) {
    PB_CHECK(id_host != nullptr, PB_ERR_INVALID, "comm_unique_id: null argument");
    PB_CHECK(nccl().ok, PB_ERR_UNSUPPORTED, "NCCL unavailable: %s", nccl().why);
    PB_NCCL(nccl().GetUniqueId(reinterpret_cast<NcclId*>(id_host)));
    return PB_OK;
}

extern "C" int pb_comm_create(const void* id_host, int32_t rank, int32_t world, pb_comm** out) {
    PB_CHECK(out != nullptr && world >= 1 && rank >= 0 && rank < world, PB_ERR_INVALID, "comm_create: bad rank/world");
    PB_CHECK(world == 1 || id_host != nullptr, PB_ERR_INVALID, "comm_create: unique id missing");
    pb_comm* c = new pb_comm();
    c->rank = rank;
    c->world = world;
    PB_CUDA(cudaGetDevice(&c->device));
    int lo = 0, hi = 0;
    PB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    PB_CUDA(cudaStreamCreateWithPriority(&c->comm_stream, cudaStreamNonBlocking, hi));
    PB_CUDA(cudaStreamCreateWithPriority(&c->side_stream, cudaStreamNonBlocking, hi));
    PB_CUDA(cudaMalloc(&c->dev_i64, 8 * sizeof(long long)));
    if (world > 1) {
        PB_CHECK(nccl().ok, PB_ERR_UNSUPPORTED, "NCCL unavailable: %s", nccl().why);
        NcclId id;
        memcpy(&id, id_host, sizeof(id));
        PB_NCCL(nccl().CommInitRank(&c->nccl, world, id, rank));
    }
    *out = c;
    return PB_OK;
}

extern "C" int pb_comm_destroy(pb_comm* c) {
    if (!c) return PB_OK;
    if (c->nccl) nccl().CommDestroy(c->nccl);
    for (cudaEvent_t e : c->events) cudaEventDestroy(e);
    if (c->comm_stream) cudaStreamDestroy(c->comm_stream);
    if (c->side_stream) cudaStreamDestroy(c->side_stream);
    if (c->dev_i64) cudaFree(c->dev_i64);
    delete c;
    return PB_OK;
}

extern "C" int pb_comm_rank(const pb_comm* c) { return c ? c->rank : 0; }
extern "C" int pb_comm_size(const pb_comm* c) { return c ? c->world : 1; }

extern "C" int64_t pb_dist_workspace_bytes(int64_t n, int D, int32_t world, int32_t rank, const pb_options* options) {
    OptScope opt_scope(options);
    return make_dist_layout(n, D, world, rank).total;
}

extern "C" int pb_dist_laplace_fit(pb_stream_t stream, pb_comm* comm, const pb_problem* prob, double tolerance,
                                   int32_t maxiter, void* workspace, int64_t workspace_bytes, double* weight,
                                   double* precision, double* posterior_mean, pb_fit_result* result_host,
                                   const pb_options* options) {
    OptScope opt_scope(options);
    Ws ws;
    DistCtx ctx;
    PB_TRY(bind_dist(prob, comm, workspace, workspace_bytes, ws, ctx));
    return laplace_fit_impl(reinterpret_cast<cudaStream_t>(stream), prob, tolerance, maxiter, 0.0, 0, ws, weight, precision,
                            posterior_mean, result_host);
}

extern "C" int64_t pb_dist_predict_scratch_bytes(int64_t n, int D, int64_t rows) {
    const int64_t ld = round_up(n > 0 ? n : 1, 16);
    rows = rows > 0 ? rows : 1;
    // V (rows x ld) | features of the chunk | partial means | digit planes of one rows x nb block column of V (INT8 path)
    return round_up(rows * ld * 8, 256) + round_up(rows * 2 * (int64_t)D * 8, 256) + round_up(64 * rows * 8, 256) + 256 +
           round_up(ozaki_scratch_bytes(rows, 1024), 256);
}

extern "C" int pb_dist_predict(pb_stream_t stream, pb_comm* comm, const pb_problem* prob, const double* precision,
                               const double* weight, void* workspace, int64_t workspace_bytes, double jitter,
                               int32_t reuse_factor, const double* X_test, int64_t n_test, int64_t chunk, void* scratch,
                               int64_t scratch_bytes, double* mean, double* variance, double* logdet_host,
                               int32_t* info_host, const pb_options* options) {
    OptScope opt_scope(options);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    Ws ws;
    DistCtx ctx;
    PB_TRY(bind_dist(prob, comm, workspace, workspace_bytes, ws, ctx));
    PB_CHECK(precision != nullptr && info_host != nullptr, PB_ERR_INVALID, "dist_predict: null argument");
    PB_CHECK(n_test >= 0 && (n_test == 0 || (X_test && weight && mean)), PB_ERR_INVALID, "dist_predict: bad test arguments");
    const int64_t n = prob->n, ldv = ws.L.ld;
    const int D = prob->D, Df = feature_dim(prob->kernel, D);
    const bool want_var = variance != nullptr && n_test > 0;
    if (n_test > 0) {
        PB_CHECK(chunk >= 1, PB_ERR_INVALID, "dist_predict: chunk must be positive");
        PB_CHECK(scratch != nullptr && scratch_bytes >= pb_dist_predict_scratch_bytes(n, D, std::min(chunk, n_test)),
                 PB_ERR_INVALID, "dist_predict: scratch too small");
        PB_CHECK((reinterpret_cast<uintptr_t>(scratch) & 255) == 0, PB_ERR_INVALID, "dist_predict: scratch must be 256-byte aligned");
    }
    const int64_t rows_cap = n_test > 0 ? std::min(chunk, n_test) : 1;
    double* V = reinterpret_cast<double*>(scratch);
    double* Zs = n_test > 0 ? reinterpret_cast<double*>(reinterpret_cast<uint8_t*>(scratch) + round_up(rows_cap * ldv * 8, 256)) : nullptr;
    double* mean_partial = n_test > 0 ? reinterpret_cast<double*>(reinterpret_cast<uint8_t*>(Zs) + round_up(rows_cap * 2 * (int64_t)D * 8, 256)) : nullptr;
    void* v_slices = n_test > 0 ? reinterpret_cast<uint8_t*>(mean_partial) + round_up(64 * rows_cap * 8, 256) + 256 : nullptr;
    const int64_t v_slice_bytes = ozaki_scratch_bytes(rows_cap, 1024);
    const double kss = prob->kernel.scale;          // kernel.elwise(x*, x*) of a stationary kernel (approximators.py:172)

    PB_TRY(features(st, prob->kernel, prob->X, n, D, D, ws.Z(), n));
    PB_CUDA(cudaMemsetAsync(ws.scalars(), 0, S_COUNT * sizeof(double), st));
    PB_TRY(precision_sqrt(st, ws, prob, precision));

    // every rank must walk the panels the same number of times: passes = max over ranks of ceil(n_test / chunk)
    long long my_passes = want_var ? (long long)ceil_div<int64_t>(n_test, rows_cap) : 0;
    long long passes = my_passes;
    if (comm->world > 1) {
        PB_CUDA(cudaMemcpyAsync(comm->dev_i64 + 1, &my_passes, sizeof(long long), cudaMemcpyHostToDevice, st));
        PB_TRY(comm_allreduce_max_i64(comm, st, comm->dev_i64 + 1, 1));
        PB_CUDA(cudaMemcpyAsync(&passes, comm->dev_i64 + 1, sizeof(long long), cudaMemcpyDeviceToHost, st));
        PB_CUDA(cudaStreamSynchronize(st));
    }
    const bool need_factor = !reuse_factor && (passes > 0 || logdet_host != nullptr);

    int64_t m_cur = 0;                              // rows of V in flight during the current pass
    PanelHook apply = [&](int64_t k, int64_t k0, int64_t w, const double* dinv, const double* P, const void* slices) -> int {
        (void)k;
        if (m_cur <= 0) return PB_OK;
        // V_k <- V_k L_kk^-T ; V[:, k1:] -= V_k L[k1:, k]^T   (B.solve(K, Kfs) at approximators.py:177, one panel at a time)
        PB_TRY(trsm_right_lt(st, P, w, ws.L.nb, dinv, V + k0, m_cur, ldv));
        const int64_t k1 = k0 + w;
        if (k1 >= n) return PB_OK;
        if (slices && m_cur >= 256) {       // the panel's rows below its diagonal block are already sliced: slice V_k and multiply
            PB_TRY(ozaki_slice(st, V + k0, m_cur, w, ldv, v_slices, v_slice_bytes));
            return ozaki_apply(st, w, v_slices, m_cur, 0, m_cur, slices, n - k1, 0, n - k1, -1.0, V + k1, ldv, false);
        }
        return gemm_nt(st, m_cur, n - k1, w, -1.0, V + k0, ldv, P + w * ws.L.nb, ws.L.nb, 1.0, V + k1, ldv, false);
    };
    PanelHook apply_and_logdet = [&](int64_t k, int64_t k0, int64_t w, const double* dinv, const double* P, const void* slices) -> int {
        logdet_add_kernel<<<1, 256, 0, st>>>(P, ws.L.nb, (int)w, ws.scalars() + S_LOGDET); pb::note_launch();
        return apply(k, k0, w, dinv, P, slices);
    };

    auto stage_chunk = [&](int64_t r0, int64_t m) -> int {       // mean of the chunk; V = s o k(X*, X) if variances are wanted
        PB_TRY(features(st, prob->kernel, X_test + r0 * D, m, D, D, Zs, rows_cap));
        PB_TRY(gram_matvec(st, prob->kernel, Zs, m, ws.Z(), n, Df, rows_cap, n, weight, mean_partial, mean + r0));
        if (want_var) PB_TRY(gram_cross(st, prob->kernel, Zs, m, ws.Z(), n, Df, rows_cap, n, V, ldv, ws.vec(V_S)));
        return PB_OK;
    };

    const long long total_passes = std::max<long long>(passes, need_factor ? 1 : 0);
    for (long long p = 0; p < total_passes; ++p) {
        const int64_t r0 = p * rows_cap;
        m_cur = (want_var && r0 < n_test) ? std::min(rows_cap, n_test - r0) : 0;
        if (m_cur > 0) PB_TRY(stage_chunk(r0, m_cur));
        if (p == 0 && need_factor) PB_TRY(bc_factor(comm, st, ws, prob, ws.vec(V_S), 1.0, jitter, apply_and_logdet));
        else PB_TRY(bc_stream(comm, st, ws, apply));
        if (m_cur > 0) PB_TRY(row_sumsq(st, V, m_cur, n, ldv, kss, variance + r0));
    }
    if (!want_var)                                   // mean-only sweep: no panels involved, chunk by chunk
        for (int64_t r0 = 0; r0 < n_test; r0 += rows_cap) PB_TRY(stage_chunk(r0, std::min(rows_cap, n_test - r0)));

    double host[S_COUNT];
    PB_CUDA(cudaMemcpyAsync(host, ws.scalars(), S_COUNT * sizeof(double), cudaMemcpyDeviceToHost, st));
    PB_CUDA(cudaMemcpyAsync(info_host, ws.info(), sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    PB_CUDA(cudaStreamSynchronize(st));
    if (!need_factor) *info_host = 0;
    if (logdet_host) *logdet_host = need_factor ? host[S_LOGDET] : NAN;
    if (host[S_BAD] > 0) {
        set_error("dist_predict: %d precisions are materially negative or NaN", (int)host[S_BAD]);
        return PB_ERR_NUMERIC;
    }
    if (*info_host != 0) {
        set_error("dist_predict: Cholesky failed at column %d", *info_host);
        return PB_ERR_NUMERIC;
    }
    return PB_OK;
}
