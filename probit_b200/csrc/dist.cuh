// Multi-GPU plumbing shared by dist.cu (communicator, block-cyclic Cholesky, sharded predict) and fit.cu (the
// row-sharded Newton / CG iterations).  One process per GPU; NCCL over NVLink 5 / NVSwitch is reached through
// dlopen (the single-GPU library has no link-time dependency on it).
#pragma once
#include "common.cuh"
#include <vector>

// The opaque handle of include/probit_b200.h.
struct pb_comm {
    void* nccl = nullptr;           // ncclComm_t (null when world == 1)
    int rank = 0, world = 1, device = 0;
    cudaStream_t comm_stream = nullptr;    // panel broadcasts (overlap the trailing updates)
    cudaStream_t side_stream = nullptr;    // look-ahead panel factorisation (highest priority)
    std::vector<cudaEvent_t> events;       // grow-only pool, reused by every factorisation
    long long* dev_i64 = nullptr;          // 8 device words for small integer reductions
};

namespace pb {

using Comm = ::pb_comm;

// All collectives are enqueued on `st`; with world == 1 they are no-ops.
int comm_allgather(Comm* c, cudaStream_t st, double* buf, int64_t per_rank);       // in place: rank r owns buf[r * per_rank ...]
int comm_allreduce_sum(Comm* c, cudaStream_t st, double* buf, int64_t count);
int comm_allreduce_max_i64(Comm* c, cudaStream_t st, long long* dev, int count);
int comm_broadcast(Comm* c, cudaStream_t st, double* buf, int64_t count, int root);
int comm_event(Comm* c, size_t index, cudaEvent_t* out);

// Rows [lo, hi) of the N x N Gram matrix held by this rank (equal chunks of nloc_max rows, the last may be short).
struct DistCtx {
    Comm* comm;
    int64_t lo, hi, nloc_max;
};

inline int64_t dist_rows_per_rank(int64_t n, int world) { return ceil_div<int64_t>(ceil_div<int64_t>(n, world), 16) * 16; }

}  // namespace pb
