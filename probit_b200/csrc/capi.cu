// Error reporting and version entry points of the C ABI.
#include "common.cuh"
#include <cstring>
#include <atomic>
#include <mutex>
#include <vector>

namespace pb {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

static std::atomic<long long> g_launches{0};
static std::atomic<int> g_profiling{0};
static std::mutex g_prof_mu;
struct GemmRec { cudaEvent_t e0, e1; double flops; };
static std::vector<GemmRec> g_gemm;

void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
bool profiling_enabled() { return g_profiling.load(std::memory_order_relaxed) != 0; }
void profile_gemm(cudaEvent_t e0, cudaEvent_t e1, double flops) {
    std::lock_guard<std::mutex> lock(g_prof_mu);
    g_gemm.push_back({e0, e1, flops});
}

// Tunables (pb_set_option); defaults chosen from measurements on B200.
static std::atomic<long long> g_pcg_min_n{24576};     // PCG Newton steps pay off once potrf >> trsv
static std::atomic<long long> g_nystrom_rank{-1};     // -1 = automatic (n/16 clamped to [256, 4096]), 0 = off
static std::atomic<double> g_cg_tol{1e-2};            // CG Newton solves: error of the step <= this * Newton tolerance
static std::atomic<int> g_potrf_nb{0};                // 0 = automatic
static std::atomic<int> g_lookahead{1};

long long opt_pcg_min_n() { return g_pcg_min_n.load(std::memory_order_relaxed); }
long long opt_nystrom_rank() { return g_nystrom_rank.load(std::memory_order_relaxed); }
double opt_cg_tol() { return g_cg_tol.load(std::memory_order_relaxed); }
int opt_potrf_nb() { return g_potrf_nb.load(std::memory_order_relaxed); }
bool opt_lookahead() { return g_lookahead.load(std::memory_order_relaxed) != 0; }

}  // namespace pb

extern "C" int pb_set_option(const char* name, double value) {
    PB_CHECK(name != nullptr, PB_ERR_INVALID, "set_option: null name");
    if (!strcmp(name, "laplace_pcg_min_n")) pb::g_pcg_min_n.store((long long)value);
    else if (!strcmp(name, "laplace_nystrom_rank")) pb::g_nystrom_rank.store((long long)value);
    else if (!strcmp(name, "laplace_cg_tol")) pb::g_cg_tol.store(value);
    else if (!strcmp(name, "potrf_block")) pb::g_potrf_nb.store((int)value);
    else if (!strcmp(name, "potrf_lookahead")) pb::g_lookahead.store(value != 0.0);
    else PB_CHECK(false, PB_ERR_INVALID, "set_option: unknown option '%s'", name);
    return PB_OK;
}

extern "C" long long pb_launch_count(void) { return pb::g_launches.load(); }

extern "C" int pb_profile_begin(void) {
    std::lock_guard<std::mutex> lock(pb::g_prof_mu);
    for (auto& r : pb::g_gemm) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    pb::g_gemm.clear();
    pb::g_profiling.store(1);
    return PB_OK;
}

// Caller must have synchronised the device.  Returns the number of profiled main-config GEMM launches,
// their summed CUDA-event duration (ms) and their summed algorithmic flops.
extern "C" int pb_profile_end(long long* gemm_launches, double* gemm_ms, double* gemm_flops) {
    pb::g_profiling.store(0);
    std::lock_guard<std::mutex> lock(pb::g_prof_mu);
    double ms = 0, fl = 0;
    for (auto& r : pb::g_gemm) {
        float t = 0;
        if (cudaEventElapsedTime(&t, r.e0, r.e1) == cudaSuccess) { ms += t; fl += r.flops; }
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
    }
    if (gemm_launches) *gemm_launches = (long long)pb::g_gemm.size();
    if (gemm_ms) *gemm_ms = ms;
    if (gemm_flops) *gemm_flops = fl;
    pb::g_gemm.clear();
    return PB_OK;
}

extern "C" int pb_version(void) { return 100; }

extern "C" const char* pb_last_error(void) { return pb::g_error; }
