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
struct GemmRec { cudaEvent_t e0, e1; double flops; long long launches; int kind; };
static long long g_int8_launches = 0;      // totals of the INT8 (ozaki.cu) launches of the last profiled region
static double g_int8_ms = 0, g_int8_flops = 0;
static std::vector<GemmRec> g_gemm;
static thread_local CaptureTally* tl_tally = nullptr;

void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
void note_launches(long long n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
long long launch_count() { return g_launches.load(std::memory_order_relaxed); }
bool profiling_enabled() { return g_profiling.load(std::memory_order_relaxed) != 0; }
void profile_gemm(cudaEvent_t e0, cudaEvent_t e1, double flops, long long launches, int kind) {
    std::lock_guard<std::mutex> lock(g_prof_mu);
    g_gemm.push_back({e0, e1, flops, launches, kind});
}
void set_capture_tally(CaptureTally* t) { tl_tally = t; }
CaptureTally* capture_tally() { return tl_tally; }

// Tunables: immutable defaults + a per-thread pointer to the options of the driver call in flight.
static pb_options default_options() {
    pb_options o;
    o.laplace_pcg_min_n = 24576;       // PCG Newton steps pay off once potrf >> trsv
    o.laplace_nystrom_rank = -1;       // automatic (n/16 clamped to [256, 4096])
    o.laplace_cg_tol = 1e-2;           // CG Newton solves: error of the step <= this * Newton tolerance
    o.negative_curvature_tol = 1e-6;
    o.potrf_block = 0;
    o.potrf_lookahead = 1;
    o.potrf_graph = 0;                 // measured slower than eager issue on B200: opt-in
    o.dist_block = 0;
    o.potrf_ozaki = -1;                // auto: INT8 tensor-core contractions for n >= 8192 (1.5 - 1.6 x the DMMA factorisation, profiles/r02_ozaki_bench_*.json)
    o.ozaki_tile = 0;
    return o;
}
static const pb_options g_defaults = default_options();
static thread_local const pb_options* tl_options = nullptr;

const pb_options& opts() { return tl_options ? *tl_options : g_defaults; }
OptScope::OptScope(const pb_options* o) : prev(tl_options) { if (o) tl_options = o; }
OptScope::~OptScope() { tl_options = prev; }

}  // namespace pb

extern "C" int pb_options_default(pb_options* options) {
    PB_CHECK(options != nullptr, PB_ERR_INVALID, "options_default: null argument");
    *options = pb::g_defaults;
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
    double ms = 0, fl = 0, ms8 = 0, fl8 = 0;
    long long launches = 0, launches8 = 0;
    for (auto& r : pb::g_gemm) {
        float t = 0;
        if (cudaEventElapsedTime(&t, r.e0, r.e1) == cudaSuccess) {
            if (r.kind == 1) { ms8 += t; fl8 += r.flops; launches8 += r.launches; }
            else { ms += t; fl += r.flops; launches += r.launches; }
        }
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
    }
    pb::g_int8_launches = launches8; pb::g_int8_ms = ms8; pb::g_int8_flops = fl8;
    if (gemm_launches) *gemm_launches = launches;
    if (gemm_ms) *gemm_ms = ms;
    if (gemm_flops) *gemm_flops = fl;
    pb::g_gemm.clear();
    return PB_OK;
}

// The INT8-sliced contractions (oz_gemm_kernel) of the region closed by the last pb_profile_end: launches, summed
// CUDA-event duration (ms) and summed FP64-EQUIVALENT flops (2 M N K; each is 28 int8 GEMMs of that shape).
extern "C" int pb_profile_int8(long long* launches, double* ms, double* fp64_equivalent_flops) {
    std::lock_guard<std::mutex> lock(pb::g_prof_mu);
    if (launches) *launches = pb::g_int8_launches;
    if (ms) *ms = pb::g_int8_ms;
    if (fp64_equivalent_flops) *fp64_equivalent_flops = pb::g_int8_flops;
    return PB_OK;
}

// Register-resident mma.sync.m16n8k8.f64 loop (SASS: DMMA.8x8x4): the FP64 tensor-core peak of THIS device, the
// denominator of every "fraction of FP64 tensor peak" bench.py prints (MEASURED_PEAKS.json carries no FP64 entry).
__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters) {
    double c[8][4];
    double a[4], b[2];
    for (int i = 0; i < 4; ++i) a[i] = 1e-3 * (threadIdx.x + i);
    for (int i = 0; i < 2; ++i) b[i] = 1e-3 * (threadIdx.x + 7 + i);
#pragma unroll
    for (int j = 0; j < 8; ++j)
        for (int i = 0; i < 4; ++i) c[j][i] = 0.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            asm volatile(
                "mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                : "+d"(c[j][0]), "+d"(c[j][1]), "+d"(c[j][2]), "+d"(c[j][3])
                : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
        }
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j)
        for (int i = 0; i < 4; ++i) s += c[j][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

extern "C" int pb_measure_fp64_tensor_peak(double* tflops_host) {
    PB_CHECK(tflops_host != nullptr, PB_ERR_INVALID, "measure_fp64_tensor_peak: null argument");
    const int grid = pb::num_sms() * 4, iters = 20000;
    double* out = nullptr;
    PB_CUDA(cudaMalloc(&out, sizeof(double) * grid * 256));
    cudaEvent_t e0, e1;
    PB_CUDA(cudaEventCreate(&e0));
    PB_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {                 // first two are warm-up
        PB_CUDA(cudaEventRecord(e0, 0));
        dmma_peak_kernel<<<grid, 256>>>(out, iters);
        PB_CUDA(cudaEventRecord(e1, 0));
        PB_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        PB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (r >= 2 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    *tflops_host = 2.0 * 16 * 8 * 8 * 8.0 * iters * (double)grid * 8 / best * 1e-9;
    return PB_OK;
}

extern "C" int pb_version(void) { return 200; }

extern "C" const char* pb_last_error(void) { return pb::g_error; }
