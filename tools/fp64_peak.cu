// FP64 peak probe for B200 (sm_100a): measures the denominators that MEASURED_PEAKS.json lacks.
//   (1) register-resident mma.sync m16n8k8.f64 (SASS DMMA.8x8x4) loop  -> FP64 tensor peak
//   (2) register-resident mma.sync m8n8k4.f64 loop
//   (3) register-resident DFMA loop                                     -> FP64 vector peak
//   (4) cublasDgemm n^3 (NT)                                            -> library GEMM comparator
//   (5) cusolverDnDpotrf                                                -> library Cholesky comparator
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo tools/fp64_peak.cu -o tools/fp64_peak -lcublas -lcusolver
// Output: one JSON object on stdout.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cublas_v2.h>
#include <cusolverDn.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "CUDA %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

template <int NACC>
__global__ void __launch_bounds__(256) dmma_m16n8k8_loop(double* out, int iters) {
    double c[NACC][4];
    double a[4], b[2];
    for (int i = 0; i < 4; ++i) a[i] = 1e-3 * (threadIdx.x + i);
    for (int i = 0; i < 2; ++i) b[i] = 1e-3 * (threadIdx.x + 7 + i);
#pragma unroll
    for (int j = 0; j < NACC; ++j) for (int i = 0; i < 4; ++i) c[j][i] = 0.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < NACC; ++j) {
            asm volatile(
                "mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                : "+d"(c[j][0]), "+d"(c[j][1]), "+d"(c[j][2]), "+d"(c[j][3])
                : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
        }
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < NACC; ++j) for (int i = 0; i < 4; ++i) s += c[j][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void __launch_bounds__(256) dmma_m8n8k4_loop(double* out, int iters) {
    double c[NACC][2];
    double a = 1e-3 * threadIdx.x, b = 1e-3 * (threadIdx.x + 3);
#pragma unroll
    for (int j = 0; j < NACC; ++j) { c[j][0] = 0; c[j][1] = 0; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < NACC; ++j) {
            asm volatile(
                "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(a), "d"(b));
        }
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < NACC; ++j) s += c[j][0] + c[j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void __launch_bounds__(256) dfma_loop(double* out, int iters) {
    double c[NACC];
    double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-9 * threadIdx.x;
#pragma unroll
    for (int j = 0; j < NACC; ++j) c[j] = j;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < NACC; ++j) c[j] = fma(c[j], a, b);
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < NACC; ++j) s += c[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
static float time_ms(F f, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); f();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        f();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

int main(int argc, char** argv) {
    int potrf_max = argc > 1 ? atoi(argv[1]) : 32768;
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 256));
    printf("{\"gpu\": \"%s\", \"sms\": %d", prop.name, sms);

    const int iters = 20000;
    for (int bps = 1; bps <= 4; bps *= 2) {   // CTAs of 256 threads per SM: 8,16,32 warps/SM
        int grid = sms * bps;
        {
            float ms = time_ms([&] { dmma_m16n8k8_loop<8><<<grid, 256>>>(out, iters); }, 3);
            double flops = 2.0 * 16 * 8 * 8 * 8.0 * iters * (double)grid * 8;
            printf(", \"dmma_m16n8k8_tflops_w%d\": %.3f", bps * 8, flops / ms * 1e-9);
        }
        {
            float ms = time_ms([&] { dmma_m8n8k4_loop<8><<<grid, 256>>>(out, iters); }, 3);
            double flops = 2.0 * 8 * 8 * 4 * 8.0 * iters * (double)grid * 8;
            printf(", \"dmma_m8n8k4_tflops_w%d\": %.3f", bps * 8, flops / ms * 1e-9);
        }
        {
            float ms = time_ms([&] { dfma_loop<16><<<grid, 256>>>(out, iters); }, 3);
            double flops = 2.0 * 16.0 * iters * (double)grid * 256;
            printf(", \"dfma_tflops_w%d\": %.3f", bps * 8, flops / ms * 1e-9);
        }
    }
    CK(cudaGetLastError());

    // cuBLAS DGEMM comparator (C = A * B^T, the K-major/K-major shape our syrk uses)
    cublasHandle_t h; cublasCreate(&h);
    for (int n : {4096, 8192}) {
        double *A, *B, *C;
        CK(cudaMalloc(&A, sizeof(double) * n * n)); CK(cudaMalloc(&B, sizeof(double) * n * n)); CK(cudaMalloc(&C, sizeof(double) * n * n));
        CK(cudaMemset(A, 0, sizeof(double) * n * n)); CK(cudaMemset(B, 0, sizeof(double) * n * n));
        double one = 1.0, zero = 0.0;
        float ms = time_ms([&] { cublasDgemm(h, CUBLAS_OP_T, CUBLAS_OP_N, n, n, n, &one, A, n, B, n, &zero, C, n); }, 3);
        printf(", \"cublas_dgemm_tn_%d_tflops\": %.3f", n, 2.0 * n * n * (double)n / ms * 1e-9);
        ms = time_ms([&] { cublasDsyrk(h, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, n, n, &one, A, n, &zero, C, n); }, 3);
        printf(", \"cublas_dsyrk_%d_tflops\": %.3f", n, 1.0 * n * n * (double)n / ms * 1e-9);
        cudaFree(A); cudaFree(B); cudaFree(C);
    }

    // cuSOLVER potrf comparator on a diagonally dominant SPD matrix (64-bit API: N=65536 overflows the int one)
    cusolverDnHandle_t sh; cusolverDnCreate(&sh);
    cusolverDnParams_t params; cusolverDnCreateParams(&params);
    for (int64_t n = 8192; n <= potrf_max; n *= 2) {
        double* A; CK(cudaMalloc(&A, sizeof(double) * (size_t)n * n));
        std::vector<double> diag(n, (double)n);
        size_t wdev = 0, whost = 0;
        cusolverDnXpotrf_bufferSize(sh, params, CUBLAS_FILL_MODE_LOWER, n, CUDA_R_64F, A, n, CUDA_R_64F, &wdev, &whost);
        void* work; CK(cudaMalloc(&work, wdev ? wdev : 8));
        std::vector<char> hwork(whost ? whost : 8);
        int* info; CK(cudaMalloc(&info, sizeof(int)));
        float best = 1e30f;
        for (int r = 0; r < 2; ++r) {
            CK(cudaMemset(A, 0, sizeof(double) * (size_t)n * n));
            CK(cudaMemcpy2D(A, sizeof(double) * (n + 1), diag.data(), sizeof(double), sizeof(double), n, cudaMemcpyHostToDevice));
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaEventRecord(e0);
            cusolverDnXpotrf(sh, params, CUBLAS_FILL_MODE_LOWER, n, CUDA_R_64F, A, n, CUDA_R_64F, work, wdev, hwork.data(), whost, info);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
        printf(", \"cusolver_dpotrf_%lld_tflops\": %.3f, \"cusolver_dpotrf_%lld_ms\": %.2f", (long long)n, (double)n * n * n / 3.0 / best * 1e-9, (long long)n, best);
        cudaFree(A); cudaFree(work); cudaFree(info);
    }
    printf("}\n");
    return 0;
}
