// FP64 peak microbenchmark for B200 (sm_100a): the denominator of the batched-path roofline.
//
// MEASURED_PEAKS.json (driver-written) records HBM copy bandwidth and bf16 GEMM throughput only;
// the batched correlation C = A'R is an FP64 contraction, and tcgen05 has no f64 kind, so the FP64
// tensor path on sm_100a is `mma.sync.m8n8k4.f64` (SASS DMMA.8x8x4 -- every wider f64 mma shape
// is decomposed into it by ptxas, checked with cuobjdump).  This tool measures, on the box:
//   1. DMMA.8x8x4 issue-bound throughput (register operands only, all SMs, several occupancies),
//   2. DFMA (vector FP64 FMA) throughput,
//   3. cuBLAS DGEMM on the C2 correlation shape (N=8192 atoms, B signals, K=1024), TN layout,
// both as a burst (best of 5 short runs) and sustained (back-to-back for ~3 s).
// Output: one JSON object on stdout.  Build: see tools/Makefile.  Not part of the product library.
#include <cublas_v2.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

template <int NACC>
__global__ void __launch_bounds__(1024) dmma_kernel(double* out, int iters, double seed) {
    double a = seed + threadIdx.x * 1e-9, b = seed * 0.5 + threadIdx.x * 1e-9;
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void __launch_bounds__(1024) dfma_kernel(double* out, int iters, double seed) {
    double a = seed + threadIdx.x * 1e-9, b = seed * 0.5;
    double c[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i] = i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) c[i] = fma(a, c[i], b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// Register-tiled DMMA (MI x NJ accumulators, MI a-fragments, NJ b-fragments: the operand pattern of a real
// warp tile).  Block 0 also reports its clock64() span so that flop/clk/SM is known independently of DVFS.
template <int MI, int NJ>
__global__ void __launch_bounds__(1024) dmma_tile_kernel(double* out, long long* cyc, int iters, double seed) {
    double a[MI], b[NJ], c[MI][NJ][2];
#pragma unroll
    for (int i = 0; i < MI; ++i) a[i] = seed + (threadIdx.x + i) * 1e-9;
#pragma unroll
    for (int j = 0; j < NJ; ++j) b[j] = seed * 0.5 + (threadIdx.x + 7 * j) * 1e-9;
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) { c[i][j][0] = 0.0; c[i][j][1] = 0.0; }
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[i][j][0]), "+d"(c[i][j][1]) : "d"(a[i]), "d"(b[j]));
    }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
    double s = 0;
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) s += c[i][j][0] + c[i][j][1];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static void time_it(F&& launch, double flop_per_launch, double* burst, double* sustained, double sustain_s) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int w = 0; w < 3; ++w) launch();
    CK(cudaDeviceSynchronize());
    double best = 1e30;
    for (int r = 0; r < 5; ++r) {
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = std::min(best, (double)ms);
    }
    *burst = flop_per_launch / (best * 1e-3) / 1e12;
    int n = std::max(1, (int)(sustain_s * 1e3 / best));
    CK(cudaEventRecord(e0));
    for (int r = 0; r < n; ++r) launch();
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    *sustained = flop_per_launch * n / (ms * 1e-3) / 1e12;
}

// --quick [device]: only the register-tiled DMMA measurement that defines the roofline denominator (about half a
// second of GPU time) -- bench.py runs it inside its own clock-sampling window so that numerator and denominator come
// from the same box, lease and clocks.  Prints {"peak_tflops": sustained, "burst_tflops": ..., "flop_per_clk_per_sm": ...}.
static int quick(int dev) {
    CK(cudaSetDevice(dev));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, dev));
    const int sms = prop.multiProcessorCount, iters = 4096, threads = 256;
    double* out; CK(cudaMalloc(&out, (size_t)sms * 1024 * sizeof(double)));
    long long* cyc; CK(cudaMalloc(&cyc, 8));
    const double flop = (double)sms * (threads / 32) * (double)iters * 8 * 4 * 512.0;
    double bu, su;
    time_it([&] { dmma_tile_kernel<8, 4><<<sms, threads>>>(out, cyc, iters, 1.0); }, flop, &bu, &su, 0.4);
    long long h; CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    const double fpc = (double)(threads / 32) * iters * 8 * 4 * 512.0 / (double)h;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"device\": %d, \"peak_tflops\": %.3f, \"burst_tflops\": %.3f, \"flop_per_clk_per_sm\": %.1f, "
           "\"kernel\": \"DMMA.8x8x4 register-tiled 8x4, 256 threads x 1 CTA/SM, sustained 0.4 s\"}\n", prop.name, sms, dev, su, bu, fpc);
    return 0;
}

int main(int argc, char** argv) {
    if (argc > 1 && !strcmp(argv[1], "--quick")) return quick(argc > 2 ? atoi(argv[2]) : 0);
    double sustain_s = argc > 1 ? atof(argv[1]) : 3.0;
    int dev = 0; CK(cudaSetDevice(dev));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, dev));
    int sms = prop.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, (size_t)sms * 8 * 1024 * sizeof(double)));
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"cc\": \"%d.%d\"", prop.name, sms, prop.major, prop.minor);

    // 1. DMMA: threads/CTA x CTAs/SM variants; 16 independent accumulators per warp.
    const int iters = 4096;
    struct Cfg { int threads; int ctas; } cfgs[] = {{128, 1}, {256, 1}, {512, 1}, {1024, 1}, {256, 2}, {1024, 2}};
    printf(", \"dmma_884\": [");
    bool first = true;
    for (auto c : cfgs) {
        double flop = (double)sms * c.ctas * (c.threads / 32) * (double)iters * 16 * 512.0;
        double bu, su;
        time_it([&] { dmma_kernel<16><<<sms * c.ctas, c.threads>>>(out, iters, 1.0); }, flop, &bu, &su, sustain_s / 2);
        printf("%s{\"threads\": %d, \"ctas_per_sm\": %d, \"burst_tflops\": %.2f, \"sustained_tflops\": %.2f}",
               first ? "" : ", ", c.threads, c.ctas, bu, su);
        first = false;
    }
    printf("]");
    // 1b. register-tiled DMMA with per-SM cycle counts
    {
        long long* cyc; CK(cudaMalloc(&cyc, 8));
        printf(", \"dmma_tile\": [");
        bool f2 = true;
        auto run = [&](auto kern, int mi, int nj, int threads) {
            double flop = (double)sms * (threads / 32) * (double)iters * mi * nj * 512.0;
            double bu, su;
            time_it([&] { kern<<<sms, threads>>>(out, cyc, iters, 1.0); }, flop, &bu, &su, 0.5);
            long long h; CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
            double fpc = (double)(threads / 32) * iters * mi * nj * 512.0 / (double)h;
            printf("%s{\"mi\": %d, \"nj\": %d, \"threads\": %d, \"burst_tflops\": %.2f, \"sustained_tflops\": %.2f, \"flop_per_clk_per_sm\": %.1f}",
                   f2 ? "" : ", ", mi, nj, threads, bu, su, fpc);
            f2 = false;
        };
        run(dmma_tile_kernel<8, 4>, 8, 4, 256);
        run(dmma_tile_kernel<8, 4>, 8, 4, 128);
        run(dmma_tile_kernel<4, 4>, 4, 4, 256);
        run(dmma_tile_kernel<4, 4>, 4, 4, 512);
        run(dmma_tile_kernel<4, 2>, 4, 2, 1024);
        run(dmma_tile_kernel<2, 2>, 2, 2, 1024);
        printf("]");
        CK(cudaFree(cyc));
    }
    // 2. DFMA
    printf(", \"dfma\": [");
    first = true;
    for (auto c : cfgs) {
        double flop = (double)sms * c.ctas * c.threads * (double)iters * 16 * 2.0;
        double bu, su;
        time_it([&] { dfma_kernel<16><<<sms * c.ctas, c.threads>>>(out, iters, 1.0); }, flop, &bu, &su, sustain_s / 2);
        printf("%s{\"threads\": %d, \"ctas_per_sm\": %d, \"burst_tflops\": %.2f, \"sustained_tflops\": %.2f}",
               first ? "" : ", ", c.threads, c.ctas, bu, su);
        first = false;
    }
    printf("]");
    CK(cudaGetLastError());

    // 3. cuBLAS DGEMM yardstick on the C2 correlation shape: C[N x B] = A'[N x K] * R[K x B]
    {
        const int N = 8192, K = 1024, B = 16384;
        double *A, *R, *C;
        CK(cudaMalloc(&A, (size_t)K * N * 8)); CK(cudaMalloc(&R, (size_t)K * B * 8)); CK(cudaMalloc(&C, (size_t)N * B * 8));
        std::vector<double> h((size_t)K * std::max(N, B));
        for (size_t i = 0; i < h.size(); ++i) h[i] = (double)((i * 2654435761u) % 1000) / 1000.0 - 0.5;
        CK(cudaMemcpy(A, h.data(), (size_t)K * N * 8, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(R, h.data(), (size_t)K * B * 8, cudaMemcpyHostToDevice));
        cublasHandle_t hd; if (cublasCreate(&hd) != CUBLAS_STATUS_SUCCESS) { fprintf(stderr, "cublasCreate failed\n"); return 3; }
        double one = 1.0, zero = 0.0, bu, su;
        time_it([&] { cublasDgemm(hd, CUBLAS_OP_T, CUBLAS_OP_N, N, B, K, &one, A, K, R, K, &zero, C, N); },
                2.0 * N * (double)B * K, &bu, &su, sustain_s);
        printf(", \"cublas_dgemm_TN_8192x16384x1024\": {\"burst_tflops\": %.2f, \"sustained_tflops\": %.2f}", bu, su);
        cublasDestroy(hd);
    }
    printf("}\n");
    return 0;
}
