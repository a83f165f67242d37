// Dependent-issue latency of FP64 FMA / ADD and of a shared-memory load on this GPU, with 1..16 warps per CTA running
// the same chain (the few-signal kernels in csrc/solve_persist.cu and csrc/solve_small.cu are bounded by these numbers).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_latency fp64_latency.cu && ./fp64_latency
#include <cstdio>
#include <cuda_runtime.h>

constexpr int CHAIN = 4096;

template <int ILP>
__global__ void dfma_chain(double* out, long long* cycles, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int c = 0; c < ILP; ++c) x[c] = threadIdx.x + c;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < CHAIN; ++i) {
#pragma unroll
        for (int c = 0; c < ILP; ++c) x[c] = fma(x[c], a, b);
    }
    const long long t1 = clock64();
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < ILP; ++c) s += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// arg-max chain with FP64 compares (DSETP) vs the same order on the bit patterns (non-negative doubles order like int64)
template <bool BITS>
__global__ void argmax_chain(double* out, long long* cycles, const double* vals) {
    __shared__ double sv[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) sv[i] = vals[i];
    __syncthreads();
    double bv = -1.0;
    int bi = 0x7fffffff;
    const long long t0 = clock64();
    for (int rep = 0; rep < CHAIN / 256; ++rep) {
#pragma unroll 16
        for (int i = 0; i < 256; ++i) {
            const double v = sv[(i + threadIdx.x) & 255];
            const int idx = i + rep;
            bool better;
            if (BITS) {
                const long long kv = __double_as_longlong(v), kb = __double_as_longlong(bv);
                better = kv > kb || (kv == kb && idx < bi);
            } else {
                better = v > bv || (v == bv && idx < bi);
            }
            if (better) { bv = v; bi = idx; }
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = bv + bi;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// special-register read in a dependent chain (ptxas rebuilds shared-memory symbol addresses from SR_CgaCtaId)
__global__ void sreg_chain(int* out, long long* cycles) {
    unsigned x = threadIdx.x;
    const long long t0 = clock64();
    for (int i = 0; i < CHAIN; ++i) {
        unsigned r;
        asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
        x = x * 3u + r;
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = (int)x;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

__global__ void lds_chain(int* out, long long* cycles, int stride) {
    __shared__ int next[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) next[i] = (i + stride) & 1023;
    __syncthreads();
    int p = threadIdx.x & 1023;
    const long long t0 = clock64();
    for (int i = 0; i < CHAIN; ++i) p = next[p];
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = p;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
    double* out;
    long long* cyc;
    cudaMalloc(&out, 148 * 1024 * sizeof(double));
    cudaMalloc(&cyc, 148 * sizeof(long long));
    long long h[148];
    printf("{\"chain\": %d", CHAIN);
    for (int warps : {1, 4, 8, 16}) {
        auto run = [&](auto kernel, const char* name, int ilp) {
            kernel<<<148, warps * 32>>>(out, cyc, 1.0000001, 1e-9);
            cudaDeviceSynchronize();
            kernel<<<148, warps * 32>>>(out, cyc, 1.0000001, 1e-9);
            cudaDeviceSynchronize();
            cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            printf(", \"%s_warps%d\": %.1f", name, warps, (double)h[0] / CHAIN / ilp * ilp);   // cycles per chain step
        };
        run(dfma_chain<1>, "dfma_ilp1_cycles_per_step", 1);
        run(dfma_chain<2>, "dfma_ilp2_cycles_per_step", 2);
        run(dfma_chain<4>, "dfma_ilp4_cycles_per_step", 4);
        {
            static double* vals = nullptr;
            if (!vals) {
                double hv[256];
                for (int i = 0; i < 256; ++i) hv[i] = (double)((i * 2654435761u) % 1000) / 7.0;
                cudaMalloc(&vals, sizeof(hv));
                cudaMemcpy(vals, hv, sizeof(hv), cudaMemcpyHostToDevice);
            }
            for (int bits = 0; bits < 2; ++bits) {
                for (int rep = 0; rep < 2; ++rep) {
                    if (bits) argmax_chain<true><<<148, warps * 32>>>(out, cyc, vals); else argmax_chain<false><<<148, warps * 32>>>(out, cyc, vals);
                    cudaDeviceSynchronize();
                }
                cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
                printf(", \"argmax_%s_cycles_per_step_warps%d\": %.1f", bits ? "int64_bits" : "fp64_compare", warps, (double)h[0] / CHAIN);
            }
        }
        sreg_chain<<<148, warps * 32>>>((int*)out, cyc);
        cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        printf(", \"sreg_cluster_ctarank_cycles_per_read_warps%d\": %.1f", warps, (double)h[0] / CHAIN);
        lds_chain<<<148, warps * 32>>>((int*)out, cyc, 33);
        cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        printf(", \"lds_cycles_per_load_warps%d\": %.1f", warps, (double)h[0] / CHAIN);
    }
    printf("}\n");
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { fprintf(stderr, "CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
