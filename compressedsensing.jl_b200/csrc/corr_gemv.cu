// Single-signal (or few-signal) residual correlation c = A'r with the |c| top-s selection fused
// in: the bandwidth-bound GEMV path of subsystems (1)+(2), FP64 and FP32 dictionaries.
//
// Replaces `mul!(P.Ar, P.A', P.r)` -> BLAS gemv('T'), `@. P.Ar = abs(P.Ar)`, `argmax` /
// `partialsortperm` (/root/reference/src/matchingpursuit.jl:181-193) for one right-hand side.
//
// Layout / mapping.  A is column-major, so the 64 atoms of one atom block are one contiguous
// run of 64*ld elements: a CTA (8 warps) streams that run once, warp w owning atoms 8w..8w+7 in
// two register-blocked groups of 4 columns (4 independent 16 B loads in flight per lane per step,
// the residual chunk loaded once from shared memory and reused by the 4 columns).  Every dot
// product is reduced wholly inside one warp in a fixed order and accumulated in FP64 even for an
// FP32 dictionary (float x float is exact in double; the FP64 pipe is idle on a bandwidth-bound
// kernel), so c_j does not depend on how atoms are partitioned over CTAs, shards or GPUs.
// The residual is staged once per CTA in shared memory as doubles.
#include "common.cuh"

namespace csb {
namespace {

constexpr int GT = 256;            // threads per CTA (8 warps x 8 atoms = PBLK)
constexpr int GW = GT / 32;
constexpr int CPW = PBLK / GW;     // atoms per warp (8)

template <typename T> struct Vec;
template <> struct Vec<float> { using type = float4; static constexpr int W = 4; };
template <> struct Vec<double> { using type = double2; static constexpr int W = 2; };

__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ double2 ldg_stream(const double2* p) {
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ void fma_vec(double& acc, const float4& a, const double* r) {
    acc = fma((double)a.x, r[0], acc); acc = fma((double)a.y, r[1], acc);
    acc = fma((double)a.z, r[2], acc); acc = fma((double)a.w, r[3], acc);
}
__device__ __forceinline__ void fma_vec(double& acc, const double2& a, const double* r) {
    acc = fma(a.x, r[0], acc); acc = fma(a.y, r[1], acc);
}

template <typename T>
__global__ void __launch_bounds__(GT) corr_gemv_kernel(CorrArgs a) {
    using V = typename Vec<T>::type;
    constexpr int W = Vec<T>::W;
    extern __shared__ double rs[];                 // [ld] residual as doubles
    __shared__ double cv[PBLK];
    const int p = blockIdx.x, sig = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ld = a.ld;
    const T* A = static_cast<const T*>(a.A);
    const T* r = static_cast<const T*>(a.R) + (size_t)sig * ld;
    for (int row = tid; row < ld; row += GT) rs[row] = (double)r[row];
    __syncthreads();

    const int nvec = ld / W;                       // ld is a multiple of 16 elements
#pragma unroll
    for (int grp = 0; grp < CPW / 4; ++grp) {
        const int local0 = warp * CPW + grp * 4;
        const int atom0 = p * PBLK + local0;
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        const V* col[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int atom = (atom0 + c < a.N) ? atom0 + c : a.N - 1;      // clamp: stay in bounds, result discarded
            col[c] = reinterpret_cast<const V*>(A + (size_t)atom * ld);
        }
#pragma unroll 2
        for (int i = lane; i < nvec; i += 32) {
            V x[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) x[c] = ldg_stream(col[c] + i);
            double rr[W];
#pragma unroll
            for (int e = 0; e < W; ++e) rr[e] = rs[i * W + e];
#pragma unroll
            for (int c = 0; c < 4; ++c) fma_vec(acc[c], x[c], rr);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            double s = acc[c];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
            if (lane == 0) cv[local0 + c] = (atom0 + c < a.N) ? fabs(s) : -1.0;
        }
    }
    __syncthreads();

    if (warp == 0) {                               // top-S of the block's 64 |c| values
        const int base = p * PBLK;
        double pv = 0.0;
        int pi = -1;
        for (int s = 0; s < a.S; ++s) {
            double bv = -1.0;
            int bi = INT_MAX;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int local = lane + 32 * e;
                const double v = cv[local];
                const int idx = base + local;
                bool ok = v >= 0.0;                                        // excludes out-of-range atoms and NaN
                if (s > 0) ok = ok && (v < pv || (v == pv && idx > pi));
                if (ok && cand_better(v, idx, bv, bi)) { bv = v; bi = idx; }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, bv, off);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                if (cand_better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
            }
            pv = bv; pi = bi;
            if (lane == 0) {
                const size_t o = ((size_t)sig * a.P + p) * a.S + s;
                a.pval[o] = bv;
                a.pidx[o] = (bi == INT_MAX) ? -1 : bi + a.idx_offset;
            }
        }
    }
}

// Reference-on-device kernel for tests: one thread per dot product, sequential FP64 sum in row
// order, then a serial selection.  Slow and obviously correct; never used by a solve.
template <typename T>
__global__ void __launch_bounds__(PBLK) corr_naive_kernel(CorrArgs a) {
    __shared__ double cv[PBLK];
    const int p = blockIdx.x, sig = blockIdx.y, local = threadIdx.x;
    const int atom = p * PBLK + local;
    const T* A = static_cast<const T*>(a.A);
    const T* r = static_cast<const T*>(a.R) + (size_t)sig * a.ld;
    double s = -1.0;
    if (atom < a.N) {
        double acc = 0.0;
        const T* col = A + (size_t)atom * a.ld;
        for (int row = 0; row < a.M; ++row) acc = fma((double)col[row], (double)r[row], acc);
        s = fabs(acc);
    }
    cv[local] = s;
    __syncthreads();
    if (local == 0) {
        double pv = 0.0;
        int pi = -1;
        for (int k = 0; k < a.S; ++k) {
            double bv = -1.0;
            int bi = INT_MAX;
            for (int l = 0; l < PBLK; ++l) {
                const double v = cv[l];
                const int idx = p * PBLK + l;
                bool ok = v >= 0.0;
                if (k > 0) ok = ok && (v < pv || (v == pv && idx > pi));
                if (ok && cand_better(v, idx, bv, bi)) { bv = v; bi = idx; }
            }
            pv = bv; pi = bi;
            const size_t o = ((size_t)sig * a.P + p) * a.S + k;
            a.pval[o] = bv;
            a.pidx[o] = (bi == INT_MAX) ? -1 : bi + a.idx_offset;
        }
    }
}

}  // namespace

cudaError_t launch_corr_gemv(const CorrArgs& a, bool f32, cudaStream_t st) {
    if (a.nsig <= 0 || a.P <= 0) return cudaSuccess;
    const size_t smem = (size_t)a.ld * sizeof(double);
    cudaError_t e;
    for (int s0 = 0; s0 < a.nsig; s0 += 65535) {          // gridDim.y limit
        CorrArgs b = a;
        const int ns = a.nsig - s0 < 65535 ? a.nsig - s0 : 65535;
        b.nsig = ns;
        b.R = static_cast<const char*>(a.R) + (size_t)s0 * a.ld * (f32 ? 4 : 8);
        b.pval = a.pval + (size_t)s0 * a.P * a.S;
        b.pidx = a.pidx + (size_t)s0 * a.P * a.S;
        dim3 grid(a.P, ns);
        if (f32) {
            e = cudaFuncSetAttribute(corr_gemv_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            corr_gemv_kernel<float><<<grid, GT, smem, st>>>(b);
        } else {
            e = cudaFuncSetAttribute(corr_gemv_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            corr_gemv_kernel<double><<<grid, GT, smem, st>>>(b);
        }
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

cudaError_t launch_corr_naive(const CorrArgs& a, bool f32, cudaStream_t st) {
    if (a.nsig <= 0 || a.P <= 0) return cudaSuccess;
    for (int s0 = 0; s0 < a.nsig; s0 += 65535) {
        CorrArgs b = a;
        const int ns = a.nsig - s0 < 65535 ? a.nsig - s0 : 65535;
        b.nsig = ns;
        b.R = static_cast<const char*>(a.R) + (size_t)s0 * a.ld * (f32 ? 4 : 8);
        b.pval = a.pval + (size_t)s0 * a.P * a.S;
        b.pidx = a.pidx + (size_t)s0 * a.P * a.S;
        dim3 grid(a.P, ns);
        if (f32) corr_naive_kernel<float><<<grid, PBLK, 0, st>>>(b);
        else corr_naive_kernel<double><<<grid, PBLK, 0, st>>>(b);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace csb
