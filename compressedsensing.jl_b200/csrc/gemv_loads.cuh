// Streaming dictionary loads shared by the GEMV-type kernels (corr_gemv.cu, solve_persist.cu).
#pragma once
#include "common.cuh"

namespace csb {

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
// Loads carrying an explicit L2 cache policy (createpolicy descriptor).  Measured on the single-signal 1024 x 8192
// solve (tools/keep_sweep.sh, profiles/): the same streaming loads run the pass in 9.4 us with an evict_normal
// descriptor against 15.1 us without one (FP64; 8.2 vs 11.0 us FP32) although both miss the L2 alike -- the policy-less
// .nc/no_allocate load is the slow path on this part.  Marking a fraction of the dictionary evict_last (to pin it
// across passes) did not produce hits and was slower for every fraction below 1.
enum : int { L2POL_NONE = 0, L2POL_NORMAL = 1, L2POL_FIRST = 2, L2POL_LAST = 3 };
__device__ __forceinline__ float4 ldg_stream(const float4* p, unsigned long long pol) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ double2 ldg_stream(const double2* p, unsigned long long pol) {
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;" : "=d"(r.x), "=d"(r.y) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ unsigned long long l2_policy(int which) {
    unsigned long long pol;
    if (which == L2POL_FIRST) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    else if (which == L2POL_LAST) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void fma_vec(double& acc, const float4& a, const double* r) {
    acc = fma((double)a.x, r[0], acc); acc = fma((double)a.y, r[1], acc);
    acc = fma((double)a.z, r[2], acc); acc = fma((double)a.w, r[3], acc);
}
__device__ __forceinline__ void fma_vec(double& acc, const double2& a, const double* r) {
    acc = fma(a.x, r[0], acc); acc = fma(a.y, r[1], acc);
}

}  // namespace csb
