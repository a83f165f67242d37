// Device helpers shared by the per-signal update kernels (update.cu, update_cluster.cu).
#pragma once
#include "common.cuh"

namespace csb {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}
// Deterministic block sum (fixed order); every thread receives the result.
template <int NT>
__device__ __forceinline__ double block_sum(double v, double* red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) s += red[w];
    return s;
}

// Global top-`take` over this signal's P*S per-block candidates -> s_cand[0..take) (atom or -1).
template <int NT>
__device__ void select_candidates(const double* __restrict__ pv, const int* __restrict__ pi, int count, int take,
                                  int* s_cand, double* s_cval, double* red_v, int* red_i) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double prev_v = 0.0;
    int prev_i = -1;
    for (int round = 0; round < take; ++round) {
        double bv = -1.0;
        int bi = INT_MAX;
        for (int c = tid; c < count; c += NT) {
            const double v = pv[c];
            const int i = pi[c];
            if (i < 0) continue;
            const bool ok = (round == 0) || (v < prev_v) || (v == prev_v && i > prev_i);
            if (ok && cand_better(v, i, bv, bi)) { bv = v; bi = i; }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, bv, off);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
            if (cand_better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
        }
        __syncthreads();
        if (lane == 0) { red_v[warp] = bv; red_i[warp] = bi; }
        __syncthreads();
        bv = red_v[0]; bi = red_i[0];
#pragma unroll
        for (int w = 1; w < NT / 32; ++w)
            if (cand_better(red_v[w], red_i[w], bv, bi)) { bv = red_v[w]; bi = red_i[w]; }
        prev_v = bv; prev_i = bi;
        if (tid == 0) { s_cand[round] = (bi == INT_MAX) ? -1 : bi; s_cval[round] = bv; }
        if (bi == INT_MAX) {            // candidates exhausted: pad the rest
            for (int r2 = round + 1 + tid; r2 < take; r2 += NT) s_cand[r2] = -1;
            break;
        }
    }
    __syncthreads();
}


}  // namespace csb
