// Single-signal (or few-signal) residual correlation c = A'r with the |c| top-s selection fused
// in: the bandwidth-bound GEMV path of subsystems (1)+(2), FP64 and FP32 dictionaries.
//
// Replaces `mul!(P.Ar, P.A', P.r)` -> BLAS gemv('T'), `@. P.Ar = abs(P.Ar)`, `argmax` /
// `partialsortperm` (/root/reference/src/matchingpursuit.jl:181-193) for one right-hand side.
//
// Layout / mapping.  A is column-major, so a CTA's atom range is one contiguous run of the dictionary: its 8 warps
// stream it once, a warp taking one register-blocked group of 4 columns at a time (4 independent 16 B loads in
// flight per lane per step, the residual chunk loaded once from shared memory and reused by the 4 columns).  Every dot
// product is reduced wholly inside one warp in a fixed order and accumulated in FP64 even for an
// FP32 dictionary (float x float is exact in double; the FP64 pipe is idle on a bandwidth-bound
// kernel), so c_j does not depend on how atoms are partitioned over CTAs, shards or GPUs.
// The residual is staged once per CTA in shared memory.
#include "common.cuh"
#include "gemv_loads.cuh"

#include <cstdlib>
#include <cstring>

namespace csb {
namespace {

constexpr int GT = 256;            // threads per CTA, HBM regime (4+ CTAs per SM)
constexpr int GT_L2 = 1024;        // threads per CTA, L2 regime (one CTA per SM: the residual is staged once per SM)

// Work distribution.  The first version gave every 64-atom block its own CTA: 2048 CTAs of 155 us each on 444 CTA
// slots is 4.6 waves, i.e. 8 % of the kernel ran with a partly empty GPU (measured 0.92 of the HBM peak).  Now the
// grid is ONE wave (SMs x resident CTAs, `corr_gemv_blocks`), CTA c owns the contiguous atom range
// [c Ng / P, (c + 1) Ng / P) in units of 4-column groups (Ng = ceil(N / 4)), and its warps pull groups from a
// shared-memory counter, so every CTA -- and every warp inside it -- finishes within one group of the others.
// One candidate record set per (CTA, signal): P = gridDim.x replaces the per-64-atom blocks of the DMMA path.
constexpr int GEMV_MAX_RANGE = 2048;   // atoms per CTA the top-s (s > 1) scratch can hold

// Two instantiations per element type.
//  * HBM regime <UNR 2, NT 256, CG 4>: dictionaries streamed from HBM; 8 x 16 B in flight per lane at 4+ CTAs per SM
//    saturate the memory system (measured 1.0 of the HBM copy peak).
//  * L2 regime <UNR 4, NT 1024, CG 2>: dictionaries that fit the L2.  The pass lasts ~10 us and a 1024 x 8192 dictionary
//    has only ~14 four-column groups per SM, so the same kernel ran latency-bound (3.6 TB/s).  Two-column groups and
//    32 warps per CTA put twice as many warps -- and twice the bytes in flight -- on every SM, one CTA per SM stages
//    the residual once per SM, and the grid is a whole multiple of the SM count (a second CTA on some SMs only would be
//    a 2x tail on so short a kernel).
// Each column has its own accumulator and is reduced inside one warp in the same order in both, so c_j is bit-identical
// whichever instantiation (and whichever shard / GPU count) computes it.
template <typename T, int UNR, int NT, int CG, bool HINT>
__global__ void __launch_bounds__(NT) corr_gemv_kernel(CorrArgs a) {
    constexpr int GW = NT / 32;
    unsigned long long pol = 0;                    // HINT: dictionary loads carry an L2 cache policy (compile-time: a
    if (HINT) pol = l2_policy(a.l2_policy);        // run-time branch in the streaming loop cost the HBM pass 13 %)
    using V = typename Vec<T>::type;
    constexpr int W = Vec<T>::W;
    extern __shared__ unsigned char gsm[];
    T* rs = reinterpret_cast<T*>(gsm);                               // [ld] residual
    double* cv = reinterpret_cast<double*>(gsm + (((size_t)a.ld * sizeof(T) + 15) / 16) * 16);   // [range] |c| (S > 1 only)
    __shared__ int next_group;
    __shared__ double red_v[GW];
    __shared__ int red_i[GW];
    const int p = blockIdx.x, sig = blockIdx.y, P = gridDim.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ld = a.ld;
    const T* A = static_cast<const T*>(a.A);
    const T* r = static_cast<const T*>(a.R) + (size_t)sig * ld;
    const long long Ng = (a.N + CG - 1) / CG;
    const int g_lo = (int)(Ng * p / P), g_hi = (int)(Ng * (p + 1) / P);
    for (int row = tid; row < ld; row += NT) rs[row] = r[row];
    if (tid == 0) next_group = g_lo;
    __syncthreads();

    const int nvec = ld / W;                       // ld is a multiple of 16 elements
    double best_v = -1.0;                          // this warp's running (|c|, atom), warp-uniform
    int best_i = INT_MAX;
    for (;;) {
        int grp = 0;
        if (lane == 0) grp = atomicAdd(&next_group, 1);
        grp = __shfl_sync(0xffffffffu, grp, 0);
        if (grp >= g_hi) break;
        const int atom0 = grp * CG;
        double acc[CG];
        const V* col[CG];
#pragma unroll
        for (int c = 0; c < CG; ++c) {
            const int atom = (atom0 + c < a.N) ? atom0 + c : a.N - 1;      // clamp: stay in bounds, result discarded
            col[c] = reinterpret_cast<const V*>(A + (size_t)atom * ld);
            acc[c] = 0.0;
        }
#pragma unroll UNR
        for (int i = lane; i < nvec; i += 32) {
            V x[CG];
#pragma unroll
            for (int c = 0; c < CG; ++c) x[c] = HINT ? ldg_stream(col[c] + i, pol) : ldg_stream(col[c] + i);
            double rr[W];
#pragma unroll
            for (int e = 0; e < W; ++e) rr[e] = (double)rs[i * W + e];
#pragma unroll
            for (int c = 0; c < CG; ++c) fma_vec(acc[c], x[c], rr);
        }
#pragma unroll
        for (int c = 0; c < CG; ++c) {
            double s = acc[c];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
            const double v = fabs(s);
            const int atom = atom0 + c;
            if (atom < a.N) {
                if (a.S == 1) { if (v >= 0.0 && cand_better(v, atom, best_v, best_i)) { best_v = v; best_i = atom; } }   // NaN never wins
                else if (lane == 0) cv[atom - g_lo * CG] = v;
            }
        }
    }

    if (a.S == 1) {
        if (lane == 0) { red_v[warp] = best_v; red_i[warp] = best_i; }
        __syncthreads();
        if (tid == 0) {
            double bv = red_v[0];
            int bi = red_i[0];
#pragma unroll
            for (int w = 1; w < GW; ++w)
                if (cand_better(red_v[w], red_i[w], bv, bi)) { bv = red_v[w]; bi = red_i[w]; }
            const size_t o = (size_t)sig * P + p;
            a.pval[o] = bv;
            a.pidx[o] = (bi == INT_MAX) ? -1 : bi + a.idx_offset;
        }
        return;
    }
    // top-S of this CTA's range: S block-wide argmax rounds with exclusion (value desc, index asc)
    __syncthreads();
    const int base = g_lo * CG;
    const int range = min(g_hi * CG, a.N) - base;
    double pv = 0.0;
    int pi = -1;
    for (int s = 0; s < a.S; ++s) {
        double bv = -1.0;
        int bi = INT_MAX;
        for (int l = tid; l < range; l += NT) {
            const double v = cv[l];
            const int idx = base + l;
            bool ok = v >= 0.0;                                        // excludes NaN
            if (s > 0) ok = ok && (v < pv || (v == pv && idx > pi));
            if (ok && cand_better(v, idx, bv, bi)) { bv = v; bi = idx; }
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
        for (int w = 1; w < GW; ++w)
            if (cand_better(red_v[w], red_i[w], bv, bi)) { bv = red_v[w]; bi = red_i[w]; }
        pv = bv; pi = bi;
        if (tid == 0) {
            const size_t o = ((size_t)sig * P + p) * a.S + s;
            a.pval[o] = bv;
            a.pidx[o] = (bi == INT_MAX) ? -1 : bi + a.idx_offset;
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

namespace {
size_t gemv_smem_bytes(int ld, bool f32, int S, int range) {
    size_t bytes = (((size_t)ld * (f32 ? 4 : 8) + 15) / 16) * 16;
    if (S > 1) bytes += (size_t)range * sizeof(double);
    return bytes;
}

// Dictionaries up to this size are treated as L2-resident (126 MB L2; the residual, candidates and the update
// kernel's working set share it).
constexpr size_t GEMV_L2_RESIDENT_BYTES = (size_t)96 << 20;
bool gemv_l2_regime(int N, int ld, bool f32) {
    const char* env = getenv("CSB200_GEMV_L2");            // test hook: 0 / 1 force the instantiation
    if (env && (env[0] == '0' || env[0] == '1')) return env[0] == '1';
    return (size_t)N * ld * (f32 ? 4 : 8) <= GEMV_L2_RESIDENT_BYTES;
}

int gemv_l2_policy(bool l2_regime) {
    const char* env = getenv("CSB200_GEMV_POLICY");        // experiment hook: none | normal | first | last
    if (env) return !strcmp(env, "none") ? L2POL_NONE : !strcmp(env, "first") ? L2POL_FIRST : !strcmp(env, "last") ? L2POL_LAST : L2POL_NORMAL;
    return l2_regime ? L2POL_NORMAL : L2POL_NONE;
}

template <typename T, int UNR, int NT, int CG, bool HINT>
cudaError_t gemv_prepare(size_t smem, int* occ) {
    cudaError_t e = cudaFuncSetAttribute(corr_gemv_kernel<T, UNR, NT, CG, HINT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess && occ) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, corr_gemv_kernel<T, UNR, NT, CG, HINT>, NT, smem);
    return e;
}
template <typename T, int UNR, int NT, int CG, bool HINT>
void gemv_launch(dim3 grid, size_t smem, cudaStream_t st, const CorrArgs& b) {
    corr_gemv_kernel<T, UNR, NT, CG, HINT><<<grid, NT, smem, st>>>(b);
}
// the four shapes of the kernel: {HBM, L2 regime} x {loads with / without an L2 policy}, per element type
template <typename T>
cudaError_t gemv_prepare_t(bool l2, bool hint, size_t smem, int* occ) {
    if (l2) return hint ? gemv_prepare<T, 4, GT_L2, 2, true>(smem, occ) : gemv_prepare<T, 4, GT_L2, 2, false>(smem, occ);
    return hint ? gemv_prepare<T, 2, GT, 4, true>(smem, occ) : gemv_prepare<T, 2, GT, 4, false>(smem, occ);
}
template <typename T>
void gemv_launch_t(bool l2, bool hint, dim3 grid, size_t smem, cudaStream_t st, const CorrArgs& b) {
    if (l2) { if (hint) gemv_launch<T, 4, GT_L2, 2, true>(grid, smem, st, b); else gemv_launch<T, 4, GT_L2, 2, false>(grid, smem, st, b); }
    else { if (hint) gemv_launch<T, 2, GT, 4, true>(grid, smem, st, b); else gemv_launch<T, 2, GT, 4, false>(grid, smem, st, b); }
}
cudaError_t gemv_prepare_any(bool f32, bool l2, size_t smem, int* occ) {
    const bool hint = gemv_l2_policy(l2) != L2POL_NONE;
    return f32 ? gemv_prepare_t<float>(l2, hint, smem, occ) : gemv_prepare_t<double>(l2, hint, smem, occ);
}
}  // namespace

// Number of CTAs (= candidate record sets per signal) of the GEMV pass: one wave of resident CTAs, at most
// GEMV_MAX_RANGE atoms per CTA.  HBM regime: at least 8 column groups per CTA.  L2 regime: a whole multiple of the SM
// count, about one column group per warp.
int corr_gemv_blocks(int N, int ld, bool f32, int S, int num_sms) {
    const bool l2 = gemv_l2_regime(N, ld, f32);
    const int cg = l2 ? 2 : 4;
    const long long Ng = ((long long)N + cg - 1) / cg;
    int occ = 1;
    const size_t smem = gemv_smem_bytes(ld, f32, S, GEMV_MAX_RANGE);
    cudaError_t e = gemv_prepare_any(f32, l2, smem, &occ);
    if (e != cudaSuccess || occ < 1) { cudaGetLastError(); occ = 1; }
    long long P = (long long)num_sms * occ;
    if (l2) {
        const long long slots = (long long)num_sms * (GT_L2 / 32);
        long long per_sm = (Ng + slots - 1) / slots;
        if (per_sm > occ) per_sm = occ;
        if (per_sm < 1) per_sm = 1;
        P = (long long)num_sms * per_sm;
        if (P > Ng) P = Ng;
    } else if (P > (Ng + 7) / 8) {
        P = (Ng + 7) / 8;
    }
    const long long pmin = ((long long)N + GEMV_MAX_RANGE - 8) / (GEMV_MAX_RANGE - 7);   // ranges are uneven by < 8 atoms
    if (P < pmin) P = pmin;
    if (P < 1) P = 1;
    return (int)P;
}

cudaError_t launch_corr_gemv(const CorrArgs& a, bool f32, cudaStream_t st) {
    if (a.nsig <= 0 || a.P <= 0) return cudaSuccess;
    const size_t smem = gemv_smem_bytes(a.ld, f32, a.S, GEMV_MAX_RANGE);
    const bool l2 = gemv_l2_regime(a.N, a.ld, f32);
    cudaError_t e = gemv_prepare_any(f32, l2, smem, nullptr);
    if (e != cudaSuccess) return e;
    for (int s0 = 0; s0 < a.nsig; s0 += 65535) {          // gridDim.y limit
        CorrArgs b = a;
        const int ns = a.nsig - s0 < 65535 ? a.nsig - s0 : 65535;
        b.nsig = ns;
        b.R = static_cast<const char*>(a.R) + (size_t)s0 * a.ld * (f32 ? 4 : 8);
        b.pval = a.pval + (size_t)s0 * a.P * a.S;
        b.pidx = a.pidx + (size_t)s0 * a.P * a.S;
        b.l2_policy = gemv_l2_policy(l2);
        dim3 grid(a.P, ns);
        if (f32) gemv_launch_t<float>(l2, b.l2_policy != L2POL_NONE, grid, smem, st, b);
        else gemv_launch_t<double>(l2, b.l2_policy != L2POL_NONE, grid, smem, st, b);
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
