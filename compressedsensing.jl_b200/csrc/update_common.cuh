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
    // pairwise tree over the warp partials: a dependent FP64 add costs ~50 cycles here, and a CTA that shares its SM with
    // the correlation pass (api.cu run_omp_split) has no other warps to hide a serial chain behind
    double p[NT / 32];
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) p[w] = red[w];
#pragma unroll
    for (int h = NT / 64; h > 0; h >>= 1)
#pragma unroll
        for (int w = 0; w < h; ++w) p[w] += p[w + h];
    return p[0];
}

// Two sums in one reduction (one barrier pair, interleaved shuffles); red: [2 * NT / 32].
template <int NT>
__device__ __forceinline__ void block_sum2(double& a, double& b, double* red) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, off);
        b += __shfl_xor_sync(0xffffffffu, b, off);
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = a; red[NT / 32 + (threadIdx.x >> 5)] = b; }
    __syncthreads();
    double p[NT / 32], q[NT / 32];
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) { p[w] = red[w]; q[w] = red[NT / 32 + w]; }
#pragma unroll
    for (int h = NT / 64; h > 0; h >>= 1)
#pragma unroll
        for (int w = 0; w < h; ++w) { p[w] += p[w + h]; q[w] += q[w + h]; }
    a = p[0]; b = q[0];
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


// Global top-`take` of one signal's DENSE |c| row (v[0..n), atom index = position + idx_offset), for the large-S
// paths (sp, oblivious, cumbabel, gomp with many atoms per update) where the correlation pass stores |A'r| itself
// instead of per-block candidates.  MSB-first radix select on the bit patterns (non-negative doubles order like
// unsigned integers) with 11-bit digits: a histogram pass finds the bucket holding the take-th largest value; as soon
// as everything at or above that bucket fits into the `cap` output slots (typically after the exponent pass and one
// mantissa pass) it is collected and a bitonic sort orders it by (value descending, index ascending) -- the order
// select_candidates produces -- and the first `take` entries are the answer.  Only massive exact ties (e.g. an
// all-zero residual) go the whole 64 bits and take the lowest indices by an ordered compaction.
// stage (optional): shared scratch of stage_elems doubles; the row is copied there once when it fits.
// hist: DENSE_HIST ints of shared memory, misc: >= 4 ints.  s_cand / s_cval: cap slots, cap a power of two >= take.
constexpr int DENSE_HIST_BITS = 11;
constexpr int DENSE_HIST = 1 << DENSE_HIST_BITS;

template <int NT>
__device__ void select_dense(const double* __restrict__ v, int n, int idx_offset, int take, int cap, int* s_cand,
                             double* s_cval, int* hist, int* misc, double* stage, int stage_elems) {
    const int tid = threadIdx.x, lane = tid & 31;
    const int want = take;
    if (take > n) take = n;
    const double* src = v;
    if (stage && stage_elems > 0 && n <= stage_elems) {
        for (int c = tid; c < n; c += NT) stage[c] = v[c];
        src = stage;
    }
    __syncthreads();
    unsigned long long prefix = 0, mask = 0;
    int rem = take, above = 0;
    bool fits = false, full = false;
    int shift = 64;
    while (!fits && shift > 0 && take > 0) {
        const int bits = shift >= DENSE_HIST_BITS + 1 ? DENSE_HIST_BITS : shift;      // 63 = 11*5 + 8: the sign bit is never set
        shift = shift == 64 ? 52 : shift - bits;
        const int nb = 1 << bits;
        for (int i = tid; i < nb; i += NT) hist[i] = 0;
        __syncthreads();
        for (int c = tid; c < n; c += NT) {
            const unsigned long long key = (unsigned long long)__double_as_longlong(src[c]);
            if ((key & mask) == prefix) atomicAdd(&hist[(int)((key >> shift) & (unsigned long long)(nb - 1))], 1);
        }
        __syncthreads();
        if (tid < 32) {
            // bins are scanned from the top; lane L owns the descending range [hi - L*per - per + 1, hi - L*per]
            const int per = (nb + 31) / 32, top = nb - 1 - lane * per;
            int mine = 0;
            for (int b = top; b > top - per && b >= 0; --b) mine += hist[b];
            int incl = mine;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) { const int o = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += o; }
            const int excl = incl - mine;
            if (excl < rem && rem <= incl) {
                int cum = excl, b = top;
                for (; b > top - per && b > 0; --b) { if (cum + hist[b] >= rem) break; cum += hist[b]; }
                misc[0] = b; misc[1] = rem - cum; misc[2] = hist[b]; misc[3] = cum;
            }
        }
        __syncthreads();
        prefix |= (unsigned long long)misc[0] << shift;
        mask |= (unsigned long long)(nb - 1) << shift;
        rem = misc[1];                                // still to take from inside the bucket
        above = take - rem;                           // elements strictly above the bucket
        fits = above + misc[2] <= cap;
        full = shift == 0;
        __syncthreads();
    }
    if (tid == 0) misc[3] = 0;
    __syncthreads();
    int count = 0;
    if (fits || take == 0) {
        for (int c = tid; c < n && take > 0; c += NT) {
            const double val = src[c];
            if (((unsigned long long)__double_as_longlong(val) & mask) >= prefix) {
                const int pos = atomicAdd(&misc[3], 1);
                s_cand[pos] = c + idx_offset; s_cval[pos] = val;
            }
        }
        __syncthreads();
        count = misc[3];
    } else {
        // more exact ties at the threshold than output slots: everything above it, then the lowest tied indices
        for (int c = tid; c < n; c += NT) {
            const double val = src[c];
            if ((unsigned long long)__double_as_longlong(val) > prefix) {
                const int pos = atomicAdd(&misc[3], 1);
                s_cand[pos] = c + idx_offset; s_cval[pos] = val;
            }
        }
        const int chunk = (n + NT - 1) / NT, lo = tid * chunk, hi = min(n, lo + chunk);
        int mine = 0;
        for (int c = lo; c < hi; ++c) mine += ((unsigned long long)__double_as_longlong(src[c]) == prefix);
        __syncthreads();
        hist[tid] = mine;
        __syncthreads();
        if (tid == 0) { int run = 0; for (int t = 0; t < NT; ++t) { const int x = hist[t]; hist[t] = run; run += x; } }
        __syncthreads();
        int rank = hist[tid];
        const int base = take - rem;
        for (int c = lo; c < hi && rank < rem; ++c)
            if ((unsigned long long)__double_as_longlong(src[c]) == prefix) { s_cand[base + rank] = c + idx_offset; s_cval[base + rank] = src[c]; ++rank; }
        __syncthreads();
        count = take;
        (void)full;
    }
    int p2 = 1;
    while (p2 < count) p2 <<= 1;
    for (int i = count + tid; i < p2; i += NT) { s_cand[i] = INT_MAX; s_cval[i] = -1.0; }
    __syncthreads();
    for (int k2 = 2; k2 <= p2; k2 <<= 1) {
        for (int j = k2 >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < p2; i += NT) {
                const int o = i ^ j;
                if (o > i) {
                    const double va = s_cval[i], vb = s_cval[o];
                    const int ia = s_cand[i], ib = s_cand[o];
                    const bool first_half = (i & k2) == 0;           // "better first" in ascending halves
                    const bool swap = first_half ? cand_better(vb, ib, va, ia) : cand_better(va, ia, vb, ib);
                    if (swap) { s_cval[i] = vb; s_cand[i] = ib; s_cval[o] = va; s_cand[o] = ia; }
                }
            }
            __syncthreads();
        }
    }
    for (int i = take + tid; i < want; i += NT) s_cand[i] = -1;      // fewer than `want` atoms exist
    __syncthreads();
}

// Small `take` (<= DENSE_SMALL_TAKE, gomp with a few atoms per update): ONE streaming pass over the dense row.  Every
// thread keeps the `take` best of its strided share in registers (sorted, insertion is rare after the first few
// elements), the NT x take survivors go to shared scratch and select_candidates merges them.  The top-`take` of the
// row is contained in the union of the per-thread top-`take` lists, so the result is exact, ties included.
constexpr int DENSE_SMALL_TAKE = 8;

template <int NT, int TK>
__device__ __forceinline__ void select_dense_small_t(const double* __restrict__ v, int n, int idx_offset, int* s_cand,
                                                     double* s_cval, double* red_v, int* red_i, double* scratch) {
    const int tid = threadIdx.x;
    double bv[TK];
    int bi[TK];
#pragma unroll
    for (int q = 0; q < TK; ++q) { bv[q] = -1.0; bi[q] = INT_MAX; }
    constexpr int U = 8;                                             // independent loads in flight per thread
    // The row holds |A'r| (the DMMA pass's EPI_ABS store): no fabs.  A thread meets its columns in ascending order, so a
    // newcomer displaces the worst kept entry only when it is strictly larger (an equal value has the higher index):
    // the common path is one compare per element; NaN never passes it.
    auto offer = [&](double val, int c) {
        if (val > bv[TK - 1]) {                                      // beats the worst entry kept so far
            double cv = val;
            int ci = c;
#pragma unroll
            for (int q = 0; q < TK; ++q) {                           // bubble the newcomer down the sorted list
                if (cand_better(cv, ci, bv[q], bi[q])) {
                    const double tv = bv[q]; const int ti = bi[q];
                    bv[q] = cv; bi[q] = ci; cv = tv; ci = ti;
                }
            }
        }
    };
    const int n_full = n - n % (NT * U);                             // whole chunks: no bounds checks
    for (int c0 = tid; c0 < n_full; c0 += NT * U) {
        double vv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) vv[u] = v[c0 + u * NT];
#pragma unroll
        for (int u = 0; u < U; ++u) offer(vv[u], c0 + u * NT);
    }
    for (int c = n_full + tid; c < n; c += NT) offer(v[c], c);
    double* cval = scratch;                                         // [NT * TK]
    int* cidx = reinterpret_cast<int*>(scratch + (size_t)NT * TK);  // [NT * TK]
#pragma unroll
    for (int q = 0; q < TK; ++q) { cval[tid * TK + q] = bv[q]; cidx[tid * TK + q] = bi[q] == INT_MAX ? -1 : bi[q] + idx_offset; }
    __syncthreads();
    select_candidates<NT>(cval, cidx, NT * TK, TK, s_cand, s_cval, red_v, red_i);
}

template <int NT>
__device__ void select_dense_small(const double* __restrict__ v, int n, int idx_offset, int take, int* s_cand,
                                   double* s_cval, double* red_v, int* red_i, double* scratch) {
    switch (take) {
        case 1: select_dense_small_t<NT, 1>(v, n, idx_offset, s_cand, s_cval, red_v, red_i, scratch); break;
        case 2: select_dense_small_t<NT, 2>(v, n, idx_offset, s_cand, s_cval, red_v, red_i, scratch); break;
        case 3: select_dense_small_t<NT, 3>(v, n, idx_offset, s_cand, s_cval, red_v, red_i, scratch); break;
        case 4: select_dense_small_t<NT, 4>(v, n, idx_offset, s_cand, s_cval, red_v, red_i, scratch); break;
        case 5: select_dense_small_t<NT, 5>(v, n, idx_offset, s_cand, s_cval, red_v, red_i, scratch); break;
        case 6: select_dense_small_t<NT, 6>(v, n, idx_offset, s_cand, s_cval, red_v, red_i, scratch); break;
        case 7: select_dense_small_t<NT, 7>(v, n, idx_offset, s_cand, s_cval, red_v, red_i, scratch); break;
        default: select_dense_small_t<NT, 8>(v, n, idx_offset, s_cand, s_cval, red_v, red_i, scratch); break;
    }
}

// Either form of the candidates a correlation pass left for signal `sig`.
template <int NT>
__device__ __forceinline__ void select_any(const StateArgs& a, int sig, int take, int cap, int* s_cand, double* s_cval,
                                           double* red_v, int* red_i, int* hist, double* stage = nullptr,
                                           int stage_elems = 0) {
    if (a.dense_ld > 0 && take <= DENSE_SMALL_TAKE && stage && stage_elems + DENSE_HIST / 2 >= NT * take * 2) {
        // the scratch handed in as histogram + staging area is one contiguous region starting at `hist`
        select_dense_small<NT>(a.pval + (size_t)sig * a.dense_ld, a.N, a.idx_offset, take, s_cand, s_cval, red_v, red_i,
                               reinterpret_cast<double*>(hist));
    } else if (a.dense_ld > 0) {
        select_dense<NT>(a.pval + (size_t)sig * a.dense_ld, a.N, a.idx_offset, take, cap, s_cand, s_cval, hist, red_i,
                         stage, stage_elems);
    } else {
        const size_t cbase = (size_t)sig * a.P * a.S;
        select_candidates<NT>(a.pval + cbase, a.pidx + cbase, a.P * a.S, take, s_cand, s_cval, red_v, red_i);
    }
}

// 16-byte loads of W consecutive rows of a dictionary column, widened to double: the gather sweeps below are
// latency-bound (a few CTAs per SM, one dependent chain per thread), so bytes in flight per load matter.
template <typename T> struct RowVec;
template <> struct RowVec<double> {
    static constexpr int W = 2;
    using V16 = double2;
    static __device__ __forceinline__ void load(const double* p, double (&o)[2]) {
        const double2 x = *reinterpret_cast<const double2*>(p);
        o[0] = x.x; o[1] = x.y;
    }
};
template <> struct RowVec<float> {
    static constexpr int W = 4;
    using V16 = float4;
    static __device__ __forceinline__ void load(const float* p, double (&o)[4]) {
        const float4 x = *reinterpret_cast<const float4*>(p);
        o[0] = x.x; o[1] = x.y; o[2] = x.z; o[3] = x.w;
    }
};

// Two consecutive rows per load (16 bytes of an FP64 column, 8 of an FP32 one): the block-append sweeps keep
// [atoms][rows] accumulators in registers and have no room for four rows per thread.
template <typename T> struct RowPair;
template <> struct RowPair<double> {
    static __device__ __forceinline__ void load(const double* p, double (&o)[2]) {
        const double2 x = *reinterpret_cast<const double2*>(p);
        o[0] = x.x; o[1] = x.y;
    }
};
template <> struct RowPair<float> {
    static __device__ __forceinline__ void load(const float* p, double (&o)[2]) {
        const float2 x = *reinterpret_cast<const float2*>(p);
        o[0] = x.x; o[1] = x.y;
    }
};

// Shared-memory working set of one signal's pursuit state (pointers into the CTA's shared memory).
template <typename T>
struct PursuitSmem {
    double* v;        // [ld]    working vector
    double* g;        // [kcap]  A_S' v
    double* hh;       // [kcap]  Q'v of the current sweep
    double* ys;       // [kcap]  accumulated R^{-1} Q'a
    double* y;        // [kcap]  R^{-1} Q'v of the current sweep
    double* zs;       // [kcap]  Q'b
    int* ssel;        // [kcap]  support, selection order
    const T** colp;   // [kcap]  columns of the active atoms
    double* Tm;       // inverse factor R^{-1} the mat-vecs read (shared or global), leading dimension ldT
    int ldT;
    double* Tsm;      // shared copy of R^{-1} or nullptr
    double* Tg;       // global copy of R^{-1} (ld = kcap) or nullptr
    int kcap;
    double* red;      // [2 * NT/32] reduction scratch (+ 2 doubles for the one-warp sums of append_atom)
    int illcond = 0;  // set by append_atom when an atom kept less than ILLCOND_RATIO of its squared norm (CTA-uniform)
    double* def_y = nullptr;          // deferred residual sweep (StateArgs::def_*): this signal's slots, nullptr = down-date r here
    double* def_gam = nullptr;
    int* def_t = nullptr;
    bool deferred = false;            // set by append_atom when it left the residual sweep to omp_residual_slice_kernel (CTA-uniform)
    const void* ring_r = nullptr;     // the signal's residual in global memory: rides through the ring behind the last column
    double* ring = nullptr;           // [RING_D][ld] per-thread cp.async ring of active columns (append_atom<.., RING_D>), 16-byte aligned
    unsigned long long l2_keep = 0;   // createpolicy L2::evict_last handle for the dictionary gathers, 0 = no hint
};

// 16-byte asynchronous copies global -> shared (LDGSTS, L1 bypassed).  Every thread copies and later reads ITS OWN rows, so
// cp.async.wait_group is the only synchronisation the ring needs; bytes in flight cost shared memory instead of registers.
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src, unsigned long long pol) {
    const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(dst_smem));
    if (pol) asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "l"(pol) : "memory");
    else asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ unsigned long long l2_evict_last_policy() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
constexpr int RING_MAX_SLOTS = 4;     // row slots per thread of the ring path: ld <= RING_MAX_SLOTS * NT * 2

// x = R^{-1} z through the STORED INVERSE has a forward error of about cond(A_S)^2 eps (x is O(1) while ||R^{-1}|| ||z|| is
// O(cond): the products cancel), where the reference's back substitution on a Givens QR gives cond(A_S) eps.  Measured
// (tools/conditioning_study.py): invisible on Gaussian dictionaries (cond(A_S) ~ 1.4), 4e-9 at cond 2e4, 5e-5 at 2e6.
// A support is marked ill-conditioned when an appended atom keeps less than this fraction of its squared norm after
// orthogonalisation (cond(A_S) >~ 10); its coefficients are then refined, see refine_coefficients.  State flag bit 16.
constexpr double ILLCOND_RATIO = 1e-2;
constexpr int FLAG_ILLCOND = 16;

// Iterative refinement of the least-squares coefficients of an ill-conditioned support (corrected semi-normal
// equations): with the explicit residual rr = b - A_S x,   x += R^{-1} R^{-T} A_S' rr,   twice.  Each step shrinks the
// error by ~cond^2 eps, so two steps reach the cond eps level of a backward-stable QR for cond(A_S) up to ~1e6 (at 1e7
// and beyond the stored inverse itself is too inaccurate to correct with; the flag stays set so the caller can tell).
// x: [t] in shared memory (in / out); uses S.v, S.g, S.hh, S.y as scratch.  b_at(row) as in append_atom.
template <typename T, int NT, typename BAt>
__device__ void refine_coefficients(PursuitSmem<T>& S, int t, int ld, BAt b_at, double* x) {
    constexpr int W = RowVec<T>::W;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __syncthreads();
    for (int rep = 0; rep < 2; ++rep) {
        for (int row = tid * W; row < ld; row += NT * W) {             // rr = b - A_S x
            double acc[W];
#pragma unroll
            for (int e = 0; e < W; ++e) acc[e] = b_at(row + e);
            for (int i = 0; i < t; ++i) {
                double a[W];
                RowVec<T>::load(S.colp[i] + row, a);
                const double xi = x[i];
#pragma unroll
                for (int e = 0; e < W; ++e) acc[e] = fma(-a[e], xi, acc[e]);
            }
#pragma unroll
            for (int e = 0; e < W; ++e) S.v[row + e] = acc[e];
        }
        __syncthreads();
        for (int i = warp; i < t; i += NT / 32) {                      // g = A_S' rr
            const T* ai = S.colp[i];
            double s0 = 0.0, s1 = 0.0;
            for (int row = lane * W; row < ld; row += 32 * W) {
                double a[W];
                RowVec<T>::load(ai + row, a);
#pragma unroll
                for (int e = 0; e < W; e += 2) { s0 = fma(a[e], S.v[row + e], s0); s1 = fma(a[e + 1], S.v[row + e + 1], s1); }
            }
            const double s = warp_sum(s0 + s1);
            if (lane == 0) S.g[i] = s;
        }
        __syncthreads();
        for (int i = tid; i < t; i += NT) {                            // hh = R^{-T} g
            double acc = 0.0;
            for (int l = 0; l <= i; ++l) acc = fma(S.Tm[l + i * S.ldT], S.g[l], acc);
            S.hh[i] = acc;
        }
        __syncthreads();
        for (int i = tid; i < t; i += NT) {                            // x += R^{-1} hh
            double acc = 0.0;
            for (int l = i; l < t; ++l) acc = fma(S.Tm[i + l * S.ldT], S.hh[l], acc);
            x[i] += acc;
        }
        __syncthreads();
    }
}

// `add_column!(AiQR, a, pos)` + the residual part of `ldiv!!` / `residual!` for ONE new atom (reference:
// src/util.jl:118-126, src/matchingpursuit.jl:152-176), by the whole CTA:
//   v = a_j;  (g = A_S'v, hh = R^{-T}g, y = R^{-1}hh, v -= A_S y) once, twice if ||v|| collapsed (DGKS);
//   rho = ||v||;  z_t = v'b / rho;  r <- r - (z_t / rho) v;  R^{-1} gains the column [-y/rho; 1/rho].
// b_at(row) / r_at(row) / r_set(row, val) abstract where the signal and residual live (global or shared).
// Returns 0 when the atom was appended (t is incremented, nr2 = ||r||^2), 1 when it is numerically dependent.
// gcol (optional): column j of the precomputed Gram matrix A'A indexed by LOCAL atom index; when given, the first
// sweep takes g = A_S'a_j from it (t scattered 8-byte loads) instead of gathering the t active atoms.
// RING_D > 0 (FP64 dictionaries, ld <= RING_MAX_SLOTS * NT * 2, S.ring set): the residual sweep of the fast path takes the
// active columns through a RING_D-deep cp.async ring -- the first RING_D columns are requested before anything else, so
// their latency hides behind the Gram look-up and the triangular mat-vecs -- with the same per-row FMA chains (even
// columns / odd columns) as the register path: results are bit-identical.
template <typename T, int NT, int RING_D = 0, typename BAt, typename RAt, typename RSet>
__device__ __forceinline__ int append_atom(PursuitSmem<T>& S, int& t, int j, const T* __restrict__ aj, int ld,
                                           BAt b_at, RAt r_at, RSet r_set, double& nr2,
                                           const double* __restrict__ gcol = nullptr, int idx_offset = 0) {
    constexpr int W = RowVec<T>::W;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    [[maybe_unused]] auto ring_issue = [&](int i) {                // request column i (if there is one) and close its group
        if constexpr (RING_D > 0) {
            if (i <= t) {                                          // pseudo-column t is the residual itself
                const T* col = i < t ? S.colp[i] : static_cast<const T*>(S.ring_r);
                double* dst = S.ring + (size_t)(i % RING_D) * ld;
                const unsigned long long pol = i < t ? S.l2_keep : 0ULL;
                for (int row = tid * 2; row < ld; row += NT * 2) cp_async16(dst + row, col + row, pol);
            }
            cp_async_commit();
        }
    };
    if constexpr (RING_D > 0) {
#pragma unroll
        for (int i = 0; i < RING_D; ++i) ring_issue(i);
    }
    double s2 = 0.0, sab = 0.0;
    for (int row = tid; row < ld; row += NT) {
        const double e = (double)aj[row];
        S.v[row] = e; s2 = fma(e, e, s2); sab = fma(e, b_at(row), sab);
    }
    block_sum2<NT>(s2, sab, S.red);                                // ||a||^2 and <a, b> in one reduction
    const double anorm2 = s2, ab = sab;
    double before2 = anorm2, rho2 = anorm2;
    // Fast path (the common case: the atom keeps at least half of its squared norm): rho^2 = ||a||^2 - ||Q'a||^2 by
    // Pythagoras and z_t = (<a,b> - <Q'a, Q'b>) / rho, so NO reduction of length M stands between the triangular
    // mat-vecs and the residual update, and v = a - A_S y is formed in the same row sweep that down-dates r.  Below the
    // threshold (where the subtraction would cancel) the explicit path with DGKS re-orthogonalisation takes over.
    bool fast = t == 0;
    for (int sweep = 0; sweep < 2 && t > 0; ++sweep) {
        if (sweep == 0 && gcol) {
            for (int i = tid; i < t; i += NT) S.g[i] = gcol[S.ssel[i] - idx_offset];   // g = (A'A)[S, j]
        } else {
            for (int i = warp; i < t; i += NT / 32) {              // g = A_S' v (four independent FMA chains per lane)
                const T* ai = S.colp[i];
                double s0 = 0.0, s1 = 0.0, s2b = 0.0, s3 = 0.0;
                int row = lane * W;
                for (; row + 32 * W < ld; row += 64 * W) {
                    double a0[W], a1[W];
                    RowVec<T>::load(ai + row, a0);
                    RowVec<T>::load(ai + row + 32 * W, a1);
#pragma unroll
                    for (int e = 0; e < W; e += 2) {
                        s0 = fma(a0[e], S.v[row + e], s0); s1 = fma(a0[e + 1], S.v[row + e + 1], s1);
                        s2b = fma(a1[e], S.v[row + 32 * W + e], s2b); s3 = fma(a1[e + 1], S.v[row + 32 * W + e + 1], s3);
                    }
                }
                if (row < ld) {
                    double a0[W];
                    RowVec<T>::load(ai + row, a0);
#pragma unroll
                    for (int e = 0; e < W; e += 2) { s0 = fma(a0[e], S.v[row + e], s0); s1 = fma(a0[e + 1], S.v[row + e + 1], s1); }
                }
                double s = warp_sum((s0 + s1) + (s2b + s3));
                if (lane == 0) S.g[i] = s;
            }
        }
        __syncthreads();
        // hh = R^{-T} g = Q'v and y = R^{-1} hh as two triangular mat-vecs with the stored inverse:
        // no substitution chain, every output element is an independent dot product.
        for (int i = tid; i < t; i += NT) {
            double h0 = 0.0, h1 = 0.0, h2 = 0.0, h3 = 0.0;
            const double* col = S.Tm + (size_t)i * S.ldT;
            int l = 0;
            for (; l + 3 <= i; l += 4) {
                h0 = fma(col[l], S.g[l], h0); h1 = fma(col[l + 1], S.g[l + 1], h1);
                h2 = fma(col[l + 2], S.g[l + 2], h2); h3 = fma(col[l + 3], S.g[l + 3], h3);
            }
            for (; l <= i; ++l) h0 = fma(col[l], S.g[l], h0);
            S.hh[i] = (h0 + h1) + (h2 + h3);
        }
        __syncthreads();
        for (int i = tid; i < t; i += NT) {
            double h0 = 0.0, h1 = 0.0, h2 = 0.0, h3 = 0.0;
            const double* rowp = S.Tm + i;
            int l = i;
            for (; l + 3 < t; l += 4) {
                h0 = fma(rowp[(size_t)l * S.ldT], S.hh[l], h0); h1 = fma(rowp[(size_t)(l + 1) * S.ldT], S.hh[l + 1], h1);
                h2 = fma(rowp[(size_t)(l + 2) * S.ldT], S.hh[l + 2], h2); h3 = fma(rowp[(size_t)(l + 3) * S.ldT], S.hh[l + 3], h3);
            }
            for (; l < t; ++l) h0 = fma(rowp[(size_t)l * S.ldT], S.hh[l], h0);
            const double acc = (h0 + h1) + (h2 + h3);
            S.y[i] = acc;
            S.ys[i] = sweep ? S.ys[i] + acc : acc;
        }
        if (sweep == 0 && warp == NT / 32 - 1) {                   // ||Q'a||^2 and <Q'a, Q'b> by the last warp, same phase
            double p = 0.0, q = 0.0;
            for (int i = lane; i < t; i += 32) { const double h = S.hh[i]; p = fma(h, h, p); q = fma(h, S.zs[i], q); }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                p += __shfl_xor_sync(0xffffffffu, p, off);
                q += __shfl_xor_sync(0xffffffffu, q, off);
            }
            if (lane == 0) { S.red[2 * (NT / 32)] = p; S.red[2 * (NT / 32) + 1] = q; }
        }
        __syncthreads();
        if (sweep == 0) {
            const double hn2 = S.red[2 * (NT / 32)];
            if (anorm2 - hn2 >= 0.5 * anorm2) { rho2 = anorm2 - hn2; fast = true; break; }
        }
        s2 = 0.0;
        for (int row = tid * W; row < ld; row += NT * W) {         // v -= A_S y, W rows per thread per 16 B load
            double acc[W];
#pragma unroll
            for (int e = 0; e < W; ++e) acc[e] = S.v[row + e];
            double acc1[W];
#pragma unroll
            for (int e = 0; e < W; ++e) acc1[e] = 0.0;
            int i = 0;
#pragma unroll 2
            for (; i + 1 < t; i += 2) {                             // two independent chains per row
                double a0[W], a1[W];
                RowVec<T>::load(S.colp[i] + row, a0);
                RowVec<T>::load(S.colp[i + 1] + row, a1);
                const double y0 = S.y[i], y1 = S.y[i + 1];
#pragma unroll
                for (int e = 0; e < W; ++e) { acc[e] = fma(-a0[e], y0, acc[e]); acc1[e] = fma(-a1[e], y1, acc1[e]); }
            }
            if (i < t) {
                double a0[W];
                RowVec<T>::load(S.colp[i] + row, a0);
                const double y0 = S.y[i];
#pragma unroll
                for (int e = 0; e < W; ++e) acc[e] = fma(-a0[e], y0, acc[e]);
            }
#pragma unroll
            for (int e = 0; e < W; ++e) { acc[e] += acc1[e]; S.v[row + e] = acc[e]; s2 = fma(acc[e], acc[e], s2); }
        }
        rho2 = block_sum<NT>(s2, S.red);
        if (rho2 >= 0.5 * before2) break;                          // DGKS: one sweep was enough
        before2 = rho2;
    }
    if (!(rho2 > 1e-26 * anorm2)) {                                // numerically dependent atom: not appended
        if constexpr (RING_D > 0) cp_async_wait<0>();
        return 1;
    }
    if (rho2 < ILLCOND_RATIO * anorm2) S.illcond = 1;
    const double rho = sqrt(rho2);
    double zt, s2r = 0.0;
    if (fast) {
        // z_t = q_t'b = (<a,b> - <Q'a, Q'b>) / rho;  r <- r - (z_t / rho) (a - A_S y), v formed on the fly
        zt = (t > 0 ? ab - S.red[2 * (NT / 32) + 1] : ab) / rho;
        const double gam = zt / rho;
        if (S.def_y) {                                             // the sweep runs later, row slice by row slice over all signals
            for (int i = tid; i < t; i += NT) S.def_y[i] = S.y[i];
            if (tid == 0) { *S.def_gam = gam; *S.def_t = t; }
            S.deferred = true;
        } else if constexpr (RING_D > 0) {
            double acc[RING_MAX_SLOTS][2], acc1[RING_MAX_SLOTS][2];
#pragma unroll
            for (int k = 0; k < RING_MAX_SLOTS; ++k) {
                const int row = tid * 2 + k * NT * 2;
                acc1[k][0] = acc1[k][1] = 0.0;
                if (row < ld) { acc[k][0] = S.v[row]; acc[k][1] = S.v[row + 1]; }
                else { acc[k][0] = acc[k][1] = 0.0; }
            }
            auto step = [&](int i, double (&ac)[RING_MAX_SLOTS][2]) {
                cp_async_wait<RING_D - 1>();                       // column i has landed (RING_D - 1 younger groups may be open)
                const double* src = S.ring + (size_t)(i % RING_D) * ld + tid * 2;
                const double yi = S.y[i];
                double2 c[RING_MAX_SLOTS];
#pragma unroll
                for (int k = 0; k < RING_MAX_SLOTS; ++k)
                    if (tid * 2 + k * NT * 2 < ld) c[k] = *reinterpret_cast<const double2*>(src + k * NT * 2);
#pragma unroll
                for (int k = 0; k < RING_MAX_SLOTS; ++k)
                    if (tid * 2 + k * NT * 2 < ld) { ac[k][0] = fma(-c[k].x, yi, ac[k][0]); ac[k][1] = fma(-c[k].y, yi, ac[k][1]); }
                ring_issue(i + RING_D);                            // the slot just read is free again
            };
            int i = 0;
            for (; i + 1 < t; i += 2) { step(i, acc); step(i + 1, acc1); }
            if (i < t) step(i, acc);
            cp_async_wait<0>();
            const double* rs = S.ring + (size_t)(t % RING_D) * ld;  // the residual came in behind the last column
#pragma unroll
            for (int k = 0; k < RING_MAX_SLOTS; ++k) {
                const int row = tid * 2 + k * NT * 2;
                if (row < ld) {
                    const double2 rv = *reinterpret_cast<const double2*>(rs + row);
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const double vq = acc[k][e] + acc1[k][e];
                        S.v[row + e] = vq;
                        const T rr = (T)((e ? rv.y : rv.x) - gam * vq);
                        r_set(row + e, rr);
                        s2r = fma((double)rr, (double)rr, s2r);
                    }
                }
            }
        } else
        for (int row = tid * W; row < ld; row += NT * W) {
            double acc[W], acc1[W];
#pragma unroll
            for (int e = 0; e < W; ++e) { acc[e] = S.v[row + e]; acc1[e] = 0.0; }
            int i = 0;
#pragma unroll 2
            for (; i + 1 < t; i += 2) {
                double a0[W], a1[W];
                RowVec<T>::load(S.colp[i] + row, a0);
                RowVec<T>::load(S.colp[i + 1] + row, a1);
                const double y0 = S.y[i], y1 = S.y[i + 1];
#pragma unroll
                for (int e = 0; e < W; ++e) { acc[e] = fma(-a0[e], y0, acc[e]); acc1[e] = fma(-a1[e], y1, acc1[e]); }
            }
            if (i < t) {
                double a0[W];
                RowVec<T>::load(S.colp[i] + row, a0);
                const double y0 = S.y[i];
#pragma unroll
                for (int e = 0; e < W; ++e) acc[e] = fma(-a0[e], y0, acc[e]);
            }
#pragma unroll
            for (int e = 0; e < W; ++e) {
                const double vq = acc[e] + acc1[e];
                S.v[row + e] = vq;                                 // q_t = v / rho is read by forward regression
                const T rr = (T)(r_at(row + e) - gam * vq);
                r_set(row + e, rr);
                s2r = fma((double)rr, (double)rr, s2r);
            }
        }
    } else {
        if constexpr (RING_D > 0) cp_async_wait<0>();              // the prefetched columns are not used on this path
        double sb = 0.0;
        for (int row = tid; row < ld; row += NT) sb += S.v[row] * b_at(row);
        zt = block_sum<NT>(sb, S.red) / rho;                       // z_t = q_t' b
        // residual: r = b - Q Q'b gains one term, r <- r - q_t z_t.  Identical to the reference's
        // from-scratch b - A_S x_S (x_S = R^{-1} Q'b) up to rounding, at one pass less over A_S.
        const double gam = zt / rho;
        for (int row = tid; row < ld; row += NT) {
            const T rr = (T)(r_at(row) - gam * S.v[row]);
            r_set(row, rr);
            s2r += (double)rr * (double)rr;
        }
    }
    if (!S.deferred) nr2 = block_sum<NT>(s2r, S.red);
    // append the column [h; rho] to R  <=>  append [-R^{-1}h / rho; 1/rho] to R^{-1}
    const double irho = 1.0 / rho;
    for (int i = tid; i < t; i += NT) {
        const double e = -S.ys[i] * irho;
        if (S.Tg) S.Tg[i + (size_t)t * S.kcap] = e;
        if (S.Tsm) S.Tsm[i + t * S.ldT] = e;
    }
    if (tid == 0) {
        if (S.Tg) S.Tg[t + (size_t)t * S.kcap] = irho;
        if (S.Tsm) S.Tsm[t + t * S.ldT] = irho;
        S.zs[t] = zt; S.ssel[t] = j; S.colp[t] = aj;
    }
    ++t;
    __syncthreads();
    return 0;
}

// ------------------------------------------------------------------------------------------------------
// Block append for GOMP (`addindex!(x, AiQR, A, indices)`, src/util.jl:129-134: l atoms per update!).
// Appending the m new atoms one at a time gathers the t active atoms 2 m times; here they are gathered
// twice per UPDATE: one sweep forms every inner product the m appends will need,
//     Gm = [A_S  a_J]' a_J                     ((t + m) x m),
// the m sequential factor updates then run on small matrices only -- for atom c (factor size tc = t + c):
//     hh = T' Gm[0:tc, c],  y = T hh,  rho^2 = ||a_c||^2 - ||hh||^2,  T gains [-y / rho; 1 / rho]
// -- and a second sweep forms all m orthogonalised directions at once,
//     v_c = a_c - A_S y_c[0:t] - sum_{c' < c} a_{c'} y_c[t + c'].
// rho^2 comes from Pythagoras instead of the explicit vector, which is only safe while the atom is not nearly
// dependent: the block stops at the first atom with rho^2 < ||a||^2 / 2 (the DGKS threshold) and the caller
// appends that atom and the rest one by one with append_atom (explicit norm, re-orthogonalisation).
// Vb: [BM][ld] shared block (the new atoms on entry, their orthogonalised directions on exit);
// Gm, Ym: [(kcap + BM)][BM] shared scratch.  Returns the number of atoms appended (<= m).
constexpr int BLOCK_MAX = 8;

// BMR: size of the per-thread register arrays (>= m; 4 or 8): with the arrays sized for 8 atoms the 16-byte loads of the
// sweeps do not fit the 128 registers two CTAs per SM allow.  Shared-memory strides stay BLOCK_MAX.
template <typename T, int NT, int BMR, typename BAt, typename RAt, typename RSet>
__device__ __forceinline__ int append_block(PursuitSmem<T>& S, int& t, int m, const int* __restrict__ J,
                                            const T* const* __restrict__ Jcol, int ld, double* __restrict__ Vb,
                                            double* __restrict__ Gm, double* __restrict__ Ym, double* __restrict__ sc,
                                            BAt b_at, RAt r_at, RSet r_set, double& nr2,
                                            const double* __restrict__ gram = nullptr, int gramN = 0, int idx_offset = 0) {
    constexpr int W = RowVec<T>::W;
    constexpr int WP = 2;                                      // rows per thread and step in the copy and in sweep 2
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int t0 = t;
    // sc: [0, BM) ||a_c||^2, [BM, 2BM) rho_c, [2BM, 3BM) v_c'b, [3BM, 4BM) ||v_c||^2
    double part[BMR];
#pragma unroll
    for (int c = 0; c < BMR; ++c) part[c] = 0.0;
    for (int row = tid * WP; row < ld; row += NT * WP) {           // the new atoms: two rows per load, all m in flight
        double e[BMR][WP];
#pragma unroll
        for (int c = 0; c < BMR; ++c)
            if (c < m) RowPair<T>::load(Jcol[c] + row, e[c]);
#pragma unroll
        for (int c = 0; c < BMR; ++c)
            if (c < m) {
#pragma unroll
                for (int w = 0; w < WP; ++w) { Vb[c * ld + row + w] = e[c][w]; part[c] = fma(e[c][w], e[c][w], part[c]); }
            }
    }
#pragma unroll
    for (int c = 0; c < BMR; ++c)
        if (c < m) { const double s = block_sum<NT>(part[c], S.red); if (tid == 0) sc[c] = s; }
    __syncthreads();
    // sweep 1: Gm[i][c] = <column i, a_c>, columns i < t0 from the dictionary, i >= t0 from the block itself --
    // or, when the dictionary's Gram matrix A'A is cached, (t0 + m) m scattered 8-byte loads and no gather at all
    if (gram) {
        for (int e = tid; e < (t0 + m) * m; e += NT) {
            const int i = e / m, c = e - i * m;
            const int col = (i < t0 ? S.ssel[i] : J[i - t0]) - idx_offset;
            Gm[i * BLOCK_MAX + c] = gram[(size_t)(J[c] - idx_offset) * gramN + col];
        }
    } else
    for (int i = warp; i < t0 + m; i += NT / 32) {
        double acc[BMR];
#pragma unroll
        for (int c = 0; c < BMR; ++c) acc[c] = 0.0;
        if (i < t0) {
            const T* ai = S.colp[i];
#pragma unroll 2
            for (int row = lane * W; row < ld; row += 32 * W) {
                double a[W];
                RowVec<T>::load(ai + row, a);
#pragma unroll
                for (int c = 0; c < BMR; ++c)
                    if (c < m) {
#pragma unroll
                        for (int e = 0; e < W; ++e) acc[c] = fma(a[e], Vb[c * ld + row + e], acc[c]);
                    }
            }
        } else {
            const double* ai = Vb + (size_t)(i - t0) * ld;
            for (int row = lane; row < ld; row += 32) {
                const double a = ai[row];
#pragma unroll
                for (int c = 0; c < BMR; ++c) if (c < m) acc[c] = fma(a, Vb[c * ld + row], acc[c]);
            }
        }
#pragma unroll
        for (int c = 0; c < BMR; ++c)
            if (c < m) { const double s = warp_sum(acc[c]); if (lane == 0) Gm[i * BLOCK_MAX + c] = s; }
    }
    __syncthreads();
    // the m factor updates, on small matrices only
    int done = 0;
    for (int c = 0; c < m; ++c) {
        const int tc = t0 + c;
        for (int i = tid; i < tc; i += NT) {
            double acc = 0.0;
            for (int l = 0; l <= i; ++l) acc = fma(S.Tm[l + i * S.ldT], Gm[l * BLOCK_MAX + c], acc);
            S.hh[i] = acc;
        }
        __syncthreads();
        double hn = 0.0;
        for (int i = tid; i < tc; i += NT) {
            double acc = 0.0;
            for (int l = i; l < tc; ++l) acc = fma(S.Tm[i + l * S.ldT], S.hh[l], acc);
            S.y[i] = acc;
            hn += S.hh[i] * S.hh[i];
        }
        const double hn2 = block_sum<NT>(hn, S.red);
        const double an2 = sc[c];
        const double rho2 = an2 - hn2;
        if (!(rho2 >= 0.5 * an2)) break;                       // nearly dependent: leave it to append_atom
        const double rho = sqrt(rho2), irho = 1.0 / rho;
        for (int i = tid; i < tc; i += NT) {
            const double e = -S.y[i] * irho;
            if (S.Tg) S.Tg[i + (size_t)tc * S.kcap] = e;
            if (S.Tsm) S.Tsm[i + tc * S.ldT] = e;
            Ym[i * BLOCK_MAX + c] = S.y[i];
        }
        if (tid == 0) {
            if (S.Tg) S.Tg[tc + (size_t)tc * S.kcap] = irho;
            if (S.Tsm) S.Tsm[tc + tc * S.ldT] = irho;
            sc[BLOCK_MAX + c] = rho;
            S.ssel[tc] = J[c]; S.colp[tc] = Jcol[c];
        }
        ++done;
        __syncthreads();
    }
    if (done == 0) return 0;
    // sweep 2: all `done` directions at once, two rows per thread and step (16-byte gathers of FP64 atoms).  In
    // place: a thread owns its rows, and the originals a_c are read from Vb before the row's results overwrite them.
    double pb[BMR], pn[BMR];
#pragma unroll
    for (int c = 0; c < BMR; ++c) { pb[c] = 0.0; pn[c] = 0.0; }
    for (int row = tid * WP; row < ld; row += NT * WP) {
        double acc[BMR][WP];
#pragma unroll
        for (int c = 0; c < BMR; ++c)
#pragma unroll
            for (int w = 0; w < WP; ++w) acc[c][w] = c < done ? Vb[c * ld + row + w] : 0.0;
#pragma unroll 8
        for (int i = 0; i < t0; ++i) {
            double a[WP];
            RowPair<T>::load(S.colp[i] + row, a);
#pragma unroll
            for (int c = 0; c < BMR; ++c)
                if (c < done) {
                    const double yc = Ym[i * BLOCK_MAX + c];
#pragma unroll
                    for (int w = 0; w < WP; ++w) acc[c][w] = fma(-a[w], yc, acc[c][w]);
                }
        }
#pragma unroll
        for (int c = 1; c < BMR; ++c)
#pragma unroll
            for (int cp = 0; cp < c; ++cp)
                if (c < done) {
                    const double yc = Ym[(t0 + cp) * BLOCK_MAX + c];
#pragma unroll
                    for (int w = 0; w < WP; ++w) acc[c][w] = fma(-Vb[cp * ld + row + w], yc, acc[c][w]);
                }
#pragma unroll
        for (int w = 0; w < WP; ++w) {
            const double bb = b_at(row + w);
#pragma unroll
            for (int c = 0; c < BMR; ++c)
                if (c < done) { pb[c] = fma(acc[c][w], bb, pb[c]); pn[c] = fma(acc[c][w], acc[c][w], pn[c]); }
        }
#pragma unroll
        for (int c = 0; c < BMR; ++c)
            if (c < done) {
#pragma unroll
                for (int w = 0; w < WP; ++w) Vb[c * ld + row + w] = acc[c][w];
            }
    }
#pragma unroll
    for (int c = 0; c < BMR; ++c)
        if (c < done) {
            const double sb = block_sum<NT>(pb[c], S.red);
            const double sn = block_sum<NT>(pn[c], S.red);
            if (tid == 0) { sc[2 * BLOCK_MAX + c] = sb; sc[3 * BLOCK_MAX + c] = sn; S.zs[t0 + c] = sb / sc[BLOCK_MAX + c]; }
        }
    __syncthreads();
    // r <- r - sum_c q_c (q_c'b), with the explicit norms ||v_c||^2
    double s2r = 0.0;
    for (int row = tid; row < ld; row += NT) {
        double acc = r_at(row);
#pragma unroll
        for (int c = 0; c < BMR; ++c)
            if (c < done) acc = fma(-sc[2 * BLOCK_MAX + c] / sc[3 * BLOCK_MAX + c], Vb[c * ld + row], acc);
        const T rr = (T)acc;
        r_set(row, rr);
        s2r += (double)rr * (double)rr;
    }
    nr2 = block_sum<NT>(s2r, S.red);
    t = t0 + done;
    __syncthreads();
    return done;
}

}  // namespace csb
