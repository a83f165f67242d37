// Whole-solve kernel for FEW signals on dictionaries up to about the L2 size: every `update!` of an omp / mp call
// for 1..8 signals in ONE cooperative launch (one CTA per SM, resident for the whole solve).
//
// Why.  With one signal the per-`update!` pipeline of the large paths -- GEMV kernel, then an 8-CTA cluster update
// kernel -- spends as long on kernel boundaries and on the update's chain of cluster barriers as on the correlation
// itself: 21 us per `update!` on the 64 MiB config-2 dictionary (0.49 of the HBM-bandwidth roofline of the
// dictionary bytes, profiles/c2s_r01.md), and with 2..23 signals every signal re-read the dictionary.  Here:
//   * CTAs [0, ns) are UPDATERS, one per signal: they keep the signal's whole pursuit state (b, r, support, R^{-1},
//     Q'b, and the columns of the active atoms) in shared memory from the first `update!` to the last and run the
//     reference's loop body (/root/reference/src/matchingpursuit.jl:26-31, 62-70, 73-82): final arg-max over the
//     workers' candidates, `i in x.nzind` check, append_atom (update_common.cuh), eps test.
//   * the other CTAs are WORKERS: worker w owns the contiguous atom range [N w / W, N (w + 1) / W), keeps as many of
//     its columns as fit in shared memory for the WHOLE solve (they are read from L2/HBM once per solve instead of once
//     per `update!`), streams the rest, and per `update!` computes c = A'r for all ns residuals in one pass over its
//     columns (register-blocked: CG columns x NS signals per warp step), fused with the |c| arg-max.
//   * the two roles hand over through global memory (L2) with self-validating 8-byte words (data + sequence number in
//     one atomic 64-bit access): workers publish (|c|, atom) per signal, the updater publishes the new residual.  No host
//     round trip, no kernel boundary, no grid-wide barrier, no memory fence in the loop.
// Every c_j is reduced inside one warp in the order corr_gemv.cu uses (lane i takes the 16-byte vectors i, i + 32, ..;
// xor-shuffle tree), so selections agree bit for bit with the multi-launch path; the update arithmetic is
// append_atom's (same as the CTA update kernel, with this kernel's block size).
// All spin loops are bounded (SPIN_LIMIT): a lost CTA turns into an error status, never a hung GPU.
#include "common.cuh"
#include "gemv_loads.cuh"
#include "update_common.cuh"

#include <cstdlib>
#include <cstring>

namespace csb {
namespace {

constexpr int PT = 512;                    // threads per CTA (128 registers per thread: CG x NS accumulators fit)
constexpr int PW = PT / 32;
constexpr unsigned SPIN_LIMIT = 1u << 24;  // polls before a wait gives up (a few seconds): an error status, never a hung GPU
constexpr int DBG_PHASES = 16;
constexpr unsigned BELL_STOP = 0xfffffffeu, BELL_ABORT = 0xffffffffu;   // doorbell values above every version number              // clock64() stamps per iteration and role (CSB200_PERSIST_DEBUG)

// ---- hand-over without fences: every 8-byte word carries its own sequence number ("LL" words) ----------------------
// word = (seq << 32) | 32 data bits, written and read as ONE 64-bit scalar access (single-copy atomic in the PTX
// memory model), so a reader that sees the expected seq sees the data that was stored with it -- no release/acquire
// fence (MEMBAR.SC/ALL.GPU + CCTL.IVALL on this part, the dominant cost of the first version of this kernel), no
// counter, no flag.  seq = (launch epoch << 16) | version: buffers are never cleared between launches.
__device__ __forceinline__ unsigned long long ll_word(unsigned data, unsigned seq) {
    return ((unsigned long long)seq << 32) | data;
}
__device__ __forceinline__ void ll_store(unsigned long long* p, unsigned long long w) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ void ll_store2(unsigned long long* p, unsigned long long w0, unsigned long long w1) {
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(w0), "l"(w1) : "memory");
}
__device__ __forceinline__ unsigned long long ll_load(const unsigned long long* p) {
    unsigned long long w;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
    return w;
}
__device__ __forceinline__ void ll_load2(const unsigned long long* p, unsigned long long& w0, unsigned long long& w1) {
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(p) : "memory");
}
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u32(unsigned* p, unsigned v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// One element of a residual as LL words: FP64 = two words (one 16-byte vector access of two atomic halves), FP32 = one.
template <typename T> struct ResLL;
template <> struct ResLL<double> {
    static constexpr int WORDS = 2;
    static __device__ __forceinline__ void store(unsigned long long* p, double v, unsigned seq) {
        const unsigned long long b = (unsigned long long)__double_as_longlong(v);
        ll_store2(p, ll_word((unsigned)b, seq), ll_word((unsigned)(b >> 32), seq));
    }
    static __device__ __forceinline__ bool load(const unsigned long long* p, unsigned seq, double& out) {
        unsigned long long w0, w1;
        ll_load2(p, w0, w1);
        if ((unsigned)(w0 >> 32) != seq || (unsigned)(w1 >> 32) != seq) return false;
        out = __longlong_as_double((long long)((w1 << 32) | (w0 & 0xffffffffull)));
        return true;
    }
};
template <> struct ResLL<float> {
    static constexpr int WORDS = 1;
    static __device__ __forceinline__ void store(unsigned long long* p, float v, unsigned seq) {
        ll_store(p, ll_word(__float_as_uint(v), seq));
    }
    static __device__ __forceinline__ bool load(const unsigned long long* p, unsigned seq, float& out) {
        const unsigned long long w = ll_load(p);
        if ((unsigned)(w >> 32) != seq) return false;
        out = __uint_as_float((unsigned)w);
        return true;
    }
};

// ---- dictionary columns kept in REGISTERS -------------------------------------------------------------------------
// A worker's 512 threads own 64 K registers (256 KiB) -- more than its shared memory.  The streaming loop needs about
// half of them, so each warp keeps REG_VECS 16-byte vectors per lane (64 registers) of dictionary for the whole solve:
// REG_VECS / NVL whole columns, NVL = vectors per lane and column (lane i holds the vectors i, i + 32, .. of a column,
// the same layout and summation order as every other column).  At 1024 x 8192 FP64 that is one more 8 KiB column per
// warp: 27 (shared memory) + 16 (registers) of a worker's 55-56 columns never leave the SM, and the L2 traffic per
// `update!` drops from 33 to 15 MiB.
constexpr int REG_VECS = 16;
__host__ __device__ constexpr int reg_vectors_for(int NS, bool f32) { return (NS <= 2 && !f32) ? REG_VECS : 0; }   // FP32 variants spill with it

template <typename T, int NS, int RQ, int NVL>
__device__ __forceinline__ void reg_column_dots(const typename Vec<T>::type (&regc)[RQ > 0 ? RQ : 1], const T* __restrict__ rs, int ld,
                                                int nvec, int lane, int ncols, int atom0, double (&best_v)[NS], int (&best_i)[NS]) {
    constexpr int W = Vec<T>::W;
    constexpr int RC = RQ / NVL;
#pragma unroll
    for (int c = 0; c < RC; ++c) {
        if (c >= ncols) break;                                            // warp-uniform
        double acc[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) acc[s] = 0.0;
#pragma unroll
        for (int q = 0; q < NVL; ++q) {
            const int i = lane + 32 * q;
            if (i < nvec) {
#pragma unroll
                for (int s = 0; s < NS; ++s) {
                    double rr[W];
#pragma unroll
                    for (int e = 0; e < W; ++e) rr[e] = (double)rs[(size_t)s * ld + i * W + e];
                    fma_vec(acc[s], regc[c * NVL + q], rr);
                }
            }
        }
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            double sum = acc[s];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
            const double v = fabs(sum);
            if (v >= 0.0 && cand_better(v, atom0 + c, best_v[s], best_i[s])) { best_v[s] = v; best_i[s] = atom0 + c; }
        }
    }
}

// ---- worker: c = A'r for NS residuals over this CTA's atom range, fused |c| arg-max --------------------------------
template <typename T, int NS, int CG>
__device__ void persist_worker(const PersistArgs& a, unsigned char* smem, double (*red_v)[PW], int (*red_i)[PW],
                               int* s_state) {
    using V = typename Vec<T>::type;
    constexpr int W = Vec<T>::W;
    constexpr int RW = ResLL<T>::WORDS;
    constexpr int RQ = reg_vectors_for(NS, sizeof(T) == 4);
    constexpr int UNR0 = CG * NS >= 32 ? 1 : (CG * NS >= 16 ? 2 : (CG * NS >= 4 ? 4 : 8));  // loads in flight vs the 128-register budget
    constexpr int UNR = (RQ > 0 && UNR0 > 1) ? UNR0 / 2 : UNR0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ld = a.ld, ns = a.ns;
    const int w = (int)blockIdx.x - ns;
    const int lo = (int)((long long)a.N * w / a.workers), hi = (int)((long long)a.N * (w + 1) / a.workers);
    const int ncol = hi - lo;
    const int ncache = ncol < a.wcache ? ncol : a.wcache;
    // register columns follow the shared-memory ones: warp q keeps columns ncache + q * rcw + [0, rcw) of the range
    const int nvl = RQ > 0 ? a.nvl : 0;                                   // vectors per lane and column: 0 (off), 4, 8 or 16
    const int rcw = nvl ? RQ / nvl : 0;
    const int nregs = nvl ? (ncol - ncache < PW * rcw ? ncol - ncache : PW * rcw) : 0;
    const int rfirst = ncache + warp * rcw;
    const int nreg_w = rfirst >= ncache + nregs ? 0 : (ncache + nregs - rfirst < rcw ? ncache + nregs - rfirst : rcw);
    const int sfirst = ncache + nregs;                                    // streamed columns: [sfirst, ncol)
    const int nstream = ncol - sfirst;
    const T* A = static_cast<const T*>(a.A);
    T* rs = reinterpret_cast<T*>(smem);                                   // [NS][ld]
    T* cache = rs + (size_t)NS * ld;                                      // [ncache][ld]: columns lo .. lo + ncache
    const unsigned epoch1 = a.epoch + 1u;
    const unsigned seq0 = a.epoch << 16;
    const unsigned long long pol = l2_policy(L2POL_NORMAL);
    const int nvec = ld / W;                                              // ld is a multiple of 16 elements
    long long* dbg = (a.dbg && w == 0) ? a.dbg + (size_t)DBG_PHASES * a.k : nullptr;   // worker 0's stamps follow the updater's

    {   // columns kept on chip for the whole solve
        const V* src = reinterpret_cast<const V*>(A + (size_t)lo * ld);
        V* dst = reinterpret_cast<V*>(cache);
        const int total = ncache * nvec;
        for (int i = tid; i < total; i += PT) dst[i] = ldg_stream(src + i, pol);
    }
    V regc[RQ > 0 ? RQ : 1];
    if constexpr (RQ > 0) {
        const int lg = nvl == 16 ? 4 : nvl == 8 ? 3 : 2;
#pragma unroll
        for (int q = 0; q < RQ; ++q) {
            const int c = q >> lg, i = lane + 32 * (q & (nvl - 1));
            regc[q] = V{};
            if (nvl && c < nreg_w && i < nvec) regc[q] = ldg_stream(reinterpret_cast<const V*>(A + (size_t)(lo + rfirst + c) * ld) + i, pol);
        }
    }
    auto reg_pass = [&](double (&bv)[NS], int (&bi)[NS]) {
        if constexpr (RQ > 0) {
            if (nreg_w <= 0) return;
            if (nvl == 16) reg_column_dots<T, NS, RQ, 16>(regc, rs, ld, nvec, lane, nreg_w, lo + rfirst, bv, bi);
            else if (nvl == 8) reg_column_dots<T, NS, RQ, 8>(regc, rs, ld, nvec, lane, nreg_w, lo + rfirst, bv, bi);
            else reg_column_dots<T, NS, RQ, 4>(regc, rs, ld, nvec, lane, nreg_w, lo + rfirst, bv, bi);
        }
    };
    for (int i = tid; i < (NS - ns) * ld; i += PT) rs[(size_t)ns * ld + i] = (T)0;     // unused signal slots
    {   // residual version 0 is b itself
        const V* src = reinterpret_cast<const V*>(static_cast<const T*>(a.B));
        V* dst = reinterpret_cast<V*>(rs);
        for (int i = tid; i < ns * nvec; i += PT) dst[i] = __ldcg(src + i);
    }
    const int gs = (nstream + CG - 1) / CG, gc = (ncache + CG - 1) / CG;

    for (int it = 0; it < a.k; ++it) {
        if (dbg && tid == 0) dbg[it * DBG_PHASES + 0] = clock64();
        if (it > 0) {
            // residual version `it` of every signal: one thread per signal watches THIS worker's doorbell of that signal
            // (the updater rings one doorbell per worker, 256 bytes apart: 147 CTAs polling one address queue up at a
            // single L2 slice -- ~3.5k cycles per hand-over in the first version of this kernel)
            const unsigned seq = seq0 | (unsigned)it;
            if (tid < ns) {
                const unsigned long long* bell = a.bell + ((size_t)tid * a.workers + w) * BELL_STRIDE;
                int state = 2;
                for (unsigned spin = 0; spin < SPIN_LIMIT; ++spin) {
                    const unsigned long long v = ll_load(bell);
                    if ((unsigned)(v >> 32) == epoch1) {
                        const unsigned ver = (unsigned)v;
                        if (ver == BELL_ABORT) break;
                        if (ver == BELL_STOP) { state = 1; break; }
                        if (ver >= (unsigned)it) { state = 0; break; }
                    }
                    if ((spin & 1023u) == 1023u && ld_relaxed_u32(a.ctrl + PERSIST_MAX_SIGNALS) == epoch1) break;   // someone gave up
                }
                if (state == 2) st_relaxed_u32(a.ctrl + PERSIST_MAX_SIGNALS, epoch1);
                s_state[tid] = state;
            }
            __syncthreads();
            bool all_stop = true, error = false;
            for (int s = 0; s < ns; ++s) { all_stop = all_stop && s_state[s] == 1; error = error || s_state[s] == 2; }
            if (all_stop || error) break;
            if (dbg && tid == 0) dbg[it * DBG_PHASES + 1] = clock64();
            const int shift = (int)(((long long)w * 64) % ld);            // workers start at different rows: all 147 of
            // them read the same words.  Four words in flight per thread (two signals x two rows), then the stragglers
            for (int s = 0; s < ns; s += 2) {
                const bool h[2] = {s_state[s] == 0, s + 1 < ns && s_state[s + 1] == 0};   // a stopped signal keeps its last residual
                if (!h[0] && !h[1]) continue;
                for (int e0 = tid; e0 < ld; e0 += 2 * PT) {
                    const unsigned long long* p[4];
                    T* d[4];
                    bool need[4], ok[4];
                    T val[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int ss = s + (u >> 1), ee = e0 + (u & 1) * PT;
                        const int e = ee + shift < ld ? ee + shift : ee + shift - ld;
                        need[u] = h[u >> 1] && ee < ld;
                        p[u] = a.r_ll + ((size_t)ss * ld + e) * RW;
                        d[u] = rs + (size_t)ss * ld + e;
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) ok[u] = need[u] ? ResLL<T>::load(p[u], seq, val[u]) : true;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (!need[u]) continue;
                        unsigned spin = 0;
                        while (!ok[u] && ++spin < SPIN_LIMIT) ok[u] = ResLL<T>::load(p[u], seq, val[u]);
                        *d[u] = val[u];
                    }
                }
            }
        }
        __syncthreads();
        if (dbg && tid == 0) dbg[it * DBG_PHASES + 2] = clock64();

        double best_v[NS];
        int best_i[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) { best_v[s] = -1.0; best_i[s] = INT_MAX; }
        // even warps take their register columns first, odd warps last: the L2 pipe is busy from the start of the pass
        // to its end, and the register columns' FMA chains run under the other warps' loads
        if (!(warp & 1)) reg_pass(best_v, best_i);
        for (int g0 = 0, rnd = 0; g0 < gs + gc; g0 += PW, ++rnd) {
            const int g = g0 + ((rnd & 1) ? PW - 1 - warp : warp);          // dealt back and forth: a warp with a streamed group
            if (g >= gs + gc) continue;                                    // gets its second group last (and a cheap one)
            const bool streamed = g < gs;                                  // streamed groups first: their loads fly longest
            const int first = streamed ? sfirst + g * CG : (g - gs) * CG;  // column offset inside the range
            const int limit = streamed ? ncol : ncache;
            double acc[CG][NS];
            const V* col[CG];
#pragma unroll
            for (int c = 0; c < CG; ++c) {
                const int cc = first + c < limit ? first + c : limit - 1;   // clamp: stay in bounds, result discarded
                col[c] = streamed ? reinterpret_cast<const V*>(A + (size_t)(lo + cc) * ld)
                                  : reinterpret_cast<const V*>(cache + (size_t)cc * ld);
#pragma unroll
                for (int s = 0; s < NS; ++s) acc[c][s] = 0.0;
            }
            if (streamed) {
#pragma unroll UNR
                for (int i = lane; i < nvec; i += 32) {
                    V x[CG];
#pragma unroll
                    for (int c = 0; c < CG; ++c) x[c] = ldg_stream(col[c] + i, pol);
#pragma unroll
                    for (int s = 0; s < NS; ++s) {
                        double rr[W];
#pragma unroll
                        for (int e = 0; e < W; ++e) rr[e] = (double)rs[(size_t)s * ld + i * W + e];
#pragma unroll
                        for (int c = 0; c < CG; ++c) fma_vec(acc[c][s], x[c], rr);
                    }
                }
            } else {
#pragma unroll UNR
                for (int i = lane; i < nvec; i += 32) {
                    V x[CG];
#pragma unroll
                    for (int c = 0; c < CG; ++c) x[c] = col[c][i];
#pragma unroll
                    for (int s = 0; s < NS; ++s) {
                        double rr[W];
#pragma unroll
                        for (int e = 0; e < W; ++e) rr[e] = (double)rs[(size_t)s * ld + i * W + e];
#pragma unroll
                        for (int c = 0; c < CG; ++c) fma_vec(acc[c][s], x[c], rr);
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < CG; ++c) {
                const int atom = lo + first + c;
                const bool valid = first + c < limit;
#pragma unroll
                for (int s = 0; s < NS; ++s) {
                    double sum = acc[c][s];
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
                    const double v = fabs(sum);
                    if (valid && v >= 0.0 && cand_better(v, atom, best_v[s], best_i[s])) { best_v[s] = v; best_i[s] = atom; }   // NaN never wins
                }
            }
        }
        if (warp & 1) reg_pass(best_v, best_i);
        if (lane == 0) {
#pragma unroll
            for (int s = 0; s < NS; ++s) { red_v[s][warp] = best_v[s]; red_i[s][warp] = best_i[s]; }
        }
        __syncthreads();
        if (dbg && tid == 0) dbg[it * DBG_PHASES + 3] = clock64();
        if (warp < ns) {                                                   // warp s folds the PW per-warp results of signal s
            double bv = lane < PW ? red_v[warp][lane] : -1.0;
            int bi = lane < PW ? red_i[warp][lane] : INT_MAX;
#pragma unroll
            for (int off = PW / 2; off > 0; off >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, bv, off);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                if (cand_better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
            }
            if (lane == 0) {
                const unsigned seq = seq0 | (unsigned)(it + 1);
                const unsigned long long b = (unsigned long long)__double_as_longlong(bv);
                // (the record address is rebuilt here on purpose: hoisted out of the loop it was spilled to local memory)
                int opaque_zero;
                asm volatile("mov.u32 %0, 0;" : "=r"(opaque_zero));
                unsigned long long* rec = a.cand_ll + ((size_t)(warp + opaque_zero) * a.workers + w) * 4;
                ll_store2(rec, ll_word((unsigned)b, seq), ll_word((unsigned)(b >> 32), seq));
                ll_store(rec + 2, ll_word((unsigned)((bi == INT_MAX) ? -1 : bi + a.idx_offset), seq));
            }
        }
        if (dbg && tid == 0) dbg[it * DBG_PHASES + 4] = clock64();
    }
}

// ---- updater: the loop body of `update!` for one signal, state resident in shared memory ---------------------------
// FP64 add / fma results are available ~40 cycles after issue on this part, so every DEPENDENT chain in the updater is
// kept short: block reductions are a 5-level shuffle tree plus a 4-level register tree over the 16 warp partials (not a
// 16-long serial sum), dot products run on 4 independent accumulators, and the two triangular mat-vecs of an append
// are done by one warp without block barriers.
__device__ __forceinline__ double tree_sum16(const double* row) {
    double v[PW];
#pragma unroll
    for (int q = 0; q < PW; ++q) v[q] = row[q];
#pragma unroll
    for (int w = PW / 2; w > 0; w >>= 1)
#pragma unroll
        for (int q = 0; q < w; ++q) v[q] += v[q + w];
    return v[0];
}
// Block reductions of up to three values with ONE barrier each: three scratch rows used in rotation (a row is not
// rewritten before two further barriers have passed).
struct Red3 {
    double* buf;      // [3][3][PW]
    int phase;
    __device__ __forceinline__ void sum3(double& a, double& b, double& c) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, off);
            b += __shfl_xor_sync(0xffffffffu, b, off);
            c += __shfl_xor_sync(0xffffffffu, c, off);
        }
        double* row = buf + (size_t)phase * 3 * PW;
        if (lane == 0) { row[warp] = a; row[PW + warp] = b; row[2 * PW + warp] = c; }
        __syncthreads();
        a = tree_sum16(row); b = tree_sum16(row + PW); c = tree_sum16(row + 2 * PW);
        phase = phase == 2 ? 0 : phase + 1;
    }
    __device__ __forceinline__ void sum1(double& a) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
        double* row = buf + (size_t)phase * 3 * PW;
        if (lane == 0) row[warp] = a;
        __syncthreads();
        a = tree_sum16(row);
        phase = phase == 2 ? 0 : phase + 1;
    }
};

// `add_column!(AiQR, a, pos)` + the residual down-date for ONE atom, specialised for this kernel (reference:
// src/util.jl:118-126, src/matchingpursuit.jl:152-176; same mathematics as append_atom in update_common.cuh: implicit Q,
// CGS with DGKS re-orthogonalisation, R^{-1} stored).  ONE signal is on the critical path of 147 waiting SMs here, so:
// a thread owns fixed rows of v / b / r (no barrier between the row-wise phases); ||a||^2, ||v||^2 and <v, b> share one
// reduction; the triangular mat-vecs run in one warp (lane = output, 4 accumulators); and the new residual leaves for
// the workers (store_r / after_store) before its norm is reduced.
template <typename T, typename StoreR, typename AfterStore>
__device__ __forceinline__ int append_fast(PursuitSmem<T>& S, int& t, int j, const T* __restrict__ ajg, int ld,
                                           const double* __restrict__ bs, double* __restrict__ rs, Red3& red,
                                           StoreR store_r, AfterStore after_store, double& nr2, T* acache, int ucache,
                                           long long* stamp = nullptr) {
    constexpr int W = RowVec<T>::W;
    using V = typename Vec<T>::type;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    auto mark = [&](int slot) { if (stamp && tid == 0) stamp[slot] = clock64(); };
    // active atom i: slot i of this CTA's shared-memory cache (address computed, so the loads are LDS -- a pointer
    // fetched from the colp table makes them generic loads at ~3x the latency) or, beyond the cache, the dictionary
    const int tc = t < ucache ? t : ucache;
    const bool cached = t < ucache;
    T* slot = acache + (size_t)(cached ? t : 0) * ld;
    double* sc = red.buf + 9 * PW;                                     // [4] <a,a>, <a,b>, ||Q'a||^2, <Q'a, Q'b>
    // v = a_j straight from the dictionary (L2), and into cache slot t while it fits: one pass, one barrier
    for (int row = tid * W; row < ld; row += PT * W) {
        const V x = *reinterpret_cast<const V*>(ajg + row);
        if (cached) *reinterpret_cast<V*>(slot + row) = x;
        double e[W];
        RowVec<T>::load(reinterpret_cast<const T*>(&x), e);
#pragma unroll
        for (int q = 0; q < W; ++q) S.v[row + q] = e[q];
    }
    const T* aj = cached ? slot : ajg;
    __syncthreads();
    mark(5);
    // one warp per dot product of length M: g_i = <a_i, v> for the t active atoms, then <v, v> and <v, b>
    auto dot_with_v = [&](auto load2) {                                // load2(row, e[W]): W values of the other operand
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        int row = lane * W;
        for (; row + 32 * W < ld; row += 64 * W) {                     // two steps per trip, four independent chains
            double e0[W], e1[W];
            load2(row, e0);
            load2(row + 32 * W, e1);
#pragma unroll
            for (int q = 0; q < W; ++q) {
                acc[q & 1] = fma(e0[q], S.v[row + q], acc[q & 1]);
                acc[2 + (q & 1)] = fma(e1[q], S.v[row + 32 * W + q], acc[2 + (q & 1)]);
            }
        }
        if (row < ld) {
            double e0[W];
            load2(row, e0);
#pragma unroll
            for (int q = 0; q < W; ++q) acc[q & 1] = fma(e0[q], S.v[row + q], acc[q & 1]);
        }
        return warp_sum((acc[0] + acc[1]) + (acc[2] + acc[3]));
    };
    auto gather_g = [&](bool with_norms) {
        const int n = t + (with_norms ? 2 : 0);
        for (int i = warp; i < n; i += PW) {
            double g;
            if (i < tc) { const T* ai = acache + (size_t)i * ld; g = dot_with_v([&](int row, double (&e)[W]) { RowVec<T>::load(ai + row, e); }); }
            else if (i < t) { const T* ai = S.colp[i]; g = dot_with_v([&](int row, double (&e)[W]) { RowVec<T>::load(ai + row, e); }); }
            else if (i == t) g = dot_with_v([&](int row, double (&e)[W]) {
#pragma unroll
                for (int q = 0; q < W; ++q) e[q] = S.v[row + q]; });
            else g = dot_with_v([&](int row, double (&e)[W]) {
#pragma unroll
                for (int q = 0; q < W; ++q) e[q] = bs[row + q]; });
            if (lane == 0) { if (i < t) S.g[i] = g; else sc[i - t] = g; }
        }
        __syncthreads();
    };
    // hh = R^{-T} g = Q'v (hh_i = sum_{l <= i} T[l, i] g_l), its squared norm and <hh, Q'b>, and y = R^{-1} hh
    // (y_i = sum_{l >= i} T[i, l] hh_l).  Half a warp per output: 16 lanes split the sum and fold it with four shuffles,
    // so each mat-vec is one short step for the whole block instead of a 32-long walk of a single warp (which was the
    // longest phase of the append: 3.3k of 15k cycles).
    auto mat_vecs = [&](int sweep) {
        const int half = lane >> 4, l16 = lane & 15;
        const int ldT = S.ldT;
        for (int i0 = 2 * warp; i0 < t; i0 += 2 * PW) {
            const int i = i0 + half;
            double acc = 0.0;
            if (i < t) {
                const double* col = S.Tm + (size_t)i * ldT;
                for (int l = l16; l <= i; l += 16) acc = fma(col[l], S.g[l], acc);
            }
#pragma unroll
            for (int off = 8; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
            if (l16 == 0 && i < t) S.hh[i] = acc;
        }
        __syncthreads();
        for (int i0 = 2 * warp; i0 < t; i0 += 2 * PW) {
            const int i = i0 + half;
            double acc = 0.0;
            if (i < t) {
                const double* rowp = S.Tm + i;
                for (int l = i + l16; l < t; l += 16) acc = fma(rowp[(size_t)l * ldT], S.hh[l], acc);
            }
#pragma unroll
            for (int off = 8; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
            if (l16 == 0 && i < t) { S.y[i] = acc; S.ys[i] = sweep ? S.ys[i] + acc : acc; }
        }
        if (sweep == 0 && warp == PW - 1) {                            // ||Q'a||^2 and <Q'a, Q'b> for the fast path
            double p = 0.0, q2 = 0.0;
            for (int i = lane; i < t; i += 32) { const double h = S.hh[i]; p = fma(h, h, p); q2 = fma(h, S.zs[i], q2); }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                p += __shfl_xor_sync(0xffffffffu, p, off);
                q2 += __shfl_xor_sync(0xffffffffu, q2, off);
            }
            if (lane == 0) { sc[2] = p; sc[3] = q2; }
        }
        __syncthreads();
    };
    // v_rows(row, out[W]): v - A_S y on W rows (two chains)
    auto v_rows = [&](int row, double (&out)[W]) {
        double acc[2][W];
#pragma unroll
        for (int q = 0; q < W; ++q) { acc[0][q] = S.v[row + q]; acc[1][q] = 0.0; }
        int i = 0;
#pragma unroll 2
        for (; i + 1 < tc; i += 2) {                                    // cached atoms: shared-memory loads, two chains
            double e0[W], e1[W];
            RowVec<T>::load(acache + (size_t)i * ld + row, e0);
            RowVec<T>::load(acache + (size_t)(i + 1) * ld + row, e1);
            const double y0 = S.y[i], y1 = S.y[i + 1];
#pragma unroll
            for (int q = 0; q < W; ++q) { acc[0][q] = fma(-e0[q], y0, acc[0][q]); acc[1][q] = fma(-e1[q], y1, acc[1][q]); }
        }
        if (i < tc) {
            double e0[W];
            RowVec<T>::load(acache + (size_t)i * ld + row, e0);
            const double y0 = S.y[i];
#pragma unroll
            for (int q = 0; q < W; ++q) acc[0][q] = fma(-e0[q], y0, acc[0][q]);
            ++i;
        }
        for (; i < t; ++i) {                                           // beyond the cache: from the dictionary (L2)
            double e0[W];
            RowVec<T>::load(S.colp[i] + row, e0);
            const double y0 = S.y[i];
#pragma unroll
            for (int q = 0; q < W; ++q) acc[0][q] = fma(-e0[q], y0, acc[0][q]);
        }
#pragma unroll
        for (int q = 0; q < W; ++q) out[q] = acc[0][q] + acc[1][q];
    };

    gather_g(true);                                                    // g, <a,a>, <a,b>
    mark(6);
    if (t > 0) mat_vecs(0);
    mark(7);
    const double anorm2 = sc[0], ab = sc[1];
    double rho2 = anorm2, vb = ab;
    // Fast path: the atom keeps at least half of its squared norm, so rho^2 = ||a||^2 - ||Q'a||^2 (Pythagoras) and
    // <v, b> = <a, b> - <Q'a, Q'b> are safe, and no reduction of length M stands between the mat-vecs and the doorbells
    bool fast = t == 0;
    if (t > 0 && anorm2 - sc[2] >= 0.5 * anorm2) { fast = true; rho2 = anorm2 - sc[2]; vb = ab - sc[3]; }
    if (!fast) {
        // explicit path: v -= A_S y with ||v||^2 and <v, b> reduced; a second sweep when ||v|| collapsed (DGKS)
        double before2 = anorm2;
        for (int sweep = 0; sweep < 2; ++sweep) {
            if (sweep == 1) { gather_g(false); mat_vecs(1); }
            double s2 = 0.0, sb = 0.0, unused = 0.0;
            for (int row = tid * W; row < ld; row += PT * W) {
                double vq[W];
                v_rows(row, vq);
#pragma unroll
                for (int q = 0; q < W; ++q) { S.v[row + q] = vq[q]; s2 = fma(vq[q], vq[q], s2); sb = fma(vq[q], bs[row + q], sb); }
            }
            red.sum3(s2, sb, unused);
            rho2 = s2; vb = sb;
            mark(sweep == 0 ? 8 : 11);
            if (rho2 >= 0.5 * before2) break;
            before2 = rho2;
        }
    }
    if (!(rho2 > 1e-26 * anorm2)) return 1;                            // numerically dependent atom: not appended
    if (rho2 < ILLCOND_RATIO * anorm2) S.illcond = 1;
    const double irho = rsqrt(rho2);                                   // 1 / rho (<= 1 ulp), no division on the critical path
    const double zt = vb * irho;                                       // z_t = q_t' b
    const double gam = zt * irho;
    double s2r = 0.0;
    for (int row = tid * W; row < ld; row += PT * W) {                 // r <- r - q_t z_t on this thread's rows
        double vq[W];
        if (fast) v_rows(row, vq);
        else {
#pragma unroll
            for (int q = 0; q < W; ++q) vq[q] = S.v[row + q];
        }
#pragma unroll
        for (int q = 0; q < W; ++q) {
            const T rr = (T)(rs[row + q] - gam * vq[q]);
            rs[row + q] = (double)rr;
            store_r(row + q, rr);
            s2r = fma((double)rr, (double)rr, s2r);
        }
    }
    after_store();                                                     // the workers can start while ||r|| is being reduced
    mark(9);
    for (int i = tid; i < t; i += PT) S.Tsm[i + t * S.ldT] = -S.ys[i] * irho;   // R^{-1} gains [-R^{-1}h / rho; 1 / rho]
    if (tid == 0) { S.Tsm[t + t * S.ldT] = irho; S.zs[t] = zt; S.ssel[t] = j; S.colp[t] = aj; }
    red.sum1(s2r);                                                     // barrier: the new column is visible
    nr2 = s2r;
    ++t;
    return 0;
}

template <typename T>
__device__ void persist_updater(const PersistArgs& a, unsigned char* smem, double* red_buf) {
    constexpr int RW = ResLL<T>::WORDS;
    constexpr int W = RowVec<T>::W;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int sig = blockIdx.x, ld = a.ld, kcap = a.kcap;
    const int ldT = kcap | 1;
    double* p = reinterpret_cast<double*>(smem);
    PursuitSmem<T> S;
    S.v = p; p += ld;
    double* bs = p; p += ld;                     // signal
    double* rs = p; p += ld;                     // residual (values are T-representable)
    S.g = p; p += kcap;
    S.hh = p; p += kcap;
    S.ys = p; p += kcap;
    S.y = p; p += kcap;
    S.zs = p; p += kcap;
    double* Tsm = p; p += (size_t)kcap * ldT;
    S.ssel = reinterpret_cast<int*>(p);
    S.colp = reinterpret_cast<const T**>(S.ssel + ((kcap + 1) & ~1));
    // [ucache][ld] after the tables, 16-byte aligned; the offset is computed arithmetically so that the compiler keeps
    // the shared address space (LDS) for these loads
    const size_t acache_off = (((size_t)(3 * ld + 5 * kcap + (size_t)kcap * ldT) * sizeof(double) +
                                (size_t)((kcap + 1) & ~1) * sizeof(int) + (size_t)kcap * sizeof(void*)) + 15) & ~(size_t)15;
    T* acache = reinterpret_cast<T*>(smem + acache_off);
    S.Tm = Tsm; S.Tsm = Tsm; S.ldT = ldT; S.Tg = nullptr; S.kcap = kcap; S.red = red_buf;
    Red3 red{red_buf, 0};
    __shared__ int s_j, s_fail;

    const T* A = static_cast<const T*>(a.A);
    const T* b = static_cast<const T*>(a.B) + (size_t)sig * ld;
    T* rg = static_cast<T*>(a.R) + (size_t)sig * ld;
    unsigned long long* rll = a.r_ll + (size_t)sig * ld * RW;
    unsigned long long* bells = a.bell + (size_t)sig * a.workers * BELL_STRIDE;
    const unsigned epoch1 = a.epoch + 1u;
    const unsigned seq0 = a.epoch << 16;
    long long* dbg = (a.dbg && sig == 0) ? a.dbg : nullptr;

    double s2 = 0.0;
    int bad = 0;
    for (int row = tid; row < ld; row += PT) {      // r = b: the state of a freshly constructed MP / OMP object
        const T e = b[row];
        bs[row] = (double)e; rs[row] = (double)e;
        s2 += (double)e * (double)e;
        bad |= !isfinite((double)e);
    }
    if (tid == 0) s_fail = 0;
    red.sum1(s2);
    double nr = sqrt(s2);
    int t = 0, flags = 0, iters = 0;
    bool done = false, failed = false;
    if (__syncthreads_or(bad)) { flags = 4; done = true; }

    // one doorbell per worker: "residual version `ver` is on its way" (the data words validate themselves)
    auto ring = [&](unsigned ver) {
        for (int c = tid; c < a.workers; c += PT) ll_store(bells + (size_t)c * BELL_STRIDE, ll_word(ver, epoch1));
    };

    for (int it = 0; it < a.k && !done; ++it) {
        if (a.mode == 0 && !(t < a.M)) {
            // `nnz(x) < size(P.A, 1) || return x` (:63): the support is full, every remaining update! is a no-op that
            // still tests eps (:79) -- the residual no longer changes, so either the first of them breaks or none does
            iters += (nr >= a.eps) ? (a.k - it) : 1;
            if (!(nr >= a.eps)) done = true;
            break;
        }
        if (dbg && tid == 0) dbg[it * DBG_PHASES + 0] = clock64();
        // candidates of this update!: ONE warp collects every worker's record (lane l waits for workers l, l + 32, ..),
        // folds them with five shuffles and hands the winner to the block through shared memory -- one barrier, no second
        // reduction stage (a 16-way pass over per-warp results by all 512 threads cost 3.6k cycles here)
        const unsigned cseq = seq0 | (unsigned)(it + 1);
        if (warp == 0) {
            double bv = -1.0;
            int bi = INT_MAX;
            bool lost = false;
            for (int c = lane; c < a.workers; c += 32) {
                const unsigned long long* rec = a.cand_ll + ((size_t)sig * a.workers + c) * 4;
                unsigned long long w0 = 0, w1 = 0, w2 = 0;
                bool ok = false;
                for (unsigned spin = 0; spin < SPIN_LIMIT && !ok; ++spin) {
                    ll_load2(rec, w0, w1);
                    w2 = ll_load(rec + 2);
                    ok = (unsigned)(w0 >> 32) == cseq && (unsigned)(w1 >> 32) == cseq && (unsigned)(w2 >> 32) == cseq;
                    if (!ok && (spin & 1023u) == 1023u && ld_relaxed_u32(a.ctrl + PERSIST_MAX_SIGNALS) == epoch1) break;
                }
                if (!ok) { lost = true; continue; }
                if (dbg) dbg[(size_t)2 * DBG_PHASES * a.k + (size_t)it * gridDim.x + c] = clock64();
                const double v = __longlong_as_double((long long)((w1 << 32) | (w0 & 0xffffffffull)));
                const int i = (int)(unsigned)w2;
                if (i >= 0 && cand_better(v, i, bv, bi)) { bv = v; bi = i; }
            }
            if (dbg && tid == 0) dbg[it * DBG_PHASES + 1] = clock64();
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, bv, off);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                if (cand_better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
            }
            lost = __any_sync(0xffffffffu, lost);
            if (lane == 0) { s_j = (bi == INT_MAX) ? -1 : bi; s_fail = lost ? 1 : 0; }
            if (dbg && tid == 0) dbg[it * DBG_PHASES + 12] = clock64();
        }
        __syncthreads();
        if (dbg && tid == 0) dbg[it * DBG_PHASES + 13] = clock64();
        if (s_fail) { failed = true; break; }
        const int j = s_j;                                               // global atom index or -1 (every thread has it)
        const unsigned rseq = seq0 | (unsigned)(it + 1);                 // the residual version this update! produces
        const bool more = it + 1 < a.k;                                  // somebody will read it
        if (dbg && tid == 0) dbg[it * DBG_PHASES + 2] = clock64();

        if (a.mode == 2) {
            // mp: x[i] += <a_i, r>;  r <- r - <a_i, r> a_i   (:26-31); one (atom, increment) record per iteration
            double c = 0.0;
            if (j >= 0) {
                const T* aj = A + (size_t)(j - a.idx_offset) * ld;
                double s = 0.0;
                for (int row = tid * W; row < ld; row += PT * W) {
                    double e[W];
                    RowVec<T>::load(aj + row, e);
#pragma unroll
                    for (int q = 0; q < W; ++q) { S.v[row + q] = e[q]; s = fma(e[q], rs[row + q], s); }
                }
                red.sum1(s);
                c = s;                                                   // dot(view(A,:,i), r)  (:29)
                s2 = 0.0;
                for (int row = tid * W; row < ld; row += PT * W) {
#pragma unroll
                    for (int q = 0; q < W; ++q) {
                        const T rr = (T)(rs[row + q] - c * S.v[row + q]);
                        rs[row + q] = (double)rr;
                        if (more) ResLL<T>::store(rll + (size_t)(row + q) * RW, rr, rseq);
                        s2 = fma((double)rr, (double)rr, s2);
                    }
                }
                if (more) ring((unsigned)(it + 1));
                red.sum1(s2);
                nr = sqrt(s2);
            } else {
                flags |= 2;
                if (more) {
                    for (int row = tid; row < ld; row += PT) ResLL<T>::store(rll + (size_t)row * RW, (T)rs[row], rseq);
                    ring((unsigned)(it + 1));
                }
            }
            if (tid == 0) {
                a.sel[(size_t)sig * a.stride + it] = j;
                a.x[(size_t)sig * a.stride + it] = c;
            }
            ++iters;
            t = iters;
            __syncthreads();
            if (dbg && tid == 0) dbg[it * DBG_PHASES + 4] = clock64();
            continue;
        }

        // omp: append the winner unless it is active already (:66) or the factorisation is full
        bool changed = false;
        if (j < 0) {
            flags |= 2;
        } else {
            int in = 0;
            for (int i = tid; i < t; i += PT) in |= (S.ssel[i] == j);
            if (!__syncthreads_or(in) && t < kcap) {
                // the new atom's column goes into this CTA's shared memory while it fits (inside append_fast, in the same
                // pass that loads v): the later sweeps over the active atoms then never leave the SM
                const T* aj = A + (size_t)(j - a.idx_offset) * ld;
                if (dbg && tid == 0) dbg[it * DBG_PHASES + 3] = clock64();
                double nr2 = 0.0;
                // the doorbells ring as soon as the new residual is on its way -- before its norm is known: if the eps test
                // then ends the solve, the workers' extra pass is discarded and they leave at the STOP ring
                const int dep = append_fast<T>(
                    S, t, j, aj, ld, bs, rs, red,
                    [&](int row, T val) { if (more) ResLL<T>::store(rll + (size_t)row * RW, val, rseq); },
                    [&]() { if (more) ring((unsigned)(it + 1)); }, nr2, acache, a.ucache, dbg ? dbg + it * DBG_PHASES : nullptr);
                if (dbg && tid == 0) dbg[it * DBG_PHASES + 10] = clock64();
                if (dep) flags |= 1; else { changed = true; nr = sqrt(nr2); }
            }
        }
        ++iters;
        if (!(nr >= a.eps)) done = true;                                 // `norm(residual!(P, x)) >= eps || break` (:79)
        if (!changed && !done && more) {                                 // r is unchanged: re-issue it under the new version
            for (int row = tid; row < ld; row += PT) ResLL<T>::store(rll + (size_t)row * RW, (T)rs[row], rseq);
            ring((unsigned)(it + 1));
        }
        __syncthreads();
        if (dbg && tid == 0) dbg[it * DBG_PHASES + 4] = clock64();
    }
    __syncthreads();
    ring(failed ? BELL_ABORT : BELL_STOP);                               // releases the workers on every exit path
    if (tid == 0 && failed) st_relaxed_u32(a.ctrl + PERSIST_MAX_SIGNALS, epoch1);

    // ---- results ----
    if (a.mode != 2) {
        for (int i = tid; i < t; i += PT) {                              // x_S = R^{-1} Q'b  (`ldiv!`, :175)
            double acc = 0.0;
            for (int l = i; l < t; ++l) acc = fma(Tsm[i + l * ldT], S.zs[l], acc);
            S.ys[i] = acc;
            a.sel[(size_t)sig * a.stride + i] = S.ssel[i];
        }
        if (S.illcond) { flags |= FLAG_ILLCOND; refine_coefficients<T, PT>(S, t, ld, [&](int row) { return bs[row]; }, S.ys); }
        else __syncthreads();
        for (int i = tid; i < t; i += PT) a.x[(size_t)sig * a.stride + i] = S.ys[i];
    }
    for (int row = tid; row < ld; row += PT) rg[row] = (T)rs[row];       // the residual where the batch keeps it
    if (tid == 0) {
        a.nnz[sig] = t;
        a.iters[sig] = iters;
        a.resnorm[sig] = nr;
        a.flags[sig] = flags | (failed ? 8 : 0);
        a.done[sig] = done ? 1 : 0;
    }
}

template <typename T, int NS, int CG>
__global__ void __launch_bounds__(PT, 1) persist_solve_kernel(PersistArgs a) {
    extern __shared__ __align__(16) unsigned char psm[];
    __shared__ double red_v[NS < 10 ? 10 : NS][PW];                      // workers: [NS][PW]; updater: Red3's [3][3][PW] + 4 scalars
    __shared__ int red_i[NS][PW];
    __shared__ int s_state[PERSIST_MAX_SIGNALS];
    if ((int)blockIdx.x < a.ns) persist_updater<T>(a, psm, &red_v[0][0]);
    else persist_worker<T, NS, CG>(a, psm, red_v, red_i, s_state);
}

size_t updater_fixed_bytes(int ld, int kcap) {
    size_t bytes = (size_t)(3 * ld + 5 * kcap + (size_t)kcap * (kcap | 1)) * sizeof(double) +
                   (size_t)((kcap + 1) & ~1) * sizeof(int) + (size_t)kcap * sizeof(void*);
    return (bytes + 15) / 16 * 16 + 16;
}

template <typename T, int NS, int CG>
cudaError_t persist_launch(const PersistArgs& a, size_t smem, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(persist_solve_kernel<T, NS, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    PersistArgs copy = a;
    void* params[] = {&copy};
    return cudaLaunchCooperativeKernel((const void*)persist_solve_kernel<T, NS, CG>, dim3(a.ns + a.workers), dim3(PT), params,
                                       smem, st);
}

template <typename T>
cudaError_t persist_launch_t(const PersistArgs& a, int NS, size_t smem, cudaStream_t st) {
    switch (NS) {
        case 1: return persist_launch<T, 1, 2>(a, smem, st);
        case 2: return persist_launch<T, 2, 2>(a, smem, st);
        case 4: return persist_launch<T, 4, 4>(a, smem, st);
        default: return persist_launch<T, 8, 4>(a, smem, st);
    }
}

int slots_for(int ns) { return ns <= 1 ? 1 : ns <= 2 ? 2 : ns <= 4 ? 4 : 8; }

}  // namespace

// Shared-memory plan of one launch; false when the shape does not fit (the caller takes the multi-launch path).
bool persist_plan(int ld, int N, int kcap, int ns, bool f32, int num_sms, PersistArgs* out, size_t* smem_out) {
    if (ns < 1 || ns > PERSIST_MAX_SIGNALS || kcap < 1 || num_sms < ns + 1) return false;
    const size_t es = f32 ? 4 : 8;
    const size_t budget = 222 * 1024;                         // of 227 KiB per CTA; the rest covers static shared memory
    const size_t upd = updater_fixed_bytes(ld, kcap);
    const size_t col = (size_t)ld * es;
    const int NS = slots_for(ns);
    const size_t wrk = (size_t)NS * col;
    if (upd > budget || wrk + col > budget) return false;
    int workers = num_sms - ns;
    if (workers > N) workers = N;
    const int per = (N + workers - 1) / workers;              // atoms of the largest worker range
    int wcache = (int)((budget - wrk) / col);
    if (wcache > per) wcache = per;
    // register columns (reg_column_dots): whole columns of 4, 8 or 16 vectors per lane
    int nvl = 0;
    {
        static const bool off = [] { const char* e = getenv("CSB200_PERSIST_REGCACHE"); return e && e[0] == '0'; }();
        const int nvec = ld / (f32 ? 4 : 2), per_lane = (nvec + 31) / 32;
        if (!off && reg_vectors_for(NS, f32) > 0 && per_lane <= REG_VECS && per > wcache) nvl = per_lane <= 4 ? 4 : per_lane <= 8 ? 8 : 16;
    }
    int ucache = (int)((budget - upd) / col);
    if (ucache > kcap) ucache = kcap;
    size_t smem = wrk + (size_t)wcache * col;
    const size_t us = upd + (size_t)ucache * col;
    if (us > smem) smem = us;
    if (out) { out->workers = workers; out->wcache = wcache; out->ucache = ucache; out->nvl = nvl; }
    if (smem_out) *smem_out = smem;
    return true;
}

cudaError_t launch_persist_solve(const PersistArgs& a, bool f32, size_t smem, cudaStream_t st) {
    const int NS = slots_for(a.ns);
    return f32 ? persist_launch_t<float>(a, NS, smem, st) : persist_launch_t<double>(a, NS, smem, st);
}

}  // namespace csb
