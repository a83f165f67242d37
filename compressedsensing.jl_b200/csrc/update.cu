// Per-signal greedy-pursuit state update: subsystems (2b)-(5) of the north star.
//
// One CTA per signal.  After a correlation pass has left, per (atom block, signal), the top-s
// (|c|, atom) candidates, this kernel does what the body of the reference's `update!` does
// between two `argmaxinner!` calls (/root/reference/src/matchingpursuit.jl:62-70, 116-123):
//   1. final selection: merge the per-block candidates into the global top-`take`
//      (value descending, lower index on ties) -- `argmax` / `partialsortperm`, :184,:192;
//   2. `i in x.nzind || ...`: atoms already active are skipped (:66, util.jl:119);
//   3. `add_column!(AiQR, a, pos)` (util.jl:123): orthogonalise the new atom against the
//      active ones and append one column to the triangular factor R;
//   4. `ldiv!(AiQR, b)` (:175): x_S = R^{-1} Q'b by back substitution;
//   5. `residual!` (:152-161) and `norm(r)` for the eps test (:79,:132).  The reference recomputes
//      r = b - A_S x_S from scratch; with x_S = R^{-1}Q'b that is r = b - QQ'b, which gains exactly
//      one term per appended atom, so r <- r - q_t (q_t'b) is applied instead (one gather pass over
//      A_S less; the two differ by rounding only, ~1e-16 ||b|| per step).  r is written where the
//      next correlation pass reads it (subsystem (5): fused into the producer of the next pass).
//
// Orthogonalisation scheme.  UpdatableQRFactorizations.jl keeps an explicit (full) Q; at the
// batched shapes a thin Q is 256 KiB..1 MiB per signal (16 GiB at the headline config), so Q is
// kept IMPLICIT: Q = A_S R^{-1}.  Appending atom a:
//      g = A_S' v,  h = R^{-T} g  (= Q'v),  y = R^{-1} h,  v <- v - A_S y  (= v - Q Q'v)
// run once, and a second time when ||v|| dropped below ||v_before||/sqrt(2) (Daniel-Gragg-
// Kaufman-Stewart "twice is enough" re-orthogonalisation).  rho = ||v|| is computed from the
// explicit vector (no 1 - h'h cancellation), R[:,t] = [h; rho], z_t = <v, b>/rho.
// Instead of R the kernel stores T = R^{-1} (upper triangular, append order): appending the column
// [h; rho] to R appends [-T h / rho; 1/rho] to T, and T h is the y above, already computed.  Every
// triangular solve of the reference's `ldiv!` thereby becomes a triangular mat-vec whose output
// elements are independent dot products -- no serial substitution chain inside the CTA.
// The factor is kept in APPEND order; the reference keeps it in sorted-index order (insertion
// at `findfirst(==(i), x.nzind)`), which is the same least-squares problem with its columns
// permuted -- the host shim sorts (index, coefficient) pairs ascending when it builds the
// SparseVector.  T (kcap x kcap), z and x live in global memory per signal; the CTA that owns a signal
// stages T and one signal-length vector v in shared memory.
#include "common.cuh"
#include "update_common.cuh"
#include <cstdlib>

namespace csb {
namespace {

// The block-append working set [bm][ld] doubles as histogram (first DENSE_HIST ints) + staging area of the dense
// top-k selection, which runs before any atom is appended: never smaller than the histogram.
__host__ __device__ inline size_t block_region_elems(int bm, int ld) {
    const size_t e = (size_t)bm * ld;
    return e < (size_t)DENSE_HIST / 2 ? (size_t)DENSE_HIST / 2 : e;
}

constexpr int MAX_TAKE = 256;       // atoms per acquisition the update kernels handle (l of gomp, k of sp / oblivious / cumbabel)
constexpr int UT = 128;            // threads per CTA
constexpr int UW = UT / 32;

// Shared-memory budget for keeping the inverse factor on chip (above it the kernel works on the copy in L2).
constexpr int T_SMEM_MAX_K = 96;

// Exact <a, r> by one warp (r staged in shared memory): the loop of append_atom's g sweep.  Every lane returns the sum.
template <typename T>
__device__ __forceinline__ double warp_dot_col(const T* __restrict__ ai, const double* sr, int ld, int lane) {
    constexpr int W = RowVec<T>::W;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int row = lane * W;
    for (; row + 32 * W < ld; row += 64 * W) {
        double a0[W], a1[W];
        RowVec<T>::load(ai + row, a0);
        RowVec<T>::load(ai + row + 32 * W, a1);
#pragma unroll
        for (int e = 0; e < W; e += 2) {
            s0 = fma(a0[e], sr[row + e], s0); s1 = fma(a0[e + 1], sr[row + e + 1], s1);
            s2 = fma(a1[e], sr[row + 32 * W + e], s2); s3 = fma(a1[e + 1], sr[row + 32 * W + e + 1], s3);
        }
    }
    if (row < ld) {
        double a0[W];
        RowVec<T>::load(ai + row, a0);
#pragma unroll
        for (int e = 0; e < W; e += 2) { s0 = fma(a0[e], sr[row + e], s0); s1 = fma(a0[e + 1], sr[row + e + 1], s1); }
    }
    return warp_sum((s0 + s1) + (s2 + s3));
}

// `argmaxinner!` (/root/reference/src/matchingpursuit.jl:181-185) from the TF32 screening pass (corr_screen_tf32.cu).
// The pass left, per atom chunk, the SCREEN_T largest |c~| of this signal with |c~_j - <a_j, r>| <= E = scr_bound ||r||.
// Every atom within 2E of the largest |c~| may be the FP64 arg-max, no other atom can: those are re-evaluated exactly
// (one warp per atom) and the winner is picked with the reference's tie-break.  A single atom in the window needs no
// arithmetic at all.  A chunk whose LAST slot is still inside the window may hold more such atoms than its list shows:
// all atoms of that chunk are re-evaluated.  If ||r|| is outside the range FP32 represents safely the bound does not
// hold and every chunk is treated that way (an exact scan of the dictionary).
// sv: shared staging area for the residual (ld doubles), or nullptr: the dots read r from global memory (T = double only).
template <typename T, int NT>
__device__ void screen_select(const StateArgs& a, int sig, const T* __restrict__ A, const T* __restrict__ r, double* sv,
                              int* s_cand, double* s_cval, double* red_v, int* red_i) {
    __shared__ int s_list[SCREEN_T * SCREEN_MAX_CHUNKS];
    __shared__ int s_n;
    __shared__ unsigned s_inc;                                     // bit c: chunk c's list may be incomplete
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nc = a.scr_nc, ld = a.ld, nchunks = nc / SCREEN_T;
    const double nr = a.resnorm[sig];
    const double rsc = a.scr_f16 ? a.rscale[sig] : 1.0;            // FP16 operands: candidates are scaled by rsc * 2^sA
    const double inv_q = a.scr_f16 ? a.scr_invqA / rsc : 1.0;
    float v = -1.0f;
    int idx = -1;
    if (tid < nc) { v = a.scr_val[(size_t)sig * nc + tid]; idx = a.scr_idx[(size_t)sig * nc + tid]; }
    if (idx < 0 || !(v >= 0.0f)) { v = -1.0f; idx = -1; }
    float m = v;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
    if (tid == 0) { s_n = 0; s_inc = 0u; }
    if (lane == 0) red_v[warp] = (double)m;
    __syncthreads();
    double v0 = red_v[0];
#pragma unroll
    for (int w = 1; w < NT / 32; ++w) v0 = fmax(v0, red_v[w]);
    const bool range_ok = nr >= SCREEN_NORM_MIN && nr <= SCREEN_NORM_MAX && v0 >= 0.0 && v0 <= 3.0e38;
    v0 *= inv_q;
    const double thr = v0 - 2.0 * (a.scr_f16 ? a.scr_bound * nr + a.scr_abs / rsc : a.scr_bound * nr);
    const bool inw = idx >= 0 && (double)v * inv_q >= thr;
    const int chunk = tid / SCREEN_T;
    if (inw && (tid % SCREEN_T) == SCREEN_T - 1) atomicOr(&s_inc, 1u << chunk);
    __syncthreads();
    const unsigned inc = range_ok ? s_inc : (nchunks >= 32 ? 0xffffffffu : ((1u << nchunks) - 1u));
    if (inw && !((inc >> chunk) & 1u)) s_list[atomicAdd(&s_n, 1)] = idx;
    __syncthreads();
    const int n = s_n;
    if (inc == 0u && n == 1) {
        if (tid == 0) {
            s_cand[0] = s_list[0]; s_cval[0] = v0;
            if (a.scr_stats) atomicAdd(&a.scr_stats[0], 1ULL);
        }
        __syncthreads();
        return;
    }
    if (sv) {
        for (int row = tid; row < ld; row += NT) sv[row] = (double)r[row];
        __syncthreads();
    } else {
        sv = const_cast<double*>(reinterpret_cast<const double*>(r));   // callers pass nullptr only for T = double
    }
    double bv = -1.0;
    int bi = INT_MAX;
    for (int c = warp; c < n; c += NT / 32) {
        const int j = s_list[c];
        const double d = fabs(warp_dot_col<T>(A + (size_t)(j - a.idx_offset) * ld, sv, ld, lane));
        if (cand_better(d, j, bv, bi)) { bv = d; bi = j; }
    }
    int scanned = 0;
    for (int c = 0; c < nchunks; ++c) {
        if (!((inc >> c) & 1u)) continue;
        const int j0 = c * a.scr_chunk_atoms, j1 = min(a.N, j0 + a.scr_chunk_atoms);
        for (int j = j0 + warp; j < j1; j += NT / 32) {
            const double d = fabs(warp_dot_col<T>(A + (size_t)j * ld, sv, ld, lane));
            if (cand_better(d, j + a.idx_offset, bv, bi)) { bv = d; bi = j + a.idx_offset; }
        }
        scanned += j1 - j0;
    }
    __syncthreads();                                               // red_v was read above
    if (lane == 0) { red_v[warp] = bv; red_i[warp] = bi; }
    __syncthreads();
    if (tid == 0) {
        bv = red_v[0]; bi = red_i[0];
#pragma unroll
        for (int w = 1; w < NT / 32; ++w)
            if (cand_better(red_v[w], red_i[w], bv, bi)) { bv = red_v[w]; bi = red_i[w]; }
        s_cand[0] = bi == INT_MAX ? -1 : bi;
        s_cval[0] = bv;
        if (a.scr_stats) {
            atomicAdd(&a.scr_stats[0], 1ULL);
            atomicAdd(&a.scr_stats[1], (unsigned long long)(n + scanned));
            if (inc) atomicAdd(&a.scr_stats[2], 1ULL);
        }
    }
    __syncthreads();
}

// NT = 128 threads, 8 CTAs/SM for one atom per update (omp); NT = 256 with the block-append working set
// (BLOCK: up to `bm` new atoms orthogonalised together, update_common.cuh append_block) for gomp.
// RING_D > 0: the residual sweep of append_atom reads the active columns through a cp.async ring of RING_D columns in shared
// memory (update_common.cuh); FP64, one atom per update, ld <= RING_MAX_SLOTS * NT * 2.  Fewer CTAs fit on an SM (the ring
// is RING_D * ld * 8 bytes) but each keeps RING_D whole columns in flight without spending registers on them.
template <typename T, int NT, bool BLOCK, int RING_D = 0>
__device__ __forceinline__ void omp_update_body(const StateArgs& a, const T* __restrict__ Acache, int t_in_smem, int bm, const int sig) {
    static_assert(RING_D == 0 || (!BLOCK && sizeof(T) == 8), "the cp.async ring serves the FP64 one-atom update only");
    extern __shared__ double dsm[];
    const int ld = a.ld, kcap = a.kcap;
    PursuitSmem<T> S;
    double* p = dsm;
    double* Vb = nullptr; double* Gm = nullptr; double* Ym = nullptr; double* sc = nullptr;
    if constexpr (BLOCK) {
        Vb = p; p += block_region_elems(bm, ld);      // [bm][ld] block of new atoms / directions
        S.v = Vb;                                     // the one-by-one fallback reuses a row of the block
        Gm = p; p += (size_t)(kcap + BLOCK_MAX) * BLOCK_MAX;
        Ym = p; p += (size_t)(kcap + BLOCK_MAX) * BLOCK_MAX;
        sc = p; p += 4 * BLOCK_MAX;
    } else {
        S.v = p; p += ld;
    }
    S.g = p; p += kcap;
    S.hh = p; p += kcap;
    S.ys = p; p += kcap;
    S.y = p; p += kcap;
    S.zs = p; p += kcap;
    S.ssel = reinterpret_cast<int*>(p);
    S.colp = reinterpret_cast<const T**>(S.ssel + ((kcap + 1) & ~1));
    double* Tsm = reinterpret_cast<double*>(S.colp + kcap);        // [kcap][ldT] inverse factor (optional)
    if constexpr (RING_D > 0) {
        double* after = Tsm + (t_in_smem ? (size_t)kcap * (kcap | 1) : 0);
        S.ring = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(after) + 15) & ~(uintptr_t)15);
    }
    if (a.upd_hints & 2) S.l2_keep = l2_evict_last_policy();
    S.ring_r = static_cast<const T*>(a.R) + (size_t)sig * a.ld;
    if constexpr (!BLOCK && sizeof(T) == 8) {
        if (a.def_y && !a.resc) {
            S.def_y = a.def_y + (size_t)sig * kcap; S.def_gam = a.def_gam + sig; S.def_t = a.def_t + sig;
        }
    }
    __shared__ double red[2 * (NT / 32) + 2];
    __shared__ int red_i[NT / 32];
    __shared__ int s_cand[MAX_TAKE];
    __shared__ double s_cval[MAX_TAKE];
    __shared__ int s_J[BLOCK_MAX];
    __shared__ const T* s_Jcol[BLOCK_MAX];

    const int tid = threadIdx.x;
    if (a.slow && !a.slow[sig]) return;                            // omp_append_warp_kernel has done this signal's update
    if (a.def_t && tid == 0) a.def_t[sig] = -1;                    // nothing deferred yet
    if (a.done[sig] && !a.ignore_done) return;                     // the reference `break`s (:79,:132)
    // forward regression (`forward_step!`, src/forward.jl:56-67): same append / solve tail as omp, different
    // acquisition (candidates carry delta2 = <a,r>^2 / rescaling) and stopping rules
    const bool ols = a.resc != nullptr;
    if (ols && !(a.nnz[sig] < a.M && a.resnorm[sig] > a.max_eps)) {          // forward.jl:57,60 -> return false
        if (tid == 0) { a.done[sig] = 1; a.iters[sig] += 1; }
        return;
    }

    const T* A = static_cast<const T*>(a.A);
    const T* b = static_cast<const T*>(a.B) + (size_t)sig * ld;
    T* r = static_cast<T*>(a.R) + (size_t)sig * ld;
    // R^{-1}: global copy (column-major, ld = kcap) plus a shared working copy with an odd leading
    // dimension (conflict-free column walks) when it fits
    S.Tg = a.Rf + (size_t)sig * kcap * kcap;
    S.Tsm = t_in_smem ? Tsm : nullptr;
    S.Tm = t_in_smem ? Tsm : S.Tg;
    S.ldT = t_in_smem ? (kcap | 1) : kcap;
    S.kcap = kcap;
    S.red = red;
    int t = a.nnz[sig];
    int flags = 0;
    bool changed = false;
    double nr2 = 0.0;
    // upd_hints bit 0: signal, residual and its TF32 copy are touched once per update -> streaming loads / stores, so that
    // they do not push the dictionary (gathered t columns per signal) out of the L2
    const bool stream = a.upd_hints & 1;
    auto b_at = [&](int row) { return (double)(stream ? __ldcs(b + row) : b[row]); };
    auto r_at = [&](int row) { return (double)(stream ? __ldcs(r + row) : r[row]); };
    // TF32 (or scaled FP16) copy read by the screening pass
    void* r32 = !a.R32 ? nullptr : a.scr_f16 ? static_cast<void*>(reinterpret_cast<__half*>(a.R32) + (size_t)sig * a.ld32)
                                             : static_cast<void*>(a.R32 + (size_t)sig * a.ld32);
    const double rsc_new = a.scr_f16 ? screen_rscale(a.resnorm[sig]) : 1.0;   // from the norm BEFORE this update
    auto r_set = [&](int row, T val) {
        if (stream) __stcs(r + row, val); else r[row] = val;
        if (r32) screen_store(r32, row, (double)val, a.scr_f16, rsc_new, stream);
    };

    for (int i = tid; i < t; i += NT) {
        const int si = a.sel[(size_t)sig * kcap + i];
        S.ssel[i] = si;
        S.zs[i] = a.z[(size_t)sig * kcap + i];
        S.colp[i] = Acache ? Acache + (size_t)i * ld : A + (size_t)(si - a.idx_offset) * ld;
    }
    if (t_in_smem)
        for (int e = tid; e < t * kcap; e += NT) { const int c = e / kcap, l = e - c * kcap; if (l <= c) Tsm[l + c * S.ldT] = S.Tg[e]; }
    __syncthreads();

    if (t < a.M) {                                                 // `nnz(x) < size(P.A, 1) || return x` (:63,:117)
        if constexpr (BLOCK) {
            // the block working set is idle during the selection: its first 8 KB serve as histogram, the rest as staging
            select_any<NT>(a, sig, a.take, MAX_TAKE, s_cand, s_cval, red, red_i, reinterpret_cast<int*>(Vb),
                           Vb + DENSE_HIST / 2, (int)block_region_elems(bm, ld) - DENSE_HIST / 2);
        } else if (a.scr_val) {                                    // screened candidates, exact FP64 decision
            screen_select<T, NT>(a, sig, A, r, S.v, s_cand, s_cval, red, red_i);
        } else {                                                   // one atom per update: always per-block candidates
            const size_t cbase = (size_t)sig * a.P * a.S;
            select_candidates<NT>(a.pval + cbase, a.pidx + cbase, a.P * a.S, a.take, s_cand, s_cval, red, red_i);
        }

        int round = 0;
        while (round < a.take) {
            if constexpr (BLOCK) {
                // gather the next group of up to bm candidates that are not active yet (in |c| order)
                int m = 0;
                bool full = false;
                while (round < a.take && m < bm) {
                    const int j = s_cand[round];
                    if (j < 0) { flags |= 2; ++round; continue; }
                    int in = 0;
                    for (int i = tid; i < t; i += NT) in |= (S.ssel[i] == j);
                    if (__syncthreads_or(in)) { ++round; continue; }      // already active (:66, util.jl:119)
                    if (t + m >= kcap || t + m >= a.M) { full = true; break; }   // capacity of UpdatableQR(T, n, k)
                    if (tid == 0) { s_J[m] = j; s_Jcol[m] = A + (size_t)(j - a.idx_offset) * ld; }
                    ++m; ++round;
                }
                __syncthreads();
                if (m > 0) {
                    int done_b = 0;
                    if (m > 1) {
                        done_b = m <= 4 ? append_block<T, NT, 4>(S, t, m, s_J, s_Jcol, ld, Vb, Gm, Ym, sc, b_at, r_at, r_set, nr2,
                                                                 a.gram, a.N, a.idx_offset)
                                        : append_block<T, NT, BLOCK_MAX>(S, t, m, s_J, s_Jcol, ld, Vb, Gm, Ym, sc, b_at, r_at,
                                                                         r_set, nr2, a.gram, a.N, a.idx_offset);
                        if (done_b) changed = true;
                    }
                    for (int c = done_b; c < m; ++c) {                   // one by one: single atom, or DGKS fallback
                        const int dep = append_atom<T, NT>(S, t, s_J[c], s_Jcol[c], ld, b_at, r_at, r_set, nr2);
                        if (dep) flags |= 1; else changed = true;
                    }
                }
                if (full) break;
            } else {
                const int j = s_cand[round++];
                if (j < 0) { flags |= 2; continue; }
                if (ols && !(a.min_delta2 < s_cval[0])) {          // forward.jl:63,68: no admissible atom -> false
                    if (tid == 0) { a.done[sig] = 1; a.iters[sig] += 1; }
                    return;
                }
                int in = 0;
                for (int i = tid; i < t; i += NT) in |= (S.ssel[i] == j);
                if (__syncthreads_or(in)) continue;                // already active: nothing to add (:66, util.jl:119)
                if (t >= kcap || t >= a.M) break;                  // capacity of UpdatableQR(T, n, k)
                const T* aj = Acache ? Acache + (size_t)t * ld : A + (size_t)(j - a.idx_offset) * ld;
                const int dep = append_atom<T, NT, RING_D>(
                    S, t, j, aj, ld, b_at, r_at, r_set, nr2,
                    (a.gram && !Acache) ? a.gram + (size_t)(j - a.idx_offset) * a.N : nullptr, a.idx_offset);
                if (dep) flags |= 1; else changed = true;
                if (ols) {
                    // the atom leaves the passive set (`P.δ²[x.nzind] = 0`, :74): c^2 / Inf == 0 from now on;
                    // q_t = v / rho feeds the rescaling down-date fused into the next correlation pass
                    const double irho = dep ? 0.0 : S.Tm[(t - 1) + (t - 1) * S.ldT];
                    double* qn = a.qnew + (size_t)sig * ld;
                    for (int row = tid; row < ld; row += NT) qn[row] = S.v[row] * irho;
                    if (tid == 0) a.resc[(size_t)(j - a.idx_offset) * a.ldr + sig] = __longlong_as_double(0x7ff0000000000000LL);
                }
            }
        }
    }
    if (ols && !changed) {
        double* qn = a.qnew + (size_t)sig * ld;
        for (int row = tid; row < ld; row += NT) qn[row] = 0.0;
    }

    double nr = a.resnorm[sig];
    if (changed) {
        const bool ill = S.illcond || (a.flags[sig] & FLAG_ILLCOND);
        if (S.illcond) flags |= FLAG_ILLCOND;
        for (int i = tid; i < t; i += NT) {                        // x_S = R^{-1} Q'b  (`ldiv!`, :175)
            double acc = 0.0;
            for (int l = i; l < t; ++l) acc = fma(S.Tm[i + l * S.ldT], S.zs[l], acc);
            S.ys[i] = acc;
            a.sel[(size_t)sig * kcap + i] = S.ssel[i];
            a.z[(size_t)sig * kcap + i] = S.zs[i];
        }
        if (ill) refine_coefficients<T, NT>(S, t, ld, b_at, S.ys);   // CTA-uniform branch
        else __syncthreads();
        for (int i = tid; i < t; i += NT) a.x[(size_t)sig * kcap + i] = S.ys[i];
        nr = sqrt(nr2);
    }
    if (tid == 0) {
        a.nnz[sig] = t;
        a.resnorm[sig] = nr;
        a.iters[sig] += 1;
        if (flags) a.flags[sig] |= flags;
        if (a.scr_f16 && changed && !S.deferred) a.rscale[sig] = rsc_new;
        if (!S.deferred) {                                         // else: the last slice of the residual sweep does both
            a.resnorm[sig] = nr;
            if (!(nr >= a.eps)) a.done[sig] = 1;                   // `norm(residual!(P, x)) >= eps || break`
        }
    }
}

// Selection + append of the screened omp loop with ONE WARP PER SIGNAL (FP64 dictionary with Gram matrix, one atom per
// update, kcap <= 32, deferred residual sweep).  The CTA-per-signal body above spends its time in a chain of ~15
// barrier-separated phases and 6-7 dependent global round trips with 8 signals per SM in flight (ncu: 1.1 ms per launch
// for 65 536 signals, independent of the support size, against ~0.25 ms of HBM traffic).  Here a signal is one warp: lane i
// owns support slot i, the block reductions become shuffles, nothing waits on a CTA barrier, and 20 signals per SM are in
// flight.  The arithmetic is the CTA path's, operation for operation -- each lane plays threads lane, lane + 32, lane + 64,
// lane + 96 of the 128-thread block in the two M-length sums, and the triangular mat-vecs are per-slot dot products
// anyway -- so every result is bit-identical.  Anything off the common path (exact scan, no candidate, Pythagoras test
// failed -> DGKS, dependent or ill-conditioned atom) is left untouched and flagged in a.slow[sig]: omp_update_kernel, launched
// right after, handles exactly those signals.
constexpr int WPB = 4;                                             // warps (signals) per CTA
constexpr int WK = 32;                                             // largest support capacity of the warp path
constexpr int W_LDT = WK | 1;
struct WarpSmem {
    double T[WK * W_LDT];
    union {                                                        // the candidate list is dead before g / hh are written
        struct { double g[WK], hh[WK]; };
        int list[SCREEN_T * SCREEN_MAX_CHUNKS];
    };
    double zs[WK];
};
static_assert(sizeof(int) * SCREEN_T * SCREEN_MAX_CHUNKS <= 2 * WK * sizeof(double), "candidate list must fit under g + hh");
__global__ void __launch_bounds__(WPB * 32, 6)
omp_append_warp_kernel(StateArgs a) {
    extern __shared__ double dsm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sig = blockIdx.x * WPB + warp;
    if (sig >= a.nsig) return;
    WarpSmem& W = reinterpret_cast<WarpSmem*>(dsm)[warp];
    const int ld = a.ld, kcap = a.kcap;
    const unsigned FULL = 0xffffffffu;
    // ---- state and candidates: independent loads, issued together
    const int done = a.done[sig];
    int t = a.nnz[sig];
    const double nr = a.resnorm[sig];
    const int oldflags = a.flags[sig];
    if (lane == 0) { a.def_t[sig] = -1; a.slow[sig] = 0; }
    if (done && !a.ignore_done) return;                            // the reference `break`s (:79)
    auto leave_to_cta = [&]() { if (lane == 0) { a.slow[sig] = 1; a.slow_list[atomicAdd(a.slow_count, 1)] = sig; } };
    if (oldflags & FLAG_ILLCOND) { leave_to_cta(); return; }       // coefficients of this support are refined on every update
    const double* A = static_cast<const double*>(a.A);
    const double* b = static_cast<const double*>(a.B) + (size_t)sig * ld;
    const double* r = static_cast<const double*>(a.R) + (size_t)sig * ld;
    double* Tg = a.Rf + (size_t)sig * kcap * kcap;
    int ssel = -1;
    double zsl = 0.0;
    if (lane < t) { ssel = a.sel[(size_t)sig * kcap + lane]; zsl = a.z[(size_t)sig * kcap + lane]; }
    for (int c = 0; c < t; ++c)
        if (lane <= c) W.T[lane + c * W_LDT] = __ldcs(Tg + lane + (size_t)c * kcap);   // streamed once per update!: do not displace the dictionary
    W.zs[lane] = zsl;
    bool changed = false;
    int stat_n = -1;                                               // atoms re-evaluated in FP64 (-1: no selection ran)
    if (t < a.M) {                                                 // `nnz(x) < size(P.A, 1) || return x` (:63)
        // ---- screen_select by one warp
        const int nc = a.scr_nc;
        float v[SCREEN_T * SCREEN_MAX_CHUNKS / 32];
        int idx[SCREEN_T * SCREEN_MAX_CHUNKS / 32];
        float m = -1.0f;
#pragma unroll
        for (int q = 0; q < SCREEN_T * SCREEN_MAX_CHUNKS / 32; ++q) {
            const int e = lane + 32 * q;
            v[q] = -1.0f; idx[q] = -1;
            if (e < nc) { v[q] = a.scr_val[(size_t)sig * nc + e]; idx[q] = a.scr_idx[(size_t)sig * nc + e]; }
            if (idx[q] < 0 || !(v[q] >= 0.0f)) { v[q] = -1.0f; idx[q] = -1; }
            m = fmaxf(m, v[q]);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL, m, off));
        double v0 = (double)m;
        const bool range_ok = nr >= SCREEN_NORM_MIN && nr <= SCREEN_NORM_MAX && v0 >= 0.0 && v0 <= 3.0e38;
        const double rsc = a.scr_f16 ? a.rscale[sig] : 1.0;
        const double inv_q = a.scr_f16 ? a.scr_invqA / rsc : 1.0;
        v0 *= inv_q;
        const double thr = v0 - 2.0 * (a.scr_f16 ? a.scr_bound * nr + a.scr_abs / rsc : a.scr_bound * nr);
        unsigned inc = 0u;
        bool inw[SCREEN_T * SCREEN_MAX_CHUNKS / 32];
#pragma unroll
        for (int q = 0; q < SCREEN_T * SCREEN_MAX_CHUNKS / 32; ++q) {
            const int e = lane + 32 * q;
            inw[q] = idx[q] >= 0 && (double)v[q] * inv_q >= thr;
            if (inw[q] && (e % SCREEN_T) == SCREEN_T - 1) inc |= 1u << (e / SCREEN_T);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) inc |= __shfl_xor_sync(FULL, inc, off);
        if (!range_ok || inc != 0u) { leave_to_cta(); return; }    // exact scan of whole chunks: the CTA path
        int n = 0;
#pragma unroll
        for (int q = 0; q < SCREEN_T * SCREEN_MAX_CHUNKS / 32; ++q) {
            const unsigned bal = __ballot_sync(FULL, inw[q]);
            if (inw[q]) W.list[n + __popc(bal & ((1u << lane) - 1u))] = idx[q];
            n += __popc(bal);
        }
        __syncwarp();
        if (n == 0) { leave_to_cta(); return; }
        int j;
        if (n == 1) {
            j = W.list[0];
        } else {
            double bv = -1.0;
            int bi = INT_MAX;
            for (int c = 0; c < n; ++c) {
                const int jc = W.list[c];
                const double d = fabs(warp_dot_col<double>(A + (size_t)(jc - a.idx_offset) * ld, r, ld, lane));
                if (cand_better(d, jc, bv, bi)) { bv = d; bi = jc; }
            }
            j = bi;
        }
        stat_n = n == 1 ? 0 : n;
        const bool in = __any_sync(FULL, lane < t && ssel == j);   // already active: nothing to add (:66, util.jl:119)
        if (!in && t < kcap && t < a.M) {
            // ---- append_atom, fast path
            const double* aj = A + (size_t)(j - a.idx_offset) * ld;
            double s2[4] = {0.0, 0.0, 0.0, 0.0}, sab[4] = {0.0, 0.0, 0.0, 0.0};
            for (int row0 = 0; row0 < ld; row0 += 128) {           // lane plays threads lane + 32 w of the 128-thread block
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    const int row = row0 + lane + 32 * w;
                    if (row < ld) {
                        const double e = aj[row];
                        s2[w] = fma(e, e, s2[w]); sab[w] = fma(e, __ldcs(b + row), sab[w]);
                    }
                }
            }
#pragma unroll
            for (int w = 0; w < 4; ++w)
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    s2[w] += __shfl_xor_sync(FULL, s2[w], off);
                    sab[w] += __shfl_xor_sync(FULL, sab[w], off);
                }
            s2[0] += s2[2]; sab[0] += sab[2]; s2[1] += s2[3]; sab[1] += sab[3];
            const double anorm2 = s2[0] + s2[1], ab = sab[0] + sab[1];
            double rho2 = anorm2, yl = 0.0, qsum = 0.0;
            if (t > 0) {
                W.g[lane] = lane < t ? a.gram[(size_t)(j - a.idx_offset) * a.N + (ssel - a.idx_offset)] : 0.0;   // g = (A'A)[S, j]
                __syncwarp();
                double hl = 0.0;
                if (lane < t) {                                    // hh = R^{-T} g
                    double h0 = 0.0, h1 = 0.0, h2 = 0.0, h3 = 0.0;
                    const double* col = W.T + lane * W_LDT;
                    int l = 0;
                    for (; l + 3 <= lane; l += 4) {
                        h0 = fma(col[l], W.g[l], h0); h1 = fma(col[l + 1], W.g[l + 1], h1);
                        h2 = fma(col[l + 2], W.g[l + 2], h2); h3 = fma(col[l + 3], W.g[l + 3], h3);
                    }
                    for (; l <= lane; ++l) h0 = fma(col[l], W.g[l], h0);
                    hl = (h0 + h1) + (h2 + h3);
                }
                W.hh[lane] = hl;
                __syncwarp();
                if (lane < t) {                                    // y = R^{-1} hh
                    double h0 = 0.0, h1 = 0.0, h2 = 0.0, h3 = 0.0;
                    const double* rowp = W.T + lane;
                    int l = lane;
                    for (; l + 3 < t; l += 4) {
                        h0 = fma(rowp[l * W_LDT], W.hh[l], h0); h1 = fma(rowp[(l + 1) * W_LDT], W.hh[l + 1], h1);
                        h2 = fma(rowp[(l + 2) * W_LDT], W.hh[l + 2], h2); h3 = fma(rowp[(l + 3) * W_LDT], W.hh[l + 3], h3);
                    }
                    for (; l < t; ++l) h0 = fma(rowp[l * W_LDT], W.hh[l], h0);
                    yl = (h0 + h1) + (h2 + h3);
                }
                double p = 0.0, q = 0.0;                           // ||Q'a||^2 and <Q'a, Q'b>
                if (lane < t) { p = fma(hl, hl, p); q = fma(hl, zsl, q); }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    p += __shfl_xor_sync(FULL, p, off);
                    q += __shfl_xor_sync(FULL, q, off);
                }
                if (!(anorm2 - p >= 0.5 * anorm2)) { leave_to_cta(); return; }   // re-orthogonalisation needed: the CTA path
                rho2 = anorm2 - p;
                qsum = q;
            }
            if (!(rho2 > 1e-26 * anorm2) || rho2 < ILLCOND_RATIO * anorm2) { leave_to_cta(); return; }
            const double rho = sqrt(rho2);
            const double zt = (t > 0 ? ab - qsum : ab) / rho;
            const double gam = zt / rho;
            if (lane < t) a.def_y[(size_t)sig * kcap + lane] = yl;
            if (lane == 0) { a.def_gam[sig] = gam; a.def_t[sig] = t; }
            const double irho = 1.0 / rho;
            if (lane < t) {
                const double e = -yl * irho;
                __stcs(Tg + lane + (size_t)t * kcap, e);
                W.T[lane + t * W_LDT] = e;
            }
            if (lane == t) { __stcs(Tg + t + (size_t)t * kcap, irho); W.T[t + t * W_LDT] = irho; W.zs[t] = zt; ssel = j; zsl = zt; }
            ++t;
            changed = true;
            __syncwarp();
        }
    }
    if (changed) {
        if (lane < t) {                                            // x_S = R^{-1} Q'b  (`ldiv!`, :175)
            double acc = 0.0;
            for (int l = lane; l < t; ++l) acc = fma(W.T[lane + l * W_LDT], W.zs[l], acc);
            a.sel[(size_t)sig * kcap + lane] = ssel;
            a.z[(size_t)sig * kcap + lane] = zsl;
            a.x[(size_t)sig * kcap + lane] = acc;
        }
    }
    if (lane == 0) {
        a.nnz[sig] = t;
        a.iters[sig] += 1;
        if (!changed && !(nr >= a.eps)) a.done[sig] = 1;           // unchanged residual: `norm(residual!(P, x)) >= eps || break`
        if (a.scr_stats && stat_n >= 0) {                          // counted here: a signal left to the CTA path is counted there
            atomicAdd(&a.scr_stats[0], 1ULL);
            if (stat_n) atomicAdd(&a.scr_stats[1], (unsigned long long)stat_n);
        }
    }
}

// Deferred residual sweep of the screened omp loop:  r <- r - gamma (a_j - A_S y)  for row slots [k0, k1) of every signal
// (slot k = rows [256 k, 256 k + 256): thread tid owns rows 2 tid + 256 k and the next, exactly the assignment of
// append_atom's own sweep, and the per-row FMA chains -- even columns, odd columns -- and the per-thread sum of squares
// are continued in the same order, so r, r32 and ||r|| are bit-identical to the undeferred update).
// Why: one update! gathers t columns of 8 KiB per signal; with every signal walking whole columns the 64 MiB FP64
// dictionary does not stay in the (two-die) L2 -- ncu: 57 % hit rate, 3 GB of dictionary re-read from DRAM per launch.
// Launched per slice over ALL signals, the live part of the dictionary is 256 rows x N x 8 B (16 MiB at config 2).
template <int NT>
__global__ void __launch_bounds__(NT, 8)
omp_residual_slice_kernel(StateArgs a, int k0, int k1, int last) {
    extern __shared__ double dsm[];
    __shared__ double red[NT / 32];
    const int sig = blockIdx.x, tid = threadIdx.x;
    const int t = a.def_t[sig];
    if (t < 0) return;
    const int ld = a.ld, kcap = a.kcap;
    double* sy = dsm;                                              // [kcap]
    const double** scol = reinterpret_cast<const double**>(sy + kcap);
    const double* A = static_cast<const double*>(a.A);
    for (int i = tid; i < t; i += NT) {
        sy[i] = a.def_y[(size_t)sig * kcap + i];
        scol[i] = A + (size_t)(a.sel[(size_t)sig * kcap + i] - a.idx_offset) * ld;
    }
    const double* aj = A + (size_t)(a.sel[(size_t)sig * kcap + t] - a.idx_offset) * ld;
    const double gam = a.def_gam[sig];
    double* r = static_cast<double*>(a.R) + (size_t)sig * ld;
    void* r32 = !a.R32 ? nullptr : a.scr_f16 ? static_cast<void*>(reinterpret_cast<__half*>(a.R32) + (size_t)sig * a.ld32)
                                             : static_cast<void*>(a.R32 + (size_t)sig * a.ld32);
    const double rsc_new = a.scr_f16 ? screen_rscale(a.resnorm[sig]) : 1.0;   // resnorm still holds the norm before this update
    double s2r = k0 == 0 ? 0.0 : a.def_s2[(size_t)sig * NT + tid];
    __syncthreads();
    for (int k = k0; k < k1; ++k) {
        const int row = tid * 2 + k * NT * 2;
        if (row < ld) {
            const double2 rv = __ldcs(reinterpret_cast<const double2*>(r + row));
            const double2 av = *reinterpret_cast<const double2*>(aj + row);
            double acc[2] = {av.x, av.y}, acc1[2] = {0.0, 0.0};
            int i = 0;
#pragma unroll 4
            for (; i + 1 < t; i += 2) {                            // two chains per row, as in append_atom
                const double2 a0 = *reinterpret_cast<const double2*>(scol[i] + row);
                const double2 a1 = *reinterpret_cast<const double2*>(scol[i + 1] + row);
                const double y0 = sy[i], y1 = sy[i + 1];
                acc[0] = fma(-a0.x, y0, acc[0]); acc[1] = fma(-a0.y, y0, acc[1]);
                acc1[0] = fma(-a1.x, y1, acc1[0]); acc1[1] = fma(-a1.y, y1, acc1[1]);
            }
            if (i < t) {
                const double2 a0 = *reinterpret_cast<const double2*>(scol[i] + row);
                const double y0 = sy[i];
                acc[0] = fma(-a0.x, y0, acc[0]); acc[1] = fma(-a0.y, y0, acc[1]);
            }
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const double vq = acc[e] + acc1[e];
                const double rr = (e ? rv.y : rv.x) - gam * vq;
                __stcs(r + row + e, rr);
                if (r32) screen_store(r32, row + e, rr, a.scr_f16, rsc_new, true);
                s2r = fma(rr, rr, s2r);
            }
        }
    }
    if (!last) { a.def_s2[(size_t)sig * NT + tid] = s2r; return; }
    const double nr2 = block_sum<NT>(s2r, red);
    if (tid == 0) {
        const double nr = sqrt(nr2);
        a.resnorm[sig] = nr;
        if (a.scr_f16) a.rscale[sig] = rsc_new;
        if (!(nr >= a.eps)) a.done[sig] = 1;                       // `norm(residual!(P, x)) >= eps || break`
    }
}

// One CTA per signal (grid = nsig), or -- StateArgs::grid_cap > 0 -- a fixed number of CTAs that walk the signals with a
// grid stride.  The capped form is what runs UNDER a correlation pass (api.cu, run_omp_split): with one CTA per SM from
// the first signal to the last there is no backlog of small CTAs to flood the SMs the moment the pass's 148 large CTAs
// retire, which kept the next pass from being placed for 0.2-0.8 ms at every boundary.
template <typename T, int NT, bool BLOCK>
__global__ void __launch_bounds__(NT, BLOCK ? 2 : 8)
omp_update_kernel(StateArgs a, const T* __restrict__ Acache, int t_in_smem, int bm) {
    for (int sig = blockIdx.x; sig < a.nsig; sig += gridDim.x) {
        omp_update_body<T, NT, BLOCK>(a, Acache, t_in_smem, bm, sig);
        if (sig + (int)gridDim.x < a.nsig) __syncthreads();        // the shared-memory state is rebuilt per signal
    }
}

// The signals omp_append_warp_kernel left behind (StateArgs::slow_list, a handful per launch): a small fixed grid walks the list.
template <typename T, int NT>
__global__ void __launch_bounds__(NT, 8)
omp_update_list_kernel(StateArgs a, int t_in_smem) {
    const int n = *a.slow_count;
    for (int i = blockIdx.x; i < n; i += gridDim.x) {
        omp_update_body<T, NT, false>(a, nullptr, t_in_smem, 0, a.slow_list[i]);
        if (i + (int)gridDim.x < n) __syncthreads();
    }
}

// The same update with the cp.async column ring (FP64 dictionary, one atom per update).  MINB CTAs per SM is what the ring
// leaves room for: 4 columns of 8 KiB + 22 KiB of state -> 4 CTAs, with 128 registers per thread available.
template <int NT, int RING_D, int MINB>
__global__ void __launch_bounds__(NT, MINB)
omp_update_ring_kernel(StateArgs a, int t_in_smem) {
    for (int sig = blockIdx.x; sig < a.nsig; sig += gridDim.x) {
        omp_update_body<double, NT, false, RING_D>(a, nullptr, t_in_smem, 0, sig);
        if (sig + (int)gridDim.x < a.nsig) __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------------
// Subspace pursuit / oblivious selection (/root/reference/src/twostage.jl:67-117,
// src/oblivious.jl:3-8).  One CTA per signal runs the body of one `update!(P::SP, x)`:
//   sp_acquisition! (:87-92)   top-k of |A'r| joins the support (atoms already in it are not duplicated),
//                              least squares on the <= 2k atoms (`solve!`, :120-123)
//   pruning (:97-101)          the nnz - k entries of smallest |x| leave (`partialsortperm(abs.(x.nzval), ..)`:
//                              ties go to the lower position, i.e. the lower atom index since nzind is sorted)
//   solve! on the k kept atoms, residual, and the stopping test of `sp` (:113-115).
// first = 1 is the initial `sp_acquisition!(P, x, P.k)` of `sp` (:108) -- which is also all of `oblivious`.
// Both least-squares problems reuse the block append of gomp (two gather sweeps per 8 atoms); the pruned
// problem is re-factorised from scratch from r = b, as the reference's `factorize!` does (qr! on a copy).

template <typename T, int NT, typename BAt, typename RAt, typename RSet>
__device__ __forceinline__ int append_list(PursuitSmem<T>& S, int& t, const int* list, int count, bool skip_active,
                                           int bm, int cap, const T* __restrict__ A, int idx_offset, int ld, double* Vb,
                                           double* Gm, double* Ym, double* sc, int* s_J, const T** s_Jcol, BAt b_at,
                                           RAt r_at, RSet r_set, double& nr2, bool& changed,
                                           const double* __restrict__ gram = nullptr, int gramN = 0) {
    const int tid = threadIdx.x;
    int flags = 0, round = 0;
    while (round < count) {
        int m = 0;
        bool full = false;
        while (round < count && m < bm) {
            const int j = list[round];
            if (j < 0) { flags |= 2; ++round; continue; }
            if (skip_active) {
                int in = 0;
                for (int i = tid; i < t; i += NT) in |= (S.ssel[i] == j);
                if (__syncthreads_or(in)) { ++round; continue; }
            }
            if (t + m >= cap) { full = true; break; }
            if (tid == 0) { s_J[m] = j; s_Jcol[m] = A + (size_t)(j - idx_offset) * ld; }
            ++m; ++round;
        }
        __syncthreads();
        if (m > 0) {
            int done_b = 0;
            if (m > 1) {
                done_b = m <= 4 ? append_block<T, NT, 4>(S, t, m, s_J, s_Jcol, ld, Vb, Gm, Ym, sc, b_at, r_at, r_set, nr2, gram,
                                                         gramN, idx_offset)
                                : append_block<T, NT, BLOCK_MAX>(S, t, m, s_J, s_Jcol, ld, Vb, Gm, Ym, sc, b_at, r_at, r_set,
                                                                 nr2, gram, gramN, idx_offset);
                if (done_b) changed = true;
            }
            for (int c = done_b; c < m; ++c) {
                const int dep = append_atom<T, NT>(S, t, s_J[c], s_Jcol[c], ld, b_at, r_at, r_set, nr2);
                if (dep) flags |= 1; else changed = true;
            }
        }
        if (full) break;
    }
    return flags;
}

// CAP: slots of the candidate / keep lists (>= k): 256 covers the usual case at 4 KiB of static shared memory, 1024 the rest
template <typename T, int NT, int CAP>
__global__ void __launch_bounds__(NT, 2)
sp_update_kernel(StateArgs a, int t_in_smem, int bm, int k, double delta, int first, int* __restrict__ ndone) {
    extern __shared__ double dsm[];
    const int ld = a.ld, kcap = a.kcap;
    PursuitSmem<T> S;
    double* p = dsm;
    double* Vb = p; p += block_region_elems(bm, ld);
    S.v = Vb;
    double* Gm = p; p += (size_t)(kcap + BLOCK_MAX) * BLOCK_MAX;
    double* Ym = p; p += (size_t)(kcap + BLOCK_MAX) * BLOCK_MAX;
    double* sc = p; p += 4 * BLOCK_MAX;
    S.g = p; p += kcap;
    S.hh = p; p += kcap;
    S.ys = p; p += kcap;
    S.y = p; p += kcap;
    S.zs = p; p += kcap;
    S.ssel = reinterpret_cast<int*>(p);
    S.colp = reinterpret_cast<const T**>(S.ssel + ((kcap + 1) & ~1));
    double* Tsm = reinterpret_cast<double*>(S.colp + kcap);
    __shared__ double red[2 * (NT / 32) + 2];
    __shared__ int red_i[NT / 32];
    __shared__ int s_cand[CAP];
    __shared__ double s_cval[CAP];
    __shared__ int s_keep[CAP];
    __shared__ int s_nkeep;
    __shared__ int s_J[BLOCK_MAX];
    __shared__ const T* s_Jcol[BLOCK_MAX];

    const int sig = blockIdx.x, tid = threadIdx.x;
    if (a.done[sig]) return;
    const T* A = static_cast<const T*>(a.A);
    const T* b = static_cast<const T*>(a.B) + (size_t)sig * ld;
    T* r = static_cast<T*>(a.R) + (size_t)sig * ld;
    S.Tg = a.Rf + (size_t)sig * kcap * kcap;
    S.Tsm = t_in_smem ? Tsm : nullptr;
    S.Tm = t_in_smem ? Tsm : S.Tg;
    S.ldT = t_in_smem ? (kcap | 1) : kcap;
    S.kcap = kcap;
    S.red = red;
    int t = first ? 0 : a.nnz[sig];
    int flags = 0;
    bool changed = false;
    double nr2 = a.resnorm[sig] * a.resnorm[sig];
    auto b_at = [&](int row) { return (double)b[row]; };
    auto r_at = [&](int row) { return (double)r[row]; };
    auto r_set = [&](int row, T val) { r[row] = val; };
    for (int i = tid; i < t; i += NT) {
        const int si = a.sel[(size_t)sig * kcap + i];
        S.ssel[i] = si;
        S.zs[i] = a.z[(size_t)sig * kcap + i];
        S.colp[i] = A + (size_t)(si - a.idx_offset) * ld;
    }
    if (t_in_smem)
        for (int e = tid; e < t * kcap; e += NT) { const int c = e / kcap, l = e - c * kcap; if (l <= c) Tsm[l + c * S.ldT] = S.Tg[e]; }
    __syncthreads();

    select_any<NT>(a, sig, k, CAP, s_cand, s_cval, red, red_i, reinterpret_cast<int*>(Vb), Vb + DENSE_HIST / 2,
                   (int)block_region_elems(bm, ld) - DENSE_HIST / 2);
    const int cap = kcap < a.M ? kcap : a.M;
    flags |= append_list<T, NT>(S, t, s_cand, k, true, bm, cap, A, a.idx_offset, ld, Vb, Gm, Ym, sc, s_J, s_Jcol,
                                b_at, r_at, r_set, nr2, changed, a.gram, a.N);
    if (!first && t > k) {
        for (int i = tid; i < t; i += NT) {                       // x' = R^{-1} Q'b on the enlarged support
            double acc = 0.0;
            for (int l = i; l < t; ++l) acc = fma(S.Tm[i + l * S.ldT], S.zs[l], acc);
            S.ys[i] = acc;
        }
        if (S.illcond) refine_coefficients<T, NT>(S, t, ld, b_at, S.ys);      // the ranking below must see accurate |x|
        else __syncthreads();
        for (int i = tid; i < t; i += NT) S.y[i] = fabs(S.ys[i]);
        S.illcond = 0;                                            // the kept atoms are re-factorised from scratch
        __syncthreads();
        const int drop = t - k;
        for (int i = tid; i < t; i += NT) {                       // rank by (|x|, atom index): the `drop` smallest leave
            const double vi = S.y[i];
            const int ji = S.ssel[i];
            int rank = 0;
            for (int l = 0; l < t; ++l) rank += (S.y[l] < vi) || (S.y[l] == vi && S.ssel[l] < ji);
            S.g[i] = rank < drop ? 0.0 : 1.0;
        }
        __syncthreads();
        if (tid == 0) {
            int n = 0;
            for (int i = 0; i < t; ++i) if (S.g[i] != 0.0) s_keep[n++] = S.ssel[i];
            s_nkeep = n;
        }
        __syncthreads();
        // `solve!` on the kept atoms: fresh factorisation, residual restarted from b
        double s2 = 0.0;
        for (int row = tid; row < ld; row += NT) { const T e = b[row]; r[row] = e; s2 += (double)e * (double)e; }
        nr2 = block_sum<NT>(s2, red);
        t = 0;
        __syncthreads();
        flags |= append_list<T, NT>(S, t, s_keep, s_nkeep, false, bm, cap, A, a.idx_offset, ld, Vb, Gm, Ym, sc, s_J,
                                    s_Jcol, b_at, r_at, r_set, nr2, changed, a.gram, a.N);
    }
    for (int i = tid; i < t; i += NT) {                            // x_S = R^{-1} Q'b
        double acc = 0.0;
        for (int l = i; l < t; ++l) acc = fma(S.Tm[i + l * S.ldT], S.zs[l], acc);
        S.ys[i] = acc;
        a.sel[(size_t)sig * kcap + i] = S.ssel[i];
        a.z[(size_t)sig * kcap + i] = S.zs[i];
    }
    if (S.illcond) { flags |= FLAG_ILLCOND; refine_coefficients<T, NT>(S, t, ld, b_at, S.ys); }   // CTA-uniform
    else __syncthreads();
    for (int i = tid; i < t; i += NT) a.x[(size_t)sig * kcap + i] = S.ys[i];
    if (tid == 0) {
        const double old = a.resnorm[sig], nr = sqrt(nr2);
        a.nnz[sig] = t;
        a.resnorm[sig] = nr;
        a.iters[sig] += 1;
        if (flags) a.flags[sig] |= flags;
        // `if resnorm <= delta || oldnorm <= resnorm break` (twostage.jl:113-115); the new x is kept either way
        if (!first && (nr <= delta || old <= nr)) { a.done[sig] = 1; if (ndone) atomicAdd(ndone, 1); }
    }
}

// Matching pursuit step (/root/reference/src/matchingpursuit.jl:26-31): i = argmax |A'r|,
// x[i] += <a_i, r>.  The reference recomputes r = b - A x from scratch at the next step; because
// x changes in one entry only, that is r <- r - <a_i, r> a_i, which is what is applied here.
template <typename T>
__global__ void __launch_bounds__(UT) mp_update_kernel(StateArgs a, int iter, int stride) {
    __shared__ double red[UW];
    __shared__ int red_i[UW];
    __shared__ int s_cand[1];
    __shared__ double s_cval[1];
    const int sig = blockIdx.x, tid = threadIdx.x, ld = a.ld;
    const T* A = static_cast<const T*>(a.A);
    T* r = static_cast<T*>(a.R) + (size_t)sig * ld;
    if (a.scr_val) {                                                   // TF32-screened candidates, exact FP64 decision
        screen_select<T, UT>(a, sig, A, r, nullptr, s_cand, s_cval, red, red_i);
    } else {
        const size_t cbase = (size_t)sig * a.P * a.S;
        select_candidates<UT>(a.pval + cbase, a.pidx + cbase, a.P * a.S, 1, s_cand, s_cval, red, red_i);
    }
    const int j = s_cand[0];
    if (j < 0) {
        if (tid == 0) { a.flags[sig] |= 2; a.sel[(size_t)sig * stride + iter] = -1; a.x[(size_t)sig * stride + iter] = 0.0; }
        return;
    }
    const T* aj = A + (size_t)(j - a.idx_offset) * ld;
    float* r32 = a.R32 ? a.R32 + (size_t)sig * a.ld32 : nullptr;
    double s = 0.0;
    for (int row = tid; row < ld; row += UT) s += (double)aj[row] * (double)r[row];
    const double c = block_sum<UT>(s, red);                            // dot(view(A,:,i), r)  (:29)
    double s2 = 0.0;
    for (int row = tid; row < ld; row += UT) {
        const T rr = (T)((double)r[row] - c * (double)aj[row]);
        r[row] = rr;
        if (r32) r32[row] = tf32_round((float)rr);
        s2 += (double)rr * (double)rr;
    }
    const double nr = sqrt(block_sum<UT>(s2, red));
    if (tid == 0) {
        a.sel[(size_t)sig * stride + iter] = j;
        a.x[(size_t)sig * stride + iter] = c;
        a.nnz[sig] = iter + 1;
        a.iters[sig] = iter + 1;
        a.resnorm[sig] = nr;
    }
}

// r = b, ||b||, counters cleared.
template <typename T>
__global__ void __launch_bounds__(UT) reset_state_kernel(StateArgs a) {
    __shared__ double red[UW];
    const int sig = blockIdx.x, tid = threadIdx.x, ld = a.ld;
    const T* b = static_cast<const T*>(a.B) + (size_t)sig * ld;
    T* r = static_cast<T*>(a.R) + (size_t)sig * ld;
    double s2 = 0.0;
    void* r32 = !a.R32 ? nullptr : a.scr_f16 ? static_cast<void*>(reinterpret_cast<__half*>(a.R32) + (size_t)sig * a.ld32)
                                             : static_cast<void*>(a.R32 + (size_t)sig * a.ld32);
    for (int row = tid; row < ld; row += UT) {
        const T e = b[row]; r[row] = e; s2 += (double)e * (double)e;
        if (r32 && !a.scr_f16) screen_store(r32, row, (double)e, 0, 1.0, false);
    }
    const double nr = sqrt(block_sum<UT>(s2, red));
    if (r32 && a.scr_f16) {                                        // the scale needs ||b|| first
        const double p = screen_rscale(nr);
        for (int row = tid; row < ld; row += UT) screen_store(r32, row, (double)b[row], 1, p, false);
        if (tid == 0) a.rscale[sig] = p;
    }
    if (tid == 0) { a.nnz[sig] = 0; a.iters[sig] = 0; a.done[sig] = 0; a.flags[sig] = 0; a.resnorm[sig] = nr; }
}

// MP warm start: r = b - A x0, stored entries applied in ascending-index order as SparseArrays does.
template <typename T>
__global__ void __launch_bounds__(UT) mp_warmstart_kernel(StateArgs a, const int* __restrict__ x0_idx,
                                                          const double* __restrict__ x0_val,
                                                          const int* __restrict__ x0_nnz, int x0_stride) {
    __shared__ double red[UW];
    const int sig = blockIdx.x, tid = threadIdx.x, ld = a.ld;
    const T* A = static_cast<const T*>(a.A);
    const T* b = static_cast<const T*>(a.B) + (size_t)sig * ld;
    T* r = static_cast<T*>(a.R) + (size_t)sig * ld;
    const int n0 = x0_nnz[sig];
    double s2 = 0.0;
    for (int row = tid; row < ld; row += UT) {
        double acc = (double)b[row];
        for (int e = 0; e < n0; ++e) {
            const int j = x0_idx[(size_t)sig * x0_stride + e];
            acc -= (double)A[(size_t)(j - a.idx_offset) * ld + row] * x0_val[(size_t)sig * x0_stride + e];
        }
        const T rr = (T)acc;
        r[row] = rr;
        s2 += (double)rr * (double)rr;
    }
    const double nr = sqrt(block_sum<UT>(s2, red));
    if (tid == 0) a.resnorm[sig] = nr;
}

__global__ void __launch_bounds__(UT) topk_from_partials_kernel(StateArgs a, int s, long long* out_idx, double* out_val) {
    __shared__ double red[UW];
    __shared__ int red_i[UW];
    __shared__ int s_cand[MAX_S];
    __shared__ double s_cval[MAX_S];
    __shared__ int s_hist[DENSE_HIST];
    const int sig = blockIdx.x;
    select_any<UT>(a, sig, s, MAX_S, s_cand, s_cval, red, red_i, s_hist);
    for (int i = threadIdx.x; i < s; i += UT) {
        out_idx[(size_t)sig * s + i] = s_cand[i];
        out_val[(size_t)sig * s + i] = s_cand[i] < 0 ? 0.0 : s_cval[i];
    }
}

// ||a_j||^2 for every atom, one warp per column (`colnorms(A)` / `sum!(abs2, ...)`, src/util.jl:2, src/forward.jl:28,105).
template <typename T>
__global__ void __launch_bounds__(256) colnorm2_kernel(const T* __restrict__ A, int ld, int N, double* __restrict__ out, int take_sqrt) {
    const int j = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (j >= N) return;
    const T* a = A + (size_t)j * ld;
    double s = 0.0;
    for (int row = lane; row < ld; row += 32) { const double e = (double)a[row]; s = fma(e, e, s); }
    s = warp_sum(s);
    if (lane == 0) out[j] = take_sqrt ? sqrt(s) : s;
}

// Babel function (`cumbabel`, src/util.jl:106-117).  "Signal" s of the batch is atom col0 + s of the dictionary and its
// candidates are the per-block top-(k+1) of |A'a_i|: merge them, drop the atom itself (`inner[i] = 0`), keep the k
// largest, prefix-sum them and fold into mu[0..k) with max (non-negative doubles order like their bit patterns).
template <int CAP>
__global__ void __launch_bounds__(UT) babel_reduce_kernel(StateArgs a, int k, int col0, unsigned long long* __restrict__ mu) {
    __shared__ double red[UW];
    __shared__ int red_i[UW];
    __shared__ int s_cand[CAP];
    __shared__ double s_cval[CAP];
    __shared__ int s_hist[DENSE_HIST];
    const int sig = blockIdx.x;
    select_any<UT>(a, sig, k + 1, CAP, s_cand, s_cval, red, red_i, s_hist);
    if (threadIdx.x == 0) {
        double run = 0.0;
        int out = 0;
        for (int c = 0; c < k + 1 && out < k; ++c) {
            if (s_cand[c] == col0 + sig) continue;                 // inner product with self does not count
            if (s_cand[c] >= 0) run += s_cval[c];                  // fewer than k other atoms: the tail adds zeros
            atomicMax(mu + out, (unsigned long long)__double_as_longlong(run));
            ++out;
        }
    }
}

__global__ void __launch_bounds__(256) ols_init_kernel(const double* __restrict__ cn2, int N, long long ldr,
                                                       double* __restrict__ resc, size_t nq, double* __restrict__ qnew) {
    const size_t stride = (size_t)gridDim.x * 256, t0 = (size_t)blockIdx.x * 256 + threadIdx.x;
    for (size_t i = t0; i < (size_t)N * ldr; i += stride) resc[i] = cn2[i / ldr];
    for (size_t i = t0; i < nq; i += stride) qnew[i] = 0.0;
}

template <typename T>
__global__ void nonfinite_check_kernel(const T* __restrict__ p, size_t n, int* flag) {
    bool bad = false;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        bad |= !isfinite((double)p[i]);
    if (bad) atomicOr(flag, 1);
}

size_t update_smem_bytes(int ld, int kcap, bool t_in_smem, int bm) {
    size_t bytes = (size_t)(5 * kcap) * sizeof(double) + (size_t)((kcap + 1) & ~1) * sizeof(int) +
                   (size_t)kcap * sizeof(void*);
    if (bm > 0) bytes += (block_region_elems(bm, ld) + 2 * (size_t)(kcap + BLOCK_MAX) * BLOCK_MAX + 4 * BLOCK_MAX) * sizeof(double);
    else bytes += (size_t)ld * sizeof(double);
    if (t_in_smem) bytes += (size_t)kcap * (kcap | 1) * sizeof(double);
    return bytes;
}

// atoms orthogonalised together by the block path for this shape (0 = block path not used)
// Dynamic shared memory one CTA may use if two are to fit on an SM: 228 KB per SM, 1 KB reserved per CTA, minus the
// kernel's static shared memory (ptxas -v: 3264 B for the gomp block kernel, 4304 B for sp_update_kernel).
constexpr size_t two_cta_dyn_smem(size_t static_bytes) { return (228 * 1024 - 2 * 1024) / 2 - static_bytes; }

int block_width(int ld, int kcap, int take) {
    if (take < 2) return 0;
    const bool t_in = kcap <= T_SMEM_MAX_K;
    int bm = take < BLOCK_MAX ? take : BLOCK_MAX;
    while (bm >= 2 && update_smem_bytes(ld, kcap, t_in, bm) > two_cta_dyn_smem(3584)) --bm;   // keep 2 CTAs per SM
    return bm >= 2 ? bm : 0;
}

template <typename T>
cudaError_t launch_omp_update_t(const StateArgs& a, cudaStream_t st, const void* Acache) {
    const int t_in_smem = a.kcap <= T_SMEM_MAX_K ? 1 : 0;
    const char* env = getenv("CSB200_GOMP_BLOCK");             // test hook: 0 disables the block append
    const int bm = (Acache || (env && env[0] == '0')) ? 0 : block_width(a.ld, a.kcap, a.take);
    const size_t smem = update_smem_bytes(a.ld, a.kcap, t_in_smem != 0, bm);
    cudaError_t e;
    if (bm > 0) {
        e = cudaFuncSetAttribute(omp_update_kernel<T, 256, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        omp_update_kernel<T, 256, true><<<a.nsig, 256, smem, st>>>(a, static_cast<const T*>(Acache), t_in_smem, bm);
    } else {
        // cp.async column ring (CSB200_UPD_RING = depth 3 / 4 / 6, 0 = off): FP64, one atom per update, short signals
        const int ring_env = [] { const char* r = getenv("CSB200_UPD_RING"); return r ? atoi(r) : UPD_RING_DEFAULT; }();
        if constexpr (sizeof(T) == 8) {
            if (ring_env > 0 && !Acache && a.ld <= RING_MAX_SLOTS * UT * 2 && a.grid_cap == 0) {
                const int depth = ring_env >= 6 ? 6 : (ring_env >= 4 ? 4 : 3);
                const size_t rsmem = smem + (size_t)depth * a.ld * sizeof(double) + 16;
                auto go = [&](auto kern) -> cudaError_t {
                    cudaError_t e2 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsmem);
                    if (e2 != cudaSuccess) return e2;
                    e2 = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
                    if (e2 != cudaSuccess) return e2;
                    kern<<<a.nsig, UT, rsmem, st>>>(a, t_in_smem);
                    return cudaGetLastError();
                };
                return depth == 6 ? go(omp_update_ring_kernel<UT, 6, 3>) : depth == 4 ? go(omp_update_ring_kernel<UT, 4, 4>)
                                                                                      : go(omp_update_ring_kernel<UT, 3, 4>);
            }
        }
        e = cudaFuncSetAttribute(omp_update_kernel<T, UT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        // An SM runs CTAs of two kernels side by side only under ONE shared-memory / L1 split: when this kernel is to run
        // under the correlation kernel (which needs the largest carve-out) it must ask for the same split.
        e = cudaFuncSetAttribute(omp_update_kernel<T, UT, false>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                 a.max_smem_carveout ? (int)cudaSharedmemCarveoutMaxShared : (int)cudaSharedmemCarveoutDefault);
        if (e != cudaSuccess) return e;
        const int grid = a.grid_cap > 0 && a.grid_cap < a.nsig ? a.grid_cap : a.nsig;
        if constexpr (sizeof(T) == 8) {
            if (a.slow) {                                          // warp-per-signal selection + append first; the CTA kernel takes the rest
                e = cudaFuncSetAttribute(omp_append_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(WPB * sizeof(WarpSmem)));
                if (e != cudaSuccess) return e;
                e = cudaFuncSetAttribute(omp_append_warp_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
                if (e != cudaSuccess) return e;
                e = cudaMemsetAsync(a.slow_count, 0, sizeof(int), st);
                if (e != cudaSuccess) return e;
                omp_append_warp_kernel<<<(a.nsig + WPB - 1) / WPB, WPB * 32, WPB * sizeof(WarpSmem), st>>>(a);
                e = cudaGetLastError();
                if (e != cudaSuccess) return e;
                e = cudaFuncSetAttribute(omp_update_list_kernel<T, UT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (e != cudaSuccess) return e;
                const int lgrid = a.nsig < 148 * 4 ? a.nsig : 148 * 4;
                omp_update_list_kernel<T, UT><<<lgrid, UT, smem, st>>>(a, t_in_smem);
            }
        }
        if (!a.slow) omp_update_kernel<T, UT, false><<<grid, UT, smem, st>>>(a, static_cast<const T*>(Acache), t_in_smem, 0);
        if constexpr (sizeof(T) == 8) {
            if (a.def_y && !a.resc) {                              // the residual sweep, slice by slice over all signals
                e = cudaGetLastError();
                if (e != cudaSuccess) return e;
                const int slots = (a.ld + 2 * UT - 1) / (2 * UT);
                const int kper = a.def_kper > 0 ? a.def_kper : 1;
                const size_t ssm = (size_t)a.kcap * (sizeof(double) + sizeof(void*));
                for (int k0 = 0; k0 < slots; k0 += kper) {
                    const int k1 = k0 + kper < slots ? k0 + kper : slots;
                    omp_residual_slice_kernel<UT><<<a.nsig, UT, ssm, st>>>(a, k0, k1, k1 == slots ? 1 : 0);
                }
            }
        }
    }
    return cudaGetLastError();
}

}  // namespace

bool omp_update_uses_block(int ld, int kcap, int take) {
    const char* env = getenv("CSB200_GOMP_BLOCK");
    return !(env && env[0] == '0') && block_width(ld, kcap, take) > 0;
}

size_t omp_update_smem_bytes(int ld, int kcap) { return update_smem_bytes(ld, kcap, kcap <= T_SMEM_MAX_K, 0); }

cudaError_t launch_omp_update(const StateArgs& a, bool f32, cudaStream_t st, const void* Acache) {
    if (a.nsig <= 0) return cudaSuccess;
    return f32 ? launch_omp_update_t<float>(a, st, Acache) : launch_omp_update_t<double>(a, st, Acache);
}

int sp_block_width(int ld, int kcap) {
    const bool t_in = kcap <= T_SMEM_MAX_K;
    int bm = BLOCK_MAX;
    const size_t stat = kcap > 2 * MAX_TAKE ? 17408 : 4608;     // static shared memory of the CAP = 1024 / 256 instantiation
    while (bm >= 2 && update_smem_bytes(ld, kcap, t_in, bm) > two_cta_dyn_smem(stat)) --bm;   // 2 CTAs per SM when possible
    if (bm < 2) { bm = 2; }
    return bm;
}

size_t sp_update_smem_bytes(int ld, int kcap) {
    return update_smem_bytes(ld, kcap, kcap <= T_SMEM_MAX_K, sp_block_width(ld, kcap));
}

cudaError_t launch_sp_update(const StateArgs& a, bool f32, int k, double delta, int first, int* ndone, cudaStream_t st) {
    if (a.nsig <= 0) return cudaSuccess;
    const int t_in_smem = a.kcap <= T_SMEM_MAX_K ? 1 : 0;
    const int bm = sp_block_width(a.ld, a.kcap);
    const size_t smem = update_smem_bytes(a.ld, a.kcap, t_in_smem != 0, bm);
    auto go = [&](auto kern) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kern<<<a.nsig, 256, smem, st>>>(a, t_in_smem, bm, k, delta, first, ndone);
        return cudaGetLastError();
    };
    const bool big = k > MAX_TAKE;                              // the list capacity follows k; the block width the batch's kcap
    if (f32) return big ? go(sp_update_kernel<float, 256, SP_MAX_K>) : go(sp_update_kernel<float, 256, MAX_TAKE>);
    return big ? go(sp_update_kernel<double, 256, SP_MAX_K>) : go(sp_update_kernel<double, 256, MAX_TAKE>);
}

cudaError_t launch_colnorms(const void* A, bool f32, int ld, int N, double* out, cudaStream_t st) {
    if (N <= 0) return cudaSuccess;
    if (f32) colnorm2_kernel<float><<<(N + 7) / 8, 256, 0, st>>>(static_cast<const float*>(A), ld, N, out, 1);
    else colnorm2_kernel<double><<<(N + 7) / 8, 256, 0, st>>>(static_cast<const double*>(A), ld, N, out, 1);
    return cudaGetLastError();
}

cudaError_t launch_babel_reduce(const StateArgs& a, int k, int col0, double* mu, cudaStream_t st) {
    if (a.nsig <= 0) return cudaSuccess;
    if (k + 1 > MAX_TAKE) babel_reduce_kernel<SP_MAX_K><<<a.nsig, UT, 0, st>>>(a, k, col0, reinterpret_cast<unsigned long long*>(mu));
    else babel_reduce_kernel<MAX_TAKE><<<a.nsig, UT, 0, st>>>(a, k, col0, reinterpret_cast<unsigned long long*>(mu));
    return cudaGetLastError();
}

cudaError_t launch_ols_init(const StateArgs& a, double* colnorm2, cudaStream_t st) {
    if (a.nsig <= 0) return cudaSuccess;
    colnorm2_kernel<double><<<(a.N + 7) / 8, 256, 0, st>>>(static_cast<const double*>(a.A), a.ld, a.N, colnorm2, 0);
    ols_init_kernel<<<148 * 8, 256, 0, st>>>(colnorm2, a.N, a.ldr, a.resc, (size_t)a.nsig * a.ld, a.qnew);
    return cudaGetLastError();
}

cudaError_t launch_mp_update(const StateArgs& a, bool f32, int iter, int stride, cudaStream_t st) {
    if (a.nsig <= 0) return cudaSuccess;
    if (f32) mp_update_kernel<float><<<a.nsig, UT, 0, st>>>(a, iter, stride);
    else mp_update_kernel<double><<<a.nsig, UT, 0, st>>>(a, iter, stride);
    return cudaGetLastError();
}

// One warp that sleeps for `ns` nanoseconds: keeps a stream's next kernel from becoming runnable for that long (see
// run_omp_split in api.cu: the update of one half must not grab the SMs before the other half's correlation pass has
// placed its CTAs).
__global__ void spacer_kernel(unsigned ns) {
    if (threadIdx.x == 0) {
        unsigned long long t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        do {
            __nanosleep(2000);
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        } while (t1 - t0 < ns);
    }
}
cudaError_t launch_spacer(unsigned ns, cudaStream_t st) {
    spacer_kernel<<<1, 32, 0, st>>>(ns);
    return cudaGetLastError();
}

cudaError_t launch_reset_state(const StateArgs& a, bool f32, cudaStream_t st) {
    if (a.nsig <= 0) return cudaSuccess;
    if (f32) reset_state_kernel<float><<<a.nsig, UT, 0, st>>>(a);
    else reset_state_kernel<double><<<a.nsig, UT, 0, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_mp_warmstart(const StateArgs& a, bool f32, const int* x0_idx, const double* x0_val,
                                const int* x0_nnz, int x0_stride, cudaStream_t st) {
    if (a.nsig <= 0) return cudaSuccess;
    if (f32) mp_warmstart_kernel<float><<<a.nsig, UT, 0, st>>>(a, x0_idx, x0_val, x0_nnz, x0_stride);
    else mp_warmstart_kernel<double><<<a.nsig, UT, 0, st>>>(a, x0_idx, x0_val, x0_nnz, x0_stride);
    return cudaGetLastError();
}

cudaError_t launch_topk_from_partials(const StateArgs& a, int s, long long* out_idx, double* out_val, cudaStream_t st) {
    if (a.nsig <= 0) return cudaSuccess;
    topk_from_partials_kernel<<<a.nsig, UT, 0, st>>>(a, s, out_idx, out_val);
    return cudaGetLastError();
}

// out[col * ld_out + row] = (double) in[col * ld_in + row] for row < rows (rows >= `rows` of `out` are left alone)
__global__ void widen_f32_kernel(const float* __restrict__ in, long long ld_in, double* __restrict__ out, long long ld_out,
                                 int rows, long long cols) {
    for (long long col = blockIdx.y; col < cols; col += gridDim.y)
        for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < rows; row += gridDim.x * blockDim.x)
            out[col * ld_out + row] = (double)in[col * ld_in + row];
}

cudaError_t launch_widen_f32(const float* in, long long ld_in, double* out, long long ld_out, int rows, long long cols,
                             cudaStream_t st) {
    if (rows <= 0 || cols <= 0) return cudaSuccess;
    dim3 grid((unsigned)((rows + 255) / 256 < 8 ? (rows + 255) / 256 : 8), (unsigned)(cols < 32768 ? cols : 32768));
    widen_f32_kernel<<<grid, 256, 0, st>>>(in, ld_in, out, ld_out, rows, cols);
    return cudaGetLastError();
}

cudaError_t launch_nonfinite_check(const void* p, size_t n, bool f32, int* flag, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    const int grid = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
    if (f32) nonfinite_check_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(p), n, flag);
    else nonfinite_check_kernel<double><<<grid, 256, 0, st>>>(static_cast<const double*>(p), n, flag);
    return cudaGetLastError();
}

}  // namespace csb
