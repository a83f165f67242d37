// Whole-solve kernel for SMALL dictionaries: one CTA carries one signal through every `update!` of an
// omp / gomp / mp call in a single launch.
//
// At BASELINE config 1 (128 x 256 FP64, k = 8: a 256 KiB dictionary) the per-iteration pipeline of the
// large paths (correlation kernel + update kernel per `update!`) is pure launch latency: 17 launches of a
// few microseconds for 0.5 Mflop of arithmetic.  Here the loop of the reference
// (/root/reference/src/matchingpursuit.jl:34-40, 73-82, 126-139) runs inside the kernel:
//   correlation c = A'r: a warp takes 4 atoms at a time, lanes stride the rows (coalesced; the dictionary
//     stays in L1/L2 across iterations), FP64 accumulation, fixed reduction order;
//   |c| argmax / top-l by block reductions with exclusion (value desc, index asc: Julia `argmax`,
//     `partialsortperm(rev = true)`);
//   append_atom (update_common.cuh): implicit-Q orthogonalisation, R^{-1} column, residual down-date;
//   eps test after the update, GOMP remainder update after an eps-break, MP increments.
// Signal, residual, support, R^{-1} and Q'b stay in shared memory from the first iteration to the last;
// the only global traffic besides the dictionary is b in and the results out.
#include "common.cuh"
#include "update_common.cuh"

#include <cooperative_groups.h>
#include <cstdlib>

namespace cg = cooperative_groups;

namespace csb {
namespace {

constexpr int ST = 256;     // threads per CTA
constexpr int SW = ST / 32;

template <typename T>
__global__ void __launch_bounds__(ST, 2) small_solve_kernel(StateArgs a, SmallSolveArgs q) {
    extern __shared__ double dsm[];
    const int ld = a.ld, kcap = a.kcap, N = a.N;
    const int ldT = kcap | 1;
    PursuitSmem<T> S;
    S.v = dsm;                                   // [ld]
    double* bs = S.v + ld;                       // [ld]  signal
    double* rs = bs + ld;                        // [ld]  residual (values are T-representable)
    double* cv = rs + ld;                        // [N]   signed correlations
    S.g = cv + N;
    S.hh = S.g + kcap;
    S.ys = S.hh + kcap;
    S.y = S.ys + kcap;
    S.zs = S.y + kcap;
    double* Tsm = S.zs + kcap;                   // [kcap][ldT]
    S.ssel = reinterpret_cast<int*>(Tsm + (size_t)kcap * ldT);
    S.colp = reinterpret_cast<const T**>(S.ssel + ((kcap + 1) & ~1));
    __shared__ double red[2 * SW + 2];
    __shared__ int red_i[SW];
    __shared__ int s_cand[MAX_S];
    S.Tm = Tsm; S.Tsm = Tsm; S.ldT = ldT; S.Tg = nullptr; S.kcap = kcap; S.red = red;

    const int sig = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const T* A = static_cast<const T*>(a.A);
    const T* b = static_cast<const T*>(a.B) + (size_t)sig * ld;
    T* rg = static_cast<T*>(a.R) + (size_t)sig * ld;

    // r = b (the state of a freshly constructed MP/OMP/GOMP object); reject NaN/Inf
    double s2 = 0.0;
    int bad = 0;
    for (int row = tid; row < ld; row += ST) {
        const double e = (double)b[row];
        bs[row] = e; rs[row] = e;
        s2 += e * e;
        bad |= !isfinite(e);
    }
    double nr = sqrt(block_sum<ST>(s2, red));
    if (__syncthreads_or(bad)) {
        if (tid == 0) { a.flags[sig] = 4; a.nnz[sig] = 0; a.iters[sig] = 0; a.resnorm[sig] = nr; }
        return;
    }
    int t = 0, flags = 0, iters = 0;
    bool done = false;

    if (q.mode == 2 && q.x0_nnz) {                                       // mp warm start: r = b - A x0
        const int n0 = q.x0_nnz[sig];
        s2 = 0.0;
        for (int row = tid; row < ld; row += ST) {
            double acc = bs[row];
            for (int e = 0; e < n0; ++e)
                acc -= (double)A[(size_t)(q.x0_idx[(size_t)sig * q.x0_stride + e] - a.idx_offset) * ld + row] *
                       q.x0_val[(size_t)sig * q.x0_stride + e];
            const T rr = (T)acc;
            rs[row] = (double)rr;
            s2 += (double)rr * (double)rr;
        }
        nr = sqrt(block_sum<ST>(s2, red));
    }

    const int loop_updates = q.mode == 1 ? q.k / q.l : q.k;
    const int rem = q.mode == 1 ? q.k % q.l : 0;
    const int total_updates = loop_updates + (rem > 0 ? 1 : 0);
    for (int it = 0; it < total_updates; ++it) {
        const bool is_rem = it == loop_updates;                          // gomp remainder: runs even after a break
        if (done && !is_rem) continue;
        const int take = q.mode == 1 ? (is_rem ? rem : q.l) : 1;
        if (q.mode != 2 && !(t < a.M)) { ++iters; if (!(nr >= q.eps)) done = true; continue; }   // :63,:117

        // ---- c = A'r : 4 atoms per warp at a time, lanes over rows ----
        for (int j0 = warp * 4; j0 < N; j0 += SW * 4) {
            double acc[4] = {0.0, 0.0, 0.0, 0.0};
            const T* col[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) col[c] = A + (size_t)(j0 + c < N ? j0 + c : N - 1) * ld;
            for (int row = lane; row < ld; row += 32) {
                const double rr = rs[row];
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[c] = fma((double)col[c][row], rr, acc[c]);
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const double s = warp_sum(acc[c]);
                if (lane == 0 && j0 + c < N) cv[j0 + c] = s;
            }
        }
        __syncthreads();

        // ---- top-`take` of |c| (value desc, index asc) ----
        double pv = 0.0;
        int pi = -1;
        for (int round = 0; round < take; ++round) {
            double bv = -1.0;
            int bi = INT_MAX;
            for (int j = tid; j < N; j += ST) {
                const double v = fabs(cv[j]);
                const bool ok = (round == 0) || (v < pv) || (v == pv && j > pi);
                if (ok && v > bv) { bv = v; bi = j; }                    // j ascends: first maximum wins; NaN never wins
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, bv, off);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                if (cand_better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
            }
            __syncthreads();
            if (lane == 0) { red[warp] = bv; red_i[warp] = bi; }
            __syncthreads();
            bv = red[0]; bi = red_i[0];
#pragma unroll
            for (int w = 1; w < SW; ++w)
                if (cand_better(red[w], red_i[w], bv, bi)) { bv = red[w]; bi = red_i[w]; }
            pv = bv; pi = bi;
            if (tid == 0) s_cand[round] = (bi == INT_MAX) ? -1 : bi;
            if (bi == INT_MAX) { for (int r2 = round + 1 + tid; r2 < take; r2 += ST) s_cand[r2] = -1; break; }
        }
        __syncthreads();

        if (q.mode == 2) {
            // ---- mp: x[i] += <a_i, r>;  r <- r - <a_i, r> a_i   (:26-31) ----
            const int j = s_cand[0];
            double c = 0.0;
            if (j >= 0) {
                c = cv[j];
                const T* aj = A + (size_t)j * ld;
                s2 = 0.0;
                for (int row = tid; row < ld; row += ST) {
                    const T rr = (T)(rs[row] - c * (double)aj[row]);
                    rs[row] = (double)rr;
                    s2 += (double)rr * (double)rr;
                }
                nr = sqrt(block_sum<ST>(s2, red));
            } else {
                flags |= 2;
            }
            if (tid == 0) {
                a.sel[(size_t)sig * q.stride + it] = j < 0 ? -1 : j + a.idx_offset;
                a.x[(size_t)sig * q.stride + it] = c;
            }
            ++iters;
            t = iters;
            __syncthreads();
            continue;
        }

        // ---- omp / gomp: append the candidates that are not active yet ----
        for (int round = 0; round < take; ++round) {
            const int jl = s_cand[round];
            if (jl < 0) { flags |= 2; continue; }
            const int j = jl + a.idx_offset;
            int in = 0;
            for (int i = tid; i < t; i += ST) in |= (S.ssel[i] == j);
            if (__syncthreads_or(in)) continue;                          // :66, util.jl:119
            if (t >= kcap || t >= a.M) break;
            double nr2 = 0.0;
            const int dep = append_atom<T, ST>(
                S, t, j, A + (size_t)jl * ld, ld, [&](int row) { return bs[row]; }, [&](int row) { return rs[row]; },
                [&](int row, T val) { rs[row] = (double)val; }, nr2);
            if (dep) flags |= 1; else nr = sqrt(nr2);
        }
        ++iters;
        if (!(nr >= q.eps)) done = true;                                 // `norm(residual!(P, x)) >= eps || break`
    }

    // ---- results ----
    if (q.mode != 2) {
        for (int i = tid; i < t; i += ST) {                              // x_S = R^{-1} Q'b  (`ldiv!`, :175)
            double acc = 0.0;
            for (int l = i; l < t; ++l) acc = fma(Tsm[i + l * ldT], S.zs[l], acc);
            S.ys[i] = acc;
            a.sel[(size_t)sig * q.stride + i] = S.ssel[i];
        }
        if (S.illcond) { flags |= FLAG_ILLCOND; refine_coefficients<T, ST>(S, t, ld, [&](int row) { return bs[row]; }, S.ys); }
        else __syncthreads();
        for (int i = tid; i < t; i += ST) a.x[(size_t)sig * q.stride + i] = S.ys[i];
    }
    for (int row = tid; row < ld; row += ST) rg[row] = (T)rs[row];
    if (tid == 0) {
        a.nnz[sig] = t;
        a.iters[sig] = iters;
        a.resnorm[sig] = nr;
        a.flags[sig] = flags;
        a.done[sig] = done ? 1 : 0;
    }
}

// ------------------------------------------------------------------------------------------------------
// Cluster-resident variant for FEW signals (round 2): one thread-block cluster per signal, the dictionary split over
// the CL CTAs' shared memory for the whole solve.
//
// With one signal the one-CTA kernel above is latency-bound: every `update!` streams the 256 KiB dictionary of
// config 1 through ONE SM (L1/L2 round trips: ~10 us per update!, 84 us per solve -- slower than one CPU core, 55-72 us).
// Here CTA c of the cluster keeps atoms [N c / CL, N (c + 1) / CL) in its shared memory (32 KiB at config 1, CL = 8), so
// a correlation pass is one 4-atom group per warp from shared memory; the CTAs exchange their top-`take` candidates
// through distributed shared memory (one store per peer, double-buffered, one cluster barrier per update!) and EVERY
// CTA then runs the same deterministic update on the same data -- winner merge, column fetch from the owner's shared
// memory into a local active-atom cache, append_atom, eps test -- so the replicas stay bit-identical and no residual
// ever has to be handed back.  Same reference semantics as the kernel above (mode 0 omp, 1 gomp, 2 mp).
constexpr int CL_MAX = 8;                       // portable cluster size

struct XRec { double c; int idx; int pad; };    // one candidate: signed correlation, global atom index (-1: none)

template <typename T, int CL>
__global__ void __launch_bounds__(ST, 1) cluster_solve_kernel(StateArgs a, SmallSolveArgs q, int nloc_max) {
    cg::cluster_group cl = cg::this_cluster();
    extern __shared__ double dsm[];
    const int ld = a.ld, kcap = a.kcap, N = a.N;
    const int ldT = kcap | 1;
    const int crank = (int)cl.block_rank();
    const int lo = (int)((long long)N * crank / CL), hi = (int)((long long)N * (crank + 1) / CL);
    const int nloc = hi - lo;
    PursuitSmem<T> S;
    S.v = dsm;                                   // [ld]
    double* bs = S.v + ld;                       // [ld]  signal
    double* rs = bs + ld;                        // [ld]  residual (values are T-representable)
    double* cv = rs + ld;                        // [nloc_max] signed correlations of the local atoms
    S.g = cv + nloc_max;
    S.hh = S.g + kcap;
    S.ys = S.hh + kcap;
    S.y = S.ys + kcap;
    S.zs = S.y + kcap;
    double* Tsm = S.zs + kcap;                   // [kcap][ldT]
    XRec* xch = reinterpret_cast<XRec*>(Tsm + (size_t)kcap * ldT);           // [2][CL][MAX_S] candidate exchange
    XRec* merged = xch + 2 * CL * MAX_S;                                      // [MAX_S] this update!'s winners
    S.ssel = reinterpret_cast<int*>(merged + MAX_S);
    S.colp = reinterpret_cast<const T**>(S.ssel + ((kcap + 1) & ~1));
    T* acache = reinterpret_cast<T*>(dsm) + 0;                                // placed below (needs 16-byte alignment)
    {
        const size_t off = ((size_t)((const char*)(S.colp + kcap) - (const char*)dsm) + 15) & ~(size_t)15;
        acache = reinterpret_cast<T*>(reinterpret_cast<char*>(dsm) + off);    // [kcap][ld] columns of the active atoms
    }
    T* Asm = acache + (size_t)kcap * ld;                                      // [nloc_max][ld] this CTA's dictionary slice
    __shared__ double red[2 * SW + 2];
    __shared__ int red_i[SW];
    S.Tm = Tsm; S.Tsm = Tsm; S.ldT = ldT; S.Tg = nullptr; S.kcap = kcap; S.red = red;

    const int sig = blockIdx.x / CL;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const T* A = static_cast<const T*>(a.A);
    const T* b = static_cast<const T*>(a.B) + (size_t)sig * ld;
    T* rg = static_cast<T*>(a.R) + (size_t)sig * ld;

    {   // this CTA's dictionary slice, once
        using V = typename RowVec<T>::V16;
        const V* src = reinterpret_cast<const V*>(A + (size_t)lo * ld);
        V* dst = reinterpret_cast<V*>(Asm);
        const int total = nloc * (ld / RowVec<T>::W);
        for (int i = tid; i < total; i += ST) dst[i] = src[i];
    }
    double s2 = 0.0;
    int bad = 0;
    for (int row = tid; row < ld; row += ST) {
        const double e = (double)b[row];
        bs[row] = e; rs[row] = e;
        s2 += e * e;
        bad |= !isfinite(e);
    }
    double nr = sqrt(block_sum<ST>(s2, red));
    int t = 0, flags = 0, iters = 0;
    bool done = false;
    if (__syncthreads_or(bad)) { flags = 4; done = true; }
    cl.sync();                                   // every CTA of the cluster runs and has its slice in place

    const int loop_updates = q.mode == 1 ? q.k / q.l : q.k;
    const int rem = q.mode == 1 ? q.k % q.l : 0;
    const int total_updates = (flags & 4) ? 0 : loop_updates + (rem > 0 ? 1 : 0);
    int parity = 0;
    for (int it = 0; it < total_updates; ++it) {
        const bool is_rem = it == loop_updates;                          // gomp remainder: runs even after a break
        if (done && !is_rem) continue;
        const int take = q.mode == 1 ? (is_rem ? rem : q.l) : 1;
        if (q.mode != 2 && !(t < a.M)) { ++iters; if (!(nr >= q.eps)) done = true; continue; }   // :63,:117

        XRec* mine = xch + (size_t)parity * CL * MAX_S;
        if (take == 1) {
            // ---- one atom per update! (omp, mp): the arg-max rides in the correlation loop, one block barrier, one
            // cluster barrier, and every thread picks the winner from the CL records itself ----
            double wbv = -1.0, wbc = 0.0;
            int wbi = INT_MAX;
            for (int j0 = warp * 4; j0 < nloc; j0 += SW * 4) {
                double acc[4] = {0.0, 0.0, 0.0, 0.0};
                const T* col[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) col[c] = Asm + (size_t)(j0 + c < nloc ? j0 + c : nloc - 1) * ld;
                for (int row = lane; row < ld; row += 32) {
                    const double rr = rs[row];
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc[c] = fma((double)col[c][row], rr, acc[c]);
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const double sc = warp_sum(acc[c]);
                    const double v = fabs(sc);
                    if (j0 + c < nloc && v >= 0.0 && cand_better(v, j0 + c, wbv, wbi)) { wbv = v; wbi = j0 + c; wbc = sc; }   // NaN never wins
                }
            }
            if (lane == 0) { red[warp] = wbc; red_i[warp] = wbi; }
            __syncthreads();
            if (tid < CL) {                                              // thread p delivers this CTA's record to CTA p
                double bc = 0.0;
                int bi = INT_MAX;
                for (int w = 0; w < SW; ++w)
                    if (red_i[w] != INT_MAX && cand_better(fabs(red[w]), red_i[w], fabs(bc), bi)) { bc = red[w]; bi = red_i[w]; }
                XRec rec;
                rec.c = bi == INT_MAX ? 0.0 : bc;
                rec.idx = bi == INT_MAX ? -1 : lo + bi + a.idx_offset;
                rec.pad = 0;
                cl.map_shared_rank(mine, tid)[crank * MAX_S] = rec;
            }
            cl.sync();                                                   // every CTA's record is in every buffer
            XRec win; win.c = 0.0; win.idx = -1; win.pad = 0;
            for (int src = 0; src < CL; ++src) {
                const XRec rec = mine[src * MAX_S];
                if (rec.idx >= 0 && (win.idx < 0 || cand_better(fabs(rec.c), rec.idx, fabs(win.c), win.idx))) win = rec;
            }
            if (tid == 0) merged[0] = win;
            __syncthreads();
            parity ^= 1;
        } else {
        // ---- c = A'r over the local slice: 4 atoms per warp at a time, lanes over rows, shared memory only ----
            for (int j0 = warp * 4; j0 < nloc; j0 += SW * 4) {
                double acc[4] = {0.0, 0.0, 0.0, 0.0};
                const T* col[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) col[c] = Asm + (size_t)(j0 + c < nloc ? j0 + c : nloc - 1) * ld;
                for (int row = lane; row < ld; row += 32) {
                    const double rr = rs[row];
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc[c] = fma((double)col[c][row], rr, acc[c]);
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const double s = warp_sum(acc[c]);
                    if (lane == 0 && j0 + c < nloc) cv[j0 + c] = s;
                }
            }
            __syncthreads();

            // ---- local top-`take` of |c| (value desc, index asc), written straight into every peer's exchange buffer ----
            double pv = 0.0;
            int pi = -1;
            for (int round = 0; round < take; ++round) {
                double bv = -1.0;
                int bi = INT_MAX;
                for (int j = tid; j < nloc; j += ST) {
                    const double v = fabs(cv[j]);
                    const bool ok = (round == 0) || (v < pv) || (v == pv && j > pi);
                    if (ok && v > bv) { bv = v; bi = j; }                    // j ascends: first maximum wins; NaN never wins
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    const double ov = __shfl_xor_sync(0xffffffffu, bv, off);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                    if (cand_better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
                }
                __syncthreads();
                if (lane == 0) { red[warp] = bv; red_i[warp] = bi; }
                __syncthreads();
                bv = red[0]; bi = red_i[0];
#pragma unroll
                for (int w = 1; w < SW; ++w)
                    if (cand_better(red[w], red_i[w], bv, bi)) { bv = red[w]; bi = red_i[w]; }
                pv = bv; pi = bi;
                if (tid < CL) {                                              // thread p delivers the record to CTA p
                    XRec rec;
                    rec.c = bi == INT_MAX ? 0.0 : cv[bi];
                    rec.idx = bi == INT_MAX ? -1 : lo + bi + a.idx_offset;
                    rec.pad = 0;
                    XRec* remote = cl.map_shared_rank(mine, tid);
                    remote[crank * MAX_S + round] = rec;
                }
                if (bi == INT_MAX) {
                    if (tid < CL) {
                        XRec* remote = cl.map_shared_rank(mine, tid);
                        for (int r2 = round + 1; r2 < take; ++r2) { XRec rec; rec.c = 0.0; rec.idx = -1; rec.pad = 0; remote[crank * MAX_S + r2] = rec; }
                    }
                    break;
                }
            }
            cl.sync();                                                       // every CTA's candidates are in every buffer

            // ---- merge: global top-`take` over the CL x take records, identically on every CTA ----
            double mv = 0.0;
            int mi = -1;
            for (int round = 0; round < take; ++round) {
                double bv = -1.0;
                int bi = INT_MAX, bslot = -1;
                for (int e = tid; e < CL * take; e += ST) {
                    const int src = e / take, rr = e - src * take;
                    const XRec rec = mine[src * MAX_S + rr];
                    if (rec.idx < 0) continue;
                    const double v = fabs(rec.c);
                    const bool ok = (round == 0) || (v < mv) || (v == mv && rec.idx > mi);
                    if (ok && cand_better(v, rec.idx, bv, bi)) { bv = v; bi = rec.idx; bslot = src * MAX_S + rr; }
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    const double ov = __shfl_xor_sync(0xffffffffu, bv, off);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                    const int os = __shfl_xor_sync(0xffffffffu, bslot, off);
                    if (cand_better(ov, oi, bv, bi)) { bv = ov; bi = oi; bslot = os; }
                }
                __syncthreads();
                if (lane == 0) { red[warp] = bv; red_i[warp] = bslot; }
                __syncthreads();
                if (tid == 0) {
                    double wv = -1.0;
                    int wi = INT_MAX, ws = -1;
                    for (int w = 0; w < SW; ++w) {
                        if (red_i[w] < 0) continue;
                        const int ci = mine[red_i[w]].idx;
                        if (cand_better(red[w], ci, wv, wi)) { wv = red[w]; wi = ci; ws = red_i[w]; }
                    }
                    XRec rec; rec.c = 0.0; rec.idx = -1; rec.pad = 0;
                    if (ws >= 0) rec = mine[ws];
                    merged[round] = rec;
                }
                __syncthreads();
                mv = fabs(merged[round].c); mi = merged[round].idx;
                if (mi < 0) { for (int r2 = round + 1 + tid; r2 < take; r2 += ST) { XRec rec; rec.c = 0.0; rec.idx = -1; rec.pad = 0; merged[r2] = rec; } break; }
            }
            __syncthreads();
            parity ^= 1;
        }

        // the column of atom j from its owner's shared memory (distributed shared memory) into cache slot `slot`
        auto fetch = [&](int j, int slot) -> const T* {
            const int owner = (int)((((long long)(j - a.idx_offset) + 1) * CL - 1) / N);        // inverse of lo = N c / CL
            const int olo = (int)((long long)N * owner / CL);
            const T* src = cl.map_shared_rank(Asm, owner) + (size_t)(j - a.idx_offset - olo) * ld;
            T* dst = acache + (size_t)slot * ld;
            using V = typename RowVec<T>::V16;
            for (int i = tid; i < ld / RowVec<T>::W; i += ST) reinterpret_cast<V*>(dst)[i] = reinterpret_cast<const V*>(src)[i];
            __syncthreads();
            return dst;
        };

        if (q.mode == 2) {
            // ---- mp: x[i] += <a_i, r>;  r <- r - <a_i, r> a_i   (:26-31) ----
            const int j = merged[0].idx;
            double c = 0.0;
            if (j >= 0) {
                c = merged[0].c;
                const T* aj = fetch(j, 0);
                s2 = 0.0;
                for (int row = tid; row < ld; row += ST) {
                    const T rr = (T)(rs[row] - c * (double)aj[row]);
                    rs[row] = (double)rr;
                    s2 += (double)rr * (double)rr;
                }
                nr = sqrt(block_sum<ST>(s2, red));
            } else {
                flags |= 2;
            }
            if (crank == 0 && tid == 0) {
                a.sel[(size_t)sig * q.stride + it] = j;
                a.x[(size_t)sig * q.stride + it] = c;
            }
            ++iters;
            t = iters;
            __syncthreads();
            continue;
        }

        // ---- omp / gomp: append the candidates that are not active yet ----
        for (int round = 0; round < take; ++round) {
            const int j = merged[round].idx;
            if (j < 0) { flags |= 2; continue; }
            int in = 0;
            for (int i = tid; i < t; i += ST) in |= (S.ssel[i] == j);
            if (__syncthreads_or(in)) continue;                          // :66, util.jl:119
            if (t >= kcap || t >= a.M) break;
            const T* aj = fetch(j, t);
            double nr2 = 0.0;
            const int dep = append_atom<T, ST>(
                S, t, j, aj, ld, [&](int row) { return bs[row]; }, [&](int row) { return rs[row]; },
                [&](int row, T val) { rs[row] = (double)val; }, nr2);
            if (dep) flags |= 1; else nr = sqrt(nr2);
        }
        ++iters;
        if (!(nr >= q.eps)) done = true;                                 // `norm(residual!(P, x)) >= eps || break`
    }

    // ---- results (identical on every CTA; rank 0 writes) ----
    if (q.mode != 2) {
        for (int i = tid; i < t; i += ST) {                              // x_S = R^{-1} Q'b  (`ldiv!`, :175)
            double acc = 0.0;
            for (int l = i; l < t; ++l) acc = fma(Tsm[i + l * ldT], S.zs[l], acc);
            S.ys[i] = acc;
        }
        if (S.illcond) { flags |= FLAG_ILLCOND; refine_coefficients<T, ST>(S, t, ld, [&](int row) { return bs[row]; }, S.ys); }
        else __syncthreads();
        if (crank == 0)
            for (int i = tid; i < t; i += ST) { a.x[(size_t)sig * q.stride + i] = S.ys[i]; a.sel[(size_t)sig * q.stride + i] = S.ssel[i]; }
    }
    if (crank == 0) {
        for (int row = tid; row < ld; row += ST) rg[row] = (T)rs[row];
        if (tid == 0) {
            a.nnz[sig] = t;
            a.iters[sig] = iters;
            a.resnorm[sig] = nr;
            a.flags[sig] = flags;
            a.done[sig] = done ? 1 : 0;
        }
    }
    cl.sync();                                   // no CTA leaves while a peer may still read its dictionary slice
}

size_t cluster_solve_smem_bytes(int ld, int N, int kcap, int CL, bool f32) {
    const size_t es = f32 ? 4 : 8;
    const int nloc_max = (N + CL - 1) / CL + 1;
    size_t bytes = (size_t)(3 * ld + nloc_max + 5 * kcap + (size_t)kcap * (kcap | 1)) * sizeof(double) +
                   (size_t)(2 * CL_MAX + 1) * MAX_S * sizeof(XRec) + (size_t)((kcap + 1) & ~1) * sizeof(int) +
                   (size_t)kcap * sizeof(void*) + 16;
    bytes += ((size_t)kcap + nloc_max) * ld * es;
    return bytes;
}

size_t small_smem_bytes(int ld, int N, int kcap) {
    return (size_t)(3 * ld + N + 5 * kcap + (size_t)kcap * (kcap | 1)) * sizeof(double) +
           (size_t)((kcap + 1) & ~1) * sizeof(int) + (size_t)kcap * sizeof(void*);
}

}  // namespace

bool small_solve_eligible(int ld, int N, int kcap, int nsig, bool f32) {
    const size_t dict_bytes = (size_t)ld * N * (f32 ? 4 : 8);
    return dict_bytes <= ((size_t)2 << 20) && N <= 4096 && kcap <= 64 && nsig <= 512 &&
           small_smem_bytes(ld, N, kcap) <= 100 * 1024;
}

// Few signals on a dictionary that fits the shared memory of one 8-CTA cluster: the cluster-resident kernel.
bool cluster_solve_eligible(int ld, int N, int kcap, int nsig, int take, bool f32) {
    return nsig >= 1 && nsig <= 8 && N >= 4 * CL_MAX && N <= 4096 && kcap <= 64 && take <= MAX_S &&
           cluster_solve_smem_bytes(ld, N, kcap, CL_MAX, f32) <= 200 * 1024;
}

template <typename T>
static cudaError_t launch_cluster_solve_t(const StateArgs& a, const SmallSolveArgs& q, cudaStream_t st) {
    constexpr int CL = CL_MAX;
    const bool f32 = sizeof(T) == 4;
    const size_t smem = cluster_solve_smem_bytes(a.ld, a.N, a.kcap, CL, f32);
    cudaError_t e = cudaFuncSetAttribute(cluster_solve_kernel<T, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(a.nsig * CL));
    cfg.blockDim = dim3(ST);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    const int nloc_max = (a.N + CL - 1) / CL + 1;
    return cudaLaunchKernelEx(&cfg, cluster_solve_kernel<T, CL>, a, q, nloc_max);
}

cudaError_t launch_cluster_solve(const StateArgs& a, const SmallSolveArgs& q, bool f32, cudaStream_t st) {
    if (a.nsig <= 0) return cudaSuccess;
    return f32 ? launch_cluster_solve_t<float>(a, q, st) : launch_cluster_solve_t<double>(a, q, st);
}

cudaError_t launch_small_solve(const StateArgs& a, const SmallSolveArgs& q, bool f32, cudaStream_t st) {
    if (a.nsig <= 0) return cudaSuccess;
    const size_t smem = small_smem_bytes(a.ld, a.N, a.kcap);
    cudaError_t e;
    if (f32) {
        e = cudaFuncSetAttribute(small_solve_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        small_solve_kernel<float><<<a.nsig, ST, smem, st>>>(a, q);
    } else {
        e = cudaFuncSetAttribute(small_solve_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        small_solve_kernel<double><<<a.nsig, ST, smem, st>>>(a, q);
    }
    return cudaGetLastError();
}

}  // namespace csb
