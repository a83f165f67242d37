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
//   * the two roles hand over through global memory: workers publish (|c|, atom) per signal and arrive on a counter,
//     the updater publishes the new residual and a per-signal sequence flag (release / acquire at gpu scope).  No host
//     round trip, no kernel boundary, no grid-wide barrier: one arrive and one flag per `update!`.
// Every c_j is reduced inside one warp in the order corr_gemv.cu uses (lane i takes the 16-byte vectors i, i + 32, ..;
// xor-shuffle tree), so selections agree bit for bit with the multi-launch path; the update arithmetic is
// append_atom's (same as the CTA update kernel, with this kernel's block size).
// All spin loops are bounded (PERSIST_TIMEOUT_NS): a lost CTA turns into an error status, never a hung GPU.
#include "common.cuh"
#include "gemv_loads.cuh"
#include "update_common.cuh"

#include <cstdlib>
#include <cstring>

namespace csb {
namespace {

constexpr int PT = 512;                    // threads per CTA (128 registers per thread: CG x NS accumulators fit)
constexpr int PW = PT / 32;
constexpr unsigned FLAG_STOP = 0x7fffffffu;
constexpr unsigned long long PERSIST_TIMEOUT_NS = 4000000000ull;      // 4 s per wait

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Spin until *p >= want (acquire).  Returns false on timeout or when another CTA has raised the error word.
__device__ __forceinline__ bool wait_ge(const unsigned* p, unsigned want, unsigned* err) {
    if (ld_acquire_gpu(p) >= want) return true;
    const unsigned long long t0 = timer_ns();
    for (unsigned spin = 1;; ++spin) {
        if (ld_acquire_gpu(p) >= want) return true;
        if ((spin & 255u) == 0) {
            if (ld_acquire_gpu(err) != 0u) return false;
            if (timer_ns() - t0 > PERSIST_TIMEOUT_NS) { atomicExch(err, 1u); return false; }
        }
    }
}

template <typename V> __device__ __forceinline__ V ld_shared_vec(const V* p) { return *p; }

// ---- worker: c = A'r for NS residuals over this CTA's atom range, fused |c| arg-max --------------------------------
template <typename T, int NS, int CG>
__device__ void persist_worker(const PersistArgs& a, unsigned char* smem, double (*red_v)[PW], int (*red_i)[PW],
                               unsigned* s_flag) {
    using V = typename Vec<T>::type;
    constexpr int W = Vec<T>::W;
    constexpr int UNR = CG * NS >= 32 ? 1 : (CG * NS >= 16 ? 2 : 4);   // loads in flight vs the 128-register budget
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ld = a.ld, ns = a.ns;
    const int w = (int)blockIdx.x - ns;
    const int lo = (int)((long long)a.N * w / a.workers), hi = (int)((long long)a.N * (w + 1) / a.workers);
    const int ncol = hi - lo;
    const int ncache = ncol < a.wcache ? ncol : a.wcache;
    const int nstream = ncol - ncache;
    const T* A = static_cast<const T*>(a.A);
    T* rs = reinterpret_cast<T*>(smem);                                   // [NS][ld]
    T* cache = rs + (size_t)NS * ld;                                      // [ncache][ld]: columns lo .. lo + ncache
    unsigned* arrive = a.sync;
    unsigned* flag = a.sync + 1;
    unsigned* err = a.sync + 1 + PERSIST_MAX_SIGNALS;
    const unsigned long long pol = l2_policy(L2POL_NORMAL);
    const int nvec = ld / W;                                              // ld is a multiple of 16 elements

    {   // columns kept on chip for the whole solve
        const V* src = reinterpret_cast<const V*>(A + (size_t)lo * ld);
        V* dst = reinterpret_cast<V*>(cache);
        const int total = ncache * nvec;
        for (int i = tid; i < total; i += PT) dst[i] = ldg_stream(src + i, pol);
    }
    for (int i = tid; i < (NS - ns) * ld; i += PT) rs[(size_t)ns * ld + i] = (T)0;     // unused signal slots
    const int gs = (nstream + CG - 1) / CG, gc = (ncache + CG - 1) / CG;

    for (int it = 0; it < a.k; ++it) {
        // residual version `it` of every signal (version 0 is b itself)
        if (tid < ns) {
            const bool ok = wait_ge(flag + tid, (unsigned)it, err);
            s_flag[tid] = ok ? ld_acquire_gpu(flag + tid) : FLAG_STOP;
        }
        __syncthreads();
        bool all_stop = true;
        for (int s = 0; s < ns; ++s) all_stop = all_stop && s_flag[s] == FLAG_STOP;
        if (all_stop) break;
        {
            const V* src = reinterpret_cast<const V*>(static_cast<const T*>(it == 0 ? a.B : (const void*)a.R));
            V* dst = reinterpret_cast<V*>(rs);
            for (int i = tid; i < ns * nvec; i += PT) dst[i] = __ldcg(src + i);       // L2: the updater just wrote it
        }
        __syncthreads();

        double best_v[NS];
        int best_i[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) { best_v[s] = -1.0; best_i[s] = INT_MAX; }
        for (int g = warp; g < gs + gc; g += PW) {
            const bool streamed = g < gs;                                  // streamed groups first: their loads fly longest
            const int first = streamed ? ncache + g * CG : (g - gs) * CG;  // column offset inside the range
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
                    for (int c = 0; c < CG; ++c) x[c] = ld_shared_vec(col[c] + i);
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
        if (lane == 0) {
#pragma unroll
            for (int s = 0; s < NS; ++s) { red_v[s][warp] = best_v[s]; red_i[s][warp] = best_i[s]; }
        }
        __syncthreads();
        if (tid < ns) {
            double bv = red_v[tid][0];
            int bi = red_i[tid][0];
            for (int q = 1; q < PW; ++q)
                if (cand_better(red_v[tid][q], red_i[tid][q], bv, bi)) { bv = red_v[tid][q]; bi = red_i[tid][q]; }
            a.cand_val[(size_t)tid * a.workers + w] = bv;
            a.cand_idx[(size_t)tid * a.workers + w] = (bi == INT_MAX) ? -1 : bi + a.idx_offset;
            __threadfence();
        }
        __syncthreads();
        if (tid == 0) { __threadfence(); atomicAdd(arrive, 1u); }
    }
}

// ---- updater: the loop body of `update!` for one signal, state resident in shared memory ---------------------------
template <typename T>
__device__ void persist_updater(const PersistArgs& a, unsigned char* smem, double* red, int* red_i) {
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
    T* acache = reinterpret_cast<T*>((reinterpret_cast<uintptr_t>(S.colp + kcap) + 15) & ~(uintptr_t)15);   // [ucache][ld]
    S.Tm = Tsm; S.Tsm = Tsm; S.ldT = ldT; S.Tg = nullptr; S.kcap = kcap; S.red = red;
    __shared__ int s_j, s_ok;

    const T* A = static_cast<const T*>(a.A);
    const T* b = static_cast<const T*>(a.B) + (size_t)sig * ld;
    T* rg = static_cast<T*>(a.R) + (size_t)sig * ld;
    unsigned* arrive = a.sync;
    unsigned* flag = a.sync + 1 + sig;
    unsigned* err = a.sync + 1 + PERSIST_MAX_SIGNALS;

    double s2 = 0.0;
    int bad = 0;
    for (int row = tid; row < ld; row += PT) {      // r = b: the state of a freshly constructed MP / OMP object
        const T e = b[row];
        bs[row] = (double)e; rs[row] = (double)e; rg[row] = e;
        s2 += (double)e * (double)e;
        bad |= !isfinite((double)e);
    }
    double nr = sqrt(block_sum<PT>(s2, red));
    int t = 0, flags = 0, iters = 0;
    bool done = false, failed = false;
    if (__syncthreads_or(bad)) { flags = 4; done = true; }

    auto publish = [&](unsigned value) {
        __syncthreads();                              // every thread's residual stores precede the release below
        if (tid == 0) { __threadfence(); st_release_gpu(flag, value); }
    };

    for (int it = 0; it < a.k && !done; ++it) {
        if (a.mode == 0 && !(t < a.M)) {
            // `nnz(x) < size(P.A, 1) || return x` (:63): the support is full, every remaining update! is a no-op that
            // still tests eps (:79) -- the residual no longer changes, so either the first of them breaks or none does
            iters += (nr >= a.eps) ? (a.k - it) : 1;
            if (!(nr >= a.eps)) done = true;
            break;
        }
        // candidates of this update! from every worker
        if (tid == 0) s_ok = wait_ge(arrive, (unsigned)a.workers * (unsigned)(it + 1), err) ? 1 : 0;
        __syncthreads();
        if (!s_ok) { failed = true; break; }
        double bv = -1.0;
        int bi = INT_MAX;
        for (int c = tid; c < a.workers; c += PT) {
            const double v = __ldcg(a.cand_val + (size_t)sig * a.workers + c);
            const int i = __ldcg(a.cand_idx + (size_t)sig * a.workers + c);
            if (i >= 0 && cand_better(v, i, bv, bi)) { bv = v; bi = i; }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, bv, off);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
            if (cand_better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) { red[warp] = bv; red_i[warp] = bi; }
        __syncthreads();
        if (tid == 0) {
            bv = red[0]; bi = red_i[0];
            for (int q = 1; q < PW; ++q)
                if (cand_better(red[q], red_i[q], bv, bi)) { bv = red[q]; bi = red_i[q]; }
            s_j = (bi == INT_MAX) ? -1 : bi;
        }
        __syncthreads();
        const int j = s_j;                                               // global atom index or -1

        if (a.mode == 2) {
            // mp: x[i] += <a_i, r>;  r <- r - <a_i, r> a_i   (:26-31); one (atom, increment) record per iteration
            double c = 0.0;
            if (j >= 0) {
                const T* aj = A + (size_t)(j - a.idx_offset) * ld;
                double s = 0.0;
                for (int row = tid; row < ld; row += PT) { const double e = (double)aj[row]; S.v[row] = e; s += e * rs[row]; }
                c = block_sum<PT>(s, red);                               // dot(view(A,:,i), r)  (:29)
                s2 = 0.0;
                for (int row = tid; row < ld; row += PT) {
                    const T rr = (T)(rs[row] - c * S.v[row]);
                    rs[row] = (double)rr; rg[row] = rr;
                    s2 += (double)rr * (double)rr;
                }
                nr = sqrt(block_sum<PT>(s2, red));
            } else {
                flags |= 2;
            }
            if (tid == 0) {
                a.sel[(size_t)sig * a.stride + it] = j;
                a.x[(size_t)sig * a.stride + it] = c;
            }
            ++iters;
            t = iters;
            if (it + 1 < a.k) publish((unsigned)(it + 1));
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
                // the new atom's column: copied into this CTA's shared memory while it fits (the later sweeps over the
                // active atoms then never leave the SM), read from the dictionary otherwise
                const T* aj = A + (size_t)(j - a.idx_offset) * ld;
                if (t < a.ucache) {
                    T* slot = acache + (size_t)t * ld;
                    using V = typename Vec<T>::type;
                    const int nvec = ld / Vec<T>::W;
                    for (int i = tid; i < nvec; i += PT) reinterpret_cast<V*>(slot)[i] = reinterpret_cast<const V*>(aj)[i];
                    __syncthreads();
                    aj = slot;
                }
                double nr2 = 0.0;
                const int dep = append_atom<T, PT>(
                    S, t, j, aj, ld, [&](int row) { return bs[row]; }, [&](int row) { return rs[row]; },
                    [&](int row, T val) { rs[row] = (double)val; rg[row] = val; }, nr2);
                if (dep) flags |= 1; else { changed = true; nr = sqrt(nr2); }
            }
        }
        (void)changed;
        ++iters;
        if (!(nr >= a.eps)) done = true;                                 // `norm(residual!(P, x)) >= eps || break` (:79)
        if (!done && it + 1 < a.k) publish((unsigned)(it + 1));
    }
    publish(FLAG_STOP);                                                  // releases the workers (and covers every exit path)

    // ---- results ----
    if (a.mode != 2) {
        for (int i = tid; i < t; i += PT) {                              // x_S = R^{-1} Q'b  (`ldiv!`, :175)
            double acc = 0.0;
            for (int l = i; l < t; ++l) acc = fma(Tsm[i + l * ldT], S.zs[l], acc);
            a.x[(size_t)sig * a.stride + i] = acc;
            a.sel[(size_t)sig * a.stride + i] = S.ssel[i];
        }
    }
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
    __shared__ double red_v[NS][PW];                                     // the updater uses the first row of each
    __shared__ int red_i[NS][PW];
    __shared__ unsigned s_flag[PERSIST_MAX_SIGNALS];
    if ((int)blockIdx.x < a.ns) persist_updater<T>(a, psm, &red_v[0][0], &red_i[0][0]);
    else persist_worker<T, NS, CG>(a, psm, red_v, red_i, s_flag);
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
        default: return persist_launch<T, 8, 2>(a, smem, st);
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
    int ucache = (int)((budget - upd) / col);
    if (ucache > kcap) ucache = kcap;
    size_t smem = wrk + (size_t)wcache * col;
    const size_t us = upd + (size_t)ucache * col;
    if (us > smem) smem = us;
    if (out) { out->workers = workers; out->wcache = wcache; out->ucache = ucache; }
    if (smem_out) *smem_out = smem;
    return true;
}

cudaError_t launch_persist_solve(const PersistArgs& a, bool f32, size_t smem, cudaStream_t st) {
    const int NS = slots_for(a.ns);
    return f32 ? persist_launch_t<float>(a, NS, smem, st) : persist_launch_t<double>(a, NS, smem, st);
}

}  // namespace csb
