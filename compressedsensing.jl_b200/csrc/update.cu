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
// The factor is kept in APPEND order; the reference keeps it in sorted-index order (insertion
// at `findfirst(==(i), x.nzind)`), which is the same least-squares problem with its columns
// permuted -- the host shim sorts (index, coefficient) pairs ascending when it builds the
// SparseVector.  R (kcap x kcap), z and x live in global memory per signal and stay L2-resident
// for the CTA that owns them; v (one signal-length vector) lives in shared memory.
#include "common.cuh"

namespace csb {
namespace {

constexpr int UT = 128;            // threads per CTA
constexpr int UW = UT / 32;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}
// Deterministic block sum (fixed order); every thread receives the result.
__device__ __forceinline__ double block_sum(double v, double* red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < UW; ++w) s += red[w];
    return s;
}

// Global top-`take` over this signal's P*S per-block candidates -> s_cand[0..take) (atom or -1).
__device__ void select_candidates(const double* __restrict__ pv, const int* __restrict__ pi, int count, int take,
                                  int* s_cand, double* s_cval, double* red_v, int* red_i) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double prev_v = 0.0;
    int prev_i = -1;
    for (int round = 0; round < take; ++round) {
        double bv = -1.0;
        int bi = INT_MAX;
        for (int c = tid; c < count; c += UT) {
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
        for (int w = 1; w < UW; ++w)
            if (cand_better(red_v[w], red_i[w], bv, bi)) { bv = red_v[w]; bi = red_i[w]; }
        prev_v = bv; prev_i = bi;
        if (tid == 0) { s_cand[round] = (bi == INT_MAX) ? -1 : bi; s_cval[round] = bv; }
        if (bi == INT_MAX) {            // candidates exhausted: pad the rest
            for (int r2 = round + 1 + tid; r2 < take; r2 += UT) s_cand[r2] = -1;
            break;
        }
    }
    __syncthreads();
}

// Column-oriented triangular solves on one warp; R is column-major with leading dimension kcap.
// forward:  solve R' h = g   (g is destroyed);   backward: solve R y = w   (w is destroyed).
__device__ __forceinline__ void warp_forward_RT(const double* R, int kcap, int t, double* g, double* h, int lane) {
    for (int l = 0; l < t; ++l) {
        const double hl = g[l] / R[l + (size_t)l * kcap];
        __syncwarp();
        if (lane == 0) h[l] = hl;
        for (int i = l + 1 + lane; i < t; i += 32) g[i] -= R[l + (size_t)i * kcap] * hl;
        __syncwarp();
    }
}
__device__ __forceinline__ void warp_backward_R(const double* R, int kcap, int t, double* w, double* y, int lane) {
    for (int l = t - 1; l >= 0; --l) {
        const double yl = w[l] / R[l + (size_t)l * kcap];
        __syncwarp();
        if (lane == 0) y[l] = yl;
        for (int i = lane; i < l; i += 32) w[i] -= R[i + (size_t)l * kcap] * yl;
        __syncwarp();
    }
}

template <typename T>
__global__ void __launch_bounds__(UT) omp_update_kernel(StateArgs a, const T* __restrict__ Acache) {
    extern __shared__ double dsm[];
    const int ld = a.ld, kcap = a.kcap;
    double* v = dsm;                 // [ld]   working vector
    double* g = v + ld;              // [kcap]
    double* hh = g + kcap;           // [kcap]  Q'v of the current sweep
    double* h = hh + kcap;           // [kcap]  accumulated Q'a  -> new column of R
    double* w = h + kcap;            // [kcap]  scratch for back substitution
    double* y = w + kcap;            // [kcap]
    double* zs = y + kcap;           // [kcap]  Q'b
    double* xs = zs + kcap;          // [kcap]  coefficients
    int* ssel = reinterpret_cast<int*>(xs + kcap);   // [kcap] support, selection order
    __shared__ double red[UW];
    __shared__ int red_i[UW];
    __shared__ int s_cand[MAX_S];
    __shared__ double s_cval[MAX_S];

    const int sig = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (a.done[sig] && !a.ignore_done) return;                     // the reference `break`s (:79,:132)

    const T* A = static_cast<const T*>(a.A);
    const T* b = static_cast<const T*>(a.B) + (size_t)sig * ld;
    T* r = static_cast<T*>(a.R) + (size_t)sig * ld;
    double* Rf = a.Rf + (size_t)sig * kcap * kcap;
    int t = a.nnz[sig];
    int flags = 0;
    bool changed = false;
    double nr2 = 0.0;

    for (int i = tid; i < t; i += UT) { ssel[i] = a.sel[(size_t)sig * kcap + i]; zs[i] = a.z[(size_t)sig * kcap + i]; }
    __syncthreads();

    // column of the i-th active atom / of candidate j about to become the t-th
    auto active_col = [&](int i) -> const T* {
        return Acache ? Acache + (size_t)i * ld : A + (size_t)(ssel[i] - a.idx_offset) * ld;
    };

    if (t < a.M) {                                                 // `nnz(x) < size(P.A, 1) || return x` (:63,:117)
        const size_t cbase = (size_t)sig * a.P * a.S;
        select_candidates(a.pval + cbase, a.pidx + cbase, a.P * a.S, a.take, s_cand, s_cval, red, red_i);

        for (int round = 0; round < a.take; ++round) {
            const int j = s_cand[round];
            if (j < 0) { flags |= 2; continue; }
            int in = 0;
            for (int i = tid; i < t; i += UT) in |= (ssel[i] == j);
            if (__syncthreads_or(in)) continue;                    // already active: nothing to add (:66, util.jl:119)
            if (t >= kcap || t >= a.M) break;                      // capacity of UpdatableQR(T, n, k)

            const T* aj = Acache ? Acache + (size_t)t * ld : A + (size_t)(j - a.idx_offset) * ld;
            double s2 = 0.0;
            for (int row = tid; row < ld; row += UT) { const double e = (double)aj[row]; v[row] = e; s2 += e * e; }
            const double anorm2 = block_sum(s2, red);
            double before2 = anorm2, rho2 = anorm2;
            for (int sweep = 0; sweep < 2 && t > 0; ++sweep) {
                for (int i = warp; i < t; i += UW) {               // g = A_S' v
                    const T* ai = active_col(i);
                    double s = 0.0;
                    for (int row = lane; row < ld; row += 32) s += (double)ai[row] * v[row];
                    s = warp_sum(s);
                    if (lane == 0) g[i] = s;
                }
                __syncthreads();
                if (warp == 0) {
                    warp_forward_RT(Rf, kcap, t, g, hh, lane);     // hh = R^{-T} g = Q'v
                    for (int i = lane; i < t; i += 32) w[i] = hh[i];
                    __syncwarp();
                    warp_backward_R(Rf, kcap, t, w, y, lane);      // y = R^{-1} hh
                }
                __syncthreads();
                s2 = 0.0;
                for (int row = tid; row < ld; row += UT) {         // v -= A_S y
                    double acc = v[row];
                    for (int i = 0; i < t; ++i) acc -= (double)active_col(i)[row] * y[i];
                    v[row] = acc;
                    s2 += acc * acc;
                }
                for (int i = tid; i < t; i += UT) h[i] = sweep ? h[i] + hh[i] : hh[i];
                rho2 = block_sum(s2, red);
                if (rho2 >= 0.5 * before2) break;                  // DGKS: one sweep was enough
                before2 = rho2;
            }
            if (!(rho2 > 1e-26 * anorm2)) { flags |= 1; continue; }   // numerically dependent atom: not appended
            const double rho = sqrt(rho2);
            double sb = 0.0;
            for (int row = tid; row < ld; row += UT) sb += v[row] * (double)b[row];
            const double zt = block_sum(sb, red) / rho;            // z_t = q_t' b
            // residual: r = b - Q Q'b gains one term, r <- r - q_t z_t.  Identical to the reference's
            // from-scratch b - A_S x_S (x_S = R^{-1} Q'b) up to rounding, at one pass less over A_S.
            const double gam = zt / rho;
            double s2r = 0.0;
            for (int row = tid; row < ld; row += UT) {
                const T rr = (T)((double)r[row] - gam * v[row]);
                r[row] = rr;
                s2r += (double)rr * (double)rr;
            }
            nr2 = block_sum(s2r, red);
            for (int i = tid; i < t; i += UT) Rf[i + (size_t)t * kcap] = h[i];
            if (tid == 0) { Rf[t + (size_t)t * kcap] = rho; zs[t] = zt; ssel[t] = j; }
            ++t;
            changed = true;
            __syncthreads();
        }
    }

    double nr = a.resnorm[sig];
    if (changed) {
        if (warp == 0) {                                           // x_S = R^{-1} Q'b  (`ldiv!`, :175)
            for (int i = lane; i < t; i += 32) w[i] = zs[i];
            __syncwarp();
            warp_backward_R(Rf, kcap, t, w, xs, lane);
        }
        __syncthreads();
        nr = sqrt(nr2);
        for (int i = tid; i < t; i += UT) {
            a.sel[(size_t)sig * kcap + i] = ssel[i];
            a.z[(size_t)sig * kcap + i] = zs[i];
            a.x[(size_t)sig * kcap + i] = xs[i];
        }
    }
    if (tid == 0) {
        a.nnz[sig] = t;
        a.resnorm[sig] = nr;
        a.iters[sig] += 1;
        if (flags) a.flags[sig] |= flags;
        if (!(nr >= a.eps)) a.done[sig] = 1;                       // `norm(residual!(P, x)) >= eps || break`
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
    const size_t cbase = (size_t)sig * a.P * a.S;
    select_candidates(a.pval + cbase, a.pidx + cbase, a.P * a.S, 1, s_cand, s_cval, red, red_i);
    const int j = s_cand[0];
    if (j < 0) {
        if (tid == 0) { a.flags[sig] |= 2; a.sel[(size_t)sig * stride + iter] = -1; a.x[(size_t)sig * stride + iter] = 0.0; }
        return;
    }
    const T* aj = A + (size_t)(j - a.idx_offset) * ld;
    double s = 0.0;
    for (int row = tid; row < ld; row += UT) s += (double)aj[row] * (double)r[row];
    const double c = block_sum(s, red);                            // dot(view(A,:,i), r)  (:29)
    double s2 = 0.0;
    for (int row = tid; row < ld; row += UT) {
        const T rr = (T)((double)r[row] - c * (double)aj[row]);
        r[row] = rr;
        s2 += (double)rr * (double)rr;
    }
    const double nr = sqrt(block_sum(s2, red));
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
    for (int row = tid; row < ld; row += UT) { const T e = b[row]; r[row] = e; s2 += (double)e * (double)e; }
    const double nr = sqrt(block_sum(s2, red));
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
    const double nr = sqrt(block_sum(s2, red));
    if (tid == 0) a.resnorm[sig] = nr;
}

__global__ void __launch_bounds__(UT) topk_from_partials_kernel(StateArgs a, int s, long long* out_idx, double* out_val) {
    __shared__ double red[UW];
    __shared__ int red_i[UW];
    __shared__ int s_cand[MAX_S];
    __shared__ double s_cval[MAX_S];
    const int sig = blockIdx.x;
    const size_t cbase = (size_t)sig * a.P * a.S;
    select_candidates(a.pval + cbase, a.pidx + cbase, a.P * a.S, s, s_cand, s_cval, red, red_i);
    for (int i = threadIdx.x; i < s; i += UT) {
        out_idx[(size_t)sig * s + i] = s_cand[i];
        out_val[(size_t)sig * s + i] = s_cand[i] < 0 ? 0.0 : s_cval[i];
    }
}

template <typename T>
__global__ void nonfinite_check_kernel(const T* __restrict__ p, size_t n, int* flag) {
    bool bad = false;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        bad |= !isfinite((double)p[i]);
    if (bad) atomicOr(flag, 1);
}

size_t update_smem_bytes(int ld, int kcap) { return (size_t)(ld + 7 * kcap) * sizeof(double) + (size_t)kcap * sizeof(int); }

}  // namespace

cudaError_t launch_omp_update(const StateArgs& a, bool f32, cudaStream_t st, const void* Acache) {
    if (a.nsig <= 0) return cudaSuccess;
    const size_t smem = update_smem_bytes(a.ld, a.kcap);
    cudaError_t e;
    if (f32) {
        e = cudaFuncSetAttribute(omp_update_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        omp_update_kernel<float><<<a.nsig, UT, smem, st>>>(a, static_cast<const float*>(Acache));
    } else {
        e = cudaFuncSetAttribute(omp_update_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        omp_update_kernel<double><<<a.nsig, UT, smem, st>>>(a, static_cast<const double*>(Acache));
    }
    return cudaGetLastError();
}

cudaError_t launch_mp_update(const StateArgs& a, bool f32, int iter, int stride, cudaStream_t st) {
    if (a.nsig <= 0) return cudaSuccess;
    if (f32) mp_update_kernel<float><<<a.nsig, UT, 0, st>>>(a, iter, stride);
    else mp_update_kernel<double><<<a.nsig, UT, 0, st>>>(a, iter, stride);
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

cudaError_t launch_nonfinite_check(const void* p, size_t n, bool f32, int* flag, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    const int grid = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
    if (f32) nonfinite_check_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(p), n, flag);
    else nonfinite_check_kernel<double><<<grid, 256, 0, st>>>(static_cast<const double*>(p), n, flag);
    return cudaGetLastError();
}

}  // namespace csb
