// Per-signal state update for the SINGLE-SIGNAL (few-signal) paths: one thread-block CLUSTER per signal.
//
// Same mathematics as omp_update_kernel (update.cu: final selection, implicit-Q append with DGKS
// re-orthogonalisation, R^{-1} update, coefficients, residual down-date -- reference
// /root/reference/src/matchingpursuit.jl:62-70, 116-123, 152-176 and src/util.jl:118-126), but where the
// batched path has tens of thousands of signals to fill the GPU with one CTA each, a single-signal solve
// (BASELINE configs 1 and 4: M up to 8192, up to 128 active atoms) left ONE CTA to gather 2 t M elements per
// iteration and the update took as long as the HBM-bound correlation pass it sits behind.
//
// Here the M rows are split over the CL CTAs of a cluster (8, the portable maximum, or 16 = the opt-in non-portable
// size for single-signal solves on long atoms, where the gather sweeps are bound by how many SMs pull from L2).  Each CTA keeps its slice of v / r / b, gathers
// its slice of the active atoms, and the length-M reductions (A_S'v, ||v||^2, v'b, ||r||^2) are finished
// across the cluster through distributed shared memory: every CTA stores its partials into every peer's
// exchange buffer (`map_shared_rank`), `cluster.sync()`, then sums the 8 partials in rank order -- a
// deterministic all-reduce that costs one hardware cluster barrier.  The small state (support, R^{-1}, Q'b)
// is replicated per CTA and updated identically, so no broadcast is needed.  Three exchanges per appended
// atom (four when a second orthogonalisation sweep is required): {A_S'v, ||a||^2}, {||v||^2, v'b}, {||r||^2}.
#include <cooperative_groups.h>

#include <atomic>
#include <cstdint>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "update_common.cuh"

namespace cg = cooperative_groups;

namespace csb {
namespace {

constexpr int CT = 256;            // threads per CTA
constexpr int T_SMEM_MAX_K_CL = 128;   // 128 x 129 doubles = 132 KB: fits beside the 1-CTA-per-SM working set

struct Exchange {
    double* buf;       // [2][CL][stride] in this CTA's shared memory
    int stride;
    int parity;
};

// Deterministic cluster-wide sum of n doubles held in `vals` (shared memory of each CTA); result in `out`
// (shared memory), identical on every CTA.
template <int CL>
__device__ __forceinline__ void cluster_allreduce(cg::cluster_group& cl, Exchange& x, const double* vals, int n,
                                                  double* out) {
    const int tid = threadIdx.x;
    const unsigned me = cl.block_rank();
    double* mine = x.buf + (size_t)x.parity * CL * x.stride;
    __syncthreads();                                   // vals complete
    for (int dst = 0; dst < CL; ++dst) {
        double* remote = cl.map_shared_rank(mine, dst);
        for (int i = tid; i < n; i += CT) remote[me * x.stride + i] = vals[i];
    }
    cl.sync();                                         // release/acquire: peers' stores are visible
    for (int i = tid; i < n; i += CT) {
        double s = 0.0;
#pragma unroll
        for (int c = 0; c < CL; ++c) s += mine[c * x.stride + i];
        out[i] = s;
    }
    x.parity ^= 1;
    __syncthreads();
}

template <typename T, int CL>
__global__ void __launch_bounds__(CT, 1) omp_update_cluster_kernel(StateArgs a, const T* __restrict__ Acache,
                                                                   int t_in_smem) {
    cg::cluster_group cl = cg::this_cluster();
    extern __shared__ double dsm[];
    const int ld = a.ld, kcap = a.kcap;
    const int Mc = ld / CL;                                    // rows owned by this CTA (ld is a multiple of 16)
    const int crank = (int)cl.block_rank();
    const int row0 = crank * Mc;
    const int xstride = kcap + 4;
    double* v = dsm;                 // [Mc]
    double* g = v + Mc;              // [kcap + 4] partials to exchange: g[0..t), then scalars
    double* gs = g + xstride;        // [kcap + 4] reduced
    double* hh = gs + xstride;       // [kcap]
    double* ys = hh + kcap;          // [kcap]
    double* y = ys + kcap;           // [kcap]
    double* zs = y + kcap;           // [kcap]
    double* xbuf = zs + kcap;        // [2][CL][xstride]
    int* ssel = reinterpret_cast<int*>(xbuf + 2 * CL * xstride);
    const T** colp = reinterpret_cast<const T**>(ssel + ((kcap + 1) & ~1));
    double* Tsm = reinterpret_cast<double*>(colp + kcap);
    __shared__ double red[CT / 32];
    __shared__ int red_i[CT / 32];
    __shared__ int s_cand[GOMP_MAX_L];
    __shared__ double s_cval[GOMP_MAX_L];

    const int sig = blockIdx.x / CL;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int W = RowVec<T>::W;                            // elements per 16-byte load
    const bool vec = (Mc % W) == 0;                            // every CTA's row slice starts 16-byte aligned
    // uniform over the cluster: every CTA of a cluster reads the same flag
    if (a.done[sig] && !a.ignore_done) return;
    cl.sync();      // every CTA of the cluster is running (its shared memory exists) before any remote store

    Exchange ex{xbuf, xstride, 0};
    const T* A = static_cast<const T*>(a.A);
    const T* b = static_cast<const T*>(a.B) + (size_t)sig * ld + row0;
    T* r = static_cast<T*>(a.R) + (size_t)sig * ld + row0;
    double* Tg = a.Rf + (size_t)sig * kcap * kcap;
    double* Tm = t_in_smem ? Tsm : Tg;
    const int ldT = t_in_smem ? (kcap | 1) : kcap;
    int t = a.nnz[sig];
    int flags = 0;
    bool changed = false;
    double nr2 = 0.0;

    for (int i = tid; i < t; i += CT) {
        const int si = a.sel[(size_t)sig * kcap + i];
        ssel[i] = si;
        zs[i] = a.z[(size_t)sig * kcap + i];
        colp[i] = (Acache ? Acache + (size_t)i * ld : A + (size_t)(si - a.idx_offset) * ld) + row0;
    }
    if (t_in_smem)
        for (int e = tid; e < t * kcap; e += CT) { const int c = e / kcap, l = e - c * kcap; if (l <= c) Tsm[l + c * ldT] = Tg[e]; }
    __syncthreads();

    if (t < a.M) {
        const size_t cbase = (size_t)sig * a.P * a.S;
        select_candidates<CT>(a.pval + cbase, a.pidx + cbase, a.P * a.S, a.take, s_cand, s_cval, red, red_i);

        for (int round = 0; round < a.take; ++round) {
            const int j = s_cand[round];
            if (j < 0) { flags |= 2; continue; }
            int in = 0;
            for (int i = tid; i < t; i += CT) in |= (ssel[i] == j);
            if (__syncthreads_or(in)) continue;
            if (t >= kcap || t >= a.M) break;

            const T* aj = (Acache ? Acache + (size_t)t * ld : A + (size_t)(j - a.idx_offset) * ld) + row0;
            double s2 = 0.0, sb0 = 0.0;
            for (int row = tid; row < Mc; row += CT) {
                const double e = (double)aj[row];
                v[row] = e; s2 += e * e;
                if (t == 0) sb0 = fma(e, (double)b[row], sb0);             // first atom: v = a_j is final, v'b rides along
            }
            s2 = block_sum<CT>(s2, red);               // this CTA's part of ||a||^2 (syncs: v is complete)
            if (t == 0) sb0 = block_sum<CT>(sb0, red);
            double anorm2 = 0.0, before2 = 0.0, rho2 = 0.0, vb = 0.0;
            bool have_norm = false;
            for (int sweep = 0; sweep < 2; ++sweep) {
                if (t > 0 && vec) {
                    // partial g = A_S[rows]' v[rows]: 16-byte loads, four atoms per warp in flight (the sweep is
                    // latency-bound: 8 CTAs pull t x Mc elements out of L2 with one dependent chain per lane)
                    for (int i0 = warp * 4; i0 < t; i0 += (CT / 32) * 4) {
                        double s[4] = {0.0, 0.0, 0.0, 0.0};
                        const T* ai[4];
#pragma unroll
                        for (int c = 0; c < 4; ++c) ai[c] = colp[i0 + c < t ? i0 + c : t - 1];
#pragma unroll 2
                        for (int row = lane * W; row < Mc; row += 32 * W) {
                            double av[4][W];
#pragma unroll
                            for (int c = 0; c < 4; ++c) RowVec<T>::load(ai[c] + row, av[c]);
#pragma unroll
                            for (int e = 0; e < W; ++e) {
                                const double ve = v[row + e];
#pragma unroll
                                for (int c = 0; c < 4; ++c) s[c] = fma(av[c][e], ve, s[c]);
                            }
                        }
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const double r4 = warp_sum(s[c]);
                            if (lane == 0 && i0 + c < t) g[i0 + c] = r4;
                        }
                    }
                } else if (t > 0) {
                    for (int i = warp; i < t; i += CT / 32) {
                        const T* ai = colp[i];
                        double s = 0.0;
                        for (int row = lane; row < Mc; row += 32) s += (double)ai[row] * v[row];
                        s = warp_sum(s);
                        if (lane == 0) g[i] = s;
                    }
                }
                if (tid == 0) { g[t] = s2; g[t + 1] = sb0; }               // ride along: ||a||^2 partial (sweep 0), v'b (t = 0)
                cluster_allreduce<CL>(cl, ex, g, t + 2, gs);
                if (!have_norm) { anorm2 = gs[t]; before2 = anorm2; rho2 = anorm2; have_norm = true; }
                if (t == 0) { vb = gs[1]; break; }
                for (int i = tid; i < t; i += CT) {                        // hh = R^{-T} g
                    double acc = 0.0;
                    for (int l = 0; l <= i; ++l) acc = fma(Tm[l + i * ldT], gs[l], acc);
                    hh[i] = acc;
                }
                __syncthreads();
                for (int i = tid; i < t; i += CT) {                        // y = R^{-1} hh
                    double acc = 0.0;
                    for (int l = i; l < t; ++l) acc = fma(Tm[i + l * ldT], hh[l], acc);
                    y[i] = acc;
                    ys[i] = sweep ? ys[i] + acc : acc;
                }
                __syncthreads();
                double p2 = 0.0, pb = 0.0;                                 // ||v||^2 and v'b partials of the swept v
                if (vec) {
                    for (int row = tid * W; row < Mc; row += CT * W) {     // v -= A_S y on the owned rows, W rows per 16 B load
                        double acc[W];
#pragma unroll
                        for (int e = 0; e < W; ++e) acc[e] = v[row + e];
#pragma unroll 8
                        for (int i = 0; i < t; ++i) {
                            double av[W];
                            RowVec<T>::load(colp[i] + row, av);
                            const double yi = y[i];
#pragma unroll
                            for (int e = 0; e < W; ++e) acc[e] = fma(-av[e], yi, acc[e]);
                        }
#pragma unroll
                        for (int e = 0; e < W; ++e) {
                            v[row + e] = acc[e];
                            p2 = fma(acc[e], acc[e], p2);
                            pb = fma(acc[e], (double)b[row + e], pb);
                        }
                    }
                } else {
                    for (int row = tid; row < Mc; row += CT) {
                        double acc = v[row];
                        for (int i = 0; i < t; ++i) acc -= (double)colp[i][row] * y[i];
                        v[row] = acc;
                        p2 += acc * acc;
                        pb = fma(acc, (double)b[row], pb);
                    }
                }
                p2 = block_sum<CT>(p2, red);
                pb = block_sum<CT>(pb, red);
                if (tid == 0) { g[0] = p2; g[1] = pb; }                    // one exchange for rho^2 and v'b
                cluster_allreduce<CL>(cl, ex, g, 2, gs);
                rho2 = gs[0];
                vb = gs[1];
                if (rho2 >= 0.5 * before2) break;                          // DGKS: one sweep was enough
                before2 = rho2;
                s2 = 0.0;
            }
            if (!(rho2 > 1e-26 * anorm2)) { flags |= 1; continue; }        // numerically dependent atom
            if (rho2 < ILLCOND_RATIO * anorm2) flags |= FLAG_ILLCOND;      // see refine_coefficients (update_common.cuh)
            const double rho = sqrt(rho2);
            const double zt = vb / rho;                                    // z_t = q_t' b
            const double gam = zt / rho;
            double s2r = 0.0;
            for (int row = tid; row < Mc; row += CT) {                     // r <- r - q_t z_t on the owned rows
                const T rr = (T)((double)r[row] - gam * v[row]);
                r[row] = rr;
                s2r += (double)rr * (double)rr;
            }
            s2r = block_sum<CT>(s2r, red);
            if (tid == 0) g[0] = s2r;
            cluster_allreduce<CL>(cl, ex, g, 1, gs);
            nr2 = gs[0];
            const double irho = 1.0 / rho;
            for (int i = tid; i < t; i += CT) {                            // new column of R^{-1}
                const double e = -ys[i] * irho;
                if (crank == 0) Tg[i + (size_t)t * kcap] = e;
                if (t_in_smem) Tsm[i + t * ldT] = e;
            }
            if (tid == 0) {
                if (crank == 0) Tg[t + (size_t)t * kcap] = irho;
                if (t_in_smem) Tsm[t + t * ldT] = irho;
                zs[t] = zt; ssel[t] = j; colp[t] = aj;
            }
            ++t;
            changed = true;
            __syncthreads();
            if (!t_in_smem) cl.sync();     // the factor lives in global memory: peers must see rank 0's new column
        }
    }

    // cluster-uniform: every CTA saw the same appends
    const bool ill = changed && ((flags & FLAG_ILLCOND) || (a.flags[sig] & FLAG_ILLCOND));
    if (changed) {
        for (int i = tid; i < t; i += CT) {                                // x_S = R^{-1} Q'b, on every CTA
            double acc = 0.0;
            for (int l = i; l < t; ++l) acc = fma(Tm[i + l * ldT], zs[l], acc);
            ys[i] = acc;
        }
        __syncthreads();
    }
    if (ill) {
        // ill-conditioned support: two steps of iterative refinement, x += R^{-1} R^{-T} A_S' (b - A_S x), as in
        // refine_coefficients (update_common.cuh) with the rows split over the cluster and one all-reduce per step
        for (int rep = 0; rep < 2; ++rep) {
            for (int row = tid; row < Mc; row += CT) {
                double acc = (double)b[row];
                for (int i = 0; i < t; ++i) acc = fma(-(double)colp[i][row], ys[i], acc);
                v[row] = acc;
            }
            __syncthreads();
            for (int i = warp; i < t; i += CT / 32) {
                const T* ai = colp[i];
                double s = 0.0;
                for (int row = lane; row < Mc; row += 32) s = fma((double)ai[row], v[row], s);
                s = warp_sum(s);
                if (lane == 0) g[i] = s;
            }
            cluster_allreduce<CL>(cl, ex, g, t, gs);
            for (int i = tid; i < t; i += CT) {
                double acc = 0.0;
                for (int l = 0; l <= i; ++l) acc = fma(Tm[l + i * ldT], gs[l], acc);
                hh[i] = acc;
            }
            __syncthreads();
            for (int i = tid; i < t; i += CT) {
                double acc = 0.0;
                for (int l = i; l < t; ++l) acc = fma(Tm[i + l * ldT], hh[l], acc);
                ys[i] += acc;
            }
            __syncthreads();
        }
    }
    if (crank == 0) {
        double nr = a.resnorm[sig];
        if (changed) {
            for (int i = tid; i < t; i += CT) {
                a.x[(size_t)sig * kcap + i] = ys[i];
                a.sel[(size_t)sig * kcap + i] = ssel[i];
                a.z[(size_t)sig * kcap + i] = zs[i];
            }
            nr = sqrt(nr2);
        }
        if (tid == 0) {
            a.nnz[sig] = t;
            a.resnorm[sig] = nr;
            a.iters[sig] += 1;
            if (flags) a.flags[sig] |= flags;
            if (!(nr >= a.eps)) a.done[sig] = 1;
        }
    }
}

size_t cluster_smem_bytes(int ld, int kcap, bool t_in_smem, int CL) {
    const int xstride = kcap + 4;
    size_t bytes = (size_t)(ld / CL + 2 * xstride + 4 * kcap + 2 * CL * xstride) * sizeof(double) +
                   (size_t)((kcap + 1) & ~1) * sizeof(int) + (size_t)kcap * sizeof(void*);
    if (t_in_smem) bytes += (size_t)kcap * (kcap | 1) * sizeof(double);
    return bytes;
}

template <typename T, int CL>
cudaError_t launch_cl(const StateArgs& a, cudaStream_t st, const void* Acache, bool probe_only) {
    const int t_in_smem = a.kcap <= T_SMEM_MAX_K_CL ? 1 : 0;
    const size_t smem = cluster_smem_bytes(a.ld, a.kcap, t_in_smem != 0, CL);
    if (smem > MAX_DYN_SMEM) return cudaErrorInvalidConfiguration;
    cudaError_t e = cudaFuncSetAttribute(omp_update_cluster_kernel<T, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (CL > 8) {
        e = cudaFuncSetAttribute(omp_update_cluster_kernel<T, CL>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return e;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(a.nsig * CL));
    cfg.blockDim = dim3(CT);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (probe_only) {                                  // can the device co-schedule one such cluster at all?
        int n = 0;
        e = cudaOccupancyMaxActiveClusters(&n, omp_update_cluster_kernel<T, CL>, &cfg);
        if (e != cudaSuccess) return e;
        return n >= 1 ? cudaSuccess : cudaErrorInvalidConfiguration;
    }
    return cudaLaunchKernelEx(&cfg, omp_update_cluster_kernel<T, CL>, a, static_cast<const T*>(Acache), t_in_smem);
}

// 16-CTA clusters: few signals (the grid stays within one wave), long atoms (each CTA still streams >= 256 rows per
// atom with 16-byte loads), and a device that can place such a cluster (cudaOccupancyMaxActiveClusters).
// CSB200_CLUSTER=8|16 overrides the size rule.
template <typename T>
bool wide_cluster(const StateArgs& a) {
    // placement depends on the per-CTA shared memory only: remember the largest size that fitted, the smallest that did not
    static std::atomic<size_t> ok_upto{0}, bad_from{SIZE_MAX};
    const char* env = getenv("CSB200_CLUSTER");
    if (env && !strcmp(env, "8")) return false;
    const bool forced = env && !strcmp(env, "16");
    constexpr int W = RowVec<T>::W;
    if (a.ld % (16 * W) != 0) return false;
    if (!forced && !(a.nsig <= 4 && a.ld >= 4096)) return false;
    const size_t smem = cluster_smem_bytes(a.ld, a.kcap, a.kcap <= T_SMEM_MAX_K_CL, 16);
    if (smem <= ok_upto.load()) return true;
    if (smem >= bad_from.load()) return false;
    const cudaError_t e = launch_cl<T, 16>(a, nullptr, nullptr, true);
    if (e != cudaSuccess) { cudaGetLastError(); if (smem < bad_from.load()) bad_from.store(smem); return false; }
    if (smem > ok_upto.load()) ok_upto.store(smem);
    return true;
}

template <typename T>
cudaError_t launch_t(const StateArgs& a, cudaStream_t st, const void* Acache) {
    if (wide_cluster<T>(a)) {
        const cudaError_t e = launch_cl<T, 16>(a, st, Acache, false);
        if (e == cudaSuccess) return e;
        cudaGetLastError();                            // placement refused at launch time: the portable size always fits
    }
    return launch_cl<T, 8>(a, st, Acache, false);
}

}  // namespace

size_t omp_update_cluster_smem_bytes(int ld, int kcap) { return cluster_smem_bytes(ld, kcap, kcap <= T_SMEM_MAX_K_CL, 8); }

cudaError_t launch_omp_update_cluster(const StateArgs& a, bool f32, cudaStream_t st, const void* Acache) {
    if (a.nsig <= 0) return cudaSuccess;
    return f32 ? launch_t<float>(a, st, Acache) : launch_t<double>(a, st, Acache);
}

}  // namespace csb
