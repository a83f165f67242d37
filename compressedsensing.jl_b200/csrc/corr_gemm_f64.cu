// Batched residual correlation  C = A' R  (FP64) with the |c| top-s selection fused into the
// epilogue -- subsystems (1)+(2) of the north star for the many-signal path.
//
// Replaces, for B signals at once, `mul!(P.Ar, P.A', P.r); @. P.Ar = abs(P.Ar); argmax(P.Ar)` /
// `partialsortperm(P.Ar, 1:k, rev=true)` (/root/reference/src/matchingpursuit.jl:181-193).
// C (N x B doubles, 4 GiB at the headline config) is never written: each consumer warp reduces
// its 64-atom x 32-signal accumulator block to the top-s (|c|, atom) records of that block.
//
// sm_100a design
//   * tcgen05 has no f64 kind, so the FP64 tensor path is mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4;
//     ptxas lowers every wider f64 mma shape to it).  What is Blackwell/Hopper-class here is the
//     data movement: both operands are K-contiguous (atoms and signals are columns), so one 2-D
//     TMA box {16 doubles x 128 columns} per operand per k-chunk lands a 128 B-row, SWIZZLE_128B
//     tile in shared memory, through a 4-stage full/empty mbarrier ring that runs 3 chunks ahead of
//     the math and straight across tile boundaries (persistent CTAs, one per SM).  All 8 warps
//     compute (256 threads x ~220 registers is the whole register file, and warps are allocated
//     in groups of 4, so there is no room for a dedicated producer warp); lane 0 of warp 0 issues
//     the TMA for chunk c+3 before it consumes chunk c.
//   * Fragment loads are LDS.128: lane (g = lane/4, q = lane%4) reads the 16 B chunk q of an
//     8-double k-group for row g and feeds .x to one DMMA and .y to the next.  Both operands use
//     the same k-permutation, so the contraction is unchanged, and with the 128 B swizzle the
//     warp-wide 512 B request touches every bank exactly 4 times (the bandwidth floor).
//   * CTA tile 128 atoms x 128 signals, warp tile 64 x 32 (64 accumulator doubles per thread):
//     shared-memory traffic is 12 LDS.128 per 64 DMMA per warp, ~19 % of the LDS bandwidth at
//     DMMA peak.  TMA zero-fills out-of-range rows/columns, so no shape padding is needed.
#include "common.cuh"
#include <cstdlib>

namespace csb {

namespace {

constexpr int TILE_N = 128;                 // atoms per CTA tile
constexpr int TILE_B = 128;                 // signals per CTA tile
constexpr int KCH = 16;                     // doubles per k-chunk: 128 B rows
constexpr int A_TILE_BYTES = TILE_N * KCH * 8;
constexpr int R_TILE_BYTES = TILE_B * KCH * 8;
constexpr int BOX_PAIR_BYTES = A_TILE_BYTES + R_TILE_BYTES;   // one 16-deep k-slice of both operands: 32 KiB
// A pipeline stage holds SUB such slices (SUB = 1: 4 stages x 32 KiB, SUB = 2: 3 stages x 64 KiB): fewer, longer
// stages halve the number of mbarrier round trips per flop.
template <int SUB> struct Pipe {
    static constexpr int STAGES = SUB == 1 ? 4 : 3;
    static constexpr int STAGE_BYTES = SUB * BOX_PAIR_BYTES;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 2 * STAGES * 8;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
// Tile order.  Atom tiles are walked in bands of `band` tiles; within a band the signal tile is the slow
// index.  band = tilesN (the default) is plain atom-fastest order: the ~148 concurrently running CTAs cover
// every atom tile of ~2.3 signal tiles.  Measured on B200 at the headline shape (profiles/gemm_band_sweep_r01.md):
// narrower bands cut DRAM reads from 6.0 GB to 1.55 GB per launch (the band stays L2-resident) but cost
// 0.8-2.8 % of kernel time, and DRAM is at 2.3 % of its bandwidth either way on this tensor-bound kernel,
// so the default keeps the faster order; CSB200_GEMM_BAND overrides it.
constexpr int DEFAULT_BAND = 0;   // 0 = tilesN
constexpr int DEFAULT_VARIANT = -1;   // -1 = choose by shape
__device__ __forceinline__ void tile_coords(int tile, int tilesN, int tilesB, int BAND, int& tn, int& tb) {
    const int per_band = BAND * tilesB;
    const int band = tile / per_band;
    const int rem = tile - band * per_band;
    const int left = tilesN - band * BAND;
    const int bw = left < BAND ? left : BAND;
    tb = rem / bw;
    tn = band * BAND + (rem - tb * bw);
}

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// Warp layout variants (CTA tile is always 128 atoms x 128 signals):
//   <8,4,2,4>  8 warps, warp tile 64 x 32, 64 accumulators/thread, ~228 registers (fills the register file)
//   <4,4,4,4> 16 warps, warp tile 32 x 32, 32 accumulators/thread, <=128 registers: twice the warps per
//             scheduler to cover LDS / mbarrier latency at the price of 8 instead of 6 LDS.128 per 32 DMMA
// Epilogues: EPI_TOPS = |c| top-S per (64-atom block, signal); EPI_STORE = plain C store (A'A);
// EPI_OLS = forward-regression criterion c^2 / rescaling with the rescaling down-date fused in (see below).
// EPI_ABS = dense |c| store, signal-major (pval[sig * ldc + atom]): the large-S paths select from it with a radix select.
enum { EPI_TOPS = 0, EPI_STORE = 1, EPI_OLS = 2, EPI_ABS = 3 };
// DUAL (EPI_OLS only): the 128 "signal" columns of a CTA tile are 64 residuals r_s (from mapR) followed by the newest
// orthonormal directions q_s of the SAME 64 signals (from mapQ), arranged so that a thread's accumulators
// acc[i][j] (j < NJ/2) = <a, r_s> and acc[i][j + NJ/2] = <a, q_s> belong to the same (atom, signal) pairs.
// Register cap: 224 per thread for the 8-warp layouts instead of the 255 a (256, 1) launch bound would allow.  The
// kernel was at 228-236 and is spill-free at 224; the 8192 registers this leaves on the SM are exactly one 128-thread x
// 64-register CTA of omp_update_kernel, which is what lets the per-signal update of one half of a batch run UNDER the
// correlation pass of the other half (api.cu, run_omp_split) instead of after it.
template <int MI, int NJ, int WM, int WN, int SUB, int EPI = EPI_TOPS, bool DUAL = false>
__global__ void __maxnreg__(WM * WN * 32 <= 256 ? 224 : 128)
corr_gemm_f64_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapR,
                     int N, int nsig, int kchunks, int tilesN, int tilesB, int band, int S, int P, int idx_offset,
                     double* __restrict__ pval, int* __restrict__ pidx, long long ldc,
                     const __grid_constant__ CUtensorMap mapQ, double* __restrict__ resc) {
    constexpr bool STORE = EPI == EPI_STORE;
    constexpr int SIG_PER_TILE = DUAL ? TILE_B / 2 : TILE_B;        // distinct signals per CTA tile
    constexpr int STAGES = Pipe<SUB>::STAGES;
    constexpr int STAGE_BYTES = Pipe<SUB>::STAGE_BYTES;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
    uint8_t* sm = smem_raw + pad;                       // 1024 B aligned: required by SWIZZLE_128B
    const uint32_t sm_base = smem_u32(sm);
    const uint32_t bar_full = sm_base + STAGES * STAGE_BYTES;
    const uint32_t bar_empty = bar_full + STAGES * 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntiles = tilesN * tilesB;
    // k-chunks this CTA streams, over all of its tiles: one flat sequence through the stage ring
    const int my_tiles = blockIdx.x < ntiles ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int total_chunks = my_tiles * kchunks;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bar_full + s * 8, 1);
            mbar_init(bar_empty + s * 8, WM * WN);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapR) : "memory");
    }
    __syncthreads();

    // TMA producer step, executed by thread 0 only: load flat chunk c into stage c % STAGES once the
    // consumers have released that stage's previous occupant (chunk c - STAGES).
    auto issue_chunk = [&](int c) {
        const int stg = c % STAGES;
        const uint32_t par = (uint32_t)((c / STAGES) & 1) ^ 1u;
        const int tseq = c / kchunks, kc = c - tseq * kchunks;
        const int tile = (int)blockIdx.x + tseq * (int)gridDim.x;
        int tn, tb;
        tile_coords(tile, tilesN, tilesB, band, tn, tb);
        mbar_wait(bar_empty + stg * 8, par);
        mbar_arrive_expect_tx(bar_full + stg * 8, STAGE_BYTES);
        const uint32_t dst = sm_base + stg * STAGE_BYTES;
#pragma unroll
        for (int u = 0; u < SUB; ++u) {          // k-slices past the end of the matrix are zero-filled by TMA
            tma_load_2d(dst + u * BOX_PAIR_BYTES, &mapA, bar_full + stg * 8, (kc * SUB + u) * KCH, tn * TILE_N);
            if constexpr (DUAL) {       // two {16 x 64} boxes: residuals into rows 0-63, directions into rows 64-127
                tma_load_2d(dst + u * BOX_PAIR_BYTES + A_TILE_BYTES, &mapR, bar_full + stg * 8, (kc * SUB + u) * KCH, tb * SIG_PER_TILE);
                tma_load_2d(dst + u * BOX_PAIR_BYTES + A_TILE_BYTES + R_TILE_BYTES / 2, &mapQ, bar_full + stg * 8, (kc * SUB + u) * KCH, tb * SIG_PER_TILE);
            } else {
                tma_load_2d(dst + u * BOX_PAIR_BYTES + A_TILE_BYTES, &mapR, bar_full + stg * 8, (kc * SUB + u) * KCH, tb * TILE_B);
            }
        }
    };
    if (threadIdx.x == 0) {
        for (int c = 0; c < STAGES - 1 && c < total_chunks; ++c) issue_chunk(c);
    }

    // ---------------------------------- DMMA consumers -----------------------------------
    static_assert(8 * MI * WM == TILE_N && 8 * NJ * WN == TILE_B, "warp layout must cover the CTA tile");
    const int wm = warp % WM;         // which (8*MI)-atom slice of the tile
    const int wn = warp / WM;         // which (8*NJ)-signal slice
    const int g = lane >> 2, q = lane & 3;
    const uint32_t a_row = (uint32_t)(wm * 8 * MI + g) * 128u;                   // + i*1024
    // signal-side fragment rows: j-th 8-column group of this warp (DUAL: groups j >= NJ/2 are the q half of the tile)
    auto r_off = [&](int j) -> uint32_t {
        const int row = DUAL ? (j / (NJ / 2)) * (TILE_B / 2) + wn * 8 * (NJ / 2) + (j % (NJ / 2)) * 8 + g
                             : wn * 8 * NJ + j * 8 + g;
        return (uint32_t)A_TILE_BYTES + (uint32_t)row * 128u;
    };
    int stage = 0;
    uint32_t phase = 0;
    int chunk = 0;                     // flat chunk index being consumed

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        int tn, tb;
        tile_coords(tile, tilesN, tilesB, band, tn, tb);
        double acc[MI][NJ][2];
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

        for (int kc = 0; kc < kchunks; ++kc, ++chunk) {
            // keep STAGES-1 chunks in flight, across tile boundaries (the next tile's first chunks
            // stream in during this tile's epilogue)
            if (threadIdx.x == 0 && chunk + STAGES - 1 < total_chunks) issue_chunk(chunk + STAGES - 1);
            __syncwarp();
            mbar_wait(bar_full + stage * 8, phase);
            const uint8_t* st0 = sm + stage * STAGE_BYTES;
#pragma unroll
            for (int hh = 0; hh < 2 * SUB; ++hh) {
                const int h = hh & 1;
                const uint8_t* st = st0 + (hh >> 1) * BOX_PAIR_BYTES;
                const uint32_t sw = (uint32_t)(((4 * h + q) ^ g) << 4);      // SWIZZLE_128B: chunk ^= row % 8
                double2 af[MI], bf[NJ];
#pragma unroll
                for (int i = 0; i < MI; ++i) af[i] = *reinterpret_cast<const double2*>(st + a_row + i * 1024 + sw);
#pragma unroll
                for (int j = 0; j < NJ; ++j) bf[j] = *reinterpret_cast<const double2*>(st + r_off(j) + sw);
#pragma unroll
                for (int i = 0; i < MI; ++i)
#pragma unroll
                    for (int j = 0; j < NJ; ++j) dmma884(acc[i][j], af[i].x, bf[j].x);
#pragma unroll
                for (int i = 0; i < MI; ++i)
#pragma unroll
                    for (int j = 0; j < NJ; ++j) dmma884(acc[i][j], af[i].y, bf[j].y);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_empty + stage * 8);
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }

        if constexpr (STORE) {
            // ---- plain store epilogue: C[sig + atom * ldc]; a thread's (e = 0, 1) pair is one 16 B store ----
#pragma unroll
            for (int i = 0; i < MI; ++i) {
                const int atom = tn * TILE_N + wm * 8 * MI + i * 8 + g;
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    const int sig = tb * TILE_B + wn * 8 * NJ + j * 8 + 2 * q;
                    if (atom < N) {
                        double* dst = pval + (size_t)atom * ldc + sig;
                        if (sig + 1 < nsig && (ldc & 1) == 0) *reinterpret_cast<double2*>(dst) = make_double2(acc[i][j][0], acc[i][j][1]);
                        else { if (sig < nsig) dst[0] = acc[i][j][0]; if (sig + 1 < nsig) dst[1] = acc[i][j][1]; }
                    }
                }
            }
            continue;
        }
        if constexpr (EPI == EPI_ABS) {
            // lanes g = 0..7 hold 8 consecutive atoms of one signal: 64-byte segments per (i, j, e)
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int sig = tb * TILE_B + wn * 8 * NJ + j * 8 + 2 * q + e;
                    if (sig < nsig) {
                        double* dst = pval + (size_t)sig * ldc + tn * TILE_N + wm * 8 * MI + g;
#pragma unroll
                        for (int i = 0; i < MI; ++i)
                            if (tn * TILE_N + wm * 8 * MI + g + i * 8 < N) __stcs(dst + i * 8, fabs(acc[i][j][e]));
                    }
                }
            }
            continue;
        }
        if constexpr (EPI == EPI_OLS) {
            // ---- forward-regression criterion (src/forward.jl:69-76, 97-114): delta2_j = <a_j, r>^2 / resc_j with
            // resc_j = ||a_j||^2 - ||Q1'a_j||^2 kept per (signal, atom) in HBM and down-dated here by <a_j, q_new>^2.
            // Active atoms hold resc = +Inf, so their delta2 is exactly 0, which is what `P.δ²[x.nzind] = 0` sets.
            // One (delta2, atom) record per (64-atom block, signal); NaN / negative quotients never win
            // (the package's findmax override skips NaN, src/util.jl:173-189; a non-negative entry always exists).
            const int atom0 = tn * TILE_N + wm * 8 * MI + g;
            const int p = tn * WM + wm;
            constexpr int NS = DUAL ? NJ / 2 : NJ;
            // resc is [atom][signal] (leading dimension ldc, even): a thread's two signals (e = 0, 1) are one 16 B
            // streaming load / store, and the 8 atoms of a j-group are loaded together (one DRAM round trip per group)
#pragma unroll
            for (int j = 0; j < NS; ++j) {
                const int sig0 = tb * SIG_PER_TILE + wn * 8 * NS + j * 8 + 2 * q;
                double2 rv[MI];
                if (sig0 < nsig) {
#pragma unroll
                    for (int i = 0; i < MI; ++i) {
                        const int idx = atom0 + i * 8;
                        rv[i] = idx < N ? __ldcs(reinterpret_cast<const double2*>(resc + (size_t)idx * ldc + sig0))
                                        : make_double2(1.0, 1.0);
                    }
                    if constexpr (DUAL) {
#pragma unroll
                        for (int i = 0; i < MI; ++i) {
                            const int idx = atom0 + i * 8;
                            const double d0 = acc[i][j + NJ / 2][0], d1 = acc[i][j + NJ / 2][1];
                            rv[i].x = fma(-d0, d0, rv[i].x);
                            rv[i].y = fma(-d1, d1, rv[i].y);
                            if (idx < N) __stcs(reinterpret_cast<double2*>(resc + (size_t)idx * ldc + sig0), rv[i]);
                        }
                    }
                }
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int sig = sig0 + e;
                    double bv = -1.0;
                    int bi = INT_MAX;
                    if (sig < nsig) {
#pragma unroll
                        for (int i = 0; i < MI; ++i) {
                            const int idx = atom0 + i * 8;
                            const double c = acc[i][j][e];
                            const double v = c * c / (e ? rv[i].y : rv[i].x);
                            if (idx < N && v >= 0.0 && v > bv) { bv = v; bi = idx; }   // idx ascends with i: first max wins
                        }
                    }
#pragma unroll
                    for (int off = 4; off < 32; off <<= 1) {
                        const double ov = __shfl_xor_sync(0xffffffffu, bv, off);
                        const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                        if (cand_better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
                    }
                    if (g == 0 && sig < nsig && p < P) {
                        const size_t o = (size_t)sig * P + p;
                        pval[o] = bv;
                        pidx[o] = (bi == INT_MAX) ? -1 : bi + idx_offset;
                    }
                }
            }
            continue;
        }
        // ---- fused epilogue: top-S of |c| over this warp's 8*MI atoms, per signal column ----
        // acc[i][j][e] = c[atom = wm*8*MI + i*8 + g][signal = wn*8*NJ + j*8 + 2q + e]
        const int atom0 = tn * TILE_N + wm * 8 * MI + g;
        const int p = tn * WM + wm;
        if (NJ == 4 && S == 1 && tn * TILE_N + TILE_N <= N) {
            // omp / mp on an interior tile (the common case): one candidate per (signal, block), no exclusion of earlier
            // picks, no bounds test.  The 8 lanes g = 0..7 of a quad column q hold candidates for the same 8 signals
            // (j, e); instead of eight 3-level shuffle trees, a transposing butterfly halves the signals a lane keeps at
            // every level (g bit 2 picks j >= 2, bit 1 picks j odd, bit 0 picks e), so 7 exchanges replace 24 and every
            // lane ends with the block winner of ONE signal and stores it itself.  FP64 compares bound this epilogue.
            double v8[8];
            int i8[8];
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    double bv = fabs(acc[0][j][e]);
                    int bi = atom0;
                    if (!(bv >= 0.0)) { bv = -1.0; bi = INT_MAX; }           // NaN never wins
#pragma unroll
                    for (int i = 1; i < MI; ++i) {
                        const double v = fabs(acc[i][j][e]);
                        if (v > bv) { bv = v; bi = atom0 + i * 8; }          // idx ascends with i: first max wins
                    }
                    v8[j * 2 + e] = bv; i8[j * 2 + e] = bi;
                }
            }
            const bool h2 = (g & 4) != 0, h1 = (g & 2) != 0, h0 = (g & 1) != 0;
            double v4[4]; int i4[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {                                    // keep j in {0,1} or {2,3}
                const double mine = h2 ? v8[c + 4] : v8[c], send = h2 ? v8[c] : v8[c + 4];
                const int mi = h2 ? i8[c + 4] : i8[c], si = h2 ? i8[c] : i8[c + 4];
                const double ov = __shfl_xor_sync(0xffffffffu, send, 16);
                const int oi = __shfl_xor_sync(0xffffffffu, si, 16);
                const bool take = cand_better(ov, oi, mine, mi);
                v4[c] = take ? ov : mine; i4[c] = take ? oi : mi;
            }
            double v2[2]; int i2[2];
#pragma unroll
            for (int c = 0; c < 2; ++c) {                                    // keep the even or the odd j of the pair
                const double mine = h1 ? v4[c + 2] : v4[c], send = h1 ? v4[c] : v4[c + 2];
                const int mi = h1 ? i4[c + 2] : i4[c], si = h1 ? i4[c] : i4[c + 2];
                const double ov = __shfl_xor_sync(0xffffffffu, send, 8);
                const int oi = __shfl_xor_sync(0xffffffffu, si, 8);
                const bool take = cand_better(ov, oi, mine, mi);
                v2[c] = take ? ov : mine; i2[c] = take ? oi : mi;
            }
            double bv; int bi;
            {                                                                // keep e = 0 or 1
                const double mine = h0 ? v2[1] : v2[0], send = h0 ? v2[0] : v2[1];
                const int mi = h0 ? i2[1] : i2[0], si = h0 ? i2[0] : i2[1];
                const double ov = __shfl_xor_sync(0xffffffffu, send, 4);
                const int oi = __shfl_xor_sync(0xffffffffu, si, 4);
                const bool take = cand_better(ov, oi, mine, mi);
                bv = take ? ov : mine; bi = take ? oi : mi;
            }
            const int jk = (h2 ? 2 : 0) + (h1 ? 1 : 0), ek = h0 ? 1 : 0;
            const int sig = tb * TILE_B + wn * 8 * NJ + jk * 8 + 2 * q + ek;
            if (sig < nsig && p < P) {
                const size_t o = (size_t)sig * P + p;
                pval[o] = bv;
                pidx[o] = (bi == INT_MAX) ? -1 : bi + idx_offset;
            }
            continue;
        }
        double pv[NJ][2];
        int pi[NJ][2];
        for (int s = 0; s < S; ++s) {
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    double bv = -1.0;
                    int bi = INT_MAX;
#pragma unroll
                    for (int i = 0; i < MI; ++i) {
                        const double v = fabs(acc[i][j][e]);
                        const int idx = atom0 + i * 8;
                        bool ok = idx < N;
                        if (s > 0) ok = ok && (v < pv[j][e] || (v == pv[j][e] && idx > pi[j][e]));
                        if (ok && v > bv) { bv = v; bi = idx; }      // idx ascends with i: first max wins
                    }
#pragma unroll
                    for (int off = 4; off < 32; off <<= 1) {
                        const double ov = __shfl_xor_sync(0xffffffffu, bv, off);
                        const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                        if (cand_better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
                    }
                    pv[j][e] = bv;
                    pi[j][e] = bi;
                    const int sig = tb * TILE_B + wn * 8 * NJ + j * 8 + 2 * q + e;
                    if (g == 0 && sig < nsig && p < P) {
                        const size_t o = ((size_t)sig * P + p) * S + s;
                        pval[o] = bv;
                        pidx[o] = (bi == INT_MAX) ? -1 : bi + idx_offset;
                    }
                }
            }
        }
    }
}

}  // namespace

namespace {
int gemm_variant() {
    static const int v = [] { const char* e = getenv("CSB200_GEMM_VARIANT"); return e ? atoi(e) : DEFAULT_VARIANT; }();
    return v;
}
}  // namespace

// atoms per candidate block emitted by the active variant (P = ceil(N / this))
int corr_gemm_f64_block() { return gemm_variant() == 1 ? 32 : 64; }

cudaError_t corr_gemm_f64_setup() {
    cudaError_t e = cudaFuncSetAttribute(corr_gemm_f64_kernel<8, 4, 2, 4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Pipe<1>::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(corr_gemm_f64_kernel<8, 4, 2, 4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Pipe<2>::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(corr_gemm_f64_kernel<8, 4, 2, 4, 1, EPI_STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, Pipe<1>::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(corr_gemm_f64_kernel<8, 4, 2, 4, 2, EPI_ABS>, cudaFuncAttributeMaxDynamicSharedMemorySize, Pipe<2>::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(corr_gemm_f64_kernel<8, 4, 2, 4, 2, EPI_OLS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Pipe<2>::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(corr_gemm_f64_kernel<8, 4, 2, 4, 2, EPI_OLS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Pipe<2>::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(corr_gemm_f64_kernel<4, 4, 4, 4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Pipe<1>::SMEM_BYTES);
}

cudaError_t launch_corr_gemm_f64(const CUtensorMap* mapA, const CUtensorMap* mapR, const CorrArgs& a,
                                 int num_sms, cudaStream_t st) {
    const int tilesN = (a.N + TILE_N - 1) / TILE_N;
    const int tilesB = (a.nsig + TILE_B - 1) / TILE_B;
    const long long ntiles = (long long)tilesN * tilesB;
    if (ntiles <= 0) return cudaSuccess;
    const int grid = (int)(ntiles < num_sms ? ntiles : num_sms);
    static const int band_env = [] { const char* e = getenv("CSB200_GEMM_BAND"); return e ? atoi(e) : 0; }();
    int band = band_env > 0 ? band_env : DEFAULT_BAND;
    if (band <= 0 || band > tilesN) band = tilesN;
    const int kslices = a.ld / KCH;
    if (a.dense_ld > 0) {
        corr_gemm_f64_kernel<8, 4, 2, 4, 2, EPI_ABS><<<grid, 256, Pipe<2>::SMEM_BYTES, st>>>(
            *mapA, *mapR, a.N, a.nsig, (kslices + 1) / 2, tilesN, tilesB, band, a.S, a.P, a.idx_offset, a.pval, a.pidx,
            a.dense_ld, *mapR, nullptr);
        return cudaGetLastError();
    }
    int variant = gemm_variant();
    // default: 64 KiB stages (2 k-slices per mbarrier round trip; measured 2.7 % faster at K = 1024) unless an odd
    // slice count would make the zero-filled tail slice a noticeable share of the work
    if (variant < 0) variant = (kslices % 2 == 0 || kslices >= 16) ? 2 : 0;
    if (variant == 1)
        corr_gemm_f64_kernel<4, 4, 4, 4, 1><<<grid, 512, Pipe<1>::SMEM_BYTES, st>>>(
            *mapA, *mapR, a.N, a.nsig, kslices, tilesN, tilesB, band, a.S, a.P, a.idx_offset, a.pval, a.pidx, 0, *mapR, nullptr);
    else if (variant == 2) {
        // measurement hook: CSB200_GEMM_NO_EPILOGUE=1 runs the main loop only (results are garbage): the cost of the
        // fused top-S epilogue is the difference -- 2.45 % of the pass at the C2 shape (profiles/gemm_epilogue_r02.md).
        // Starting warps 4-7 late so that the two warps of a scheduler reach the epilogue at different times did not
        // recover any of it (15.88 ms for every skew from 0 to 24k cycles): one warp per scheduler cannot issue DMMA
        // faster than it already does next to its partner.
        static const bool no_epi = [] { const char* e = getenv("CSB200_GEMM_NO_EPILOGUE"); return e && e[0] == '1'; }();
        corr_gemm_f64_kernel<8, 4, 2, 4, 2><<<grid, 256, Pipe<2>::SMEM_BYTES, st>>>(
            *mapA, *mapR, a.N, a.nsig, (kslices + 1) / 2, tilesN, tilesB, band, no_epi ? 0 : a.S, a.P, a.idx_offset, a.pval, a.pidx, 0, *mapR, nullptr);
    }
    else
        corr_gemm_f64_kernel<8, 4, 2, 4, 1><<<grid, 256, Pipe<1>::SMEM_BYTES, st>>>(
            *mapA, *mapR, a.N, a.nsig, kslices, tilesN, tilesB, band, a.S, a.P, a.idx_offset, a.pval, a.pidx, 0, *mapR, nullptr);
    return cudaGetLastError();
}

cudaError_t launch_gemm_f64_store(const CUtensorMap* mapA, const CUtensorMap* mapR, int N, int nsig, int ld,
                                  double* C, long long ldc, int num_sms, cudaStream_t st) {
    const int tilesN = (N + TILE_N - 1) / TILE_N;
    const int tilesB = (nsig + TILE_B - 1) / TILE_B;
    const long long ntiles = (long long)tilesN * tilesB;
    if (ntiles <= 0) return cudaSuccess;
    const int grid = (int)(ntiles < num_sms ? ntiles : num_sms);
    corr_gemm_f64_kernel<8, 4, 2, 4, 1, EPI_STORE><<<grid, 256, Pipe<1>::SMEM_BYTES, st>>>(
        *mapA, *mapR, N, nsig, ld / KCH, tilesN, tilesB, tilesN, 1, 1, 0, C, nullptr, ldc, *mapR, nullptr);
    return cudaGetLastError();
}

// Forward-regression pass: one (delta2, atom) candidate per (64-atom block, signal).  mapQ == nullptr: first
// step (no direction yet; mapR has 128-column boxes).  Otherwise mapR / mapQ have 64-column boxes and resc is
// down-dated by <a_j, q_s>^2 in the same launch.  resc: [N][ldr] (atom-major, ldr even >= nsig).
cudaError_t launch_corr_gemm_f64_ols(const CUtensorMap* mapA, const CUtensorMap* mapR, const CUtensorMap* mapQ,
                                     const CorrArgs& a, double* resc, long long ldr, int num_sms, cudaStream_t st) {
    const int sig_per_tile = mapQ ? TILE_B / 2 : TILE_B;
    const int tilesN = (a.N + TILE_N - 1) / TILE_N;
    const int tilesB = (a.nsig + sig_per_tile - 1) / sig_per_tile;
    const long long ntiles = (long long)tilesN * tilesB;
    if (ntiles <= 0) return cudaSuccess;
    const int grid = (int)(ntiles < num_sms ? ntiles : num_sms);
    const int kchunks = (a.ld / KCH + 1) / 2;
    if (mapQ)
        corr_gemm_f64_kernel<8, 4, 2, 4, 2, EPI_OLS, true><<<grid, 256, Pipe<2>::SMEM_BYTES, st>>>(
            *mapA, *mapR, a.N, a.nsig, kchunks, tilesN, tilesB, tilesN, 1, a.P, a.idx_offset, a.pval, a.pidx, ldr, *mapQ, resc);
    else
        corr_gemm_f64_kernel<8, 4, 2, 4, 2, EPI_OLS, false><<<grid, 256, Pipe<2>::SMEM_BYTES, st>>>(
            *mapA, *mapR, a.N, a.nsig, kchunks, tilesN, tilesB, tilesN, 1, a.P, a.idx_offset, a.pval, a.pidx, ldr, *mapR, resc);
    return cudaGetLastError();
}

}  // namespace csb
