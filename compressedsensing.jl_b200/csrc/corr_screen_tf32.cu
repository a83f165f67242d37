// Screening pass of the batched omp / mp solves:  C~ = R' A  on the 5th-generation tensor cores (tcgen05, kind::tf32,
// accumulators in tensor memory), reduced in the epilogue to the SCREEN_T largest |c~| per (signal, atom chunk).
//
// Why.  `argmaxinner!` (/root/reference/src/matchingpursuit.jl:181-185) needs the POSITION of max |A'r|, not the
// N correlations.  The FP64 DMMA pass (corr_gemm_f64.cu) computes all N of them to 53 bits at 35 TFLOP/s; this pass
// computes them to ~11 bits at tcgen05 rates and hands the update kernel a short list that provably contains the
// FP64 arg-max: with operands rounded to TF32 (cvt.rna, relative error <= 2^-11 each) and FP32 accumulation,
//       |c~_j - <a_j, r>|  <=  E = screen_kappa(M) * max_j ||a_j|| * ||r||          (Cauchy-Schwarz over the rounding errors)
// so every atom whose |c~| lies within 2E of the largest |c~| is a possible arg-max and nothing else is.  The update
// kernel (update.cu, screen_select) re-evaluates exactly those atoms in FP64 and picks the winner with the reference's
// tie-break; a list that may be incomplete (its last slot still inside the window, a residual outside the FP32 range)
// falls back to an exact FP64 scan of all atoms for that signal.  The selected support is therefore the FP64 one.
//
// sm_100a design (one CTA per SM, 320 threads, persistent over work units = 128 signals x one atom chunk):
//   warp 0   TMA producer: per K-block of 128 bytes (32 TF32 floats or 64 scaled FP16 halves) one box {K-block x 128 signals} of the
//            residual copy and one box {K-block x 256 atoms} of the dictionary copy (SWIZZLE_128B, K-major for both operands:
//            signals and atoms are columns of column-major matrices), 4-stage full/empty mbarrier ring of 48 KiB stages
//            running ahead across tile and unit boundaries;
//   warp 1   one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (or kind::tf32) (M = 128 signals, N = 256 atoms,
//            32 bytes of K) four times per stage and commits the stage back to the producer; the 128 x 256 FP32 accumulator
//            lives in TMEM, double-buffered (2 x 256 of the 512 columns) so the epilogue of tile i runs under the MMAs of tile i + 1;
//   warps 2-9  epilogue, two warps per TMEM lane quadrant (128 columns of the tile each): a TMEM lane IS a signal, so after
//            tcgen05.ld.32x32b every thread scans the atoms of ITS signal in its own registers -- no shuffles -- and keeps
//            the SCREEN_T largest (value, atom) across all tiles of the unit as an unsorted set + its minimum; at the end of
//            the unit the two warps of a quadrant merge through shared memory, sort, and one 64-byte record per
//            (signal, chunk) leaves the SM.  The N x B matrix is never written.
#include "common.cuh"
#include <cstdlib>

namespace csb {

namespace {

constexpr int ST_SIG = 128;                    // signals per tile  (UMMA M, TMEM lanes)
constexpr int ST_ATOM = 256;                   // atoms per tile    (UMMA N, TMEM columns per accumulator stage)
constexpr int ST_KB = 32;                      // floats per K-block: 128-byte swizzle rows
// Pipeline depth: 4 stages (198 KiB) when the pass has the SM to itself; 3 stages (150 KiB) leave ~77 KiB of shared memory
// for CTAs of omp_update_kernel, which run under the pass in the overlapped schedule (api.cu, run_omp_screen).
constexpr int ST_A_BYTES = ST_SIG * ST_KB * 4;     // 16 KiB of R32
constexpr int ST_B_BYTES = ST_ATOM * ST_KB * 4;    // 32 KiB of A32
constexpr int ST_STAGE_BYTES = ST_A_BYTES + ST_B_BYTES;
constexpr int ST_BAR_BYTES = 128;
// Several epilogue warps per TMEM lane quadrant (each scans its share of a tile's 256 columns); at the end of a work unit the
// others hand their top-SCREEN_T to the first through this area: [quadrant][warp - 1][value | index][slot][lane]
constexpr int ST_EPI_WARPS = 8;                                    // 16 measured the same (0.946 vs 0.955 ms per pass)
constexpr int ST_EPI_PER_QUAD = ST_EPI_WARPS / 4;                  // warps sharing a quadrant split the tile's columns evenly
constexpr int ST_MERGE_BYTES = 4 * (ST_EPI_PER_QUAD - 1) * 2 * SCREEN_T * 32 * 4;
constexpr int st_smem_bytes(int stages) { return stages * ST_STAGE_BYTES + ST_BAR_BYTES + ST_MERGE_BYTES + 1024 /* alignment slack */; }
constexpr int ST_THREADS = 64 + ST_EPI_WARPS * 32;
constexpr int ST_TMEM_COLS = 512;

__device__ __forceinline__ uint32_t s_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void bar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// Shared-memory matrix descriptor of a K-major operand tile stored as 128-byte rows under SWIZZLE_128B (what the TMA
// boxes above produce): start address (>> 4), leading byte offset (unused for swizzled K-major, 1), stride byte offset =
// 1024 B between 8-row groups (>> 4), descriptor version 1 (sm_100), layout type 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// Instruction descriptor: D = F32 (bit 4), A = B = TF32 (2 at bits 7 and 10), both K-major, N >> 3 at bit 17, M >> 4 at bit 24.
constexpr uint32_t UMMA_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(ST_ATOM >> 3) << 17) | ((uint32_t)(ST_SIG >> 4) << 24);

// kind::f16 with FP16 operands (format code 0) and the same FP32 accumulator: K = 16 halves = 32 bytes per instruction
constexpr uint32_t UMMA_IDESC_F16 = (1u << 4) | ((uint32_t)(ST_ATOM >> 3) << 17) | ((uint32_t)(ST_SIG >> 4) << 24);
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(UMMA_IDESC_F16), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(UMMA_IDESC), "r"(accumulate) : "memory");
}
// Arrives on the mbarrier once every tcgen05.mma issued so far by this thread has completed (implies fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

#define TM_REGS(r, o) "=r"(r[o + 0]), "=r"(r[o + 1]), "=r"(r[o + 2]), "=r"(r[o + 3]), "=r"(r[o + 4]), "=r"(r[o + 5]), "=r"(r[o + 6]), "=r"(r[o + 7])
// 32 lanes x 32 consecutive 32-bit columns: thread l of the warp receives columns [col, col + 32) of TMEM lane (base lane + l).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : TM_REGS(r, 0), TM_REGS(r, 8), TM_REGS(r, 16), TM_REGS(r, 24)
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// The epilogue's running top-SCREEN_T as an UNSORTED set plus its minimum.  A warp executes the insertion whenever ANY of its 32
// signals needs it (~60 times per 256-atom tile on average, every element in the first tile of a unit); a sorted insertion
// is a dependent chain of 7 compare-exchange steps (~225 cycles), which made the epilogue, not the MMAs, the critical
// path once the FP16 pass halved the MMA time.  Replacing the minimum is 8 independent selects and a 3-level min tree; the
// list is sorted once per work unit (top_sort).  Which of several EQUAL minima is evicted is immaterial to the solve: an
// equal value left behind sits in the last sorted slot, and a last slot inside the window means an exact scan (screen_select).
__device__ __forceinline__ void top_replace(float (&v)[SCREEN_T], int (&id)[SCREEN_T], float& vmin, float x, int idx) {
    bool done = false;
#pragma unroll
    for (int p = 0; p < SCREEN_T; ++p) {
        const bool hit = !done && v[p] == vmin;
        v[p] = hit ? x : v[p];
        id[p] = hit ? idx : id[p];
        done = done || hit;
    }
    static_assert(SCREEN_T == 8, "min tree written for 8 slots");
    vmin = fminf(fminf(fminf(v[0], v[1]), fminf(v[2], v[3])), fminf(fminf(v[4], v[5]), fminf(v[6], v[7])));
}
// descending by value, lower atom index first among equal values (the order the sorted insertion produced)
__device__ __forceinline__ void top_sort(float (&v)[SCREEN_T], int (&id)[SCREEN_T]) {
    auto cx = [&](int a, int b) {                                  // after: slot a holds the entry that comes first
        const bool swap = v[b] > v[a] || (v[b] == v[a] && (unsigned)id[b] < (unsigned)id[a]);
        const float tv = swap ? v[a] : v[b]; const int ti = swap ? id[a] : id[b];
        v[a] = swap ? v[b] : v[a]; id[a] = swap ? id[b] : id[a];
        v[b] = tv; id[b] = ti;
    };
    // 19-comparator sorting network for 8 inputs
    cx(0, 1); cx(2, 3); cx(4, 5); cx(6, 7);
    cx(0, 2); cx(1, 3); cx(4, 6); cx(5, 7);
    cx(1, 2); cx(5, 6); cx(0, 4); cx(3, 7);
    cx(1, 5); cx(2, 6);
    cx(1, 4); cx(3, 6);
    cx(2, 4); cx(3, 5);
    cx(3, 4);
}

// F16: the operands are FP16 (residuals scaled per signal, dictionary scaled as a whole, both by powers of two: api.cu) -- a
// K-block is then 64 halves (the same 128-byte swizzle rows and the same stage bytes), one stage feeds four K = 16
// instructions, and everything else is unchanged.
template <int ST_STAGES, bool F16 = false>
__global__ void __launch_bounds__(ST_THREADS, 1)
corr_screen_tf32_kernel(const __grid_constant__ CUtensorMap mapR, const __grid_constant__ CUtensorMap mapA,
                        int N, int nsig, int kblocks, int tilesN, int chunks, int tiles_per_chunk, int units,
                        int idx_offset, float* __restrict__ cval, int* __restrict__ cidx) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t pad = (1024u - (s_u32(smem_raw) & 1023u)) & 1023u;
    uint8_t* sm = smem_raw + pad;                                  // SWIZZLE_128B tiles need 1024-byte alignment
    const uint32_t sm_base = s_u32(sm);
    const uint32_t bar_full = sm_base + ST_STAGES * ST_STAGE_BYTES;
    const uint32_t bar_empty = bar_full + ST_STAGES * 8;
    const uint32_t bar_tfull = bar_empty + ST_STAGES * 8;          // accumulator stage is complete (MMA -> epilogue)
    const uint32_t bar_tempty = bar_tfull + 2 * 8;                 // accumulator stage has been read (epilogue -> MMA)
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sm + ST_STAGES * ST_STAGE_BYTES + 96);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < ST_STAGES; ++s) { bar_init(bar_full + s * 8, 1); bar_init(bar_empty + s * 8, 1); }
        for (int s = 0; s < 2; ++s) { bar_init(bar_tfull + s * 8, 1); bar_init(bar_tempty + s * 8, ST_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapR) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
    }
    if (warp == 1) {                                               // the allocating warp also deallocates
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(s_u32(const_cast<uint32_t*>(tmem_slot))), "r"(ST_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {                                           // ---- TMA producer
            int stage = 0; uint32_t phase = 0;
            for (int u = blockIdx.x; u < units; u += gridDim.x) {
                const int sig_tile = u / chunks, chunk = u - sig_tile * chunks;
                const int t0 = chunk * tiles_per_chunk;
                const int t1 = t0 + tiles_per_chunk < tilesN ? t0 + tiles_per_chunk : tilesN;
                for (int t = t0; t < t1; ++t)
                    for (int kb = 0; kb < kblocks; ++kb) {
                        bar_wait(bar_empty + stage * 8, phase ^ 1);
                        const uint32_t full = bar_full + stage * 8;
                        const uint32_t dst = sm_base + stage * ST_STAGE_BYTES;
                        bar_arrive_expect_tx(full, ST_STAGE_BYTES);
                        tma_2d(dst, &mapR, full, kb * (F16 ? 2 * ST_KB : ST_KB), sig_tile * ST_SIG);
                        tma_2d(dst + ST_A_BYTES, &mapA, full, kb * (F16 ? 2 * ST_KB : ST_KB), t * ST_ATOM);
                        if (++stage == ST_STAGES) { stage = 0; phase ^= 1; }
                    }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {                                           // ---- MMA issuer
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int u = blockIdx.x; u < units; u += gridDim.x) {
                const int sig_tile = u / chunks, chunk = u - sig_tile * chunks;
                const int t0 = chunk * tiles_per_chunk;
                const int t1 = t0 + tiles_per_chunk < tilesN ? t0 + tiles_per_chunk : tilesN;
                for (int t = t0; t < t1; ++t) {
                    bar_wait(bar_tempty + acc * 8, acc_phase ^ 1);                  // the epilogue has drained this stage
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(acc * ST_ATOM);
                    for (int kb = 0; kb < kblocks; ++kb) {
                        bar_wait(bar_full + stage * 8, phase);
                        tc_fence_after();
                        const uint32_t sa = sm_base + stage * ST_STAGE_BYTES;
                        const uint64_t adesc = umma_desc(sa), bdesc = umma_desc(sa + ST_A_BYTES);
#pragma unroll
                        for (int k = 0; k < ST_KB / 8; ++k) {                       // K = 8 TF32 (16 FP16) = 32 bytes per instruction
                            if constexpr (F16) umma_f16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), (uint32_t)((kb | k) != 0));
                            else umma_tf32(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), (uint32_t)((kb | k) != 0));
                        }
                        umma_commit(bar_empty + stage * 8);                         // stage free once these MMAs have read it
                        if (++stage == ST_STAGES) { stage = 0; phase ^= 1; }
                    }
                    umma_commit(bar_tfull + acc * 8);
                    acc ^= 1; if (acc == 0) acc_phase ^= 1;
                }
            }
        }
    } else {                                                       // ---- epilogue warps: TMEM lane quadrant = warp % 4
        // (a warp may only touch the 32 TMEM lanes of its quadrant.)  Warps 2-5 scan the first COLS_PER_WARP columns of every tile,
        // warps 6-9 the next, ...: the scan -- one compare and one (rarely taken, but warp-divergent) branch per element -- is the
        // critical path of the FP16 pass (ncu source view: ~80 % of the samples in these warps, tensor pipe 41 % active with
        // four of them: 1.29 ms per pass; eight: 0.96 ms), so it is spread over ST_EPI_PER_QUAD warps per scheduler.
        const int quad = warp & 3, half = (warp - 2) >> 2;             // `half`: which share of the columns (0 .. ST_EPI_PER_QUAD - 1)
        constexpr int COLS_PER_WARP = ST_ATOM / ST_EPI_PER_QUAD;
        float* mrg_base = reinterpret_cast<float*>(sm + ST_STAGES * ST_STAGE_BYTES + ST_BAR_BYTES) +
                          quad * ((ST_EPI_PER_QUAD - 1) * 2 * SCREEN_T * 32);
        int acc = 0; uint32_t acc_phase = 0;
        for (int u = blockIdx.x; u < units; u += gridDim.x) {
            const int sig_tile = u / chunks, chunk = u - sig_tile * chunks;
            const int t0 = chunk * tiles_per_chunk;
            const int t1 = t0 + tiles_per_chunk < tilesN ? t0 + tiles_per_chunk : tilesN;
            float v[SCREEN_T]; int id[SCREEN_T];
            float vmin = -1.0f;
#pragma unroll
            for (int p = 0; p < SCREEN_T; ++p) { v[p] = -1.0f; id[p] = -1; }
            for (int t = t0; t < t1; ++t) {
                bar_wait(bar_tfull + acc * 8, acc_phase);
                tc_fence_after();
                const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * ST_ATOM);
#pragma unroll 1
                for (int c = half * (COLS_PER_WARP / 32); c < (half + 1) * (COLS_PER_WARP / 32); ++c) {
                    uint32_t r[32];
                    tmem_ld32(trow + (uint32_t)(c * 32), r);
                    tmem_ld_wait();
                    const int base = t * ST_ATOM + c * 32;
#pragma unroll
                    for (int e = 0; e < 32; ++e) {
                        const float x = fabsf(__uint_as_float(r[e]));
                        if (x > vmin && base + e < N) top_replace(v, id, vmin, x, base + e + idx_offset);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) bar_arrive(bar_tempty + acc * 8);
                acc ^= 1; if (acc == 0) acc_phase ^= 1;
            }
            // merge the two column halves: the second warp of the quadrant publishes its set, the first absorbs it
            if (half > 0) {
                float* mrg_v = mrg_base + (half - 1) * (2 * SCREEN_T * 32);
                int* mrg_i = reinterpret_cast<int*>(mrg_v + SCREEN_T * 32);
#pragma unroll
                for (int p = 0; p < SCREEN_T; ++p) { mrg_v[p * 32 + lane] = v[p]; mrg_i[p * 32 + lane] = id[p]; }
                asm volatile("bar.sync %0, %1;" ::"r"(1 + quad), "n"(32 * ST_EPI_PER_QUAD) : "memory");      // published
                asm volatile("bar.sync %0, %1;" ::"r"(1 + quad), "n"(32 * ST_EPI_PER_QUAD) : "memory");      // consumed: the area may be rewritten
                continue;
            }
            asm volatile("bar.sync %0, %1;" ::"r"(1 + quad), "n"(32 * ST_EPI_PER_QUAD) : "memory");
#pragma unroll 1
            for (int w = 0; w < ST_EPI_PER_QUAD - 1; ++w) {
                const float* mrg_v = mrg_base + w * (2 * SCREEN_T * 32);
                const int* mrg_i = reinterpret_cast<const int*>(mrg_v + SCREEN_T * 32);
#pragma unroll
                for (int p = 0; p < SCREEN_T; ++p) {
                    const float x = mrg_v[p * 32 + lane];
                    const int xi = mrg_i[p * 32 + lane];
                    if (x > vmin) top_replace(v, id, vmin, x, xi);
                }
            }
            asm volatile("bar.sync %0, %1;" ::"r"(1 + quad), "n"(32 * ST_EPI_PER_QUAD) : "memory");
            top_sort(v, id);
            const int sig = sig_tile * ST_SIG + quad * 32 + lane;
            if (sig < nsig) {
                float4* ov = reinterpret_cast<float4*>(cval + ((size_t)sig * chunks + chunk) * SCREEN_T);
                int4* oi = reinterpret_cast<int4*>(cidx + ((size_t)sig * chunks + chunk) * SCREEN_T);
#pragma unroll
                for (int p = 0; p < SCREEN_T; p += 4) {
                    ov[p / 4] = make_float4(v[p], v[p + 1], v[p + 2], v[p + 3]);
                    oi[p / 4] = make_int4(id[p], id[p + 1], id[p + 2], id[p + 3]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ST_TMEM_COLS) : "memory");
    }
}

// out[r + c * ld_out] = tf32(in[r + c * ld_in]) for r < rows, 0 for rows <= r < ld_out.
template <typename T>
__global__ void __launch_bounds__(256) to_tf32_kernel(const T* __restrict__ in, long long ld_in, float* __restrict__ out,
                                                     long long ld_out, int rows, long long cols) {
    const long long total = ld_out * cols;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long c = e / ld_out;
        const int r = (int)(e - c * ld_out);
        out[e] = r < rows ? tf32_round((float)in[r + c * ld_in]) : 0.0f;
    }
}

// FP16 copy of a column-major matrix for the kind::f16 pass: out[r + c * ld_out] = half(in[r + c * ld_in] * scale_c), zero below
// `rows`.  scale_c = `scale` for every column (the dictionary), or -- per_col_out given -- the power of two that brings the
// column's own norm into [2^11, 2^12) (screen_rscale; residuals), stored there.  One CTA per column (grid-stride).
__global__ void __launch_bounds__(256) to_f16_kernel(const double* __restrict__ in, long long ld_in, __half* __restrict__ out,
                                                    long long ld_out, int rows, long long cols, double scale,
                                                    double* __restrict__ per_col_out) {
    __shared__ double red[8];
    for (long long c = blockIdx.x; c < cols; c += gridDim.x) {
        const double* x = in + c * ld_in;
        double sc = scale;
        if (per_col_out) {
            double s2 = 0.0;
            for (int r = threadIdx.x; r < rows; r += 256) s2 = fma(x[r], x[r], s2);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, off);
            __syncthreads();
            if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s2;
            __syncthreads();
            double tot = 0.0;
            for (int w = 0; w < 8; ++w) tot += red[w];
            sc = screen_rscale(sqrt(tot));
            if (threadIdx.x == 0) per_col_out[c] = sc;
        }
        for (int r = threadIdx.x; r < ld_out; r += 256) out[c * ld_out + r] = r < rows ? __double2half(x[r] * sc) : __double2half(0.0);
    }
}

}  // namespace

// Atom chunks per signal tile; a work unit is (128 signals) x (one chunk), dealt round-robin to one persistent CTA per SM.
// Cost model, fitted to config 2 (profiles/screen_f16_r02.md): a unit costs its tiles plus ~1.3 tiles' worth of overhead (the
// insertion storm while the unit's top-8 fills up, merge, sort, store), and the pass takes ceil(units / SMs) rounds of
// units.  65 536 signals: 2 chunks, 7 rounds of 16 tiles (0.95 ms); 8192 signals (the strong-scaling share of one of 8
// GPUs): the earlier rule ("at least 4 rounds") took 16 chunks, 7 rounds of 2-tile units, 199 us, where 4 chunks (2 rounds
// of 8-tile units) model at 160 us.  One chunk is never taken when two are possible: with only 8 candidates per signal
// the update falls back to exact scans more often (measured: +0.25 ms per update!).  A single round is avoided when two
// rounds cost within 10 % of it: with one round, an SM that is briefly busy with another stream's kernel (the next chunk's
// input check in the pipelined one-shot call) delays the whole pass by a unit.
int screen_chunks_for(int N, int nsig, int num_sms) {
    const int tilesN = (N + ST_ATOM - 1) / ST_ATOM;
    const long long sig_tiles = (nsig + ST_SIG - 1) / ST_SIG;
    static const int forced = [] { const char* e = getenv("CSB200_SCREEN_CHUNKS"); return e ? atoi(e) : 0; }();
    constexpr double UNIT_OVERHEAD_TILES = 1.3;
    int best = 1, best2 = 0;                                       // best overall, best among the choices with >= 2 rounds
    double cost = 1e300, cost2 = 1e300;
    for (int c = 1; c <= SCREEN_MAX_CHUNKS; c *= 2) {
        if (c > tilesN) break;
        const int tpc = (tilesN + c - 1) / c;
        if ((tilesN + tpc - 1) / tpc != c) continue;               // a chunk count that would leave an empty chunk
        if (forced == c) return c;
        if (c == 1 && tilesN >= 2) continue;
        const long long rounds = (sig_tiles * c + num_sms - 1) / num_sms;
        const double t = (double)rounds * (tpc + UNIT_OVERHEAD_TILES);
        if (t < cost) { cost = t; best = c; }
        if (rounds >= 2 && t < cost2) { cost2 = t; best2 = c; }
    }
    return best2 && cost2 <= 1.10 * cost ? best2 : best;
}

int screen_chunk_atoms(int N, int chunks) {
    const int tilesN = (N + ST_ATOM - 1) / ST_ATOM;
    return (tilesN + chunks - 1) / chunks * ST_ATOM;
}

cudaError_t corr_screen_setup() {
    cudaError_t e = cudaFuncSetAttribute(corr_screen_tf32_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, st_smem_bytes(4));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(corr_screen_tf32_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, st_smem_bytes(3));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(corr_screen_tf32_kernel<3>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(corr_screen_tf32_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, st_smem_bytes(4));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(corr_screen_tf32_kernel<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, st_smem_bytes(3));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(corr_screen_tf32_kernel<3, true>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    return e;
}

cudaError_t launch_corr_screen(const CUtensorMap* mapR32, const CUtensorMap* mapA32, int N, int nsig, int ld32, int chunks,
                               int idx_offset, float* cval, int* cidx, int num_sms, cudaStream_t st, int stages, bool f16) {
    if (nsig <= 0 || N <= 0) return cudaSuccess;
    const int tilesN = (N + ST_ATOM - 1) / ST_ATOM;
    const int tpc = (tilesN + chunks - 1) / chunks;
    const int sig_tiles = (nsig + ST_SIG - 1) / ST_SIG;
    const int units = sig_tiles * chunks;
    const int grid = units < num_sms ? units : num_sms;
    if (f16) {                                                     // ld32 counts halves here: 64 per K-block
        if (stages == 3)
            corr_screen_tf32_kernel<3, true><<<grid, ST_THREADS, st_smem_bytes(3), st>>>(*mapR32, *mapA32, N, nsig, ld32 / (2 * ST_KB), tilesN,
                                                                                         chunks, tpc, units, idx_offset, cval, cidx);
        else
            corr_screen_tf32_kernel<4, true><<<grid, ST_THREADS, st_smem_bytes(4), st>>>(*mapR32, *mapA32, N, nsig, ld32 / (2 * ST_KB), tilesN,
                                                                                         chunks, tpc, units, idx_offset, cval, cidx);
        return cudaGetLastError();
    }
    if (stages == 3)
        corr_screen_tf32_kernel<3><<<grid, ST_THREADS, st_smem_bytes(3), st>>>(*mapR32, *mapA32, N, nsig, ld32 / ST_KB, tilesN, chunks,
                                                                               tpc, units, idx_offset, cval, cidx);
    else
        corr_screen_tf32_kernel<4><<<grid, ST_THREADS, st_smem_bytes(4), st>>>(*mapR32, *mapA32, N, nsig, ld32 / ST_KB, tilesN, chunks,
                                                                               tpc, units, idx_offset, cval, cidx);
    return cudaGetLastError();
}

cudaError_t launch_to_tf32(const void* in, bool f32, long long ld_in, float* out, long long ld_out, int rows, long long cols,
                           cudaStream_t st) {
    if (cols <= 0) return cudaSuccess;
    const long long total = ld_out * cols;
    const int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    if (f32) to_tf32_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(in), ld_in, out, ld_out, rows, cols);
    else to_tf32_kernel<double><<<grid, 256, 0, st>>>(static_cast<const double*>(in), ld_in, out, ld_out, rows, cols);
    return cudaGetLastError();
}

cudaError_t launch_to_f16(const double* in, long long ld_in, void* out, long long ld_out, int rows, long long cols, double scale,
                          double* per_col_scale, cudaStream_t st) {
    if (cols <= 0) return cudaSuccess;
    const int grid = (int)(cols < 148 * 16 ? cols : 148 * 16);
    to_f16_kernel<<<grid, 256, 0, st>>>(in, ld_in, static_cast<__half*>(out), ld_out, rows, cols, scale, per_col_scale);
    return cudaGetLastError();
}

}  // namespace csb
