// Shared declarations for libcsb200 (B200 / sm_100a greedy pursuit).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <limits.h>

namespace csb {

// Atoms are grouped in blocks of PBLK for the fused |c| top-s epilogue: every correlation
// kernel emits, per (atom block, signal), the top-s candidates of that block.  One record
// per candidate: |c| as double (FP32 dictionaries widen) and the GLOBAL 0-based atom index.
constexpr int PBLK = 64;
constexpr int MAX_S = 64;          // candidates per 64-atom block the fused DMMA epilogue can emit (= the block size)
constexpr int GOMP_MAX_L = 256;    // atoms per gomp update! (l); larger l: CSB200_ERR_UNSUPPORTED
constexpr int ROW_ALIGN = 16;      // leading dimensions are padded to 16 elements (128 B for f64)
// TF32 screening pass of the batched omp / mp solves (corr_screen_tf32.cu): per (signal, atom chunk) the SCREEN_T largest
// |c~|, with |c~_j - <a_j, r>| <= screen_kappa(M) * max_j ||a_j|| * ||r||.  Operands are rounded to TF32 (2^-11 each, so at
// most 2^-10 + 2^-22 per product; summed with Cauchy-Schwarz: sum |a_i r_i| <= ||a|| ||r||); the FP32 accumulation of M
// exact products adds at most M * 2^-22 of sum |a_i r_i| even if every add TRUNCATES (one ulp each); 5 % margin on top.
// M = 1024: 1.28e-3; M = 4096: 2.05e-3.  tests/test_gpu_screen.py measures the actual error (Gaussian data: < 0.1 of the
// bound; all-positive operands, where truncation would add up: see the test) against it.
// omp_update_kernel experiments (round 2): defaults of CSB200_UPD_RING (cp.async column ring depth, 0 = register path) and
// CSB200_UPD_HINTS (bit 0 streaming b / r / r32, bit 1 L2 evict_last on the dictionary gathers)
constexpr int UPD_RING_DEFAULT = 0;
constexpr int UPD_HINTS_DEFAULT = 0;
// CSB200_UPD_DEFER: 0 = omp_update_kernel down-dates r itself; k = 1, 2, 4: the residual sweep of the screened omp loop runs as
// separate launches over slices of k * 256 rows of ALL signals (dictionary rows of one slice stay in the L2)
constexpr int UPD_DEFER_DEFAULT = 2;
// CSB200_UPD_WARP=1 (needs the deferred sweep, a Gram matrix and k <= 32): selection + append by one warp per signal
constexpr int UPD_WARP_DEFAULT = 1;
// CSB200_SCREEN_F16: the screening pass of batched omp on FP16 operands (kind::f16) instead of TF32
constexpr int SCREEN_F16_DEFAULT = 1;
constexpr int SCREEN_T = 8;
constexpr int SCREEN_MAX_CHUNKS = 16;
constexpr int SCREEN_MAX_ROWS = 8192;
__host__ __device__ inline double screen_kappa(int M) { return 1.05 * (9.765625e-4 + 2.384185791015625e-7 + (double)M * 2.384185791015625e-7); }
constexpr double SCREEN_NORM_MIN = 1e-18, SCREEN_NORM_MAX = 1e18;   // residual norms outside: exact scan (FP32 range)

// FP16 screening (round 2; kind::f16: twice the tensor rate of kind::tf32, half the operand bytes, the SAME 11 significant
// bits): the residual of a signal is stored as half(r * p) with p the power of two that puts the norm the residual had BEFORE the
// update into [2^11, 2^12) (norms only shrink in omp, so |entries| <= 2^12 << 65504), the dictionary as half(A * 2^sA) with
// 2^sA max_j ||a_j|| in [2^11, 2^12).  Rounding of normal halves is <= 2^-11 relative as with TF32; entries below 2^-14 are
// charged the full 2^-14 (valid whether the tensor core keeps or flushes subnormals): on the residual side that is
// sqrt(M) 2^-14 max||a|| / p in true units (StateArgs::scr_abs / p; 1e-3 of the relative term in a normal iteration, dominant
// only when an update shrinks the residual by > 1e3, where it widens the window up to an exact scan), on the dictionary side
// a relative sqrt(M) 2^-25 (folded into scr_bound).
__host__ __device__ inline double screen_kappa_f16(int M) {
    return screen_kappa(M) + 1.05 * sqrt((double)M) * 2.98023223876953125e-8;
}
__host__ __device__ inline double screen_rscale(double nr_old) {
    if (!(nr_old >= 1e-30 && nr_old <= 1e30)) return 1.0;          // outside SCREEN_NORM_MIN..MAX anyway: exact scan
    return ldexp(1.0, 11 - ilogb(nr_old));
}
#ifdef __CUDACC__
__device__ __forceinline__ float tf32_round(float x) {          // round to nearest, ties away: 10 explicit mantissa bits
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}
// one entry of the screening pass's copy of a residual: TF32-rounded float, or scaled half (base points at the signal's row 0)
__device__ __forceinline__ void screen_store(void* base, int row, double val, int f16, double p, bool streaming) {
    if (f16) {
        const __half h = __double2half(val * p);
        if (streaming) __stcs(reinterpret_cast<unsigned short*>(base) + row, __half_as_ushort(h));
        else reinterpret_cast<__half*>(base)[row] = h;
    } else {
        const float f = tf32_round((float)val);
        if (streaming) __stcs(reinterpret_cast<float*>(base) + row, f);
        else reinterpret_cast<float*>(base)[row] = f;
    }
}
#endif

// Candidate order shared by every kernel: larger |c| first, lower index on ties
// (Julia `argmax` = first maximal index, src/matchingpursuit.jl:184;
//  `partialsortperm(.., rev=true)` = Base.Order.Perm tie-break, :192).
__host__ __device__ __forceinline__ bool cand_better(double v, int i, double bv, int bi) {
    return (v > bv) || (v == bv && i < bi);
}

struct CorrArgs {
    const void* A;       // dictionary, ld x N
    const void* R;       // residuals,  ld x nsig
    int M, ld, N, nsig;
    int S;               // candidates per atom block
    int P;               // number of atom blocks = ceil(N / PBLK)
    int idx_offset;      // global index of this shard's first atom
    double* pval;        // [nsig][P][S]
    int* pidx;           // [nsig][P][S]
    int dense_ld = 0;    // > 0: store |c| itself, pval[sig * dense_ld + atom] (DMMA path only), and leave pidx alone
    int l2_policy = 0;   // GEMV: L2 cache policy of the dictionary loads (set by the launcher; see corr_gemv.cu)
};

// correlation kernels (corr_gemm_f64.cu, corr_gemv.cu)
cudaError_t launch_corr_gemm_f64(const CUtensorMap* mapA, const CUtensorMap* mapR, const CorrArgs& a,
                                 int num_sms, cudaStream_t st);
cudaError_t corr_gemm_f64_setup();   // one-time cudaFuncSetAttribute
int corr_gemm_f64_block();           // atoms per candidate block of the DMMA GEMM epilogue (32 or 64)
// Plain store epilogue of the same DMMA GEMM: C[sig + atom * ldc] = <A[:, atom], R[:, sig]> (used for A'A).
cudaError_t launch_gemm_f64_store(const CUtensorMap* mapA, const CUtensorMap* mapR, int N, int nsig, int ld,
                                  double* C, long long ldc, int num_sms, cudaStream_t st);
// Forward-regression pass (EPI_OLS epilogue): candidates are (delta2, atom); mapQ == nullptr on the first step.
cudaError_t launch_corr_gemm_f64_ols(const CUtensorMap* mapA, const CUtensorMap* mapR, const CUtensorMap* mapQ,
                                     const CorrArgs& a, double* resc, long long ldr, int num_sms, cudaStream_t st);
// TF32 screening pass (corr_screen_tf32.cu): cval / cidx [nsig][chunks][SCREEN_T]; ld32 = rows of the FP32 operands (multiple of 32)
int screen_chunks_for(int N, int nsig, int num_sms);
int screen_chunk_atoms(int N, int chunks);      // atoms covered by one chunk (whole 256-atom tiles)
cudaError_t corr_screen_setup();
cudaError_t launch_corr_screen(const CUtensorMap* mapR32, const CUtensorMap* mapA32, int N, int nsig, int ld32, int chunks,
                               int idx_offset, float* cval, int* cidx, int num_sms, cudaStream_t st, int stages = 4, bool f16 = false);
cudaError_t launch_to_f16(const double* in, long long ld_in, void* out, long long ld_out, int rows, long long cols, double scale,
                          double* per_col_scale, cudaStream_t st);
cudaError_t launch_to_tf32(const void* in, bool f32, long long ld_in, float* out, long long ld_out, int rows, long long cols,
                           cudaStream_t st);
// GEMV pass: a.P = corr_gemv_blocks(...) CTAs per signal, each emitting the top-S of its contiguous atom range.
int corr_gemv_blocks(int N, int ld, bool f32, int S, int num_sms);
cudaError_t launch_corr_gemv(const CorrArgs& a, bool f32, cudaStream_t st);
cudaError_t launch_corr_naive(const CorrArgs& a, bool f32, cudaStream_t st);

// per-signal state (update.cu)
struct StateArgs {
    const void* A;        // ld x N
    const void* B;        // ld x nsig   right-hand sides
    void* R;              // ld x nsig   residuals (output)
    int M, ld, N, nsig, kcap;
    int S;                // candidates stored per atom block in pval/pidx
    int P;
    int take;             // atoms to append in this update (1 for omp, l or k%l for gomp)
    int idx_offset;       // this shard's first atom (A is indexed with global index - idx_offset)
    int ignore_done;      // gomp remainder step runs even after an eps-break (matchingpursuit.jl:134-137)
    double eps;
    const double* pval;   // [nsig][P][S]
    const int* pidx;
    int* nnz;             // [nsig]
    int* sel;             // [nsig][kcap]  support in selection order (global atom index)
    double* Rf;           // [nsig][kcap*kcap] column-major upper-triangular INVERSE factor R^{-1} (append order)
    double* z;            // [nsig][kcap]  Q'b
    double* x;            // [nsig][kcap]  coefficients aligned with sel
    double* resnorm;      // [nsig]
    int* iters;           // [nsig]
    int* done;            // [nsig]  eps-break flag
    int* flags;           // [nsig]  bit0: dependent atom skipped, bit1: no candidate, bit2: non-finite input
    const double* gram;   // optional N x N Gram matrix A'A (ld = N), FP64 dictionaries only; nullptr = not available
    // forward regression (src/forward.jl): non-null resc switches the update kernel to `forward_step!` semantics
    double* resc = nullptr;     // [N][ldr]  OLS rescaling ||a_j||^2 - ||Q1'a_j||^2 per (atom, signal); +Inf marks active atoms
    long long ldr = 0;          // leading dimension of resc (even, >= nsig)
    double* qnew = nullptr;     // [nsig][ld] newest orthonormal direction q_t of each signal (zero if none was added)
    int dense_ld = 0;           // > 0: pval is the dense |A'r| matrix [nsig][dense_ld] (no pidx); 0: per-block candidates
    double max_eps = 0.0;       // forward_step! returns false unless ||r|| > max_eps   (:60)
    double min_delta2 = 0.0;    // ... and unless min_delta^2 < max_j delta2_j          (:63)
    // Deferred, row-sliced residual sweep of the screened omp loop (update.cu, omp_residual_slice_kernel): omp_update_kernel
    // leaves y = R^{-1}Q'a_j, gamma = z_t / rho and the old support size here instead of down-dating r itself
    double* def_y = nullptr;    // [nsig][kcap]
    double* def_gam = nullptr;  // [nsig]
    int* def_t = nullptr;       // [nsig] active columns to combine (support size before the append); -1 = nothing deferred
    int* slow = nullptr;        // [nsig] warp-per-signal append (omp_append_warp_kernel): 1 = left to omp_update_kernel; nullptr = warp path off
    int* slow_list = nullptr;   // [nsig] the signals with slow[sig] == 1, in arrival order
    int* slow_count = nullptr;  // [1] entries of slow_list (cleared before every launch of the warp kernel)
    int def_kper = 1;           // row slots of 256 rows per slice launch
    double* def_s2 = nullptr;   // [nsig][128] per-thread running sum of squares of the new residual, carried from slice to slice
    int upd_hints = 0;          // omp_update_kernel cache hints: bit 0 streaming loads / stores of b, r, r32; bit 1 L2 evict_last on
                                // the dictionary gathers of the cp.async ring
    int grid_cap = 0;           // > 0: omp_update_kernel runs with at most this many CTAs, each walking several signals
    int max_smem_carveout = 0;  // launch hint: ask for the SM's largest shared-memory carve-out, i.e. the configuration the
                                // DMMA correlation kernel runs under, so that CTAs of both kernels can share an SM
    // TF32 screening (corr_screen_tf32.cu): non-null scr_val switches omp_update_kernel's selection to "exact FP64
    // re-evaluation of the screened candidates" and makes it keep the TF32 copy of the residual up to date
    const float* scr_val = nullptr;   // [nsig][scr_nc] |c~| descending per chunk of SCREEN_T
    const int* scr_idx = nullptr;     // [nsig][scr_nc] global atom index or -1
    int scr_nc = 0;                   // chunks * SCREEN_T
    int scr_chunk_atoms = 0;          // atoms per chunk: chunk c covers local atoms [c * scr_chunk_atoms, (c + 1) * scr_chunk_atoms)
    double scr_bound = 0.0;           // screen_kappa(M) * max_j ||a_j||: E = scr_bound * ||r||
    float* R32 = nullptr;             // [nsig][ld32] TF32-rounded residuals (the screening pass's operand); scr_f16: halves
    int ld32 = 0;                     // elements per signal of R32 (floats, or halves when scr_f16)
    int scr_f16 = 0;                  // 1: FP16 operands (see screen_rscale); candidates arrive scaled by rscale[sig] / scr_invqA
    double* rscale = nullptr;         // [nsig] power of two the stored residual is scaled by
    double scr_abs = 0.0;             // additive term of the bound: E = scr_bound ||r|| + scr_abs / rscale
    double scr_invqA = 1.0;           // 2^-sA
    unsigned long long* scr_stats = nullptr;   // optional counters: [0] signal-updates screened, [1] candidates re-evaluated, [2] exact scans
};
// Acache (optional): ld x kcap buffer holding the active atoms' columns in selection order, with the
// candidate's column already stored in slot nnz (column-sharded mode: the atom may live on a peer).
cudaError_t launch_omp_update(const StateArgs& a, bool f32, cudaStream_t st, const void* Acache = nullptr);
// One 8-CTA cluster per signal (update_cluster.cu): the single-/few-signal paths.
cudaError_t launch_omp_update_cluster(const StateArgs& a, bool f32, cudaStream_t st, const void* Acache = nullptr);
// Whole-solve kernel for small dictionaries (solve_small.cu): mode 0 omp, 1 gomp, 2 mp.
struct SmallSolveArgs {
    int mode, k, l;
    double eps;
    int stride;                 // slots per signal in sel / x
    const int* x0_idx;          // mp warm start (device pointers) or nullptr
    const double* x0_val;
    const int* x0_nnz;
    int x0_stride;
};
bool small_solve_eligible(int ld, int N, int kcap, int nsig, bool f32);
cudaError_t launch_small_solve(const StateArgs& a, const SmallSolveArgs& q, bool f32, cudaStream_t st);
// Cluster-resident variant for <= 8 signals (dictionary split over the shared memory of an 8-CTA cluster per signal).
bool cluster_solve_eligible(int ld, int N, int kcap, int nsig, int take, bool f32);
cudaError_t launch_cluster_solve(const StateArgs& a, const SmallSolveArgs& q, bool f32, cudaStream_t st);
// Whole-solve cooperative kernel for 1..PERSIST_MAX_SIGNALS signals (solve_persist.cu): mode 0 omp, 2 mp.
constexpr int PERSIST_MAX_SIGNALS = 8;
struct PersistArgs {
    const void* A;        // dictionary, ld x N
    const void* B;        // signals,   ld x ns
    void* R;              // residuals, ld x ns (rewritten every update!, read by the workers)
    int M, ld, N, ns, kcap, idx_offset;
    int mode, k, stride;  // stride: slots per signal in sel / x
    double eps;
    int workers;          // GEMV CTAs; grid = ns + workers
    int wcache;           // dictionary columns a worker keeps in shared memory for the whole solve
    int ucache;           // active-atom columns an updater keeps in shared memory
    int nvl;              // register columns of the workers: 16-byte vectors per lane and column (0: none)
    unsigned long long* cand_ll;   // [ns][workers][4] per-worker arg-max of |c| as self-validating words {v lo, v hi, atom, -}
    unsigned long long* r_ll;      // [ns][ld][1 or 2] the residual handed to the workers, same word format
    unsigned long long* bell;      // [ns][workers][BELL_STRIDE] one doorbell per (signal, worker): (epoch + 1) << 32 | residual version
    unsigned* ctrl;                // [0, PERSIST_MAX_SIGNALS) "signal has stopped", then "a CTA gave up"; value = epoch + 1
    unsigned epoch;                // launch number, 16 bits: sequence numbers are (epoch << 16) | version
    long long* dbg;                // optional clock64() stamps (CSB200_PERSIST_DEBUG)
    int* nnz; int* sel; double* x; double* resnorm; int* iters; int* done; int* flags;
};
constexpr size_t PERSIST_CTRL_WORDS = 1 + PERSIST_MAX_SIGNALS;
constexpr int BELL_STRIDE = 32;            // words between doorbells: 256 bytes, i.e. different L2 slices
// fills workers / wcache / ucache and the dynamic shared memory size; false: the shape does not fit this kernel
bool persist_plan(int ld, int N, int kcap, int ns, bool f32, int num_sms, PersistArgs* out, size_t* smem_out);
cudaError_t launch_persist_solve(const PersistArgs& a, bool f32, size_t smem, cudaStream_t st);
size_t omp_update_smem_bytes(int ld, int kcap);            // dynamic shared memory the kernels above need
bool omp_update_uses_block(int ld, int kcap, int take);    // gomp: the block-append variant (which can read dense |A'r|) runs
size_t omp_update_cluster_smem_bytes(int ld, int kcap);
constexpr size_t MAX_DYN_SMEM = 227 * 1024;
// resc[j][s] = ||a_j||^2 for every signal (`sum!(abs2, P.rescaling', P.A)`, src/forward.jl:105), qnew = 0
cudaError_t launch_ols_init(const StateArgs& a, double* colnorm2, cudaStream_t st);
cudaError_t launch_mp_update(const StateArgs& a, bool f32, int iter, int stride, cudaStream_t st);
// Subspace pursuit `update!` / oblivious selection (update.cu sp_update_kernel); k <= SP_MAX_K atoms per acquisition.
constexpr int SP_MAX_K = 1024;
cudaError_t launch_sp_update(const StateArgs& a, bool f32, int k, double delta, int first, int* ndone, cudaStream_t st);
size_t sp_update_smem_bytes(int ld, int kcap);
// dictionary analysis (src/util.jl:2, 96-117): column 2-norms; Babel-function fold over a chunk of atoms
cudaError_t launch_colnorms(const void* A, bool f32, int ld, int N, double* out, cudaStream_t st);
cudaError_t launch_babel_reduce(const StateArgs& a, int k, int col0, double* mu, cudaStream_t st);
cudaError_t launch_reset_state(const StateArgs& a, bool f32, cudaStream_t st);
cudaError_t launch_spacer(unsigned ns, cudaStream_t st);   // a one-warp kernel that sleeps for ns nanoseconds
cudaError_t launch_mp_warmstart(const StateArgs& a, bool f32, const int* x0_idx, const double* x0_val,
                                const int* x0_nnz, int x0_stride, cudaStream_t st);
cudaError_t launch_topk_from_partials(const StateArgs& a, int s, long long* out_idx, double* out_val,
                                      cudaStream_t st);
cudaError_t launch_nonfinite_check(const void* p, size_t n, bool f32, int* flag, cudaStream_t st);
// FP32 -> FP64 copy of a column-major matrix (batched solves on FP32 dictionaries run on an FP64 twin)
cudaError_t launch_widen_f32(const float* in, long long ld_in, double* out, long long ld_out, int rows, long long cols,
                             cudaStream_t st);

}  // namespace csb
