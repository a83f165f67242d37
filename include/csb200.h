/* csb200.h -- C ABI of libcsb200.so: B200 (sm_100a) greedy pursuit (mp / omp / gomp).
 *
 * This is the drop-in boundary for the hot path of SebastianAment/CompressedSensing.jl.
 * The reference has no FFI of its own (pure Julia over BLAS + UpdatableQRFactorizations);
 * each entry point below replaces the Julia function cited next to it, and the Julia shim
 * (compressedsensing.jl_b200/julia/CompressedSensingB200.jl) `ccall`s exactly these symbols.
 * See INTEGRATION.md for the binding a maintainer would add.
 *
 * Conventions
 *   - Plain C linkage, plain pointers and sizes; no C++/torch types; no exceptions escape.
 *   - Matrices are column-major with a leading dimension in ELEMENTS (Julia `Matrix` layout:
 *     each atom / each signal is one contiguous column).
 *   - Indices crossing the boundary are 0-based int64; -1 pads unused slots.
 *   - Every function returns CSB200_OK (0) or a negative csb200_status.
 *   - The caller owns every host buffer; the library owns all device memory behind the
 *     opaque handles.  Calls on different handles may run concurrently from different host
 *     threads; calls on one handle are serialised internally.
 *   - There is NO CPU fallback: without an sm_100 device the create calls fail with
 *     CSB200_ERR_UNSUPPORTED_ARCH / CSB200_ERR_CUDA.
 */
#ifndef CSB200_H
#define CSB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CSB200_VERSION 100 /* 0.1.0 */

typedef enum csb200_status {
    CSB200_OK = 0,
    CSB200_ERR_INVALID_ARG = -1,      /* null pointer, negative size, k/l out of range            */
    CSB200_ERR_NEGATIVE_EPS = -2,     /* eps < 0: the shim rethrows the reference's String
                                         "ε = ... has to be non-negative" (matchingpursuit.jl:74,127) */
    CSB200_ERR_NONFINITE_INPUT = -3,  /* NaN/Inf in A or b              */
    CSB200_ERR_CUDA = -4,             /* a CUDA runtime call failed; csb200_last_error() has text */
    CSB200_ERR_OOM = -5,              /* device allocation failed                                  */
    CSB200_ERR_UNSUPPORTED_ARCH = -6, /* device is not compute capability 10.x                     */
    CSB200_ERR_UNSUPPORTED = -7,      /* shape / option outside what this build handles            */
    CSB200_ERR_NCCL = -8,             /* NCCL call failed (column-sharded mode)                    */
    CSB200_ERR_DIM_MISMATCH = -9      /* signal length != rows of the dictionary (Julia: DimensionMismatch) */
} csb200_status;

typedef enum csb200_dtype { CSB200_F64 = 0, CSB200_F32 = 1 } csb200_dtype;

typedef struct csb200_dict csb200_dict;   /* a dictionary A resident on one GPU                    */
typedef struct csb200_batch csb200_batch; /* per-batch solver state (signals, residuals, QR, ...)  */

int csb200_version(void);
const char* csb200_strerror(int status);
/* Text of the last CUDA/NCCL failure on the calling thread ("" if none). */
const char* csb200_last_error(void);
/* Number of visible CUDA devices, or a negative status. */
int csb200_device_count(void);

/* ---- dictionary ------------------------------------------------------------------------
 * Replaces the `A` field captured by the constructors MP(A,b) / OMP(A,b,k) / GOMP(A,b,l)
 * (src/matchingpursuit.jl:18-24, 54-60, 108-114).  A is M x N column-major (M rows =
 * signal length, N atoms), lda >= M elements.  The dictionary is copied to `device`
 * (HBM-resident for the life of the handle) and checked for NaN/Inf.
 * FP32 dictionaries: batches created with max_signals >= 24 (and csb200_fr) are solved on an FP64 copy of the
 * dictionary kept on the handle (built on first use; exact float x float products, FP64 accumulation); signals are
 * still passed as FP32.  Smaller batches stream the FP32 dictionary (HBM-bound GEMV path).
 * n_offset / n_total describe a column shard: this handle holds atoms
 * [n_offset, n_offset + N) of a dictionary with n_total atoms; pass 0 and N when unsharded.
 */
int csb200_dict_create(const void* A, int64_t M, int64_t N, int64_t lda, int dtype, int device,
                       csb200_dict** out);
int csb200_dict_create_shard(const void* A, int64_t M, int64_t N, int64_t lda, int dtype, int device,
                             int64_t n_offset, int64_t n_total, csb200_dict** out);
/* Multi-GPU handle (SURVEY.md 8b "Threading": fan-out is library-internal and invisible to the caller).  The
 * dictionary is uploaded to devices[0] and replicated device-to-device onto devices[1..ndev-1] (an entry may repeat:
 * each entry is one worker with its own replica, workspace and stream).  The one-shot entry points below
 * (csb200_omp / gomp / mp / fr / sp / oblivious) split the signals of a call into contiguous ranges, one per worker
 * (at least 256 signals each; fewer signals use fewer workers), and solve them concurrently on internal host
 * threads; signals are independent, so there is no data-path communication and results are bit-identical to the
 * single-device call.  Everything else (batch API, dictionary analysis) runs on devices[0].
 * csb200_dict_devices: number of workers; fills devices[0..min(n, capacity)). */
int csb200_dict_create_multi(const void* A, int64_t M, int64_t N, int64_t lda, int dtype, const int* devices, int ndev,
                             csb200_dict** out);
int csb200_dict_devices(const csb200_dict* dict, int* devices, int capacity);
int csb200_dict_destroy(csb200_dict* dict);
/* Release the device workspace the one-shot calls (csb200_omp/gomp/mp) keep on the handle for reuse. */
int csb200_dict_trim(csb200_dict* dict);
int csb200_dict_shape(const csb200_dict* dict, int64_t* M, int64_t* N, int* dtype, int* device);

/* ---- batch state -----------------------------------------------------------------------
 * One batch holds up to max_signals right-hand sides and, per signal, what the reference
 * keeps in P.r, P.Ar, P.AiQR and x (src/matchingpursuit.jl:44-60): the residual, the
 * support in selection order, the triangular factor of the active atoms, the coefficients.
 * max_sparsity bounds the support size (the `k` of UpdatableQR(T, n, k), :58).
 */
int csb200_batch_create(csb200_dict* dict, int64_t max_signals, int64_t max_sparsity, csb200_batch** out);
int csb200_batch_destroy(csb200_batch* batch);
/* Copy nsig signals (columns of Bmat, M x nsig, ldb >= M elements, dict dtype) host -> device. */
int csb200_batch_upload(csb200_batch* batch, const void* Bmat, int64_t ldb, int64_t nsig);
/* Same, but Bmat is a DEVICE pointer on the dictionary's GPU (device -> device copy). */
int csb200_batch_upload_device(csb200_batch* batch, const void* dBmat, int64_t ldb, int64_t nsig);

/* Solve on the signals currently resident in the batch (no host<->device traffic).
 *   omp  : src/matchingpursuit.jl:73-82   (k update!s, stop once ||r|| < eps; eps >= 0)
 *   gomp : src/matchingpursuit.jl:126-139 (k / l update!s of l atoms, then one of k % l)
 *   mp   : src/matchingpursuit.jl:34-40   (exactly `iters` update!s; optional warm start)
 * Each call blocks until the device work has finished.
 * Notes on arithmetic (same mathematics as the reference, rounding-level differences):
 *   - The residual is DOWN-DATED (r -= q_t q_t'b) and its norm accumulated from it; the reference recomputes
 *     r = b - A x.  The two agree to ~1e-16 ||b||, so an eps-break is decided identically except when eps itself lies at
 *     that rounding level (the default eps(T) on noise-free data solved past its true sparsity): there the decision is
 *     noise in the reference too and nnz / iters may differ by the trailing no-information updates.
 *   - FP32 dictionaries: see csb200_dict_create -- batches of >= 24 signals compute in FP64 on exact FP32 inputs, smaller
 *     ones in FP32 storage with FP64 accumulation; both stay inside the FP32 bound (2e-5) of the reference's Float32
 *     arithmetic, but a signal's last bits can depend on how many signals share its batch.
 */
int csb200_batch_omp(csb200_batch* batch, int64_t k, double eps);
int csb200_batch_gomp(csb200_batch* batch, int64_t l, int64_t k, double eps);
/* x0_*: optional warm start (may be NULL): x0_nnz[s] entries per signal, stored at
 * x0_idx/x0_val[s * x0_stride + j] (0-based atom indices). */
int csb200_batch_mp(csb200_batch* batch, int64_t iters, const int64_t* x0_idx, const double* x0_val,
                    const int64_t* x0_nnz, int64_t x0_stride);

/* Forward regression == OLS == OOMP == ORMP: `fr(A, b, max_eps, min_delta, k)`,
 * src/forward.jl:44-51, each step being `forward_step!` (:56-67): stop unless nnz < M and ||r|| > max_eps; pick
 * argmax_j <a_j, r>^2 / (||a_j||^2 - ||Q1'a_j||^2) over the passive atoms (first index on ties, :62 with the
 * findmax override of src/util.jl:173-189); append it if min_delta^2 < that maximum, else stop.  FP64, unsharded
 * dictionaries only (CSB200_ERR_UNSUPPORTED otherwise).  Outputs as for omp (selection order). */
int csb200_batch_fr(csb200_batch* batch, int64_t k, double max_eps, double min_delta);

/* Subspace pursuit and oblivious selection.
 *   sp        : `sp(A, b, k, delta = 1e-12; maxiter = 16k)`, src/twostage.jl:105-117 -- initial top-k acquisition,
 *               then up to maxiter `update!`s (:94-103: add the top-k of |A'r|, least squares on <= 2k atoms, keep the
 *               k largest |x|, least squares again) until ||r|| <= delta or ||r|| stops decreasing, per signal.
 *               Needs 2k <= M (the reference errors otherwise) and a batch with max_sparsity >= 2k; k <= 256.
 *   oblivious : `oblivious(A, b, k)`, src/oblivious.jl:3-8 -- the k atoms most correlated with b and their
 *               least-squares coefficients (k <= min(M, 256)).
 * Outputs as for omp; a signal's support is reported in the order the final factorisation appended it. */
int csb200_batch_sp(csb200_batch* batch, int64_t k, double delta, int64_t maxiter);
int csb200_batch_oblivious(csb200_batch* batch, int64_t k);

/* Copy results device -> host.  `stride` = slots per signal in sel_idx / coef (>= the k or
 * iters of the last solve).  Any output pointer may be NULL.
 *   sel_idx[s*stride + j]  j-th atom appended for signal s, in SELECTION order (-1 padded).
 *                          For mp: the atom chosen at iteration j (atoms may repeat).
 *   coef[s*stride + j]     omp/gomp: least-squares coefficient of that atom;
 *                          mp: the increment <a_i, r> added at iteration j (src/matchingpursuit.jl:29).
 *   nnz[s]                 number of valid slots.
 *   resnorm[s]             ||b - A x||_2 after the last update.
 *   iters[s]               update!s executed for this signal (an eps-break stops the count).
 */
int csb200_batch_download(csb200_batch* batch, int64_t stride, int64_t* sel_idx, double* coef,
                          int64_t* nnz, double* resnorm, int64_t* iters);

/* Per-signal diagnostics of the last solve (nsig 32-bit words), a bit set:
 *   1   an atom that won the arg-max was numerically dependent on the active set (rho^2 <= 1e-26 ||a||^2) and was NOT
 *       appended -- the reference's add_column! would divide by a ~0 diagonal here;
 *   2   the correlation pass produced no candidate (all-NaN correlations cannot happen on finite input; kept for safety);
 *   16  the active set is ill-conditioned (an appended atom kept < 1e-2 of its squared norm after orthogonalisation,
 *       cond(A_S) >~ 10): the coefficients were refined by two steps of iterative refinement, which restores the
 *       ~cond(A_S) eps accuracy of a backward-stable QR (measured for cond up to 2e8, profiles/conditioning_r02.md). */
int csb200_batch_flags(csb200_batch* batch, int32_t* flags);

/* Timing of the dominant kernel (the correlation pass), measured with CUDA events on the
 * batch's own stream.  enable != 0 turns per-launch event recording on and clears the
 * counters; csb200_batch_corr_time returns the summed duration (ms) and launch count since. */
int csb200_batch_profile(csb200_batch* batch, int enable);
int csb200_batch_corr_time(csb200_batch* batch, double* total_ms, int64_t* launches, int64_t* other_launches);
/* Device time (ms, CUDA events on the batch's stream) of the last omp/gomp/mp solve on this batch:
 * first kernel enqueued to last kernel finished; excludes upload/download. */
int csb200_batch_last_solve_ms(csb200_batch* batch, double* ms);

/* ---- one-shot host-buffer entry points (what `omp(A,b,k)` etc. bind to) ------------------
 * upload + solve + download on a temporary batch.  Outputs as in csb200_batch_download with
 * stride = k (omp/gomp) or iters (mp).
 */
int csb200_omp(csb200_dict* dict, const void* Bmat, int64_t ldb, int64_t nsig, int64_t k, double eps,
               int64_t* sel_idx, double* coef, int64_t* nnz, double* resnorm, int64_t* iters);
int csb200_gomp(csb200_dict* dict, const void* Bmat, int64_t ldb, int64_t nsig, int64_t l, int64_t k,
                double eps, int64_t* sel_idx, double* coef, int64_t* nnz, double* resnorm, int64_t* iters);
int csb200_fr(csb200_dict* dict, const void* Bmat, int64_t ldb, int64_t nsig, int64_t k, double max_eps,
              double min_delta, int64_t* sel_idx, double* coef, int64_t* nnz, double* resnorm, int64_t* iters);
int csb200_sp(csb200_dict* dict, const void* Bmat, int64_t ldb, int64_t nsig, int64_t k, double delta,
              int64_t maxiter, int64_t* sel_idx, double* coef, int64_t* nnz, double* resnorm, int64_t* iters);
int csb200_oblivious(csb200_dict* dict, const void* Bmat, int64_t ldb, int64_t nsig, int64_t k,
                     int64_t* sel_idx, double* coef, int64_t* nnz, double* resnorm);
int csb200_mp(csb200_dict* dict, const void* Bmat, int64_t ldb, int64_t nsig, int64_t iters_k,
              const int64_t* x0_idx, const double* x0_val, const int64_t* x0_nnz, int64_t x0_stride,
              int64_t* sel_idx, double* coef, double* resnorm);

/* ---- dictionary analysis (src/util.jl:2, 96-117) ----------------------------
 * colnorms : out[j] = ||A[:, j]||_2 (what `normalize!` divides by, src/util.jl:59-61).
 * cumbabel : mu[i-1] = Babel function mu_1(i) for i = 1..k -- max over atoms of the sum of the i largest
 *            |<a_j, a_l>|, l != j; `coherence(A)` = `babel(A, 1)` = mu[0].  1 <= k <= min(N, 255). */
int csb200_dict_colnorms(csb200_dict* dict, double* out);
int csb200_dict_cumbabel(csb200_dict* dict, int64_t k, double* mu);

/* ---- batched result format (host-side helper, no GPU work) ---------------------------------
 * Assemble the selection-order outputs of csb200_omp / csb200_gomp (sel_idx, coef, nnz with `stride` slots per
 * signal) into compressed-sparse-column arrays of the N x nsig coefficient matrix -- Julia's SparseMatrixCSC
 * (colptr, rowval, nzval): within each column (signal) row indices ascend, as in the reference's SparseVector.
 * index_base = 1 for Julia, 0 for C / Python.  colptr has nsig + 1 entries; rowval / nzval need sum(nnz). */
int csb200_assemble_csc(int64_t nsig, int64_t stride, const int64_t* sel_idx, const double* coef, const int64_t* nnz,
                        int64_t index_base, int64_t* colptr, int64_t* rowval, double* nzval);

/* ---- column-sharded single-dictionary mode (one process per GPU, NCCL over NVLink) -------
 * Each rank holds a csb200_dict_create_shard() slice.  csb200_comm_* wrap one NCCL
 * communicator; the unique id (128 bytes) is produced on rank 0 and distributed by the host
 * program (torch.distributed / MPI / files).  csb200_omp_sharded runs `omp` for ONE signal
 * with the per-iteration exchange done on the device stream (no host round trip):
 * all ranks return the same result.
 * Exchange transport: by default the ranks map each other's mailboxes (CUDA IPC over NVLink peer access, set up
 * collectively inside the first csb200_omp_sharded call) and one kernel per iteration stores the candidate record
 * into every peer, publishes a sequence number and picks the winner; if any rank cannot map its peers, or
 * CSB200_SHARD_EXCHANGE=nccl is set (on ALL ranks), the records travel through ncclAllGather instead.  Both
 * transports give bit-identical results.  csb200_comm_exchange_mode: transport of the last solve (1 peer memory,
 * 0 NCCL).  A rank that waits longer than CSB200_PEER_TIMEOUT_S (default 60) for a peer fails with CSB200_ERR_NCCL.
 */
#define CSB200_NCCL_ID_BYTES 128
typedef struct csb200_comm csb200_comm;
int csb200_comm_unique_id(void* id_bytes);
int csb200_comm_create(const void* id_bytes, int rank, int nranks, int device, csb200_comm** out);
int csb200_comm_destroy(csb200_comm* comm);
int csb200_comm_exchange_mode(const csb200_comm* comm);
/* Device-time breakdown (CUDA events on the communicator's stream) of the last csb200_omp_sharded call on this rank:
 * ms5 = {whole solve (first correlation pass enqueued -> last update finished), correlation passes, exchange, update,
 * gaps}; the last three are recorded only while CSB200_SHARD_TIMING is 1 (also printed on stderr) or 2.  iters = k. */
int csb200_comm_last_timing(csb200_comm* comm, double* ms5, int64_t* iters);
int csb200_omp_sharded(csb200_dict* shard, csb200_comm* comm, const void* b, int64_t k, double eps,
                       int64_t* sel_idx, double* coef, int64_t* nnz, double* resnorm, int64_t* iters,
                       double* corr_ms);

/* ---- test / debug hooks ------------------------------------------------------------------
 * One correlation pass over the current residuals with the production kernel (impl 0 = auto,
 * 1 = DMMA GEMM, 2 = GEMV) or a naive one-thread-per-dot kernel (impl 3); returns, per signal,
 * the top-`s` atoms (index, |c|) in (value desc, index asc) order.  Used by tests/ only. */
int csb200_debug_corr_topk(csb200_batch* batch, int impl, int64_t s, int64_t* idx, double* val);
/* One screening pass (tcgen05 kernel of corr_screen_tf32.cu, the first stage of large batched omp solves; operands in scaled FP16,
 * or in TF32 with CSB200_SCREEN_F16=0) over the
 * current residuals: val / idx receive [nsig][chunks][8] candidates (|c~| descending per chunk, 0-based atom or -1),
 * *chunks the number of atom chunks, *bound the proven error bound per unit of ||r||: |c~_j - <a_j, r>| <= bound ||r||.
 * val and idx must hold nsig * 16 * 8 entries.  Used by tests/ only. */
int csb200_debug_screen_pass(csb200_batch* batch, float* val, int32_t* idx, int64_t* chunks, double* bound);
/* Which path the last csb200_batch_omp / csb200_batch_mp on this batch took (*path: 0 few-signal / GEMV paths, 1 FP64 DMMA loop, 2 FP64
 * DMMA with the two-half overlap, 3 TF32 screening + exact FP64 re-evaluation, 4 the same with scaled FP16 operands) and the screening counters since the
 * last reset: stats3[0] signal-updates decided from screened candidates, [1] candidates re-evaluated in FP64 (windows
 * holding more than one atom), [2] exact scans of all atoms (incomplete list or residual outside the FP32 range). */
int csb200_batch_screen_stats(csb200_batch* batch, int64_t* path, uint64_t* stats3, int reset);
/* Copy the current residual matrix (M x nsig, dict dtype, ld = M) to the host. */
int csb200_debug_get_residual(csb200_batch* batch, void* out);
/* Few-signal solves (< 24 signals) are captured into a CUDA graph on their second run with the same algorithm, k,
 * l, eps and signal count on a batch and replayed afterwards (CSB200_GRAPH=0 disables it).  Number of solves of this
 * batch that ran as a graph launch. */
int csb200_debug_graph_replays(csb200_batch* batch, int64_t* replays);

#ifdef __cplusplus
}
#endif
#endif /* CSB200_H */
