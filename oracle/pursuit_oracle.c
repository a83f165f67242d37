/* CPU oracle for the greedy-pursuit hot path of CompressedSensing.jl in plain C  --  TEST INFRASTRUCTURE ONLY.
 *
 * A second, independent restatement (the first is oracle/pursuit_oracle.py) of
 *   /root/reference/src/matchingpursuit.jl:26-40   (MP:   update!, mp)
 *   /root/reference/src/matchingpursuit.jl:62-91   (OMP:  update!, omp)
 *   /root/reference/src/matchingpursuit.jl:116-148 (GOMP: update!, gomp)
 *   /root/reference/src/matchingpursuit.jl:152-193 (residual!, addindex!, ldiv!!, argmaxinner!)
 *   /root/reference/src/util.jl:118-134            (addindex! with the sorted insert + QR insert)
 * for Float64 dictionaries.  It exists (a) to cross-check the NumPy oracle with different code, and (b) as the CPU
 * baseline of bench.py: the reference's algorithm, one signal per thread on all host cores (pthreads pulling signals
 * from an atomic counter; this image's gcc has no libgomp), which is how a Julia user would spread independent
 * `omp(A, b, k)` calls over `Threads.@threads`.
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load it; nothing under compressedsensing.jl_b200/.
 *
 * PARITY STATUS: value-level parity with the reference is UNPINNED (no golden vectors in the reference, no Julia in
 * this image; see the header of pursuit_oracle.py).  tests/test_oracle_c.py pins this file to the NumPy oracle.
 *
 * The least-squares engine restates the contract of the un-vendored UpdatableQRFactorizations v1.0.0
 * (Manifest.toml:446-450) exactly as oracle/updatable_qr.py does: thin QR of A[:, sort(S)] under column insertion at
 * the sorted position (Gram-Schmidt twice + Givens rotations, the "insert a column" update util.jl:121 names).
 * Indices are 0-based.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

typedef struct {
    int64_t M, cap, t;
    double* Q;      /* M x cap, column-major, orthonormal columns 0..t-1 */
    double* R;      /* cap x cap, column-major, upper triangular t x t   */
    double* w;      /* cap  */
    double* w2;     /* cap  */
    double* v;      /* M    */
    double* z;      /* cap  */
} uqr_t;

static int uqr_init(uqr_t* F, int64_t M, int64_t cap) {
    memset(F, 0, sizeof *F);
    F->M = M; F->cap = cap; F->t = 0;
    F->Q = (double*)malloc(sizeof(double) * (size_t)M * (size_t)cap);
    F->R = (double*)calloc((size_t)cap * (size_t)cap, sizeof(double));
    F->w = (double*)malloc(sizeof(double) * (size_t)cap);
    F->w2 = (double*)malloc(sizeof(double) * (size_t)cap);
    F->v = (double*)malloc(sizeof(double) * (size_t)M);
    F->z = (double*)malloc(sizeof(double) * (size_t)cap);
    return F->Q && F->R && F->w && F->w2 && F->v && F->z ? 0 : -1;
}
static void uqr_free(uqr_t* F) { free(F->Q); free(F->R); free(F->w); free(F->w2); free(F->v); free(F->z); }

static double dot(const double* a, const double* b, int64_t n) {
    double s = 0.0;
#pragma omp simd reduction(+ : s)
    for (int64_t i = 0; i < n; ++i) s += a[i] * b[i];
    return s;
}

/* `add_column!(F, a, pos)` (util.jl:123): see oracle/updatable_qr.py for the derivation. */
static void uqr_add_column(uqr_t* F, const double* a, int64_t pos) {
    const int64_t M = F->M, cap = F->cap, t = F->t;
    double* Q = F->Q; double* R = F->R; double* v = F->v;
    memcpy(v, a, sizeof(double) * (size_t)M);
    for (int64_t i = 0; i < t; ++i) F->w[i] = dot(Q + i * M, a, M);
    for (int64_t i = 0; i < t; ++i) { const double wi = F->w[i]; const double* q = Q + i * M; for (int64_t r = 0; r < M; ++r) v[r] -= q[r] * wi; }
    for (int64_t i = 0; i < t; ++i) F->w2[i] = dot(Q + i * M, v, M);               /* "twice is enough" */
    for (int64_t i = 0; i < t; ++i) { const double wi = F->w2[i]; const double* q = Q + i * M; for (int64_t r = 0; r < M; ++r) v[r] -= q[r] * wi; F->w[i] += wi; }
    const double rho = sqrt(dot(v, v, M));
    double* qn = Q + t * M;
    for (int64_t r = 0; r < M; ++r) qn[r] = v[r] / rho;
    /* R~ = [R[:, :pos]  [w; rho]  R[:, pos:]] : shift columns pos..t-1 one to the right */
    for (int64_t c = t; c > pos; --c) {
        memcpy(R + c * cap, R + (c - 1) * cap, sizeof(double) * (size_t)t);
        R[t + c * cap] = 0.0;
    }
    for (int64_t i = 0; i < t; ++i) R[i + pos * cap] = F->w[i];
    R[t + pos * cap] = rho;
    for (int64_t c = 0; c < pos; ++c) R[t + c * cap] = 0.0;
    /* Givens rotations on row pairs (i-1, i), i = t .. pos+1, push the spike of column `pos` back up */
    for (int64_t i = t; i > pos; --i) {
        const double x = R[(i - 1) + pos * cap], y = R[i + pos * cap];
        const double h = hypot(x, y);
        if (h == 0.0) continue;
        const double c = x / h, s = y / h;
        for (int64_t col = 0; col <= t; ++col) {
            const double u = R[(i - 1) + col * cap], l = R[i + col * cap];
            R[(i - 1) + col * cap] = c * u + s * l;
            R[i + col * cap] = -s * u + c * l;
        }
        R[i + pos * cap] = 0.0;
        double* qa = Q + (i - 1) * M; double* qb = Q + i * M;
        for (int64_t r = 0; r < M; ++r) {
            const double u = qa[r], l = qb[r];
            qa[r] = c * u + s * l;
            qb[r] = -s * u + c * l;
        }
    }
    F->t = t + 1;
}

/* `ldiv!(F, r)` (matchingpursuit.jl:175): y = argmin ||A_S y - b||, logical (sorted) order. */
static void uqr_solve(uqr_t* F, const double* b, double* y) {
    const int64_t M = F->M, cap = F->cap, t = F->t;
    for (int64_t i = 0; i < t; ++i) F->z[i] = dot(F->Q + i * M, b, M);
    for (int64_t i = t - 1; i >= 0; --i) {
        double acc = F->z[i];
        for (int64_t c = i + 1; c < t; ++c) acc -= F->R[i + c * cap] * y[c];
        y[i] = acc / F->R[i + i * cap];
    }
}

/* `residual!` (matchingpursuit.jl:158-161): copyto!(r, b); mul!(r, A, x, -1, 1) -- SparseArrays walks the stored
 * entries in ascending index order: av = v * (-1); r[i] += A[i, j] * av (separate multiply and add). */
__attribute__((optimize("fp-contract=off")))
static void residual(const double* A, int64_t M, int64_t lda, const double* b, const int64_t* nzind, const double* nzval,
                     int64_t nnz, double* r) {
    memcpy(r, b, sizeof(double) * (size_t)M);
    for (int64_t e = 0; e < nnz; ++e) {
        const double av = nzval[e] * -1.0;
        const double* a = A + nzind[e] * lda;
        for (int64_t i = 0; i < M; ++i) { const double p = a[i] * av; r[i] += p; }
    }
}

/* `argmaxinner!(P)` (:181-185): |A'r|, first maximal index (strict > scan; the findmax override of util.jl:173-189
 * ignores NaN).  Also returns the signed winner correlation for mp. */
static int64_t argmaxinner(const double* A, int64_t M, int64_t N, int64_t lda, const double* r, double* signed_c) {
    int64_t best = 0;
    double bv = -1.0, bc = 0.0;
    for (int64_t j = 0; j < N; ++j) {
        const double c = dot(A + j * lda, r, M);
        const double v = fabs(c);
        if (v > bv) { bv = v; best = j; bc = c; }
    }
    if (signed_c) *signed_c = bc;
    return best;
}

static int64_t find_sorted(const int64_t* ind, int64_t n, int64_t j, int* found) {
    int64_t lo = 0;
    while (lo < n && ind[lo] < j) ++lo;
    *found = lo < n && ind[lo] == j;
    return lo;
}

/* `addindex!(x, AiQR, a, i)` (util.jl:118-126): no-op when already stored; x[i] = NaN -> sorted insert; QR insert. */
static int addindex(int64_t* nzind, double* nzval, int64_t* nnz, uqr_t* F, const double* A, int64_t lda, int64_t j) {
    int found;
    const int64_t pos = find_sorted(nzind, *nnz, j, &found);
    if (found) return 0;
    for (int64_t e = *nnz; e > pos; --e) { nzind[e] = nzind[e - 1]; nzval[e] = nzval[e - 1]; }
    nzind[pos] = j; nzval[pos] = NAN;
    *nnz += 1;
    uqr_add_column(F, A + j * lda, pos);
    return 1;
}

static double norm2(const double* r, int64_t M) { return sqrt(dot(r, r, M)); }

/* omp (algo 0) and gomp (algo 1, l atoms per update): matchingpursuit.jl:73-82 / 126-139.
 * order: atoms in the order they were appended (cap slots, -1 padded); nzind / nzval: the SparseVector. */
static int solve_omp_gomp(int algo, const double* A, int64_t M, int64_t N, int64_t lda, const double* b, int64_t k,
                          int64_t l, double eps, int64_t cap, int64_t* order, int64_t* nzind, double* nzval,
                          int64_t* nnz_out, double* resnorm, int64_t* iters_out) {
    /* capacity: omp passes k to the OMP constructor (:75), gomp does not (:128 -> M) */
    int64_t qcap = algo == 0 ? (k < M ? k : M) : M;
    if (qcap > N) qcap = N;
    if (qcap < 1) qcap = 1;
    uqr_t F;
    double* r = (double*)malloc(sizeof(double) * (size_t)M);
    double* absc = algo == 1 ? (double*)malloc(sizeof(double) * (size_t)N) : NULL;
    if (uqr_init(&F, M, qcap) || !r || (algo == 1 && !absc)) { uqr_free(&F); free(r); free(absc); return -1; }
    int64_t nnz = 0, nord = 0, iters = 0;
    double nr = norm2(b, M);
    for (int64_t e = 0; e < cap; ++e) order[e] = -1;
    const int64_t loops = algo == 0 ? k : k / l;
    const int64_t rem = algo == 0 ? 0 : k % l;
    for (int64_t it = 0; it < loops + (rem > 0); ++it) {
        const int is_rem = it == loops;
        const int64_t take = algo == 0 ? 1 : (is_rem ? rem : l);
        if (nnz < M) {                                                   /* :63 / :117 */
            residual(A, M, lda, b, nzind, nzval, nnz, r);                /* :64 / :118 */
            if (algo == 0) {
                const int64_t j = argmaxinner(A, M, N, lda, r, NULL);    /* :65 over ALL atoms */
                if (nnz < qcap && addindex(nzind, nzval, &nnz, &F, A, lda, j)) {   /* :66-67 */
                    if (nord < cap) order[nord++] = j;
                    uqr_solve(&F, b, nzval);                             /* :68 */
                }
            } else {
                for (int64_t j = 0; j < N; ++j) absc[j] = fabs(dot(A + j * lda, r, M));
                double pv = 0.0; int64_t pi = -1;                        /* partialsortperm(rev=true): value desc, index asc */
                for (int64_t round = 0; round < take && round < N; ++round) {
                    double bv = -1.0; int64_t bi = -1;
                    for (int64_t j = 0; j < N; ++j) {
                        const double v = absc[j];
                        const int ok = round == 0 || v < pv || (v == pv && j > pi);
                        if (ok && v > bv) { bv = v; bi = j; }
                    }
                    if (bi < 0) break;
                    pv = bv; pi = bi;
                    if (nnz < qcap && nnz < M && addindex(nzind, nzval, &nnz, &F, A, lda, bi) && nord < cap) order[nord++] = bi;
                }
                uqr_solve(&F, b, nzval);                                 /* :121 one solve after all inserts */
            }
        }
        residual(A, M, lda, b, nzind, nzval, nnz, r);                    /* :79 / :132 */
        nr = norm2(r, M);
        ++iters;
        if (!is_rem && !(nr >= eps)) { it = loops - 1; }                 /* break; the remainder update still runs (:134-137) */
    }
    *nnz_out = nnz; *resnorm = nr; *iters_out = iters;
    uqr_free(&F); free(r); free(absc);
    return 0;
}

/* mp (matchingpursuit.jl:26-40): exactly k updates, x[i] += <a_i, r>. */
static int solve_mp(const double* A, int64_t M, int64_t N, int64_t lda, const double* b, int64_t k, int64_t cap,
                    int64_t* order, int64_t* nzind, double* nzval, int64_t* nnz_out, double* resnorm,
                    int64_t* iters_out) {
    double* r = (double*)malloc(sizeof(double) * (size_t)M);
    if (!r) return -1;
    int64_t nnz = 0;
    for (int64_t e = 0; e < cap; ++e) order[e] = -1;
    for (int64_t it = 0; it < k; ++it) {
        residual(A, M, lda, b, nzind, nzval, nnz, r);                    /* :27 */
        double c;
        const int64_t j = argmaxinner(A, M, N, lda, r, &c);              /* :28 */
        c = dot(A + j * lda, r, M);                                      /* :29 recomputed signed dot */
        int found;
        const int64_t pos = find_sorted(nzind, nnz, j, &found);
        if (found) nzval[pos] += c;
        else {
            for (int64_t e = nnz; e > pos; --e) { nzind[e] = nzind[e - 1]; nzval[e] = nzval[e - 1]; }
            nzind[pos] = j; nzval[pos] = c; ++nnz;
        }
        if (it < cap) order[it] = j;
    }
    residual(A, M, lda, b, nzind, nzval, nnz, r);
    *nnz_out = nnz; *resnorm = norm2(r, M); *iters_out = k;
    free(r);
    return 0;
}

/* Batch driver: signal s is column s of B (ldb); outputs have `cap` slots per signal (cap >= min(k, M) for omp / gomp,
 * cap >= k for mp).  algo: 0 omp, 1 gomp, 2 mp.  threads <= 0: one per online core.  Returns 0, or -1 (allocation),
 * -2 (bad argument).  *threads_used reports the team size. */
typedef struct {
    int algo;
    const double* A; int64_t M, N, lda;
    const double* B; int64_t ldb, nsig, k, l; double eps; int64_t cap;
    int64_t* order; int64_t* nzind; double* nzval; int64_t* nnz; double* resnorm; int64_t* iters;
    atomic_llong next;
    atomic_int status;
} job_t;

static void* worker(void* arg) {
    job_t* J = (job_t*)arg;
    for (;;) {
        const int64_t s = (int64_t)atomic_fetch_add(&J->next, 1);
        if (s >= J->nsig) break;
        int rc;
        if (J->algo == 2)
            rc = solve_mp(J->A, J->M, J->N, J->lda, J->B + s * J->ldb, J->k, J->cap, J->order + s * J->cap,
                          J->nzind + s * J->cap, J->nzval + s * J->cap, J->nnz + s, J->resnorm + s, J->iters + s);
        else
            rc = solve_omp_gomp(J->algo, J->A, J->M, J->N, J->lda, J->B + s * J->ldb, J->k, J->l, J->eps, J->cap,
                                J->order + s * J->cap, J->nzind + s * J->cap, J->nzval + s * J->cap, J->nnz + s,
                                J->resnorm + s, J->iters + s);
        if (rc) atomic_store(&J->status, -1);
    }
    return NULL;
}

int cs_oracle_solve_batch(int algo, const double* A, int64_t M, int64_t N, int64_t lda, const double* B, int64_t ldb,
                          int64_t nsig, int64_t k, int64_t l, double eps, int64_t cap, int64_t* order, int64_t* nzind,
                          double* nzval, int64_t* nnz, double* resnorm, int64_t* iters, int threads, int* threads_used) {
    if (!A || !B || M <= 0 || N <= 0 || lda < M || ldb < M || nsig < 0 || k < 0 || cap < 1 || !(eps >= 0) ||
        algo < 0 || algo > 2 || (algo == 1 && l < 1))
        return -2;
    if (threads <= 0) threads = (int)sysconf(_SC_NPROCESSORS_ONLN);
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    if ((int64_t)threads > nsig) threads = nsig > 0 ? (int)nsig : 1;
    job_t J = {algo, A, M, N, lda, B, ldb, nsig, k, l, eps, cap, order, nzind, nzval, nnz, resnorm, iters, 0, 0};
    pthread_t tid[256];
    int started = 0;
    for (int i = 1; i < threads; ++i)
        if (pthread_create(&tid[started], NULL, worker, &J) == 0) ++started;
    worker(&J);                                          /* the calling thread works too */
    for (int i = 0; i < started; ++i) pthread_join(tid[i], NULL);
    if (threads_used) *threads_used = started + 1;
    return atomic_load(&J.status);
}
