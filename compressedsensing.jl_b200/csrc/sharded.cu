// Column-sharded single-dictionary OMP: one process per GPU, NCCL over NVLink (see include/csb200.h).
//
// BASELINE config 4 (8192 x 1 048 576 FP32 = 32 GiB) does not fit one GPU's sensible share, and the
// correlation c = A'r is separable by columns: rank g owns atoms [n_offset, n_offset + N_local).
// Everything else in `update!` is O(M k) and is REPLICATED: every rank keeps b, r, the support, R^{-1},
// Q'b and runs the same deterministic update kernel on the same inputs, so replicas stay bit-identical
// and nothing but the candidate atom ever crosses NVLink.
//
// Per iteration, all on one stream with no host round trip:
//   1. fused GEMV + |c| argmax over the local shard                    (corr_gemv.cu)
//   2. local_best: reduce the per-block candidates to one record  { |c|, global index, atom column }
//   3. ncclAllGather of the records (16 B + M elements per rank; 8 x 32 KB at config 4).  NCCL has no
//      MAXLOC and the winner's rank is data dependent, so a broadcast would need a host decision;
//      gathering every rank's candidate column instead keeps the exchange root-free and on-stream.
//   4. global_pick: every rank picks the same winner (largest |c|, lowest global index on ties -- Julia
//      `argmax`, src/matchingpursuit.jl:184) and stores its column in slot nnz of the active-atom cache
//   5. the per-signal update kernel (update.cu) reading atoms from that cache.
//
// Steps 2-4 are ONE kernel when the ranks can map each other's memory (`peer_exchange_kernel`, the default):
// every rank owns a mailbox in device memory, exported with cudaIpcGetMemHandle and mapped by all peers
// (NVLink / NVSwitch peer access).  The kernel reduces the local candidates, stores its record straight
// into slot [parity][rank] of EVERY rank's mailbox, publishes a sequence number (bar.sync, fence.sys, st.release.sys),
// spins (ld.acquire.sys, bounded by a timeout) until all ranks' numbers have arrived in its own mailbox and
// picks the winner -- no NCCL launch, no staging copy, one NVLink one-way latency per iteration.  Two
// parity slots suffice: a rank cannot publish iteration s + 1 before every peer has published s, and a
// peer publishes s only after it has consumed s - 1.  NCCL remains the bootstrap (the handles travel
// through one all-gather) and the transport of last resort (CSB200_SHARD_EXCHANGE=nccl, or no peer access).
// NCCL is dlopen'ed so that single-GPU users carry no dependency on it.
#include "../../include/csb200.h"
#include "common.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

using namespace csb;

// accessors implemented in api.cu
extern "C" {
int csb200_internal_dict_info(const csb200_dict* d, const void** dA, int64_t* M, int64_t* N, int64_t* ld, int* dtype,
                              int* device, int64_t* n_offset, int64_t* n_total);
void csb200_internal_set_error(const char* msg);
}

namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

NcclApi& nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        // An NCCL that the host program has loaded already (e.g. the one PyTorch bundles) is reused; otherwise the
        // system library is loaded with RTLD_LOCAL, so that its symbols never shadow a different NCCL version another
        // library brings along later (a RTLD_GLOBAL 2.27 made a later `import torch` fail on a 2.28-only symbol).
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            api.handle = dlopen(n, RTLD_NOW | RTLD_NOLOAD);
            if (api.handle) break;
        }
        for (const char* n : names) {
            if (api.handle) break;
            api.handle = dlopen(n, RTLD_NOW | RTLD_LOCAL);
        }
        if (!api.handle) return;
        api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.handle, "ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.handle, "ncclCommInitRank");
        api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.handle, "ncclCommDestroy");
        api.AllGather = (decltype(api.AllGather))dlsym(api.handle, "ncclAllGather");
        api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.handle, "ncclGetErrorString");
        api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather && api.GetErrorString;
    });
    return api;
}

int fail_nccl(ncclResult_t r, const char* what) {
    char buf[256];
    snprintf(buf, sizeof buf, "%s: %s", what, nccl().GetErrorString ? nccl().GetErrorString(r) : "nccl error");
    csb200_internal_set_error(buf);
    return CSB200_ERR_NCCL;
}
int fail_cuda(cudaError_t e, const char* what) {
    char buf[256];
    snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(e));
    csb200_internal_set_error(buf);
    cudaGetLastError();
    return e == cudaErrorMemoryAllocation ? CSB200_ERR_OOM : CSB200_ERR_CUDA;
}
#define CU_TRY(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) return fail_cuda(e__, #expr); } while (0)
#define NC_TRY(expr) do { ncclResult_t r__ = (expr); if (r__ != ncclSuccess) return fail_nccl(r__, #expr); } while (0)

// One exchange record: header {|c| as double, global atom index, pad} followed by the atom's column.
constexpr int REC_HDR = 16;

template <typename T>
__global__ void __launch_bounds__(256) local_best_kernel(const double* __restrict__ pval, const int* __restrict__ pidx,
                                                         int P, const T* __restrict__ A, int ld, int idx_offset,
                                                         unsigned char* __restrict__ rec) {
    __shared__ double sv[8];
    __shared__ int si[8];
    __shared__ int s_best;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double bv = -1.0;
    int bi = INT_MAX;
    for (int c = tid; c < P; c += 256) {
        const int i = pidx[c];
        if (i >= 0 && cand_better(pval[c], i, bv, bi)) { bv = pval[c]; bi = i; }
    }
    for (int off = 16; off > 0; off >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, off);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
        if (cand_better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { sv[warp] = bv; si[warp] = bi; }
    __syncthreads();
    if (tid == 0) {
        bv = sv[0]; bi = si[0];
        for (int w = 1; w < 8; ++w) if (cand_better(sv[w], si[w], bv, bi)) { bv = sv[w]; bi = si[w]; }
        *reinterpret_cast<double*>(rec) = bv;
        *reinterpret_cast<int*>(rec + 8) = (bi == INT_MAX) ? -1 : bi;
        *reinterpret_cast<int*>(rec + 12) = 0;
        s_best = (bi == INT_MAX) ? -1 : bi;
    }
    __syncthreads();
    T* col = reinterpret_cast<T*>(rec + REC_HDR);
    if (s_best >= 0) {
        const T* src = A + (size_t)(s_best - idx_offset) * ld;
        for (int row = tid; row < ld; row += 256) col[row] = src[row];
    } else {
        for (int row = tid; row < ld; row += 256) col[row] = (T)0;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) global_pick_kernel(const unsigned char* __restrict__ recs, int nranks,
                                                          size_t rec_bytes, int ld, const int* __restrict__ nnz,
                                                          int kcap, T* __restrict__ Acache, double* __restrict__ cand_val,
                                                          int* __restrict__ cand_idx) {
    __shared__ int s_rank;
    if (threadIdx.x == 0) {
        double bv = -1.0;
        int bi = INT_MAX, br = -1;
        for (int g = 0; g < nranks; ++g) {
            const unsigned char* rec = recs + (size_t)g * rec_bytes;
            const double v = *reinterpret_cast<const double*>(rec);
            const int i = *reinterpret_cast<const int*>(rec + 8);
            if (i >= 0 && cand_better(v, i, bv, bi)) { bv = v; bi = i; br = g; }
        }
        cand_val[0] = bv;
        cand_idx[0] = (br < 0) ? -1 : bi;
        s_rank = br;
    }
    __syncthreads();
    const int t = nnz[0];
    if (s_rank < 0 || t >= kcap) return;
    const T* src = reinterpret_cast<const T*>(recs + (size_t)s_rank * rec_bytes + REC_HDR);
    T* dst = Acache + (size_t)t * ld;
    for (int row = threadIdx.x; row < ld; row += 256) dst[row] = src[row];
}


// ---- peer-memory exchange (see the header comment) ---------------------------------------------------------------
constexpr int MAX_PEERS = 16;
constexpr size_t BOX_ERR_OFF = 128;        // int: a wait timed out
constexpr size_t BOX_DATA_OFF = 256;       // flags: unsigned long long [MAX_PEERS] at offset 0
constexpr int PX_THREADS = 1024;

struct PeerArgs {
    unsigned char* box[MAX_PEERS];         // every rank's mailbox as mapped in this process (box[rank] = own)
    int rank, nranks;
    unsigned long long seq;                // sequence number of this exchange (monotone over the communicator's life)
    unsigned long long slot_bytes;         // capacity of one record slot
    unsigned long long timeout_ns;
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

template <typename T>
__global__ void __launch_bounds__(PX_THREADS) peer_exchange_kernel(PeerArgs pa, const double* __restrict__ pval,
                                                                   const int* __restrict__ pidx, int P,
                                                                   const T* __restrict__ A, int ld, int idx_offset,
                                                                   const int* __restrict__ nnz, int kcap,
                                                                   T* __restrict__ Acache, double* __restrict__ cand_val,
                                                                   int* __restrict__ cand_idx) {
    __shared__ double sv[PX_THREADS / 32];
    __shared__ int si[PX_THREADS / 32];
    __shared__ double s_bv;
    __shared__ int s_best, s_rank;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // (a) local best of the GEMV's per-CTA candidates
    double bv = -1.0;
    int bi = INT_MAX;
    for (int c = tid; c < P; c += PX_THREADS) {
        const int i = pidx[c];
        if (i >= 0 && cand_better(pval[c], i, bv, bi)) { bv = pval[c]; bi = i; }
    }
    for (int off = 16; off > 0; off >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, off);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
        if (cand_better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { sv[warp] = bv; si[warp] = bi; }
    __syncthreads();
    if (tid == 0) {
        bv = sv[0]; bi = si[0];
        for (int w = 1; w < PX_THREADS / 32; ++w) if (cand_better(sv[w], si[w], bv, bi)) { bv = sv[w]; bi = si[w]; }
        s_bv = bv;
        s_best = (bi == INT_MAX) ? -1 : bi;
    }
    __syncthreads();
    // (b) push {|c|, index, column} into slot [parity][rank] of every rank's mailbox (remote stores over NVLink)
    const int best = s_best;
    const size_t slot = BOX_DATA_OFF + ((size_t)(pa.seq & 1) * pa.nranks + pa.rank) * pa.slot_bytes;
    const int nvec = (int)((size_t)ld * sizeof(T) / 16);            // ld is a multiple of 16 elements
    const uint4* src = best >= 0 ? reinterpret_cast<const uint4*>(A + (size_t)(best - idx_offset) * ld) : nullptr;
    for (int v = tid; v < nvec; v += PX_THREADS) {
        const uint4 x = src ? __ldg(src + v) : make_uint4(0u, 0u, 0u, 0u);
        for (int g = 0; g < pa.nranks; ++g)
            *reinterpret_cast<uint4*>(pa.box[g] + slot + REC_HDR + (size_t)v * 16) = x;
    }
    if (tid < pa.nranks) {
        const unsigned long long vb = (unsigned long long)__double_as_longlong(s_bv);
        uint4 h;
        h.x = (unsigned)vb; h.y = (unsigned)(vb >> 32); h.z = (unsigned)best; h.w = 0u;
        *reinterpret_cast<uint4*>(pa.box[tid] + slot) = h;
    }
    __syncthreads();                       // CTA-scope order: every thread's stores precede the publishing threads' fence
    unsigned char* mine = pa.box[pa.rank];
    if (tid < pa.nranks) {
        // (c) publish: my record for exchange `seq` is complete in rank tid's mailbox.  The system-scope fence is
        // cumulative over the stores the barrier ordered before it, so only the <= 16 publishing threads pay for it
        // (all 1024 threads fencing made membar the kernel's top stall: ncu, profiles/c4_exchange_r01f.md).
        __threadfence_system();
        st_release_sys(reinterpret_cast<unsigned long long*>(pa.box[tid]) + pa.rank, pa.seq);
        // (d) wait for rank tid's record in my own mailbox
        const unsigned long long* f = reinterpret_cast<const unsigned long long*>(mine) + tid;
        const unsigned long long t0 = global_timer_ns();
        while (ld_acquire_sys(f) < pa.seq) {
            if (global_timer_ns() - t0 > pa.timeout_ns) { atomicExch(reinterpret_cast<int*>(mine + BOX_ERR_OFF), 1); break; }
        }
    }
    __syncthreads();
    // (e) every rank picks the same winner from identical records (value desc, global index asc)
    const unsigned char* recs = mine + BOX_DATA_OFF + (size_t)(pa.seq & 1) * pa.nranks * pa.slot_bytes;
    if (tid == 0) {
        double wv = -1.0;
        int wi = INT_MAX, wr = -1;
        for (int g = 0; g < pa.nranks; ++g) {
            const uint4 h = __ldcg(reinterpret_cast<const uint4*>(recs + (size_t)g * pa.slot_bytes));
            const double v = __longlong_as_double((long long)(((unsigned long long)h.y << 32) | h.x));
            const int i = (int)h.z;
            if (i >= 0 && cand_better(v, i, wv, wi)) { wv = v; wi = i; wr = g; }
        }
        cand_val[0] = wv;
        cand_idx[0] = (wr < 0) ? -1 : wi;
        s_rank = wr;
    }
    __syncthreads();
    const int t = nnz[0];
    if (s_rank < 0 || t >= kcap) return;
    const uint4* wsrc = reinterpret_cast<const uint4*>(recs + (size_t)s_rank * pa.slot_bytes + REC_HDR);
    uint4* dst = reinterpret_cast<uint4*>(Acache + (size_t)t * ld);
    for (int v = tid; v < nvec; v += PX_THREADS) dst[v] = __ldcg(wsrc + v);      // L1 may hold the slot's previous use
}

struct PeerBox {
    unsigned char* local = nullptr;
    unsigned char* peer[MAX_PEERS] = {};
    size_t slot_bytes = 0, total_bytes = 0;
    unsigned long long seq = 0;
    bool tried = false, ready = false;
};

}  // namespace

struct csb200_comm {
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1, device = 0;
    cudaStream_t stream = nullptr;
    std::mutex mu;
    PeerBox px;
    int last_mode = 0;          // exchange used by the last solve: 0 NCCL all-gather, 1 peer-memory mailboxes
    // per-solve resources kept across solves (allocation, 2-4 events per iteration and their release cost a 128-iteration
    // solve 4-14 % of its wall time when they were created and destroyed inside every call)
    unsigned char* scratch = nullptr;
    size_t scratch_bytes = 0;
    std::vector<cudaEvent_t> ev_pool;
    // device-time breakdown of the last solve (ms): whole solve, correlation passes, exchange, update, gaps; the
    // last three only when CSB200_SHARD_TIMING is set (1: also printed on stderr, 2: silent)
    double last_ms[5] = {0, 0, 0, 0, 0};
    int64_t last_iters = 0;
    unsigned long long* d_agree = nullptr;      // [2 + 2 * nranks]: {seq, ready} send slot + gathered copies
};

namespace {

void peerbox_release(csb200_comm* c) {
    PeerBox& px = c->px;
    for (int g = 0; g < c->nranks && g < MAX_PEERS; ++g)
        if (px.peer[g] && g != c->rank) cudaIpcCloseMemHandle(px.peer[g]);
    if (px.local) cudaFree(px.local);
    for (auto& q : px.peer) q = nullptr;
    px.local = nullptr;
    px.ready = false;
    cudaGetLastError();
}

// Collective (every rank of the communicator calls it from its first sharded solve): allocate the mailbox, swap IPC
// handles through one NCCL all-gather, map the peers, agree through a second all-gather on whether EVERY rank
// succeeded.  Any failure on any rank leaves all ranks on the NCCL exchange -- never a mixed state.
int peerbox_setup(csb200_comm* c, size_t rec_bytes) {
    PeerBox& px = c->px;
    px.tried = true;
    if (c->nranks > MAX_PEERS) return CSB200_OK;
    struct Hello { cudaIpcMemHandle_t h; unsigned long long slot; int ok; int pad; };
    static_assert(sizeof(Hello) % 8 == 0, "Hello layout");
    const int G = c->nranks;
    cudaStream_t st = c->stream;
    size_t slot = rec_bytes > ((size_t)1 << 20) ? rec_bytes : ((size_t)1 << 20);
    slot = (slot + 255) / 256 * 256;
    Hello me;
    memset(&me, 0, sizeof me);
    me.slot = slot;
    px.total_bytes = BOX_DATA_OFF + 2 * (size_t)G * slot;
    bool ok = cudaMalloc(&px.local, px.total_bytes) == cudaSuccess;
    if (!ok) px.local = nullptr;
    ok = ok && cudaMemset(px.local, 0, px.total_bytes) == cudaSuccess;
    ok = ok && cudaDeviceSynchronize() == cudaSuccess;
    ok = ok && (G == 1 || cudaIpcGetMemHandle(&me.h, px.local) == cudaSuccess);
    cudaGetLastError();
    me.ok = ok ? 1 : 0;

    Hello* dbuf = nullptr;                       // [1 + G]: send, recv
    CU_TRY(cudaMalloc(&dbuf, sizeof(Hello) * (size_t)(1 + G)));
    std::vector<Hello> all(G);
    auto gather = [&]() -> int {
        cudaError_t e = cudaMemcpyAsync(dbuf, &me, sizeof me, cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) return fail_cuda(e, "peer exchange setup: upload");
        ncclResult_t nr = nccl().AllGather(dbuf, dbuf + 1, sizeof(Hello), ncclChar, c->comm, st);
        if (nr != ncclSuccess) return fail_nccl(nr, "peer exchange setup: ncclAllGather");
        e = cudaMemcpyAsync(all.data(), dbuf + 1, sizeof(Hello) * (size_t)G, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) return fail_cuda(e, "peer exchange setup: download");
        return CSB200_OK;
    };
    int rc = gather();
    if (rc) { cudaFree(dbuf); peerbox_release(c); return rc; }
    bool all_ok = true;
    for (int g = 0; g < G; ++g) all_ok = all_ok && all[g].ok == 1 && all[g].slot == slot;
    if (all_ok) {
        for (int g = 0; g < G; ++g) {
            if (g == c->rank) { px.peer[g] = px.local; continue; }
            void* ptr = nullptr;
            if (cudaIpcOpenMemHandle(&ptr, all[g].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                all_ok = false;
                break;
            }
            px.peer[g] = static_cast<unsigned char*>(ptr);
        }
    }
    me.ok = all_ok ? 1 : 0;
    rc = gather();                               // also orders every rank's memset before any peer's first store
    cudaFree(dbuf);
    if (rc) { peerbox_release(c); return rc; }
    for (int g = 0; g < G; ++g) all_ok = all_ok && all[g].ok == 1;
    if (!all_ok) { peerbox_release(c); return CSB200_OK; }
    px.slot_bytes = slot;
    px.seq = 0;
    px.ready = true;
    return CSB200_OK;
}

bool want_peer_exchange() {
    const char* e = getenv("CSB200_SHARD_EXCHANGE");        // "nccl" forces the all-gather; must agree on all ranks
    return !(e && (strcmp(e, "nccl") == 0 || strcmp(e, "NCCL") == 0));
}

}  // namespace

extern "C" {

int csb200_comm_unique_id(void* id_bytes) {
    if (!id_bytes) return CSB200_ERR_INVALID_ARG;
    if (!nccl().ok) { csb200_internal_set_error("libnccl.so.2 could not be loaded"); return CSB200_ERR_NCCL; }
    static_assert(sizeof(ncclUniqueId) == CSB200_NCCL_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId id;
    NC_TRY(nccl().GetUniqueId(&id));
    memcpy(id_bytes, &id, sizeof id);
    return CSB200_OK;
}

int csb200_comm_create(const void* id_bytes, int rank, int nranks, int device, csb200_comm** out) {
    if (!id_bytes || !out || nranks < 1 || rank < 0 || rank >= nranks) return CSB200_ERR_INVALID_ARG;
    *out = nullptr;
    if (!nccl().ok) { csb200_internal_set_error("libnccl.so.2 could not be loaded"); return CSB200_ERR_NCCL; }
    CU_TRY(cudaSetDevice(device));
    csb200_comm* c = new (std::nothrow) csb200_comm;
    if (!c) return CSB200_ERR_OOM;
    c->rank = rank; c->nranks = nranks; c->device = device;
    ncclUniqueId id;
    memcpy(&id, id_bytes, sizeof id);
    ncclResult_t r = nccl().CommInitRank(&c->comm, nranks, id, rank);
    if (r != ncclSuccess) { delete c; return fail_nccl(r, "ncclCommInitRank"); }
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { nccl().CommDestroy(c->comm); delete c; return fail_cuda(e, "cudaStreamCreate"); }
    *out = c;
    return CSB200_OK;
}

int csb200_comm_destroy(csb200_comm* c) {
    if (!c) return CSB200_OK;
    cudaSetDevice(c->device);
    if (c->stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
    for (auto e : c->ev_pool) cudaEventDestroy(e);
    if (c->scratch) cudaFree(c->scratch);
    if (c->d_agree) cudaFree(c->d_agree);
    peerbox_release(c);
    if (c->comm && nccl().ok) nccl().CommDestroy(c->comm);
    delete c;
    return CSB200_OK;
}

int csb200_comm_exchange_mode(const csb200_comm* c) { return c ? c->last_mode : CSB200_ERR_INVALID_ARG; }

int csb200_comm_last_timing(csb200_comm* c, double* ms5, int64_t* iters) {
    if (!c || !ms5) return CSB200_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    for (int i = 0; i < 5; ++i) ms5[i] = c->last_ms[i];
    if (iters) *iters = c->last_iters;
    return CSB200_OK;
}

int csb200_omp_sharded(csb200_dict* shard, csb200_comm* c, const void* b, int64_t k, double eps, int64_t* sel_idx,
                       double* coef, int64_t* nnz_out, double* resnorm, int64_t* iters_out, double* corr_ms) {
    if (!shard || !c || !b || k < 0) return CSB200_ERR_INVALID_ARG;
    if (!(eps >= 0)) return CSB200_ERR_NEGATIVE_EPS;
    const void* dA; int64_t M, N, ld, n_offset, n_total; int dtype, device;
    int rc = csb200_internal_dict_info(shard, &dA, &M, &N, &ld, &dtype, &device, &n_offset, &n_total);
    if (rc) return rc;
    if (device != c->device) { csb200_internal_set_error("shard and communicator live on different devices"); return CSB200_ERR_INVALID_ARG; }
    std::lock_guard<std::mutex> lk(c->mu);
    CU_TRY(cudaSetDevice(device));
    const bool f32 = dtype == CSB200_F32;
    const size_t es = f32 ? 4 : 8;
    int64_t kcap = k < M ? k : M;
    if (n_total < kcap) kcap = n_total;
    if (kcap < 1) kcap = 1;
    int num_sms = 148;
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, device);
    const int P = corr_gemv_blocks((int)N, (int)ld, dtype == CSB200_F32, 1, num_sms);
    const size_t rec_bytes = (REC_HDR + (size_t)ld * es + 15) / 16 * 16;
    cudaStream_t st = c->stream;
    bool peer = want_peer_exchange();
    if (peer && !c->px.tried) { rc = peerbox_setup(c, rec_bytes); if (rc) return rc; }
    peer = peer && c->px.ready && rec_bytes <= c->px.slot_bytes;
    // The mailbox sequence number is per-rank host state: a rank that left an earlier solve early (error, timeout)
    // would otherwise run behind its peers for good and read stale slots.  Every solve therefore starts by agreeing
    // on {max sequence number, "all ranks take the peer path"} through one 16-byte-per-rank all-gather.
    if (c->nranks > 1) {
        if (!c->d_agree) CU_TRY(cudaMalloc(&c->d_agree, sizeof(unsigned long long) * (size_t)(2 + 2 * c->nranks)));
        unsigned long long me[2] = {c->px.seq, peer ? 1ull : 0ull};
        std::vector<unsigned long long> all(2 * (size_t)c->nranks);
        CU_TRY(cudaMemcpyAsync(c->d_agree, me, sizeof me, cudaMemcpyHostToDevice, st));
        NC_TRY(nccl().AllGather(c->d_agree, c->d_agree + 2, 2 * sizeof(unsigned long long), ncclChar, c->comm, st));
        CU_TRY(cudaMemcpyAsync(all.data(), c->d_agree + 2, all.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaStreamSynchronize(st));
        for (int g = 0; g < c->nranks; ++g) {
            if (all[2 * g] > c->px.seq) c->px.seq = all[2 * g];
            peer = peer && all[2 * g + 1] == 1ull;
        }
    }
    c->last_mode = peer ? 1 : 0;
    PeerArgs pa;
    memset(&pa, 0, sizeof pa);
    if (peer) {
        for (int g = 0; g < c->nranks; ++g) pa.box[g] = c->px.peer[g];
        pa.rank = c->rank; pa.nranks = c->nranks; pa.slot_bytes = c->px.slot_bytes;
        const char* te = getenv("CSB200_PEER_TIMEOUT_S");
        const double ts = te ? atof(te) : 60.0;
        pa.timeout_ns = (unsigned long long)((ts > 0 ? ts : 60.0) * 1e9);
    }

    // device scratch (one allocation)
    struct Off { size_t b, r, acache, pval, pidx, cval, cidx, nnz, sel, T, z, x, res, it, done, flags, send, recv, end; } o;
    size_t p = 0;
    auto take = [&](size_t bytes) { size_t at = p; p += (bytes + 255) / 256 * 256; return at; };
    o.b = take(ld * es); o.r = take(ld * es); o.acache = take((size_t)ld * kcap * es);
    o.pval = take((size_t)P * 8); o.pidx = take((size_t)P * 4); o.cval = take(8); o.cidx = take(4);
    o.nnz = take(4); o.sel = take(kcap * 4); o.T = take((size_t)kcap * kcap * 8); o.z = take(kcap * 8); o.x = take(kcap * 8);
    o.res = take(8); o.it = take(4); o.done = take(4); o.flags = take(4);
    o.send = take(rec_bytes); o.recv = take(rec_bytes * c->nranks); o.end = p;
    if (o.end > c->scratch_bytes) {                              // grow-only scratch owned by the communicator
        if (c->scratch) { cudaStreamSynchronize(st); cudaFree(c->scratch); c->scratch = nullptr; c->scratch_bytes = 0; }
        CU_TRY(cudaMalloc(&c->scratch, o.end));
        c->scratch_bytes = o.end;
    }
    unsigned char* base = c->scratch;
    const char* tenv = getenv("CSB200_SHARD_TIMING");           // per-phase device times: 1 = also on stderr, 2 = silent
    const bool timing = tenv && (tenv[0] == '1' || tenv[0] == '2');
    const bool timing_print = tenv && tenv[0] == '1';
    const size_t nev = (timing ? 4 : 2) * (size_t)k + 1;        // + one event after the last update
    while (c->ev_pool.size() < nev) {
        cudaEvent_t e = nullptr;
        CU_TRY(cudaEventCreate(&e));
        c->ev_pool.push_back(e);
    }
    cudaEvent_t* ev = c->ev_pool.data();
    int status = CSB200_OK;

    do {
        cudaError_t e = cudaMemsetAsync(base, 0, o.end, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(base + o.b, b, (size_t)M * es, cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) { status = fail_cuda(e, "upload signal"); break; }

        CorrArgs ca;
        ca.A = dA; ca.R = base + o.r; ca.M = (int)M; ca.ld = (int)ld; ca.N = (int)N; ca.nsig = 1; ca.S = 1; ca.P = P;
        ca.idx_offset = (int)n_offset; ca.pval = (double*)(base + o.pval); ca.pidx = (int*)(base + o.pidx);
        StateArgs sa;
        sa.A = dA; sa.B = base + o.b; sa.R = base + o.r; sa.M = (int)M; sa.ld = (int)ld; sa.N = (int)N; sa.nsig = 1;
        sa.kcap = (int)kcap; sa.S = 1; sa.P = 1; sa.take = 1; sa.idx_offset = (int)n_offset; sa.ignore_done = 0; sa.eps = eps;
        sa.pval = (double*)(base + o.cval); sa.pidx = (int*)(base + o.cidx);
        sa.nnz = (int*)(base + o.nnz); sa.sel = (int*)(base + o.sel); sa.Rf = (double*)(base + o.T);
        sa.z = (double*)(base + o.z); sa.x = (double*)(base + o.x); sa.resnorm = (double*)(base + o.res);
        sa.iters = (int*)(base + o.it); sa.done = (int*)(base + o.done); sa.flags = (int*)(base + o.flags);
        sa.gram = nullptr;

        int* dflag = (int*)(base + o.flags);
        e = launch_nonfinite_check(base + o.b, (size_t)ld, f32, dflag, st);
        int hflag = 0;
        if (e == cudaSuccess) e = cudaMemcpyAsync(&hflag, dflag, 4, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { status = fail_cuda(e, "nonfinite check"); break; }
        if (hflag) { status = CSB200_ERR_NONFINITE_INPUT; break; }
        e = launch_reset_state(sa, f32, st);
        if (e != cudaSuccess) { status = fail_cuda(e, "reset_state"); break; }

        for (int64_t it = 0; it < k && status == CSB200_OK; ++it) {
            cudaEventRecord(ev[2 * it], st);
            e = launch_corr_gemv(ca, f32, st);
            cudaEventRecord(ev[2 * it + 1], st);
            if (e != cudaSuccess) { status = fail_cuda(e, "corr_gemv"); break; }
            if (peer) {
                pa.seq = ++c->px.seq;
                if (f32) peer_exchange_kernel<float><<<1, PX_THREADS, 0, st>>>(pa, ca.pval, ca.pidx, P, (const float*)dA, (int)ld, (int)n_offset, sa.nnz, (int)kcap, (float*)(base + o.acache), (double*)(base + o.cval), (int*)(base + o.cidx));
                else peer_exchange_kernel<double><<<1, PX_THREADS, 0, st>>>(pa, ca.pval, ca.pidx, P, (const double*)dA, (int)ld, (int)n_offset, sa.nnz, (int)kcap, (double*)(base + o.acache), (double*)(base + o.cval), (int*)(base + o.cidx));
                e = cudaGetLastError();
                if (e != cudaSuccess) { status = fail_cuda(e, "peer_exchange"); break; }
            } else {
                if (f32) local_best_kernel<float><<<1, 256, 0, st>>>(ca.pval, ca.pidx, P, (const float*)dA, (int)ld, (int)n_offset, base + o.send);
                else local_best_kernel<double><<<1, 256, 0, st>>>(ca.pval, ca.pidx, P, (const double*)dA, (int)ld, (int)n_offset, base + o.send);
                ncclResult_t nr = nccl().AllGather(base + o.send, base + o.recv, rec_bytes, ncclChar, c->comm, st);
                if (nr != ncclSuccess) { status = fail_nccl(nr, "ncclAllGather"); break; }
                if (f32) global_pick_kernel<float><<<1, 256, 0, st>>>(base + o.recv, c->nranks, rec_bytes, (int)ld, sa.nnz, (int)kcap, (float*)(base + o.acache), (double*)(base + o.cval), (int*)(base + o.cidx));
                else global_pick_kernel<double><<<1, 256, 0, st>>>(base + o.recv, c->nranks, rec_bytes, (int)ld, sa.nnz, (int)kcap, (double*)(base + o.acache), (double*)(base + o.cval), (int*)(base + o.cidx));
            }
            if (timing) cudaEventRecord(ev[2 * k + 2 * it], st);
            e = launch_omp_update_cluster(sa, f32, st, base + o.acache);
            if (timing) cudaEventRecord(ev[2 * k + 2 * it + 1], st);
            if (e != cudaSuccess) { status = fail_cuda(e, "omp_update"); break; }
        }
        if (status) break;
        cudaEventRecord(ev[nev - 1], st);
        std::vector<int> hsel(kcap);
        std::vector<double> hx(kcap);
        int hn = 0, hit = 0;
        double hres = 0;
        e = cudaMemcpyAsync(hsel.data(), base + o.sel, kcap * 4, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(hx.data(), base + o.x, kcap * 8, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(&hn, base + o.nnz, 4, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(&hit, base + o.it, 4, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(&hres, base + o.res, 8, cudaMemcpyDeviceToHost, st);
        int perr = 0;
        if (peer && e == cudaSuccess) e = cudaMemcpyAsync(&perr, c->px.local + BOX_ERR_OFF, 4, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { status = fail_cuda(e, "download"); break; }
        if (perr) {
            cudaMemsetAsync(c->px.local + BOX_ERR_OFF, 0, 4, st);
            csb200_internal_set_error("peer-memory exchange timed out waiting for a rank (CSB200_PEER_TIMEOUT_S)");
            status = CSB200_ERR_NCCL;
            break;
        }
        for (int64_t j = 0; j < k; ++j) {
            if (sel_idx) sel_idx[j] = j < hn ? hsel[j] : -1;
            if (coef) coef[j] = j < hn ? hx[j] : 0.0;
        }
        if (nnz_out) *nnz_out = hn;
        if (iters_out) *iters_out = hit;
        if (resnorm) *resnorm = hres;
        {
            double tot = 0;
            for (int64_t it = 0; it < k; ++it) { float ms = 0; cudaEventElapsedTime(&ms, ev[2 * it], ev[2 * it + 1]); tot += ms; }
            if (corr_ms) *corr_ms = tot;
            c->last_ms[0] = c->last_ms[2] = c->last_ms[3] = c->last_ms[4] = 0.0;
            c->last_ms[1] = tot;
            c->last_iters = k;
            if (k > 0) {
                float ms = 0;
                cudaEventElapsedTime(&ms, ev[0], ev[nev - 1]);
                c->last_ms[0] = ms;         // first correlation pass enqueued -> last update finished
            }
        }
        if (timing && k > 0) {
            double gemv = 0, exch = 0, upd = 0, gap = 0;
            for (int64_t it = 0; it < k; ++it) {
                float ms = 0;
                cudaEventElapsedTime(&ms, ev[2 * it], ev[2 * it + 1]); gemv += ms;
                cudaEventElapsedTime(&ms, ev[2 * it + 1], ev[2 * k + 2 * it]); exch += ms;
                cudaEventElapsedTime(&ms, ev[2 * k + 2 * it], ev[2 * k + 2 * it + 1]); upd += ms;
                if (it + 1 < k) { cudaEventElapsedTime(&ms, ev[2 * k + 2 * it + 1], ev[2 * (it + 1)]); gap += ms; }
            }
            c->last_ms[2] = exch; c->last_ms[3] = upd; c->last_ms[4] = gap;
            if (timing_print)
                fprintf(stderr, "[csb200 shard rank %d/%d] per iteration (us): gemv %.1f  exchange[%s] %.1f  update %.1f  gap %.1f\n",
                        c->rank, c->nranks, 1e3 * gemv / k, peer ? "peer-memory" : "local-best+allgather+pick", 1e3 * exch / k, 1e3 * upd / k, 1e3 * gap / k);
        }
    } while (0);
    cudaStreamSynchronize(st);
    return status;
}

}  // extern "C"
