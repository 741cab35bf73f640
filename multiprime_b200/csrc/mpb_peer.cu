// mpb_peer.cu — small-vector all-reduce over NVLink peer memory for the sequence-sharded walk.
//
// A walk round of a sharded run is  scan (local sequences) -> sum of the count vector over the shards -> advance.  The
// vector is small (4 counters per candidate, a few hundred candidates) and there are a dozen dependent rounds per window
// batch: what matters is latency, not bandwidth.  Going through a collective library costs a host call and a kernel of
// its own per round; here the exchange is ONE single-block kernel on the scan's stream:
//     push   every rank stores its vector into slot [parity][rank] of every peer's receive buffer (posted NVLink writes)
//     signal fence, then a release store of the round's sequence number into flag[rank] of every peer
//     wait   acquire loads on the own flags until every peer's number has arrived
//     sum    the own receive slots (local HBM), in rank order, back into the vector
// Two parities of receive slots make one barrier per round enough: a rank can only be one round ahead of the slowest
// peer (it needs that peer's signal to pass), so the slot it overwrites two rounds later has been summed by everybody.
// The number of elements is read from device memory (the walk's candidate count) and is the same on every rank because
// the walk is deterministic on the summed counts; rounds enqueued past the end of the walk have zero candidates and
// return at once on every rank — no rank ever waits for a peer that is not coming.  A wait still gives up after
// PEER_TIMEOUT_NS (a dead peer must not hang the GPU), raises the walk's error flag and later rounds return at once.
//
// The receive buffers are plain cudaMalloc memory, opened in the peers through CUDA IPC (one process per GPU) or used by
// address (shards on threads of one process: the tests).
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include <unistd.h>

#include "mpb200.h"
#include "mpb_host.h"
#include "mpb_peer.h"

#define fail mpb_fail
#define CK MPB_CK

#define PEER_FLAG_WORDS 512  // flags live in the first 4 KB of a buffer (one 64-bit word per source rank, padded)
#define PEER_TIMEOUT_NS 4000000000ull
#define PEER_MAGIC 0x6d70625f70656572ull

struct mpb_peer {
    mpb_ctx* ctx;
    int rank, world;
    int64_t cap;                    // elements per receive slot
    unsigned long long* local;      // own buffer: flags + [2][world][cap]
    unsigned long long* base[MPB_PEER_MAX_WORLD];
    bool opened[MPB_PEER_MAX_WORLD];  // base[p] came from cudaIpcOpenMemHandle
    unsigned long long* seq;        // device: rounds done
    bool connected;
};

struct PeerHandle {  // 128 bytes
    uint64_t magic;
    int64_t pid;
    int32_t device, rank;
    uint64_t ptr;
    int64_t cap;
    cudaIpcMemHandle_t ipc;  // 64 bytes
    uint64_t pad[3];
};
static_assert(sizeof(PeerHandle) == MPB_PEER_HANDLE_BYTES, "handle layout");

struct PeerDev {
    unsigned long long* base[MPB_PEER_MAX_WORLD];
    unsigned long long* seq;
    long long cap;
    int rank, world;
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

#define PEER_THREADS 1024
__global__ void __launch_bounds__(PEER_THREADS)
k_peer_allreduce(PeerDev pd, unsigned long long* __restrict__ data, const int* __restrict__ n_items, int mult,
                 int phases, int* __restrict__ err) {
    __shared__ unsigned long long s_seq;
    const long long n = (long long)n_items[0] * mult;
    if (n <= 0) return;  // the same on every rank
    if (*((volatile int*)err) & MPB_ERR_PEER_TIMEOUT) return;  // a peer is gone: do not wait for it round after round
    if (n > pd.cap) {
        if (threadIdx.x == 0) atomicOr(err, MPB_ERR_PEER_CAP);
        return;
    }
    // phases: 1 = push + signal, 2 = wait + sum, 3 = both (a round).  The split form lets ONE stream play all the ranks
    // of a group in turn (every rank's phase 1, then every rank's phase 2): the single-GPU test of this kernel.
    if (threadIdx.x == 0) {
        s_seq = pd.seq[0] + ((phases & 1) ? 1ull : 0ull);
        pd.seq[0] = s_seq;
    }
    __syncthreads();
    const unsigned long long seq = s_seq;
    const long long par = (long long)(seq & 1ull);
    if (phases & 1) {
        for (int p = 0; p < pd.world; ++p) {
            unsigned long long* dst = pd.base[p] + PEER_FLAG_WORDS + (par * pd.world + pd.rank) * pd.cap;
            for (long long i = threadIdx.x; i < n; i += PEER_THREADS) dst[i] = data[i];
        }
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x < pd.world) st_release_sys(pd.base[threadIdx.x] + pd.rank, seq);
    }
    if (!(phases & 2)) return;
    if (threadIdx.x < pd.world) {
        const unsigned long long* mine = pd.base[pd.rank] + threadIdx.x;
        const unsigned long long t0 = timer_ns();
        unsigned spin = 0;
        while (ld_acquire_sys(mine) < seq) {
            if ((++spin & 0xFFu) == 0 && timer_ns() - t0 > PEER_TIMEOUT_NS) {
                atomicOr(err, MPB_ERR_PEER_TIMEOUT);
                break;
            }
        }
    }
    __syncthreads();
    const unsigned long long* src = pd.base[pd.rank] + PEER_FLAG_WORDS + par * pd.world * pd.cap;
    for (long long i = threadIdx.x; i < n; i += PEER_THREADS) {
        unsigned long long s = 0;
        for (int p = 0; p < pd.world; ++p) s += __ldcv(src + p * pd.cap + i);
        data[i] = s;
    }
}

static size_t peer_bytes(int world, int64_t cap) { return ((size_t)PEER_FLAG_WORDS + (size_t)2 * world * cap) * 8; }

extern "C" int mpb_peer_create(mpb_ctx* ctx, int rank, int world, int64_t cap_elems, mpb_peer** out) {
    if (!ctx || !out) return fail(MPB_EINVAL, "NULL argument");
    if (world < 1 || world > MPB_PEER_MAX_WORLD || rank < 0 || rank >= world)
        return fail(MPB_EINVAL, "peer group: rank %d of %d (at most %d ranks)", rank, world, MPB_PEER_MAX_WORLD);
    if (cap_elems < 1) return fail(MPB_EINVAL, "cap_elems");
    CK(cudaSetDevice(ctx->device));
    mpb_peer* p = new mpb_peer();
    memset(p, 0, sizeof *p);
    p->ctx = ctx;
    p->rank = rank;
    p->world = world;
    p->cap = cap_elems;
    cudaError_t e = cudaMalloc(&p->local, peer_bytes(world, cap_elems));
    if (e == cudaSuccess) e = cudaMalloc(&p->seq, 8);
    if (e == cudaSuccess) e = cudaMemset(p->local, 0, PEER_FLAG_WORDS * 8);
    if (e == cudaSuccess) e = cudaMemset(p->seq, 0, 8);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        mpb_peer_free(p);
        return fail(MPB_ENOMEM, "peer buffer of %zu bytes: %s", peer_bytes(world, cap_elems), cudaGetErrorString(e));
    }
    p->base[rank] = p->local;
    *out = p;
    return 0;
}

extern "C" int mpb_peer_handle(mpb_peer* p, void* handle_out) {
    if (!p || !handle_out) return fail(MPB_EINVAL, "NULL argument");
    CK(cudaSetDevice(p->ctx->device));
    PeerHandle h;
    memset(&h, 0, sizeof h);
    h.magic = PEER_MAGIC;
    h.pid = (int64_t)getpid();
    h.device = p->ctx->device;
    h.rank = p->rank;
    h.ptr = (uint64_t)(uintptr_t)p->local;
    h.cap = p->cap;
    CK(cudaIpcGetMemHandle(&h.ipc, p->local));
    memcpy(handle_out, &h, sizeof h);
    return 0;
}

extern "C" int mpb_peer_connect(mpb_peer* p, const void* handles) {
    if (!p || !handles) return fail(MPB_EINVAL, "NULL argument");
    if (p->connected) return fail(MPB_EINVAL, "peer group already connected");
    CK(cudaSetDevice(p->ctx->device));
    const PeerHandle* hs = (const PeerHandle*)handles;
    for (int r = 0; r < p->world; ++r) {
        const PeerHandle& h = hs[r];
        if (h.magic != PEER_MAGIC || h.rank != r || h.cap != p->cap)
            return fail(MPB_EINVAL, "peer handle %d is malformed (rank %d, cap %lld)", r, h.rank, (long long)h.cap);
        if (r == p->rank) continue;
        if (h.pid == (int64_t)getpid()) {  // shards on threads of one process: the address is valid as it is
            if (h.device != p->ctx->device) {
                int can = 0;
                CK(cudaDeviceCanAccessPeer(&can, p->ctx->device, h.device));
                if (!can) return fail(MPB_ECUDA, "device %d cannot access device %d", p->ctx->device, h.device);
                const cudaError_t e = cudaDeviceEnablePeerAccess(h.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                    return fail(MPB_ECUDA, "peer access %d -> %d: %s", p->ctx->device, h.device, cudaGetErrorString(e));
                cudaGetLastError();
            }
            p->base[r] = (unsigned long long*)(uintptr_t)h.ptr;
        } else {
            void* q = nullptr;
            const cudaError_t e = cudaIpcOpenMemHandle(&q, h.ipc, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) return fail(MPB_ECUDA, "cudaIpcOpenMemHandle (rank %d): %s", r, cudaGetErrorString(e));
            p->base[r] = (unsigned long long*)q;
            p->opened[r] = true;
        }
    }
    p->connected = true;
    return 0;
}

extern "C" int64_t mpb_peer_cap(mpb_peer* p) { return p ? p->cap : 0; }

static int peer_launch(mpb_peer* p, unsigned long long* data_d, const int* n_items_d, int mult, int phases, int* err_d) {
    if (!p->connected) return fail(MPB_EINVAL, "peer group not connected");
    mpb_ctx* ctx = p->ctx;
    PeerDev pd;
    for (int r = 0; r < MPB_PEER_MAX_WORLD; ++r) pd.base[r] = r < p->world ? p->base[r] : nullptr;
    pd.seq = p->seq;
    pd.cap = p->cap;
    pd.rank = p->rank;
    pd.world = p->world;
    MPB_LAUNCH(ctx, k_peer_allreduce, 1, PEER_THREADS, 0, pd, data_d, n_items_d, mult, phases, err_d);
    return 0;
}

int mpb_peer_allreduce_launch(mpb_peer* p, unsigned long long* data_d, const int* n_items_d, int mult, int* err_d) {
    return peer_launch(p, data_d, n_items_d, mult, 3, err_d);
}

// Stand-alone form: sum over the ranks of data[0 .. n) (int64, device memory, n <= capacity), in place, on the
// context's stream.  Every rank of the group must make the same calls in the same order.
extern "C" int mpb_peer_allreduce(mpb_peer* p, int64_t* data_dev, int64_t n) { return mpb_peer_allreduce_phases(p, data_dev, n, 3); }

// phases 1 (push + signal) and 2 (wait + sum) of a round as separate calls: see k_peer_allreduce
extern "C" int mpb_peer_allreduce_phases(mpb_peer* p, int64_t* data_dev, int64_t n, int phases) {
    if (!p || !data_dev) return fail(MPB_EINVAL, "NULL argument");
    if (phases < 1 || phases > 3) return fail(MPB_EINVAL, "phases");
    if (n < 1 || n > p->cap) return fail(MPB_EINVAL, "n=%lld outside 1..%lld", (long long)n, (long long)p->cap);
    mpb_ctx* ctx = p->ctx;
    CK(cudaSetDevice(ctx->device));
    int* scratch = nullptr;  // [0] = n, [1] = error flags
    CK(cudaMallocAsync(&scratch, 8, ctx->stream));
    const int init[2] = {(int)n, 0};
    CK(cudaMemcpyAsync(scratch, init, 8, cudaMemcpyHostToDevice, ctx->stream));
    int rc = peer_launch(p, (unsigned long long*)data_dev, scratch, 1, phases, scratch + 1);
    int got[2] = {0, 0};
    if (!rc) {
        CK(cudaMemcpyAsync(got, scratch, 8, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    CK(cudaFreeAsync(scratch, ctx->stream));
    if (rc) return rc;
    if (got[1] & MPB_ERR_PEER_TIMEOUT) return fail(MPB_ECUDA, "peer all-reduce: a peer did not arrive");
    if (got[1]) return fail(MPB_ECUDA, "peer all-reduce: error flags %d", got[1]);
    return 0;
}

extern "C" void mpb_peer_free(mpb_peer* p) {
    if (!p) return;
    cudaSetDevice(p->ctx->device);
    cudaStreamSynchronize(p->ctx->stream);
    for (int r = 0; r < p->world; ++r)
        if (p->opened[r] && p->base[r]) cudaIpcCloseMemHandle(p->base[r]);
    if (p->local) cudaFree(p->local);
    if (p->seq) cudaFree(p->seq);
    delete p;
}
