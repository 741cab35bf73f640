// mpb_host.h — host-side plumbing shared by the translation units of libmpb200 (not part of the public ABI)
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "mpb200.h"

int mpb_fail(int code, const char* fmt, ...);

#define MPB_CK(call)                                                                                   \
    do {                                                                                               \
        cudaError_t e__ = (call);                                                                      \
        if (e__ != cudaSuccess)                                                                        \
            return mpb_fail(e__ == cudaErrorMemoryAllocation ? MPB_ENOMEM : MPB_ECUDA, "%s:%d %s: %s", \
                            __FILE__, __LINE__, #call, cudaGetErrorString(e__));                       \
    } while (0)

struct ProfRec {
    const char* name;
    cudaEvent_t e0, e1;
    double units;  // algorithmic work units of this launch (kernel specific; evals for the scan kernels)
};
#define MPB_CTX_PINNED_INTS 256
struct mpb_ctx {
    int device;
    cudaStream_t stream;
    cudaStream_t copy_stream;  // H2D chunks of mpb_msa_upload overlap the plane build on `stream`
    int* pinned;               // small pinned scratch (MPB_CTX_PINNED_INTS ints): device -> host flags without a sync
    int64_t launches;
    int sm_count;
    bool profile;
    std::vector<ProfRec> recs;
    std::map<std::string, double> extra_units;  // units attributed to a kernel name after the fact (device-side counts)
    double pending_units;
};
static inline int mpb_ctx_device(mpb_ctx* c) { return c->device; }
static inline cudaStream_t mpb_ctx_stream(mpb_ctx* c) { return c->stream; }
static inline int mpb_ctx_sms(mpb_ctx* c) { return c->sm_count; }

// every kernel goes through MPB_LAUNCH: counted, and (when profiling is on) bracketed by CUDA events on the stream
#define MPB_LAUNCH_NAMED(ctx, name, kern, grid, block, smem, ...)                 \
    do {                                                                          \
        ProfRec pr__ = {name, nullptr, nullptr, (ctx)->pending_units};            \
        if ((ctx)->profile) {                                                     \
            MPB_CK(cudaEventCreate(&pr__.e0));                                    \
            MPB_CK(cudaEventCreate(&pr__.e1));                                    \
            MPB_CK(cudaEventRecord(pr__.e0, (ctx)->stream));                      \
        }                                                                         \
        kern<<<grid, block, smem, (ctx)->stream>>>(__VA_ARGS__);                  \
        (ctx)->launches++;                                                        \
        MPB_CK(cudaGetLastError());                                               \
        if ((ctx)->profile) {                                                     \
            MPB_CK(cudaEventRecord(pr__.e1, (ctx)->stream));                      \
            (ctx)->recs.push_back(pr__);                                          \
        }                                                                         \
        (ctx)->pending_units = 0;                                                 \
    } while (0)
#define MPB_LAUNCH(ctx, kern, grid, block, smem, ...) MPB_LAUNCH_NAMED(ctx, #kern, kern, grid, block, smem, __VA_ARGS__)

// ---- handles ---------------------------------------------------------------------------------------------------
struct mpb_msa {
    mpb_ctx* ctx;
    int64_t n_seq, nsp, n_col;
    int ncw;            // column words incl. the trailing zero word
    uint32_t* planes;   // row view   [ncw][nsp] uint4{A,C,G,T}: a window of one sequence = two 128-bit loads
    uint32_t* colp;     // column view [(ncw-1)*32*4 + 2][nwords]: row (col*4 + base) holds, per 32-sequence word, the
                        // bit "sequence s has base b at column col"; the last two rows are all-ones / all-zeros
    int64_t nwords;     // nsp / 32
    uint8_t* cons;      // [(ncw-1)*32] a frequent base (0..3) of every column, from a sample of the rows: the reference
                        // k-mer of the column-domain window passes (any choice is valid, a good one saves work)
    int32_t* lens;      // [nsp]
    bool short_rows;    // some row ends before the alignment does (unaligned input)
    int* err;           // device error flags
    int64_t row0;       // global index of local sequence 0 (sequence-sharded runs)
};
#define MPB_COLP_ONES(m) ((uint32_t)(((m)->ncw - 1) * 32 * 4))
#define MPB_COLP_ZEROS(m) ((uint32_t)(((m)->ncw - 1) * 32 * 4 + 1))

struct mpb_hist {
    mpb_msa* msa;
    int k, v, nw, log2cap;
    uint64_t* keys;   // [nw][cap]
    uint32_t* cnt;    // [nw][cap]
    uint64_t* first;  // [nw][cap]
    uint32_t* elist;  // [nw][cap] slots in claim order: the table readers walk the occupied slots only
    int32_t* win_pos; // device copy
    std::vector<int32_t> h_win_pos;
    unsigned long long* gap_n;        // [nw]
    unsigned long long* iupac_gap_n;  // [nw]
    unsigned long long* n_entries;    // [nw] distinct table entries
    int32_t* exc;                     // [2*exc_max]
    unsigned long long* exc_n;
    int64_t exc_max;
    // per (window, 32-sequence word) row classes and the patched windows of the special rows, by-products of the table
    // build that the column scan (mpb_cscan.cu) needs: a plain row's window is the column cut, a special row's is not
    // (terminal-gap patching, IUPAC cells, ragged end)
    uint32_t* spec_bits;  // [nw][nwords] special rows (and padding rows): not evaluated by the column kernel
    uint32_t* gap_bits;   // [nw][nwords] gap rows (more than v gaps after patching), plain or special
    uint4* spec_win;      // [nw][spec_cap] patched planes (A,C,G,T) of the special rows that are not gap rows
    int32_t* spec_row;    // [nw][spec_cap] their local sequence index
    unsigned long long* spec_n;  // [nw] (may exceed spec_cap: then the build is repeated with a larger capacity)
    int64_t spec_cap;
    // mpb_hist_summary results, kept on the device for the walk
    unsigned long long* freq;  // [nw][4][k]
    unsigned long long* nn;    // [nw][k-1][16]
    bool have_summary;
};

static inline bool mpb_is_device_ptr(const void* p) {
    if (!p) return false;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

// input that may live on host or device: dev() is a device pointer valid on the ctx stream
struct InBuf {
    mpb_ctx* ctx;
    void* tmp = nullptr;
    const void* d = nullptr;
    int rc = 0;
    InBuf(mpb_ctx* c, const void* hd, size_t bytes) : ctx(c) {
        if (!hd || bytes == 0) return;
        if (mpb_is_device_ptr(hd)) {
            d = hd;
            return;
        }
        cudaError_t e = cudaMallocAsync(&tmp, bytes, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(tmp, hd, bytes, cudaMemcpyHostToDevice, ctx->stream);
        if (e != cudaSuccess) rc = mpb_fail(MPB_ECUDA, "staging input: %s", cudaGetErrorString(e));
        d = tmp;
    }
    ~InBuf() {
        if (tmp) cudaFreeAsync(tmp, ctx->stream);
    }
    template <class T>
    const T* dev() const {
        return (const T*)d;
    }
};

// output that may live on host or device; finish() copies back (async) — caller syncs when any output is host
struct OutBuf {
    mpb_ctx* ctx;
    void* tmp = nullptr;
    void* d = nullptr;
    void* host = nullptr;
    size_t bytes;
    int rc = 0;
    OutBuf(mpb_ctx* c, void* hd, size_t nbytes) : ctx(c), bytes(nbytes) {
        if (!hd || nbytes == 0) return;
        if (mpb_is_device_ptr(hd)) {
            d = hd;
            return;
        }
        host = hd;
        cudaError_t e = cudaMallocAsync(&tmp, nbytes, ctx->stream);
        if (e != cudaSuccess) rc = mpb_fail(MPB_ENOMEM, "staging output: %s", cudaGetErrorString(e));
        d = tmp;
    }
    ~OutBuf() {
        if (tmp) cudaFreeAsync(tmp, ctx->stream);
    }
    template <class T>
    T* dev() const {
        return (T*)d;
    }
    bool is_host() const { return host != nullptr; }
    cudaError_t finish() {
        if (host) return cudaMemcpyAsync(host, tmp, bytes, cudaMemcpyDeviceToHost, ctx->stream);
        return cudaSuccess;
    }
};

int mpb_check_flags(mpb_ctx* ctx, int* dflags);
