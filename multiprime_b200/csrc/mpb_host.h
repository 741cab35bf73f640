// mpb_host.h — host-side plumbing shared by the translation units of libmpb200 (not part of the public ABI)
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

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
    double units;  // algorithmic work units of this launch (kernel specific; evals for k_scan)
};
struct mpb_ctx {
    int device;
    cudaStream_t stream;
    int64_t launches;
    int sm_count;
    bool profile;
    std::vector<ProfRec> recs;
    double pending_units;
};
static inline int mpb_ctx_device(mpb_ctx* c) { return c->device; }
static inline cudaStream_t mpb_ctx_stream(mpb_ctx* c) { return c->stream; }
static inline int mpb_ctx_sms(mpb_ctx* c) { return c->sm_count; }

// every kernel goes through MPB_LAUNCH: counted, and (when profiling is on) bracketed by CUDA events on the stream
#define MPB_LAUNCH(ctx, kern, grid, block, smem, ...)                             \
    do {                                                                          \
        ProfRec pr__ = {#kern, nullptr, nullptr, (ctx)->pending_units};           \
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
