// mpb_walk_dev.cu — the refinement walk of a window batch resident on the device.
//
// The reference's coverage_stast loop (core:860-920) alternates a data-dependent refinement step with a scan of every
// sequence (mis_primer_check).  Round 1 drove it from the host: 11 dependent launch -> sync -> D2H -> Python callback
// round trips per window batch.  Here the tracks (mpb_walk_core.h: seeds, NN arrays, counters) live in HBM and a
// round is a chain of kernels on one stream:
//     k_walk_advance   one thread per track: take the previous round's counts, advance (mpb_walk_consume), stage the
//                      next candidates (mpb_walk_emit)
//     k_walk_compact   one block: prefix sum over the tracks -> compact candidate list in track order (deterministic:
//                      the shards of a sequence-sharded run must agree on it element by element)
//     k_cscan_plan / k_cscan / k_cscan_special    (mpb_cscan.cu) read the candidate count from device memory
// The host only enqueues; it learns the number of live tracks from a pinned word written by an asynchronous copy and
// stops enqueueing when it reads zero (rounds enqueued past the end are no-ops: zero candidates).
// Sequence-sharded runs sum the count vector over the shards between scan and advance: one single-block kernel over
// NVLink peer memory on the same stream (mpb_peer.cu, mpb_walk_dev_set_peer), or the caller's own all-reduce.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <vector>

#include "mpb200.h"
#include "mpb_host.h"
#include "mpb_cscan.h"
#include "mpb_cscan_plan.cuh"
#include "mpb_peer.h"
#include "mpb_walk_core.h"

#define fail mpb_fail
#define CK MPB_CK
#define LAUNCH MPB_LAUNCH

struct mpb_walk_dev {
    mpb_hist* h;
    int k, dnum, degeneracy, n_win, max_cands, rounds_enqueued, max_rounds;
    uint32_t fmask, rmask;
    mpb_track* tracks;      // [2 * n_win]
    int32_t* ntracks;       // [n_win]
    int64_t* cover;         // [n_win]
    uint8_t* trace;         // [2 * n_win][MPB_WALK_MAX_ROUNDS][32]
    mpb_cand* cands;        // [max_cands] compact candidate list of the current round
    mpb_cand* stage;        // [2 * n_win][k - 1] per-track staging
    int32_t* n_emit;        // [2 * n_win]
    uint32_t* plans;        // [max_cands][CSCAN_PLAN_WORDS]
    unsigned long long* counts;  // [max_cands][4]
    int* n_cand;            // candidates of the current round
    int* live_dev;          // [max_rounds + 1] live tracks after each advance
    unsigned long long* totals;  // [2] rounds with candidates, candidates scanned
    int* live_host;         // pinned mirror of live_dev
    int* err;               // track error flags (OR)
    bool fused;             // rounds go through k_walk_round (advance + compact + plans in one block)
    mpb_peer* peer;         // sharded run: the scan is followed by the peer-memory all-reduce of counts
};

__global__ void k_walk_seed(int n_win, int k, const int32_t* __restrict__ win_idx, const unsigned long long* __restrict__ freq,
                            const unsigned long long* __restrict__ nn, int tensors_by_hist_window,
                            const uint64_t* __restrict__ mm_key, mpb_track* __restrict__ tracks, int32_t* __restrict__ ntracks) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_win) return;
    const long long src = tensors_by_hist_window ? win_idx[w] : w;
    ntracks[w] = mpb_walk_seed(win_idx[w], k, reinterpret_cast<const int64_t*>(freq) + src * 4 * k,
                               reinterpret_cast<const int64_t*>(nn) + src * (k - 1) * 16, mm_key[w], &tracks[2 * w],
                               &tracks[2 * w + 1]);
    if (ntracks[w] == 1) tracks[2 * w + 1].state = 2;
}

// one thread per track slot: take the counts of the track's candidates of the previous round, advance, and stage the
// next candidates in the track's own region stage[s * (k - 1) ..] (n_emit[s] of them).
__device__ __forceinline__ void walk_advance_slot(int s, int n_win, int k, int dnum, int degeneracy, int round,
                                                  mpb_track* __restrict__ tracks, const int32_t* __restrict__ ntracks,
                                                  const int64_t* __restrict__ cover, const unsigned long long* __restrict__ counts,
                                                  mpb_cand* __restrict__ stage, int32_t* __restrict__ n_emit,
                                                  uint8_t* __restrict__ trace, int* __restrict__ err) {
    int n = 0;
    if ((s & 1) < ntracks[s >> 1]) {
        mpb_track& t = tracks[s];
        if (t.state != 2) {
            if (round > 0)
                mpb_walk_consume(t, k, reinterpret_cast<const int64_t*>(counts) + (long long)t.first_cand * 4, cover[s >> 1],
                                 dnum, degeneracy, trace + (long long)s * MPB_WALK_MAX_ROUNDS * 32);
            if (t.state != 2) n = mpb_walk_emit(t, k, stage + (long long)s * (k - 1)) | (1 << 30);  // bit 30: still live
            if (t.err) atomicOr(err, t.err);
        }
    }
    n_emit[s] = n;
}

__global__ void k_walk_advance(int n_win, int k, int dnum, int degeneracy, int round, mpb_track* __restrict__ tracks,
                               const int32_t* __restrict__ ntracks, const int64_t* __restrict__ cover,
                               const unsigned long long* __restrict__ counts, mpb_cand* __restrict__ stage,
                               int32_t* __restrict__ n_emit, uint8_t* __restrict__ trace, int* __restrict__ err) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= 2 * n_win) return;
    walk_advance_slot(s, n_win, k, dnum, degeneracy, round, tracks, ntracks, cover, counts, stage, n_emit, trace, err);
}

// ONE block: exclusive prefix sum of n_emit over the track slots -> every track's first candidate, candidates copied to
// their compact positions.  The slot order is the track order, so the candidate list is identical on every rank of a
// sequence-sharded run (the count vectors are summed element by element) and from run to run.
#define COMPACT_THREADS 1024
__device__ __forceinline__ int walk_compact_block(int n_slots, int k, int round, mpb_track* __restrict__ tracks,
                                                  const mpb_cand* __restrict__ stage, const int32_t* __restrict__ n_emit,
                                                  mpb_cand* __restrict__ cands, int* __restrict__ n_cand, int* __restrict__ live,
                                                  unsigned long long* __restrict__ totals, int* s_warp, int* s_live) {
    const int per = (n_slots + COMPACT_THREADS - 1) / COMPACT_THREADS;
    const int lo = threadIdx.x * per, hi = min(n_slots, lo + per);
    int sum = 0, alive = 0;
    for (int s = lo; s < hi; ++s) {
        sum += n_emit[s] & 0xFFFF;
        alive += n_emit[s] >> 30;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = sum;
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    const int alive_w = __reduce_add_sync(0xffffffffu, alive);
    if (lane == 31) s_warp[warp] = incl;
    if (lane == 0) s_live[warp] = alive_w;
    __syncthreads();
    if (warp == 0) {
        int v = s_warp[lane];
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += u;
        }
        s_warp[lane] = v;  // inclusive over warps
        int l = s_live[lane];
        l = __reduce_add_sync(0xffffffffu, l);
        if (lane == 0) s_live[0] = l;
    }
    __syncthreads();
    int base = incl - sum + (warp > 0 ? s_warp[warp - 1] : 0);
    for (int s = lo; s < hi; ++s) {
        const int n = n_emit[s] & 0xFFFF;
        if (n_emit[s] >> 30) {
            tracks[s].first_cand = base;
            for (int i = 0; i < n; ++i) cands[base + i] = stage[(long long)s * (k - 1) + i];
            base += n;
        }
    }
    const int total = s_warp[COMPACT_THREADS / 32 - 1];
    if (threadIdx.x == 0) {
        n_cand[0] = total;
        live[round] = s_live[0];
        if (total > 0) {
            totals[0] += 1ull;
            totals[1] += (unsigned long long)total;
        }
    }
    return total;
}

__global__ void __launch_bounds__(COMPACT_THREADS)
k_walk_compact(int n_slots, int k, int round, mpb_track* __restrict__ tracks, const mpb_cand* __restrict__ stage,
               const int32_t* __restrict__ n_emit, mpb_cand* __restrict__ cands, int* __restrict__ n_cand,
               int* __restrict__ live, unsigned long long* __restrict__ totals) {
    __shared__ int s_warp[COMPACT_THREADS / 32];
    __shared__ int s_live[COMPACT_THREADS / 32];
    walk_compact_block(n_slots, k, round, tracks, stage, n_emit, cands, n_cand, live, totals, s_warp, s_live);
}

// A whole round's bookkeeping in ONE block (batches of up to WALK_FUSED_MAX windows): advance every track, compact the
// staged candidates in track order, build the scan plans of the new candidates and clear their counters — three dependent
// launches of a few hundred threads each otherwise, a dozen times per window batch.
#define WALK_FUSED_MAX 2048
__global__ void __launch_bounds__(COMPACT_THREADS)
k_walk_round(int n_win, int k, int dnum, int degeneracy, int round, mpb_track* __restrict__ tracks,
             const int32_t* __restrict__ ntracks, const int64_t* __restrict__ cover, unsigned long long* __restrict__ counts,
             mpb_cand* __restrict__ stage, int32_t* __restrict__ n_emit, uint8_t* __restrict__ trace, mpb_cand* __restrict__ cands,
             int* __restrict__ n_cand, int* __restrict__ live, unsigned long long* __restrict__ totals,
             const int32_t* __restrict__ win_pos, int nw, uint32_t fmask, uint32_t rmask, uint32_t ones_row,
             uint32_t* __restrict__ plans, int* __restrict__ err, int* __restrict__ scan_err) {
    __shared__ int s_warp[COMPACT_THREADS / 32];
    __shared__ int s_live[COMPACT_THREADS / 32];
    for (int s = threadIdx.x; s < 2 * n_win; s += COMPACT_THREADS)
        walk_advance_slot(s, n_win, k, dnum, degeneracy, round, tracks, ntracks, cover, counts, stage, n_emit, trace, err);
    __syncthreads();  // (orders the block's global writes: n_emit, stage, tracks)
    const int total = walk_compact_block(2 * n_win, k, round, tracks, stage, n_emit, cands, n_cand, live, totals, s_warp, s_live);
    __syncthreads();
    for (int c = threadIdx.x; c < total; c += COMPACT_THREADS)
        cscan_plan_one(c, cands, win_pos, nw, k, fmask, rmask, ones_row, plans, counts, 1, scan_err);
}

extern "C" int mpb_walk_dev_begin(mpb_hist* h, int dnum, int degeneracy, uint32_t fmask, uint32_t rmask, int32_t n_win,
                                  const int32_t* win_idx, const int64_t* cover_number, const uint64_t* mm_key,
                                  const int64_t* freq_hd, const int64_t* nn_hd, mpb_walk_dev** out) {
    if (!h || !win_idx || !cover_number || !mm_key || !out) return fail(MPB_EINVAL, "NULL argument");
    if (n_win < 1) return fail(MPB_EINVAL, "no windows to walk");
    if ((freq_hd == nullptr) != (nn_hd == nullptr)) return fail(MPB_EINVAL, "freq and nn come together");
    if (!freq_hd && !h->have_summary) return fail(MPB_EINVAL, "mpb_hist_summary has not run on this handle");
    for (int i = 0; i < n_win; ++i)
        if (win_idx[i] < 0 || win_idx[i] >= h->nw) return fail(MPB_EINVAL, "win_idx[%d]=%d outside the batch", i, win_idx[i]);
    mpb_ctx* ctx = h->msa->ctx;
    CK(cudaSetDevice(ctx->device));
    const int k = h->k;
    mpb_walk_dev* w = new mpb_walk_dev();
    memset(w, 0, sizeof *w);
    w->h = h;
    w->k = k;
    w->dnum = dnum;
    w->degeneracy = degeneracy;
    w->n_win = n_win;
    w->fmask = fmask;
    w->rmask = rmask;
    w->max_cands = 2 * n_win * (k - 1);
    w->max_rounds = MPB_WALK_MAX_ROUNDS;
    w->fused = n_win <= WALK_FUSED_MAX;
    cudaError_t e = cudaMallocAsync(&w->tracks, (size_t)2 * n_win * sizeof(mpb_track), ctx->stream);
    if (e == cudaSuccess) e = cudaMallocAsync(&w->ntracks, (size_t)n_win * 4, ctx->stream);
    if (e == cudaSuccess) e = cudaMallocAsync(&w->cover, (size_t)n_win * 8, ctx->stream);
    if (e == cudaSuccess) e = cudaMallocAsync(&w->trace, (size_t)2 * n_win * MPB_WALK_MAX_ROUNDS * 32, ctx->stream);
    if (e == cudaSuccess) e = cudaMallocAsync(&w->cands, (size_t)w->max_cands * sizeof(mpb_cand), ctx->stream);
    if (e == cudaSuccess) e = cudaMallocAsync(&w->stage, (size_t)w->max_cands * sizeof(mpb_cand), ctx->stream);
    if (e == cudaSuccess) e = cudaMallocAsync(&w->n_emit, (size_t)2 * n_win * 4, ctx->stream);
    if (e == cudaSuccess) e = cudaMallocAsync(&w->plans, (size_t)w->max_cands * CSCAN_PLAN_WORDS * 4, ctx->stream);
    if (e == cudaSuccess) e = cudaMallocAsync(&w->counts, (size_t)w->max_cands * 4 * 8, ctx->stream);
    if (e == cudaSuccess) e = cudaMallocAsync(&w->n_cand, 8, ctx->stream);
    if (e == cudaSuccess) e = cudaMallocAsync(&w->live_dev, (size_t)(w->max_rounds + 1) * 4, ctx->stream);
    if (e == cudaSuccess) e = cudaMallocAsync(&w->totals, 16, ctx->stream);
    if (e == cudaSuccess) e = cudaMallocAsync(&w->err, 4, ctx->stream);
    if (e != cudaSuccess) {
        mpb_walk_dev_free(w);
        return fail(MPB_ENOMEM, "walk state for %d windows: %s", n_win, cudaGetErrorString(e));
    }
    w->live_host = ctx->pinned;  // one walk at a time per context
    for (int i = 0; i <= w->max_rounds; ++i) w->live_host[i] = -1;
    CK(cudaMemsetAsync(w->n_cand, 0, 8, ctx->stream));
    CK(cudaMemsetAsync(w->live_dev, 0, (size_t)(w->max_rounds + 1) * 4, ctx->stream));
    CK(cudaMemsetAsync(w->totals, 0, 16, ctx->stream));
    CK(cudaMemsetAsync(w->err, 0, 4, ctx->stream));
    CK(cudaMemsetAsync(w->counts, 0, (size_t)w->max_cands * 4 * 8, ctx->stream));
    CK(cudaMemcpyAsync(w->cover, cover_number, (size_t)n_win * 8, cudaMemcpyHostToDevice, ctx->stream));
    InBuf wi(ctx, win_idx, (size_t)n_win * 4), mk(ctx, mm_key, (size_t)n_win * 8);
    InBuf fq(ctx, freq_hd, (size_t)n_win * 4 * k * 8), nq(ctx, nn_hd, (size_t)n_win * (k - 1) * 16 * 8);
    if (wi.rc || mk.rc || fq.rc || nq.rc) {
        mpb_walk_dev_free(w);
        return MPB_ECUDA;
    }
    LAUNCH(ctx, k_walk_seed, (unsigned)((n_win + 63) / 64), 64, 0, n_win, k, wi.dev<int32_t>(),
           freq_hd ? fq.dev<unsigned long long>() : h->freq, freq_hd ? nq.dev<unsigned long long>() : h->nn,
           freq_hd ? 0 : 1, mk.dev<uint64_t>(), w->tracks, w->ntracks);
    // no synchronisation: the inputs are pageable host arrays (staged by the runtime before the copy calls return) or
    // device memory; the InBuf staging buffers are freed in stream order
    *out = w;
    return 0;
}

extern "C" int mpb_walk_dev_advance(mpb_walk_dev* w) {
    if (!w) return fail(MPB_EINVAL, "NULL argument");
    if (w->rounds_enqueued >= w->max_rounds) return fail(MPB_EOVERFLOW, "more than %d walk rounds", w->max_rounds);
    mpb_ctx* ctx = w->h->msa->ctx;
    CK(cudaSetDevice(ctx->device));
    const int r = w->rounds_enqueued;
    if (w->fused) {
        mpb_msa* m = w->h->msa;
        LAUNCH(ctx, k_walk_round, 1, COMPACT_THREADS, 0, w->n_win, w->k, w->dnum, w->degeneracy, r, w->tracks, w->ntracks,
               w->cover, w->counts, w->stage, w->n_emit, w->trace, w->cands, w->n_cand, w->live_dev, w->totals,
               w->h->win_pos, w->h->nw, w->fmask, w->rmask, MPB_COLP_ONES(m), w->plans, w->err, m->err);
    } else {
        LAUNCH(ctx, k_walk_advance, (unsigned)((2 * w->n_win + 63) / 64), 64, 0, w->n_win, w->k, w->dnum, w->degeneracy, r,
               w->tracks, w->ntracks, w->cover, w->counts, w->stage, w->n_emit, w->trace, w->err);
        LAUNCH(ctx, k_walk_compact, 1, COMPACT_THREADS, 0, 2 * w->n_win, w->k, r, w->tracks, w->stage, w->n_emit, w->cands,
               w->n_cand, w->live_dev, w->totals);
    }
    CK(cudaMemcpyAsync(&w->live_host[r], &w->live_dev[r], 4, cudaMemcpyDeviceToHost, ctx->stream));
    w->rounds_enqueued = r + 1;
    return 0;
}

extern "C" int mpb_walk_dev_scan(mpb_walk_dev* w) {
    if (!w) return fail(MPB_EINVAL, "NULL argument");
    mpb_ctx* ctx = w->h->msa->ctx;
    CK(cudaSetDevice(ctx->device));
    int rc = mpb_cscan_launch(w->h, w->fmask, w->rmask, w->cands, w->n_cand, w->max_cands, w->plans, w->counts, 1, nullptr,
                              nullptr, w->fused ? 1 : 0);
    if (rc || !w->peer) return rc;
    return mpb_peer_allreduce_launch(w->peer, w->counts, w->n_cand, 4, w->err);
}

extern "C" int mpb_walk_dev_set_peer(mpb_walk_dev* w, mpb_peer* peer) {
    if (!w) return fail(MPB_EINVAL, "NULL argument");
    if (peer && (int64_t)w->max_cands * 4 > mpb_peer_cap(peer))
        return fail(MPB_EINVAL, "a round of this walk may hold %lld counters, the peer group carries %lld",
                    (long long)w->max_cands * 4, (long long)mpb_peer_cap(peer));
    w->peer = peer;
    return 0;
}

extern "C" int mpb_walk_dev_round(mpb_walk_dev* w) {
    int rc = mpb_walk_dev_advance(w);
    if (rc) return rc;
    return mpb_walk_dev_scan(w);
}

extern "C" int mpb_walk_dev_counts(mpb_walk_dev* w, void** counts_dev, int64_t* n_elems) {
    if (!w || !counts_dev || !n_elems) return fail(MPB_EINVAL, "NULL argument");
    *counts_dev = w->counts;
    *n_elems = (int64_t)w->max_cands * 4;
    return 0;
}

extern "C" int64_t mpb_walk_dev_live(mpb_walk_dev* w) {
    if (!w) return -1;
    int64_t last = -1;
    for (int r = 0; r < w->rounds_enqueued; ++r) {
        const int v = *((volatile int*)&w->live_host[r]);
        if (v < 0) break;
        last = v;
    }
    return last;
}

extern "C" int mpb_walk_dev_max_rounds(mpb_walk_dev* w) { return w ? w->max_rounds : 0; }

// Block (spin on the pinned mirror) until the number of live tracks after advance number `round` is known.
extern "C" int mpb_walk_dev_wait(mpb_walk_dev* w, int round, int64_t* live) {
    if (!w || !live || round < 0 || round >= w->rounds_enqueued) return fail(MPB_EINVAL, "no such round");
    mpb_ctx* ctx = w->h->msa->ctx;
    volatile int* p = (volatile int*)&w->live_host[round];
    for (unsigned spin = 0;; ++spin) {
        const int v = *p;
        if (v >= 0) {
            *live = v;
            return 0;
        }
        if ((spin & 0x3FFu) == 0x3FFu) {
            const cudaError_t e = cudaStreamQuery(ctx->stream);
            if (e == cudaSuccess) {  // stream drained: the copy has landed (or never will)
                if (*p >= 0) continue;
                return fail(MPB_ECUDA, "walk round %d never reported", round);
            }
            if (e != cudaErrorNotReady) return fail(MPB_ECUDA, "walk: %s", cudaGetErrorString(e));
        }
    }
}

// Single-process driver: enqueue rounds, staying at most `lag` rounds ahead of the device, until no track is live.
extern "C" int mpb_walk_dev_run(mpb_walk_dev* w, int lag, int64_t* rounds_out) {
    if (!w) return fail(MPB_EINVAL, "NULL argument");
    if (lag < 0) lag = 0;
    for (;;) {
        int rc = mpb_walk_dev_advance(w);
        if (rc) return rc;
        const int r = w->rounds_enqueued - 1;
        if (r >= lag) {
            int64_t live = 0;
            rc = mpb_walk_dev_wait(w, r - lag, &live);
            if (rc) return rc;
            if (live == 0) break;
        }
        rc = mpb_walk_dev_scan(w);
        if (rc) return rc;
    }
    if (rounds_out) *rounds_out = w->rounds_enqueued;
    return 0;
}

extern "C" int mpb_walk_dev_finish(mpb_walk_dev* w, uint8_t* out_sets, int64_t* out_counts, uint8_t* out_seeds,
                                   int64_t* out_seed_cover, int32_t* out_ntracks, int64_t trace_cap, uint8_t* trace_sets,
                                   int64_t* trace_off, int64_t* stats) {
    if (!w || !out_sets || !out_counts || !out_seeds || !out_seed_cover || !out_ntracks || !trace_off)
        return fail(MPB_EINVAL, "NULL argument");
    mpb_ctx* ctx = w->h->msa->ctx;
    CK(cudaSetDevice(ctx->device));
    const int n_win = w->n_win;
    std::vector<mpb_track> tracks((size_t)2 * n_win);
    std::vector<int32_t> ntr(n_win);
    std::vector<uint8_t> trace(trace_sets ? (size_t)2 * n_win * MPB_WALK_MAX_ROUNDS * 32 : 0);
    unsigned long long totals[2] = {0, 0};
    int err = 0;
    CK(cudaMemcpyAsync(tracks.data(), w->tracks, tracks.size() * sizeof(mpb_track), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(ntr.data(), w->ntracks, (size_t)n_win * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (trace_sets) CK(cudaMemcpyAsync(trace.data(), w->trace, trace.size(), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(totals, w->totals, 16, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(&err, w->err, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (err & 1) return fail(MPB_EINVAL, "refinement would re-add a base (the reference raises KeyError)");
    if (err & 2) return fail(MPB_EOVERFLOW, "more than %d refinement rounds in one window", MPB_WALK_MAX_ROUNDS);
    if (err & 4) return fail(MPB_EOVERFLOW, "candidate buffer overflow");
    if (err & MPB_ERR_PEER_TIMEOUT) return fail(MPB_ECUDA, "sharded walk: a peer's counts did not arrive (peer all-reduce timed out)");
    if (err & MPB_ERR_PEER_CAP) return fail(MPB_EOVERFLOW, "sharded walk: count vector longer than the peer group's capacity");
    // the last scanned round's candidates are not in totals yet when live tracks remain (caller stopped early)
    for (int s = 0; s < 2 * n_win; ++s)
        if ((s & 1) < ntr[s >> 1] && tracks[s].state != 2) return fail(MPB_EINVAL, "walk not finished: track %d still live", s);
    int64_t tr = 0;
    for (int wdx = 0; wdx < n_win; ++wdx) {
        trace_off[wdx] = tr;
        const mpb_track* a = &tracks[2 * (size_t)wdx];
        const int pick = ntr[wdx] == 2 ? mpb_walk_pick(a[0], a[1]) : 0;
        const mpb_track& t = a[pick];
        memcpy(out_sets + (int64_t)wdx * 32, t.sets, 32);
        out_counts[wdx * 5 + 0] = t.init;
        out_counts[wdx * 5 + 1] = t.fm;
        out_counts[wdx * 5 + 2] = t.rm;
        out_counts[wdx * 5 + 3] = pick;
        out_counts[wdx * 5 + 4] = t.perfect;
        out_ntracks[wdx] = ntr[wdx];
        for (int ti = 0; ti < 2; ++ti) {
            out_seed_cover[wdx * 2 + ti] = ti < ntr[wdx] ? a[ti].seed_cover : -1;
            if (ti < ntr[wdx]) memcpy(out_seeds + ((int64_t)wdx * 2 + ti) * 32, a[ti].seed, 32);
            else memset(out_seeds + ((int64_t)wdx * 2 + ti) * 32, 0, 32);
        }
        for (int ti = 0; ti < ntr[wdx]; ++ti)
            for (int r = 0; r < a[ti].n_trace; ++r) {
                if (trace_sets && tr < trace_cap)
                    memcpy(trace_sets + tr * 32, &trace[((size_t)(2 * wdx + ti) * MPB_WALK_MAX_ROUNDS + r) * 32], 32);
                ++tr;
            }
    }
    trace_off[n_win] = tr;
    if (stats) {
        stats[0] = (int64_t)totals[0];
        stats[1] = (int64_t)totals[1];
        stats[2] = tr;
    }
    ctx->extra_units["k_cscan"] += (double)totals[1] * (double)w->h->msa->n_seq;
    if (trace_sets && tr > trace_cap) return fail(MPB_EOVERFLOW, "trace capacity %lld < %lld", (long long)trace_cap, (long long)tr);
    return 0;
}

extern "C" void mpb_walk_dev_free(mpb_walk_dev* w) {
    if (!w) return;
    cudaStream_t st = w->h->msa->ctx->stream;
    void* ptrs[] = {w->tracks, w->ntracks, w->cover, w->trace, w->cands, w->stage, w->n_emit, w->plans, w->counts,
                    w->n_cand, w->live_dev, w->totals, w->err};
    for (void* p : ptrs)
        if (p) cudaFreeAsync(p, st);
    delete w;
}
