// mpb_dimer.cu — primer-dimer predicates on the GPU (core:457-503 dimer_check, finDimer_V4.py:191-224,
// get_Maxprimerset_V1.3.py:193-215, get_multiPrime.py:419-437).
//
// The reference enumerates, for a primer pair (i, j): every 3' end e of i (suffix lengths high to low, each suffix
// expanded in product order) x every expansion p of j, takes the LEFTMOST occurrence idx of RC(e) in p and tests
//     Loss(len, GC(e), 0, d2 = len(p) - len(e) - idx) >= threshold   or   (dG(e) < -5 and d2 == 0).
// Here both primers' expansions are materialised once (2-bit packed), the Loss comparison is a host-built truth
// table [len][GC][d2] (so it is decided by the reference's own float expression), dG(e) is summed in the reference's
// operation order without FMA, and each (e, p) pair is one thread-iteration.  The first hit in reference order is
// the minimum of e_index * n_p + p_index.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "mpb200.h"
#include "mpb_host.h"

#define DIMER_MAXLEN 32

struct mpb_dimer {
    mpb_ctx* ctx;
    int n;
    uint8_t* sets;       // [n][32] 4-bit sets
    int32_t* lens;       // [n]
    int64_t* off_p;      // [n+1] expansion offsets
    int64_t* off_e;      // [n+1] end offsets
    uint64_t* exp;       // packed expansions
    uint64_t* end_rc;    // packed reverse complement of each end
    uint32_t* end_info;  // len | gc << 8 | dgflag << 16
    uint8_t* table;      // [33][33][33] loss >= threshold
    uint32_t* pent;      // [n][32] 5-mer sets (built on the first grid call)
    int min_end;
    std::vector<int64_t> h_off_p, h_off_e;
};

__constant__ uint8_t d_fold[16] = {1, 1, 1, 2, 1, 2, 2, 3, 1, 2, 2, 3, 2, 3, 3, 4};
#define O2(a, b) ((a) | ((b) << 2))
#define O3(a, b, c) ((a) | ((b) << 2) | ((c) << 4))
#define O4(a, b, c, d) ((a) | ((b) << 2) | ((c) << 4) | ((d) << 6))
// expansion order of each base set (core:105-107), see mpb_device.cuh
__constant__ uint8_t d_order[16] = {0, 0, 1, O2(0, 1), 2, O2(0, 2), O2(2, 1), O3(2, 0, 1), 3, O2(0, 3), O2(1, 3),
                                    O3(0, 3, 1), O2(2, 3), O3(2, 0, 3), O3(2, 3, 1), O4(0, 3, 2, 1)};

// e-th expansion (product order, leftmost position slowest) of sets[a..b) -> bases packed 2 bits each, position a at
// bits 0..1
__device__ __forceinline__ uint64_t expand_packed(const uint8_t* sets, int a, int b, uint64_t e) {
    uint64_t out = 0;
    for (int i = b - 1; i >= a; --i) {
        const int code = sets[i];
        const unsigned n = d_fold[code];
        const unsigned d = (unsigned)(e % n);
        e /= n;
        const uint64_t base = (d_order[code] >> (2 * d)) & 3u;
        out |= base << (2 * (i - a));
    }
    return out;
}

__device__ __forceinline__ int find_owner(const int64_t* off, int n, int64_t idx) {
    int lo = 0, hi = n;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (off[mid] <= idx) lo = mid;
        else hi = mid;
    }
    return lo;
}

__global__ void k_dimer_expand(const uint8_t* __restrict__ sets, const int32_t* __restrict__ lens,
                               const int64_t* __restrict__ off_p, int n, uint64_t* __restrict__ exp) {
    const int64_t total = off_p[n];
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int i = find_owner(off_p, n, idx);
        exp[idx] = expand_packed(sets + (int64_t)i * DIMER_MAXLEN, 0, lens[i], (uint64_t)(idx - off_p[i]));
    }
}

// dg consts: [0..15] stack[next][cur] = freedom*hbonds+penalty, [16..19] init (A,C,G,T), [20] terminal TA, [21] per-base
// salt term, [22] symmetry, [23] threshold: dG(e) < -5 (after round(.,2))  <=>  g <= consts[23]
__global__ void k_dimer_ends(const uint8_t* __restrict__ sets, const int32_t* __restrict__ lens,
                             const int64_t* __restrict__ off_e, int n, int min_end, int max_end, int init_both,
                             const double* __restrict__ cst, uint64_t* __restrict__ end_rc,
                             uint32_t* __restrict__ end_info) {
    const int64_t total = off_e[n];
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int i = find_owner(off_e, n, idx);
        const uint8_t* S = sets + (int64_t)i * DIMER_MAXLEN;
        const int k = lens[i];
        int64_t r = idx - off_e[i];
        // suffix lengths from high to low: min(max_end, k) .. min_end   (max_end <= 0 means k + max_end)
        int L = max_end > 0 ? (max_end < k ? max_end : k) : k + max_end;
        for (; L >= min_end; --L) {
            int64_t cnt = 1;
            for (int q = k - L; q < k; ++q) cnt *= d_fold[S[q]];
            if (r < cnt) break;
            r -= cnt;
        }
        const uint64_t e = expand_packed(S, k - L, k, (uint64_t)r);  // position j of the end at bits 2j
        // reverse complement: rc[t] = 3 - e[L-1-t]
        uint64_t rc = 0;
        int gc = 0;
        for (int t = 0; t < L; ++t) {
            const uint64_t b = (e >> (2 * (L - 1 - t))) & 3u;
            rc |= (3u - b) << (2 * t);
            gc += (b == 1) | (b == 2);
        }
        // dG of the plain end (core:466-485 with a single expansion), reference operation order, no FMA
        double g = 0.0;
        for (int q = 0; q + 1 < L; ++q) {
            const int cur = (int)((e >> (2 * q)) & 3u), nxt = (int)((e >> (2 * (q + 1))) & 3u);
            g = __dadd_rn(g, cst[nxt * 4 + cur]);
        }
        const int b0 = (int)(e & 3u), bl = (int)((e >> (2 * (L - 1))) & 3u);
        double t = init_both ? __dadd_rn(cst[16 + b0], cst[16 + bl]) : cst[16 + b0];
        const bool ta = L >= 2 && ((e >> (2 * (L - 2))) & 3u) == 3u && bl == 0;  // end[-2:] == "TA"
        if (ta) t = __dadd_rn(t, cst[20]);
        g = __dadd_rn(g, t);
        g = __dsub_rn(g, __dmul_rn(cst[21], (double)L));
        bool sym = (L % 2) == 0;
        for (int q = 0; sym && q < L / 2; ++q) sym = (((e >> (2 * q)) & 3u) + ((e >> (2 * (L / 2 + q))) & 3u)) == 3u;
        if (sym) g = __dadd_rn(g, cst[22]);
        const uint32_t flag = g <= cst[23] ? 1u : 0u;
        end_rc[idx] = rc;
        end_info[idx] = (uint32_t)L | ((uint32_t)gc << 8) | (flag << 16);
    }
}

// leftmost occurrence of the L-base pattern rc inside the kj-base string p (both 2-bit packed), or -1
__device__ __forceinline__ int find_leftmost(uint64_t p, int kj, uint64_t rc, int L) {
    const uint64_t mask = L >= 32 ? ~0ull : ((1ull << (2 * L)) - 1ull);
    for (int o = 0; o + L <= kj; ++o)
        if (((p >> (2 * o)) & mask) == rc) return o;
    return -1;
}

// block per pair: first hit (in reference order) among ends(i) x expansions(j).  Any block size works (the order
// index decides, not the thread): 128 threads when there are many pairs, 512 when a few pairs of highly degenerate
// primers (the self-dimer gate of the window pipeline) would otherwise leave most of the GPU idle.
#define PAIR_THREADS 128
#define PAIR_THREADS_WIDE 512
__global__ void __launch_bounds__(PAIR_THREADS_WIDE)
k_dimer_pairs(const int32_t* __restrict__ pi, const int32_t* __restrict__ pj, const int32_t* __restrict__ lens,
              const int64_t* __restrict__ off_p, const int64_t* __restrict__ off_e, const uint64_t* __restrict__ exp,
              const uint64_t* __restrict__ end_rc, const uint32_t* __restrict__ end_info,
              const uint8_t* __restrict__ table, long long* __restrict__ first_hit, int32_t* __restrict__ hit_d2) {
    __shared__ unsigned long long best;
    __shared__ int best_d2;
    const int q = blockIdx.x;
    const int i = pi[q], j = pj[q];
    const int64_t ne = off_e[i + 1] - off_e[i], np = off_p[j + 1] - off_p[j];
    const int kj = lens[j];
    const unsigned long long total = (unsigned long long)ne * (unsigned long long)np;
    if (threadIdx.x == 0) {
        best = ~0ull;
        best_d2 = -1;
    }
    __syncthreads();
    for (unsigned long long base = 0; base < total; base += blockDim.x) {
        const unsigned long long idx = base + threadIdx.x;
        if (idx < total) {
            const int64_t e = (int64_t)(idx / (unsigned long long)np), p = (int64_t)(idx % (unsigned long long)np);
            const uint32_t info = end_info[off_e[i] + e];
            const int L = info & 255, gc = (info >> 8) & 255;
            const int o = find_leftmost(exp[off_p[j] + p], kj, end_rc[off_e[i] + e], L);
            if (o >= 0) {
                const int d2 = kj - L - o;
                if (table[(L * 33 + gc) * 33 + d2] || (d2 == 0 && (info >> 16))) {
                    const unsigned long long old = atomicMin(&best, idx);
                    (void)old;
                }
            }
        }
        __syncthreads();
        if (best != ~0ull) break;  // later chunks only hold larger order indices
        __syncthreads();
    }
    __syncthreads();
    if (best != ~0ull) {
        const unsigned long long idx = best;
        if ((idx % blockDim.x) == threadIdx.x) {
            const int64_t e = (int64_t)(idx / (unsigned long long)np), p = (int64_t)(idx % (unsigned long long)np);
            const int L = end_info[off_e[i] + e] & 255;
            best_d2 = kj - L - find_leftmost(exp[off_p[j] + p], kj, end_rc[off_e[i] + e], L);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        first_hit[q] = best == ~0ull ? -1ll : (long long)best;
        if (hit_d2) hit_d2[q] = best_d2;
    }
}

extern "C" int mpb_dimer_prepare(mpb_ctx* ctx, const uint8_t* sets, const int32_t* lens, int32_t n, int min_end,
                                 int max_end, int init_both, const uint8_t* loss_table, const double* dg_consts,
                                 mpb_dimer** out) {
    if (!ctx || !sets || !lens || !loss_table || !dg_consts || !out) return mpb_fail(MPB_EINVAL, "NULL argument");
    if (n < 1 || min_end < 1 || (max_end > 0 && max_end < min_end) || max_end > DIMER_MAXLEN)
        return mpb_fail(MPB_EINVAL, "bad n=%d or end range %d..%d", n, min_end, max_end);
    MPB_CK(cudaSetDevice(mpb_ctx_device(ctx)));
    cudaStream_t st = mpb_ctx_stream(ctx);
    mpb_dimer* d = new mpb_dimer;
    d->ctx = ctx;
    d->n = n;
    d->sets = nullptr;
    d->lens = nullptr;
    d->off_p = d->off_e = nullptr;
    d->exp = d->end_rc = nullptr;
    d->end_info = nullptr;
    d->table = nullptr;
    d->pent = nullptr;
    d->min_end = min_end;
    d->h_off_p.assign(n + 1, 0);
    d->h_off_e.assign(n + 1, 0);
    static const int fold[16] = {1, 1, 1, 2, 1, 2, 2, 3, 1, 2, 2, 3, 2, 3, 3, 4};
    for (int i = 0; i < n; ++i) {
        const int k = lens[i];
        if (k < min_end || k > DIMER_MAXLEN) {
            delete d;
            return mpb_fail(MPB_EINVAL, "primer %d has length %d (supported: %d..%d)", i, k, min_end, DIMER_MAXLEN);
        }
        int64_t deg = 1, ends = 0, suf = 1;
        for (int q = k - 1; q >= 0; --q) {
            const int f = fold[sets[(int64_t)i * DIMER_MAXLEN + q] & 15];
            if (f == 0 || sets[(int64_t)i * DIMER_MAXLEN + q] == 0) {
                delete d;
                return mpb_fail(MPB_EINVAL, "primer %d holds a gap / empty set", i);
            }
            deg *= f;
            suf *= f;
            const int L = k - q;
            const int lmax = max_end > 0 ? max_end : k + max_end;
            if (L >= min_end && L <= lmax) ends += suf;
            if (deg > (1ll << 40)) break;
        }
        if (deg > (1ll << 24)) {
            delete d;
            return mpb_fail(MPB_EINVAL, "primer %d expands to more than 2^24 sequences", i);
        }
        d->h_off_p[i + 1] = d->h_off_p[i] + deg;
        d->h_off_e[i + 1] = d->h_off_e[i] + ends;
    }
    const int64_t tp = d->h_off_p[n], te = d->h_off_e[n];
    double* cst = nullptr;
    cudaError_t e = cudaMallocAsync(&d->sets, (size_t)n * DIMER_MAXLEN, st);
    if (e == cudaSuccess) e = cudaMallocAsync(&d->lens, (size_t)n * 4, st);
    if (e == cudaSuccess) e = cudaMallocAsync(&d->off_p, (size_t)(n + 1) * 8, st);
    if (e == cudaSuccess) e = cudaMallocAsync(&d->off_e, (size_t)(n + 1) * 8, st);
    if (e == cudaSuccess) e = cudaMallocAsync(&d->exp, (size_t)tp * 8, st);
    if (e == cudaSuccess) e = cudaMallocAsync(&d->end_rc, (size_t)(te ? te : 1) * 8, st);
    if (e == cudaSuccess) e = cudaMallocAsync(&d->end_info, (size_t)(te ? te : 1) * 4, st);
    if (e == cudaSuccess) e = cudaMallocAsync(&d->table, 33 * 33 * 33, st);
    if (e == cudaSuccess) e = cudaMallocAsync(&cst, 24 * 8, st);
    if (e != cudaSuccess) return mpb_fail(MPB_ENOMEM, "dimer tables: %s", cudaGetErrorString(e));
    MPB_CK(cudaMemcpyAsync(d->sets, sets, (size_t)n * DIMER_MAXLEN, cudaMemcpyHostToDevice, st));
    MPB_CK(cudaMemcpyAsync(d->lens, lens, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    MPB_CK(cudaMemcpyAsync(d->off_p, d->h_off_p.data(), (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, st));
    MPB_CK(cudaMemcpyAsync(d->off_e, d->h_off_e.data(), (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, st));
    MPB_CK(cudaMemcpyAsync(d->table, loss_table, 33 * 33 * 33, cudaMemcpyHostToDevice, st));
    MPB_CK(cudaMemcpyAsync(cst, dg_consts, 24 * 8, cudaMemcpyHostToDevice, st));
    const int sm = mpb_ctx_sms(ctx);
    unsigned g1 = (unsigned)((tp + 255) / 256), g2 = (unsigned)((te + 255) / 256);
    if (g1 > (unsigned)sm * 32) g1 = sm * 32;
    if (g2 > (unsigned)sm * 32) g2 = sm * 32;
    MPB_LAUNCH(ctx, k_dimer_expand, g1, 256, 0, d->sets, d->lens, d->off_p, n, d->exp);
    if (te > 0)
        MPB_LAUNCH(ctx, k_dimer_ends, g2, 256, 0, d->sets, d->lens, d->off_e, n, min_end, max_end, init_both, cst,
                   d->end_rc, d->end_info);
    MPB_CK(cudaStreamSynchronize(st));
    cudaFreeAsync(cst, st);
    *out = d;
    return 0;
}

extern "C" void mpb_dimer_free(mpb_dimer* d) {
    if (!d) return;
    cudaStream_t st = mpb_ctx_stream(d->ctx);
    void* bufs[] = {d->sets, d->lens, d->off_p, d->off_e, d->exp, d->end_rc, d->end_info, d->table, d->pent};
    for (void* b : bufs)
        if (b) cudaFreeAsync(b, st);
    delete d;
}

extern "C" int mpb_dimer_counts(mpb_dimer* d, int64_t* off_p, int64_t* off_e) {
    if (!d) return mpb_fail(MPB_EINVAL, "NULL argument");
    if (off_p) memcpy(off_p, d->h_off_p.data(), (size_t)(d->n + 1) * 8);
    if (off_e) memcpy(off_e, d->h_off_e.data(), (size_t)(d->n + 1) * 8);
    return 0;
}

extern "C" int mpb_dimer_pairs(mpb_dimer* d, const int32_t* pi, const int32_t* pj, int64_t n_pairs,
                               int64_t* first_hit, int32_t* hit_d2) {
    if (!d || !pi || !pj || !first_hit) return mpb_fail(MPB_EINVAL, "NULL argument");
    if (n_pairs < 1) return 0;
    for (int64_t q = 0; q < n_pairs; ++q)
        if (pi[q] < 0 || pi[q] >= d->n || pj[q] < 0 || pj[q] >= d->n)
            return mpb_fail(MPB_EINVAL, "pair %lld outside the primer table", (long long)q);
    mpb_ctx* ctx = d->ctx;
    MPB_CK(cudaSetDevice(mpb_ctx_device(ctx)));
    cudaStream_t st = mpb_ctx_stream(ctx);
    int32_t *dpi, *dpj, *dd2;
    long long* dfh;
    MPB_CK(cudaMallocAsync(&dpi, n_pairs * 4, st));
    MPB_CK(cudaMallocAsync(&dpj, n_pairs * 4, st));
    MPB_CK(cudaMallocAsync(&dd2, n_pairs * 4, st));
    MPB_CK(cudaMallocAsync(&dfh, n_pairs * 8, st));
    MPB_CK(cudaMemcpyAsync(dpi, pi, n_pairs * 4, cudaMemcpyHostToDevice, st));
    MPB_CK(cudaMemcpyAsync(dpj, pj, n_pairs * 4, cudaMemcpyHostToDevice, st));
    const int pair_threads = n_pairs < 4ll * mpb_ctx_sms(ctx) ? PAIR_THREADS_WIDE : PAIR_THREADS;
    MPB_LAUNCH(ctx, k_dimer_pairs, (unsigned)n_pairs, pair_threads, 0, dpi, dpj, d->lens, d->off_p, d->off_e, d->exp,
               d->end_rc, d->end_info, d->table, dfh, dd2);
    MPB_CK(cudaMemcpyAsync(first_hit, dfh, n_pairs * 8, cudaMemcpyDeviceToHost, st));
    if (hit_d2) MPB_CK(cudaMemcpyAsync(hit_d2, dd2, n_pairs * 4, cudaMemcpyDeviceToHost, st));
    MPB_CK(cudaStreamSynchronize(st));
    cudaFreeAsync(dpi, st);
    cudaFreeAsync(dpj, st);
    cudaFreeAsync(dd2, st);
    cudaFreeAsync(dfh, st);
    return 0;
}

// ------------------------------------------------------------------------------------------------------
// all-pairs grid (finDimer_V4.py:191-224: every primer i against every primer j >= i)
// ------------------------------------------------------------------------------------------------------
// A pair can only form a dimer if the reverse complement of some expansion of i's LAST 5 bases occurs in some
// expansion of j (every longer 3' end ends with those 5 bases, so its reverse complement starts with theirs).
// pent[j]   : 1024-bit set of the 5-mers that occur in any expansion of primer j (position-wise compatible)
// tail5[i]  : up to TAIL_MAX packed reverse complements of the expansions of i's last min_end bases
#define TAIL_MAX 16

__global__ void k_dimer_pent(const uint8_t* __restrict__ sets, const int32_t* __restrict__ lens, int n, int m,
                             uint32_t* __restrict__ pent /* [n][32] when m == 5 */) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint8_t* S = sets + (int64_t)j * DIMER_MAXLEN;
    const int k = lens[j];
    uint32_t* out = pent + (int64_t)j * 32;
    for (int w = 0; w < 32; ++w) out[w] = 0;
    for (int o = 0; o + m <= k; ++o) {
        // enumerate the expansions of S[o..o+m)
        int total = 1;
        for (int q = 0; q < m; ++q) total *= d_fold[S[o + q]];
        for (int e = 0; e < total; ++e) {
            const uint32_t code = (uint32_t)expand_packed(S, o, o + m, (uint64_t)e);
            out[code >> 5] |= 1u << (code & 31);
        }
    }
}

// thread = (i, j) pair with j >= i; survivors are appended to a queue
__global__ void k_dimer_prefilter(const int64_t* __restrict__ off_e, const uint64_t* __restrict__ end_rc,
                                  const uint32_t* __restrict__ end_info, const uint32_t* __restrict__ pent, int n,
                                  int row0, int row1, int min_end, int2* __restrict__ queue,
                                  unsigned long long* __restrict__ qn, long long qcap) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int i = row0 + blockIdx.y;
    if (i >= row1 || j < i || j >= n) return;
    // the ends of i are stored longest first; the shortest length (min_end) block is last
    const uint32_t* P = pent + j * 32;
    const uint32_t mask = (1u << (2 * min_end)) - 1u;
    bool hit = false;
    for (int64_t e = off_e[i + 1] - 1; e >= off_e[i]; --e) {
        if ((int)(end_info[e] & 255) != min_end) break;
        const uint32_t code = (uint32_t)end_rc[e] & mask;
        if ((P[code >> 5] >> (code & 31)) & 1u) {
            hit = true;
            break;
        }
    }
    if (hit) {
        const unsigned long long slot = atomicAdd(qn, 1ull);
        if ((long long)slot < qcap) queue[slot] = make_int2(i, (int)j);
    }
}

// Rows [row0, row1) of the upper-triangular pair grid.  Outputs (host arrays of capacity max_hits): the pairs that
// form a dimer with their first-hit order index and distance 2; *n_hits their number; *n_tested the pairs that
// survived the 5-mer prefilter.  Returns MPB_EOVERFLOW when max_hits / the internal queue is too small (call again
// with fewer rows).
extern "C" int mpb_dimer_grid(mpb_dimer* d, int32_t row0, int32_t row1, int64_t max_hits, int32_t* hit_i,
                              int32_t* hit_j, int64_t* hit_order, int32_t* hit_d2, int64_t* n_hits,
                              int64_t* n_tested) {
    if (!d || !hit_i || !hit_j || !hit_order || !hit_d2 || !n_hits) return mpb_fail(MPB_EINVAL, "NULL argument");
    if (row0 < 0 || row1 > d->n || row0 >= row1) return mpb_fail(MPB_EINVAL, "bad row range");
    mpb_ctx* ctx = d->ctx;
    MPB_CK(cudaSetDevice(mpb_ctx_device(ctx)));
    cudaStream_t st = mpb_ctx_stream(ctx);
    if (d->min_end != 5) return mpb_fail(MPB_EINVAL, "the pair grid needs min_end == 5");
    if (!d->pent) {
        MPB_CK(cudaMallocAsync(&d->pent, (size_t)d->n * 32 * 4, st));
        MPB_LAUNCH(ctx, k_dimer_pent, (unsigned)((d->n + 127) / 128), 128, 0, d->sets, d->lens, d->n, 5, d->pent);
    }
    const long long qcap = 1ll << 26;
    int2* queue;
    unsigned long long* qn;
    MPB_CK(cudaMallocAsync(&queue, qcap * sizeof(int2), st));
    MPB_CK(cudaMallocAsync(&qn, 8, st));
    MPB_CK(cudaMemsetAsync(qn, 0, 8, st));
    dim3 grid((unsigned)((d->n + 255) / 256), (unsigned)(row1 - row0));
    MPB_LAUNCH(ctx, k_dimer_prefilter, grid, 256, 0, d->off_e, d->end_rc, d->end_info, d->pent, d->n, row0, row1, 5,
               queue, qn, qcap);
    unsigned long long nq = 0;
    MPB_CK(cudaMemcpyAsync(&nq, qn, 8, cudaMemcpyDeviceToHost, st));
    MPB_CK(cudaStreamSynchronize(st));
    if ((long long)nq > qcap) {
        cudaFreeAsync(queue, st);
        cudaFreeAsync(qn, st);
        return mpb_fail(MPB_EOVERFLOW, "dimer grid: %llu candidate pairs in one band, use fewer rows", nq);
    }
    if (n_tested) *n_tested = (int64_t)nq;
    *n_hits = 0;
    int rc = 0;
    if (nq > 0) {
        std::vector<int2> hq(nq);
        MPB_CK(cudaMemcpyAsync(hq.data(), queue, nq * sizeof(int2), cudaMemcpyDeviceToHost, st));
        MPB_CK(cudaStreamSynchronize(st));
        std::sort(hq.begin(), hq.end(), [](const int2& a, const int2& b) { return a.x < b.x || (a.x == b.x && a.y < b.y); });
        std::vector<int32_t> pi(nq), pj(nq), d2(nq);
        std::vector<int64_t> fh(nq);
        for (size_t q = 0; q < nq; ++q) {
            pi[q] = hq[q].x;
            pj[q] = hq[q].y;
        }
        rc = mpb_dimer_pairs(d, pi.data(), pj.data(), (int64_t)nq, fh.data(), d2.data());
        if (rc == 0) {
            int64_t nh = 0;
            for (size_t q = 0; q < nq; ++q) {
                if (fh[q] < 0) continue;
                if (nh < max_hits) {
                    hit_i[nh] = pi[q];
                    hit_j[nh] = pj[q];
                    hit_order[nh] = fh[q];
                    hit_d2[nh] = d2[q];
                }
                ++nh;
            }
            *n_hits = nh;
            if (nh > max_hits) rc = mpb_fail(MPB_EOVERFLOW, "dimer grid: %lld hits exceed max_hits", (long long)nh);
        }
    }
    cudaFreeAsync(queue, st);
    cudaFreeAsync(qn, st);
    return rc;
}
