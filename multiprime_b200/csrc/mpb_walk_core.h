// mpb_walk_core.h — the per-window control logic of multiPrime-core (seeds core:579-600, the NN-array refinement walk
// core:860-1089, NM-vs-MM choice core:816) written ONCE for host and device: the device walk (mpb_walk_dev.cu: one
// thread per track, state resident in HBM, chained with the candidate scan without a host round trip) and the host
// driver of the CPU tests (mpb_walk in mpb_walk.cu, scan as a callback) compile the same functions.
// Everything here touches O(k) numbers per track.
#pragma once
#include <stdint.h>

#include "mpb200.h"

#ifdef __CUDACC__
#define MPB_HD __host__ __device__ __forceinline__
#else
#define MPB_HD inline
#endif

#define MPB_WALK_MAX_ROUNDS 48  // trace slots per track; a track adds one base per round (<= 3 per position)

// mpb_cand (include/mpb200.h): one candidate of a scan round — the primer `allow` of window `win`; when trial >= 0
// the scan also counts the rows that match the primer perfectly WITH base (trial >> 8) at position (trial & 255): the
// reference's coverage_renew look-up of the "single new base" pattern (core:954-956), which differs from the refined
// primer in that one position only.

struct mpb_track {
    int32_t win;    // index into the walk's window list
    int32_t state;  // 0 seed, 1 refine, 2 done
    int32_t first_cand, n_opt;
    int32_t n_trace, err;
    int64_t init, fm, rm, seed_cover, perfect;
    uint32_t allow[4];
    uint8_t seed[32], sets[32];
    int8_t opt_j[MPB_MAX_K], opt_cand[MPB_MAX_K];  // junction of option i; its candidate (offset from first_cand) or -1
    int64_t nn_cov[MPB_MAX_K];
    int64_t nn[MPB_MAX_K - 1][16];
};

MPB_HD int mpb_npos4(const int64_t* v) { return (v[0] > 0) + (v[1] > 0) + (v[2] > 0) + (v[3] > 0); }

// first index != skip of np.argsort(vals)[::-1] with a stable ascending sort: descending values, ties highest index first
MPB_HD int mpb_next_best(const int64_t* v, int skip) {
    int best = -1;
    for (int i = 3; i >= 0; --i) {
        if (i == skip) continue;
        if (best < 0 || v[i] > v[best]) best = i;
    }
    return best;
}

// core:579-593: max-sum path, first maximum wins.  freq[4][k], nn[k-1][16]
MPB_HD void mpb_viterbi(const int64_t* freq, const int64_t* nn, int k, uint8_t* path) {
    int64_t score[4];
    int8_t back[MPB_MAX_K][4];
    for (int b = 0; b < 4; ++b) score[b] = freq[b * k + 0];
    for (int t = 1; t < k; ++t) {
        int64_t nw[4];
        for (int cur = 0; cur < 4; ++cur) {
            int64_t best = 0;
            int arg = -1;
            for (int prev = 0; prev < 4; ++prev) {
                const int64_t val = score[prev] + nn[(t - 1) * 16 + prev * 4 + cur];
                if (arg < 0 || val > best) {
                    best = val;
                    arg = prev;
                }
            }
            nw[cur] = best + freq[cur * k + t];
            back[t][cur] = (int8_t)arg;
        }
        for (int b = 0; b < 4; ++b) score[b] = nw[b];
    }
    int cur = 0;
    for (int b = 1; b < 4; ++b)
        if (score[b] > score[cur]) cur = b;
    path[k - 1] = (uint8_t)cur;
    for (int t = k - 1; t >= 1; --t) {
        cur = back[t][cur];
        path[t - 1] = (uint8_t)cur;
    }
}

// the tracks of one window: NM (Viterbi seed) and, when it differs, MM (most frequent gap-free haplotype).
// Returns the number of tracks written to t0 / t1 (core:775-843).
MPB_HD int mpb_walk_seed(int win, int k, const int64_t* freq, const int64_t* nn, uint64_t mm_key, mpb_track* t0,
                         mpb_track* t1) {
    uint8_t nm[32], mm[32];
    for (int i = 0; i < 32; ++i) nm[i] = mm[i] = 0;
    mpb_viterbi(freq, nn, k, nm);
    const bool has_mm = mm_key != MPB_KEY_EMPTY;
    bool same = false;
    if (has_mm) {
        const uint64_t mask = (1ull << k) - 1ull;
        const uint64_t b0 = mm_key & mask, b1 = (mm_key >> k) & mask;
        same = true;
        for (int i = 0; i < k; ++i) {
            mm[i] = (uint8_t)(((b0 >> i) & 1ull) | (((b1 >> i) & 1ull) << 1));
            same = same && mm[i] == nm[i];
        }
    }
    const int nt = (has_mm && !same) ? 2 : 1;
    for (int ti = 0; ti < nt; ++ti) {
        mpb_track* t = ti == 0 ? t0 : t1;
        const uint8_t* sd = ti == 0 ? nm : mm;
        t->win = win;
        t->state = 0;
        t->first_cand = 0;
        t->n_opt = 0;
        t->n_trace = 0;
        t->err = 0;
        t->init = t->fm = t->rm = t->seed_cover = t->perfect = 0;
        for (int x = 0; x < 4; ++x) t->allow[x] = 0;
        for (int i = 0; i < 32; ++i) {
            t->seed[i] = i < k ? sd[i] : 0;
            t->sets[i] = i < k ? (uint8_t)(1u << sd[i]) : 0;
            if (i < k) t->allow[sd[i]] |= 1u << i;
        }
        for (int j = 0; j < k - 1; ++j) {
            for (int x = 0; x < 16; ++x) t->nn[j][x] = nn[j * 16 + x];
            t->nn_cov[j] = t->nn[j][sd[j] * 4 + sd[j + 1]];
        }
    }
    return nt;
}

// core:922-1080 for ONE junction j tied at the minimum NN coverage: which position gets which base.
// Returns false when the reference would leave the primer unchanged at this junction.
// pos / base: the refinement; kind: 0 position 0 (row merge in layer 0), 1 last position (column merge in layer j),
// 2 middle (column merge in layer jj, row merge in layer jj + 1), jj: the layer of a middle refinement.
MPB_HD bool mpb_walk_option(const mpb_track& t, int k, int j, int* pos, int* base, int* kind, int* jj_out) {
    const int last = k - 2;
    const int row = t.seed[j], col = t.seed[j + 1];
    const int64_t* L = t.nn[j];
    int jj = -1;
    if (j == 0) {
        int64_t column0[4];
        for (int x = 0; x < 4; ++x) column0[x] = L[x * 4 + col];
        if (mpb_npos4(column0) > 1) {
            *pos = 0;
            *base = mpb_next_best(column0, row);
            *kind = 0;
            *jj_out = 0;
            return true;
        }
        if (mpb_npos4(L + row * 4) > 1 && k > 2) jj = 0;
        else return false;
    } else if (j == last) {
        if (mpb_npos4(L + row * 4) > 1) {
            *pos = j + 1;
            *base = mpb_next_best(L + row * 4, col);
            *kind = 1;
            *jj_out = j;
            return true;
        }
        return false;
    } else {
        jj = j;
    }
    // middle rule on layers jj, jj + 1 (position jj + 1)
    const int mrow = t.seed[jj], mcol = t.seed[jj + 1], ncol = t.seed[jj + 2];
    const int64_t* L0 = t.nn[jj];
    const int64_t* L1 = t.nn[jj + 1];
    int64_t m[4];
    for (int x = 0; x < 4; ++x) m[x] = L0[mrow * 4 + x] < L1[x * 4 + ncol] ? L0[mrow * 4 + x] : L1[x * 4 + ncol];
    if (mpb_npos4(m) <= 1) return false;
    *pos = jj + 1;
    *base = mpb_next_best(m, mcol);
    *kind = 2;
    *jj_out = jj;
    return true;
}

// Emit the candidates of the next scan round for a live track.  out: room for up to k candidates; returns how many.
// State 0: the seed itself.  State 1: one candidate per junction tied at the minimal NN coverage whose refinement is
// defined (the refined primer, with the single-new-base trial folded in).
MPB_HD int mpb_walk_emit(mpb_track& t, int k, mpb_cand* out) {
    if (t.state == 0) {
        out[0].win = t.win;
        out[0].trial = -1;
        for (int x = 0; x < 4; ++x) out[0].allow[x] = t.allow[x];
        t.n_opt = 0;
        return 1;
    }
    int64_t lowest = t.nn_cov[0];
    for (int j = 1; j < k - 1; ++j) lowest = t.nn_cov[j] < lowest ? t.nn_cov[j] : lowest;
    int n = 0, no = 0;
    for (int j = 0; j < k - 1; ++j) {
        if (t.nn_cov[j] != lowest) continue;
        int pos, base, kind, jj;
        t.opt_j[no] = (int8_t)j;
        if (mpb_walk_option(t, k, j, &pos, &base, &kind, &jj)) {
            if (t.sets[pos] & (1u << base)) t.err = 1;  // the reference raises KeyError here
            const uint32_t bit = 1u << pos;
            out[n].win = t.win;
            out[n].trial = pos | (base << 8);
            for (int x = 0; x < 4; ++x) out[n].allow[x] = x == base ? (t.allow[x] | bit) : t.allow[x];
            t.opt_cand[no] = (int8_t)n;
            ++n;
        } else {
            t.opt_cand[no] = -1;
        }
        ++no;
    }
    t.n_opt = no;
    return n;
}

MPB_HD void mpb_walk_trace(mpb_track& t, uint8_t* trace /* [MPB_WALK_MAX_ROUNDS][32] of this track */) {
    if (t.n_trace < MPB_WALK_MAX_ROUNDS)
        for (int i = 0; i < 32; ++i) trace[t.n_trace * 32 + i] = t.sets[i];
    else
        t.err = 2;
    ++t.n_trace;
}

MPB_HD long long mpb_degeneracy(const uint8_t* sets, int k, int* ndeg) {
    const int fold[16] = {0, 1, 1, 2, 1, 2, 2, 3, 1, 2, 2, 3, 2, 3, 3, 4};
    long long d = 1;
    int n = 0;
    for (int i = 0; i < k; ++i) {
        d *= fold[sets[i] & 15];
        n += fold[sets[i] & 15] > 1;
        if (d > (1ll << 40)) d = 1ll << 40;
    }
    *ndeg = n;
    return d;
}

// Take the counts of the candidates emitted by mpb_walk_emit (counts[c*4 + {perfect, F_mis, R_mis, trial perfect}],
// c relative to the track's first candidate) and advance the track (core:881-906).  total = cover_number of the window.
MPB_HD void mpb_walk_consume(mpb_track& t, int k, const int64_t* c, int64_t total, int dnum, int degeneracy,
                             uint8_t* trace) {
    if (t.state == 0) {
        t.init = c[0];
        t.fm = c[1];
        t.rm = c[2];
        t.seed_cover = t.init;
        t.perfect = c[0];
        mpb_walk_trace(t, trace);
        t.state = (t.init + t.fm < total || t.init + t.rm < total) ? 1 : 2;
        return;
    }
    int best = 0;
    int64_t best_gain = 0;
    bool have = false;
    for (int oi = 0; oi < t.n_opt; ++oi) {
        int64_t gain = t.init;
        if (t.opt_cand[oi] >= 0) gain += c[t.opt_cand[oi] * 4 + 3];
        if (!have || gain > best_gain) {
            have = true;
            best_gain = gain;
            best = oi;
        }
    }
    bool changed = false;  // "NN coverage vector changed" (core:899)
    if (have && t.opt_cand[best] >= 0) {
        int pos, base, kind, jj;
        const int j = t.opt_j[best];
        mpb_walk_option(t, k, j, &pos, &base, &kind, &jj);
        const int64_t* cc = c + t.opt_cand[best] * 4;
        if (kind == 0) {  // position 0: row `base` of layer 0 merges into row seed[0]
            const int row = t.seed[0], col = t.seed[1];
            int64_t* L = t.nn[0];
            for (int y = 0; y < 4; ++y) {
                L[row * 4 + y] += L[base * 4 + y];
                L[base * 4 + y] = 0;
            }
            const int64_t nv = L[row * 4 + col];
            changed = nv != t.nn_cov[0];
            t.nn_cov[0] = nv;
        } else if (kind == 1) {  // last position: column `base` of layer j merges into column seed[j + 1]
            const int row = t.seed[j], col = t.seed[j + 1];
            int64_t* L = t.nn[j];
            for (int x = 0; x < 4; ++x) {
                L[x * 4 + col] += L[x * 4 + base];
                L[x * 4 + base] = 0;
            }
            const int64_t nv = L[row * 4 + col];
            changed = nv != t.nn_cov[j];
            t.nn_cov[j] = nv;
        } else {  // middle: column merge in layer jj, row merge in layer jj + 1
            const int row = t.seed[jj], col = t.seed[jj + 1], nrow = t.seed[jj + 1], ncol = t.seed[jj + 2];
            int64_t* L0 = t.nn[jj];
            int64_t* L1 = t.nn[jj + 1];
            for (int x = 0; x < 4; ++x) {
                L0[x * 4 + col] += L0[x * 4 + base];
                L0[x * 4 + base] = 0;
            }
            for (int y = 0; y < 4; ++y) {
                L1[nrow * 4 + y] += L1[base * 4 + y];
                L1[base * 4 + y] = 0;
            }
            const int64_t v0 = L0[row * 4 + col], v1 = L1[nrow * 4 + ncol];
            changed = v0 != t.nn_cov[jj] || v1 != t.nn_cov[jj + 1];
            t.nn_cov[jj] = v0;
            t.nn_cov[jj + 1] = v1;
        }
        t.sets[pos] |= (uint8_t)(1u << base);
        t.allow[base] |= 1u << pos;
        t.perfect = cc[0];
        t.fm = cc[1];
        t.rm = cc[2];
    }
    t.init = best_gain;
    mpb_walk_trace(t, trace);
    int ndeg = 0;
    const long long deg = mpb_degeneracy(t.sets, k, &ndeg);
    const int64_t mx = t.fm > t.rm ? t.fm : t.rm;
    if (mx == total) t.state = 2;
    else if (!changed) t.state = 2;
    else if (2 * deg > degeneracy || 3.0 * (double)deg / 2 > (double)degeneracy || ndeg == dnum) t.state = 2;
    else if (t.init + t.fm < total || t.init + t.rm < total) t.state = 1;
    else t.state = 2;
}

// core:816: the NM track only when its summed mismatch coverage is strictly larger
MPB_HD int mpb_walk_pick(const mpb_track& nm, const mpb_track& mm) {
    return ((nm.init + nm.fm) + (nm.init + nm.rm) > (mm.init + mm.fm) + (mm.init + mm.rm)) ? 0 : 1;
}
