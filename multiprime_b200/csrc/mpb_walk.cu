// mpb_walk.cu — the per-window control logic of multiPrime-core in native host code:
//   * mpb_walk: seeds (core:579-600), the NN-array refinement walk (core:860-1089) for all windows of a batch in
//     lock step with the candidate scan, NM-vs-MM choice (core:816); the scan itself is a callback (mpb_scan on the
//     GPU plus, in sequence-sharded runs, the count all-reduce), so the walk holds no device code,
//   * mpb_primer_props: Tm (k_tm on the device), GC / di-nucleotide / hairpin filters (core:387-416, 507-521).
// Everything here touches O(k) numbers per window; the O(sequences) work stays in the kernels.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <array>
#include <vector>

#include "mpb200.h"
#include "mpb_host.h"

#include "mpb_walk_core.h"

// Host driver of the shared walk (mpb_walk_core.h): all windows of a batch in lock step, one scan call per round.
// The product path on the GPU is mpb_walk_dev.cu (same core, state resident in HBM); this driver serves the CPU
// tests (scan = a stand-in) and any caller that wants to own the scan.
extern "C" int mpb_walk(int k, int v, int dnum, int degeneracy, int32_t n_win, const int64_t* cover_number,
                        const int64_t* freq, const int64_t* nn, const uint64_t* mm_key, mpb_scan_cb scan, void* user,
                        uint8_t* out_sets, int64_t* out_counts, uint8_t* out_seeds, int64_t* out_seed_cover,
                        int32_t* out_ntracks, int64_t trace_cap, uint8_t* trace_sets, int64_t* trace_off, int64_t* stats) {
    if (!cover_number || !freq || !nn || !mm_key || !scan || !out_sets || !out_counts || !out_seeds ||
        !out_seed_cover || !out_ntracks || !trace_off)
        return mpb_fail(MPB_EINVAL, "NULL argument");
    if (k < 3 || k > MPB_MAX_K || n_win < 0) return mpb_fail(MPB_EINVAL, "bad k or n_win");
    (void)v;
    std::vector<mpb_track> tracks((size_t)n_win * 2);
    std::vector<uint8_t> trace((size_t)n_win * 2 * MPB_WALK_MAX_ROUNDS * 32, 0);
    std::vector<int> ntr(n_win, 0);
    for (int w = 0; w < n_win; ++w) {
        ntr[w] = mpb_walk_seed(w, k, freq + (int64_t)w * 4 * k, nn + (int64_t)w * (k - 1) * 16, mm_key[w],
                               &tracks[2 * w], &tracks[2 * w + 1]);
        out_ntracks[w] = ntr[w];
        for (int ti = 0; ti < ntr[w]; ++ti) memcpy(out_seeds + ((int64_t)w * 2 + ti) * 32, tracks[2 * w + ti].seed, 32);
    }
    int64_t rounds = 0, cands_total = 0;
    std::vector<mpb_cand> cands;
    std::vector<int64_t> counts;
    mpb_cand buf[MPB_MAX_K];
    for (;;) {
        cands.clear();
        for (int w = 0; w < n_win; ++w)
            for (int ti = 0; ti < ntr[w]; ++ti) {
                mpb_track& t = tracks[2 * w + ti];
                if (t.state == 2) continue;
                t.first_cand = (int32_t)cands.size();
                const int n = mpb_walk_emit(t, k, buf);
                cands.insert(cands.end(), buf, buf + n);
                if (t.err == 1) return mpb_fail(MPB_EINVAL, "refinement would re-add a base (the reference raises KeyError)");
            }
        bool any = false;
        for (int w = 0; w < n_win && !any; ++w)
            for (int ti = 0; ti < ntr[w]; ++ti) any = any || tracks[2 * w + ti].state != 2;
        if (!any) break;
        const int64_t nc = (int64_t)cands.size();
        counts.assign((size_t)nc * 4 + 4, 0);
        if (nc > 0) {
            const int rc = scan(user, cands.data(), nc, counts.data());
            if (rc) return mpb_fail(rc, "scan callback failed");
        }
        ++rounds;
        cands_total += nc;
        for (int w = 0; w < n_win; ++w)
            for (int ti = 0; ti < ntr[w]; ++ti) {
                mpb_track& t = tracks[2 * w + ti];
                if (t.state == 2) continue;
                mpb_walk_consume(t, k, &counts[(size_t)t.first_cand * 4], cover_number[w], dnum, degeneracy,
                                 &trace[(size_t)(2 * w + ti) * MPB_WALK_MAX_ROUNDS * 32]);
                if (t.err == 2) return mpb_fail(MPB_EOVERFLOW, "more than %d refinement rounds in one window", MPB_WALK_MAX_ROUNDS);
            }
    }
    int64_t tr = 0;
    for (int w = 0; w < n_win; ++w) {
        trace_off[w] = tr;
        const mpb_track* a = &tracks[2 * w];
        const int pick = ntr[w] == 2 ? mpb_walk_pick(a[0], a[1]) : 0;
        const mpb_track& t = a[pick];
        memcpy(out_sets + (int64_t)w * 32, t.sets, 32);
        out_counts[w * 5 + 0] = t.init;
        out_counts[w * 5 + 1] = t.fm;
        out_counts[w * 5 + 2] = t.rm;
        out_counts[w * 5 + 3] = pick;
        out_counts[w * 5 + 4] = t.perfect;
        for (int ti = 0; ti < 2; ++ti) out_seed_cover[w * 2 + ti] = ti < ntr[w] ? a[ti].seed_cover : -1;
        for (int ti = 0; ti < ntr[w]; ++ti)
            for (int r = 0; r < a[ti].n_trace; ++r) {
                if (trace_sets && tr < trace_cap)
                    memcpy(trace_sets + tr * 32, &trace[((size_t)(2 * w + ti) * MPB_WALK_MAX_ROUNDS + r) * 32], 32);
                ++tr;
            }
    }
    trace_off[n_win] = tr;
    if (stats) {
        stats[0] = rounds;
        stats[1] = cands_total;
        stats[2] = tr;
    }
    if (trace_sets && tr > trace_cap) return mpb_fail(MPB_EOVERFLOW, "trace capacity %lld < %lld", (long long)trace_cap, (long long)tr);
    return 0;
}

namespace {
const int FOLD[16] = {0, 1, 1, 2, 1, 2, 2, 3, 1, 2, 2, 3, 2, 3, 3, 4};
int degeneracy_of(const uint8_t* sets, int k, int* ndeg) {
    long long d = 1;
    int n = 0;
    for (int i = 0; i < k; ++i) {
        d *= FOLD[sets[i] & 15];
        n += FOLD[sets[i] & 15] > 1;
        if (d > (1ll << 40)) d = 1ll << 40;
    }
    if (ndeg) *ndeg = n;
    return d > 0x7fffffff ? 0x7fffffff : (int)d;
}
}  // namespace

// ------------------------------------------------------------------------------------------------------
// properties of finished primers
// ------------------------------------------------------------------------------------------------------
namespace {

// Python round(x, 2): correctly rounded decimal -> nearest double.  Fast path when x*100 is clear of a tie.
double round2(double x) {
    const double y = x * 100.0;
    if (y > -1e13 && y < 1e13) {
        const long long fl = (long long)y - (y < (double)(long long)y ? 1 : 0);  // floor
        const double f = y - (double)fl;
        if (f < 0.5 - 1e-6) return (double)fl / 100.0;
        if (f > 0.5 + 1e-6) return (double)(fl + 1) / 100.0;
    }
    char buf[64];
    snprintf(buf, sizeof buf, "%.2f", x);  // glibc: exact, ties to even on the exact binary value
    return strtod(buf, nullptr);
}
double round3(double x) {
    char buf[64];
    snprintf(buf, sizeof buf, "%.3f", x);
    return strtod(buf, nullptr);
}

bool near_tie2(double x) {
    const double y = x * 100.0;
    return fabs((y - floor(y)) - 0.5) < 1e-6;
}

uint8_t comp_set(uint8_t s) { return (uint8_t)(((s & 1) << 3) | ((s & 2) << 1) | ((s & 4) >> 1) | ((s & 8) >> 3)); }

// core:410-416: some expansion contains XXXX, (XY)x4 or (XYZ)x3 (X != Y, Y != Z)
bool has_repeat(const uint8_t* sets, int k) {
    uint32_t allow[4] = {0, 0, 0, 0};
    for (int i = 0; i < k; ++i)
        for (int b = 0; b < 4; ++b)
            if ((sets[i] >> b) & 1) allow[b] |= 1u << i;
    auto test = [&](const int* pat, int n) {
        if (n > k) return false;
        uint32_t hit = (n == 32) ? 1u : ((1u << (k - n + 1)) - 1u);
        for (int t = 0; t < n && hit; ++t) hit &= allow[pat[t]] >> t;
        return hit != 0;
    };
    int pat[12];
    for (int i = 0; i < 4; ++i) {
        for (int t = 0; t < 4; ++t) pat[t] = i;
        if (test(pat, 4)) return true;
        for (int j = 0; j < 4; ++j) {
            if (i != j) {
                for (int t = 0; t < 8; ++t) pat[t] = (t & 1) ? j : i;
                if (test(pat, 8)) return true;
            }
            for (int kk = 0; kk < 4; ++kk)
                if (i != j && j != kk) {
                    for (int t = 0; t < 9; ++t) pat[t] = (t % 3 == 0) ? i : (t % 3 == 1 ? j : kk);
                    if (test(pat, 9)) return true;
                }
        }
    }
    return false;
}

// core:387-398: a 5-mer whose reverse complement can occur at least `distance` bases downstream
bool has_hairpin(const uint8_t* sets, int k, int distance) {
    for (int n = 0; n <= k - 5 - 5 - distance; ++n) {
        uint8_t target[5];
        for (int t = 0; t < 5; ++t) target[t] = comp_set(sets[n + 4 - t]);
        for (int o = n + 5 + distance; o + 5 <= k; ++o) {
            bool ok = true;
            for (int t = 0; t < 5 && ok; ++t) ok = (target[t] & sets[o + t]) != 0;
            if (ok) return true;
        }
    }
    return false;
}

}  // namespace

// core:666-687 on the host copy of the alignment (the device restatement is mpb_window_slow in mpb_device.cuh)
extern "C" int mpb_window_cells(const uint8_t* packed, int64_t row_stride, const int32_t* lens, int32_t n_col, int k,
                                int64_t n, const int64_t* seq, const int32_t* pos, uint8_t* cells, int32_t* out_len) {
    if (!packed || !seq || !pos || !cells || !out_len) return mpb_fail(MPB_EINVAL, "NULL argument");
    if (k < 1 || k > 32 || n < 0 || n_col < 0 || row_stride * 2 < n_col) return mpb_fail(MPB_EINVAL, "bad k, n or stride");
    for (int64_t i = 0; i < n; ++i) {
        const uint8_t* row = packed + seq[i] * row_stride;
        const int len = lens ? lens[seq[i]] : n_col;
        const int p = pos[i];
        if (p < 0 || len < 0 || len > n_col) return mpb_fail(MPB_EINVAL, "window %lld outside the row", (long long)i);
        auto cell = [&](int c) { return (row[c >> 1] >> (4 * (c & 1))) & 15; };
        uint8_t w[32], buf[32];
        int m = std::min(std::max(len - p, 0), k);  // cells the row holds inside the window
        bool allgap = true;
        for (int j = 0; j < m; ++j) {
            w[j] = (uint8_t)cell(p + j);
            allgap = allgap && w[j] == 0;
        }
        const int left_end = std::min(p, len);
        auto bases_left = [&](int g) {  // the last g bases left of the window, nearest first, into buf
            int got = 0;
            for (int c = left_end - 1; c >= 0 && got < g; --c)
                if (cell(c)) buf[got++] = (uint8_t)cell(c);
            return got == g;
        };
        if (!(m == k && allgap)) {
            if (m > 0 && w[0] == 0) {  // leading gap run
                int g = 0;
                while (g < m && w[g] == 0) ++g;
                if (bases_left(g))
                    for (int j = 0; j < g; ++j) w[j] = buf[g - 1 - j];
            }
            if (m > 0 && w[m - 1] == 0) {  // trailing gap run
                int g = 0;
                while (g < m && w[m - 1 - g] == 0) ++g;
                int got = 0;
                for (int c = p + k; c < len && got < g; ++c)
                    if (cell(c)) buf[got++] = (uint8_t)cell(c);
                if (got == g)
                    for (int j = 0; j < g; ++j) w[m - g + j] = buf[j];
            }
        }
        if (m < k) {  // row ends inside the window: extend to the left
            const int g = k - m;
            if (bases_left(g)) {
                for (int j = m - 1; j >= 0; --j) w[j + g] = w[j];
                for (int j = 0; j < g; ++j) w[j] = buf[g - 1 - j];
                m = k;
            }
        }
        uint8_t* o = cells + i * 32;
        for (int j = 0; j < 32; ++j) o[j] = j < m ? w[j] : 0;
        out_len[i] = m;
    }
    return 0;
}

// flags: 1 GC out of range, 2 di-nucleotide repeat, 4 hairpin, 64 Tm mean needs the exact host replay,
// 128 GC mean needs the exact host replay (a rounding tie could not be excluded in double arithmetic)
extern "C" int mpb_primer_props(mpb_ctx* ctx, const uint8_t* sets, int k, int32_t n, double gc_lo, double gc_hi,
                                int distance, const double* tm_consts3, double* tm_avg, double* gc, int32_t* flags,
                                int32_t* deg_out, int32_t* ndeg_out) {
    if (!ctx || !sets || !tm_consts3 || !tm_avg || !gc || !flags) return mpb_fail(MPB_EINVAL, "NULL argument");
    if (k < 3 || k > 32 || n < 0) return mpb_fail(MPB_EINVAL, "bad k or n");
    if (n == 0) return 0;
    std::vector<int64_t> degs(n);
    for (int i = 0; i < n; ++i) {
        int nd = 0;
        const int d = degeneracy_of(sets + (int64_t)i * 32, k, &nd);
        if (d > (1 << 20)) return mpb_fail(MPB_EEXPAND, "primer %d expands to more than 2^20 sequences", i);
        if (deg_out) deg_out[i] = d;
        if (ndeg_out) ndeg_out[i] = nd;
        degs[i] = d;
    }
    // per-expansion Tm, rounded and summed on the device (k_tm_sets)
    std::vector<int64_t> tm_sum(n);
    std::vector<int32_t> tm_ties(n);
    int rc = mpb_tm_sets(ctx, sets, k, n, tm_consts3, tm_sum.data(), tm_ties.data());
    if (rc) return rc;
    double gc_frac[33];  // round(g / k, 3) of core:405 for every possible G/C count
    for (int g = 0; g <= k; ++g) gc_frac[g] = round3((double)g / (double)k);
    for (int i = 0; i < n; ++i) {
        const uint8_t* S = sets + (int64_t)i * 32;
        int fl = 0;
        // Tm: mean over expansions of round(tm, 2), rounded to 2 (core:849-852); the device summed the rounded values
        // as integer hundredths
        {
            const double m = (double)((long double)tm_sum[i] / 100.0L / (long double)degs[i]);
            if (tm_ties[i] > 0 || near_tie2(m)) fl |= 64;
            tm_avg[i] = round2(m);
        }
        // GC: mean over expansions of round(gc/len, 3), rounded to 2 (core:401-407), via the GC-count distribution
        {
            uint64_t dist[34] = {1};  // dist[g]: expansions of the prefix with g G/C bases (< 4^27: exact)
            for (int p = 0; p < k; ++p) {
                const int s = S[p] & 15;
                const uint64_t n_gc = ((s >> 1) & 1) + ((s >> 2) & 1), n_at = (s & 1) + ((s >> 3) & 1);
                for (int g = p + 1; g >= 1; --g) dist[g] = dist[g] * n_at + dist[g - 1] * n_gc;
                dist[0] *= n_at;
            }
            long double tot = 0.0L, acc = 0.0L;
            for (int g = 0; g <= k; ++g) {
                tot += (long double)dist[g];
                if (dist[g] > 0) acc += (long double)dist[g] * (long double)gc_frac[g];
            }
            const double m = (double)(acc / tot);
            if (near_tie2(m)) fl |= 128;
            gc[i] = round2(m);
        }
        if (!(gc_lo <= gc[i] && gc[i] <= gc_hi)) fl |= 1;
        if (has_repeat(S, k)) fl |= 2;
        if (has_hairpin(S, k, distance)) fl |= 4;
        flags[i] = fl;
    }
    return 0;
}
