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

namespace {

const int FOLD[16] = {0, 1, 1, 2, 1, 2, 2, 3, 1, 2, 2, 3, 2, 3, 3, 4};
// expansion order of each base set (core:105-107): ORD[set] lists base indices A,C,G,T = 0..3
const int8_t ORD[16][4] = {{-1, -1, -1, -1}, {0, -1, -1, -1}, {1, -1, -1, -1}, {0, 1, -1, -1},  {2, -1, -1, -1}, {0, 2, -1, -1},
                           {2, 1, -1, -1},   {2, 0, 1, -1},   {3, -1, -1, -1}, {0, 3, -1, -1},  {1, 3, -1, -1},  {0, 3, 1, -1},
                           {2, 3, -1, -1},   {2, 0, 3, -1},   {2, 3, 1, -1},   {0, 3, 2, 1}};

struct Opt {
    bool valid;
    int pos, base;
    int nl;
    int lidx[2];
    int64_t layer[2][16];
    std::vector<int64_t> cov;
};

struct Track {
    int win;  // index into the batch
    int k;
    uint8_t seed[32];
    uint8_t sets[32];
    uint32_t allow[4];
    std::vector<std::array<int64_t, 16>> nn;
    std::vector<int64_t> nn_cov;
    int64_t init = 0, fm = 0, rm = 0, seed_cover = 0, perfect = 0;
    int state = 0;  // 0 seed, 1 refine, 2 done
    std::vector<Opt> opts;
    std::vector<std::array<uint8_t, 32>> trace;
    int first_cand = 0;  // index of this track's first candidate in the current round
};

int npos4(const int64_t* v) { return (v[0] > 0) + (v[1] > 0) + (v[2] > 0) + (v[3] > 0); }

// np.argsort(vals)[::-1] for a stable ascending sort: descending values, ties highest index first
void order_desc(const int64_t* v, int* out) {
    int idx[4] = {0, 1, 2, 3};
    std::stable_sort(idx, idx + 4, [&](int a, int b) { return v[a] < v[b]; });
    for (int i = 0; i < 4; ++i) out[i] = idx[3 - i];
}

int first_not(const int* ord, int skip) {
    for (int i = 0; i < 4; ++i)
        if (ord[i] != skip) return ord[i];
    return -1;
}

// core:579-593: max-sum path, first maximum wins
void viterbi(const int64_t* freq /*[4][k]*/, const int64_t* nn /*[k-1][16]*/, int k, uint8_t* path) {
    int64_t score[4];
    std::vector<std::array<int8_t, 4>> back(k);
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
        memcpy(score, nw, sizeof nw);
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

// core:922-1080: the refinement the reference would try at every junction tied at the minimum NN coverage
void refine_options(Track& t) {
    const int k = t.k, last = k - 2;
    t.opts.clear();
    int64_t lowest = t.nn_cov[0];
    for (int j = 1; j < k - 1; ++j) lowest = std::min(lowest, t.nn_cov[j]);
    for (int j = 0; j < k - 1; ++j) {
        if (t.nn_cov[j] != lowest) continue;
        const int row = t.seed[j], col = t.seed[j + 1];
        const int64_t* L = t.nn[j].data();
        Opt o;
        o.valid = false;
        auto middle = [&](int jj) {
            const int nrow = t.seed[jj + 1], ncol = t.seed[jj + 2];
            const int64_t* L0 = t.nn[jj].data();
            const int64_t* L1 = t.nn[jj + 1].data();
            int64_t m[4];
            for (int x = 0; x < 4; ++x) m[x] = std::min(L0[row * 4 + x], L1[x * 4 + ncol]);
            if (npos4(m) <= 1) return;
            int ord[4];
            order_desc(m, ord);
            const int idx = first_not(ord, col);
            o.valid = true;
            o.pos = jj + 1;
            o.base = idx;
            o.nl = 2;
            o.lidx[0] = jj;
            o.lidx[1] = jj + 1;
            memcpy(o.layer[0], L0, sizeof o.layer[0]);
            memcpy(o.layer[1], L1, sizeof o.layer[1]);
            for (int x = 0; x < 4; ++x) {
                o.layer[0][x * 4 + col] += L0[x * 4 + idx];
                o.layer[0][x * 4 + idx] = 0;
            }
            for (int y = 0; y < 4; ++y) {
                o.layer[1][nrow * 4 + y] += L1[idx * 4 + y];
                o.layer[1][idx * 4 + y] = 0;
            }
            o.cov = t.nn_cov;
            o.cov[jj] = o.layer[0][row * 4 + col];
            o.cov[jj + 1] = o.layer[1][nrow * 4 + ncol];
        };
        if (j == 0) {
            int64_t column0[4];
            for (int x = 0; x < 4; ++x) column0[x] = L[x * 4 + col];
            if (npos4(column0) > 1) {  // position 0
                int ord[4];
                order_desc(column0, ord);
                const int idx = first_not(ord, row);
                o.valid = true;
                o.pos = 0;
                o.base = idx;
                o.nl = 1;
                o.lidx[0] = 0;
                memcpy(o.layer[0], L, sizeof o.layer[0]);
                for (int y = 0; y < 4; ++y) {
                    o.layer[0][row * 4 + y] += L[idx * 4 + y];
                    o.layer[0][idx * 4 + y] = 0;
                }
                o.cov = t.nn_cov;
                o.cov[0] = o.layer[0][row * 4 + col];
            } else if (npos4(L + row * 4) > 1) {
                middle(0);
            }
        } else if (j == last) {
            if (npos4(L + row * 4) > 1) {
                int ord[4];
                order_desc(L + row * 4, ord);
                const int idx = first_not(ord, col);
                o.valid = true;
                o.pos = j + 1;
                o.base = idx;
                o.nl = 1;
                o.lidx[0] = j;
                memcpy(o.layer[0], L, sizeof o.layer[0]);
                for (int x = 0; x < 4; ++x) {
                    o.layer[0][x * 4 + col] += L[x * 4 + idx];
                    o.layer[0][x * 4 + idx] = 0;
                }
                o.cov = t.nn_cov;
                o.cov[j] = o.layer[0][row * 4 + col];
            }
        } else {
            middle(j);
        }
        t.opts.push_back(std::move(o));
    }
}

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

void push_trace(Track& t) {
    std::array<uint8_t, 32> a{};
    memcpy(a.data(), t.sets, 32);
    t.trace.push_back(a);
}

}  // namespace

extern "C" int mpb_walk(int k, int v, int dnum, int degeneracy, uint32_t fmask, uint32_t rmask, int32_t n_win,
                        const int32_t* win_pos, const int64_t* cover_number, const int64_t* freq, const int64_t* nn,
                        const uint64_t* mm_key, mpb_scan_cb scan, void* user, uint8_t* out_sets, int64_t* out_counts,
                        uint8_t* out_seeds, int64_t* out_seed_cover, int32_t* out_ntracks, int64_t trace_cap,
                        uint8_t* trace_sets, int64_t* trace_off, int64_t* stats) {
    if (!win_pos || !cover_number || !freq || !nn || !mm_key || !scan || !out_sets || !out_counts || !out_seeds ||
        !out_seed_cover || !out_ntracks || !trace_off)
        return mpb_fail(MPB_EINVAL, "NULL argument");
    if (k < 3 || k > MPB_MAX_K || n_win < 0) return mpb_fail(MPB_EINVAL, "bad k or n_win");
    std::vector<Track> tracks;
    tracks.reserve((size_t)n_win * 2);
    std::vector<int> first_track(n_win + 1, 0);
    for (int w = 0; w < n_win; ++w) {
        first_track[w] = (int)tracks.size();
        uint8_t nm[32] = {0}, mm[32] = {0};
        viterbi(freq + (int64_t)w * 4 * k, nn + (int64_t)w * (k - 1) * 16, k, nm);
        bool has_mm = mm_key[w] != MPB_KEY_EMPTY;
        bool same = false;
        if (has_mm) {
            const uint64_t key = mm_key[w];
            const uint64_t mask = (1ull << k) - 1ull;
            const uint64_t b0 = key & mask, b1 = (key >> k) & mask;
            same = true;
            for (int i = 0; i < k; ++i) {
                mm[i] = (uint8_t)(((b0 >> i) & 1ull) | (((b1 >> i) & 1ull) << 1));
                same = same && mm[i] == nm[i];
            }
        }
        const int nt = (has_mm && !same) ? 2 : 1;
        out_ntracks[w] = nt;
        for (int ti = 0; ti < nt; ++ti) {
            Track t;
            t.win = w;
            t.k = k;
            memset(t.seed, 0, 32);
            memset(t.sets, 0, 32);
            memcpy(t.seed, ti == 0 ? nm : mm, k);
            memset(t.allow, 0, sizeof t.allow);
            for (int i = 0; i < k; ++i) {
                t.sets[i] = (uint8_t)(1u << t.seed[i]);
                t.allow[t.seed[i]] |= 1u << i;
            }
            t.nn.resize(k - 1);
            t.nn_cov.resize(k - 1);
            for (int j = 0; j < k - 1; ++j) {
                memcpy(t.nn[j].data(), nn + ((int64_t)w * (k - 1) + j) * 16, 16 * sizeof(int64_t));
                t.nn_cov[j] = t.nn[j][t.seed[j] * 4 + t.seed[j + 1]];
            }
            memcpy(out_seeds + ((int64_t)w * 2 + ti) * 32, t.seed, 32);
            tracks.push_back(std::move(t));
        }
    }
    first_track[n_win] = (int)tracks.size();
    std::vector<int> live(tracks.size());
    for (size_t i = 0; i < tracks.size(); ++i) live[i] = (int)i;
    int64_t rounds = 0, cands_total = 0;
    std::vector<int32_t> cpos, cpos_sorted;
    std::vector<uint32_t> callow, callow_sorted;
    std::vector<int64_t> counts, counts_sorted;
    std::vector<int> order;
    while (!live.empty()) {
        cpos.clear();
        callow.clear();
        for (int ti : live) {
            Track& t = tracks[ti];
            t.first_cand = (int)cpos.size();
            const int32_t pos = win_pos[t.win];
            if (t.state == 0) {
                cpos.push_back(pos);
                callow.insert(callow.end(), t.allow, t.allow + 4);
            } else {
                refine_options(t);
                for (const Opt& o : t.opts) {
                    if (!o.valid) continue;
                    if (t.sets[o.pos] & (1u << o.base))
                        return mpb_fail(MPB_EINVAL, "refinement would re-add a base (the reference raises KeyError)");
                    const uint32_t bit = 1u << o.pos;
                    cpos.push_back(pos);  // primer with position pos := base alone (coverage_renew look-up)
                    for (int x = 0; x < 4; ++x) callow.push_back(x == o.base ? (t.allow[x] | bit) : (t.allow[x] & ~bit));
                    cpos.push_back(pos);  // primer with the base added
                    for (int x = 0; x < 4; ++x) callow.push_back(x == o.base ? (t.allow[x] | bit) : t.allow[x]);
                }
            }
        }
        const int64_t nc = (int64_t)cpos.size();
        order.resize(nc);
        for (int64_t i = 0; i < nc; ++i) order[i] = (int)i;
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cpos[a] < cpos[b]; });
        cpos_sorted.resize(nc);
        callow_sorted.resize(nc * 4);
        for (int64_t i = 0; i < nc; ++i) {
            cpos_sorted[i] = cpos[order[i]];
            memcpy(&callow_sorted[i * 4], &callow[(int64_t)order[i] * 4], 16);
        }
        counts_sorted.assign(nc * 3, 0);
        if (nc > 0) {
            const int rc = scan(user, cpos_sorted.data(), callow_sorted.data(), nc, counts_sorted.data());
            if (rc) return mpb_fail(rc, "scan callback failed");
        }
        counts.resize(nc * 3);
        for (int64_t i = 0; i < nc; ++i) memcpy(&counts[(int64_t)order[i] * 3], &counts_sorted[i * 3], 24);
        ++rounds;
        cands_total += nc;
        std::vector<int> next;
        for (int ti : live) {
            Track& t = tracks[ti];
            const int64_t total = cover_number[t.win];
            const int64_t* c = &counts[(int64_t)t.first_cand * 3];
            if (t.state == 0) {
                t.init = c[0];
                t.fm = c[1];
                t.rm = c[2];
                t.seed_cover = t.init;
                t.perfect = c[0];  // expansion rows matching the primer exactly (core:853 perfect_coverage)
                push_trace(t);
                t.state = 1;
                if (t.init + t.fm < total || t.init + t.rm < total) next.push_back(ti);
                else t.state = 2;
                continue;
            }
            int best = 0, ci = 0, best_ci = -1;
            int64_t best_gain = 0;
            bool have = false;
            for (size_t oi = 0; oi < t.opts.size(); ++oi) {
                int64_t gain = t.init;
                int my_ci = -1;
                if (t.opts[oi].valid) {
                    gain += c[ci * 3 + 0];
                    my_ci = ci;
                    ci += 2;
                }
                if (!have || gain > best_gain) {
                    have = true;
                    best_gain = gain;
                    best = (int)oi;
                    best_ci = my_ci;
                }
            }
            const Opt& o = t.opts[best];
            std::vector<int64_t> cov_new = t.nn_cov;
            if (o.valid) {
                t.sets[o.pos] |= (uint8_t)(1u << o.base);
                t.allow[o.base] |= 1u << o.pos;
                for (int l = 0; l < o.nl; ++l) memcpy(t.nn[o.lidx[l]].data(), o.layer[l], sizeof o.layer[l]);
                cov_new = o.cov;
                t.perfect = c[(best_ci + 1) * 3 + 0];
                t.fm = c[(best_ci + 1) * 3 + 1];
                t.rm = c[(best_ci + 1) * 3 + 2];
            }
            t.init = best_gain;
            push_trace(t);
            int ndeg = 0;
            const long long deg = degeneracy_of(t.sets, k, &ndeg);
            if (std::max(t.fm, t.rm) == total) {
                t.state = 2;
            } else if (cov_new == t.nn_cov) {
                t.state = 2;
            } else if (2 * deg > degeneracy || 3.0 * deg / 2 > degeneracy || ndeg == dnum) {
                t.state = 2;
            } else {
                t.nn_cov = cov_new;
                if (t.init + t.fm < total || t.init + t.rm < total) next.push_back(ti);
                else t.state = 2;
            }
        }
        live.swap(next);
    }
    // choose the track (core:816: NM only when strictly better), emit results and the call trace
    int64_t tr = 0;
    for (int w = 0; w < n_win; ++w) {
        trace_off[w] = tr;
        const int a = first_track[w], nt = first_track[w + 1] - first_track[w];
        int pick = a;
        if (nt == 2) {
            const Track &nmt = tracks[a], &mmt = tracks[a + 1];
            pick = ((nmt.init + nmt.fm) + (nmt.init + nmt.rm) > (mmt.init + mmt.fm) + (mmt.init + mmt.rm)) ? a : a + 1;
        }
        const Track& t = tracks[pick];
        memcpy(out_sets + (int64_t)w * 32, t.sets, 32);
        out_counts[w * 5 + 0] = t.init;
        out_counts[w * 5 + 1] = t.fm;
        out_counts[w * 5 + 2] = t.rm;
        out_counts[w * 5 + 3] = pick - a;
        out_counts[w * 5 + 4] = t.perfect;
        for (int ti = 0; ti < 2; ++ti) out_seed_cover[w * 2 + ti] = ti < nt ? tracks[a + ti].seed_cover : -1;
        for (int ti = 0; ti < nt; ++ti)
            for (const auto& s : tracks[a + ti].trace) {
                if (trace_sets && tr < trace_cap) memcpy(trace_sets + tr * 32, s.data(), 32);
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
    // all expansions of all primers, product order (leftmost position slowest), as bases 0..3
    std::vector<int64_t> off(n + 1, 0);
    for (int i = 0; i < n; ++i) {
        int nd = 0;
        const int d = degeneracy_of(sets + (int64_t)i * 32, k, &nd);
        if (d > (1 << 20)) return mpb_fail(MPB_EEXPAND, "primer %d expands to more than 2^20 sequences", i);
        if (deg_out) deg_out[i] = d;
        if (ndeg_out) ndeg_out[i] = nd;
        off[i + 1] = off[i] + d;
    }
    std::vector<uint8_t> seqs((size_t)off[n] * k);
    for (int i = 0; i < n; ++i) {
        // odometer over the positions, rightmost fastest: each expansion is its predecessor with a changed suffix
        const uint8_t* S = sets + (int64_t)i * 32;
        const int64_t d = off[i + 1] - off[i];
        int digit[32];
        uint8_t* q = &seqs[(size_t)off[i] * k];
        for (int p = 0; p < k; ++p) {
            digit[p] = 0;
            q[p] = (uint8_t)ORD[S[p] & 15][0];
        }
        for (int64_t e = 1; e < d; ++e) {
            uint8_t* nx = q + k;
            memcpy(nx, q, (size_t)k);
            for (int p = k - 1; p >= 0; --p) {
                const int code = S[p] & 15;
                if (++digit[p] < FOLD[code]) {
                    nx[p] = (uint8_t)ORD[code][digit[p]];
                    break;
                }
                digit[p] = 0;
                nx[p] = (uint8_t)ORD[code][0];
            }
            q = nx;
        }
    }
    std::vector<double> tm(off[n]);
    int rc = mpb_tm(ctx, seqs.data(), k, off[n], tm_consts3, tm.data(), nullptr, nullptr);
    if (rc) return rc;
    double gc_frac[33];  // round(g / k, 3) of core:405 for every possible G/C count
    for (int g = 0; g <= k; ++g) gc_frac[g] = round3((double)g / (double)k);
    for (int i = 0; i < n; ++i) {
        const uint8_t* S = sets + (int64_t)i * 32;
        int fl = 0;
        // Tm: mean over expansions of round(tm, 2), rounded to 2 (core:849-852)
        {
            long double acc = 0.0L;
            const int64_t d = off[i + 1] - off[i];
            for (int64_t e = off[i]; e < off[i + 1]; ++e) acc += (long double)round2(tm[e]);
            const double m = (double)(acc / (long double)d);
            if (near_tie2(m)) fl |= 64;
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
