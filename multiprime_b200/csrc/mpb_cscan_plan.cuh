// mpb_cscan_plan.cuh — plan of one candidate of the column scan (shared by k_cscan_plan and the fused walk round)
#pragma once
#include <stdint.h>

#include "mpb200.h"
#include "mpb_cscan.h"

// A plan turns a candidate into lists of column-plane rows (row = column * 4 + base):
//   hdr[0] n_tri   plain positions (one allowed base), three per step, padded with the all-ones row (never mismatches)
//   hdr[1] n_deg   entries of the degenerate positions (2..4 allowed bases each; bit 31 marks a position's last entry)
//   hdr[2] n_sf / hdr[3] n_sr   entries of the F- / R-strict positions (same format)
//   hdr[4] trial row or CSCAN_NONE, hdr[5] window index, hdr[6] offset of deg, hdr[7] offset of sf, sr follows
// tri region at word 8: [n_tri][4].
__device__ __forceinline__ void cscan_plan_one(int c, const mpb_cand* __restrict__ cands, const int32_t* __restrict__ win_pos,
                                               int nw, int k, uint32_t fmask, uint32_t rmask, uint32_t ones_row,
                                               uint32_t* __restrict__ plans, unsigned long long* __restrict__ counts,
                                               int zero_counts, int* __restrict__ err) {
    const mpb_cand cd = cands[c];
    uint32_t* P = plans + (size_t)c * CSCAN_PLAN_WORDS;
    if (zero_counts)
        for (int i = 0; i < 4; ++i) counts[(size_t)c * 4 + i] = 0;
    if (cd.win < 0 || cd.win >= nw) {
        atomicOr(err, MPB_ERR_BAD_CAND);
        P[0] = P[1] = P[2] = P[3] = 0;
        P[4] = CSCAN_NONE;
        P[5] = 0;
        return;
    }
    const uint32_t p = (uint32_t)win_pos[cd.win];
    int n1 = 0;
    uint32_t simple[MPB_MAX_K + 2];
    uint32_t* deg = P + 8 + 4 * ((MPB_MAX_K + 2) / 3);
    int nd = 0;
    for (int i = 0; i < k; ++i) {
        const uint32_t set = ((cd.allow[0] >> i) & 1u) | (((cd.allow[1] >> i) & 1u) << 1) | (((cd.allow[2] >> i) & 1u) << 2) |
                             (((cd.allow[3] >> i) & 1u) << 3);
        const int nb = __popc(set);
        if (nb == 1) {
            simple[n1++] = (p + i) * 4 + (__ffs(set) - 1);
        } else if (nb == 0) {
            // no base allowed: every row mismatches here — the all-zeros row never matches
            simple[n1++] = ones_row + 1;
        } else {
            uint32_t s = set;
            while (s) {
                const int b = __ffs(s) - 1;
                s &= s - 1;
                deg[nd++] = ((p + i) * 4 + b) | (s ? 0u : 0x80000000u);
            }
        }
    }
    const int ntri = (n1 + 2) / 3;
    for (int i = n1; i < ntri * 3; ++i) simple[i] = ones_row;
    for (int t = 0; t < ntri; ++t) {
        P[8 + t * 4 + 0] = simple[t * 3 + 0];
        P[8 + t * 4 + 1] = simple[t * 3 + 1];
        P[8 + t * 4 + 2] = simple[t * 3 + 2];
        P[8 + t * 4 + 3] = 0;
    }
    // the degenerate list was written behind the widest possible tri region; move it right behind the actual one
    uint32_t* dst = P + 8 + 4 * ntri;
    for (int i = 0; i < nd; ++i) dst[i] = deg[i];
    int off = 8 + 4 * ntri + nd;
    int ns[2];
    for (int side = 0; side < 2; ++side) {
        const uint32_t sm = side == 0 ? fmask : rmask;
        int n = 0;
        for (int i = 0; i < k; ++i) {
            if (!((sm >> i) & 1u)) continue;
            uint32_t set = ((cd.allow[0] >> i) & 1u) | (((cd.allow[1] >> i) & 1u) << 1) | (((cd.allow[2] >> i) & 1u) << 2) |
                           (((cd.allow[3] >> i) & 1u) << 3);
            if (set == 0) {
                P[off + n++] = (ones_row + 1) | 0x80000000u;
                continue;
            }
            while (set) {
                const int b = __ffs(set) - 1;
                set &= set - 1;
                P[off + n++] = ((p + i) * 4 + b) | (set ? 0u : 0x80000000u);
            }
        }
        ns[side] = n;
        off += n;
    }
    P[0] = (uint32_t)ntri;
    P[1] = (uint32_t)nd;
    P[2] = (uint32_t)ns[0];
    P[3] = (uint32_t)ns[1];
    P[4] = cd.trial >= 0 ? (p + (uint32_t)(cd.trial & 255)) * 4 + (uint32_t)((cd.trial >> 8) & 3) : CSCAN_NONE;
    P[5] = (uint32_t)cd.win;
    P[6] = (uint32_t)(8 + 4 * ntri);
    P[7] = (uint32_t)(8 + 4 * ntri + nd);
}
