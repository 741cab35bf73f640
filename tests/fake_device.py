"""A pure-Python stand-in for multiprime_b200._lib (Context / Msa / Hist), built on oracle primitives.

TEST INFRASTRUCTURE ONLY: it lets the CPU test-suite drive the host logic of multiprime_b200.core (gates, seeds,
refinement walk, filters, writers) without a GPU, and documents the contract of every libmpb200 entry point
(key encoding, count semantics).  The GPU tests run the same cases against the real library.
"""
from __future__ import annotations

import numpy as np

from multiprime_b200.iupac import BASES, CODE_CHARS
from oracle import mp_oracle as o

KEY_EMPTY = 0xFFFFFFFFFFFFFFFF
KEY_IUPAC = 0xFFFFFFFFFFFFFFFE
KEY_BASE5 = 1 << 54
MAX_K = 27


def hap_key(hap: str) -> int:
    """table key of a plain haplotype string (ACGT-)"""
    k = len(hap)
    if "-" not in hap:
        b0 = b1 = 0
        for i, ch in enumerate(hap):
            b = BASES.index(ch)
            b0 |= (b & 1) << i
            b1 |= (b >> 1) << i
        return b0 | (b1 << k)
    return KEY_BASE5 + sum("ACGT-".index(ch) * 5 ** i for i, ch in enumerate(hap))


class Context:
    def __init__(self, device=0, stream=None):
        self.launches = 0

    @classmethod
    def shared(cls, device=0, stream=None):
        return cls(device, stream)

    def tm(self, seqs2bit, consts3, want_hs=False):
        tm, dh, ds = [], [], []
        for row in np.asarray(seqs2bit):
            s = "".join(BASES[b] for b in row)
            tm.append(o.tm_unrounded(s))
            h, e = o.delta_h_s(s)
            dh.append(h)
            ds.append(e)
        if want_hs:
            return np.array(tm), np.array(dh), np.array(ds)
        return np.array(tm)

    def primer_props(self, sets_arr, k, gc_lo, gc_hi, distance, consts3):
        from statistics import mean
        n = len(sets_arr)
        tm, gc = np.zeros(n), np.zeros(n)
        flags, deg, ndeg = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.int32)
        for i in range(n):
            p = "".join(CODE_CHARS[c] for c in sets_arr[i, :k])
            tm[i] = round(mean([o.tm(e) for e in o.expand(p)]), 2)
            gc[i] = o.gc_content(p)
            flags[i] = (0 if gc_lo <= gc[i] <= gc_hi else 1) | (2 if o.has_repeat(p) else 0) | \
                       (4 if o.has_hairpin(p, distance) else 0)
            deg[i], ndeg[i] = o.degeneracy(p), o.n_degenerate(p)
        return tm, gc, flags, deg, ndeg

    def pair_cover(self, uf, ur, pf, pr):
        return np.array([int(np.unpackbits((uf[a] | ur[b]).view(np.uint8)).sum()) for a, b in zip(pf, pr)], np.int32)

    def pair_cover3(self, bits, pf, pr):
        bits = np.asarray(bits)
        return np.array([int(np.unpackbits((bits[a, 0] | bits[a, 2] | bits[b, 1] | bits[b, 2]).view(np.uint8)).sum())
                         for a, b in zip(pf, pr)], np.int32)

    def dimer_flags(self, sets_list):
        return np.array([o.self_dimer("".join(CODE_CHARS[c] for c in s)) for s in sets_list], bool)

    def sync(self):
        pass

    def close(self):
        pass


class Msa:
    def __init__(self, ctx, packed4, n_seq, n_col, row_bytes=None, lens=None):
        self.ctx = ctx
        self.n_seq, self.n_col = n_seq, n_col
        codes = np.empty((n_seq, packed4.shape[1] * 2), np.uint8)
        codes[:, 0::2] = packed4 & 15
        codes[:, 1::2] = packed4 >> 4
        lens = np.full(n_seq, n_col) if lens is None else lens
        self.rows = ["".join(CODE_CHARS[c] for c in codes[i, :lens[i]]) for i in range(n_seq)]
        self.row0 = 0

    def close(self):
        pass

    def set_row0(self, row0):
        self.row0 = row0

    def seq_attr(self):
        lead = np.array([len(s) - len(s.lstrip("-")) for s in self.rows], np.int32)
        rstrip = np.array([len(s.rstrip("-")) for s in self.rows], np.int32)
        return lead, rstrip

    # per-column code patterns of the bit-sliced prefilter (mpb_prefilter.cu BS_LO / BS_HI)
    BS_LO = [0x414, 0x1900, 0x32, 0x13, 0xc40, 0xc4, 0x602, 0x1088, 0x4a0, 0xd, 0x1401, 0x409, 0x1a0, 0x248, 0x1204, 0x184,
             0x100c, 0x1028, 0x1104, 0x1018, 0x58, 0x1006, 0x118, 0x881, 0xc8, 0x482, 0x504]
    BS_HI = [0x608, 0x62, 0x1110, 0x811, 0x1802, 0xa04, 0x43, 0x320, 0x640, 0x86, 0x1a00, 0x40a, 0x809, 0x222, 0xa40, 0xe0,
             0x806, 0x29, 0x1300, 0x501, 0x221, 0x841, 0x211, 0x1044, 0x460, 0x484, 0x920]

    def prefilter(self, k, v, win_pos, code="bs"):
        """coarse view of every counted item (mpb_window_prefilter): code="bs" the 13-bit GF(2)-linear code of the
        bit-sliced kernel (mpb_prefilter.cu), code="row" the 16-bit multiplicative hash of the row-domain kernel"""
        s0, s1 = np.zeros(len(win_pos)), np.zeros(len(win_pos))
        low = {c: ("A" if i & 1 else "C" if i & 2 else "G" if i & 4 else "T") for i, c in enumerate(CODE_CHARS) if i}
        low["-"] = "A"
        for wi, p in enumerate(win_pos):
            bins = {}
            for s in self.rows:
                w = o.window_kmer(s, int(p), k)
                items = ["".join(low[ch] for ch in w)] if w.count("-") > v else [e.replace("-", "A") for e in o.expand(w)]
                for it in items:
                    if code == "bs":
                        c = 0
                        for j, ch in enumerate(it):
                            if ch in "CT":
                                c ^= self.BS_LO[j]
                            if ch in "GT":
                                c ^= self.BS_HI[j]
                    else:
                        lo = sum(1 << j for j, ch in enumerate(it) if ch in "CT")       # pre_code() of mpb200.cu
                        hi = sum(1 << j for j, ch in enumerate(it) if ch in "GT")
                        x = (lo ^ ((hi << 7) & 0xFFFFFFFF) ^ (hi >> 9)) & 0xFFFFFFFF
                        c = ((x * 0x9E3779B1) & 0xFFFFFFFF) >> 16
                    bins[c] = bins.get(c, 0) + 1
            cs = np.array(list(bins.values()), float)
            s0[wi], s1[wi] = cs.sum(), (cs * np.log2(cs)).sum()
        return s0, s1

    def seq_attr_hist(self):
        lead, rstrip = self.seq_attr()
        return (np.bincount(lead, minlength=self.n_col + 1).astype(np.int64),
                np.bincount(rstrip, minlength=self.n_col + 1).astype(np.int64))

    def hist(self, k, v, win_pos, log2_cap=0, empty=False):
        return Hist(self, k, v, win_pos, empty)

    def scan(self, k, v, fmask, rmask, cand_pos, cand_allow, bits_slot=None, counts_out=None, bits_out=None):
        cand_allow = np.asarray(cand_allow).reshape(-1, 4)
        nc = len(cand_pos)
        counts = np.zeros((nc, 3), np.int64)
        words = (self.n_seq + 31) // 32
        bits = None
        if bits_slot is not None:
            bits = np.zeros((int(max(bits_slot)) + 1 if nc else 0, 3, words), np.uint32)
        cache = {}
        for ci in range(nc):
            p = int(cand_pos[ci])
            if p not in cache:
                cache[p] = [o.window_kmer(s, p, k) for s in self.rows]
            allow = [int(x) for x in cand_allow[ci]]
            for si, w in enumerate(cache[p]):
                assert len(w) == k
                isgap = w.count("-") > v
                non_f = non_r = False
                if not isgap:
                    for hap in o.expand(w):
                        m = 0
                        for i, ch in enumerate(hap):
                            if ch == "-" or not (allow[BASES.index(ch)] >> i) & 1:
                                m |= 1 << i
                        within = bin(m).count("1") <= v
                        okf = within and not (m & fmask)
                        okr = within and not (m & rmask)
                        counts[ci, 0] += m == 0
                        counts[ci, 1] += okf and m != 0
                        counts[ci, 2] += okr and m != 0
                        non_f |= not okf
                        non_r |= not okr
                if bits is not None and bits_slot[ci] >= 0:
                    sl = bits_slot[ci]
                    for j, flag in enumerate((non_f, non_r, isgap)):
                        if flag:
                            bits[sl, j, si >> 5] |= np.uint32(1 << (si & 31))
        return counts, bits

    def pattern_hits(self, allow, lens, max_hits=1 << 20):
        """mpb_pattern_hits: exact occurrences of degenerate patterns; only single-base cells can match"""
        single = {"A": 0, "C": 1, "G": 2, "T": 3}
        allow = np.asarray(allow).reshape(-1, 4)
        out = []
        for p, (al, ln) in enumerate(zip(allow, lens)):
            for r, row in enumerate(self.rows):
                for x in range(0, len(row) - int(ln) + 1):
                    if all(row[x + i] in single and (int(al[single[row[x + i]]]) >> i) & 1 for i in range(int(ln))):
                        out.append((p, r, x))
        out.sort()
        a = np.array(out, np.int32).reshape(-1, 3)
        return a[:, 0], a[:, 1], a[:, 2]

    def seqkeys(self, k, win_pos):
        out = np.empty((len(win_pos), self.n_seq), np.uint64)
        for wi, p in enumerate(win_pos):
            for si, s in enumerate(self.rows):
                w = o.window_kmer(s, int(p), k)
                out[wi, si] = KEY_IUPAC if any(ch not in "ACGT-" for ch in w) else hap_key(w)
        return out


class Hist:
    def __init__(self, msa, k, v, win_pos, empty=False):
        self.msa, self.k, self.v = msa, k, v
        self.win_pos = [int(p) for p in win_pos]
        self.nw = len(self.win_pos)
        self.tables = []      # per window: {key: [count, first]}
        self.gap_n = []
        self.exc = []
        self.n_iupac_gap = []
        if empty:             # owner tables of a sharded run: filled through merge_segments / add_counts
            self.tables = [{} for _ in self.win_pos]
            self.gap_n = [0] * self.nw
            self.n_iupac_gap = [0] * self.nw
            return
        for wi, p in enumerate(self.win_pos):
            tab, gaps, nig = {}, 0, 0
            for si, s in enumerate(msa.rows):
                w = o.window_kmer(s, p, k)
                assert len(w) == k
                iupac = any(ch not in "ACGT-" for ch in w)
                if w.count("-") > v:
                    gaps += 1
                    if iupac:
                        nig += 1
                        self.exc.append((wi, si))
                    else:
                        e = tab.setdefault(hap_key(w), [0, (msa.row0 + si) << 16])
                        e[0] += 1
                else:
                    for ei, hap in enumerate(o.expand(w)):
                        e = tab.setdefault(hap_key(hap), [0, ((msa.row0 + si) << 16) | ei])
                        e[0] += 1
            self.tables.append(tab)
            self.gap_n.append(gaps)
            self.n_iupac_gap.append(nig)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        pass

    def close(self):
        pass

    def _is_cover(self, key):
        if key < KEY_BASE5:
            return True
        x, g = key - KEY_BASE5, 0
        for _ in range(self.k):
            g += x % 5 == 4
            x //= 5
        return g <= self.v

    def counts(self):
        return (np.array(self.gap_n, np.int64), np.array(self.n_iupac_gap, np.int64),
                np.array([len(t) for t in self.tables], np.int64))

    def add_counts(self, gap_n, n_iupac_gap):
        self.gap_n = [int(x) for x in gap_n]
        self.n_iupac_gap = [int(x) for x in n_iupac_gap]

    def summary(self):
        out = self.stats()
        out["freq"], out["nn"] = self.tensors(np.ones(self.nw, np.uint8))
        return out

    def export_at(self, order, counts, comm=None):
        keys, cnt, first = [], [], []
        for wi, n in zip(order, counts):
            tab = self.tables[int(wi)]
            assert n in (0, len(tab))
            if n:
                for key, (c, f) in tab.items():
                    keys.append(key)
                    cnt.append(c)
                    first.append(f)
        return np.array(keys, np.uint64), np.array(cnt, np.uint32), np.array(first, np.uint64)

    def merge_segments(self, seg_off, keys, cnt, first):
        for s in range(len(seg_off) - 1):
            wi = s % self.nw
            for i in range(int(seg_off[s]), int(seg_off[s + 1])):
                e = self.tables[wi].setdefault(int(keys[i]), [0, int(first[i])])
                e[0] += int(cnt[i])
                e[1] = min(e[1], int(first[i]))

    def cscan(self, fmask, rmask, cands, bits_slot=None, counts_out=None, bits_out=None):
        """mpb_cscan: mpb_scan's three counts plus the perfect matches that carry the trial base"""
        nc = len(cands)
        pos = [self.win_pos[int(w)] for w in cands["win"]]
        counts3, bits = self.msa.scan(self.k, self.v, fmask, rmask, pos, cands["allow"], bits_slot=bits_slot)
        counts = np.zeros((nc, 4), np.int64)
        counts[:, :3] = counts3
        for ci in range(nc):
            tr = int(cands["trial"][ci])
            if tr < 0:
                continue
            tp, tb = tr & 255, (tr >> 8) & 3
            allow = [int(x) & ~(1 << tp) for x in cands["allow"][ci]]
            allow[tb] |= 1 << tp
            c1, _ = self.msa.scan(self.k, self.v, fmask, rmask, [pos[ci]], np.array([allow], np.uint32))
            counts[ci, 3] = c1[0, 0]
        return counts, bits

    def walk(self, dnum, degeneracy, fmask, rmask, win_idx, cover_number, mm_key, freq=None, nn=None, comm=None,
             want_trace=True, lag=2):
        """the host driver of the shared walk (mpb_walk) over this stand-in's scan"""
        from multiprime_b200 import _lib
        win_idx = np.asarray(win_idx)
        if freq is None:
            f_all, n_all = self.tensors(np.ones(self.nw, np.uint8))
            freq, nn = f_all[win_idx], n_all[win_idx].reshape(len(win_idx), self.k - 1, 16)

        def scan_fn(cands):
            c = cands.copy()
            c["win"] = win_idx[c["win"]]
            counts, _ = self.cscan(fmask, rmask, c)
            return comm.allreduce_sum(counts) if comm is not None else counts

        return _lib.walk(self.k, self.v, dnum, degeneracy, cover_number, freq, nn, mm_key, scan_fn, want_trace)

    def stats(self):
        nw = self.nw
        out = dict(gap_n=np.array(self.gap_n, np.int64), ent=np.zeros((nw, 4)), nuniq=np.zeros((nw, 3), np.int64),
                   mm_key=np.full(nw, KEY_EMPTY, np.uint64), mm_cnt=np.zeros(nw, np.int64),
                   mm_first=np.full(nw, KEY_EMPTY, np.uint64), n_iupac_gap=np.array(self.n_iupac_gap, np.int64))
        for wi, tab in enumerate(self.tables):
            best = None
            for key, (c, f) in tab.items():
                if self._is_cover(key):
                    out["ent"][wi, 0] += c
                    out["ent"][wi, 1] += c * np.log2(c)
                    out["nuniq"][wi, 0] += 1
                else:
                    out["ent"][wi, 2] += c
                    out["ent"][wi, 3] += c * np.log2(c)
                    out["nuniq"][wi, 1] += 1
                if key < KEY_BASE5:
                    out["nuniq"][wi, 2] += 1
                    if best is None or c > best[0] or (c == best[0] and f < best[1]):
                        best = (c, f, key)
            if best:
                out["mm_cnt"][wi], out["mm_first"][wi], out["mm_key"][wi] = best
        return out

    def tensors(self, sel):
        k = self.k
        freq = np.zeros((self.nw, 4, k), np.int64)
        nn = np.zeros((self.nw, k - 1, 4, 4), np.int64)
        for wi, tab in enumerate(self.tables):
            if not sel[wi]:
                continue
            for key, (c, _) in tab.items():
                if not self._is_cover(key):
                    continue
                hap = key_string(key, k)
                for i, ch in enumerate(hap):
                    if ch != "-":
                        freq[wi, BASES.index(ch), i] += c
                        if i and hap[i - 1] != "-":
                            nn[wi, i - 1, BASES.index(hap[i - 1]), BASES.index(ch)] += c
        return freq, nn

    def dump(self, w, max_n):
        items = sorted(self.tables[w].items(), key=lambda kv: kv[1][1])
        return (np.array([k for k, _ in items], np.uint64), np.array([v[0] for _, v in items], np.uint32),
                np.array([v[1] for _, v in items], np.uint64))

    def export(self, sel, counts):
        off = np.zeros(self.nw + 1, np.int64)
        keys, cnt, first = [], [], []
        for wi, tab in enumerate(self.tables):
            if sel[wi]:
                assert len(tab) == counts[wi]
                for key, (c, f) in tab.items():
                    keys.append(key)
                    cnt.append(c)
                    first.append(f)
            off[wi + 1] = len(keys)
        return off, np.array(keys, np.uint64), np.array(cnt, np.uint32), np.array(first, np.uint64)

    def merge(self, win_off, keys, cnt, first):
        for wi in range(self.nw):
            for i in range(int(win_off[wi]), int(win_off[wi + 1])):
                e = self.tables[wi].setdefault(int(keys[i]), [0, int(first[i])])
                e[0] += int(cnt[i])
                e[1] = min(e[1], int(first[i]))

    def match(self, q_win, q_allow):
        q_allow = np.asarray(q_allow).reshape(-1, 4)
        out = np.zeros(len(q_win), np.int64)
        for qi, wi in enumerate(q_win):
            allow = [int(x) for x in q_allow[qi]]
            for key in self.tables[int(wi)]:
                if key < KEY_BASE5:
                    hap = key_string(key, self.k)
                    out[qi] += all((allow[BASES.index(ch)] >> i) & 1 for i, ch in enumerate(hap))
        return out

    def exceptions(self):
        return (np.array([w for w, _ in self.exc], np.int32), np.array([s for _, s in self.exc], np.int32))


def key_string(key: int, k: int) -> str:
    key = int(key)
    if key < KEY_BASE5:
        mask = (1 << k) - 1
        b0, b1 = key & mask, (key >> k) & mask
        return "".join(BASES[((b0 >> i) & 1) | (((b1 >> i) & 1) << 1)] for i in range(k))
    x = key - KEY_BASE5
    out = []
    for _ in range(k):
        out.append("ACGT-"[x % 5])
        x //= 5
    return "".join(out)


class Dimer:
    """stand-in for _lib.Dimer: literal enumeration of (3' end, expansion) pairs in the reference's order"""

    def __init__(self, ctx, sets_list, min_end, max_end, init_both, loss_table, dg_consts):
        from oracle import dimer_oracle as dor
        self.primers = ["".join(CODE_CHARS[c] for c in s) for s in sets_list]
        self.n = len(self.primers)
        self.table = loss_table
        self.init_both = init_both
        self._dg = dor.delta_g
        self.ends = []
        for p in self.primers:
            k = len(p)
            top = min(max_end, k) if max_end > 0 else k + max_end
            lst = []
            for L in range(top, min_end - 1, -1):
                lst.extend(o.expand(p[k - L:]))
            self.ends.append(lst)
        self.exps = [o.expand(p) for p in self.primers]
        self.off_p = np.concatenate([[0], np.cumsum([len(e) for e in self.exps])]).astype(np.int64)
        self.off_e = np.concatenate([[0], np.cumsum([len(e) for e in self.ends])]).astype(np.int64)

    def _first(self, i, j):
        n_p = len(self.exps[j])
        for ei, end in enumerate(self.ends[i]):
            target = o.rc(end)
            gc = end.count("G") + end.count("C")
            for pi, p in enumerate(self.exps[j]):
                idx = p.find(target)
                if idx >= 0:
                    d2 = len(p) - len(end) - idx
                    if self.table[len(end), gc, d2] or (d2 == 0 and self._dg(end, self.init_both) < -5):
                        return ei * n_p + pi, d2
        return -1, -1

    def pairs(self, pi, pj):
        res = [self._first(int(a), int(b)) for a, b in zip(pi, pj)]
        return np.array([r[0] for r in res], np.int64), np.array([r[1] for r in res], np.int32)

    def grid(self, row0, row1, max_hits=1 << 22):
        out = []
        for i in range(row0, row1):
            for j in range(i, self.n):
                h, d2 = self._first(i, j)
                if h >= 0:
                    out.append((i, j, h, d2))
        a = np.array(out, np.int64).reshape(-1, 4)
        return a[:, 0].astype(np.int32), a[:, 1].astype(np.int32), a[:, 2], a[:, 3].astype(np.int32), 0

    def close(self):
        pass
