"""CPU restatement of the multiPrime-core (MC-EDPD) per-window algorithm.

TEST INFRASTRUCTURE ONLY.  Nothing under ``multiprime_b200/`` may import this
module; it is the checker for ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.

Parity status: PINNED.  ``tests/golden/core_*.json`` were produced by running the
live reference (``/root/reference/scripts/multiPrime-core_V20.py``, imported by
path) with ``tests/golden/make_golden.py``; ``tests/test_oracle_golden.py``
checks this restatement against every one of those records.

The restatement is written with plain strings / dicts / lists (no numpy, no
pandas) and follows the reference function by function; every function cites the
reference lines (``core`` = scripts/multiPrime-core_V20.py) it restates.
"""
from __future__ import annotations

import math
from itertools import product
from statistics import mean

# ----------------------------------------------------------------------------------------------
# alphabets and tables (core:105-110, 129-183)
# ----------------------------------------------------------------------------------------------
# expansion order of every IUPAC code: core:105-107
EXPAND = {
    "-": "-", "A": "A", "G": "G", "C": "C", "T": "T",
    "R": "AG", "Y": "CT", "M": "AC", "K": "GT", "S": "GC", "W": "AT",
    "H": "ATC", "B": "GTC", "V": "GAC", "D": "GAT", "N": "ATGC",
}
BASES = "ACGT"                      # core:185 (index order of freq / NN tensors)
BASE_IDX = {"A": 0, "C": 1, "G": 2, "T": 3}
# set -> IUPAC letter (the reference does this through float "scores", core:109-124;
# the score of a code is the sum of the scores of its bases, all sums are distinct)
SET2CODE = {frozenset(v): k for k, v in EXPAND.items() if k != "-"}
CODE2SET = {k: frozenset(v) for k, v in EXPAND.items() if k != "-"}
COMPLEMENT = str.maketrans("ATGCRYMKSWHBVDN", "TACGYRKMSWDVBHN")  # core:218

# nearest-neighbour tables, indexed [next][current] with A,C,G,T = 0..3 (core:154-174)
H_NN = [[-7.9, -8.5, -8.2, -7.2], [-8.4, -8, -9.8, -8.2], [-7.8, -10.6, -8, -8.5], [-7.2, -7.8, -8.4, -7.9]]
S_NN = [[-22.2, -22.7, -22.2, -21.3], [-22.4, -19.9, -24.4, -22.2], [-21, -27.2, -19.9, -22.7],
        [-20.4, -21, -22.4, -22.2]]
H_INIT = {"A": 2.3, "T": 2.3, "C": 0.1, "G": 0.1}
S_INIT = {"A": 4.1, "T": 4.1, "C": -2.8, "G": -2.8}
# stacking free energy pieces (core:129-147)
DG_FREEDOM = [[-0.7, -0.81, -0.65, -0.65], [-0.67, -0.72, -0.8, -0.65], [-0.69, -0.87, -0.72, -0.81],
              [-0.61, -0.69, -0.67, -0.7]]
DG_PENALTY = [[0.4, 0.575, 0.33, 0.73], [0.23, 0.32, 0.17, 0.33], [0.41, 0.45, 0.32, 0.575],
              [0.33, 0.41, 0.23, 0.4]]
DG_HBONDS = [[2, 2.5, 2.5, 2], [2.5, 3, 3, 2.5], [2.5, 3, 3, 2.5], [2, 2.5, 2.5, 2]]
DG_INIT = {"A": 0.98, "T": 0.98, "C": 1.03, "G": 1.03}

PRIMER_NM = 100          # core:177
MONO_MM = 50             # core:178
DIV_MM = 1.5             # core:179
DNTP_MM = 0.25           # core:180


def rc(seq: str) -> str:
    """reverse complement, core:221-222"""
    return seq.translate(COMPLEMENT)[::-1]


def expand(primer: str) -> list[str]:
    """All plain strings a degenerate string stands for, leftmost position slowest
    (core:368-380: itertools.product over per-position alternative lists)."""
    return ["".join(t) for t in product(*[EXPAND[c] for c in primer])]


def fold(c: str) -> int:
    """floor(score) of core:210-215: 1 for a plain base, n for an n-fold code"""
    return len(EXPAND[c])


def degeneracy(primer) -> int:
    """core:210-211 score_trans"""
    d = 1
    for c in primer:
        d *= fold(c)
    return d


def n_degenerate(primer) -> int:
    """core:214-215 dege_number"""
    return sum(1 for c in primer if fold(c) > 1)


def is_self_complementary(seq: str) -> bool:
    """core:237-246 symmetry(): even length and first half == RC(reversed second half)...
    which reduces to first half == complement of the reversed... restated literally."""
    if len(seq) % 2 == 1:
        return False
    half = len(seq) // 2
    return seq[:half] == rc(seq[half:][::-1])


def mismatch_positions(primer: str, hap: str) -> list[int]:
    """core:229-233 Y_distance.  Position i matches iff hap[i] is one of the bases the primer
    code at i stands for; '-' never matches (exhaustive check of the float-score trick)."""
    return [i for i, (c, h) in enumerate(zip(primer, hap)) if h not in CODE2SET[c]]


# ----------------------------------------------------------------------------------------------
# thermodynamics
# ----------------------------------------------------------------------------------------------
def delta_h_s(seq: str):
    """core:249-261"""
    dh = 0
    ds = 0
    for n in range(len(seq) - 1):
        nxt, cur = BASE_IDX[seq[n + 1]], BASE_IDX[seq[n]]
        dh += H_NN[nxt][cur]
        ds += S_NN[nxt][cur]
    dh += H_INIT[seq[0]] + H_INIT[seq[-1]]
    ds += S_INIT[seq[0]] + S_INIT[seq[-1]]
    if is_self_complementary(seq):
        ds += -1.4
    return dh * 1000, ds


def salt_correction() -> float:
    """core:293-326.  The three-line "Eq 16" expression is cut after its first line by a missing
    line continuation, so the correction is a + b*ln(Mg) with `a` from the R<6 branch: a constant."""
    free_div = (DIV_MM - DNTP_MM) / 1000.0
    ratio = math.sqrt(free_div) / (MONO_MM / 1000)
    assert 0.22 <= ratio < 6.0
    mono = MONO_MM / 1000.0
    a = 3.92 * pow(10, -5) * (0.843 - (0.352 * math.sqrt(mono) * math.log(mono, math.e)))
    b = -9.11 * pow(10, -6)
    return a + (b * math.log(free_div, math.e))


def tm_unrounded(seq: str) -> float:
    """core:328-335 before round()"""
    dh, ds = delta_h_s(seq)
    denom = 1 * pow(10, 9) if is_self_complementary(seq) else 4 * pow(10, 9)
    return 1 / ((1 / (dh / (ds + 1.9872 * math.log(PRIMER_NM / denom, math.e)))) + salt_correction()) - 273.15


def tm(seq: str) -> float:
    """core:282-336 Calc_Tm_v2"""
    return round(tm_unrounded(seq), 2)


def delta_g(sequence: str) -> float:
    """core:466-485: max over expansions of the stacking dG, rounded to 2"""
    vals = []
    for seq in expand(sequence):
        g = 0
        for n in range(len(seq) - 1):
            i, j = BASE_IDX[seq[n + 1]], BASE_IDX[seq[n]]
            g += DG_FREEDOM[i][j] * DG_HBONDS[i][j] + DG_PENALTY[i][j]
        if sequence[-2:] == "TA":
            g += DG_INIT[seq[0]] + DG_INIT[seq[-1]] + 0.4
        else:
            g += DG_INIT[seq[0]] + DG_INIT[seq[-1]]
        g -= (0.175 * math.log(50 / 1000, math.e) + 0.20) * len(seq)
        if is_self_complementary(seq):
            g += 0.4
        vals.append(g)
    return round(max(vals), 2)


def penalty_points(length, gc, d1, d2) -> float:
    """core:192-193"""
    return math.log10((2 ** length * 2 ** gc) / ((2 ** d1 - 0.9) * (2 ** d2 - 0.9)))


# ----------------------------------------------------------------------------------------------
# primer filters
# ----------------------------------------------------------------------------------------------
def _repeat_patterns() -> set[str]:
    """core:196-207 (patterns containing '#' can never match a primer and are dropped)"""
    pats = set()
    for i in BASES:
        pats.add(i * 4)
        for j in BASES:
            if i != j:
                pats.add((i + j) * 4)
            for k in BASES:
                if i != j and j != k:           # python chain `i != j != k`
                    pats.add((i + j + k) * 3)
    return pats


REPEATS = _repeat_patterns()


def gc_content(primer: str) -> float:
    """core:401-407"""
    vals = [round((e.count("G") + e.count("C")) / len(e), 3) for e in expand(primer)]
    return round(mean(vals), 2)


def has_repeat(primer: str) -> bool:
    """core:410-416"""
    return any(pat in e for e in expand(primer) for pat in REPEATS)


def has_hairpin(primer: str, distance: int) -> bool:
    """core:387-398"""
    n = 0
    while n <= len(primer) - 5 - 5 - distance:
        kmers = expand(primer[n:n + 5])
        tails = expand(primer[n + 5 + distance:])
        for kmer in kmers:
            target = rc(kmer)
            for tail in tails:
                if target in tail:
                    return True
        n += 1
    return False


def information(primer: str, gc_lo: float, gc_hi: float, distance: int):
    """core:507-521 primer_pre_filter: GC (float) when clean, else the failed filters joined by '|'"""
    notes = []
    gc = gc_content(primer)
    if not gc_lo <= gc <= gc_hi:
        notes.append("GC_out_of_range (" + str(gc) + ")")
    if has_repeat(primer):
        notes.append("di_nucleotide")
    if has_hairpin(primer, distance):
        notes.append("hairpin")
    return gc if not notes else "|".join(notes)


def suffix_ends(primer: str, lo: int = 5, count: int = 14) -> list[str]:
    """core:457-464 current_end: expansions of primer[-i:] for i = lo .. lo+count-1
    (for i > len(primer) the slice is the whole primer again)"""
    ends = []
    for i in range(lo, lo + count):
        s = primer[-i:]
        if s:
            ends.extend(expand(s))
    return ends


def self_dimer(primer: str) -> bool:
    """core:487-503 dimer_check"""
    ends = sorted(suffix_ends(primer), key=len, reverse=True)
    exps = expand(primer)
    for end in ends:
        target = rc(end)
        for p in exps:
            idx = p.find(target)
            if idx >= 0:
                d2 = len(p) - len(end) - idx
                loss = penalty_points(len(end), end.count("G") + end.count("C"), 0, d2)
                if loss >= 3 or (delta_g(end) < -5 and d2 == 0):
                    return True
    return False


# ----------------------------------------------------------------------------------------------
# input handling
# ----------------------------------------------------------------------------------------------
def parse_msa(path: str):
    """core:441-455 -> (ids, sequences).  ids keep the leading '>'"""
    seqs: dict[str, str] = {}
    cur = None
    allowed = set("ACGTRYMKSWHBVD")
    with open(path) as fh:
        for line in fh:
            if line.startswith("#"):
                continue
            if line.startswith(">"):
                cur = line.strip().split(" ")[0]
            else:
                s = "".join(c if c in allowed else "-" for c in line.strip().upper())
                seqs[cur] = seqs.get(cur, "") + s
    return list(seqs.keys()), list(seqs.values())


def region(seqs: list[str], fraction: float):
    """core:617-640 seq_attribute: [start, stop) over which windows are tried.
    np.quantile(method="higher"/"lower") picks sorted[ceil/floor((n-1)*q)] (float64 product)."""
    starts = sorted(len(s) - len(s.lstrip("-")) for s in seqs)
    stops = sorted(len(s.rstrip("-")) for s in seqs)
    v = (len(seqs) - 1) * fraction
    return starts[int(math.ceil(v))], stops[int(math.floor(v))]


def strict_positions(coordinate: str, k: int):
    """core:1091-1101 get_Y: 0-based indices where a mismatch disqualifies F / R coverage"""
    f, r = set(), set()
    for tok in coordinate.split(","):
        y = int(tok.strip())
        if y > 0:
            f.add(y)
            r.add(k - y)
        else:
            f.add(k + y + 1)
            r.add(-y + 1)
    return f, r


class Params:
    """constructor mapping of core:1185-1189 (-n -> dnum, -d -> degeneracy, -c -> coordinate ...)"""

    def __init__(self, k=18, dnum=4, degeneracy=10, variation=1, entropy=3.6, gc="0.2,0.7", size=100,
                 fraction=0.8, coordinate="1,2,-1", away=4):
        self.k = k
        self.dnum = dnum
        self.degeneracy = degeneracy
        self.variation = variation
        self.entropy = entropy
        self.gc = gc
        self.gc_lo, self.gc_hi = (float(x) for x in gc.split(","))
        self.size = size
        self.fraction = fraction
        self.coordinate = coordinate
        self.away = away
        self.strict_f, self.strict_r = strict_positions(coordinate, k)

    def entropy_threshold(self, region_len: int) -> float:
        """core:642-649"""
        if region_len < 5000:
            return self.entropy
        if region_len < 10000:
            return self.entropy * 0.95
        return self.entropy * 0.9


# ----------------------------------------------------------------------------------------------
# per-window pieces
# ----------------------------------------------------------------------------------------------
def window_kmer(s: str, p: int, k: int) -> str:
    """core:666-687: the k-mer sequence `s` contributes to the window starting at column p."""
    w = s[p:p + k]
    if w != "-" * k:
        if w.startswith("-"):
            body = w.lstrip("-")
            g = len(w) - len(body)
            left = s[0:p].replace("-", "")
            if len(left) >= g:
                w = left[len(left) - g:] + body
        if w.endswith("-"):
            body = w.rstrip("-")
            g = len(w) - len(body)
            right = s[p + k:].replace("-", "")
            if len(right) >= g:
                w = body + right[0:g]
    if len(w) < k:
        g = k - len(w)
        left = s[0:p].replace("-", "")
        if len(left) >= g:
            w = left[len(left) - g:] + w
    return w


class WindowTally:
    """state built by the sequence loop of get_primers, core:653-711"""

    def __init__(self):
        self.cover: dict[str, int] = {}          # expanded haplotype -> number of expansion rows
        self.cover_mm: dict[str, int] = {}       # the gap-free ones
        self.ids_of: dict[str, list] = {}        # haplotype -> ids (non_gap_seq_id)
        self.cover_number = 0                    # sequences with <= v gaps
        self.gap_seq: dict[str, int] = {}        # raw k-mer with > v gaps -> count
        self.gap_ids: dict[str, list] = {}       # expanded gap k-mer -> ids
        self.gap_n = 0
        self.gap_fail = False


def tally_window(ids, seqs, p: int, prm: Params) -> WindowTally:
    t = WindowTally()
    n_total = len(seqs)
    for sid, s in zip(ids, seqs):
        w = window_kmer(s, p, prm.k)
        if w.count("-") > prm.variation:
            t.gap_seq[w] = t.gap_seq.get(w, 0) + 1
            t.gap_n += 1
            if round(t.gap_n / n_total, 2) >= (1 - prm.fraction):
                t.gap_fail = True
                break
            for e in expand(w):
                t.gap_ids.setdefault(e, []).append(sid)
        else:
            t.cover_number += 1
            for e in expand(w):
                t.cover[e] = t.cover.get(e, 0) + 1
                t.ids_of.setdefault(e, []).append(sid)
                if "-" not in e:
                    t.cover_mm[e] = t.cover_mm.get(e, 0) + 1
    if round(t.gap_n / n_total, 2) >= (1 - prm.fraction):
        t.gap_fail = True
    return t


def entropies(t: WindowTally):
    """core:602-614"""
    c_bit = 0
    t_bit = 0
    tot = t.cover_number + t.gap_n
    for c in t.cover.values():
        c_bit += (c / t.cover_number) * math.log((c / t.cover_number), 2)
        t_bit += (c / tot) * math.log((c / tot), 2)
    for g in t.gap_seq.values():
        t_bit += (g / tot) * math.log((g / tot), 2)
    return round(-c_bit, 2), round(-t_bit, 2)


def base_counts(t: WindowTally, k: int):
    """core:541-554 state_matrix over expansion rows: (freq[4][k], set of bases seen)"""
    freq = [[0] * k for _ in range(4)]
    seen = set()
    for hap, c in t.cover.items():
        for col, ch in enumerate(hap):
            if ch != "-":
                freq[BASE_IDX[ch]][col] += c
                seen.add(ch)
    return freq, seen


def dinuc_counts(t: WindowTally, k: int):
    """core:556-577 trans_matrix: nn[c][x][y] = rows with base x at column c and y at c+1"""
    nn = [[[0] * 4 for _ in range(4)] for _ in range(k - 1)]
    for hap, c in t.cover.items():
        for col in range(k - 1):
            x, y = hap[col], hap[col + 1]
            if x != "-" and y != "-":
                nn[col][BASE_IDX[x]][BASE_IDX[y]] += c
    return nn


def _argmax_first(vals):
    best = 0
    for i in range(1, len(vals)):
        if vals[i] > vals[best]:
            best = i
    return best


def viterbi_seed(freq, nn, k: int) -> list[int]:
    """core:579-593: max-sum path, first maximum on ties"""
    score = [freq[b][0] for b in range(4)]
    paths = [[b] for b in range(4)]
    for t in range(1, k):
        new_score, new_paths = [], []
        for cur in range(4):
            cand = [score[prev] + nn[t - 1][prev][cur] + freq[cur][t] for prev in range(4)]
            best = _argmax_first(cand)
            new_score.append(cand[best])
            new_paths.append(paths[best] + [cur])
        score, paths = new_score, new_paths
    return paths[_argmax_first(score)]


def majority_seed(cover_mm: dict) -> list[int]:
    """core:595-600: most frequent gap-free haplotype, first seen wins ties (stable sort)"""
    best_key, best = None, -1
    for key, c in cover_mm.items():
        if c > best:
            best_key, best = key, c
    return [BASE_IDX[ch] for ch in best_key]


def _argsort_desc(vals) -> list[int]:
    """np.argsort(vals)[::-1] with the stable ascending sort the reference's numpy 1.21 gives
    for 4 elements: ties come out highest index first."""
    return sorted(range(len(vals)), key=lambda i: vals[i])[::-1]


def _npos(vals) -> int:
    return sum(1 for v in vals if v > 0)


def coverage_scan(primer: str, t: WindowTally, hap_universe, prm: Params):
    """core:1103-1130 mis_primer_check over the haplotypes that are not expansions of primer"""
    mine = set(expand(primer))
    f_mis = r_mis = 0
    f_non, r_non = {}, {}
    for hap in hap_universe:
        if hap in mine:
            continue
        pos = mismatch_positions(primer, hap)
        if len(pos) > prm.variation:
            f_non[hap] = t.ids_of[hap]
            r_non[hap] = t.ids_of[hap]
            continue
        if prm.strict_f.intersection(pos):
            f_non[hap] = t.ids_of[hap]
        else:
            f_mis += t.cover[hap]
        if prm.strict_r.intersection(pos):
            r_non[hap] = t.ids_of[hap]
        else:
            r_mis += t.cover[hap]
    return f_mis, f_non, r_mis, r_non


def _union(code: str, base: str) -> str:
    return SET2CODE[CODE2SET[code] | {base}]    # KeyError mirrors the reference's table KeyError


def refine_step(primer: list[str], init_cov: int, cover: dict, seed: list[int], nn_cov: list[int], nn):
    """core:922-1089 refine_by_NN_array: one more degenerate base at (one of) the weakest junction(s)."""
    k = len(primer)
    last = k - 2
    lowest = min(nn_cov)
    outcomes = []
    for j in [i for i, c in enumerate(nn_cov) if c == lowest]:
        cov_j = list(nn_cov)
        nn_j = [[row[:] for row in layer] for layer in nn]
        pr = list(primer)
        gained = init_cov
        row, col = seed[j], seed[j + 1]

        def add_base(pos, idx):
            nonlocal gained
            trial = list(pr)
            trial[pos] = BASES[idx]
            for e in expand("".join(trial)):
                if e in cover:
                    gained += cover[e]
            pr[pos] = _union(pr[pos], BASES[idx])

        def middle(jj):
            # refine position jj+1 from row `row` of layer jj and column seed[jj+2] of layer jj+1
            nrow, ncol = seed[jj + 1], seed[jj + 2]
            m = [min(nn_j[jj][row][x], nn_j[jj + 1][x][ncol]) for x in range(4)]
            if _npos(m) > 1:
                for idx in _argsort_desc(m):
                    if idx != col:
                        add_base(jj + 1, idx)
                        for x in range(4):
                            nn_j[jj][x][col] += nn_j[jj][x][idx]
                            nn_j[jj][x][idx] = 0
                        for y in range(4):
                            nn_j[jj + 1][nrow][y] += nn_j[jj + 1][idx][y]
                            nn_j[jj + 1][idx][y] = 0
                        cov_j[jj] = nn_j[jj][row][col]
                        cov_j[jj + 1] = nn_j[jj + 1][nrow][ncol]
                        break

        if j == 0:
            column0 = [nn_j[0][x][col] for x in range(4)]
            if _npos(column0) > 1:                              # core:941-965 position 0
                for idx in _argsort_desc(column0):
                    if idx != row:
                        add_base(0, idx)
                        for y in range(4):
                            nn_j[0][row][y] += nn_j[0][idx][y]
                            nn_j[0][idx][y] = 0
                        cov_j[0] = nn_j[0][row][col]
                        break
            elif _npos(nn_j[0][row]) > 1:                       # core:967-1001 position 1
                middle(0)
        elif j == last:                                         # core:1004-1031 last position
            rowvals = list(nn_j[j][row])
            if _npos(rowvals) > 1:
                for idx in _argsort_desc(rowvals):
                    if idx != col:
                        add_base(j + 1, idx)
                        for x in range(4):
                            nn_j[j][x][col] += nn_j[j][x][idx]
                            nn_j[j][x][idx] = 0
                        cov_j[j] = nn_j[j][row][col]
                        break
        else:                                                   # core:1032-1072
            middle(j)
        outcomes.append((gained, pr, cov_j, nn_j))
    best = 0
    for i in range(1, len(outcomes)):
        if outcomes[i][0] > outcomes[best][0]:
            best = i
    gained, pr, cov_j, nn_j = outcomes[best]
    return pr, gained, cov_j, nn_j, degeneracy(pr), n_degenerate(pr)


def run_track(seed: list[int], init_cov: int, t: WindowTally, hap_universe, nn, prm: Params, trace=None):
    """core:860-920 coverage_stast: seed primer -> refined primer and its mismatch coverage"""
    k = prm.k
    primer = [BASES[b] for b in seed]
    nn_cov = [nn[j][seed[j]][seed[j + 1]] for j in range(k - 1)]
    total = t.cover_number
    f_mis, f_non, r_mis, r_non = coverage_scan("".join(primer), t, hap_universe, prm)
    if trace is not None:
        trace.append("".join(primer))
    while init_cov + f_mis < total or init_cov + r_mis < total:
        primer, init_cov, cov_new, nn, deg, ndeg = refine_step(primer, init_cov, t.cover, seed, nn_cov, nn)
        f_mis, f_non, r_mis, r_non = coverage_scan("".join(primer), t, hap_universe, prm)
        if trace is not None:
            trace.append("".join(primer))
        if max(f_mis, r_mis) == total:
            break
        if cov_new == nn_cov:
            break
        if 2 * deg > prm.degeneracy or 3 * deg / 2 > prm.degeneracy or ndeg == prm.dnum:
            break
        nn_cov = cov_new
    final = "".join(primer)
    info = information(final, prm.gc_lo, prm.gc_hi, prm.away)
    return final, init_cov + f_mis, init_cov + r_mis, info, f_non, r_non


def design_window(ids, seqs, p: int, prm: Params, entropy_threshold: float, trace=None):
    """core:651-757 get_primers + core:759-858 degenerate_by_NN_algorithm.

    Returns None when the window is rejected, else a dict with
      row      [p, cBit, tBit, primer, ndeg, nonsense, perfect, F_mis, R_mis, Tm, information]
      non_cov  [F dict, R dict]   haplotype -> ids
      gap_ids  dict               expanded gap k-mer -> ids
    `trace` (optional list) receives every primer handed to the coverage scan, in call order.
    """
    k = prm.k
    t = tally_window(ids, seqs, p, prm)
    if t.gap_fail or len(t.cover) < 1:
        return None
    c_bit, t_bit = entropies(t)
    if t_bit > entropy_threshold:
        return None
    freq, seen = base_counts(t, k)
    if len(seen) < 4:
        return None
    if any(sum(freq[b][c] for b in range(4)) == 0 for c in range(k)):
        return None
    hap_universe = list(t.cover.keys())           # snapshot, core:765
    keys = set(hap_universe)                      # grows by the defaultdict look-ups of core:787/800/809/835
    nn = dinuc_counts(t, k)
    nm_seed = viterbi_seed(freq, nn, k)

    def cover_of(seed):
        s = "".join(BASES[b] for b in seed)
        keys.add(s)
        return t.cover.get(s, 0)

    if t.cover_mm:
        mm_seed = majority_seed(t.cover_mm)
        if nm_seed == mm_seed:
            res = run_track(nm_seed, cover_of(nm_seed), t, hap_universe, nn, prm, trace)
        else:
            r_nm = run_track(nm_seed, cover_of(nm_seed), t, hap_universe, nn, prm, trace)
            r_mm = run_track(mm_seed, cover_of(mm_seed), t, hap_universe, nn, prm, trace)
            res = r_nm if (r_nm[1] + r_nm[2]) > (r_mm[1] + r_mm[2]) else r_mm
    else:
        res = run_track(nm_seed, cover_of(nm_seed), t, hap_universe, nn, prm, trace)
    primer, f_cov, r_cov, info, f_non, r_non = res
    exps = expand(primer)
    nonsense = len(set(exps) - keys)
    tm_avg = round(mean([tm(e) for e in exps]), 2)
    perfect = sum(t.cover.get(e, 0) for e in exps)
    if self_dimer(primer):
        return None
    row = [p, c_bit, t_bit, primer, n_degenerate(primer), nonsense, perfect, f_cov, r_cov, tm_avg, info]
    return {"row": row, "non_cov": [f_non, r_non], "gap_ids": t.gap_ids}


def run_core(path: str, prm: Params, positions=None):
    """core:1133-1180 run(): all windows of the conserved region (or the given positions)."""
    ids, seqs = parse_msa(path)
    start, stop = region(seqs, prm.fraction)
    if stop - start < prm.size:
        raise SystemExit(1)
    thr = prm.entropy_threshold(stop - start)
    out = []
    for p in (positions if positions is not None else range(start, stop - prm.k)):
        r = design_window(ids, seqs, p, prm, thr)
        if r is not None:
            out.append(r)
    return start, stop, out


TSV_HEADER = ["Position", "Entropy of cover (bit)", "Entropy of total (bit)", "Optimal_primer",
              "primer_degenerate_number", "nonsense_primer_number", "Optimal_coverage", "Mis-F-coverage",
              "Mis-R-coverage", "Tm", "Information"]
