"""CPU restatement of the dimer family of the reference: finDimer, get_Maxprimerset (V1.3) and get_multiPrime pairing.

TEST INFRASTRUCTURE ONLY (see oracle/mp_oracle.py).  Parity status: PINNED — tests/test_oracle_dimer_golden.py
reproduces the fixtures that tests/golden/make_golden.py recorded from the live reference scripts
(finDimer_V4.py, get_Maxprimerset_V1.3.py, get_multiPrime.py): dimer_findimer.json, cover_maxprimerset.json,
pairs_get_multiprime.json.

Plain strings and loops; every function cites the reference lines it restates
(fd = scripts/finDimer_V4.py, ms = scripts/get_Maxprimerset_V1.3.py, gm = scripts/get_multiPrime.py).
"""
from __future__ import annotations

import math
from bisect import bisect_left
from statistics import mean

from oracle.mp_oracle import (BASE_IDX, DG_FREEDOM, DG_HBONDS, DG_INIT, DG_PENALTY, REPEATS, expand,
                              is_self_complementary, penalty_points, rc)


# ----------------------------------------------------------------------------------------------
# shared pieces
# ----------------------------------------------------------------------------------------------
def delta_g(sequence: str, init_both: bool = True) -> float:
    """fd:171-189 / ms:171-189 (initiation term of both ends) and gm:398-416 (first base only)"""
    vals = []
    for seq in expand(sequence):
        g = 0
        for n in range(len(seq) - 1):
            i, j = BASE_IDX[seq[n + 1]], BASE_IDX[seq[n]]
            g += DG_FREEDOM[i][j] * DG_HBONDS[i][j] + DG_PENALTY[i][j]
        init = DG_INIT[seq[0]] + DG_INIT[seq[-1]] if init_both else DG_INIT[seq[0]]
        if sequence[-2:] == "TA":
            g += init + 0.4
        else:
            g += init
        g -= (0.175 * math.log(50 / 1000, math.e) + 0.20) * len(seq)
        if is_self_complementary(seq):
            g += 0.4
        vals.append(g)
    return round(max(vals), 2)


def ends_5_to_18(primer: str) -> list[str]:
    """fd:162-169 current_end: expansions of primer[-i:] for i = 5..18"""
    out = []
    for i in range(5, 19):
        s = primer[-i:]
        if s:
            out.extend(expand(s))
    return out


# ----------------------------------------------------------------------------------------------
# finDimer (fd:191-224, rows in V5 order = position order)
# ----------------------------------------------------------------------------------------------
def find_dimers(primers: dict, threshold: float = 3.96) -> list[tuple]:
    """primers: {sequence: header} as fd:138-146 builds it.  One row per (i, j >= i) pair: the first hit."""
    plist = list(primers.keys())
    rows = []
    for pos, pi in enumerate(plist):
        ends = sorted(list(ends_5_to_18(pi)), key=len, reverse=True)
        for pj in plist[pos:]:
            hit = None
            for end in ends:
                target = rc(end)
                for p in expand(pj):
                    idx = p.find(target)
                    if idx >= 0:
                        gc = end.count("G") + end.count("C")
                        d2 = len(p) - len(end) - idx
                        loss = penalty_points(len(end), gc, 0, d2)
                        dg = delta_g(end)
                        if loss >= threshold or (dg < -5 and d2 == 0):
                            hit = (primers[pi], pi, end, dg, len(end), 0, gc, primers[pj], pj, d2, loss)
                            break
                if hit:
                    break
            if hit:
                rows.append(hit)
    return rows


# ----------------------------------------------------------------------------------------------
# get_Maxprimerset V1.3
# ----------------------------------------------------------------------------------------------
def dimer_examination(primer_f: str, primer_r: str, accepted: set) -> bool:
    """ms:193-215: the expansions of F and R plus everything accepted so far; ends = suffixes 5..len-1"""
    total = set(expand(primer_f) + expand(primer_r)) | set(accepted)
    ends = set()
    for cp in total:
        for a in range(5, len(cp)):
            ends.update(expand(cp[-a:]))
    for end in sorted(ends, key=len, reverse=True):
        target = rc(end)
        for primer in total:
            idx = primer.find(target)
            if idx >= 0:
                gc = end.count("G") + end.count("C")
                d2 = len(primer) - len(end) - idx
                if penalty_points(len(end), gc, 0, d2) >= 3 or (delta_g(end) < -5 and d2 == 0):
                    return True
    return False


COVER_COLUMNS = ["#Primer", "Primer_rank", "Primer_F", "Primer_R", "PCR_product (Length:Tm:Coverage)",
                 "Coverage number with error in top N", "Primer position (representative sequence)"]


def _table(rows) -> str:
    lines = ["\t".join(COVER_COLUMNS)]
    for r in rows:
        lines.append("\t".join("" if x is None else str(x) for x in (r + [None] * 7)[:7]))
    return "\n".join(lines) + "\n"


def greedy_maximal(primers: list, step: int = 5):
    """ms:291-356 -> (output table text, .next.xls text, stdout lines)"""
    accepted: set = set()
    rows, nxt, out = [], [], []
    for row in primers:
        if len(row) <= 1:
            out.append("Non primers: virus {} missing!".format(row[0]))
            nxt.append("\t".join(row))
            continue
        col = 1
        while col <= len(row) - step:
            if dimer_examination(row[col], row[col + 1], accepted):
                col += step
                if col > len(row) - step:
                    rows.append([row[0]])
                    out.append("virus {} missing!".format(row[0]))
                    nxt.append("\t".join(row))
                    break
            else:
                rows.append([row[0], str(col)] + row[col:col + 5])
                accepted |= set(expand(row[col]) + expand(row[col + 1]))
                break
    return _table(rows), "".join(x + "\n" for x in nxt), out


def greedy_maximum(primers: list, step: int = 5):
    """ms:218-282 incl. the loop-variable behaviour of the reference (the outer `for` hands out the next index after a
    backtrack).  Returns (table text or None, stdout lines, exit code)."""
    accepted: set = set()
    before, jdict, rows = {}, {}, []
    blank = 0
    col = 1
    for it in range(len(primers)):
        rp = it
        if len(primers[rp]) <= 1:
            blank += 1
            continue
        while col <= len(primers[rp]) - step:
            row = primers[rp]
            if dimer_examination(row[col], row[col + 1], accepted):
                col += step
                while col > len(primers[rp]) - step:
                    rp -= 1
                    if rp < blank:
                        return None, ["Non maximum primer set. Try maximal primer set!"], 1
                    col = jdict[rp] + step
                    accepted = before[rp]
                    rows.pop()
            else:
                rows.append([row[0], str(col)] + row[col:col + 5])
                before[rp] = set(accepted)
                accepted = accepted | set(expand(row[col]) + expand(row[col + 1]))
                jdict[rp] = col
                col = 1
                break
    return _table(rows), [], 0


def sort_clusters(lines: list[str]) -> list[list[str]]:
    """ms:368-371: split on tabs, drop empty fields, stable sort by number of fields"""
    return sorted([list(filter(None, ln.strip().split("\t"))) for ln in lines], key=len)


# ----------------------------------------------------------------------------------------------
# get_multiPrime pairing (gm:303-662)
# ----------------------------------------------------------------------------------------------
def _gc_mean(seq: str) -> float:
    """gm:450-456 (mean of the rounded fractions, not rounded)"""
    return mean([round((e.count("G") + e.count("C")) / len(e), 3) for e in expand(seq)])


def _hairpin(primer: str, distance: int) -> bool:
    """gm:373-384"""
    n = 0
    while n <= len(primer) - 5 - 5 - distance:
        for kmer in expand(primer[n:n + 5]):
            for tail in expand(primer[n + 5 + distance:]):
                if rc(kmer) in tail:
                    return True
        n += 1
    return False


def _repeat(primer: str) -> bool:
    return any(pat in e for e in expand(primer) for pat in REPEATS)


def _gc_clamp(primer: str) -> bool:
    """gm:467-473"""
    return any(_gc_mean(primer[-i:]) > 0.6 for i in range(4, 17))


def _term_degenerate(primer: str, term: int) -> bool:
    """gm:439-448"""
    if term == 0:
        return False
    d = 1
    for ch in primer[-term:]:
        d *= len(expand(ch))
    return d > 1


def _fr_dimer(f: str, r: str) -> bool:
    """gm:419-437"""
    ends = set(ends_5_to_18(f)) | set(ends_5_to_18(r))
    for pp in (f, r):
        for end in ends:
            target = rc(end)
            for p in expand(pp):
                idx = p.find(target)
                if idx >= 0:
                    d2 = len(p) - len(end) - idx
                    loss = penalty_points(len(end), end.count("G") + end.count("C"), 0, d2)
                    if loss > 3.6 or (delta_g(end, init_both=False) < -5 and d2 == 0):
                        return True
    return False


def pair_candidates(tsv_text: str, gap_id: dict, non_cover_id: dict, number: int, out_path: str, size="250,500",
                    fraction=0.6, diff_tm=4, term=4, distance=4, adaptor="TCTTTCCCTACACGACGCTCTTCCGATCT,"
                                                                          "TCTTTCCCTACACGACGCTCTTCCGATCT", gc="0.4,0.6"):
    """gm:599-662 run() -> (one-line txt, xls text, fa text, stdout lines).  `gc` is the class default the reference
    always uses (its -g flag never reaches the constructor)."""
    primers = {}
    for line in tsv_text.splitlines():
        if line.startswith("Pos") or not line.strip():
            continue
        f = line.strip().split("\t")
        primers[int(f[0])] = [f[3], round(float(f[9]), 2)]
    lo, hi = (float(x) for x in gc.split(","))
    cand = sorted(p for p, (seq, _) in primers.items()
                  if not _hairpin(seq, distance) and not (_gc_mean(seq) > hi or _gc_mean(seq) < lo) and not _repeat(seq))
    min_len, max_len = (int(x) for x in size.split(","))
    ad = adaptor.split(",")
    out = ["Candidata degenerate primer number is: {}".format(len(cand))]
    if int(cand[-1]) - int(cand[0]) < min_len:
        out.append("Max PCR product legnth < min len!")
        return out_path + "\n", None, None, out
    pairs = []

    def one_pass(threshold, echo):
        for s in range(len(cand)):
            if echo:
                out.append(str(s))
            f = primers[cand[s]][0]
            if _hairpin(ad[0] + f, distance) or _term_degenerate(f, term) or _gc_clamp(f):
                continue
            a = bisect_left(cand, cand[s] + min_len)
            b = len(cand) - 1 if cand[s] + max_len > cand[-1] else bisect_left(cand, cand[s] + max_len) - 1
            for t in range(a, b + 1):
                r = rc(primers[cand[t]][0])
                if _hairpin(ad[1] + r, distance) or _term_degenerate(r, term) or _gc_clamp(r):
                    continue
                dist = int(cand[t]) - int(cand[s]) + 1
                if dist > max_len:
                    out.append("Error! PCR product greater than max length !")
                    break
                if not min_len <= dist <= max_len:
                    continue
                if _fr_dimer(f, r):
                    out.append("Dimer detection between Primer-F and Primer-R!")
                    continue
                tm_f, tm_r = primers[cand[s]][1], primers[cand[t]][1]
                if abs(tm_f - tm_r) > diff_tm:
                    continue
                un = []
                sp, tp = str(cand[s]), str(cand[t])
                for d in (gap_id[sp], non_cover_id[sp][0], gap_id[tp], non_cover_id[tp][1]):
                    for lst in d.values():
                        un.extend(set(lst))
                n_un = len(set(un))
                if n_un / number > threshold:
                    continue
                cov = number - n_un
                pairs.append((f, r, str(dist) + ":" + str(round(mean([tm_f, tm_r]), 2)) + ":" + str(round(cov / number, 4)),
                              cov, str(cand[s]) + ":" + str(cand[t])))

    thr = 1 - fraction
    one_pass(thr, True)
    if len(pairs) < 10:
        thr += 0.1
        one_pass(thr, False)
    pid = out_path.split("/")[-1].rstrip(".txt")
    txt = out_path + "\t"
    xls = "\t".join(["Primer_F_seq", "Primer_R_seq", "Product length:Tm:coverage_percentage", "Target number",
                     "Primer_start_end"]) + "\n"
    fa = ""
    for i in sorted(pairs, key=lambda k: k[3], reverse=True):
        txt += "\t".join(map(str, i)) + "\t"
        xls += "\t".join(map(str, i)) + "\n"
        a, b = i[4].split(":")
        fa += ">" + pid + "_" + a + "F\n" + i[0] + "\n>" + pid + "_" + b + "R\n" + i[1] + "\n"
    return txt + "\n", xls, fa, out
