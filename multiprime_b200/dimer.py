"""Primer-dimer predicates (core:457-503 dimer_check; finDimer / get_Maxprimerset use the same family).

INTERIM host enumeration with numpy over 2-bit packed expansions; exact (it enumerates the same (end, expansion)
pairs the reference enumerates) but O(deg^2): the dimer-grid kernel replaces it for degenerate-heavy inputs."""
from __future__ import annotations

from functools import lru_cache
from math import log10

import numpy as np

from .iupac import ORDER, comp_set


def penalty_points(length, gc, d1, d2):
    """core:192-193"""
    return log10((2 ** length * 2 ** gc) / ((2 ** d1 - 0.9) * (2 ** d2 - 0.9)))


@lru_cache(maxsize=None)
def _loss_table(k: int, threshold: float):
    """hit[length][gc][d2] = Penalty_points(length, gc, 0, d2) >= threshold"""
    t = np.zeros((k + 1, k + 1, k + 1), dtype=bool)
    for L in range(1, k + 1):
        for gc in range(0, L + 1):
            for d2 in range(0, k - L + 1):
                t[L, gc, d2] = penalty_points(L, gc, 0, d2) >= threshold
    return t


def _pack_expansions(sets) -> np.ndarray:
    """all expansions as 2-bit packed int64 (position i at bits 2i..2i+1), reference product order"""
    acc = np.zeros(1, dtype=np.int64)
    for i, s in enumerate(sets):
        alts = np.asarray(ORDER[s], dtype=np.int64) << (2 * i)
        acc = (acc[:, None] | alts[None, :]).reshape(-1)
    return acc


def self_dimer(sets, threshold: float = 3.0) -> bool:
    """core:487-503: some 3' end (5..18 nt, expanded) whose reverse complement occurs in some expansion of the primer
    with Loss >= threshold at its LEFTMOST occurrence.  (The dG clause of core:501 needs d2 == 0, where Loss >= 3.5 > 3
    already holds, so it never decides at threshold 3.)"""
    assert threshold <= 3.5
    k = len(sets)
    table = _loss_table(k, threshold)
    prim = _pack_expansions(sets)
    for L in sorted({min(i, k) for i in range(5, 19)}, reverse=True):
        suffix = sets[k - L:]
        # reverse complement of every end expansion, packed
        rc = [comp_set(s) for s in reversed(suffix)]
        # GC count is invariant under reverse complement; expansions of rc(suffix) enumerate RC(e) for all e
        targets = _pack_expansions(rc)
        gc = np.zeros(len(targets), dtype=np.int64)
        for i in range(L):
            b = (targets >> (2 * i)) & 3
            gc += (b == 1) | (b == 2)
        mask = (1 << (2 * L)) - 1
        found = np.zeros((len(targets), len(prim)), dtype=bool)
        for o in range(0, k - L + 1):                       # ascending offsets: first match = leftmost
            m = ((prim[None, :] >> (2 * o)) & mask) == targets[:, None]
            new = m & ~found
            if new.any():
                d2 = k - L - o
                ok = table[L, gc, d2]
                if (new & ok[:, None]).any():
                    return True
                found |= m
    return False
