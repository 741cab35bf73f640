"""Primer-dimer predicates (core:457-503 dimer_check; finDimer / get_Maxprimerset / get_multiPrime use the same
family) on top of the device dimer engine (csrc/mpb_dimer.cu).

The host only prepares what must be decided by the reference's own float expressions: the Loss truth table
(core:192-193 evaluated for every (length, GC, d2)) and the constants of the stacking dG (core:466-485)."""
from __future__ import annotations

import math
from functools import lru_cache
from math import log10

import numpy as np

from . import _lib


def penalty_points(length, gc, d1, d2):
    """core:192-193"""
    return log10((2 ** length * 2 ** gc) / ((2 ** d1 - 0.9) * (2 ** d2 - 0.9)))


@lru_cache(maxsize=None)
def loss_table(threshold: float, strict: bool = False) -> np.ndarray:
    """table[length][gc][d2] = Loss >= threshold (or > threshold when strict, get_multiPrime.py:435)"""
    t = np.zeros((33, 33, 33), dtype=np.uint8)
    for L in range(1, 33):
        for gc in range(0, L + 1):
            for d2 in range(0, 33 - L):
                x = penalty_points(L, gc, 0, d2)
                t[L, gc, d2] = (x > threshold) if strict else (x >= threshold)
    return t


def _dg_round_threshold() -> float:
    """largest double g with round(g, 2) < -5 (the reference compares the rounded dG, core:485/501)"""
    x = -5.005
    while not round(x, 2) < -5:
        x = math.nextafter(x, -math.inf)
    while round(math.nextafter(x, math.inf), 2) < -5:
        x = math.nextafter(x, math.inf)
    return x


@lru_cache(maxsize=None)
def dg_consts() -> tuple:
    """the 24 constants mpb_dimer_prepare wants, evaluated with the reference's expressions (core:129-147, 466-485)"""
    freedom = [[-0.7, -0.81, -0.65, -0.65], [-0.67, -0.72, -0.8, -0.65], [-0.69, -0.87, -0.72, -0.81],
               [-0.61, -0.69, -0.67, -0.7]]
    penalty = [[0.4, 0.575, 0.33, 0.73], [0.23, 0.32, 0.17, 0.33], [0.41, 0.45, 0.32, 0.575], [0.33, 0.41, 0.23, 0.4]]
    hbonds = [[2, 2.5, 2.5, 2], [2.5, 3, 3, 2.5], [2.5, 3, 3, 2.5], [2, 2.5, 2.5, 2]]
    c = []
    for i in range(4):
        for j in range(4):
            c.append(freedom[i][j] * hbonds[i][j] + penalty[i][j])
    c += [0.98, 1.03, 1.03, 0.98]                       # adjust_initiation A, C, G, T
    c.append(0.4)                                       # adjust_terminal_TA
    Na = 50
    c.append(0.175 * math.log(Na / 1000, math.e) + 0.20)
    c.append(0.4)                                       # symmetry_correction
    c.append(_dg_round_threshold())
    return tuple(c)


def dimer_flags(ctx, sets_list, threshold: float = 3.0, min_end: int = 5, max_end: int = 18) -> np.ndarray:
    """core:487-503 dimer_check for every primer of the list against itself -> bool array"""
    if not sets_list:
        return np.zeros(0, bool)
    eng = _lib.Dimer(ctx, sets_list, min_end, max_end, True, loss_table(threshold), dg_consts())
    try:
        idx = np.arange(len(sets_list), dtype=np.int32)
        hit, _ = eng.pairs(idx, idx)
    finally:
        eng.close()
    return hit >= 0
