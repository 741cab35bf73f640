"""IUPAC base-set helpers shared by the host modules.  A cell / primer position is a 4-bit set:
A=1, C=2, G=4, T=8 (so bases A,C,G,T are bit indices 0..3, the index order of the reference's tensors, core:185)."""
from __future__ import annotations

from itertools import product

BASES = "ACGT"
CODE_CHARS = "-ACMGRSVTWYHKDBN"                # index = 4-bit set
CHAR_CODE = {c: i for i, c in enumerate(CODE_CHARS)}
FOLD = [bin(i).count("1") for i in range(16)]
# alternatives of each set in the reference's expansion order (core:105-107), as base indices
ORDER = {
    0: (), 1: (0,), 2: (1,), 4: (2,), 8: (3,),
    5: (0, 2), 10: (1, 3), 3: (0, 1), 12: (2, 3), 6: (2, 1), 9: (0, 3),
    11: (0, 3, 1), 14: (2, 3, 1), 7: (2, 0, 1), 13: (2, 0, 3), 15: (0, 3, 2, 1),
}


def sets_of(primer: str) -> list[int]:
    return [CHAR_CODE[c] for c in primer]


def primer_string(sets) -> str:
    return "".join(CODE_CHARS[s] for s in sets)


def degeneracy(sets) -> int:
    """core:210-211 score_trans"""
    d = 1
    for s in sets:
        d *= FOLD[s]
    return d


def n_degenerate(sets) -> int:
    """core:214-215 dege_number"""
    return sum(1 for s in sets if FOLD[s] > 1)


def allow_masks(sets) -> list[int]:
    """[mask_A, mask_C, mask_G, mask_T]: bit i set when the base is allowed at position i"""
    m = [0, 0, 0, 0]
    for i, s in enumerate(sets):
        for b in range(4):
            if (s >> b) & 1:
                m[b] |= 1 << i
    return m


def comp_set(s: int) -> int:
    """complement of a base set (A<->T, C<->G) = the nibble bit-reversed"""
    return ((s & 1) << 3) | ((s & 2) << 1) | ((s & 4) >> 1) | ((s & 8) >> 3)


def rc_sets(sets) -> list[int]:
    return [comp_set(s) for s in reversed(sets)]


def expand_keys(sets) -> list[tuple]:
    """all plain sequences (tuples of base indices) in the reference's product order (core:368-380)"""
    return list(product(*[ORDER[s] for s in sets]))


def expand_array(sets):
    """expand_keys as a (number of expansions, len(sets)) uint8 array, same order (leftmost position slowest)"""
    import numpy as np
    opts = [np.array(ORDER[s], np.uint8) for s in sets]
    n = 1
    for o in opts:
        n *= len(o)
    out = np.empty((n, len(sets)), np.uint8)
    rep = n
    for j, o in enumerate(opts):          # column j: every option repeated `rep` times, the pattern tiled
        rep //= len(o)
        out[:, j] = np.tile(np.repeat(o, rep), n // (rep * len(o)))
    return out


def expand_strings(sets) -> list[str]:
    """like expand_keys, as strings; a gap cell (0) stays '-'"""
    return ["".join(t) for t in product(*[[BASES[b] for b in ORDER[s]] if s else ["-"] for s in sets])]
