"""Deterministic synthetic MSA generator (the "10^6 x 600" north-star workload and its small siblings).

Mutation model after SURVEY.md section 8(d) "C4 synthetic" (conserved / variable 60-column blocks,
8 clades, terminal and internal gap runs, sparse 2-fold IUPAC cells), but drawn in fixed row chunks
of CHUNK rows, each chunk from its own PCG64 stream keyed by (seed, chunk index), so that any row
range can be produced without generating the rows before it (needed to shard 10^6 rows over ranks).

Cells are returned as 4-bit base sets: A=1, C=2, G=4, T=8, IUPAC = OR of its bases, gap = 0.
"""
from __future__ import annotations

import numpy as np

CHUNK = 8192
CODE_CHARS = "-ACMGRSVTWYHKDBN"          # index = 4-bit set (A=1,C=2,G=4,T=8)
_CLADE_P = [.40, .20, .12, .10, .08, .05, .03, .02]
# 2-fold codes containing base b (A,C,G,T): A->R,M,W  C->Y,M,S  G->R,K,S  T->Y,K,W
_TWOFOLD = np.array([[1 | 4, 1 | 2, 1 | 8], [2 | 8, 1 | 2, 2 | 4], [1 | 4, 4 | 8, 2 | 4], [2 | 8, 4 | 8, 1 | 8]],
                    dtype=np.uint8)


def _plan(seed: int, n_col: int):
    rng = np.random.Generator(np.random.PCG64([seed, 0xC4]))
    root = rng.integers(0, 4, n_col)
    mu = np.where((np.arange(n_col) // 60) % 2 == 0, 0.01, 0.12)
    cols = np.stack([rng.choice(n_col, 6, replace=False) for _ in range(8)])
    subs = np.stack([rng.integers(1, 4, 6) for _ in range(8)])
    return root, mu, cols, subs


def synth_codes(n_seq: int, n_col: int = 600, seed: int = 20240923, row0: int = 0,
                gap_rate: float = 0.002, iupac_rate: float = 1e-4, term_gap: float = 0.05) -> np.ndarray:
    """rows [row0, row0+n_seq) of the synthetic alignment as uint8 4-bit sets, shape (n_seq, n_col)"""
    root, mu, ccols, csubs = _plan(seed, n_col)
    out = np.empty((n_seq, n_col), dtype=np.uint8)
    r = row0
    while r < row0 + n_seq:
        c = r // CHUNK
        lo, hi = c * CHUNK, (c + 1) * CHUNK
        block = _chunk(seed, c, n_col, root, mu, ccols, csubs, gap_rate, iupac_rate, term_gap)
        a, b = max(r, lo), min(row0 + n_seq, hi)
        out[a - row0:b - row0] = block[a - lo:b - lo]
        r = b
    return out


def _chunk(seed, c, n_col, root, mu, ccols, csubs, gap_rate, iupac_rate, term_gap):
    rng = np.random.Generator(np.random.PCG64([seed, 1, c]))
    n = CHUNK
    clade = rng.choice(8, n, p=_CLADE_P)
    x = np.broadcast_to(root.astype(np.uint8), (n, n_col)).copy()
    for k in range(8):
        rows = np.nonzero(clade == k)[0]
        x[np.ix_(rows, ccols[k])] = ((root[ccols[k]] + csubs[k]) % 4).astype(np.uint8)
    mut = rng.random((n, n_col), dtype=np.float32) < mu.astype(np.float32)
    shift = rng.integers(1, 4, (n, n_col), dtype=np.uint8)
    x = np.where(mut, (x + shift) % 4, x).astype(np.uint8)
    code = (np.uint8(1) << x).astype(np.uint8)
    # terminal gap runs
    lead = rng.random(n) < term_gap
    trail = rng.random(n) < term_gap
    lead_len = rng.integers(1, 31, n)
    trail_len = rng.integers(1, 31, n)
    col = np.arange(n_col)
    code[(col[None, :] < (lead_len * lead)[:, None])] = 0
    code[(col[None, :] >= (n_col - trail_len * trail)[:, None])] = 0
    # internal gap runs, geometric length with mean 3
    starts = np.argwhere(rng.random((n, n_col), dtype=np.float32) < np.float32(gap_rate / 3))
    glen = rng.geometric(1 / 3, len(starts))
    for d in range(int(glen.max()) if len(glen) else 0):
        sel = (glen > d) & (starts[:, 1] + d < n_col)
        code[starts[sel, 0], starts[sel, 1] + d] = 0
    # sparse 2-fold IUPAC cells
    amb = np.argwhere((rng.random((n, n_col), dtype=np.float32) < np.float32(iupac_rate)) & (code != 0))
    pick = rng.integers(0, 3, len(amb))
    base = x[amb[:, 0], amb[:, 1]]
    code[amb[:, 0], amb[:, 1]] = _TWOFOLD[base, pick]
    return code


def synth_codes_parallel(n_seq: int, n_col: int = 600, seed: int = 20240923, row0: int = 0, procs: int = 0,
                         **kw) -> np.ndarray:
    """synth_codes() with the chunks drawn by a process pool (same result, chunk streams are independent)"""
    import os
    from concurrent.futures import ProcessPoolExecutor
    procs = procs or min(32, os.cpu_count() or 1)
    bounds = list(range(row0 - row0 % CHUNK, row0 + n_seq, CHUNK))
    jobs = [(max(b, row0), min(b + CHUNK, row0 + n_seq)) for b in bounds]
    if procs <= 1 or len(jobs) <= 1:
        return synth_codes(n_seq, n_col, seed, row0, **kw)
    out = np.empty((n_seq, n_col), dtype=np.uint8)
    # (forks: call this before CUDA / NCCL threads exist in the process, as bench.py does)
    with ProcessPoolExecutor(procs) as ex:
        futs = [(a, b, ex.submit(synth_codes, b - a, n_col, seed, a, **kw)) for a, b in jobs]
        for a, b, f in futs:
            out[a - row0:b - row0] = f.result()
    return out


def codes_to_strings(codes: np.ndarray) -> list[str]:
    lut = np.frombuffer(CODE_CHARS.encode(), dtype=np.uint8)
    return [row.tobytes().decode() for row in lut[codes]]


def seq_ids(n_seq: int, row0: int = 0) -> list[str]:
    return [">s%07d" % (row0 + i) for i in range(n_seq)]


def write_fasta(path: str, codes: np.ndarray, row0: int = 0) -> None:
    with open(path, "w") as fh:
        for sid, s in zip(seq_ids(len(codes), row0), codes_to_strings(codes)):
            fh.write(sid + "\n" + s + "\n")
