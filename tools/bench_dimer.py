#!/usr/bin/env python
"""secondary metric (BASELINE.json configs[4]): primer pairs per second of the all-pairs dimer grid (finDimer) on P
synthetic 18-mer primers with up to 3 degenerate positions; prints one JSON line"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from multiprime_b200 import _lib
from multiprime_b200.dimer import dg_consts, loss_table

P = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
rng = np.random.default_rng(5)
sets = (1 << rng.integers(0, 4, (P, 18))).astype(np.uint8)
amb = rng.random((P, 18)) < 0.06                               # sprinkle 2-fold codes
sets[amb] |= (1 << rng.integers(0, 4, int(amb.sum()))).astype(np.uint8)
sets_list = [row.tolist() for row in sets]
ctx = _lib.Context(0)
t0 = time.perf_counter()
eng = _lib.Dimer(ctx, sets_list, 5, 18, True, loss_table(3.96), dg_consts())
t_prep = time.perf_counter() - t0
band = max(1, min(P, (1 << 25) // P * 8))
t0 = time.perf_counter()
hits = tested = 0
for r0 in range(0, P, band):
    hi, hj, ho, hd, nt = eng.grid(r0, min(P, r0 + band), max_hits=1 << 24)
    hits += len(hi)
    tested += nt
dt = time.perf_counter() - t0
pairs = P * (P + 1) // 2
print(json.dumps({"metric": "dimer_pairs_per_sec", "value": pairs / dt, "unit": "pairs/s", "primers": P, "pairs": pairs,
                  "seconds": dt, "prepare_seconds": t_prep, "pairs_after_5mer_prefilter": tested, "dimer_pairs": hits,
                  "launches": ctx.launches}))
eng.close()
ctx.close()
