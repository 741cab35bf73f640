import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
import torch
from multiprime_b200 import _lib, core, synth
from tests.helpers import load_case, case_alignment
from tests.parity import alignment_arrays
from tests.loopback_comm import run_shards

name = "synth300"
case = load_case(name)
ids, seqs = case_alignment(case, name)
_, codes, lens = alignment_arrays(ids, seqs)
n, L = codes.shape
world = 2
positions0 = [r["pos"] for r in case["records"]][:12]
cap = {}
orig_walk = _lib.Hist.walk
import threading
def walk_spy(self, dnum, degeneracy, fmask, rmask, win_idx, cover_number, mm_key, freq=None, nn=None, comm=None, want_trace=True, lag=2):
    res = orig_walk(self, dnum, degeneracy, fmask, rmask, win_idx, cover_number, mm_key, freq, nn, comm, want_trace, lag)
    cap[threading.get_ident()] = dict(win_idx=np.array(win_idx), cover=np.array(cover_number), mm=np.array(mm_key),
                                      pos=np.array(self.win_pos)[np.array(win_idx)], freq=None if freq is None else np.array(freq), res=res)
    return res
_lib.Hist.walk = walk_spy

def shard(rank, comm):
    lo, hi = rank * n // world, (rank + 1) * n // world
    app = core.NN_degenerate(seq_file=None, nproc=1, outfile="", alignment=(ids[lo:hi], codes[lo:hi], lens[lo:hi]), row0=lo,
                              comm=comm, device=0, **case["params"])
    recs = app.design(positions0)
    return [r["row"] for r in recs], cap[threading.get_ident()]

res = run_shards(world, shard)
app = core.NN_degenerate(seq_file=None, nproc=1, outfile="", alignment=(ids, codes, lens), device=0, **case["params"])
rows1 = [r["row"] for r in app.design(positions0)]
c1 = cap[threading.get_ident()]
print("single rows", rows1[:3])
for r in range(world):
    rows, c = res[r]
    print("rank", r, "rows", rows[:3])
    o1 = np.argsort(c1["pos"]); o2 = np.argsort(c["pos"])
    print(" pos equal", np.array_equal(c1["pos"][o1], c["pos"][o2]), "cover equal", np.array_equal(c1["cover"][o1], c["cover"][o2]),
          "mm equal", np.array_equal(c1["mm"][o1], c["mm"][o2]))
    print(" walk sets equal", np.array_equal(c1["res"]["sets"][o1], c["res"]["sets"][o2]), "counts equal",
          np.array_equal(c1["res"]["counts"][o1], c["res"]["counts"][o2]))
    print(" counts", c["res"]["counts"][o2][:3].tolist(), c1["res"]["counts"][o1][:3].tolist())
    print(" stats", c["res"]["stats"].tolist(), c1["res"]["stats"].tolist())
