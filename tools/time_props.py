"""micro-timing of mpb_primer_props and mpb_scan host overhead on the bench workload's final primers"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from multiprime_b200 import core, synth, _lib
import bench
n = 200000
codes = synth.synth_codes_parallel(n, 600)
app = core.NN_degenerate(seq_file=None, outfile="", packed=(synth.seq_ids(n), core.pack4(codes), 600, None), sidecars=False, **bench.PARAMS)
pos = list(range(app.start_position, app.stop_position - bench.K))
recs = app.design(pos)
from multiprime_b200.iupac import sets_of
arr = np.zeros((len(recs), 32), np.uint8)
for i, r in enumerate(recs):
    arr[i, :18] = sets_of(r["row"][3])
for rep in range(3):
    t = time.perf_counter(); out = app.ctx.primer_props(arr, 18, 0.2, 0.7, 4, core.TM_CONSTS); dt = time.perf_counter() - t
    print("props ms", 1000 * dt, "n", len(recs), "deg sum", int(out[3].sum()), "flags64/128", int((out[2] & 64 > 0).sum()), int((out[2] & 128 > 0).sum()))
t = time.perf_counter(); x = app._primer_props(arr, 18, 0.2, 0.7); print("_primer_props ms", 1000 * (time.perf_counter() - t))
