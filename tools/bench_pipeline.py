"""core -> pairing on the north-star workload without files in between: the candidate scan leaves the
per-sequence coverage bit vectors of every chosen primer in HBM and the pairing step (get_multiPrime semantics) takes its
pair coverage from them.  Prints one JSON line: pairs/s of the pair-coverage kernel and of the whole pairing step.
usage: python tools/bench_pipeline.py [n_seq] [steps]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiprime_b200 import core, pairing, synth
import bench

n_seq = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
packed = core.pack4(synth.synth_codes_parallel(n_seq, 600))
ids = synth.seq_ids(n_seq)
app = core.NN_degenerate(seq_file=None, outfile="", packed=(ids, packed, 600, None), device=0, sidecars=False,
                         want_trace=False, keep_bits=True, **bench.PARAMS)
pos = list(range(app.start_position, app.stop_position - bench.K))
out = "/tmp/pipeline.candidate.primers.txt"
res = {}
for it in range(steps + 1):
    t0 = time.perf_counter()
    recs = app.design(pos)
    t1 = time.perf_counter()
    devnull = open(os.devnull, "w")
    so, sys.stdout = sys.stdout, devnull                     # the reference prints one line per start candidate
    try:
        pf = pairing.Primers_filter.from_core(app, recs, out, adaptor=",", size="150,400", fraction=0.6, diff_Tm=4,
                                              position=0, distance=4, GC="0.1,0.9")
        app.ctx.profile_read(None)
        app.ctx.profile(True)
        rows = pf.run()
        ms, n_launch, units = app.ctx.profile_read("k_pair_cover3")
        app.ctx.profile(False)
    finally:
        sys.stdout = so
    t2 = time.perf_counter()
    res = {"core_ms": 1000 * (t1 - t0), "pairing_ms": 1000 * (t2 - t1), "rows_in": len(recs), "pairs_tested": units,
           "pairs_out": len(rows), "k_pair_cover3_ms": ms,
           "pair_cover_GBps": units * 4 * ((n_seq + 31) // 32) * 4 / (ms / 1000) / 1e9 if ms else None}
print(json.dumps({"workload": "synthetic %d x 600, core -> pairing in one process, bit vectors in HBM" % n_seq, **res,
                  "pairs_per_s_kernel": res["pairs_tested"] / (res["k_pair_cover3_ms"] / 1000) if res["k_pair_cover3_ms"] else None,
                  "pairs_per_s_step": res["pairs_tested"] / (res["pairing_ms"] / 1000)}))
