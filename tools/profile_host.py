"""cProfile of one resident step of the bench workload (host-side hot spots)"""
import cProfile
import pstats
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from multiprime_b200 import core, synth
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
codes = synth.synth_codes_parallel(n, 600)
packed = core.pack4(codes)
ids = synth.seq_ids(n)
app = core.NN_degenerate(seq_file=None, outfile="", packed=(ids, packed, 600, None), sidecars=False, **bench.PARAMS)
pos = list(range(app.start_position, app.stop_position - bench.K))
app.design(pos)
pr = cProfile.Profile()
pr.enable()
app.design(pos)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
