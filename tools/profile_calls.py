"""host time per libmpb200 entry point over a few steps of the bench workload (no profiler: perf_counter around every
ctypes call).  usage: python tools/profile_calls.py [n_seq] [steps]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from multiprime_b200 import _lib, core, synth
import bench

n_seq = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
codes = synth.synth_codes_parallel(n_seq, 600)
packed = core.pack4(codes)
del codes
lib = _lib.load()
acc = {}


class Timed:
    def __init__(self, name, fn):
        self.name, self.fn = name, fn

    def __call__(self, *a):
        t0 = time.perf_counter()
        r = self.fn(*a)
        dt = time.perf_counter() - t0
        e = acc.setdefault(self.name, [0, 0.0])
        e[0] += 1
        e[1] += dt
        return r


class Proxy:
    def __getattr__(self, name):
        return Timed(name, getattr(lib, name))


_lib._lib = Proxy()
ids = synth.seq_ids(n_seq)
app = core.NN_degenerate(seq_file=None, outfile="", packed=(ids, packed, 600, None), device=0, sidecars=False,
                         want_trace=False, keep_bits=True, **bench.PARAMS)
pos = list(range(app.start_position, app.stop_position - bench.K))
for _ in range(2):
    app.design(pos)
acc.clear()
app.stats["phase_ms"] = {}
t0 = time.perf_counter()
for _ in range(steps):
    app.design(pos)
app.ctx.sync()
total = (time.perf_counter() - t0) / steps * 1000
print("ms/step %.2f" % total)
for name, (n, t) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    print("%-28s calls/step %6.1f  ms/step %8.3f" % (name, n / steps, 1000 * t / steps))
print({k: round(v / steps, 2) for k, v in app.stats["phase_ms"].items()})
