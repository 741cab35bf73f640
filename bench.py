#!/usr/bin/env python
"""bench.py — candidate x sequence mismatch evaluations per second of the degenerate-primer candidate scan.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run, one rank/GPU)
    python bench.py --impl reference ...                     (the CPU arm: the oracle port on the host cores)

Workload (BASELINE.json configs[3], the configuration the metric is quoted on): synthetic 10^6-sequence x 600-column
alignment (multiprime_b200/synth.py), k=18, degeneracy <= 256 (-n 8), <= 3 mismatches, other flags default.
One step = one full pass of the hot path over every window of the conserved region: window k-mer extraction +
haplotype tables, gates, base/dinucleotide tensors, seeds, the NN-array refinement walk with one candidate scan per
round, Tm, filters -> the rows of the reference's .out TSV.  `value` counts exactly the evaluations the reference
makes: (calls to mis_primer_check) x (sequences), summed over windows, divided by the step time.
With N GPUs every rank holds n_seq sequences of the same synthetic family (weak scaling).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K, DNUM, DEG, VAR = 18, 8, 256, 3
PARAMS = dict(primer_length=K, coverage=0.8, number_of_dege_bases=DNUM, score_of_dege_bases=DEG, product_len=100,
              position="1,2,-1", variation=VAR, raw_entropy_threshold=3.6, distance=4, GC="0.2,0.7", nproc=1)
BYTES_PER_EVAL = K / 2 + 0.25                     # one k-column window in 4-bit cells + 2 result bits


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n-seq", type=int, default=1_000_000, help="sequences per GPU")
    ap.add_argument("--n-col", type=int, default=600)
    ap.add_argument("--cpu-sample-seqs", type=int, default=50000)
    ap.add_argument("--cpu-sample-windows", type=int, default=32)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)"""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 7:
                    self.rows.append(f)
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows)}


def scan_traffic():
    """DRAM bytes (read + write) of one k_scan launch from the committed `ncu --set full` capture"""
    path = os.path.join(ROOT, "profiles", "k_scan_traffic.json")
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh)["traffic_bytes_per_launch"]
    return None


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (the Python reference cannot travel to the GPU box)
# ----------------------------------------------------------------------------------------------------------
_CPU_DATA = {}


def _oracle_window(p):
    """one window of the bounded sample through the oracle port; the alignment is inherited from the parent (fork)"""
    from oracle import mp_oracle as o
    prm = o.Params(k=K, dnum=DNUM, degeneracy=DEG, variation=VAR, entropy=3.6, gc="0.2,0.7", size=100, fraction=0.8,
                   coordinate="1,2,-1", away=4)
    trace = []
    o.design_window(_CPU_DATA["ids"], _CPU_DATA["seqs"], p, prm, 3.6, trace)
    return len(trace)


def cpu_sample(n_seq: int, n_col: int, n_windows: int, procs: int):
    """oracle over a bounded sample: the first n_seq synthetic sequences, n_windows windows spread over the region.
    Returns (evals, seconds).  Windows are dealt to `procs` forked workers (the reference itself is single-process:
    its pool is inert, core:1143; this is the best case for the CPU side)."""
    from multiprime_b200 import synth
    from oracle import mp_oracle as o
    if _CPU_DATA.get("key") != (n_seq, n_col):
        codes = synth.synth_codes(n_seq, n_col)
        _CPU_DATA.update(key=(n_seq, n_col), ids=synth.seq_ids(n_seq), seqs=synth.codes_to_strings(codes))
    start, stop = o.region(_CPU_DATA["seqs"], 0.8)
    all_pos = list(range(start, stop - K))
    pos = [all_pos[int(i * (len(all_pos) - 1) / max(1, n_windows - 1))] for i in range(n_windows)]
    t0 = time.perf_counter()
    if procs <= 1:
        calls = [_oracle_window(p) for p in pos]
    else:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(procs) as pool:
            calls = pool.map(_oracle_window, pos, chunksize=1)
    dt = time.perf_counter() - t0
    return sum(calls) * n_seq, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_windows = max(args.cpu_sample_windows, 2 * cores)            # keep every core busy
    procs = min(cores, n_windows)
    vals = []
    for i in range(args.warmup + args.steps):
        ev, dt = cpu_sample(args.cpu_sample_seqs, args.n_col, n_windows, procs)
        if i >= args.warmup:
            vals.append((ev, dt))
    ev = sum(v[0] for v in vals)
    dt = sum(v[1] for v in vals)
    value = ev / dt
    sample = "oracle port (oracle/mp_oracle.py), first %d synthetic sequences x %d windows spread over the region, " \
             "%d worker processes" % (args.cpu_sample_seqs, n_windows, procs)
    line = {"impl": "reference", "metric": "candidate_x_sequence_evals_per_sec", "value": value, "unit": "evals/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": "synthetic MSA %dx%d k=%d d<=%d v<=%d (bounded sample)" %
                       (args.n_seq, args.n_col, K, DEG, VAR)},
            "cpu_baseline": {"value": value, "unit": "evals/s", "cores": procs, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------------
def run_b200(args):
    from multiprime_b200 import core, synth

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    n_seq, n_col = args.n_seq, args.n_col
    # synthetic input first: the generator forks worker processes, which must happen before CUDA / NCCL threads exist
    codes = synth.synth_codes_parallel(n_seq, n_col, row0=rank * n_seq, procs=max(1, (os.cpu_count() or 8) // world))
    packed = core.pack4(codes)
    del codes
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pinned = torch.from_numpy(packed).pin_memory()
    packed_pinned = pinned.numpy()
    ids = synth.seq_ids(n_seq, rank * n_seq)
    stream = torch.cuda.current_stream().cuda_stream

    comm = None
    if world > 1:
        from multiprime_b200.comm import TorchComm
        comm = TorchComm(torch.device("cuda", local))

    def make_app():
        return core.NN_degenerate(seq_file=None, outfile="", packed=(ids, packed_pinned, n_col, None), device=local,
                                  sidecars=False, want_trace=False, keep_bits=True, stream=stream, comm=comm,
                                  row0=rank * n_seq, **PARAMS)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    app = make_app()
    positions = list(range(app.start_position, app.stop_position - K))
    h2d = packed_pinned.nbytes

    def step_resident():
        return app.design(positions)

    e2e_init = {}

    def step_e2e():
        a = make_app()                                  # H2D of the packed alignment + plane build + region
        for kk, vv in a.init_ms.items():
            e2e_init[kk] = e2e_init.get(kk, 0.0) + vv
        recs = a.design(list(range(a.start_position, a.stop_position - K)))
        a.close()
        return recs

    results = {}
    sampler = ClockSampler(local)
    for name, fn in (("value", step_resident), ("e2e", step_e2e)):
        for _ in range(args.warmup):
            fn()
        e2e_init.clear()
        app.ctx.profile_read(None)
        app.ctx.profile(name == "value")
        app.stats.update(evals=0, scan_calls=0, candidates=0, phase_ms={})
        launches0 = app.ctx.launches
        if name == "value":
            sampler.start()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        nrows = 0
        for _ in range(args.steps):
            nrows = len(fn())
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = max(e0.elapsed_time(e1), 0.0)
        ms = max(ms, 1000 * wall) if name == "e2e" else ms
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        results[name] = {"ms": float(t.item()), "rows": nrows}
        if name == "value":
            sampler.stop_flag.set()
            results["launches"] = app.ctx.launches - launches0
            results["scan"] = app.ctx.profile_read("k_cscan")
            results["kernel_ms"] = {kn: app.ctx.profile_read(kn)[0] / args.steps for kn in
                                    ("k_prefilter", "k_prefilter_sums", "k_hist", "k_hist_summary", "k_hist_match",
                                     "k_cscan", "k_cscan_plan", "k_cscan_special", "k_walk_advance", "k_walk_seed",
                                     "k_tm", "k_dimer_pairs", "k_dimer_expand", "k_dimer_ends")}
            results["hist"] = app.ctx.profile_read("k_hist")
            results["evals_per_step"] = app.stats["evals"] / args.steps
            results["scan_calls"] = app.stats["scan_calls"] / args.steps
            results["candidates"] = app.stats["candidates"] / args.steps
            results["phases"] = {k: round(v / args.steps, 2) for k, v in app.stats["phase_ms"].items()}
            app.ctx.profile(False)
    evals_all = float(results["evals_per_step"])      # already global: calls x (sequences of ALL shards)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    ms_step = results["value"]["ms"] / args.steps
    value = evals_all / (ms_step / 1000)
    e2e_ms = results["e2e"]["ms"] / args.steps
    peak, peak_src = peaks()
    scan_ms, scan_n, scan_units = results["scan"]
    achieved = scan_units * BYTES_PER_EVAL / (scan_ms / 1000) / 1e9 if scan_ms > 0 else 0.0
    line = {
        "metric": "candidate_x_sequence_evals_per_sec", "value": value, "unit": "evals/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": "synthetic MSA %dx%d per GPU (multiprime_b200/synth.py seed 20240923), k=%d, -n %d -d %d "
                               "-v %d, %d windows, %d rows out" % (n_seq, n_col, K, DNUM, DEG, VAR, len(positions),
                                                                  results["value"]["rows"]),
                   "parallelism": "sequence shards x%d, all-reduce of coverage counts per scan round" % world,
                   "l2": "inputs (%.0f MB of bit-planes + GB-sized haplotype tables) exceed the 126 MB L2" %
                         (n_seq * n_col / 2 / 1e6),
                   "evals_per_step": evals_all, "scan_launches_per_step": results["scan_calls"],
                   "scan_candidates_per_step": results["candidates"]},
        "clocks": sampler.summary(),
        "e2e": {"value": evals_all / (e2e_ms / 1000), "unit": "evals/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": int(results["value"]["rows"] * 120), "ms_per_step": e2e_ms,
                "setup_ms_per_step": {kk: round(vv / args.steps, 2) for kk, vv in e2e_init.items()}},
        "gpu_launches": int(results["launches"]),
        "roofline": {"bound": "hbm", "kernel": "k_cscan", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": scan_traffic(), "peak_source": peak_src,
                     "launches": scan_n, "avg_launch_ms": scan_ms / max(1, scan_n),
                     "evals_in_launches": scan_units,
                     "note": "achieved = scanned candidate x sequence pairs x %.2f B / event-timed k_scan time; traffic = "
                             "dram read+write bytes of one launch (ncu --set full, profiles/): far BELOW the algorithmic "
                             "bytes because a loaded window serves all candidates of its window and neighbouring windows "
                             "share words through L1/L2 - the kernel is bound by integer issue (ncu: 77%% issue active, "
                             "2.7%% of peak DRAM throughput)" % BYTES_PER_EVAL},
        "kernels": {"k_hist_ms_per_step": results["hist"][0] / args.steps,
                    "k_scan_ms_per_step": scan_ms / args.steps,
                    "ms_per_step": {kk: round(vv, 3) for kk, vv in results["kernel_ms"].items()}},
        "host_phases_ms_per_step": results["phases"],
    }
    if not args.no_cpu_baseline and world == 1:       # the CPU baseline is timed on rank 0 at N = 1 only
        evc, dtc = cpu_sample(args.cpu_sample_seqs, n_col, args.cpu_sample_windows, 1)
        line["cpu_baseline"] = {"value": evc / dtc, "unit": "evals/s", "cores": 1, "kind": "port",
                                "sample": "oracle/mp_oracle.py, first %d synthetic sequences x %d windows, %.1f s" %
                                          (args.cpu_sample_seqs, args.cpu_sample_windows, dtc)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
