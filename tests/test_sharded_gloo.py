"""world_size-2 run of the sequence-sharded path on CPU (gloo + the fake device): every rank must produce the rows
of the single-process run, with one all-reduce per scan round and an all-gather of haplotype entries per batch."""
import os
import socket

import pytest
import torch.multiprocessing as mp

from tests.helpers import load_case


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, limit, q):
    import numpy as np
    import torch.distributed as dist
    from multiprime_b200 import core
    from multiprime_b200.comm import TorchComm
    from tests import fake_device
    from tests.helpers import case_alignment
    from tests.parity import alignment_arrays
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    case = load_case(name)
    ids, seqs = case_alignment(case, name)
    n = len(ids)
    lo, hi = rank * n // world, (rank + 1) * n // world
    _, codes, lens = alignment_arrays(ids[lo:hi], seqs[lo:hi])
    full_cols = max(len(s) for s in seqs)
    if codes.shape[1] < full_cols:
        codes = np.pad(codes, ((0, 0), (0, full_cols - codes.shape[1])))
    app = core.NN_degenerate(seq_file=None, nproc=1, outfile="", alignment=(ids[lo:hi], codes, lens), row0=lo,
                             comm=TorchComm(), _backend=fake_device, **case["params"])
    recs = case["records"][:limit]
    got = {r["row"][0]: r for r in app.design([r["pos"] for r in recs])}
    bad = []
    for rec in recs:
        g = got.get(rec["pos"])
        if (g is None) != (rec["row"] is None) or (g is not None and (g["row"] != rec["row"] or g["trace"] != rec["trace"])):
            bad.append((rec["pos"], g and g["row"], rec["row"]))
    # the JSON side files of a sharded run: shards hold disjoint id lists, merged in rank order (core.run)
    from tests.helpers import digest
    non_cov = {r["row"][0]: r["non_cov"] for r in got.values()}
    gap_ids = {r["row"][0]: r["gap_ids"] for r in got.values()}
    non_cov, gap_ids = core._merge_sidecars(app.comm.allgather_object((non_cov, gap_ids)))
    for rec in recs:
        if rec["row"] is not None and (digest(non_cov[rec["pos"]][0]) != rec["f_non"] or
                                       digest(non_cov[rec["pos"]][1]) != rec["r_non"] or
                                       digest(gap_ids[rec["pos"]]) != rec["gap_ids"]):
            bad.append((rec["pos"], "side files"))
    q.put((rank, app.start_position, app.stop_position, bad, sum(1 for r in recs if r["row"] is not None)))
    dist.destroy_process_group()


@pytest.mark.parametrize("name,limit", [("synth_iupac", 60), ("c3_tmsa", 12)])
def test_two_ranks_equal_single(name, limit):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, name, limit, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = []
    import queue
    import time
    t0 = time.time()
    while len(res) < len(procs) and time.time() - t0 < 300:
        try:
            res.append(q.get(timeout=2))
        except queue.Empty:
            if any(p.exitcode not in (None, 0) for p in procs):
                break
    for p in procs:
        p.join(10)
        if p.is_alive():
            p.kill()
    assert len(res) == len(procs), "a rank died: exit codes %s" % [p.exitcode for p in procs]
    case = load_case(name)
    for rank, start, stop, bad, n_acc in res:
        assert (start, stop) == (case["start"], case["stop"])
        assert not bad, (rank, bad[:3])
        assert n_acc > 0


@pytest.mark.parametrize("name,world", [("synth_iupac", 3), ("c3_tmsa", 4)])
def test_window_owners_threads(name, world):
    """window ownership with 3 and 4 shards (windows per owner differ by one, some owners idle on short batches): the
    shards run on threads with the loop-back communicator and the stand-in device"""
    from tests import fake_device
    from tests.loopback_comm import run_shards
    from tests.test_gpu_sharded import _shard_rows
    res = run_shards(world, lambda rank, comm: _shard_rows(name, rank, world, comm, _backend=fake_device))
    case = load_case(name)
    for rank, (start, stop, bad, n_acc) in enumerate(res):
        assert (start, stop) == (case["start"], case["stop"])
        assert not bad, (rank, bad[:3])
        assert n_acc > 0


def test_exception_records_beyond_the_inline_room(monkeypatch):
    """gap rows holding IUPAC cells travel with the counter gather; a shard with more of them than the inline room
    triggers the second, padded gather — same rows either way"""
    from multiprime_b200 import core
    from tests import fake_device
    from tests.loopback_comm import run_shards
    from tests.test_gpu_sharded import _shard_rows
    monkeypatch.setattr(core.NN_degenerate, "EXC_INLINE", 1)
    world, name = 3, "synth_iupac"
    res = run_shards(world, lambda rank, comm: _shard_rows(name, rank, world, comm, _backend=fake_device))
    case = load_case(name)
    for rank, (start, stop, bad, n_acc) in enumerate(res):
        assert (start, stop) == (case["start"], case["stop"])
        assert not bad, (rank, bad[:3])
