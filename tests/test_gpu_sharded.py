"""GPU suite, sequence shards: two processes share cuda:0 (gloo for the collectives), each holding half of the rows of
a golden case in the REAL library; both must produce the reference's rows.  (The NCCL path of the same code is what
`bench.py --gpus N` runs.)"""
import os
import queue
import socket
import time

import pytest
import torch.multiprocessing as mp

from tests.helpers import load_case

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, q):
    import numpy as np
    import torch.distributed as dist
    from multiprime_b200 import core
    from multiprime_b200.comm import TorchComm
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
                             comm=TorchComm(), device=0, **case["params"])
    recs = case["records"]
    got = {r["row"][0]: r for r in app.design([r["pos"] for r in recs])}
    bad = []
    for rec in recs:
        g = got.get(rec["pos"])
        if (g is None) != (rec["row"] is None) or (g is not None and (g["row"] != rec["row"] or g["trace"] != rec["trace"])):
            bad.append((rec["pos"], g and g["row"], rec["row"]))
    q.put((rank, app.start_position, app.stop_position, bad))
    app.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["synth300", "c2_k18"])
def test_two_shards_one_gpu(name):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, name, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = []
    t0 = time.time()
    while len(res) < 2 and time.time() - t0 < 300:
        try:
            res.append(q.get(timeout=2))
        except queue.Empty:
            if any(p.exitcode not in (None, 0) for p in procs):
                break
    for p in procs:
        p.join(10)
        if p.is_alive():
            p.kill()
    assert len(res) == 2, "a rank died: exit codes %s" % [p.exitcode for p in procs]
    case = load_case(name)
    for rank, start, stop, bad in res:
        assert (start, stop) == (case["start"], case["stop"])
        assert not bad, (rank, bad[:3])


def _shard_rows(name, rank, world, comm, _expect_peer=None, **extra):
    """design() of one shard of a golden case -> (start, stop, mismatching records)"""
    import numpy as np
    from multiprime_b200 import core
    from tests.helpers import case_alignment
    from tests.parity import alignment_arrays
    case = load_case(name)
    ids, seqs = case_alignment(case, name)
    n = len(ids)
    lo, hi = rank * n // world, (rank + 1) * n // world
    _, codes, lens = alignment_arrays(ids[lo:hi], seqs[lo:hi])
    full_cols = max(len(s) for s in seqs)
    if codes.shape[1] < full_cols:
        codes = np.pad(codes, ((0, 0), (0, full_cols - codes.shape[1])))
    app = core.NN_degenerate(seq_file=None, nproc=1, outfile="", alignment=(ids[lo:hi], codes, lens), row0=lo,
                             comm=comm, **extra, **case["params"])
    if _expect_peer is not None:
        assert (app.peer is not None) == _expect_peer
    recs = case["records"]
    got = {r["row"][0]: r for r in app.design([r["pos"] for r in recs])}
    bad = []
    for rec in recs:
        g = got.get(rec["pos"])
        if (g is None) != (rec["row"] is None) or (g is not None and (g["row"] != rec["row"] or g["trace"] != rec["trace"])):
            bad.append((rec["pos"], g and g["row"], rec["row"]))
    res = (app.start_position, app.stop_position, bad, sum(1 for r in recs if r["row"] is not None))
    app.close()
    return res


@pytest.mark.parametrize("name,world", [("synth_iupac", 2), ("c2_k18", 2), ("c3_tmsa", 3)])
def test_device_collectives_loopback(name, world):
    """the DEVICE branches of the sharded path on one GPU: shards on threads, collectives on device tensors
    (export_dev -> all-to-all -> merge from device pointers; in-place all-reduce of the walk's count vector)"""
    from tests.loopback_comm import run_shards
    res = run_shards(world, lambda rank, comm: _shard_rows(name, rank, world, comm, device=0))
    case = load_case(name)
    for rank, (start, stop, bad, n_acc) in enumerate(res):
        assert (start, stop) == (case["start"], case["stop"])
        assert not bad, (rank, bad[:3])
        assert n_acc > 0


def test_peer_allreduce_kernel():
    """k_peer_allreduce on one GPU: four group members on ONE stream, the two phases of every round played in turn (all
    ranks push and signal, then all ranks wait and sum) — rounds of different lengths, the sequence numbers and slot
    parities carrying over from round to round, the sums compared with numpy.  (Across processes the same kernel runs
    both phases at once: test_two_ranks_nccl.)"""
    import numpy as np
    from multiprime_b200 import _lib
    world, cap, rounds = 4, 4096, [1, 5, 1000, 4096, 7, 4096, 33]
    rng = np.random.default_rng(5)
    ctx = _lib.Context(0)
    peers = [_lib.Peer(ctx, r, world, cap) for r in range(world)]
    handles = np.stack([p.handle() for p in peers])
    for p in peers:
        p.connect(handles)
    bufs = [_lib.DevBuf(ctx, (cap,), np.int64) for _ in range(world)]
    for n in rounds:
        data = [rng.integers(0, 1 << 40, n).astype(np.int64) for _ in range(world)]
        for r in range(world):
            _lib.check(_lib.load().mpb_ctx_memcpy(ctx.h, _lib.C.c_void_p(bufs[r].p), _lib.ptr(data[r]), n * 8))
        for r in range(world):
            peers[r].allreduce(bufs[r].p, n, phases=1)
        for r in range(world):
            peers[r].allreduce(bufs[r].p, n, phases=2)
        want = sum(data)
        for r in range(world):
            assert (bufs[r].to_host()[:n] == want).all(), (r, n)
    with pytest.raises(_lib.MpbError):
        peers[0].allreduce(bufs[0].p, cap + 1)
    for b in bufs:
        b.close()
    for p in peers:
        p.close()
    ctx.close()


def _nccl_worker(rank, world, port, name, q):
    import torch
    import torch.distributed as dist
    from multiprime_b200.comm import TorchComm
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    res = _shard_rows(name, rank, world, TorchComm(torch.device("cuda", rank)), device=rank,
                      stream=torch.cuda.current_stream().cuda_stream, _expect_peer=os.environ.get("MPB_PEER", "1") != "0")
    q.put((rank,) + res)
    dist.destroy_process_group()


@pytest.mark.parametrize("name,peer", [("synth_iupac", "1"), ("c2_k18", "1"), ("c2_k18", "0")])
def test_two_ranks_nccl(name, peer, monkeypatch):
    """the NCCL path bench.py --gpus N runs: two ranks on two GPUs, rows and traces of the golden case; the walk's counts
    summed through peer memory opened over CUDA IPC (MPB_PEER=1, the default) or by NCCL all-reduces (MPB_PEER=0)"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    monkeypatch.setenv("MPB_PEER", peer)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, name, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = []
    t0 = time.time()
    while len(res) < 2 and time.time() - t0 < 300:
        try:
            res.append(q.get(timeout=2))
        except queue.Empty:
            if any(p.exitcode not in (None, 0) for p in procs):
                break
    for p in procs:
        p.join(10)
        if p.is_alive():
            p.kill()
    assert len(res) == 2, "a rank died: exit codes %s" % [p.exitcode for p in procs]
    case = load_case(name)
    for rank, start, stop, bad, n_acc in res:
        assert (start, stop) == (case["start"], case["stop"])
        assert not bad, (rank, bad[:3])


def test_cli_under_torchrun(tmp_path):
    """`torchrun ... scripts/multiPrime-core.py`: two ranks (sharing cuda:0, gloo) write the reference CLI's files"""
    import json
    import subprocess
    import sys
    from multiprime_b200 import synth
    from tests.helpers import GOLDEN
    g = json.load(open(os.path.join(GOLDEN, "cli_core.json")))
    n, L, seed, gr, ir = g["synth"]
    fa = tmp_path / "in.fa"
    synth.write_fasta(str(fa), synth.synth_codes(n, L, seed=seed, gap_rate=gr, iupac_rate=ir))
    out = tmp_path / "mine.out"
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, MPB_DIST_BACKEND="gloo")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
                          os.path.join(root, "scripts", "multiPrime-core.py"), "-i", str(fa), "-o", str(out)] + g["args"],
                         capture_output=True, text=True, env=env, timeout=600)
    assert res.returncode == 0, res.stderr[-3000:]
    assert out.read_text() == g["tsv"]
    assert json.load(open(str(out) + ".non_coverage_seq_id_json")) == g["non_cov"]
    assert json.load(open(str(out) + ".gap_seq_id_json")) == g["gap"]
