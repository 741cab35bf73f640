"""CPU suite: the host logic of multiprime_b200.core driven through tests/fake_device.py (no GPU), compared with
the reference's golden records window by window."""
import pytest

from tests import fake_device
from tests.parity import check_case


@pytest.mark.parametrize("name,limit", [("synth_iupac", 60), ("synth300", 40), ("c2_k18", 12), ("c3_tmsa", 14),
                                        ("c1_testfa", 45)])
def test_host_logic_fake_device(name, limit):
    check_case(name, backend=fake_device, max_windows=limit)


def test_window_cells_native_matches_host_restatement():
    """mpb_window_cells (native host code) == NN_degenerate._window_cells (Python restatement of core:666-687) on
    aligned rows with terminal / internal gap runs and on ragged rows, windows inside, across and past the row end"""
    import numpy as np

    from multiprime_b200 import _lib, core, synth

    rng = np.random.default_rng(5)
    for ragged in (False, True):
        codes = synth.synth_codes(300, 120, seed=77, gap_rate=0.05, iupac_rate=0.02, term_gap=0.5)
        lens = rng.integers(0, 121, len(codes)).astype(np.int32) if ragged else np.full(len(codes), 120, np.int32)
        for s, n in enumerate(lens):
            codes[s, n:] = 0
        app = object.__new__(core.NN_degenerate)          # only the host copy of the alignment is needed
        app._codes, app._packed4, app.lens, app._row_cache = None, core.pack4(codes), lens, {}
        for k in (5, 18, 27):
            app.primer_length = k
            seq = rng.integers(0, len(codes), 4000)
            pos = rng.integers(0, 125, 4000)
            cells, got = _lib.window_cells(app._packed4, lens, 120, k, seq, pos)
            for i, (s, p) in enumerate(zip(seq.tolist(), pos.tolist())):
                want = app._window_cells(s, p)
                assert got[i] == len(want), (ragged, k, s, p)
                assert bytes(cells[i, :got[i]]) == want, (ragged, k, s, p)
                assert not cells[i, got[i]:].any()


def test_primer_props_host_part(tmp_path):
    """what mpb_primer_props computes on the HOST (degeneracy, GC content, di-nucleotide / hairpin flags, mean and
    rounding of the per-expansion Tm sums the device returns): host build of mpb_walk.cu with a stub for the device call"""
    import ctypes as C
    import os
    import subprocess

    import numpy as np

    from multiprime_b200 import core
    from multiprime_b200.iupac import expand_keys

    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.dirname(here)
    so = str(tmp_path / "libprops_host.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-x", "c++", "-std=c++17", "-I", os.path.join(root, "include"),
                           "-I", os.path.join(root, "multiprime_b200", "csrc"), "-I", "/usr/local/cuda/include", "-o", so,
                           os.path.join(here, "native", "props_host.cpp")])
    lib = C.CDLL(so)
    rng = np.random.default_rng(9)
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    for k in (3, 18, 27):
        n = 40
        sets = np.zeros((n, 32), np.uint8)
        for i in range(n):
            row = [1 << int(b) for b in rng.integers(0, 4, k)]
            for j in rng.choice(k, int(rng.integers(0, min(k, 7))), replace=False):      # degenerate positions
                row[j] = int(rng.integers(1, 16))
            sets[i, :k] = row
        deg, ndeg, flags = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.int32)
        tm, gc = np.zeros(n), np.zeros(n)
        assert lib.props_host(P(sets), k, n, P(deg), P(ndeg), P(tm), P(gc), P(flags)) == 0
        for i in range(n):
            row = sets[i, :k].tolist()
            assert deg[i] == len(expand_keys(row))
            assert ndeg[i] == sum(1 for c in row if bin(c).count("1") > 1)
            assert tm[i] == round(50.0 + 0.01 * (i % 7), 2)          # the mean of identical rounded values
            if not flags[i] & 128:              # (128: a rounding tie the caller replays exactly)
                assert gc[i] == core.gc_content(row)
                assert bool(flags[i] & 1) == (not 0.4 <= gc[i] <= 0.6)
            assert bool(flags[i] & 2) == core.has_repeat(row)
            assert bool(flags[i] & 4) == core.has_hairpin(row, 4)


def test_expand_array_and_exact_mean_match_their_definitions():
    """the numpy forms used on the hot path against the plain ones: itertools.product order, statistics.mean exactness"""
    import random
    from statistics import mean
    import numpy as np
    from multiprime_b200.core import exact_mean
    from multiprime_b200.iupac import expand_array, expand_keys
    rnd = random.Random(11)
    for _ in range(100):
        sets = [rnd.choice([1, 2, 4, 8, 1, 2, 4, 8, 5, 10, 3, 12, 6, 9, 11, 14, 7, 13, 15]) for _ in range(rnd.randint(3, 10))]
        want = np.asarray(expand_keys(sets), np.uint8).reshape(-1, len(sets))
        got = expand_array(sets)
        assert got.shape == want.shape and (got == want).all(), sets
    for _ in range(200):
        vals = [round(rnd.uniform(1.5, 99.0), 2) for _ in range(rnd.randint(1, 300))]
        assert exact_mean(vals) == mean(vals)
        vals = [round(rnd.uniform(0.01, 0.99), 2) for _ in range(rnd.randint(1, 60))]       # below 1: the 2^60 path
        assert exact_mean(vals) == mean(vals)
        vals = [round(rnd.uniform(-50.0, 2000.0), 2) for _ in range(rnd.randint(1, 60))]    # mixed: falls back as needed
        assert exact_mean(vals) == mean(vals)


def test_peer_group_is_all_or_nothing(monkeypatch):
    """_lib.Peer.of: when one rank cannot create its buffer, every rank ends up without a peer group (the walk then
    all-reduces through the communicator) — decided with the communicator's own collectives, no device involved"""
    import threading
    import numpy as np
    from multiprime_b200 import _lib
    from tests.loopback_comm import _Shared, ThreadComm

    class FakePeer(_lib.Peer):
        closed = []

        def __init__(self, ctx, rank, world, cap_elems=1 << 18):
            if rank == 1:
                raise _lib.MpbError(-3, "no memory for the peer buffer")
            self.ctx, self.rank, self.world, self.cap, self.h = ctx, rank, world, cap_elems, object()

        def handle(self):
            return np.full(_lib.PEER_HANDLE_BYTES // 8, self.rank + 1, np.int64)

        def connect(self, handles):
            raise AssertionError("nobody connects when a rank has no buffer")

        def close(self):
            FakePeer.closed.append(self.rank)
            self.h = None

    world, shared, out = 3, _Shared(3), [None] * 3

    class Ctx:
        pass

    def body(rank):
        try:
            comm = ThreadComm.__new__(ThreadComm)
            comm.sh, comm.rank, comm.world, comm.device = shared, rank, world, None
            comm.peer_key = ("test", 1)
            out[rank] = FakePeer.of(Ctx(), comm)
        except BaseException as exc:          # a dead shard would leave the others at a barrier
            out[rank] = exc
            shared.barrier.abort()

    threads = [threading.Thread(target=body, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert out == [None, None, None]
    assert sorted(FakePeer.closed) == [0, 2]
