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
