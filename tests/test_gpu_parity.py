"""GPU suite: libmpb200 (through the C ABI) against the reference's golden records and against the oracle."""
import numpy as np
import pytest

from tests.parity import check_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["synth_iupac", "synth300", "c2_k18", "c2_k22", "c3_tmsa", "c1_testfa"])
def test_golden_case(name):
    stats = check_case(name)
    assert stats["scan_calls"] > 0


def test_smoke_entry():
    import __graft_entry__ as g
    g.smoke()


def test_tm_kernel_bit_exact():
    """Tm / dH / dS of random 18..22-mers equal the oracle's floats bit for bit (same operation order, no FMA)"""
    from multiprime_b200 import _lib, core
    from oracle import mp_oracle as o
    rng = np.random.default_rng(3)
    ctx = _lib.Context(0)
    for k in (16, 18, 21, 22):
        seqs = rng.integers(0, 4, (500, k)).astype(np.uint8)
        seqs[0, :k - k % 2] = np.array([2, 1] * (k // 2))   # self-complementary when k is even
        tm, dh, ds = ctx.tm(seqs, core.TM_CONSTS, want_hs=True)
        for i, row in enumerate(seqs):
            s = "".join("ACGT"[b] for b in row)
            assert tm[i] == o.tm_unrounded(s), (s, tm[i], o.tm_unrounded(s))
            assert (dh[i], ds[i]) == o.delta_h_s(s)
    ctx.close()
