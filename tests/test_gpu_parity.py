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


def test_self_dimer_engine_vs_oracle():
    """core:487-503 on random degenerate primers (up to 192 expansions): device flags == oracle"""
    import json
    import os
    import random
    from multiprime_b200 import _lib, dimer
    from multiprime_b200.iupac import sets_of
    from oracle import mp_oracle as o
    from tests.helpers import GOLDEN
    random.seed(11)
    kat = json.load(open(os.path.join(GOLDEN, "kat.json")))["filters"]
    primers = list(kat)
    codes = "ACGTRYMKSWHBVDN"
    while len(primers) < 260:
        k = random.choice([16, 18, 20, 22, 25])
        p = "".join(random.choice("ACGT" * 7 + codes) for _ in range(k))
        if o.degeneracy(p) <= 192:
            primers.append(p)
    ctx = _lib.Context(0)
    flags = dimer.dimer_flags(ctx, [sets_of(p) for p in primers])
    for p, f in zip(primers, flags):
        assert bool(f) == o.self_dimer(p), p
    ctx.close()
