"""Differential test on CPU: random small alignments and random parameters through the product's host logic (native
walk + gates, fake device underneath) against the oracle, window by window.  The oracle is pinned to the reference by
the golden vectors; this widens the parameter space (k, variation, coordinates, degeneracy caps, gap / IUPAC rates)."""
import numpy as np
import pytest

from multiprime_b200 import core, synth
from oracle import mp_oracle as o
from tests import fake_device

CONFIGS = [
    # (seed, n_seq, n_col, k, dnum, degeneracy, variation, coordinate, gap_rate, iupac_rate, fraction)
    (101, 40, 90, 12, 3, 8, 0, "1,-1", 0.004, 0.0, 0.8),
    (102, 55, 100, 16, 5, 32, 2, "2,3,-2", 0.006, 0.004, 0.7),
    (103, 30, 110, 19, 6, 48, 3, "-1", 0.002, 0.006, 0.8),
    (104, 64, 90, 21, 4, 10, 1, "1,2,-1", 0.01, 0.002, 0.6),
    (105, 70, 130, 24, 8, 256, 2, "3", 0.003, 0.003, 0.8),
    (106, 48, 80, 27, 2, 4, 1, "1,2,-1", 0.0, 0.0, 0.8),
    (107, 33, 95, 9, 4, 16, 1, "1,-1,-2", 0.008, 0.01, 0.75),
    (108, 16, 120, 18, 6, 64, 2, "1,2,-1", 0.02, 0.0, 0.5),
]


@pytest.mark.parametrize("cfg", CONFIGS, ids=[str(c[0]) for c in CONFIGS])
def test_random_alignment_matches_oracle(cfg):
    _run(cfg, fake_device)


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", CONFIGS, ids=[str(c[0]) for c in CONFIGS])
def test_random_alignment_matches_oracle_gpu(cfg):
    """the same comparison with libmpb200 underneath (k from 9 to 27, variation 0..3, assorted strict positions)"""
    _run(cfg, None)


def _run(cfg, backend):
    seed, n, L, k, dnum, deg, v, coord, gr, ir, frac = cfg
    codes = synth.synth_codes(n, L, seed=seed, gap_rate=gr, iupac_rate=ir, term_gap=0.15)
    ids, seqs = synth.seq_ids(n), synth.codes_to_strings(codes)
    prm = o.Params(k=k, dnum=dnum, degeneracy=deg, variation=v, entropy=3.6, gc="0.2,0.7", size=10, fraction=frac,
                   coordinate=coord, away=4)
    start, stop = o.region(seqs, frac)
    if stop - start < 10 + k:
        pytest.skip("region too short")
    app = core.NN_degenerate(seq_file=None, primer_length=k, coverage=frac, number_of_dege_bases=dnum,
                             score_of_dege_bases=deg, product_len=10, position=coord, variation=v,
                             raw_entropy_threshold=3.6, distance=4, GC="0.2,0.7", nproc=1, outfile="",
                             alignment=(ids, codes, np.full(n, L, np.int32)), _backend=backend)
    assert (app.start_position, app.stop_position) == (start, stop)
    positions = list(range(start, stop - k))
    got = {r["row"][0]: r for r in app.design(positions)}
    thr = prm.entropy_threshold(stop - start)
    n_acc = 0
    for p in positions:
        trace = []
        want = o.design_window(ids, seqs, p, prm, thr, trace)
        if want is None:
            assert p not in got, (cfg, p, got[p]["row"] if p in got else None)
            continue
        n_acc += 1
        assert p in got, (cfg, p, want["row"])
        assert got[p]["row"] == want["row"], (cfg, p, got[p]["row"], want["row"])
        assert got[p]["trace"] == trace, (cfg, p)
        assert got[p]["non_cov"] == want["non_cov"] and got[p]["gap_ids"] == want["gap_ids"], (cfg, p)
    app.close()
    print("accepted windows:", n_acc, "of", len(positions))
