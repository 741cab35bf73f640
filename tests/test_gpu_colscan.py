"""GPU suite: the column scan (mpb_cscan) against the row kernel (mpb_scan) on identical candidates, the full-region
goldens of the live reference, and the benchmarked regime (>= 2^18 sequences: N/2-slot tables, prefilter pruning)
against the oracle."""
import numpy as np
import pytest

from tests.parity import check_case

pytestmark = pytest.mark.gpu


def _random_candidates(rng, codes, pos, k, per_window=4):
    """per window: a k-mer of a random row, relaxed at a few positions; every second one with a trial (the primer
    with one more base at one position, counted like the reference's coverage_renew look-up)"""
    from multiprime_b200.iupac import allow_masks
    n = codes.shape[0]
    wins, allows, trials = [], [], []
    for wi, p in enumerate(pos):
        for c in range(per_window):
            row = codes[rng.integers(0, n), p:p + k]
            sets = [int(x) if x else 1 << int(rng.integers(0, 4)) for x in row]
            sets = sets + [1] * (k - len(sets))
            for j in rng.integers(0, k, int(rng.integers(0, 5))):
                sets[j] |= 1 << int(rng.integers(0, 4))
            if c == per_window - 1:
                sets[int(rng.integers(0, k))] = 0                      # a position nothing matches
            trial = -1
            if c % 2 == 1:
                tp = int(rng.integers(0, k))
                free = [b for b in range(4) if not (sets[tp] >> b) & 1]
                if free and sets[tp]:
                    tb = free[int(rng.integers(0, len(free)))]
                    sets[tp] |= 1 << tb
                    trial = tp | (tb << 8)
            wins.append(wi)
            allows.append(allow_masks(sets))
            trials.append(trial)
    return np.array(wins, np.int32), np.array(allows, np.uint32), np.array(trials, np.int32)


@pytest.mark.parametrize("n,L,k,v,kw", [
    (3000, 200, 18, 3, dict(seed=3)),                                            # the bench's flags
    (1500, 160, 22, 2, dict(seed=4, gap_rate=0.02, iupac_rate=0.004)),          # many patched / IUPAC rows
    (700, 150, 9, 0, dict(seed=5, gap_rate=0.05, iupac_rate=0.01, term_gap=0.6)),
    (2100, 130, 27, 5, dict(seed=6, gap_rate=0.01)),                            # 5-bit counter
    (40000, 120, 18, 1, dict(seed=7)),                                           # several word tiles per block
])
def test_column_scan_equals_row_scan(n, L, k, v, kw):
    from multiprime_b200 import _lib, core, synth
    rng = np.random.default_rng(n + k)
    codes = synth.synth_codes(n, L, **kw)
    ragged = kw.get("term_gap", 0) > 0.5
    lens = None
    if ragged:                                                                   # unaligned input: rows end early
        lens = rng.integers(L // 2, L + 1, n).astype(np.int32)
        for s, m in enumerate(lens):
            codes[s, m:] = 0
    ctx = _lib.Context(0)
    msa = _lib.Msa(ctx, core.pack4(codes), n, L, lens=lens)
    pos = sorted(set(rng.integers(0, L - k - (20 if ragged else 0), 24).tolist()))
    fmask, rmask = core.strict_masks("1,2,-1", k)
    wins, allows, trials = _random_candidates(rng, codes, pos, k)
    cands = _lib.make_cands(wins, allows, trials)
    slots = np.arange(len(cands), dtype=np.int32)
    slots[::3] = -1
    slots[slots >= 0] = np.arange((slots >= 0).sum())
    with msa.hist(k, v, pos) as h:
        got, gbits = h.cscan(fmask, rmask, cands, bits_slot=slots)
        got2, _ = h.cscan(fmask, rmask, cands)
    cand_pos = np.array(pos, np.int32)[wins]
    want, wbits = msa.scan(k, v, fmask, rmask, cand_pos, allows, bits_slot=slots)
    assert (got[:, :3] == want).all()
    assert (got2 == got).all()
    assert (gbits == wbits).all()
    # the trial count is the perfect count of "the primer with that position := the one base"
    sel = np.nonzero(trials >= 0)[0]
    t_allow = allows[sel].copy()
    for j, ci in enumerate(sel):
        tp, tb = int(trials[ci]) & 255, int(trials[ci]) >> 8
        t_allow[j] &= ~np.uint32(1 << tp)
        t_allow[j, tb] |= np.uint32(1 << tp)
    tw, _ = msa.scan(k, v, fmask, rmask, cand_pos[sel], t_allow)
    assert (got[sel, 3] == tw[:, 0]).all()
    assert got[:, 0].sum() > 0
    msa.close()
    ctx.close()


@pytest.mark.parametrize("name", ["c2f_k18", "c2f_k19", "c2f_k20", "c2f_k21", "c2f_k22", "c3f_tmsa"])
def test_golden_full_region(name):
    """every window of the region against the live reference's records: 1000_fasta.msa for k = 18..22 (BASELINE
    configs[1]) and the whole Cluster_0_20727.tmsa with the YAML flags (configs[2])"""
    stats = check_case(name)
    assert stats["scan_calls"] > 0


def test_golden_c4_first_10k_rows():
    """the first 10^4 rows of the north-star synthetic alignment with the benchmark's flags (-l 18 -n 8 -d 256 -v 3):
    rows, call traces and side files of all 582 windows equal the live reference's"""
    stats = check_case("c4_10k")
    assert stats["accepted"] == 174


def test_benchmarked_regime_vs_oracle():
    """2^18 synthetic sequences with the benchmark's flags: tables of N/2 slots, the whole-window prefilter pruning
    most windows.  All 582 windows go through design(); 24 windows spread over accepted / rejected / edge-of-gate
    cases are compared row for row, trace for trace and side file for side file with the oracle."""
    from multiprime_b200 import core, synth
    from oracle import mp_oracle as o
    n, L, k = 1 << 18, 600, 18
    codes = synth.synth_codes_parallel(n, L)
    ids = synth.seq_ids(n)
    app = core.NN_degenerate(seq_file=None, primer_length=k, coverage=0.8, number_of_dege_bases=8,
                             score_of_dege_bases=256, product_len=100, position="1,2,-1", variation=3,
                             raw_entropy_threshold=3.6, distance=4, GC="0.2,0.7", nproc=1, outfile="",
                             alignment=(ids, codes, np.full(n, L, np.int32)), device=0, sidecars=False)
    assert app._table_log2cap(k, n) == 17                       # the N/2 regime of the 10^6 benchmark
    positions = list(range(app.start_position, app.stop_position - k))
    recs = {r["row"][0]: r for r in app.design(positions)}
    assert app.stats["prefiltered"] > 200 and len(recs) > 100
    accepted = sorted(recs)
    rejected = [p for p in positions if p not in recs]
    # edge of the gate: rejected windows next to accepted ones, accepted windows with the highest entropy
    edge_rej = [p for p in rejected if (p - 1 in recs) or (p + 1 in recs)]
    by_ent = sorted(accepted, key=lambda p: -recs[p]["row"][2])
    pick = accepted[::max(1, len(accepted) // 8)][:8] + by_ent[:6] + edge_rej[:6] + rejected[::max(1, len(rejected) // 4)][:4]
    pick = sorted(set(pick))
    assert len(pick) >= 16
    seqs = synth.codes_to_strings(codes)
    prm = o.Params(k=k, dnum=8, degeneracy=256, variation=3, entropy=3.6, gc="0.2,0.7", size=100, fraction=0.8,
                   coordinate="1,2,-1", away=4)
    start, stop = o.region(seqs, 0.8)
    assert (app.start_position, app.stop_position) == (start, stop)
    # the side files need the per-sequence keys: a second pass over the picked windows with sidecars on
    app.sidecars = True
    full = {r["row"][0]: r for r in app.design(pick)}
    for p in pick:
        trace = []
        want = o.design_window(ids, seqs, p, prm, prm.entropy_threshold(stop - start), trace)
        if want is None:
            assert p not in recs and p not in full, p
            continue
        assert p in recs, p
        assert recs[p]["row"] == want["row"], (p, recs[p]["row"], want["row"])
        assert recs[p]["trace"] == trace, p
        assert full[p]["row"] == want["row"]
        assert full[p]["non_cov"] == want["non_cov"] and full[p]["gap_ids"] == want["gap_ids"], p
    app.close()
