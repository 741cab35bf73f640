"""GPU suite: the column-domain window passes (k_prefilter_bs, k_hist_col: the default) against the row-domain kernels
(MPB_WINPASS=row) on identical inputs — table contents, per-window counters, statistics, and, through the column scan on
random candidates, the row classes and patched special windows both builds leave behind.  The two prefilters use
different code functions: their item counts must agree exactly, and both must stay below the exact entropy."""
import numpy as np
import pytest

from tests.test_gpu_colscan import _random_candidates

pytestmark = pytest.mark.gpu


def _snapshot(msa, k, v, pos, codes, rng_seed, with_prefilter):
    from multiprime_b200 import _lib, core
    out = {}
    if with_prefilter:
        out["pre"] = msa.prefilter(k, v, pos)
    fmask, rmask = core.strict_masks("1,2,-1", k)
    rng = np.random.default_rng(rng_seed)
    wins, allows, trials = _random_candidates(rng, codes, pos, k, per_window=3)
    cands = _lib.make_cands(wins, allows, trials)
    slots = np.arange(len(cands), dtype=np.int32)
    with msa.hist(k, v, pos) as h:
        out["counts"] = h.counts()
        st = h.stats()
        out["stats"] = {name: st[name] for name in ("gap_n", "nuniq", "mm_key", "mm_cnt", "mm_first", "n_iupac_gap")}
        out["ent"] = st["ent"]
        n_ent = out["counts"][2]
        out["tables"] = [h.dump(w, int(n_ent[w]) + 1) for w in range(len(pos))]
        ew, es = h.exceptions()
        out["exc"] = sorted(zip(ew.tolist(), es.tolist()))
        out["scan"] = h.cscan(fmask, rmask, cands, bits_slot=slots)
    return out


@pytest.mark.parametrize("n,L,k,v,kw", [
    (5000, 200, 18, 3, dict(seed=3)),                                            # the bench's flags
    (4200, 170, 22, 2, dict(seed=4, gap_rate=0.02, iupac_rate=0.004)),          # many patched / IUPAC rows
    (900, 150, 9, 0, dict(seed=5, gap_rate=0.05, iupac_rate=0.01, term_gap=0.6)),   # ragged rows, short k
    (2100, 130, 27, 5, dict(seed=6, gap_rate=0.01)),                            # longest k
    (1300, 90, 5, 1, dict(seed=8, gap_rate=0.03, iupac_rate=0.002)),            # k below the prefilter's range
    (300, 80, 4, 4, dict(seed=9, gap_rate=0.2)),                                # variation >= k: all-gap rows are cover rows
    (70000, 120, 18, 1, dict(seed=7, gap_rate=0.004, term_gap=0.3)),            # many blocks, terminal gap runs
])
def test_column_passes_equal_row_passes(n, L, k, v, kw, monkeypatch):
    from multiprime_b200 import _lib, core, synth
    rng = np.random.default_rng(n + k)
    codes = synth.synth_codes(n, L, **kw)
    ragged = kw.get("term_gap", 0) > 0.5
    lens = None
    if ragged:
        lens = rng.integers(L // 2, L + 1, n).astype(np.int32)
        for s, m in enumerate(lens):
            codes[s, m:] = 0
    ctx = _lib.Context(0)
    msa = _lib.Msa(ctx, core.pack4(codes), n, L, lens=lens)
    hi = L - k - (L // 2 if ragged else 0)
    pos = sorted(set(rng.integers(0, hi, 40).tolist()) | set(range(10, min(hi, 10 + 40))))   # scattered + a dense run
    snaps = {}
    for mode in ("row", "col"):
        monkeypatch.setenv("MPB_WINPASS", mode)
        snaps[mode] = _snapshot(msa, k, v, pos, codes, 11, with_prefilter=k >= 8)
    a, b = snaps["row"], snaps["col"]
    if k >= 8:
        assert (a["pre"][0] == b["pre"][0]).all()                               # items counted: every expansion, every gap row
        ent, n_ig = a["ent"], a["stats"]["n_iupac_gap"]
        exact = -((ent[:, 1] - ent[:, 0] * np.log2(n)) + (ent[:, 3] - ent[:, 2] * np.log2(n))) / n
        ok = n_ig == 0                                                           # (those rows are not table entries)
        for snap in (a, b):
            bound = (snap["pre"][0] * np.log2(n) - snap["pre"][1]) / n
            assert (bound[ok] <= exact[ok] + 1e-9).all()
            assert (bound[ok] >= 0.5 * exact[ok] - 1e-9).all()                   # and it is not a trivial bound
    for x, y in zip(a["counts"], b["counts"]):
        assert (x == y).all()
    for name in a["stats"]:
        assert (a["stats"][name] == b["stats"][name]).all(), name
    assert np.allclose(a["ent"], b["ent"], rtol=1e-12, atol=1e-9)
    for w, (ta, tb) in enumerate(zip(a["tables"], b["tables"])):
        for x, y in zip(ta, tb):                                                # keys, counts, first-seen order
            assert (x == y).all(), w
    assert a["exc"] == b["exc"]
    assert (a["scan"][0] == b["scan"][0]).all()
    assert (a["scan"][1] == b["scan"][1]).all()
    assert a["counts"][2].sum() > 0
    msa.close()
    ctx.close()


def test_bitsliced_prefilter_is_tight_on_plain_rows():
    """gap-free, IUPAC-free alignment: every row is one item, the exact window entropy comes from numpy; the 8192-bin
    code histogram must bound it from below and, where the window has few distinct k-mers, reach it"""
    from multiprime_b200 import _lib, core, synth
    n, L, k = 40000, 150, 18
    codes = synth.synth_codes(n, L, seed=12, gap_rate=0.0, iupac_rate=0.0, term_gap=0.0)
    ctx = _lib.Context(0)
    msa = _lib.Msa(ctx, core.pack4(codes), n, L)
    pos = list(range(0, L - k, 3))
    s0, s1 = msa.prefilter(k, 2, pos)
    assert (s0 == n).all()
    bound = (s0 * np.log2(n) - s1) / n
    base = np.log2(codes).astype(np.uint64)                                      # one-hot 1,2,4,8 -> 0..3
    for wi, p in enumerate(pos):
        key = np.zeros(n, np.uint64)
        for j in range(k):
            key = key * np.uint64(4) + base[:, p + j]
        _, cnt = np.unique(key, return_counts=True)
        exact = float(-(cnt / n * np.log2(cnt / n)).sum())
        assert bound[wi] <= exact + 1e-9, (p, bound[wi], exact)
        if len(cnt) < 300:
            assert bound[wi] >= exact - 0.05, (p, bound[wi], exact, len(cnt))
        else:
            assert bound[wi] >= min(exact, 12.0) - 1.5, (p, bound[wi], exact, len(cnt))
    msa.close()
    ctx.close()


@pytest.mark.parametrize("name", ["c1_testfa", "synth_iupac", "c2_k18", "c3_tmsa"])
def test_goldens_with_row_passes(name, monkeypatch):
    """the reference's records through the row-domain kernels too (the default path runs the column-domain ones)"""
    from tests.parity import check_case
    monkeypatch.setenv("MPB_WINPASS", "row")
    assert check_case(name)["scan_calls"] > 0
