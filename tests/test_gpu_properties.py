"""GPU suite: size-independent properties at sizes the oracle cannot reach, and edge cases of the C ABI."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

K, V = 18, 3


def _msa(ctx, codes):
    from multiprime_b200 import _lib, core
    return _lib.Msa(ctx, core.pack4(codes), codes.shape[0], codes.shape[1])


def test_scan_and_tables_are_additive_over_sequences():
    """every per-window quantity the walk consumes is a sum over sequences: scanning rows A, rows B and rows A+B must
    give counts(A) + counts(B) == counts(A+B); the same for gap counts, base / dinucleotide tensors and prefilter item
    counts (200 000 synthetic sequences, incl. gap-edge and IUPAC rows)"""
    from multiprime_b200 import _lib, synth
    from multiprime_b200.iupac import allow_masks, sets_of
    n = 200_000
    codes = synth.synth_codes(n, 600, seed=77)
    cut = 83_111
    ctx = _lib.Context(0)
    parts = [_msa(ctx, codes[:cut]), _msa(ctx, codes[cut:]), _msa(ctx, codes)]
    pos = np.array([0, 3, 17, 40, 58, 125, 300, 577], np.int32)
    fmask, rmask = 0b110, (1 << 17) | (1 << 16) | 0b100
    rng = np.random.default_rng(5)
    cand_pos, cand_allow = [], []
    for p in pos:
        row = codes[rng.integers(0, n), p:p + K]
        sets = [int(c) if c else 1 for c in row]
        for extra in range(3):                                  # the k-mer itself and two degenerate relaxations
            s2 = list(sets)
            for j in rng.integers(0, K, extra * 3):
                s2[j] |= 1 << int(rng.integers(0, 4))
            cand_pos.append(int(p))
            cand_allow.append(allow_masks(s2))
    res = [m.scan(K, V, fmask, rmask, cand_pos, cand_allow)[0] for m in parts]
    assert (res[0] + res[1] == res[2]).all()
    assert res[2][:, 0].sum() > 0 and (res[2] >= 0).all()
    stats, tens, pre = [], [], []
    for m in parts:
        with m.hist(K, V, pos) as h:
            st = h.stats()
            stats.append(st)
            tens.append(h.tensors(np.ones(len(pos), np.uint8)))
        pre.append(m.prefilter(K, V, pos))
    assert (stats[0]["gap_n"] + stats[1]["gap_n"] == stats[2]["gap_n"]).all()
    assert (tens[0][0] + tens[1][0] == tens[2][0]).all() and (tens[0][1] + tens[1][1] == tens[2][1]).all()
    assert (pre[0][0] + pre[1][0] == pre[2][0]).all()
    # items = expansion rows of cover rows + gap rows; base counts of column 0 sum to the cover expansion rows
    assert (stats[2]["ent"][:, 0] + stats[2]["ent"][:, 2] + stats[2]["n_iupac_gap"] == pre[2][0]).all()
    # the prefilter bound never exceeds the exact (approximate-sum) total entropy
    n_all = float(n)
    ent = stats[2]["ent"]
    t_bit = -((ent[:, 1] - ent[:, 0] * np.log2(n_all)) + (ent[:, 3] - ent[:, 2] * np.log2(n_all))) / n_all
    bound = (pre[2][0] * np.log2(n_all) - pre[2][1]) / n_all
    assert (bound <= t_bit + 1e-9).all()
    for m in parts:
        m.close()
    ctx.close()


def test_bits_agree_with_counts():
    """per-sequence non-cover bits of mpb_scan: rows not flagged F-non-cover and not gap rows are exactly the rows
    counted as perfect or F-mis-covered (no IUPAC cells in this input, so one expansion per row)"""
    from multiprime_b200 import _lib, synth
    from multiprime_b200.iupac import allow_masks
    n = 50_000
    codes = synth.synth_codes(n, 300, seed=3, iupac_rate=0.0)
    ctx = _lib.Context(0)
    m = _msa(ctx, codes)
    pos = [5, 130, 250]
    allow = [allow_masks([int(c) if c else 1 for c in codes[7, p:p + K]]) for p in pos]
    counts, bits = m.scan(K, V, 0b110, 1 << 17, pos, allow, bits_slot=[0, 1, 2])
    for i in range(3):
        unpack = lambda w: np.unpackbits(w.view(np.uint8), bitorder="little")[:n].astype(bool)
        non_f, non_r, gap = unpack(bits[i, 0]), unpack(bits[i, 1]), unpack(bits[i, 2])
        assert (~non_f & ~gap).sum() == counts[i, 0] + counts[i, 1]
        assert (~non_r & ~gap).sum() == counts[i, 0] + counts[i, 2]
        assert not (non_f & gap).any()
    m.close()
    ctx.close()


def test_abi_edge_cases():
    from multiprime_b200 import _lib, core, synth
    ctx = _lib.Context(0)
    codes = synth.synth_codes(64, 100, seed=1)
    m = _msa(ctx, codes)
    with pytest.raises(_lib.MpbError):          # primer longer than the 64-bit keys allow
        m.hist(28, 1, [0])
    with pytest.raises(_lib.MpbError):          # window outside the alignment
        m.hist(18, 1, [100])
    with pytest.raises(_lib.MpbError):
        m.scan(18, 1, 0, 0, [500], [[1, 1, 1, 1]])
    counts, _ = m.scan(18, 1, 0, 0, [], np.zeros((0, 4), np.uint32))          # empty candidate list
    assert counts.shape == (0, 3)
    # k = 27 (the largest supported): table keys still distinguish haplotypes
    with m.hist(27, 2, [0, 10]) as h:
        st = h.stats()
        assert (st["nuniq"][:, 0] >= 1).all()
    # a sequence with fewer than k bases is refused
    short = np.zeros((4, 40), np.uint8)
    short[:, :10] = 1
    ms = _lib.Msa(ctx, core.pack4(short), 4, 40, lens=np.array([40, 40, 40, 12], np.int32))
    with pytest.raises(_lib.MpbError):
        ms.hist(18, 1, [30])
    ms.close()
    m.close()
    ctx.close()


def test_single_sequence_alignment():
    """degenerate input: one sequence, no variation -> every window yields the sequence's own k-mer"""
    from multiprime_b200 import core
    from multiprime_b200.iupac import CHAR_CODE
    seq = "ACGTTGCAAGCTTAGGCTAACGGATCCATGGCAATTCGTAGCTAGGATCCGATTACAGGCTTAAGGCCTTAACGGTTAACCGGATATCGCGCGATATCCGG" * 2
    codes = np.array([[CHAR_CODE[c] for c in seq]], np.uint8)
    app = core.NN_degenerate(seq_file=None, primer_length=18, coverage=0.8, number_of_dege_bases=4, score_of_dege_bases=10,
                             product_len=50, position="1,2,-1", variation=1, raw_entropy_threshold=3.6, distance=4,
                             GC="0.2,0.7", nproc=1, outfile="", alignment=([">one"], codes, np.array([len(seq)], np.int32)))
    recs = app.design(range(0, 40))
    assert recs, "no window accepted"
    for r in recs:
        p = r["row"][0]
        assert r["row"][3] == seq[p:p + 18] and r["row"][6] == 1 and r["row"][1] == -0.0
    app.close()


def test_small_tables_overflow_and_are_rebuilt():
    """a table with too few slots reports MPB_EOVERFLOW (probe runs are bounded) and _lib.Hist rebuilds the batch with
    doubled tables until it fits: the statistics equal those of the default-sized tables"""
    from multiprime_b200 import _lib, synth
    codes = synth.synth_codes(3000, 200, seed=9, gap_rate=0.004, iupac_rate=0.001)
    ctx = _lib.Context(0)
    m = _msa(ctx, codes)
    pos = [0, 30, 75, 120, 150]                      # conserved and variable blocks
    with m.hist(K, V, pos) as h0:
        want = h0.stats()
        want_t = h0.tensors(np.ones(len(pos), np.uint8))
    assert want["nuniq"][:, :2].sum(axis=1).max() > 256          # some window cannot fit 2^7 slots
    with pytest.raises(_lib.MpbError) as info:                   # the C ABI itself reports the overflow ...
        h = _lib.C.c_void_p()
        wp = np.array(pos, np.int32)
        _lib.check(_lib.load().mpb_hist_build(m.h, K, V, _lib.ptr(wp), len(pos), 7, _lib.C.byref(h)))
    assert info.value.code == -4
    with m.hist(K, V, pos, 7) as h1:                             # ... and the binding grows the tables
        got = h1.stats()
        got_t = h1.tensors(np.ones(len(pos), np.uint8))
    for key in ("gap_n", "nuniq", "mm_key", "mm_cnt", "mm_first", "n_iupac_gap"):
        assert (got[key] == want[key]).all(), key
    assert np.allclose(got["ent"], want["ent"], rtol=1e-12, atol=1e-9)
    assert (got_t[0] == want_t[0]).all() and (got_t[1] == want_t[1]).all()
    m.close()
    ctx.close()
