"""Pin oracle/mp_oracle.py to the live reference: every golden record (tests/golden/core_*.json, made by
tests/golden/make_golden.py from multiPrime-core_V20.py) must be reproduced exactly."""
import json
import os

import pytest

from oracle import mp_oracle as o
from tests.helpers import GOLDEN, case_alignment, digest, load_case, oracle_params

CASES = ["synth300", "synth_iupac", "c2_k18", "c2_k20", "c2_k22", "c3_tmsa", "c1_testfa",
         # every window of the region: 1000_fasta.msa k = 18..22, the whole Cluster_0_20727.tmsa, 10^4 synthetic rows
         "c2f_k18", "c2f_k19", "c2f_k20", "c2f_k21", "c2f_k22", "c3f_tmsa", "c4_10k"]


@pytest.mark.parametrize("name", CASES)
def test_core_windows(name):
    case = load_case(name)
    ids, seqs = case_alignment(case, name)
    prm = oracle_params(case["params"])
    start, stop = o.region(seqs, prm.fraction)
    assert (start, stop) == (case["start"], case["stop"])
    thr = prm.entropy_threshold(stop - start)
    for rec in case["records"]:
        trace = []
        got = o.design_window(ids, seqs, rec["pos"], prm, thr, trace)
        assert trace == rec["trace"], (name, rec["pos"])
        if rec["row"] is None:
            assert got is None, (name, rec["pos"])
            continue
        assert got is not None, (name, rec["pos"])
        assert got["row"] == rec["row"], (name, rec["pos"])
        assert digest(got["non_cov"][0]) == rec["f_non"]
        assert digest(got["non_cov"][1]) == rec["r_non"]
        assert digest(got["gap_ids"]) == rec["gap_ids"]


def test_kat():
    with open(os.path.join(GOLDEN, "kat.json")) as fh:
        kat = json.load(fh)
    for s, v in kat["tm"].items():
        assert o.tm(s) == v
    for s, v in kat["dh_ds"].items():
        assert list(o.delta_h_s(s)) == v
    for s, v in kat["dg"].items():
        assert o.delta_g(s) == v
    for a, v in kat["loss"]:
        assert o.penalty_points(*a) == v
    for key, (f, r) in kat["get_y"].items():
        coord, k = key.split("|")
        gf, gr = o.strict_positions(coord, int(k))
        assert (sorted(gf), sorted(gr)) == (f, r)
    for s, (d, n) in kat["deg"].items():
        assert (o.degeneracy(s), o.n_degenerate(s)) == (d, n)
    for p, v in kat["filters"].items():
        assert o.information(p, 0.2, 0.7, 4) == v["info"]
        assert o.self_dimer(p) == v["self_dimer"]
        assert o.has_hairpin(p, 4) == v["hairpin"]
        assert o.has_repeat(p) == v["repeat"]
        assert o.gc_content(p) == v["gc"]


def test_survey_kats():
    """the known answers listed in SURVEY.md 8(c)"""
    assert o.tm("ATGAAGACCATCATTGCC") == 51.19
    assert o.delta_h_s("GGTACGGCCTCAGACATC") == (-141000.0, -379.7999999999999)
    assert o.tm("GC" * 9) == 76.6
    assert o.salt_correction() == 0.00010318549324165211
    assert o.delta_g("GCAACTGTTACC") == -9.24
    assert o.penalty_points(12, 6, 0, 0) == 7.418539921951662
    assert o.degeneracy("GGTAYGGYYTCAGRCATC") == 16
    assert o.strict_positions("1,2,-1", 18) == ({1, 2, 18}, {16, 17, 2})
