"""Pin oracle/dimer_oracle.py to the live reference: finDimer_V4 rows, get_Maxprimerset_V1.3 outputs (three modes) and
get_multiPrime outputs (two cases), as recorded by tests/golden/make_golden.py."""
import json
import os

import pytest

from oracle import dimer_oracle as d
from tests.helpers import GOLDEN


def _load(name):
    with open(os.path.join(GOLDEN, name)) as fh:
        return json.load(fh)


def test_findimer_rows():
    g = _load("dimer_findimer.json")
    primers = {}
    for i, p in enumerate(g["primers"]):
        primers[p] = ">P%03d" % i                 # keyed by sequence: a duplicate keeps the last header (fd:138-146)
    rows = d.find_dimers(primers, g["threshold"])
    assert [list(r) for r in rows] == g["rows"]


@pytest.mark.parametrize("mode", ["T", "F", "F_easy"])
def test_maxprimerset(mode):
    g = _load("cover_maxprimerset.json")
    lines = g["input_easy"] if mode == "F_easy" else g["input"]
    primers = d.sort_clusters(lines)
    want = g[mode]
    assert "".join("\t".join(r) + "\n" for r in primers) == want["sort"]
    ref_msgs = [ln for ln in want["stdout"].splitlines() if not ln.startswith("INFO")]
    if mode == "T":
        table, nxt, msgs = d.greedy_maximal(primers)
        assert table == want["out"] and nxt == want["next"] and msgs == ref_msgs
    else:
        table, msgs, rc = d.greedy_maximum(primers)
        assert rc == want["rc"] and msgs == ref_msgs
        if want["out"] is not None:
            assert table == want["out"]


@pytest.mark.parametrize("tag", ["a", "b"])
def test_get_multiprime(tag):
    g = _load("pairs_get_multiprime.json")
    want = g[tag]
    args = dict(zip(want["args"][0::2], want["args"][1::2]))
    kw = dict(size=args["-s"], fraction=float(args["-f"]), diff_tm=int(args["-t"]))
    if "-e" in args:
        kw["term"] = int(args["-e"])
    if "-a" in args:
        kw["adaptor"] = args["-a"]
    out_path = "<TMP>/Cluster_%s.candidate.primers.txt" % tag
    txt, xls, fa, msgs = d.pair_candidates(g["core_tsv"], g["core_gap"], g["core_non_cov"], g["synth"][0], out_path, **kw)
    assert txt == want["txt"] and xls == want["xls"] and fa == want["fa"]
    assert msgs == [ln for ln in want["stdout"].splitlines() if not ln.startswith("INFO")]
