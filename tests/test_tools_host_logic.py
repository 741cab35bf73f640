"""CPU suite: the host logic of the finDimer / get_Maxprimerset / get_multiPrime drop-ins on the fake dimer engine
(tests/fake_device.py), against the reference's goldens.  The GPU suite runs the same goldens on libmpb200."""
import json
import os

import pytest

from tests import fake_device
from tests.helpers import GOLDEN


def _load(name):
    with open(os.path.join(GOLDEN, name)) as fh:
        return json.load(fh)


def test_findimer(tmp_path):
    from multiprime_b200 import findimer
    g = _load("dimer_findimer.json")
    fa = tmp_path / "p.fa"
    fa.write_text("".join(">P%03d\n%s\n" % (i, p) for i, p in enumerate(g["primers"])))
    app = findimer.Dimer(primer_file=str(fa), outfile=str(tmp_path / "o.txt"), threshold=g["threshold"], nproc=1,
                         _backend=fake_device)
    rows = app.run()
    assert [list(r) for r in rows] == g["rows"]


@pytest.mark.parametrize("mode", ["T", "F_easy"])
def test_maxprimerset(tmp_path, mode, capsys):
    from multiprime_b200 import maxprimerset
    g = _load("cover_maxprimerset.json")
    lines = g["input_easy"] if mode == "F_easy" else g["input"]
    inp = tmp_path / "candidate_primers_sets.txt"
    inp.write_text("\n".join(lines) + "\n")
    out = tmp_path / "final_maxprimers_set.xls"
    maxprimerset.main(["-i", str(inp), "-o", str(out), "-s", "5", "-m", mode[0]], _backend=fake_device)
    want = g[mode]
    assert out.read_text() == want["out"]
    assert (tmp_path / "sort.candidate_primers_sets.txt").read_text() == want["sort"]
    if want["next"] is not None:
        assert (tmp_path / "final_maxprimers_set.next.xls").read_text() == want["next"]


@pytest.mark.parametrize("tag", ["a", "b"])
def test_get_multiprime(tmp_path, tag, capsys):
    from multiprime_b200 import pairing, synth
    g = _load("pairs_get_multiprime.json")
    n, L, seed, gr, ir = g["synth"]
    fa = tmp_path / "in.fa"
    synth.write_fasta(str(fa), synth.synth_codes(n, L, seed=seed, gap_rate=gr, iupac_rate=ir))
    core_out = tmp_path / "c.out"
    core_out.write_text(g["core_tsv"])
    json.dump(g["core_non_cov"], open(str(core_out) + ".non_coverage_seq_id_json", "w"))
    json.dump(g["core_gap"], open(str(core_out) + ".gap_seq_id_json", "w"))
    out = tmp_path / ("Cluster_%s.candidate.primers.txt" % tag)
    want = g[tag]
    pairing.main(["-i", str(core_out), "-r", str(fa), "-o", str(out)] + want["args"], _backend=fake_device)
    stem = str(out).strip(".txt")
    assert out.read_text().replace(str(tmp_path), "<TMP>") == want["txt"]
    assert open(stem + ".xls").read() == want["xls"]
    assert open(stem + ".fa").read() == want["fa"]
