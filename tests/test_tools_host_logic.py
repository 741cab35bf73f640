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


@pytest.mark.parametrize("tag", ["a", "b"])
def test_get_multiprime_from_bit_vectors(tmp_path, tag, capsys):
    """the core step writes per-sequence bit vectors instead of the JSON side files and the pairing step
    takes its pair coverage from them — the reference's three output files must come out byte for byte"""
    from multiprime_b200 import cli_core, core, pairing, synth
    g = _load("pairs_get_multiprime.json")
    n, L, seed, gr, ir = g["synth"]
    fa = tmp_path / "in.fa"
    synth.write_fasta(str(fa), synth.synth_codes(n, L, seed=seed, gap_rate=gr, iupac_rate=ir))
    core_out = tmp_path / "c.out"
    ids, codes, lens = core.parse_msa(str(fa))
    app = core.NN_degenerate(seq_file=None, primer_length=18, coverage=0.8, number_of_dege_bases=4, score_of_dege_bases=10,
                             raw_entropy_threshold=3.6, product_len=100, position="1,2,-1", variation=1, distance=4,
                             GC="0.2,0.7", nproc=1, outfile=str(core_out), alignment=(ids, codes, lens),
                             sidecar_format="bits", want_trace=False, _backend=fake_device)
    app.run()
    assert core_out.read_text() == g["core_tsv"]
    assert os.path.exists(str(core_out) + ".coverage_bits.npz")
    assert not os.path.exists(str(core_out) + ".gap_seq_id_json")
    out = tmp_path / ("Cluster_%s.candidate.primers.txt" % tag)
    want = g[tag]
    pairing.main(["-i", str(core_out), "-r", str(fa), "-o", str(out)] + want["args"], _backend=fake_device)
    stem = str(out).strip(".txt")
    assert out.read_text().replace(str(tmp_path), "<TMP>") == want["txt"]
    assert open(stem + ".xls").read() == want["xls"]
    assert open(stem + ".fa").read() == want["fa"]


@pytest.mark.parametrize("world", [2, 3])
def test_findimer_row_bands_over_ranks(tmp_path, world):
    """the dimer grid's row bands dealt over ranks (threads + loop-back communicator + the stand-in
    engine) and the hit lists gathered give the reference's rows"""
    from multiprime_b200 import findimer
    from tests.loopback_comm import run_shards
    g = _load("dimer_findimer.json")
    fa = tmp_path / "p.fa"
    fa.write_text("".join(">P%03d\n%s\n" % (i, p) for i, p in enumerate(g["primers"])))

    def shard(rank, comm):
        app = findimer.Dimer(primer_file=str(fa), outfile=str(tmp_path / ("o%d.txt" % rank)), threshold=g["threshold"],
                             nproc=1, comm=comm, _backend=fake_device)
        return app.find(rows_per_band=20)

    for rows in run_shards(world, shard):
        assert [list(r) for r in rows] == g["rows"]


@pytest.mark.parametrize("tag", ["a", "b"])
def test_get_degeprimer(tmp_path, tag, capsys):
    """get_degePrimer drop-in against get_degePrimer_V6.py on a DegePrime table built from the C2 rows
    (tests/golden/make_golden.py degeprimer); the reference FASTA only supplies the sequence count (1000)"""
    from multiprime_b200 import degeprimer
    g = _load("pairs_get_degeprimer.json")
    inp = tmp_path / "degeprime.out"
    inp.write_text(g["table"])
    ref = tmp_path / "ref.fa"
    ref.write_text("".join(">s%d\nACGT\n" % i for i in range(1000)))
    out = tmp_path / "Cluster.candidate.primers.txt"
    want = g["runs"][tag]
    degeprimer.main(["-i", str(inp), "-r", str(ref), "-o", str(out)] + want["args"])
    assert out.read_text().replace(str(tmp_path), "<TMP>") == want["txt"]
    got = [ln for ln in capsys.readouterr().out.splitlines() if not ln.startswith("INFO")]
    assert got == want["stdout"]
