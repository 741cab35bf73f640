"""GPU suite: libmpb200 (through the C ABI) against the reference's golden records and against the oracle."""
import numpy as np
import pytest

from tests.parity import check_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["synth_iupac", "synth300", "c2_k18", "c2_k20", "c2_k22", "c3_tmsa", "c1_testfa"])
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


def test_findimer_golden(tmp_path):
    """finDimer: rows of the reference (finDimer_V4 == V5 as sets) on 154 primers, incl. duplicates and degenerate ones"""
    import json
    import os
    from multiprime_b200 import findimer
    from tests.helpers import GOLDEN
    g = json.load(open(os.path.join(GOLDEN, "dimer_findimer.json")))
    fa = tmp_path / "p.fa"
    fa.write_text("".join(">P%03d\n%s\n" % (i, p) for i, p in enumerate(g["primers"])))
    out = tmp_path / "o.txt"
    app = findimer.Dimer(primer_file=str(fa), outfile=str(out), threshold=g["threshold"], nproc=1)
    rows = app.run()
    assert [list(r) for r in rows] == g["rows"]                 # V5 order: (i, j) ascending
    text = out.read_text().splitlines()
    assert text[0].split("\t") == findimer.HEADERS
    assert [ln.split("\t") for ln in text[1:]] == [[str(x) for x in r] for r in g["rows"]]
    assert (tmp_path / "o.txt.dimer_num").read_text().startswith("SeqName\tPrimer_ID\tDimer-primer_ID\tRowSum\n")


def test_core_cli_bytes(tmp_path):
    """the drop-in CLI against the reference CLI's own output files on the same FASTA: TSV byte-identical,
    JSON side files equal as dictionaries (the reference's key order depends on PYTHONHASHSEED)"""
    import json
    import os
    import subprocess
    import sys
    from multiprime_b200 import synth
    from tests.helpers import GOLDEN
    g = json.load(open(os.path.join(GOLDEN, "cli_core.json")))
    n, L, seed, gr, ir = g["synth"]
    fa = tmp_path / "in.fa"
    synth.write_fasta(str(fa), synth.synth_codes(n, L, seed=seed, gap_rate=gr, iupac_rate=ir))
    out = tmp_path / "mine.out"
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "scripts", "multiPrime-core.py"), "-i", str(fa), "-o",
                          str(out)] + g["args"], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    assert res.stdout.strip().startswith("INFO ") and "Total times:" in res.stdout
    assert out.read_text() == g["tsv"]
    assert json.load(open(str(out) + ".non_coverage_seq_id_json")) == g["non_cov"]
    assert json.load(open(str(out) + ".gap_seq_id_json")) == g["gap"]


@pytest.mark.parametrize("mode", ["T", "F", "F_easy"])
def test_maxprimerset_golden(tmp_path, mode):
    """get_Maxprimerset drop-in against get_Maxprimerset_V1.3.py: output, .next.xls and sort.* byte-identical, same
    exit code (maximum mode without a solution exits 1 with the reference's message)"""
    import json
    import os
    import subprocess
    import sys
    from tests.helpers import GOLDEN
    g = json.load(open(os.path.join(GOLDEN, "cover_maxprimerset.json")))
    lines = g["input_easy"] if mode == "F_easy" else g["input"]
    inp = tmp_path / "candidate_primers_sets.txt"
    inp.write_text("\n".join(lines) + "\n")
    out = tmp_path / "final_maxprimers_set.xls"
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "scripts", "get_Maxprimerset.py"), "-i", str(inp), "-o",
                          str(out), "-s", "5", "-m", mode[0]], capture_output=True, text=True)
    want = g[mode]
    assert res.returncode == want["rc"], res.stderr[-2000:]
    assert (tmp_path / "sort.candidate_primers_sets.txt").read_text() == want["sort"]
    if want["out"] is not None:
        assert out.read_text() == want["out"]
    if want["next"] is not None:
        assert (tmp_path / "final_maxprimers_set.next.xls").read_text() == want["next"]
    ref_msgs = [ln for ln in want["stdout"].splitlines() if not ln.startswith("INFO")]
    got_msgs = [ln for ln in res.stdout.splitlines() if not ln.startswith("INFO")]
    assert got_msgs == ref_msgs


def test_primer_props_vs_oracle():
    """mpb_primer_props (Tm mean, GC, di-nucleotide / hairpin flags) on random degenerate primers == oracle"""
    import random
    from statistics import mean
    from multiprime_b200 import _lib, core
    from multiprime_b200.iupac import sets_of
    from oracle import mp_oracle as o
    random.seed(23)
    codes = "ACGTRYMKSWHBVDN"
    ctx = _lib.Context(0)
    for k in (18, 20, 22):
        primers = []
        while len(primers) < 150:
            p = "".join(random.choice("ACGT" * 6 + codes) for _ in range(k))
            if random.random() < 0.3:                       # plant repeats / hairpins
                i = random.randrange(0, k - 9)
                p = p[:i] + random.choice(["ACACACAC", "GGGG", "CAGCAGCAG", "TTTT", "ATGATGATG"]) + p[i + 9:]
                p = p[:k].ljust(k, "A")
            if o.degeneracy(p) <= 256:
                primers.append(p)
        arr = np.zeros((len(primers), 32), np.uint8)
        for i, p in enumerate(primers):
            arr[i, :k] = sets_of(p)
        tm, gc, flags, deg, ndeg = ctx.primer_props(arr, k, 0.2, 0.7, 4, core.TM_CONSTS)
        for i, p in enumerate(primers):
            assert not flags[i] & (64 | 128) or True
            assert (deg[i], ndeg[i]) == (o.degeneracy(p), o.n_degenerate(p))
            if not flags[i] & 64:
                assert tm[i] == round(mean([o.tm(e) for e in o.expand(p)]), 2), p
            if not flags[i] & 128:
                assert gc[i] == o.gc_content(p), p
            assert bool(flags[i] & 2) == o.has_repeat(p), p
            assert bool(flags[i] & 4) == o.has_hairpin(p, 4), p
    ctx.close()


@pytest.mark.parametrize("tag", ["a", "b"])
def test_get_multiprime_golden(tmp_path, tag):
    """get_multiPrime drop-in on the reference core's own output: the three output files byte-identical (second case
    takes the `fewer than 10 pairs -> repeat with a looser threshold` branch), stdout lines equal"""
    import json
    import os
    import subprocess
    import sys
    from multiprime_b200 import synth
    from tests.helpers import GOLDEN
    g = json.load(open(os.path.join(GOLDEN, "pairs_get_multiprime.json")))
    n, L, seed, gr, ir = g["synth"]
    fa = tmp_path / "in.fa"
    synth.write_fasta(str(fa), synth.synth_codes(n, L, seed=seed, gap_rate=gr, iupac_rate=ir))
    core_out = tmp_path / "c.out"
    core_out.write_text(g["core_tsv"])
    json.dump(g["core_non_cov"], open(str(core_out) + ".non_coverage_seq_id_json", "w"))
    json.dump(g["core_gap"], open(str(core_out) + ".gap_seq_id_json", "w"))
    out = tmp_path / ("Cluster_%s.candidate.primers.txt" % tag)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    want = g[tag]
    res = subprocess.run([sys.executable, os.path.join(root, "scripts", "get_multiPrime.py"), "-i", str(core_out), "-r",
                          str(fa), "-o", str(out)] + want["args"], capture_output=True, text=True)
    assert res.returncode == want["rc"], res.stderr[-2000:]
    stem = str(out).strip(".txt")
    assert out.read_text().replace(str(tmp_path), "<TMP>") == want["txt"]
    assert open(stem + ".xls").read() == want["xls"]
    assert open(stem + ".fa").read() == want["fa"]
    strip = lambda text: [ln for ln in text.replace(str(tmp_path), "<TMP>").splitlines() if not ln.startswith("INFO")]
    assert strip(res.stdout) == strip(want["stdout"])


@pytest.mark.parametrize("tag,world", [("a", 1), ("b", 1), ("a", 2)])
def test_core_to_pairing_through_bit_vectors(tmp_path, tag, world):
    """the core CLI writes per-sequence bit vectors (--sidecars bits) instead of the JSON side files
    and the get_multiPrime drop-in takes its pair coverage from them on the device; with two ranks (torchrun, gloo, one
    GPU) both steps run sequence-sharded.  The reference's three output files come out byte for byte."""
    import json
    import os
    import subprocess
    import sys
    from multiprime_b200 import synth
    from tests.helpers import GOLDEN
    g = json.load(open(os.path.join(GOLDEN, "pairs_get_multiprime.json")))
    n, L, seed, gr, ir = g["synth"]
    fa = tmp_path / "in.fa"
    synth.write_fasta(str(fa), synth.synth_codes(n, L, seed=seed, gap_rate=gr, iupac_rate=ir))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    core_out = tmp_path / "c.out"
    env = dict(os.environ, MPB_DIST_BACKEND="gloo")
    launch = [sys.executable]
    if world > 1:
        import socket
        sk = socket.socket()
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
        sk.close()
        launch = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                  "--master-addr", "127.0.0.1", "--master-port", str(port)]
    res = subprocess.run(launch + [os.path.join(root, "scripts", "multiPrime-core.py"), "-i", str(fa), "-o", str(core_out),
                                   "-l", "18", "-n", "4", "-d", "10", "-v", "1", "-p", "1", "--sidecars", "bits"],
                         capture_output=True, text=True, env=env, timeout=600)
    assert res.returncode == 0, res.stderr[-3000:]
    assert core_out.read_text() == g["core_tsv"]
    assert not os.path.exists(str(core_out) + ".gap_seq_id_json")
    out = tmp_path / ("Cluster_%s.candidate.primers.txt" % tag)
    want = g[tag]
    res = subprocess.run(launch + [os.path.join(root, "scripts", "get_multiPrime.py"), "-i", str(core_out), "-r", str(fa),
                                   "-o", str(out)] + want["args"], capture_output=True, text=True, env=env, timeout=600)
    assert res.returncode == want["rc"], res.stderr[-3000:]
    stem = str(out).strip(".txt")
    assert out.read_text().replace(str(tmp_path), "<TMP>") == want["txt"]
    assert open(stem + ".xls").read() == want["xls"]
    assert open(stem + ".fa").read() == want["fa"]
    strip = lambda text: [ln for ln in text.replace(str(tmp_path), "<TMP>").splitlines() if not ln.startswith("INFO")]
    assert strip(res.stdout) == strip(want["stdout"])


@pytest.mark.parametrize("mode", ["bs", "row"])
def test_prefilter_matches_definition(mode, monkeypatch):
    """mpb_window_prefilter (bit-sliced cluster kernel / row-domain kernel): item counts and sum(c log2 c) of the coarse
    bins equal the plain-Python definition, and the bound never exceeds the reference's total entropy"""
    monkeypatch.setenv("MPB_WINPASS", "row" if mode == "row" else "col")
    from multiprime_b200 import _lib, core
    from tests import fake_device
    from tests.helpers import case_alignment, load_case
    from tests.parity import alignment_arrays
    case = load_case("synth_iupac")
    ids, seqs = case_alignment(case, "synth_iupac")
    _, codes, lens = alignment_arrays(ids, seqs)
    k, v = case["params"]["primer_length"], case["params"]["variation"]
    pos = [r["pos"] for r in case["records"]]
    ctx = _lib.Context(0)
    msa = _lib.Msa(ctx, core.pack4(codes), len(ids), codes.shape[1])
    s0, s1 = msa.prefilter(k, v, pos)
    fmsa = fake_device.Msa(None, core.pack4(codes), len(ids), codes.shape[1])
    f0, f1 = fmsa.prefilter(k, v, pos, code=mode)
    assert (s0 == f0).all()
    assert np.allclose(s1, f1, rtol=1e-12, atol=1e-9)
    n = len(ids)
    bound = (s0 * np.log2(n) - s1) / n
    for rec, b in zip(case["records"], bound):
        if rec["row"] is not None:
            assert b <= rec["row"][2] + 0.006          # row[2] = Entropy of total (bit), rounded to 2 decimals
    msa.close()
    ctx.close()
