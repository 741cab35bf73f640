#!/usr/bin/env python
"""Generate the golden vectors that pin oracle/ (and, through it, the CUDA path) to the live reference.

Runs ONLY in the build container: it imports the reference scripts by path from /root/reference
(never copied) and writes small JSON / npz fixtures next to this file.

    PYTHONHASHSEED=0 python tests/golden/make_golden.py [case ...]

Outputs
  core_<case>.json     per-window records of NN_degenerate.get_primers (row, sidecar digests, the
                       primers handed to mis_primer_check in call order) + region start/stop
  msa_<case>.npz       the input alignment as 4-bit base sets + ids (inputs must travel to the GPU box)
  kat.json             known-answer vectors of the scalar formulas (Tm, dH/dS, dG, Loss, get_Y, ...)
  dimer_*.json / cover_*.json   finDimer_V4 / get_Maxprimerset_V1.3 outputs on committed primer sets
"""
from __future__ import annotations

import hashlib
import importlib.util
import json
import os
import sys
import tempfile
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

warnings.filterwarnings("ignore")
from multiprime_b200 import synth  # noqa: E402

CHARS = synth.CODE_CHARS
CHAR2CODE = {c: i for i, c in enumerate(CHARS)}


def load_ref(name, fname):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, "scripts", fname))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def digest(obj) -> str:
    """order-insensitive over dict keys (the reference's key order depends on PYTHONHASHSEED), order
    preserving inside the id lists"""
    return hashlib.sha256(json.dumps(obj, sort_keys=True).encode()).hexdigest()[:16]


CASES = {
    # name: (input, kwargs for NN_degenerate, window positions)
    "c2_k18": ("test_data/1000_fasta.msa",
               dict(primer_length=18, coverage=0.8, number_of_dege_bases=6, score_of_dege_bases=64,
                    raw_entropy_threshold=3.6, product_len=100, position="1,2,-1", variation=2, distance=4,
                    GC="0.2,0.7"),
               list(range(45, 85)) + list(range(300, 320)) + list(range(740, 757))),
    "c2_k22": ("test_data/1000_fasta.msa",
               dict(primer_length=22, coverage=0.8, number_of_dege_bases=6, score_of_dege_bases=64,
                    raw_entropy_threshold=3.6, product_len=100, position="1,2,-1", variation=2, distance=4,
                    GC="0.2,0.7"),
               list(range(45, 65)) + list(range(500, 510))),
    "c2_k20": ("test_data/1000_fasta.msa",
               dict(primer_length=20, coverage=0.8, number_of_dege_bases=6, score_of_dege_bases=64,
                    raw_entropy_threshold=3.6, product_len=100, position="1,2,-1", variation=2, distance=4,
                    GC="0.2,0.7"),
               list(range(200, 222)) + list(range(640, 650))),
    "c3_tmsa": ("test_data/results/Clusters_msa/Cluster_0_20727.tmsa",
                dict(primer_length=18, coverage=0.7, number_of_dege_bases=4, score_of_dege_bases=10,
                     raw_entropy_threshold=3.6, product_len=150, position="2,3,-1", variation=1, distance=4,
                     GC="0.2,0.7"),
                list(range(29, 60)) + list(range(590, 610)) + list(range(1480, 1500))),
    "c1_testfa": ("test_data/test.fa",
                  dict(primer_length=18, coverage=0.8, number_of_dege_bases=4, score_of_dege_bases=4,
                       raw_entropy_threshold=3.6, product_len=100, position="1,2,-1", variation=1, distance=4,
                       GC="0.2,0.7"),
                  list(range(0, 30)) + list(range(7290, 7310)) + list(range(8100, 8112))),
    # every window of the region: 1000_fasta.msa for k = 18..22, the whole
    # Cluster_0_20727.tmsa, and the first 10^4 rows of the north-star synthetic alignment with the C4 flags
    **{"c2f_k%d" % kk: ("test_data/1000_fasta.msa",
                        dict(primer_length=kk, coverage=0.8, number_of_dege_bases=6, score_of_dege_bases=64,
                             raw_entropy_threshold=3.6, product_len=100, position="1,2,-1", variation=2, distance=4,
                             GC="0.2,0.7"), None) for kk in (18, 19, 20, 21, 22)},
    "c3f_tmsa": ("test_data/results/Clusters_msa/Cluster_0_20727.tmsa",
                 dict(primer_length=18, coverage=0.7, number_of_dege_bases=4, score_of_dege_bases=10,
                      raw_entropy_threshold=3.6, product_len=150, position="2,3,-1", variation=1, distance=4,
                      GC="0.2,0.7"), None),
    "c4_10k": ("@synth:10000:600:20240923",
               dict(primer_length=18, coverage=0.8, number_of_dege_bases=8, score_of_dege_bases=256,
                    raw_entropy_threshold=3.6, product_len=100, position="1,2,-1", variation=3, distance=4,
                    GC="0.2,0.7"), None),
    "synth300": ("@synth:300:240:7",
                 dict(primer_length=18, coverage=0.8, number_of_dege_bases=8, score_of_dege_bases=256,
                      raw_entropy_threshold=3.6, product_len=100, position="1,2,-1", variation=3, distance=4,
                      GC="0.2,0.7"),
                 None),
    "synth_iupac": ("@synth:200:200:11:0.004:0.003",
                    dict(primer_length=20, coverage=0.8, number_of_dege_bases=6, score_of_dege_bases=64,
                         raw_entropy_threshold=3.6, product_len=100, position="1,-1", variation=2, distance=4,
                         GC="0.2,0.7"),
                    None),
}


# full-window cases read the alignment fixture of an earlier case instead of storing a second copy
MSA_ALIAS = {**{"c2f_k%d" % kk: "c2_k18" for kk in (18, 19, 20, 21, 22)}, "c3f_tmsa": "c3_tmsa"}


def materialise(inp: str, tmp: str) -> str:
    if not inp.startswith("@synth:"):
        return os.path.join(REF, inp)
    f = inp.split(":")[1:]
    n, L, seed = int(f[0]), int(f[1]), int(f[2])
    kw = {}
    if len(f) > 3:
        kw = dict(gap_rate=float(f[3]), iupac_rate=float(f[4]))
    codes = synth.synth_codes(n, L, seed=seed, **kw)
    path = os.path.join(tmp, "synth_%d_%d_%d.fa" % (n, L, seed))
    synth.write_fasta(path, codes)
    return path


def run_case(core, name):
    inp, kw, positions = CASES[name]
    with tempfile.TemporaryDirectory() as tmp:
        path = materialise(inp, tmp)
        app = core.NN_degenerate(seq_file=path, nproc=1, outfile=os.path.join(tmp, "x.out"), **kw)
        ids = list(app.seq_dict.keys())
        seqs = list(app.seq_dict.values())
        start, stop = int(app.start_position), int(app.stop_position)
        if positions is None:
            positions = list(range(start, stop - kw["primer_length"]))
        trace = []
        orig = app.mis_primer_check

        def wrapped(all_primers, primer, cover, ngsi):
            trace.append(primer)
            return orig(all_primers, primer, cover, ngsi)

        app.mis_primer_check = wrapped
        records = []
        for p in positions:
            trace.clear()
            app.get_primers(app.seq_dict, p)
            r = app.resQ.get()
            if r is None:
                records.append({"pos": p, "row": None, "trace": list(trace)})
                continue
            row = [r[0][0]] + list(r[0][1])
            row = [x.item() if hasattr(x, "item") else x for x in row]
            records.append({"pos": p, "row": row, "trace": list(trace),
                            "f_non": digest(r[1][1][0]), "r_non": digest(r[1][1][1]),
                            "gap_ids": digest(dict(r[2][1])),
                            "n_f_non": len(r[1][1][0]), "n_r_non": len(r[1][1][1])})
        L = max(len(s) for s in seqs)
        codes = np.zeros((len(seqs), L), dtype=np.uint8)
        lens = np.array([len(s) for s in seqs], dtype=np.int32)
        for i, s in enumerate(seqs):
            codes[i, :len(s)] = [CHAR2CODE[c] for c in s]
        packed = (codes[:, 0::2] | (np.pad(codes, ((0, 0), (0, L % 2)))[:, 1::2] << 4)).astype(np.uint8)
        alias = MSA_ALIAS.get(name)
        if not inp.startswith("@synth:") and alias is None:
            np.savez_compressed(os.path.join(HERE, "msa_%s.npz" % name), packed=packed, n_col=L, lens=lens,
                                ids=np.array(ids))
        with open(os.path.join(HERE, "core_%s.json" % name), "w") as fh:
            blob = {"input": inp, "params": kw, "start": start, "stop": stop, "n_seq": len(seqs), "records": records}
            if alias:
                blob["msa"] = alias
            json.dump(blob, fh, separators=(",", ":"))
        acc = sum(1 for r in records if r["row"] is not None)
        print(name, "windows", len(records), "accepted", acc, "region", start, stop)


def make_kat(core):
    kat = {"tm": {}, "dh_ds": {}, "dg": {}, "loss": [], "get_y": {}, "deg": {}, "filters": {}}
    seqs = ["ATGAAGACCATCATTGCC", "GGTACGGCCTCAGACATC", "ACGTACGTACGTACGT", "A" * 18, "GC" * 9, "TTTAAACAGCCTGTGGGT",
            "GCGCGCGC", "ACGTTGCA", "TTTTTTTTTTTTTTTTTTTTTT", "CAGTCAGTCAGTCAGTCAGT", "AATTAATT"]
    for s in seqs:
        kat["tm"][s] = core.Calc_Tm_v2(s)
        kat["dh_ds"][s] = list(core.Calc_deltaH_deltaS(s))
    app = core.NN_degenerate.__new__(core.NN_degenerate)
    app.distance = 4
    app.GC = ["0.2", "0.7"]
    for s in ["GCAACTGTTACC", "GCATC", "GGGTA", "ACGTTA", "ACGCGT", "T" * 14, "GCRYC", "ACGTWSTA", "GGCC", "ACGT",
              "CCGGYTA"]:
        kat["dg"][s] = app.deltaG(s)
    for a in [(12, 6, 0, 0), (5, 3, 0, 0), (5, 2, 0, 3), (18, 9, 0, 0), (5, 0, 0, 13), (7, 4, 0, 1), (9, 9, 0, 2)]:
        kat["loss"].append([list(a), core.Penalty_points(*a)])
    for coord, k in [("1,2,-1", 18), ("2,3,-1", 18), ("1,-1", 20), ("-2,4", 22)]:
        app.position, app.primer_length = coord, k
        f, r = app.get_Y()
        kat["get_y"]["%s|%d" % (coord, k)] = [sorted(f), sorted(r)]
    for s in ["GGTAYGGYYTCAGRCATC", "ACGTNNAC", "HBVDACGT", "AAAA"]:
        kat["deg"][s] = [core.score_trans(s), core.dege_number(s)]
    primers = ["TTTMAAMCAGCCTGTGGG", "ACYCACCCAMAGGGCCCA", "ATGAAGACYRTCATTGCY", "DATGGAWAAGCTTRCCGA",
               "ACACACACGGTTGGCCAA", "GGGGATCGATCGATTAGC", "ACGTTTTTCCCCCAAACGT", "CAGCAGCAGTTGACCATG",
               "GCGGCCGCTTTTGCGGCCGC", "AARYTTRCYGACCTCWAY", "ATATATATGCGTACGTAC", "ACTGACCCGGGTCAGTTT"]
    for p in primers:
        kat["filters"][p] = {"info": app.primer_pre_filter(p), "self_dimer": bool(app.dimer_check(p)),
                             "hairpin": bool(app.hairpin_check(p)), "repeat": bool(app.di_nucleotide(p)),
                             "gc": app.GC_fraction(p)}
    with open(os.path.join(HERE, "kat.json"), "w") as fh:
        json.dump(kat, fh, indent=1)
    print("kat written")


def make_dimer():
    """finDimer_V4 (Dimer.dimer_check called per position; V4 == V5 as sets) on primers harvested from the core cases"""
    fd = load_ref("findimer", "finDimer_V4.py")
    primers = []
    for name in CASES:
        path = os.path.join(HERE, "core_%s.json" % name)
        if os.path.exists(path):
            for rec in json.load(open(path))["records"]:
                if rec["row"]:
                    primers.append(rec["row"][3])
    core = load_ref("mpcore2", "multiPrime-core_V20.py")
    seen, keep = set(), []
    for p in primers:
        if p not in seen and core.score_trans(p) <= 16:
            seen.add(p)
            keep.append(p)
    keep = keep[:110]
    keep += ["ACGTACGTACGTACGTAC", "GGGGCCCCGGGGCCCCAT", "ATATATATGCGCGCGCAT", keep[3]]     # palindromes + a duplicate
    import random
    rnd = random.Random(4)
    comp = str.maketrans("ACGTRYMKSWHBVD", "TGCAYRKMSWDVBH")
    for t in range(40):                       # partners built to pair with the 3' end of an existing primer
        src = keep[rnd.randrange(100)]
        L = rnd.randrange(5, 13)
        tail = rnd.randrange(0, 4)
        rc = src[-L:].translate(comp)[::-1]
        body = "".join(rnd.choice("ACGT") for _ in range(20 - L - tail))
        keep.append(body + rc + "".join(rnd.choice("ACGT") for _ in range(tail)))
    with tempfile.TemporaryDirectory() as tmp:
        fa = os.path.join(tmp, "p.fa")
        with open(fa, "w") as fh:
            for i, p in enumerate(keep):
                fh.write(">P%03d\n%s\n" % (i, p))
        app = fd.Dimer(primer_file=fa, outfile=os.path.join(tmp, "o.txt"), threshold=3.96, nproc=1)
        rows = []
        for pos in range(len(app.primers_list)):
            app.dimer_check(pos)
            while True:
                r = app.resQ.get()
                if r is None:
                    break
                rows.append(list(r))
    with open(os.path.join(HERE, "dimer_findimer.json"), "w") as fh:
        json.dump({"primers": keep, "threshold": 3.96, "rows": rows}, fh, indent=0)
    print("finDimer rows", len(rows), "primers", len(keep))


def make_cover():
    """get_Maxprimerset_V1.3 (the script the pipeline calls) on clusters drawn from the finDimer primer pool"""
    import random
    import subprocess
    pool = json.load(open(os.path.join(HERE, "dimer_findimer.json")))["primers"]
    rnd = random.Random(9)
    lines = ["/data/empty_cluster.candidate.primers.txt"]
    for c in range(40):
        fields = ["/data/Cluster_%d.candidate.primers.txt" % c]
        for _ in range(rnd.randrange(1, 7)):
            f, r = rnd.choice(pool), rnd.choice(pool)
            start = rnd.randrange(0, 900)
            ln = rnd.randrange(150, 900)
            fields += [f, r, "%d:%.2f:%.3f" % (ln, rnd.uniform(48, 56), rnd.uniform(0.7, 1.0)), str(rnd.randrange(200, 500)),
                       "%d:%d" % (start, start + ln)]
        lines.append("\t".join(fields) + "\t")
    easy = []
    for c in range(14):
        fields = ["/data/Easy_%d.candidate.primers.txt" % c]
        for _ in range(rnd.randrange(5, 11)):
            f, r = rnd.choice(pool), rnd.choice(pool)
            fields += [f, r, "300:50.00:0.900", "400", "10:310"]
        easy.append("\t".join(fields) + "\t")
    blob = {"input": lines, "input_easy": easy}
    for mode in ("T", "F", "F_easy"):
        with tempfile.TemporaryDirectory() as tmp:
            inp = os.path.join(tmp, "candidate_primers_sets.txt")
            open(inp, "w").write("\n".join(easy if mode == "F_easy" else lines) + "\n")
            out = os.path.join(tmp, "final_maxprimers_set.xls")
            res = subprocess.run([sys.executable, os.path.join(REF, "scripts", "get_Maxprimerset_V1.3.py"), "-i", inp,
                                  "-o", out, "-s", "5", "-m", mode[0]], capture_output=True, text=True)
            blob[mode] = {"rc": res.returncode, "stdout": res.stdout,
                          "out": open(out).read() if os.path.exists(out) else None,
                          "sort": open(os.path.join(tmp, "sort.candidate_primers_sets.txt")).read()}
            nxt = os.path.join(tmp, "final_maxprimers_set.next.xls")
            blob[mode]["next"] = open(nxt).read() if os.path.exists(nxt) else None
            print("cover mode", mode, "rc", res.returncode, "rows", (blob[mode]["out"] or "").count("\n") - 1,
                  res.stderr[-300:])
    with open(os.path.join(HERE, "cover_maxprimerset.json"), "w") as fh:
        json.dump(blob, fh)


def make_pairs():
    """get_multiPrime.py (V8) on the reference core's own output for a 80 x 420 synthetic alignment"""
    import subprocess
    n, L, seed, gr, ir = 80, 420, 17, 0.003, 0.001
    blob = {"synth": [n, L, seed, gr, ir]}
    with tempfile.TemporaryDirectory() as tmp:
        fa = os.path.join(tmp, "in.fa")
        synth.write_fasta(fa, synth.synth_codes(n, L, seed=seed, gap_rate=gr, iupac_rate=ir))
        core_out = os.path.join(tmp, "c.out")
        subprocess.run([sys.executable, os.path.join(REF, "scripts", "multiPrime-core.py"), "-i", fa, "-o", core_out,
                        "-l", "18", "-n", "4", "-d", "10", "-v", "1", "-p", "1"], check=True, capture_output=True)
        blob["core_tsv"] = open(core_out).read()
        blob["core_non_cov"] = json.load(open(core_out + ".non_coverage_seq_id_json"))
        blob["core_gap"] = json.load(open(core_out + ".gap_seq_id_json"))
        for tag, extra in (("a", ["-s", "100,300", "-f", "0.3", "-t", "6", "-e", "2"]),
                           ("b", ["-s", "150,260", "-f", "0.97", "-t", "3", "-a", ","])):
            out = os.path.join(tmp, "Cluster_%s.candidate.primers.txt" % tag)
            res = subprocess.run([sys.executable, os.path.join(REF, "scripts", "get_multiPrime.py"), "-i", core_out, "-r", fa,
                                  "-o", out] + extra, capture_output=True, text=True)
            stem = out.strip(".txt")
            blob[tag] = {"args": extra, "rc": res.returncode, "stdout": res.stdout.replace(tmp, "<TMP>"),
                         "txt": open(out).read().replace(tmp, "<TMP>"),
                         "xls": open(stem + ".xls").read() if os.path.exists(stem + ".xls") else None,
                         "fa": open(stem + ".fa").read() if os.path.exists(stem + ".fa") else None}
            print("pairs", tag, "rc", res.returncode, "rows", (blob[tag]["xls"] or "").count("\n") - 1, res.stderr[-200:])
    with open(os.path.join(HERE, "pairs_get_multiprime.json"), "w") as fh:
        json.dump(blob, fh)


def pcr_primer_sets():
    """primer pairs cut from test_data/test.fa itself (so that they amplify), some made degenerate, one pair whose forward
    primer occurs twice in a genome, one with N, one that matches nothing"""
    seqs = []
    with open(os.path.join(REF, "test_data", "test.fa")) as fh:
        for line in fh:
            if not line.startswith(">"):
                seqs.append(line.strip())
    comp = str.maketrans("ATGC", "TACG")
    rc = lambda x: x.translate(comp)[::-1]
    s0, s4 = seqs[0], seqs[4]
    pairs = {}
    with_base = {"A": "RMWNHVD", "C": "YMSHBVN", "G": "RKSBVDN", "T": "YKWHBDN"}

    def dege(seq, spots):
        """seq with the bases at `spots` replaced by the (spot-th, cyclically) IUPAC code that contains them"""
        out = list(seq)
        for n, i in enumerate(spots):
            out[i] = with_base[seq[i]][(n + i) % 7]
        return "".join(out)

    f, r = s0[100:118], rc(s0[400:420])
    pairs["plain_0_100"] = (f, r)
    pairs["dege_0_100"] = (dege(f, [5, 11]), dege(r, [3, 9]))
    f4, r4 = s4[2000:2020], rc(s4[2300:2318])
    pairs["dege_4_2000"] = (dege(f4, [2, 10, 15]), dege(r4, [8]))
    # the most frequent 12-mer of genome 0 as forward primer: it occurs more than once
    from collections import Counter
    cnt = Counter(s0[i:i + 12] for i in range(len(s0) - 300))
    kmer, n = [(km, c) for km, c in cnt.most_common() if len(set(km)) > 2][0]
    assert n >= 2
    first = s0.index(kmer)
    pairs["repeat_0"] = (kmer, rc(s0[first + 40:first + 58]))
    pairs["nomatch"] = ("ACGTACGTACGTACGTAC", "TTTTTTTTGGGGGGGGCC")
    pairs["hbvdn_0_3000"] = (dege(s0[3000:3019], [0, 6, 12, 17]), dege(rc(s0[3200:3220]), [1, 7, 13]))
    return pairs


def make_pcr():
    """extract_PCR_product_V1.py on test_data/test.fa: formats seq and fa, outputs recorded file by file"""
    import subprocess
    pairs = pcr_primer_sets()
    blob = {"pairs": pairs, "runs": {}}
    ref = os.path.join(REF, "test_data", "test.fa")
    script = os.path.join(REF, "scripts", "extract_PCR_product_V1.py")

    def collect(tmp, outdir, cov):
        files = {}
        for fn in sorted(os.listdir(outdir)):
            files[fn] = open(os.path.join(outdir, fn)).read()
        return {"files": files, "coverage": open(cov).read()}

    for name, (f, r) in pairs.items():
        with tempfile.TemporaryDirectory() as tmp:
            outdir, cov = os.path.join(tmp, "PCR"), os.path.join(tmp, "Coverage.xls")
            res = subprocess.run([sys.executable, script, "-r", ref, "-i", f + "," + r, "-f", "seq", "-o", outdir, "-s", cov,
                                  "-p", "1"], capture_output=True, text=True)
            blob["runs"]["seq:" + name] = dict(collect(tmp, outdir, cov), rc=res.returncode)
            print("pcr", name, res.returncode, blob["runs"]["seq:" + name]["coverage"].splitlines()[0][:120], res.stderr[-200:])
    with tempfile.TemporaryDirectory() as tmp:
        fa = os.path.join(tmp, "primers.fa")
        with open(fa, "w") as fh:
            for name, (f, r) in pairs.items():
                fh.write(">%s_F\n%s\n>%s_R\n%s\n" % (name, f, name, r))
        outdir, cov = os.path.join(tmp, "PCR"), os.path.join(tmp, "Coverage.xls")
        res = subprocess.run([sys.executable, script, "-r", ref, "-i", fa, "-f", "fa", "-o", outdir, "-s", cov, "-p", "1"],
                             capture_output=True, text=True)
        blob["runs"]["fa"] = dict(collect(tmp, outdir, cov), rc=res.returncode, primers_fa=open(fa).read())
        print("pcr fa", res.returncode, res.stderr[-300:])
    with open(os.path.join(HERE, "pcr_product.json"), "w") as fh:
        json.dump(blob, fh)


def make_degeprimer():
    """get_degePrimer_V6.py on a DegePrime-format table built from the full C2 golden rows (k = 18) and
    test_data/1000.fasta as the reference FASTA; two parameter sets"""
    import subprocess
    case = json.load(open(os.path.join(HERE, "core_c2f_k18.json")))
    lines = ["Pos\tTotalSeq\tUniqueMers\tEntropy\tPrimerDeg\tPrimerSeq\tPrimerMatching"]
    for rec in case["records"]:
        if rec["row"]:
            r = rec["row"]
            lines.append("\t".join(map(str, [r[0], case["n_seq"], r[5] + 1, r[2], r[4], r[3], max(r[7], r[8])])))
    table = "\n".join(lines) + "\n"
    blob = {"table": table, "runs": {}}
    ref = os.path.join(REF, "test_data", "1000.fasta")
    for tag, extra in (("a", ["-s", "100,300", "-f", "0.3", "-e", "2", "-m", "0"]),
                       ("b", ["-s", "150,400", "-f", "0.6", "-a", ",", "-d", "3"])):
        with tempfile.TemporaryDirectory() as tmp:
            inp, out = os.path.join(tmp, "degeprime.out"), os.path.join(tmp, "Cluster.candidate.primers.txt")
            open(inp, "w").write(table)
            res = subprocess.run([sys.executable, os.path.join(REF, "scripts", "get_degePrimer_V6.py"), "-i", inp, "-r", ref,
                                  "-o", out] + extra, capture_output=True, text=True)
            blob["runs"][tag] = {"args": extra, "rc": res.returncode, "txt": open(out).read().replace(tmp, "<TMP>"),
                                 "stdout": [ln for ln in res.stdout.splitlines() if not ln.startswith("INFO")]}
            print("degeprimer", tag, res.returncode, blob["runs"][tag]["txt"].count(":") , res.stderr[-200:])
    with open(os.path.join(HERE, "pairs_get_degeprimer.json"), "w") as fh:
        json.dump(blob, fh)


def make_cli():
    """the reference CLI end to end on a small synthetic alignment: TSV text + the two JSON side files"""
    import subprocess
    codes = synth.synth_codes(60, 140, seed=3, gap_rate=0.004, iupac_rate=0.002)
    with tempfile.TemporaryDirectory() as tmp:
        fa = os.path.join(tmp, "in.fa")
        synth.write_fasta(fa, codes)
        out = os.path.join(tmp, "ref.out")
        subprocess.run([sys.executable, os.path.join(REF, "scripts", "multiPrime-core.py"), "-i", fa, "-o", out,
                        "-l", "18", "-n", "6", "-d", "64", "-v", "2", "-p", "1"], check=True, capture_output=True)
        blob = {"args": ["-l", "18", "-n", "6", "-d", "64", "-v", "2", "-p", "1"], "synth": [60, 140, 3, 0.004, 0.002],
                "tsv": open(out).read(), "non_cov": json.load(open(out + ".non_coverage_seq_id_json")),
                "gap": json.load(open(out + ".gap_seq_id_json"))}
    with open(os.path.join(HERE, "cli_core.json"), "w") as fh:
        json.dump(blob, fh)
    print("cli rows", blob["tsv"].count("\n") - 1)


def main():
    assert os.environ.get("PYTHONHASHSEED") == "0", "run with PYTHONHASHSEED=0"
    core = load_ref("mpcore", "multiPrime-core_V20.py")
    which = sys.argv[1:] or (["kat"] + list(CASES))
    for name in which:
        if name == "kat":
            make_kat(core)
        elif name == "dimer":
            make_dimer()
        elif name == "cli":
            make_cli()
        elif name == "cover":
            make_cover()
        elif name == "pairs":
            make_pairs()
        elif name == "pcr":
            make_pcr()
        elif name == "degeprimer":
            make_degeprimer()
        elif name in CASES:
            run_case(core, name)


if __name__ == "__main__":
    main()
