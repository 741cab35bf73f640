"""extract_PCR_product drop-in against the reference script's own outputs on test_data/test.fa
(tests/golden/pcr_product.json, made by tests/golden/make_golden.py pcr): plain, degenerate (2-, 3- and 4-fold codes),
a forward primer that occurs twice in a genome, a pair without product; formats seq and fa."""
import json
import os

import numpy as np
import pytest

from tests.helpers import GOLDEN, case_alignment, load_case


def _test_fa(tmp_path):
    """test_data/test.fa rebuilt from the committed fixture of the C1 case (one line per sequence, as the original)"""
    case = load_case("c1_testfa")
    ids, seqs = case_alignment(case, "c1_testfa")
    z = np.load(os.path.join(GOLDEN, "msa_c1_testfa.npz"), allow_pickle=True)
    heads = [str(x) for x in z["headers"]]
    seqs = [list(s) for s in seqs]
    for r, c in z["n_cells"].tolist():               # the fixture stores N as a gap cell (core:453)
        seqs[r][c] = "N"
    seqs = ["".join(s) for s in seqs]
    fa = tmp_path / "test.fa"
    fa.write_text("".join(h + "\n" + s + "\n" for h, s in zip(heads, seqs)))
    return str(fa)


def _run_all(tmp_path, backend):
    from multiprime_b200 import pcr_product
    g = json.load(open(os.path.join(GOLDEN, "pcr_product.json")))
    fa = _test_fa(tmp_path)
    for run, want in g["runs"].items():
        outdir, cov = tmp_path / ("out_" + run.replace(":", "_")), tmp_path / ("cov_" + run.replace(":", "_"))
        if run.startswith("seq:"):
            f, r = g["pairs"][run[4:]]
            argv = ["-r", fa, "-i", f + "," + r, "-f", "seq", "-o", str(outdir), "-s", str(cov), "-p", "1"]
        else:
            pf = tmp_path / "primers.fa"
            pf.write_text(want["primers_fa"])
            argv = ["-r", fa, "-i", str(pf), "-f", "fa", "-o", str(outdir), "-s", str(cov), "-p", "1"]
        pcr_product.main(argv, _backend=backend)
        got = {fn: open(os.path.join(outdir, fn)).read() for fn in sorted(os.listdir(outdir))}
        assert got == want["files"], run
        assert open(cov).read() == want["coverage"], run


def test_pcr_product_host_logic(tmp_path, capsys):
    from tests import fake_device
    _run_all(tmp_path, fake_device)


@pytest.mark.gpu
def test_pcr_product_gpu(tmp_path, capsys):
    _run_all(tmp_path, None)
