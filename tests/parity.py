"""the parity check shared by the CPU (fake device) and GPU (libmpb200) suites: run multiprime_b200.core on a
golden case and compare every window with the reference's record"""
from __future__ import annotations

import numpy as np

from multiprime_b200 import core
from multiprime_b200.iupac import CHAR_CODE
from tests.helpers import case_alignment, digest, load_case


def alignment_arrays(ids, seqs):
    L = max(len(s) for s in seqs)
    codes = np.zeros((len(seqs), L), np.uint8)
    lens = np.array([len(s) for s in seqs], np.int32)
    for i, s in enumerate(seqs):
        codes[i, :len(s)] = [CHAR_CODE[c] for c in s]
    return ids, codes, lens


def make_app(case, name, backend=None, **extra):
    ids, seqs = case_alignment(case, name)
    kw = dict(case["params"])
    return core.NN_degenerate(seq_file=None, nproc=1, outfile="", alignment=alignment_arrays(ids, seqs),
                              _backend=backend, **kw, **extra)


def check_case(name, backend=None, max_windows=None, **extra):
    case = load_case(name)
    app = make_app(case, name, backend, **extra)
    assert (app.start_position, app.stop_position) == (case["start"], case["stop"])
    records = case["records"][:max_windows] if max_windows else case["records"]
    got = {r["row"][0]: r for r in app.design([r["pos"] for r in records])}
    for rec in records:
        g = got.get(rec["pos"])
        if rec["row"] is None:
            assert g is None, (name, rec["pos"], g and g["row"])
            continue
        assert g is not None, (name, rec["pos"], rec["row"])
        assert g["row"] == rec["row"], (name, rec["pos"], g["row"], rec["row"])
        assert g["trace"] == rec["trace"], (name, rec["pos"])
        assert digest(g["non_cov"][0]) == rec["f_non"], (name, rec["pos"], "F side file")
        assert digest(g["non_cov"][1]) == rec["r_non"], (name, rec["pos"], "R side file")
        assert digest(g["gap_ids"]) == rec["gap_ids"], (name, rec["pos"], "gap side file")
    app.close()
    return app.stats
