"""shared helpers for the tests: golden loading, synthetic inputs, digests"""
from __future__ import annotations

import hashlib
import json
import os

import numpy as np

from multiprime_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def digest(obj) -> str:
    return hashlib.sha256(json.dumps(obj, sort_keys=True).encode()).hexdigest()[:16]


def load_case(name: str):
    with open(os.path.join(GOLDEN, "core_%s.json" % name)) as fh:
        return json.load(fh)


def case_alignment(case: dict, name: str):
    """(ids, strings) of the alignment a golden case was generated from"""
    inp = case["input"]
    if inp.startswith("@synth:"):
        f = inp.split(":")[1:]
        n, L, seed = int(f[0]), int(f[1]), int(f[2])
        kw = dict(gap_rate=float(f[3]), iupac_rate=float(f[4])) if len(f) > 3 else {}
        codes = synth.synth_codes(n, L, seed=seed, **kw)
        return synth.seq_ids(n), synth.codes_to_strings(codes)
    z = np.load(os.path.join(GOLDEN, "msa_%s.npz" % case.get("msa", name)))
    packed, L, lens = z["packed"], int(z["n_col"]), z["lens"]
    codes = np.empty((packed.shape[0], packed.shape[1] * 2), dtype=np.uint8)
    codes[:, 0::2] = packed & 15
    codes[:, 1::2] = packed >> 4
    strs = synth.codes_to_strings(codes[:, :L])
    return [str(x) for x in z["ids"]], [s[:n] for s, n in zip(strs, lens)]


def oracle_params(kw: dict):
    from oracle import mp_oracle as o
    return o.Params(k=kw["primer_length"], dnum=kw["number_of_dege_bases"], degeneracy=kw["score_of_dege_bases"],
                    variation=kw["variation"], entropy=kw["raw_entropy_threshold"], gc=kw["GC"],
                    size=kw["product_len"], fraction=kw["coverage"], coordinate=kw["position"],
                    away=kw["distance"])
