"""Drop-in for scripts/get_multiPrime.py (get_multiPrime_V8.py): pair forward / reverse candidates of one cluster.

Same flags and output files (<out>, <out>.strip(".txt") + ".xls" / ".fa").  Per candidate pair the reference applies, in
this order (get_multiPrime.py:509-582): adaptor hairpin, 3'-degenerate, GC clamp, product size, F-R dimer
(Loss > 3.6 or dG < -5 at distance 0, dG with the first base's initiation term only), |dTm|, pair coverage from the
id lists of the two JSON side files.  Here the dimer test runs on the GPU dimer engine (all pairs of a start in one
batch) and the coverage is popcount(uncovered_F | uncovered_R) on per-sequence bit vectors (mpb_pair_cover);
the cheap string filters stay on the host, evaluated on base sets without expansion.

Quirks kept on purpose: `-g` is parsed but the class default "0.4,0.6" is what filters (get_multiPrime.py:665-671 never
passes it); the output names use str.strip(".txt"); when fewer than 10 pairs pass, the whole pairing is repeated with
the threshold raised by 0.1 and the pairs are APPENDED (duplicates included)."""
from __future__ import annotations

import argparse
import json
import os
import time
from bisect import bisect_left

import numpy as np

from . import _lib
from .core import exact_mean, has_hairpin, has_repeat
from .dimer import dg_consts, loss_table
from .iupac import FOLD, sets_of


def parseArg(argv=None):
    parser = argparse.ArgumentParser(description="For degenerate primer design")
    parser.add_argument("-i", "--input", type=str, required=True, help="Input file: multiPrime out.", metavar="<file>")
    parser.add_argument("-r", "--ref", type=str, required=True,
                        help="Reference sequence file: all the sequence in 1 fasta.", metavar="<str>")
    parser.add_argument("-g", "--gc", type=str, default="0.2,0.7", help="Filter primers by GC content.", metavar="<str>")
    parser.add_argument("-f", "--fraction", type=float, default=0.6, help="Filter primers by match fraction. Default: 0.6.",
                        metavar="<float>")
    parser.add_argument("-e", "--end", type=int, default=4,
                        help="No degenerate base within the last N bases. Default: 4.", metavar="<int>")
    parser.add_argument("-p", "--proc", type=int, default=20, help="accepted for compatibility", metavar="<int>")
    parser.add_argument("-s", "--size", type=str, default="250,500", help="Filter primers by PRODUCT size. Default [250,500].",
                        metavar="<str>")
    parser.add_argument("-d", "--dist", type=int, default=4, help="Hairpin: distance of the minimal paired bases. Default: 4.",
                        metavar="<int>")
    parser.add_argument("-t", "--Tm", type=int, default=4, help="Max difference of Tm between primer-F and primer-R. "
                                                                "Default: 4.", metavar="<int>")
    parser.add_argument("-a", "--adaptor", type=str,
                        default="TCTTTCCCTACACGACGCTCTTCCGATCT,TCTTTCCCTACACGACGCTCTTCCGATCT",
                        help="Adaptor sequences F,R (',' for none).", metavar="<str>")
    parser.add_argument("-m", "--maxseq", type=int, default=0, help="Limit of sequence number. Default: 0 (all).",
                        metavar="<int>")
    parser.add_argument("-o", "--out", type=str, required=True, help="Output file: candidate primers.", metavar="<file>")
    parser.add_argument("--device", type=int, default=0, help=argparse.SUPPRESS)
    return parser.parse_args(argv)


def rc_string(seq: str) -> str:
    return seq.translate(str.maketrans("ATGCRYMKSWHBVDN", "TACGYRKMSWDVBHN"))[::-1]


def gc_mean(sets) -> float:
    """get_multiPrime.py:450-456 GC_fraction: exact mean over expansions of round(gc/len, 3) — NOT rounded"""
    k = len(sets)
    dist = [1]
    for s in sets:
        n_gc = ((s >> 1) & 1) + ((s >> 2) & 1)
        n_at = (s & 1) + ((s >> 3) & 1)
        new = [0] * (len(dist) + 1)
        for g, m in enumerate(dist):
            new[g] += m * n_at
            new[g + 1] += m * n_gc
        dist = new
    vals = []
    for g, m in enumerate(dist):
        if m:
            vals.append((round(g / k, 3), m))
    total = sum(m for _, m in vals)
    acc = sum(m * int(v * 1152921504606846976.0) for v, m in vals)
    return acc / (total << 60)


def gc_clamp(sets, num=4, length=13) -> bool:
    """get_multiPrime.py:467-473"""
    for i in range(num, num + length):
        if gc_mean(sets[-i:]) > 0.6:
            return True
    return False


def term_degenerate(sets, term: int) -> bool:
    """get_multiPrime.py:439-448"""
    if term == 0:
        return False
    d = 1
    for s in sets[-term:]:
        d *= FOLD[s]
    return d > 1


class Primers_filter(object):
    """get_multiPrime.py:303-321 constructor arguments"""

    def __init__(self, ref_file, primer_file, adaptor, rep_seq_number=500, distance=4, outfile="", diff_Tm=5,
                 size="300,700", position=9, GC="0.4,0.6", nproc=10, fraction=0.6, device=0, comm=None, coverage=None,
                 rows=None, number=None, _backend=None):
        self.nproc = nproc
        self.primer_file = primer_file
        self.adaptor = adaptor
        self.size = size
        self.outfile = os.path.abspath(outfile)
        self.distance = distance
        self.Input_file = ref_file
        self.fraction = fraction
        self.GC = GC
        self.diff_Tm = diff_Tm
        self.rep_seq_number = rep_seq_number
        self.rows = rows                          # the core step's rows handed over in-process (else read from primer_file)
        self.number = number if number is not None else self.get_number()
        self.position = position
        from .comm import NoComm
        self.comm = comm or NoComm()              # sequence shards: every rank holds the bit vectors of its sequences
        self.coverage = coverage                  # (positions, bits[n, 3, words]) handed over in-process, or None
        self.primers, self.gap_id, self.non_cover_id = self.parse_primers()
        self._backend = _backend or _lib          # tests inject tests/fake_device.py
        self.ctx = self._backend.Context.shared(device) if hasattr(self._backend.Context, "shared") else \
            self._backend.Context(device)
        self.pre_filter_primers = self.pre_filter()

    @classmethod
    def from_core(cls, app, recs, outfile, **kw):
        """pair the rows of a core run (multiprime_b200.core.NN_degenerate, keep_bits=True) in the same process: the
        scan's bit vectors are read where design() left them (in HBM), no files in between"""
        pos = np.concatenate([np.asarray(p, np.int32) for p, _ in app.bit_vectors]) if app.bit_vectors else np.zeros(0, np.int32)
        bits = app.bit_vectors[0][1] if len(app.bit_vectors) == 1 else app.coverage_bits()[1]
        return cls(ref_file=None, primer_file=None, outfile=outfile, rows=[r["row"] for r in recs], coverage=(pos, bits),
                   number=app.total_sequence_number, comm=app.comm if app.comm.world > 1 else None,
                   device=getattr(app.ctx, "device", 0), **kw)

    def parse_primers(self):
        primer_dict = {}
        if self.rows is not None:
            for r in self.rows:
                primer_dict[int(r[0])] = [r[3], round(int(r[6]) / self.number, 2), int(r[7]), int(r[8]),
                                          round(float(r[9]), 2)]
            return primer_dict, None, None
        with open(self.primer_file) as f:
            for line in f:
                if line.startswith("Pos"):
                    continue
                i = line.strip().split("\t")
                primer_dict[int(i[0])] = [i[3], round(int(i[6]) / self.number, 2), int(i[7]), int(i[8]),
                                          round(float(i[9]), 2)]
        # coverage information of the core step: its per-sequence bit vectors when they exist (handed over in-process,
        # or <input>.coverage_bits*.npz written above core.SIDE_JSON_MAX sequences), else the reference's JSON side files
        from .core import bits_file
        if self.coverage is None:
            path = bits_file(self.primer_file, self.comm.rank, self.comm.world)
            if os.path.exists(path):
                z = np.load(path)
                if int(z["world"]) != self.comm.world:
                    raise SystemExit("Error: %s was written by %d ranks" % (path, int(z["world"])))
                self.coverage = (z["positions"], z["bits"])
        if self.coverage is not None:
            return primer_dict, None, None
        if self.comm.world > 1:
            raise SystemExit("Error: a sharded pairing run needs the core step's coverage_bits files")
        with open(self.primer_file + ".gap_seq_id_json") as g:
            gap_dict = json.load(g)
        with open(self.primer_file + ".non_coverage_seq_id_json") as n:
            non_cover_dict = json.load(n)
        return primer_dict, gap_dict, non_cover_dict

    def get_number(self):
        """get_multiPrime.py:348-357: newline count / 2, capped by -m"""
        with open(self.Input_file, encoding="utf-8") as f:
            seq_number = int(f.read().count("\n") / 2)
        if seq_number > self.rep_seq_number != 0:
            return self.rep_seq_number
        return seq_number

    def pre_filter(self):
        lo, hi = (float(x) for x in self.GC.split(","))
        keep = []
        for pos, info in self.primers.items():
            sets = sets_of(info[0])
            if has_hairpin(sets, self.distance):
                continue
            gc = gc_mean(sets)
            if gc > hi or gc < lo:
                continue
            if has_repeat(sets):
                continue
            keep.append(pos)
        return sorted(keep)

    @staticmethod
    def closest(my_list, my_number1, my_number2):
        index_left = bisect_left(my_list, my_number1)
        if my_number2 > my_list[-1]:
            index_right = len(my_list) - 1
        else:
            index_right = bisect_left(my_list, my_number2) - 1
        return index_left, index_right

    # -- device-side tables ---------------------------------------------------------------------------------
    def _uncovered_bits(self, cand):
        """per candidate position: bit vectors of the ids left uncovered as forward / as reverse primer"""
        ids = {}
        for pos in cand:
            sp = str(pos)
            for dct in (self.gap_id[sp], self.non_cover_id[sp][0], self.non_cover_id[sp][1]):
                for lst in dct.values():
                    for x in lst:
                        if x not in ids:
                            ids[x] = len(ids)
        words = max(1, (len(ids) + 31) // 32)
        uf = np.zeros((len(cand), words), np.uint32)
        ur = np.zeros((len(cand), words), np.uint32)
        for r, pos in enumerate(cand):
            sp = str(pos)
            for arr, dcts in ((uf, (self.gap_id[sp], self.non_cover_id[sp][0])),
                              (ur, (self.gap_id[sp], self.non_cover_id[sp][1]))):
                for dct in dcts:
                    for lst in dct.values():
                        for x in lst:
                            j = ids[x]
                            arr[r, j >> 5] |= np.uint32(1 << (j & 31))
        return uf, ur

    def _pair_uncovered(self, cand, pairs):
        """sequences a (forward, reverse) pair leaves uncovered (get_multiPrime.py:556-569), for all pairs at once"""
        if not pairs:
            return []
        pf, pr = [p[0] for p in pairs], [p[1] for p in pairs]
        if self.coverage is None:                 # from the id lists of the JSON side files
            uf, ur = self._uncovered_bits(cand)
            return self.ctx.pair_cover(uf, ur, pf, pr)
        # from the scan's own bit vectors: popcount(F | gap | R' | gap') on the device; sequence shards add up
        positions, bits = self.coverage
        row_of = {int(p): i for i, p in enumerate(np.asarray(positions).tolist())}
        rows = np.array([row_of[int(p)] for p in cand], np.int32)
        unc = self.ctx.pair_cover3(bits, rows[pf], rows[pr]).astype(np.int64)
        return self.comm.allreduce_sum(unc) if self.comm.world > 1 else unc

    def run(self):
        min_len, max_len = (int(x) for x in self.size.split(","))
        cand = self.pre_filter_primers
        adaptor = self.adaptor.split(",")
        chief = self.comm.rank == 0
        say = print if chief else (lambda *a, **k: None)
        say("Candidata degenerate primer number is: {}".format(len(cand)))
        if int(cand[-1]) - int(cand[0]) < min_len:
            say("Max PCR product legnth < min len!")
            if chief:
                with open(self.outfile, "w") as fo:
                    fo.write(str(self.outfile) + "\n")
            return []
        n = len(cand)
        fwd = [self.primers[p][0] for p in cand]
        rev = [rc_string(s) for s in fwd]
        fsets = [sets_of(s) for s in fwd]
        rsets = [sets_of(s) for s in rev]
        ad_f, ad_r = sets_of(adaptor[0]), sets_of(adaptor[1])
        # per-primer filters of the forward use (get_multiPrime.py:510-519) and of the reverse use (:526-534)
        ok_f = [not (has_hairpin(ad_f + fsets[i], self.distance) or term_degenerate(fsets[i], self.position)
                     or gc_clamp(fsets[i])) for i in range(n)]
        ok_r = [not (has_hairpin(ad_r + rsets[i], self.distance) or term_degenerate(rsets[i], self.position)
                     or gc_clamp(rsets[i])) for i in range(n)]
        # candidate (start, stop) pairs in the reference's loop order
        pairs = []
        per_start = []
        for s in range(n):
            lst = []
            if ok_f[s]:
                a, b = self.closest(cand, cand[s] + min_len, cand[s] + max_len)
                for t in range(a, b + 1):
                    if not ok_r[t]:
                        continue
                    distance = int(cand[t]) - int(cand[s]) + 1
                    if distance > max_len:
                        lst.append((t, "break"))
                        break
                    if min_len <= distance <= max_len:
                        lst.append((t, len(pairs)))
                        pairs.append((s, t))
            per_start.append(lst)
        # F-R dimer check (get_multiPrime.py:419-437) of all pairs in one batch on the GPU
        dimer = np.zeros(len(pairs), bool)
        if pairs:
            eng = self._backend.Dimer(self.ctx, fsets + rsets, 5, 18, False, loss_table(3.6, True), dg_consts())
            try:
                idx = np.arange(2 * n, dtype=np.int32)
                self_hit = eng.pairs(idx, idx)[0] >= 0
                ps = np.array([p[0] for p in pairs], np.int32)
                pt = np.array([p[1] for p in pairs], np.int32) + n
                if self.comm.world > 1:        # the pair list is dealt round-robin to the ranks, the hit flags added up
                    mine = np.arange(self.comm.rank, len(pairs), self.comm.world)
                    part = np.zeros(len(pairs), np.int64)
                    part[mine] = (eng.pairs(ps[mine], pt[mine])[0] >= 0) | (eng.pairs(pt[mine], ps[mine])[0] >= 0)
                    cross = self.comm.allreduce_sum(part) > 0
                else:
                    cross = (eng.pairs(ps, pt)[0] >= 0) | (eng.pairs(pt, ps)[0] >= 0)
                dimer = cross | self_hit[ps] | self_hit[pt]
            finally:
                eng.close()
        uncovered = self._pair_uncovered(cand, pairs)
        out = []

        def one_pass(threshold, echo):
            for s in range(n):
                if echo:                       # get_multiPrime.py:620 prints the index in the first pass only
                    say(s)
                for t, q in per_start[s]:
                    if q == "break":
                        say("Error! PCR product greater than max length !")
                        break
                    if dimer[q]:
                        say("Dimer detection between Primer-F and Primer-R!")
                        continue
                    tm_f, tm_r = self.primers[cand[s]][4], self.primers[cand[t]][4]
                    if abs(tm_f - tm_r) > self.diff_Tm:
                        continue
                    non_cover = int(uncovered[q])
                    if non_cover / self.number > threshold:
                        continue
                    all_coverage = self.number - non_cover
                    line = (fwd[s], rev[t],
                            str(int(cand[t]) - int(cand[s]) + 1) + ":" + str(round(exact_mean([tm_f, tm_r]), 2)) + ":" +
                            str(round(all_coverage / self.number, 4)), all_coverage, str(cand[s]) + ":" + str(cand[t]))
                    out.append(line)

        coverage_threshold = 1 - self.fraction
        one_pass(coverage_threshold, True)
        if len(out) < 10:
            coverage_threshold += 0.1
            one_pass(coverage_threshold, False)
        if not chief:
            return out
        ID = str(self.outfile)
        primer_ID = str(self.outfile).split("/")[-1].rstrip(".txt")
        with open(self.outfile, "w") as fo, open(self.outfile.strip(".txt") + ".xls", "w") as fo_xls, \
                open(self.outfile.strip(".txt") + ".fa", "w") as fa:
            fo_xls.write("\t".join(["Primer_F_seq", "Primer_R_seq", "Product length:Tm:coverage_percentage",
                                    "Target number", "Primer_start_end"]) + "\n")
            fo.write(ID + "\t")
            for i in sorted(out, key=lambda r: r[3], reverse=True):
                fo.write("\t".join(map(str, i)) + "\t")
                fo_xls.write("\t".join(map(str, i)) + "\n")
                a, b = i[4].split(":")
                fa.write(">" + primer_ID + "_" + a + "F\n" + i[0] + "\n>" + primer_ID + "_" + b + "R\n" + i[1] + "\n")
            fo.write("\n")
        return out


def _shard_setup(args):
    """under torchrun: one rank per GPU (NCCL; MPB_DIST_BACKEND=gloo for ranks that share a GPU), every rank reads the
    coverage_bits shard the core step wrote for it (get_multiPrime.py:509-582 / SURVEY.md 8e: pair coverage is a
    popcount over sequences, so sequence shards add up; the F-R dimer tests are dealt to the ranks)"""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return {"device": args.device}, 0
    import torch
    import torch.distributed as dist
    from .comm import TorchComm
    rank, local = int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", "0"))
    backend = os.environ.get("MPB_DIST_BACKEND", "nccl")
    device = local if backend == "nccl" else args.device
    if backend == "nccl":
        torch.cuda.set_device(device)
        dist.init_process_group("nccl", device_id=torch.device("cuda", device))
    else:
        dist.init_process_group(backend)
    return {"device": device, "comm": TorchComm()}, rank


def main(argv=None, _backend=None):
    e1 = time.time()
    args = parseArg(argv)
    extra, rank = _shard_setup(args)
    app = Primers_filter(ref_file=args.ref, primer_file=args.input, adaptor=args.adaptor, rep_seq_number=args.maxseq,
                         distance=args.dist, outfile=args.out, size=args.size, position=args.end, fraction=args.fraction,
                         diff_Tm=args.Tm, nproc=args.proc, _backend=_backend, **extra)
    app.run()
    if "comm" in extra:
        import torch.distributed as dist
        dist.destroy_process_group()
    e2 = time.time()
    if rank == 0:
        print("INFO {} Total times: {}".format(time.strftime("%Y-%m-%d %H:%M:%S", time.localtime(time.time())),
                                               round(float(e2 - e1), 2)))


if __name__ == "__main__":
    main()
