"""Drop-in for scripts/get_degePrimer.py (get_degePrimer_V6.py): pair the candidates of a DegePrime table (the
multi-DegePrime workflow's counterpart of get_multiPrime).

Same flags and output file (one line: absolute path, then `F  R  length  min(matching)  start:stop` per pair, sorted by
the matching number, stable).  The per-primer filters (hairpin with adaptor, 3'-degenerate, GC clamp, GC range,
di-nucleotide repeats) are the ones of pairing.py, evaluated on base sets without expansion.  Two quirks of the reference
are kept: `-g` is parsed but the class default "0.4,0.6" filters (get_degePrimer_V6.py:553-557 never passes it), its
hairpin check only tests the first expansion of every 5-mer (a consumed generator, see has_hairpin), and its
pair dimer check can never fire — `current_end` builds its set with `end_seq.union(...)`, which discards the result
(get_degePrimer_V6.py:323-329), so every pair that reaches the check is kept.  Nothing here needs the GPU."""
from __future__ import annotations

import os
import sys
import time
from bisect import bisect_left
from optparse import OptionParser

from .core import has_repeat
from .iupac import ORDER, sets_of
from .pairing import gc_clamp, gc_mean, rc_string, term_degenerate


def has_hairpin(sets, distance: int) -> bool:
    """get_degePrimer_V6.py:296-316.  Unlike get_multiPrime's check, `degenerate_seq` is a GENERATOR here
    (get_degePrimer_V6.py:283-294) and the tail's expansions are consumed while the FIRST expansion of the 5-mer is
    tested: only that first expansion (first alternative of every degenerate position) is ever compared."""
    k = len(sets)
    for n in range(0, k - 5 - 5 - distance + 1):
        first = [ORDER[s][0] for s in sets[n:n + 5]]                  # base indices A,C,G,T = 0..3
        target = [1 << (3 - b) for b in reversed(first)]               # reverse complement, as base sets
        tail = sets[n + 5 + distance:]
        for o in range(0, len(tail) - 5 + 1):
            if all(target[t] & tail[o + t] for t in range(5)):
                return True
    return False


def argsParse(argv=None):
    parser = OptionParser('Usage: %prog -i [input] -r [sequence.fa] -o [output] \n \
                Options: {-f [0.6] -m [500] -n [200] -e [4] -p [9] -s [250,500] -g [0.4,0.6] -d [4] -a ","}.')
    parser.add_option('-i', '--input', dest='input', help='Input file: degeprimer out.')
    parser.add_option('-r', '--ref', dest='ref', help='Reference sequence file: all the sequence in 1 fasta.')
    parser.add_option('-g', '--gc', dest='gc', default="0.4,0.6", help="Filter primers by GC content. Default [0.4,0.6].")
    parser.add_option('-f', '--fraction', dest='fraction', default="0.6", type="float",
                      help="Filter primers by match fraction. Default: 0.6.")
    parser.add_option('-e', '--end', dest='end', default="4", type="int",
                      help="No degenerate base within the last N bases. Default: 4.")
    parser.add_option('-p', '--proc', dest='proc', default="10", type="int", help="accepted for compatibility")
    parser.add_option('-s', '--size', dest='size', default="250,500", help="Filter primers by PRODUCT size. Default [250,500].")
    parser.add_option('-d', '--dist', dest='dist', default=4, type="int",
                      help='Hairpin: distance of the minimal paired bases. Default: 4.')
    parser.add_option('-a', '--adaptor', dest='adaptor',
                      default="TCTTTCCCTACACGACGCTCTTCCGATCT,TCTTTCCCTACACGACGCTCTTCCGATCT", type="str",
                      help='Adaptor sequences F,R ("," for none).')
    parser.add_option('-m', '--maxseq', dest='maxseq', default=500, type="int", help='Limit of sequence number. Default: 500.')
    parser.add_option('-o', '--out', dest='out', help='Output file: candidate primers.')
    args = sys.argv[1:] if argv is None else argv
    (options, rest) = parser.parse_args(args)
    if len(args) == 0:
        parser.print_help()
        sys.exit(1)
    for value, msg in ((options.input, "Input file must be specified !!!"), (options.ref, "Reference file must be specified !!!"),
                       (options.out, "No output file provided !!!")):
        if value is None:
            parser.print_help()
            print(msg)
            sys.exit(1)
    return options


class Primers_filter(object):
    """get_degePrimer_V6.py:233-249 constructor arguments"""

    def __init__(self, ref_file, primer_file, adaptor, rep_seq_number=500, distance=4, outfile="", size="300,700",
                 position=9, GC="0.4,0.6", nproc=10, fraction=0.6):
        self.nproc = nproc
        self.primer_file = primer_file
        self.adaptor = adaptor
        self.size = size
        self.outfile = os.path.abspath(outfile)
        self.distance = distance
        self.Input_file = ref_file
        self.fraction = fraction
        self.GC = GC
        self.rep_seq_number = rep_seq_number
        self.number = self.get_number()
        self.position = position
        self.primers = self.parse_primers()
        self.pre_filter_primers = self.pre_filter()

    def get_number(self):
        """get_degePrimer_V6.py:268-278"""
        with open(self.Input_file, encoding="utf-8") as f:
            seq_number = int(f.read().count("\n") / 2)
        if seq_number > self.rep_seq_number != 0:
            print(seq_number, self.rep_seq_number)
            return self.rep_seq_number
        return seq_number

    def parse_primers(self):
        """get_degePrimer_V6.py:251-265: DegePrime table, primer in column 6, matching sequences in column 7"""
        primer_dict = {}
        with open(self.primer_file) as f:
            for line in f:
                if line.startswith("Pos"):
                    continue
                i = line.strip().split("\t")
                primer_dict[int(i[0])] = [i[5], round(int(i[6]) / self.number, 2), int(i[6])]
        return primer_dict

    def pre_filter(self):
        """get_degePrimer_V6.py:430-449"""
        lo, hi = (float(x) for x in self.GC.split(","))
        keep = []
        for pos, (primer, coverage, _) in self.primers.items():
            sets = sets_of(primer)
            if has_hairpin(sets, self.distance):
                continue
            gc = gc_mean(sets)
            if gc > hi or gc < lo:
                continue
            if has_repeat(sets):
                continue
            if coverage < self.fraction:
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

    def primer_pairs(self):
        """get_degePrimer_V6.py:460-508"""
        min_len, max_len = (int(x) for x in self.size.split(","))
        cand = self.pre_filter_primers
        adaptor = self.adaptor.split(",")
        out = []
        if not cand or int(cand[-1]) - int(cand[0]) < min_len:
            return out
        ad_f, ad_r = sets_of(adaptor[0]), sets_of(adaptor[1])
        fwd = [self.primers[p][0] for p in cand]
        rev = [rc_string(s) for s in fwd]
        fsets = [sets_of(s) for s in fwd]
        rsets = [sets_of(s) for s in rev]
        bad = lambda ad, sets: (has_hairpin(ad + sets, self.distance) or term_degenerate(sets, self.position)
                                or gc_clamp(sets))
        ok_r = {}
        for s in range(len(cand)):
            if bad(ad_f, fsets[s]):
                continue
            a, b = self.closest(cand, cand[s] + min_len, cand[s] + max_len)
            if a > b:
                break                      # the reference stops the whole search here
            for t in range(a, b + 1):
                if t not in ok_r:
                    ok_r[t] = not bad(ad_r, rsets[t])
                if not ok_r[t]:
                    continue
                distance = int(cand[t]) - int(cand[s]) + 1
                if distance > max_len:
                    break
                if min_len <= distance <= max_len:
                    # (the reference's dimer_check cannot fire, see the module docstring)
                    out.append((fwd[s], rev[t], distance, min(self.primers[cand[s]][2], self.primers[cand[t]][2]),
                                str(cand[s]) + ":" + str(cand[t])))
        return out

    def run(self):
        pairs = sorted(self.primer_pairs(), key=lambda k: k[3], reverse=True)
        with open(self.outfile, "w") as fo:
            fo.write(str(self.outfile) + "\t")
            for i in pairs:
                fo.write("\t".join(map(str, i)) + "\t")
            fo.write("\n")
        return pairs


def main(argv=None):
    e1 = time.time()
    options = argsParse(argv)
    app = Primers_filter(ref_file=options.ref, primer_file=options.input, adaptor=options.adaptor,
                         rep_seq_number=options.maxseq, distance=options.dist, outfile=options.out, size=options.size,
                         position=options.end, fraction=options.fraction, nproc=options.proc)
    app.run()
    e2 = time.time()
    print("INFO {} Total times: {}".format(time.strftime("%Y-%m-%d %H:%M:%S", time.localtime(time.time())),
                                           round(float(e2 - e1), 2)))


if __name__ == "__main__":
    main()
