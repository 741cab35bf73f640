"""Drop-in for scripts/get_Maxprimerset.py (get_Maxprimerset_V1.3.py): greedy primer-set cover with dimer examination.

Same flags (-i -a -s -m -o) and output files: <out>, <out minus .xls>.next.xls (maximal mode) and sort.<input>
next to the input.  The walk over clusters is sequential as in the reference; what it asks at every step —
"does the candidate pair dimerise with itself or with anything accepted so far" (V1.3:193-215) — is answered by the
GPU dimer engine over (3' ends of a) x (expansions of b) for the pairs (a, b) that involve a new primer (pairs
of two accepted primers were already tested when the later one was accepted)."""
from __future__ import annotations

import optparse
import re
import sys
import time

from . import _lib
from .dimer import dg_consts, loss_table
from .iupac import sets_of

COLUMNS = ["#Primer", "Primer_rank", "Primer_F", "Primer_R", "PCR_product (Length:Tm:Coverage)",
           "Coverage number with error in top N", "Primer position (representative sequence)"]


def argsParse(argv=None):
    parser = optparse.OptionParser()
    parser.add_option('-i', '--input', dest='input', help='Input file: primers.')
    parser.add_option('-a', '--adaptor', dest='adaptor',
                      default="TCTTTCCCTACACGACGCTCTTCCGATCT,TCTTTCCCTACACGACGCTCTTCCGATCT", type="str",
                      help='Adaptor sequence, which is used for NGS next (accepted, unused as in the reference).')
    parser.add_option('-s', '--step', dest='step', default=5, type="int",
                      help='distance between primers; column number of primer1_F to primer2_F.')
    parser.add_option('-m', '--method', dest='method', default="T", type="str",
                      help='which method: maximal or maximum. If -m [T] use maximal; else maximum')
    parser.add_option('-o', '--out', dest='out', help='Prefix of out file: candidate primers')
    parser.add_option('--device', dest='device', default=0, type="int", help=optparse.SUPPRESS_HELP)
    args = sys.argv[1:] if argv is None else argv
    (options, rest) = parser.parse_args(args)
    if len(args) == 0:
        parser.print_help()
        sys.exit(1)
    elif options.input is None:
        parser.print_help()
        print("Input file must be specified !!!")
        sys.exit(1)
    elif options.out is None:
        parser.print_help()
        print("No output file provided !!!")
        sys.exit(1)
    return options


class DimerExaminer:
    """V1.3:193-215 dimer_examination over a fixed universe of primers, held on the device"""

    def __init__(self, ctx, primers, _backend=None):
        self.index = {}
        uniq = []
        for p in primers:
            if p not in self.index:
                self.index[p] = len(uniq)
                uniq.append(p)
        self.eng = (_backend or _lib).Dimer(ctx, [sets_of(p) for p in uniq], 5, -1, True, loss_table(3.0), dg_consts())
        self.cache = {}
        self.queries = 0

    def _hits(self, pairs):
        need = [pr for pr in pairs if pr not in self.cache]
        if need:
            hit, _ = self.eng.pairs([a for a, _ in need], [b for _, b in need])
            self.queries += len(need)
            for pr, h in zip(need, hit.tolist()):
                self.cache[pr] = h >= 0
        return any(self.cache[pr] for pr in pairs)

    def examine(self, primer_f, primer_r, accepted: list) -> bool:
        """True when the pair cannot join the accepted primers"""
        new = [self.index[primer_f], self.index[primer_r]]
        pairs = [(a, b) for a in new for b in new]
        for c in accepted:
            for a in new:
                pairs.append((a, c))
                pairs.append((c, a))
        return self._hits(list(dict.fromkeys(pairs)))

    def close(self):
        self.eng.close()


def _fmt_rows(rows):
    lines = ["\t".join(COLUMNS)]
    for r in rows:
        lines.append("\t".join("" if r.get(c) is None else str(r[c]) for c in COLUMNS))
    return "\n".join(lines) + "\n"


def greedy_maximal_primers(primers, step, exam: DimerExaminer, output, next_candidate):
    """V1.3:291-356: clusters in order; first compatible pair of each; a cluster without one is logged and skipped"""
    accepted = []
    rows = []
    for row in primers:
        if len(row) <= 1:
            print("Non primers: virus {} missing!".format(row[0]))
            next_candidate.write("\t".join(row) + "\n")
            continue
        column_pointer = 1
        while column_pointer <= len(row) - step:
            if exam.examine(row[column_pointer], row[column_pointer + 1], accepted):
                column_pointer += step
                if column_pointer > len(row) - step:
                    rows.append({"#Primer": row[0]})
                    print("virus {} missing!".format(row[0]))
                    next_candidate.write("\t".join(row) + "\n")
                    break
            else:
                rows.append(dict(zip(COLUMNS, [row[0], str(column_pointer)] + row[column_pointer:column_pointer + 5])))
                accepted.extend(dict.fromkeys([exam.index[row[column_pointer]], exam.index[row[column_pointer + 1]]]))
                break
    with open(output, "w") as fh:
        fh.write(_fmt_rows(rows))
    return rows


def greedy_primers(primers, step, exam: DimerExaminer, output):
    """V1.3:218-282, the 'maximum' variant with backtracking — including the reference's loop-variable behaviour:
    the outer `for` keeps handing out the next row index after a backtrack, so rows between the re-placed row and the
    row that triggered the backtrack are not revisited."""
    row_num = len(primers)
    accepted_before = {}        # row -> accepted list before that row was placed
    jdict = {}
    accepted = []
    rows = []
    blank_row = 0
    column_pointer = 1
    for it in range(row_num):
        row_pointer = it
        if len(primers[row_pointer]) <= 1:
            blank_row += 1
            continue
        while column_pointer <= len(primers[row_pointer]) - step:
            row = primers[row_pointer]
            if exam.examine(row[column_pointer], row[column_pointer + 1], accepted):
                column_pointer += step
                while column_pointer > len(primers[row_pointer]) - step:       # backtrack_to_previous_row
                    row_pointer -= 1
                    if row_pointer < blank_row:
                        print("Non maximum primer set. Try maximal primer set!")
                        sys.exit(1)
                    column_pointer = jdict[row_pointer] + step
                    accepted = accepted_before[row_pointer]
                    rows.pop()
            else:
                rows.append(dict(zip(COLUMNS, [row[0], str(column_pointer)] + row[column_pointer:column_pointer + 5])))
                accepted_before[row_pointer] = list(accepted)
                accepted = accepted + list(dict.fromkeys([exam.index[row[column_pointer]],
                                                          exam.index[row[column_pointer + 1]]]))
                jdict[row_pointer] = column_pointer
                column_pointer = 1
                break
    with open(output, "w") as fh:
        fh.write(_fmt_rows(rows))
    return rows


def main(argv=None, _backend=None):
    e1 = time.time()
    backend = _backend or _lib
    options = argsParse(argv)
    if re.search("/", options.input):
        sort_dir = options.input.split("/")
        sort = '/'.join(sort_dir[:-1]) + "/sort." + sort_dir[-1]
    else:
        sort = "sort." + options.input
    with open(options.input, "r") as primers_file, open(sort, "w") as f:
        primers = list(sorted([list(filter(None, line.strip().split('\t'))) for line in primers_file], key=len))
        for i in primers:
            f.write('\t'.join(i) + "\n")
    step = options.step
    universe = []
    for row in primers:
        col = 1
        while col <= len(row) - step:
            universe.extend(row[col:col + 2])
            col += step
    ctx = backend.Context(options.device)
    exam = DimerExaminer(ctx, universe, backend) if universe else None
    try:
        if options.method == "T":
            next_candidate = options.out.rstrip(".xls") + ".next.xls"
            with open(next_candidate, "w") as next_candidate_txt:
                greedy_maximal_primers(primers, step, exam, options.out, next_candidate_txt)
        else:
            greedy_primers(primers, step, exam, options.out)
    finally:
        if exam:
            exam.close()
        ctx.close()
    e2 = time.time()
    print("INFO {} Total times: {}".format(time.strftime("%Y-%m-%d %H:%M:%S", time.localtime(time.time())),
                                           round(float(e2 - e1), 2)))


if __name__ == "__main__":
    main()
