"""Command line of the drop-in multiPrime-core.py: the reference's flags and defaults (core:60-102), its output files
and its closing `INFO ... Total times` line."""
from __future__ import annotations

import argparse
import time


def parseArg(argv=None):
    parser = argparse.ArgumentParser(description="For degenerate primer design")
    parser.add_argument("-i", "--input", type=str, required=True,
                        help="Input file: multi-alignment output (muscle or others).", metavar="<file>")
    parser.add_argument("-l", "--plen", type=int, default=18, help="Length of primer. Default: 18.", metavar="<int>")
    parser.add_argument("-n", "--dnum", type=int, default=4, help="Max number of degenerate. Default: 4.",
                        metavar="<int>")
    parser.add_argument("-d", "--degeneracy", type=int, default=10, help="Max degeneracy of primer. Default: 10.",
                        metavar="<int>")
    parser.add_argument("-v", "--variation", type=int, default=1, help="Max mismatch number of primer. Default: 1",
                        metavar="<int>")
    parser.add_argument("-e", "--entropy", type=float, default=3.6,
                        help="Entropy threshold of a primer-length window; windows above it are skipped. Default: 3.6.",
                        metavar="<float>")
    parser.add_argument("-g", "--gc", type=str, default="0.2,0.7",
                        help="Filter primers by GC content. Default [0.2,0.7].", metavar="<str>")
    parser.add_argument("-s", "--size", type=int, default=100,
                        help="Filter primers by mini PRODUCT size. Default 100.", metavar="<int>")
    parser.add_argument("-f", "--fraction", type=float, default=0.8,
                        help="Filter primers by match fraction. Default: 0.8.", metavar="<float>")
    parser.add_argument("-c", "--coordinate", type=str, default="1,2,-1",
                        help="Primer positions where a mismatch disqualifies coverage-with-error "
                             "(>0: from the 5' end, <0: from the 3' end). Default: 1,2,-1.", metavar="<str>")
    parser.add_argument("-p", "--proc", type=int, default=20,
                        help="Number of process to launch (accepted for compatibility; the scan runs on the GPU). "
                             "Default: 20.", metavar="<int>")
    parser.add_argument("-a", "--away", type=int, default=4,
                        help="Filter hairpin structure: minimal distance between the paired bases. Default: 4.",
                        metavar="<int>")
    parser.add_argument("-o", "--out", type=str, required=True, help="output file", metavar="<file>")
    parser.add_argument("--device", type=int, default=0, help=argparse.SUPPRESS)
    parser.add_argument("--sidecars", default="auto", choices=["auto", "json", "bits"], help=argparse.SUPPRESS)
    return parser.parse_args(argv)


def _shard_setup(args):
    """Launched under torchrun (RANK / WORLD_SIZE set): every rank parses the alignment, keeps a contiguous block of
    its rows on its own GPU and joins the process group (NCCL; MPB_DIST_BACKEND=gloo for ranks that share one GPU).
    Returns the extra keyword arguments for NN_degenerate."""
    import os
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return {"device": args.device}, 0
    import torch
    import torch.distributed as dist
    from .comm import TorchComm
    from .core import parse_msa
    rank, local = int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", "0"))
    backend = os.environ.get("MPB_DIST_BACKEND", "nccl")
    device = local if backend == "nccl" else args.device
    if backend == "nccl":
        torch.cuda.set_device(device)
        dist.init_process_group("nccl", device_id=torch.device("cuda", device))
    else:
        dist.init_process_group(backend)
    ids, codes, lens = parse_msa(args.input)
    n = len(ids)
    if n < world:
        raise SystemExit("Error: %d sequences cannot be sharded over %d ranks; run a single process." % (n, world))
    lo, hi = rank * n // world, (rank + 1) * n // world
    extra = {"device": device, "alignment": (ids[lo:hi], codes[lo:hi], lens[lo:hi]), "row0": lo, "comm": TorchComm()}
    if backend == "nccl":            # the library's kernels and torch's collectives share one stream
        extra["stream"] = torch.cuda.current_stream().cuda_stream
    return extra, rank


def main(argv=None):
    e1 = time.time()
    args = parseArg(argv)
    from .core import NN_degenerate
    extra, rank = _shard_setup(args)
    app = NN_degenerate(seq_file=args.input, primer_length=args.plen, coverage=args.fraction,
                        number_of_dege_bases=args.dnum, score_of_dege_bases=args.degeneracy,
                        raw_entropy_threshold=args.entropy, product_len=args.size, position=args.coordinate,
                        variation=args.variation, distance=args.away, GC=args.gc, nproc=args.proc, outfile=args.out,
                        sidecar_format=args.sidecars, want_trace=False, **extra)
    app.run()
    app.close()
    if "comm" in extra:
        import torch.distributed as dist
        dist.destroy_process_group()
    e2 = time.time()
    if rank == 0:
        print("INFO {} Total times: {}".format(time.strftime("%Y-%m-%d %H:%M:%S", time.localtime(time.time())),
                                               round(float(e2 - e1), 2)))


if __name__ == "__main__":
    main()
