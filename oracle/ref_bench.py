"""Time the UNMODIFIED Python reference (dmitrymyl/ortho2align) on directories of its own JSON files:

    python oracle/ref_bench.py --align DIR --bg DIR --out DIR --cores N [--fitting hist] [--root baseline/_ref]

`build_orthologs` through the reference's own TimeoutProcessPool (`parallel.py:296-488`, `pipeline.py:846-860`) with
`cores` worker processes, then `get_best_orthologs`; prints one JSON line with the wall seconds of both. `pebble` is not
installable offline: the fork-pool stand-in of oracle/ref_harness.py provides the three calls the reference makes. That
pool cannot return "dropped" lncRNAs (see ref_harness._SerialPool), so the caller feeds inputs where every lncRNA has
an ortholog (BASELINE config 2 shape) and runs this script under a timeout.
TEST INFRASTRUCTURE: executed only by bench.py's `cpu_baseline_python` leg. The reference tree is taken from --root
(`baseline/_ref`, the offline pip install of /root/reference that travels to the GPU box) or /root/reference.
"""
import argparse
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--align", required=True)
    ap.add_argument("--bg", required=True)
    ap.add_argument("--out", required=True)
    ap.add_argument("--cores", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--fitting", default="hist")
    ap.add_argument("--root", default=None)
    args = ap.parse_args()
    root = args.root or (os.path.join(ROOT, "baseline", "_ref") if os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "ortho2align"))
                         else "/root/reference")
    os.environ["O2A_REFERENCE_ROOT"] = root
    sys.path.insert(0, ROOT)
    from oracle import ref_harness as rh
    if not rh.reference_available():
        print(json.dumps({"unavailable": f"no reference tree at {root}"}))
        return
    ref = rh.import_reference("fork" if args.cores > 1 else "serial")
    t0 = time.perf_counter()
    ref.pipeline.build_orthologs(args.align, args.bg, args.fitting, args.out, threshold=0.05, gapopen=5, gapextend=2, fdr=True,
                                 cores=args.cores, timeout=None, silent=True, stats_filename=os.path.join(args.out, "..", "ref.stats.txt"))
    t1 = time.perf_counter()
    j = lambda n: os.path.join(args.out, n)
    ref.pipeline.get_best_orthologs(j("significant.query_orthologs.bed"), j("significant.query_orthologs.tsv"),
                                    j("significant.subject_orthologs.bed"), j("significant.subject_orthologs.tsv"),
                                    "block_length", "max", j("best.q.bed"), j("best.q.tsv"), j("best.s.bed"), j("best.s.tsv"))
    t2 = time.perf_counter()
    n_orth = sum(1 for _ in open(j("significant.query_orthologs.bed")))
    print(json.dumps({"build_orthologs_s": t1 - t0, "get_best_orthologs_s": t2 - t1, "cores": args.cores, "orthologs": n_orth,
                      "reference_root": root, "pool": "fork stand-in for pebble.ProcessPool" if args.cores > 1 else "in-process"}))


if __name__ == "__main__":
    main()
