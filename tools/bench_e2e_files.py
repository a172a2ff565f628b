"""File-level end to end (SURVEY.md §8d "E2E" boundary): directories of JSON -> the 7 + 4 output files, through
the drop-in `build_orthologs` + `get_best_orthologs`.

    python tools/bench_e2e_files.py [--lnc 100] [--cores 0] [--runner gpu|oracle] [--reference]

Prints one JSON line with the wall time of each phase (ingest, compute, records + files, best) and HSPs/s for
a cold run (JSON parse) and a warm run (columnar sidecar). `--runner oracle` swaps the GPU call for the CPU
oracle (host-logic profiling where there is no GPU); `--reference` also times the unmodified reference with
`cores` processes where /root/reference exists (in-process pebble stand-in: the reference's own pool cannot
return its results, see oracle/ref_harness.py)."""
import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lnc", type=int, default=100)
    ap.add_argument("--cores", type=int, default=0)
    ap.add_argument("--runner", default="gpu", choices=["gpu", "oracle"])
    ap.add_argument("--reference", action="store_true")
    ap.add_argument("--ref-lnc", type=int, default=8)
    args = ap.parse_args()
    from ortho2align_b200 import pipeline, synth
    cores = args.cores or len(os.sched_getaffinity(0))
    if args.runner == "oracle":
        from oracle import oracle as O
        pipeline._run_shards = lambda batches, devices, fitter, fdr, alpha, go, ge: [O.build_batch(b, fitter, fdr, alpha, go, ge) for b in batches]
    b, m = synth.workload(2, n_lnc=args.lnc, with_meta=True, seed=1)
    items = synth.to_json_dicts(b, m)
    out = {"lncRNAs": args.lnc, "hsps": b.n_hsp, "cores": cores, "runner": args.runner}
    with tempfile.TemporaryDirectory(prefix="o2a_e2e_") as tmp:
        ad, bd = os.path.join(tmp, "align"), os.path.join(tmp, "bg")
        os.makedirs(ad), os.makedirs(bd)
        nbytes = 0
        for name, alns, bg in items:
            for d, obj in ((ad, alns), (bd, bg)):
                p = os.path.join(d, name + ".json")
                with open(p, "w") as f:
                    json.dump(obj, f)
                nbytes += os.path.getsize(p)
        out["json_mb"] = round(nbytes / 1e6, 1)
        del items

        def run(tag, **kw):
            od = os.path.join(tmp, "out_" + tag)
            pipeline.PHASES.clear()
            t0 = time.perf_counter()
            pipeline.build_orthologs(ad, bd, "hist", od, fdr=True, cores=cores, silent=True,
                                     stats_filename=os.path.join(tmp, tag + ".stats"), **kw)
            t1 = time.perf_counter()
            j = lambda n: os.path.join(od, n)
            pipeline.get_best_orthologs(j("significant.query_orthologs.bed"), j("significant.query_orthologs.tsv"),
                                        j("significant.subject_orthologs.bed"), j("significant.subject_orthologs.tsv"),
                                        "block_length", "max", j("best.q.bed"), j("best.q.tsv"), j("best.s.bed"), j("best.s.tsv"))
            t2 = time.perf_counter()
            n_orth = sum(1 for _ in open(j("significant.query_orthologs.bed")))
            out[tag] = {"build_orthologs_s": t1 - t0, "get_best_orthologs_s": t2 - t1, "hsps_per_s": b.n_hsp / (t2 - t0),
                        "orthologs": n_orth, "phases_s": dict(pipeline.PHASES)}
            return od

        if args.runner == "gpu":
            run("warmup")                                    # CUDA context, library load
        cache = os.path.join(tmp, "columnar.o2ac")
        run("cold_json", cache=cache)
        run("warm_sidecar", cache=cache)
        if args.reference:
            from oracle import ref_harness as rh
            if rh.reference_available():
                sub_a, sub_b = os.path.join(tmp, "ra"), os.path.join(tmp, "rb")
                os.makedirs(sub_a), os.makedirs(sub_b)
                names = sorted(os.listdir(ad))[:args.ref_lnc]
                for n in names:
                    os.symlink(os.path.join(ad, n), os.path.join(sub_a, n))
                    os.symlink(os.path.join(bd, n), os.path.join(sub_b, n))
                h = sum(len(a["HSPs"]) for n in names for a in json.load(open(os.path.join(ad, n))))
                t0 = time.perf_counter()
                rh.run_reference_build(sub_a, sub_b, os.path.join(tmp, "ref_out"), fitting="hist", fdr=True, cores=1, pool="serial")
                dt = time.perf_counter() - t0
                out["reference_1_core"] = {"lncRNAs": len(names), "hsps": h, "seconds": dt, "hsps_per_s": h / dt}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
