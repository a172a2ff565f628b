"""Producer-side throughput: BLAST tables -> to_genomic + cut_coordinates -> columnar.

    python tools/bench_producer.py [--lnc 10000] [--steps 10] [--warmup 3] [--cpu-lnc 2000] [--tables 200]

Prints one JSON line:
  value      : input HSPs/s through o2a_cut_batch with inputs resident in HBM (CUDA events recorded by the
               library on its stream around the count pass, the scan and the write pass)
  roofline   : algorithmic bytes = 24 B read per input HSP (4 x int32 + float64) + 25 B written per kept HSP
               (4 x int32 + float64 + flag byte) over the kernel time, against MEASURED_PEAKS.json hbm_gbs
  e2e        : pinned host arrays in, host arrays out (H2D + kernels + D2H), wall clock
  cpu_baseline: the oracle port (C, OpenMP over alignments) on a bounded sample, same box
  blast_reader: native table reader (libo2a_ingest.so) MB/s and HSPs/s, 1 thread and all threads
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lnc", type=int, default=10000)
    ap.add_argument("--hits", type=int, default=2000)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--cpu-lnc", type=int, default=2000)
    ap.add_argument("--tables", type=int, default=200)
    args = ap.parse_args()
    import torch
    from ortho2align_b200 import producer, synth
    from ortho2align_b200.engine import Engine
    if not torch.cuda.is_available():
        raise SystemExit("no CUDA device: the producer path has no CPU fallback")
    dev = torch.device("cuda:0")
    t0 = time.perf_counter()
    rel, rg = synth.producer_workload(args.lnc, seed=1, hits_per_region=args.hits)
    gen_s = time.perf_counter() - t0
    n_hsp, n_aln = len(rel["score"]), len(rg["qleft"])
    eng = Engine(0)
    eng.set_timing(True)
    d_rel = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in rel.items()}
    d_rg = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in rg.items()}
    torch.cuda.synchronize()
    s = eng.cut_struct(d_rel, d_rg, device=True)
    for _ in range(args.warmup):
        kept = eng.run_cut(s)
    l0 = eng.launch_count()
    stage = {k: 0.0 for k in eng.CUT_STAGES}
    for _ in range(args.steps):
        eng.run_cut(s)
        for k, v in eng.cut_ms().items():
            stage[k] += v
    launches = eng.launch_count() - l0
    stage = {k: v / args.steps for k, v in stage.items()}
    kern_ms = stage["count"] + stage["scan"] + stage["write"]
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6548.8))
    alg_bytes = 24.0 * n_hsp + 25.0 * kept + 8.0 * 9 * n_aln
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
    # end to end: pinned host arrays -> host arrays
    def pin(a):
        a = np.ascontiguousarray(a)
        t = torch.empty(a.shape, dtype=torch.from_numpy(a[:0]).dtype, pin_memory=True)
        t.numpy()[...] = a
        return t
    keep_alive = {k: pin(v) for k, v in {**rel, **rg}.items()}
    h_rel = {k: keep_alive[k].numpy() for k in rel}
    h_rg = {k: keep_alive[k].numpy() for k in rg}
    hs = eng.cut_struct(h_rel, h_rg)
    eng.run_cut(hs); eng.fetch_cut(pinned=True)
    e2e_steps = max(2, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        eng.run_cut(hs)
        out = eng.fetch_cut(pinned=True)
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    h2d = sum(int(v.numpy().nbytes) for v in keep_alive.values())
    d2h = sum(int(v.nbytes) for v in out.values())
    # CPU port on a bounded sample of the same workload
    from oracle import oracle as O
    from oracle import producer as OP
    c_aln = min(n_aln, args.cpu_lnc * 5)
    c_h = int(rel["hsp_off"][c_aln])
    c_rel = {k: (v[:c_aln + 1] if k == "hsp_off" else v[:c_h]) for k, v in rel.items()}
    c_rg = {k: v[:c_aln] for k, v in rg.items()}
    threads = O.host_threads()
    best = None
    for _ in range(3):
        OP.cut_batch(c_rel, c_rg, n_threads=threads)
        best = OP.LAST_CALL_SECONDS if best is None else min(best, OP.LAST_CALL_SECONDS)
    # BLAST table reader
    rng = np.random.default_rng(4)
    tables = [synth.blast_tabular(*synth.relative_hits(rng, args.hits, 120_000, 2_100_000)) for _ in range(args.tables)]
    tb = sum(len(t) for t in tables)
    reader = {"tables": args.tables, "hits_per_table": args.hits, "mb": round(tb / 1e6, 1)}
    raw = [t.encode() for t in tables]
    for nt in (1, threads):
        bt = None
        for _ in range(3):
            t0 = time.perf_counter()
            r = producer.parse_blast_tables(raw, threads=nt)
            dt = time.perf_counter() - t0
            bt = dt if bt is None else min(bt, dt)
        assert not r["status"].any() and len(r["score"]) == args.tables * args.hits
        reader[f"native_{nt}_threads"] = {"hsps_per_s": len(r["score"]) / bt, "mb_per_s": tb / 1e6 / bt}
    t0 = time.perf_counter()
    for t in tables[:10]:
        OP.parse_blast(t)
    reader["python_restatement_1_thread"] = {"hsps_per_s": 10 * args.hits / (time.perf_counter() - t0)}
    line = {
        "metric": "input HSPs/sec through to_genomic + cut_coordinates (o2a_cut_batch)", "unit": "HSPs/s",
        "value": n_hsp / (kern_ms * 1e-3), "ms_per_step": kern_ms, "steps": args.steps, "warmup": args.warmup,
        "n_gpus": 1, "dtype": "int32 coordinates / f64 scores", "data": "synthetic (synth.producer_workload)",
        "config": {"workload": f"{args.lnc} genes x 5 syntenic regions x ~{args.hits} BLAST hits, gene-boundary cut",
                   "alignments": n_aln, "input_hsps": n_hsp, "kept_hsps": int(kept), "cut_hsps": int((out['cut'] & 1).sum()),
                   "l2": f"inputs {24 * n_hsp / 1e9:.2f} GB larger than L2; no flush needed", "gen_seconds": round(gen_s, 1)},
        "stage_ms": stage, "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "algorithmic_bytes": alg_bytes},
        "e2e": {"value": n_hsp / e2e_s, "unit": "HSPs/s", "ms_per_step": e2e_s * 1e3, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h},
        "cpu_baseline": {"value": c_h / best, "unit": "HSPs/s", "cores": threads, "kind": "port",
                         "sample": f"first {c_aln} alignments ({c_h} HSPs), best of 3"},
        "blast_reader": reader,
    }
    print(json.dumps(line))


if __name__ == "__main__":
    main()
