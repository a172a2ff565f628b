#!/usr/bin/env python
"""bench.py — HSPs/sec through build_orthologs + best-ortholog selection (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 2] [--lnc L]

A "step" = one pass of the hot path (background fit -> score filter -> p -> q -> chaining -> best
ortholog) over one batch of synthetic lncRNAs of the named BASELINE config (default configs[1]:
20k lncRNAs x 5 regions x ~2k HSPs, 10k background scores each, hist fitter).
  value : device-timed (CUDA events) with the columnar inputs already resident in HBM
  e2e   : the same call through the C-ABI with pinned HOST buffers: H2D of all inputs and D2H of
          the result (chains + per-lncRNA selection) inside the timed region
  roofline : the streaming kernel (k_filter), algorithmic bytes / CUDA-event duration vs measured peak
  cpu_baseline : the oracle port (C, OpenMP, all host threads) on a bounded sample, rank 0, N=1
`--impl reference` times that oracle port alone (the reference is pure Python and cannot travel to
the GPU box).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from ortho2align_b200 import synth  # noqa: E402
from ortho2align_b200.columnar import concat_batches  # noqa: E402

METRIC = "HSPs/sec through build_orthologs+best-ortholog"
LAYOUTS = {0: "float64 scores, int32 background", 8: "int8 scores + exception lists, int8 background (lossless)",
           16: "int16 scores + exception lists, int16 background (lossless)",
           32: "int32 scores + exception lists, int32 background (lossless)"}
UNIT = "HSPs/s"
CONFIG_NAMES = {1: "C1 strRNA 96 lncRNAs, 1k bg", 2: "C2 20k lncRNAs x 5 regions x ~2k HSPs, 10k bg/lncRNA",
                3: "C3 as C2 with 1M bg/lncRNA", 4: "C4 dense chaining 5k regions x 50k HSPs",
                5: "C5 100k lncRNAs, ~2e7 HSPs"}


def _gen_block(args):
    cfg, lo, hi, seed, total = args
    return synth.workload(cfg, n_lnc=total, seed=seed, lnc_range=(lo, hi))


def generate(cfg, n_lnc, seed, procs):
    """Synthetic batch of `n_lnc` lncRNAs of config `cfg`, generated block-parallel (fork pool)."""
    block = 64
    spans = [(cfg, lo, min(lo + block * 4, n_lnc), seed, n_lnc) for lo in range(0, n_lnc, block * 4)]
    if procs > 1 and len(spans) > 1:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(procs) as pool:
            parts = pool.map(_gen_block, spans)
    else:
        parts = [_gen_block(s) for s in spans]
    return concat_batches(parts)


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons of one GPU every 100 ms (NVML) while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40,
                 "sw_thermal_slowdown": 0x20, "hw_power_brake": 0x80, "sync_boost": 0x10,
                 "applications_clocks_setting": 0x2}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop_evt.wait(0.1)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


GRAPH_STATES = {0: "plain launches", 1: "captured in this step", 2: "CUDA graph replay (one launch per step)"}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel, config, n_lnc, fitter):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed `ncu --set full` capture
    of this workload (profiles/ncu_traffic.json, written by tools/ncu_summary.py from the .ncu-rep), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            table = json.load(f)
    except (OSError, ValueError):
        return None
    e = table.get(f"{kernel}|C{config}|{n_lnc}|{fitter}")
    return float(e["dram_bytes"]) if e else None


def python_reference_rate(config, fitter, cores, n_lnc, timeout_s=240):
    """The UNMODIFIED Python reference (baseline/_ref = the offline pip install of /root/reference, or /root/reference)
    on `n_lnc` lncRNAs of `config` written as its own JSON files: build_orthologs through its TimeoutProcessPool with
    `cores` processes + get_best_orthologs (oracle/ref_bench.py, run as a subprocess under a timeout). HSPs/s or a reason."""
    import subprocess
    import tempfile
    tmp_root = "/dev/shm" if os.path.isdir("/dev/shm") else None
    with tempfile.TemporaryDirectory(prefix="o2a_pyref_", dir=tmp_root) as tmp:
        ad, bd, od = os.path.join(tmp, "align"), os.path.join(tmp, "bg"), os.path.join(tmp, "out")
        os.makedirs(ad), os.makedirs(bd)
        step = 8
        spans = [(config, lo, min(lo + step, n_lnc), n_lnc, ad, bd) for lo in range(0, n_lnc, step)]
        import multiprocessing as mp
        with mp.get_context("fork").Pool(min(cores, 32, len(spans))) as pool:
            parts = pool.map(_write_json_block, spans)
        n_hsp = sum(p[0] for p in parts)
        try:
            res = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "ref_bench.py"), "--align", ad, "--bg", bd,
                                  "--out", od, "--cores", str(cores), "--fitting", fitter],
                                 capture_output=True, text=True, timeout=timeout_s)
        except subprocess.TimeoutExpired:
            return {"unavailable": f"the reference did not finish {n_lnc} lncRNAs within {timeout_s} s"}
        line = (res.stdout.strip().splitlines() or ["{}"])[-1]
        try:
            d = json.loads(line)
        except ValueError:
            return {"unavailable": f"reference run failed: {res.stderr.strip()[-300:]}"}
        if "unavailable" in d:
            return d
        secs = d["build_orthologs_s"] + d["get_best_orthologs_s"]
        return {"value": n_hsp / secs, "unit": UNIT, "cores": cores, "kind": "reference (unmodified Python package, pebble stand-in)",
                "sample": f"{n_lnc} lncRNAs ({n_hsp} HSPs) as JSON directories -> 7 + 4 files, {secs:.2f} s wall", **d}


def cpu_port_rate(batch, fitter, threads, repeats=1):
    """Oracle port (TEST INFRASTRUCTURE used as the reported CPU baseline) on `batch`: HSPs/s."""
    from oracle import oracle as O
    best = None
    for _ in range(repeats):
        O.build_batch(batch, fitter, True, n_threads=threads, compact=False)
        dt = O.LAST_CALL_SECONDS          # the C call only (no Python-side result packing)
        best = dt if best is None else min(best, dt)
    return batch.n_hsp / best, best


def run_reference_arm(args, rank):
    if rank != 0:
        return
    from oracle import oracle as O
    threads = O.host_threads()
    sample_lnc = args.ref_lnc
    batch = generate(args.config, sample_lnc, args.seed, min(threads, 16))
    for _ in range(args.warmup):
        O.build_batch(batch, args.fitter, True, n_threads=threads, compact=False)
    dt = 0.0
    for _ in range(args.steps):
        O.build_batch(batch, args.fitter, True, n_threads=threads, compact=False)
        dt += O.LAST_CALL_SECONDS
    dt /= args.steps
    v = batch.n_hsp / dt
    sample = f"{sample_lnc} lncRNAs ({batch.n_hsp} HSPs) of the workload per step"
    line = {"metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": CONFIG_NAMES[args.config], "fitter": args.fitter, "fdr": True,
                       "boundary": "columnar arrays in host RAM -> result arrays (CORE boundary, SURVEY.md §8d)"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# Secondary measurements (`--stage producer | ingest | files | annotate`): the rows SURVEY.md 8f adds around the hot path.
# They live here because their CPU-baseline legs run the oracle / the reference harness, which only tests/,
# smoke() and bench.py may touch. `tools/bench_*.py` are thin wrappers.
# --------------------------------------------------------------------------------------------------
def stage_producer(argv):
    ap = argparse.ArgumentParser(prog="bench.py --stage producer")
    ap.add_argument("--lnc", type=int, default=10000)
    ap.add_argument("--hits", type=int, default=2000)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--cpu-lnc", type=int, default=2000)
    ap.add_argument("--tables", type=int, default=200)
    args = ap.parse_args(argv)
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



def stage_annotate(argv):
    """annotate_orthologs' interval join: o2a_annotate_batch on synthetic orthologs x annotation."""
    ap = argparse.ArgumentParser(prog="bench.py --stage annotate")
    ap.add_argument("--orthologs", type=int, default=4_000_000)
    ap.add_argument("--annotation", type=int, default=4_000_000)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--cpu-orthologs", type=int, default=20000)
    args = ap.parse_args(argv)
    import torch
    from ortho2align_b200 import synth
    from ortho2align_b200.engine import Engine
    if not torch.cuda.is_available():
        raise SystemExit("no CUDA device: the annotate path has no CPU fallback")
    dev = torch.device("cuda:0")
    t0 = time.perf_counter()
    q, a = synth.annotate_workload(args.orthologs, args.annotation, seed=1, span=1_000_000_000)
    gen_s = time.perf_counter() - t0
    n_q, n_a = len(q["start"]), len(a["start"])
    eng = Engine(0)
    eng.set_timing(True)
    dq = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in q.items()}
    da = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in a.items()}
    torch.cuda.synchronize()
    s = eng.ann_struct(dq, da, 0, True, device=True)
    for _ in range(args.warmup):
        pairs = eng.run_annotate(s)
    l0 = eng.launch_count()
    stage = {k: 0.0 for k in eng.ANN_STAGES}
    for _ in range(args.steps):
        eng.run_annotate(s)
        for k, v in eng.annotate_ms().items():
            stage[k] += v
    launches = eng.launch_count() - l0
    stage = {k: v / args.steps for k, v in stage.items()}
    kern_ms = stage["prefmax"] + stage["count"] + stage["write"]
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6548.8))
    # algorithmic bytes: every ortholog (17 B) and annotation range (17 B) read once, 8 B offset + 20 B per pair written
    alg_bytes = 17.0 * n_q + 17.0 * n_a + 8.0 * n_q + 20.0 * pairs
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9

    def pin(v):
        v = np.ascontiguousarray(v)
        t = torch.empty(v.shape, dtype=torch.from_numpy(v[:0]).dtype, pin_memory=True)
        t.numpy()[...] = v
        return t
    keep_q, keep_a = {k: pin(v) for k, v in q.items()}, {k: pin(v) for k, v in a.items()}
    hs = eng.ann_struct({k: v.numpy() for k, v in keep_q.items()}, {k: v.numpy() for k, v in keep_a.items()}, 0, True)
    eng.run_annotate(hs); eng.fetch_annotation(pinned=True)
    e2e_steps = max(2, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        eng.run_annotate(hs)
        out = eng.fetch_annotation(pinned=True)
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    h2d = sum(int(v.numpy().nbytes) for v in list(keep_q.values()) + list(keep_a.values()))
    d2h = sum(int(v.nbytes) for v in out.values())
    # CPU port (the reference's loops, literally: O(chromosome) backward scan per ortholog) on a bounded sample
    from oracle import annotate as OA
    from oracle import oracle as O
    threads = O.host_threads()
    c_q = min(n_q, args.cpu_orthologs)
    sel = np.linspace(0, n_q - 1, c_q).astype(np.int64) if c_q else np.zeros(0, np.int64)
    cq = {k: np.ascontiguousarray(v[sel]) for k, v in q.items()}
    best = None
    for _ in range(2):
        exp = OA.annotate_batch(cq, a, 0, True, n_threads=threads)
        best = OA.LAST_CALL_SECONDS if best is None else min(best, OA.LAST_CALL_SECONDS)
    # the sample doubles as a parity check of the timed call's result
    off = out["nb_off"]
    for j, i in enumerate(sel[:2000].tolist()):
        assert np.array_equal(out["nb_idx"][off[i]:off[i + 1]], exp["nb_idx"][exp["nb_off"][j]:exp["nb_off"][j + 1]]), i
    line = {
        "metric": "orthologs/sec through find_neighbours + calc_JI + calc_OC (o2a_annotate_batch)", "unit": "orthologs/s",
        "value": n_q / (kern_ms * 1e-3), "ms_per_step": kern_ms, "steps": args.steps, "warmup": args.warmup,
        "n_gpus": 1, "dtype": "int32 coordinates / f64 JI, OC", "data": "synthetic (synth.annotate_workload)",
        "config": {"workload": f"{n_q} orthologs x {n_a} annotation ranges on 24 chromosomes of 1 Gb, distance 0, strandness on",
                   "pairs": int(pairs), "l2": "flushed by the inputs of the other side between passes: "
                   f"{17 * (n_q + n_a) / 1e6:.0f} MB of columns + {20 * pairs / 1e6:.0f} MB of pairs per step",
                   "gen_seconds": round(gen_s, 1)},
        "stage_ms": stage, "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "algorithmic_bytes": alg_bytes,
                     "note": "binary searches dominate: latency-bound, not a streaming kernel"},
        "e2e": {"value": n_q / e2e_s, "unit": "orthologs/s", "ms_per_step": e2e_s * 1e3, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h},
        "cpu_baseline": {"value": c_q / best, "unit": "orthologs/s", "cores": threads, "kind": "port",
                         "sample": f"{c_q} evenly spaced orthologs against the full annotation, best of 2"},
    }
    print(json.dumps(line))


def stage_ingest(argv):
    ap = argparse.ArgumentParser(prog="bench.py --stage ingest")
    ap.add_argument("--lnc", type=int, default=48)
    ap.add_argument("--threads", type=int, default=0)
    args = ap.parse_args(argv)
    import tempfile
    from ortho2align_b200 import ingest
    from ortho2align_b200.pipeline import _load_lncrna
    b, m = synth.workload(2, n_lnc=args.lnc, with_meta=True, seed=1)
    items = synth.to_json_dicts(b, m)
    with tempfile.TemporaryDirectory(prefix="o2a_ingest_") as tmp:
        ad, bd = os.path.join(tmp, "align"), os.path.join(tmp, "bg")
        os.makedirs(ad), os.makedirs(bd)
        af, bf, nbytes = [], [], 0
        for name, alns, bg in items:
            pa, pb = os.path.join(ad, name + ".json"), os.path.join(bd, name + ".json")
            with open(pa, "w") as f:
                json.dump(alns, f)
            with open(pb, "w") as f:
                json.dump(bg, f)
            nbytes += os.path.getsize(pa) + os.path.getsize(pb)
            af.append(pa); bf.append(pb)
        out = {"lncRNAs": args.lnc, "hsps": b.n_hsp, "json_mb": round(nbytes / 1e6, 1)}
        t0 = time.perf_counter()
        for a, c in zip(af, bf):
            _load_lncrna(a, c)
        dt = time.perf_counter() - t0
        out["python_json_pack"] = {"hsps_per_s": b.n_hsp / dt, "mb_per_s": nbytes / 1e6 / dt, "threads": 1}
        nthreads = args.threads or len(os.sched_getaffinity(0))
        for nt in (1, nthreads):
            best = None
            for _ in range(3):
                t0 = time.perf_counter()
                r = ingest.load_files(af, bf, threads=nt)
                dt = time.perf_counter() - t0
                best = dt if best is None else min(best, dt)
                assert r.batch.n_hsp == b.n_hsp
                r.close()
            out[f"native_{nt}_threads"] = {"hsps_per_s": b.n_hsp / best, "mb_per_s": nbytes / 1e6 / best, "threads": nt}
        # columnar sidecar: written once, then memory-mapped and copied into fresh arrays (what a run touches)
        r = ingest.load_files(af, bf, threads=nthreads)
        side = os.path.join(tmp, "columnar.o2ac")
        t0 = time.perf_counter()
        ingest.save_sidecar(side, r, af, bf)
        out["sidecar_write_s"] = time.perf_counter() - t0
        out["sidecar_mb"] = round(os.path.getsize(side) / 1e6, 1)
        r.close()
        best = None
        for _ in range(3):
            t0 = time.perf_counter()
            s = ingest.open_sidecar(side, af, bf)
            total = sum(int(getattr(s.batch, k).sum(dtype="int64")) for k in ("bg", "qstart", "qend", "sstart", "send"))
            total += float(s.batch.score.sum())
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        out["sidecar_read"] = {"hsps_per_s": b.n_hsp / best, "mb_per_s": os.path.getsize(side) / 1e6 / best, "threads": 1,
                               "note": "open + manifest check + one full pass over every mapped array"}
        try:
            sys.path.insert(0, ROOT)
            from oracle import ref_harness as rh
            if rh.reference_available():
                ref = rh.import_reference()
                t0 = time.perf_counter()
                for a in af[:8]:
                    with open(a) as f:
                        [ref.genomicranges.GenomicRangesAlignment.from_dict(d) for d in json.load(f)]
                dt = time.perf_counter() - t0
                frac = sum(len(al["HSPs"]) for it in items[:8] for al in it[1])
                out["reference_json_from_dict"] = {"hsps_per_s": frac / dt, "threads": 1, "sample_lncRNAs": 8}
        except Exception as e:  # pragma: no cover
            out["reference_json_from_dict"] = {"error": str(e)}
        print(json.dumps(out))



def stage_files(argv):
    ap = argparse.ArgumentParser(prog="bench.py --stage files")
    ap.add_argument("--lnc", type=int, default=100)
    ap.add_argument("--cores", type=int, default=0)
    ap.add_argument("--runner", default="gpu", choices=["gpu", "oracle"])
    ap.add_argument("--reference", action="store_true")
    ap.add_argument("--ref-lnc", type=int, default=8)
    ap.add_argument("--config", type=int, default=2, help="shape of the synthetic lncRNAs (2: 5 regions x ~2k HSPs; 5: genome-scale mix)")
    ap.add_argument("--tmp", default=None, help="where the JSON directories are written (default: /dev/shm if present)")
    args = ap.parse_args(argv)
    print(json.dumps(files_to_files(args.config, args.lnc, args.cores, args.runner, args.reference, args.ref_lnc, args.tmp)))


def _write_json_block(a):
    cfg, lo, hi, total, ad, bd = a
    b, m = synth.workload(cfg, n_lnc=total, with_meta=True, seed=1, lnc_range=(lo, hi))
    nbytes = 0
    for name, alns, bg in synth.to_json_dicts(b, m):
        for d, obj in ((ad, alns), (bd, bg)):
            p = os.path.join(d, name + ".json")
            with open(p, "w") as f:
                json.dump(obj, f)
            nbytes += os.path.getsize(p)
    return b.n_hsp, nbytes


def files_to_files(config, n_lnc, cores=0, runner="gpu", reference=False, ref_lnc=8, tmp_root=None):
    """Directories of the reference's JSON files -> the 7 + 4 output files through the drop-in build_orthologs +
    get_best_orthologs (cold: native JSON parse; warm: columnar sidecar). Returns the report dict."""
    import tempfile
    from ortho2align_b200 import pipeline
    cores = cores or len(os.sched_getaffinity(0))
    args = argparse.Namespace(lnc=n_lnc, runner=runner, reference=reference, ref_lnc=ref_lnc)
    if tmp_root is None and os.path.isdir("/dev/shm"):
        tmp_root = "/dev/shm"
    if args.runner == "oracle":
        from oracle import oracle as O
        pipeline._run_shards = lambda batches, devices, fitter, fdr, alpha, go, ge: [
            O.build_batch(b.to_columnar() if hasattr(b, "to_columnar") else b, fitter, fdr, alpha, go, ge) for b in batches]
    out = {"workload": CONFIG_NAMES[config], "lncRNAs": args.lnc, "cores": cores, "runner": args.runner}
    with tempfile.TemporaryDirectory(prefix="o2a_e2e_", dir=tmp_root) as tmp:
        ad, bd = os.path.join(tmp, "align"), os.path.join(tmp, "bg")
        os.makedirs(ad), os.makedirs(bd)
        t0 = time.perf_counter()
        step = 64 if config != 1 else 1
        spans = [(config, lo, min(lo + step, args.lnc), args.lnc, ad, bd) for lo in range(0, args.lnc, step)]
        if cores > 1 and len(spans) > 1:
            import multiprocessing as mp
            with mp.get_context("fork").Pool(min(cores, 32)) as pool:
                parts = pool.map(_write_json_block, spans)
        else:
            parts = [_write_json_block(sp) for sp in spans]
        n_hsp = sum(p[0] for p in parts)
        out["hsps"] = n_hsp
        out["json_mb"] = round(sum(p[1] for p in parts) / 1e6, 1)
        out["write_json_s"] = round(time.perf_counter() - t0, 2)
        b = argparse.Namespace(n_hsp=n_hsp)

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
            run("warmup")                                    # CUDA context, library load, workspace
        run("cold_json")                                     # JSON parse -> files, nothing cached
        cache = os.path.join(tmp, "columnar.o2ac")
        run("cold_json_writing_sidecar", cache=cache)        # the same plus writing the columnar sidecar
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
    return out



def strong_scaling(args, eng, params, dev, rank, world, procs, config, n_lnc, steps, warmup):
    """ONE fixed workload (BASELINE configs[4] by default: 100 k lncRNAs, ~2e7 HSPs) LPT-sharded by lncRNA over the
    ranks (ortho2align_b200.distributed.my_shard = pipeline.lpt_shards by HSP count); every rank runs its shard, the
    per-lncRNA results are gathered on every rank (distributed.gather_best_orthologs: counts + one padded all_gather)
    inside the timed region, and rank 0 checks them against the whole workload run on one GPU. Returns the dict that
    goes into the JSON line (rank 0) or None."""
    import torch
    import torch.distributed as dist
    from ortho2align_b200 import distributed as D
    from ortho2align_b200.engine import DeviceBatch
    with_kde = args.fitter == "kde"
    t0 = time.perf_counter()
    n_lnc = n_lnc or {1: 96, 2: 20_000, 3: 20_000, 4: 5_000, 5: 100_000}[config]
    batch = generate(config, n_lnc, args.seed, procs)                       # the same workload on every rank
    t_gen = time.perf_counter() - t0
    idx, shard = D.my_shard(batch, rank, world)
    db = DeviceBatch(shard, dev, aos=True)
    st = db.struct(with_kde)
    for _ in range(max(warmup, 1)):
        eng.run(st, params)
    res = eng.fetch(survivors=False)
    if world > 1:
        D.gather_best_orthologs(res, idx, batch.n_lnc, dev)                 # warm-up of the collective
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    stream = torch.cuda.current_stream()
    eng.set_timing(False)                 # no stage events: the shard's launch sequence is replayed as one CUDA graph
    for _ in range(3):
        eng.run(st, params)               # eager -> capture -> replay
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = eng.launch_count()
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(steps):
        c = eng.run(st, params)
    e1.record(stream)
    torch.cuda.synchronize()
    graph_state = eng.graph_state()
    eng.set_timing(True)
    dev_ms = e0.elapsed_time(e1) / steps
    # the job's result exchange: every rank learns which lncRNAs have an ortholog (what get_best_orthologs consumes)
    res = eng.fetch(survivors=False)
    has = D.gather_best_orthologs(res, idx, batch.n_lnc, dev) if world > 1 else (res.best_aln >= 0).astype(np.int64)
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - t0) * 1e3 / steps
    launches = (eng.launch_count() - l0) // steps
    t = torch.tensor([dev_ms, wall_ms], dtype=torch.float64, device=dev)
    tot = torch.tensor([c.n_survivors, c.n_retained, c.n_chain_hsps, c.n_chains, shard.n_hsp], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    out = None
    if rank == 0:
        # the whole workload on this GPU: the sharded job must reproduce it
        full = DeviceBatch(batch, dev, aos=True)
        cf = eng.run(full.struct(with_kde), params)
        rf = eng.fetch(survivors=False)
        same = (int(tot[0]) == cf.n_survivors and int(tot[1]) == cf.n_retained and int(tot[2]) == cf.n_chain_hsps
                and int(tot[3]) == cf.n_chains and np.array_equal(has, (rf.best_aln >= 0).astype(np.int64)))
        out = {"scaling": "strong", "workload": CONFIG_NAMES[config], "lncRNAs": batch.n_lnc, "hsps": batch.n_hsp,
               "n_gpus": world, "steps": steps, "ms_per_step": float(t[0]), "value": batch.n_hsp / (float(t[0]) * 1e-3),
               "unit": UNIT, "wall_ms_per_step_incl_result_gather": float(t[1]),
               "value_incl_result_gather": batch.n_hsp / (float(t[1]) * 1e-3),
               "launches_per_step": int(launches), "graph": GRAPH_STATES[graph_state], "hsps_on_rank0": shard.n_hsp,
               "identical_to_single_gpu": bool(same), "shard": "LPT by HSP count (pipeline.lpt_shards)",
               "gen_seconds": round(t_gen, 1)}
        assert same, "sharded result differs from the single-GPU result"
        del full
    del db
    torch.cuda.empty_cache()
    return out


def pcie_ceiling(dev, world, mb=256, reps=10):
    """Plain cudaMemcpyAsync of a pinned buffer on every rank AT THE SAME TIME: {"h2d"|"d2h": {"per_rank_gbs", "aggregate_gbs"}}.
    Collective over the ranks (barrier + all_gather). This is the box's host<->device ceiling the e2e arm runs against."""
    import torch
    import torch.distributed as dist
    n = mb << 20
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True); h.fill_(1)
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    out = {}
    for name, (src, dst) in (("h2d", (h, d)), ("d2h", (d, h))):
        for _ in range(3):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        gbs = torch.tensor([n * reps / dt / 1e9], dtype=torch.float64, device=dev)
        allg = torch.zeros(world, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_gather_into_tensor(allg, gbs)
        else:
            allg = gbs
        out[name] = {"per_rank_gbs": [round(float(x), 2) for x in allg.tolist()], "aggregate_gbs": round(float(allg.sum()), 2)}
    del h, d
    return out


def stage_pcie(argv):
    """Host->device ceiling of the box: every rank copies a pinned buffer to its GPU at the same time (plain
    cudaMemcpyAsync); prints per-rank and aggregate GB/s. Run under torchrun with N = 1, 2, 4, 8."""
    ap = argparse.ArgumentParser(prog="bench.py --stage pcie")
    ap.add_argument("--mb", type=int, default=512)
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args(argv)
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0")); local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    out = pcie_ceiling(dev, world, args.mb, args.reps)
    if rank == 0:
        print(json.dumps({"stage": "pcie", "n_gpus": world, "mb": args.mb, "reps": args.reps, **out}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    if "--stage" in sys.argv:
        k = sys.argv.index("--stage")
        stage = sys.argv[k + 1] if k + 1 < len(sys.argv) else ""
        rest = sys.argv[1:k] + sys.argv[k + 2:]
        fn = {"producer": stage_producer, "ingest": stage_ingest, "files": stage_files, "annotate": stage_annotate,
              "pcie": stage_pcie}.get(stage)
        if stage != "build":
            if fn is None:
                raise SystemExit("--stage must be one of build, producer, ingest, files, annotate, pcie")
            return fn(rest)
        sys.argv = [sys.argv[0]] + rest
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--lnc", type=int, default=None, help="lncRNAs per GPU (default: the config's full size)")
    ap.add_argument("--fitter", default="hist", choices=["hist", "kde", "empirical"])
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--cpu-lnc", type=int, default=20000, help="lncRNAs in the cpu_baseline sample")
    ap.add_argument("--ref-lnc", type=int, default=20000, help="lncRNAs per step of --impl reference")
    ap.add_argument("--e2e-steps", type=int, default=None)
    ap.add_argument("--e2e-engines", type=int, default=2, choices=[1, 2],
                    help="library contexts per GPU in the e2e arm: with 2, two host threads alternate batches, so the "
                         "uploads of one batch overlap the last kernels and the result download of the other")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default): every rank owns a full shard of --config; strong: ONE fixed workload "
                         "(--strong-config, default BASELINE configs[4]) LPT-sharded over the ranks, result checked against 1 GPU")
    ap.add_argument("--strong-config", type=int, default=5)
    ap.add_argument("--strong-lnc", type=int, default=None)
    ap.add_argument("--no-strong", action="store_true", help="N > 1 default run: skip the attached strong-scaling measurement")
    ap.add_argument("--batches", type=int, default=1,
                    help="device-resident arm: the workload as this many batches of --lnc lncRNAs each, all resident in HBM, "
                         "one step = one pass over all of them (config 3 at full size: --lnc 1000 --batches 20)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-python-reference", action="store_true", help="skip timing the unmodified Python reference (baseline/_ref)")
    ap.add_argument("--pyref-lnc", type=int, default=768, help="lncRNAs in the Python reference's sample (768 ~ 5 s on 16 cores)")
    ap.add_argument("--no-files", action="store_true", help="skip the directories -> files figure (N = 1 default run)")
    ap.add_argument("--files-lnc", type=int, default=100_000, help="lncRNAs of the directories -> files figure (config 5 shape)")
    ap.add_argument("--no-e2e", action="store_true", help="device-resident arm only (profiling runs under ncu)")
    ap.add_argument("--gather", default="final", choices=["step", "final"],
                    help="N > 1: all_gather of the result counters after every step (asynchronous), or once for all "
                         "steps before the timed region closes")
    ap.add_argument("--score-bits", type=int, default=0, choices=[0, 8, 16, 32],
                    help="score layout of the device-resident arm: 0 = float64 (the reference's dtype), "
                         "8/16/32 = lossless narrow ints + exception lists")
    ap.add_argument("--e2e-score-bits", type=int, default=0, choices=[0, 8, 16, 32],
                    help="force the score width of the e2e arm's host layout (0 = what the ingest's planner picks: the "
                         "narrowest lossless width that sends at most 10 %% of the scores to the exception lists)")
    ap.add_argument("--e2e-coords", default="zerocopy,gather",
                    help="comma list of coordinate modes to measure in the e2e arm (zerocopy, gather, upload); the "
                         "fastest is the one the headline e2e uses")
    ap.add_argument("--no-e2e-f64", dest="e2e_f64", action="store_false",
                    help="skip the second e2e figure (float64 / int32 host buffers)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    full = {1: 96, 2: 20_000, 3: 20_000, 4: 5_000, 5: 100_000}[args.config]
    n_lnc = args.lnc or full
    ncpu = os.cpu_count() or 8
    procs = max(1, min(32, ncpu // max(world, 1)))
    t_gen = time.perf_counter()
    # weak scaling: every rank owns its own full-size shard of lncRNAs (independent units, no exchange)
    batch = generate(args.config, n_lnc, args.seed + 1000 * rank, procs) if args.scaling == "weak" else None
    t_gen = time.perf_counter() - t_gen

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from ortho2align_b200.engine import DeviceBatch, Engine

    eng = Engine(local_rank)
    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)
    eng.set_timing(True)
    params = eng.params(args.fitter, True, 0.05, 5, 2, "block_length", "max")
    with_kde = args.fitter == "kde"

    if args.scaling == "strong":
        sampler = ClockSampler(local_rank)
        sampler.start()
        ss = strong_scaling(args, eng, params, dev, rank, world, procs, args.strong_config, args.strong_lnc,
                            args.steps, args.warmup)
        clocks = sampler.stop()
        if rank == 0:
            line = {"metric": METRIC, "value": ss["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
                    "warmup": args.warmup, "ms_per_step": ss["ms_per_step"], "higher_is_better": True, "scaling": "strong",
                    "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                    "config": {"workload": ss["workload"], "lncRNAs": ss["lncRNAs"], "hsps": ss["hsps"], "fitter": args.fitter,
                               "fdr": True, "alpha": 0.05, "parallelism": f"one workload LPT-sharded by lncRNA over {world} GPU(s)",
                               "l2": "every step re-reads its shard from HBM; C5 per-GPU shards below L2 size at N >= 4 are noted in DESIGN.md"},
                    "strong_scaling": ss, "gpu_launches": ss["launches_per_step"] * args.steps, "clocks": clocks,
                    "e2e": None, "roofline": None, "cpu_baseline": None}
            print(json.dumps(line), flush=True)
        if world > 1:
            dist.destroy_process_group()
        eng.close()
        return

    # ---------------- device-resident arm (value) ----------------
    bg_bits = args.score_bits if args.score_bits in (8, 16) else 32
    dbatch = DeviceBatch(batch, dev, aos=True, score_bits=args.score_bits, bg_bits=bg_bits)
    dstruct = dbatch.struct(with_kde)
    # --batches B: B - 1 more batches of the same shape (other seeds), generated one at a time and kept in HBM only
    more, more_hsp, more_bg, more_exc = [], 0, 0, 0
    for bi in range(1, args.batches):
        t0 = time.perf_counter()
        hb_i = generate(args.config, n_lnc, args.seed + 1000 * rank + 7919 * bi, procs)
        t_gen += time.perf_counter() - t0
        db_i = DeviceBatch(hb_i, dev, aos=True, score_bits=args.score_bits, bg_bits=bg_bits)
        more.append((db_i, db_i.struct(with_kde)))
        more_hsp += hb_i.n_hsp; more_bg += hb_i.n_bg
        more_exc += int(db_i.t["exc_pos"].numel()) if args.score_bits else 0
        del hb_i
    # the only collective of the path: every rank learns every shard's result counts (NCCL all_gather, 40 B per
    # rank and step). It is issued asynchronously from its own ring slot, so the next step's kernels do not wait
    # for the other ranks; all of them are waited for before the timed region closes.
    n_slots = args.steps + max(args.warmup, 1) + 1
    counts_host = torch.zeros((n_slots, 5), dtype=torch.int64).pin_memory()
    counts_dev = torch.zeros((n_slots, 5), dtype=torch.int64, device=dev)
    gathered = torch.zeros((n_slots, 5 * world), dtype=torch.int64, device=dev)
    gathered_all = torch.zeros((world, n_slots, 5), dtype=torch.int64, device=dev)
    pending = []

    steps_done = [0]

    class _Sum:
        n_survivors = n_retained = n_chain_hsps = n_chains = n_failed_lnc = 0

    def step_device(acc=None):
        c = eng.run(dstruct, params)
        if more:
            tot = _Sum()
            for cc in [c]:
                for f in ("n_survivors", "n_retained", "n_chain_hsps", "n_chains", "n_failed_lnc"):
                    setattr(tot, f, getattr(tot, f) + int(getattr(cc, f)))
            if acc is not None:
                for k, v in eng.stage_ms().items():
                    acc[k] = acc.get(k, 0.0) + v
            for _, st_i in more:
                cc = eng.run(st_i, params)
                for f in ("n_survivors", "n_retained", "n_chain_hsps", "n_chains", "n_failed_lnc"):
                    setattr(tot, f, getattr(tot, f) + int(getattr(cc, f)))
                if acc is not None:
                    for k, v in eng.stage_ms().items():
                        acc[k] = acc.get(k, 0.0) + v
            c = tot
        elif acc is not None:
            for k, v in eng.stage_ms().items():
                acc[k] = acc.get(k, 0.0) + v
        if world > 1 and args.gather == "final":
            k = steps_done[0] % n_slots
            steps_done[0] += 1
            counts_host[k].numpy()[:] = (c.n_survivors, c.n_retained, c.n_chain_hsps, c.n_chains, c.n_failed_lnc)
        elif world > 1:
            k = len(pending) % n_slots
            counts_host[k].numpy()[:] = (c.n_survivors, c.n_retained, c.n_chain_hsps, c.n_chains, c.n_failed_lnc)
            counts_dev[k].copy_(counts_host[k], non_blocking=True)
            pending.append(dist.all_gather_into_tensor(gathered[k], counts_dev[k], async_op=True))
        return c

    def drain_gathers():
        if world > 1 and args.gather == "final" and steps_done[0]:
            # one exchange for all the steps since the last drain: slot k of `gathered` = every rank's counters of step k
            counts_dev.copy_(counts_host, non_blocking=True)
            dist.all_gather_into_tensor(gathered_all.view(-1), counts_dev.view(-1))
            gathered.copy_(gathered_all.permute(1, 0, 2).reshape(n_slots, 5 * world))
            steps_done[0] = 0
        for w in pending:
            w.wait()                 # stream-level wait (no host synchronisation)
        del pending[:]

    for _ in range(max(args.warmup, 1)):
        step_device()
    drain_gathers()
    # NVML is initialised BEFORE the barrier: with N ranks doing it at once it takes a variable number of
    # milliseconds, and a rank that enters the timed region late makes every other rank wait for it in the closing
    # exchange (measured: 0.5 ms per step of pure start skew at 8 GPUs).
    sampler = ClockSampler(local_rank)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    sampler.start()
    # The K headline steps run without the library's stage events: a device-resident batch with unchanged arguments is
    # then replayed as ONE CUDA graph (captured during the warm-up; `graph` in the line says whether that happened).
    # The per-stage kernel times of the roofline come from K more steps with the events on, right after the timed region.
    eng.set_timing(False)
    for _ in range(3):
        step_device()                # eager -> capture -> replay
    drain_gathers()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    launches0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(args.steps):
        c = step_device()
    e_own = torch.cuda.Event(enable_timing=True)
    e_own.record(stream)             # this rank's own K steps end here; what follows waits for the slowest rank
    drain_gathers()
    e1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop()
    launches = eng.launch_count() - launches0
    graph_state = eng.graph_state()
    ms_total = e0.elapsed_time(e1)
    eng.set_timing(True)
    score_ms = []
    stage_acc = {}
    step_device()
    for _ in range(args.steps):
        before = stage_acc.get("filter", 0.0)
        step_device(stage_acc)
        score_ms.append(stage_acc["filter"] - before)
    drain_gathers()
    torch.cuda.synchronize()
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    per_rank = None
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # diagnostics: every rank's own step time (without the closing exchange) and the sum of its kernel stages
        mine = torch.tensor([e0.elapsed_time(e_own) / args.steps, sum(stage_acc.values()) / args.steps],
                            dtype=torch.float64, device=dev)
        allr = torch.zeros((world, 2), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allr.view(-1), mine)
        per_rank = {"own_ms_per_step": [round(float(x), 4) for x in allr[:, 0].tolist()],
                    "kernel_stage_ms_per_step": [round(float(x), 4) for x in allr[:, 1].tolist()]}
    ms_step = float(t.item()) / args.steps
    hsps_all = batch.n_hsp * world                       # one batch per rank: what the e2e arm moves per step
    hsps_value = (batch.n_hsp + more_hsp) * world        # the device-resident arm passes over every batch per step
    value = hsps_value / (ms_step * 1e-3)

    # roofline of the dominant kernel: k_front = score filter + p-values + ranks + q + keep flags, one warp per alignment
    # (stage "filter" of the library's event timeline; the alignments it leaves to k_pq / k_rank_big are stage "pq").
    # Algorithmic bytes (SURVEY.md 8d, K3+K4a+K5): the score of every HSP read once (8 B float64, or 1/2/4 B narrow +
    # 12 B per exception entry) + per SURVIVOR 4 B position + 8 B score + 8 B p + 8 B q + 1 B keep written.
    peak, peak_src = measured_peak()
    score_bytes = {0: 8.0, 8: 1.0, 16: 2.0, 32: 4.0}[args.score_bits]
    n_exc = (int(dbatch.t["exc_pos"].numel()) if args.score_bits else 0) + more_exc
    alg_bytes = score_bytes * (batch.n_hsp + more_hsp) + 12.0 * n_exc + 29.0 * c.n_survivors
    avg_score_ms = float(np.mean(score_ms))
    achieved = alg_bytes / (avg_score_ms * 1e-3) / 1e9
    kernel = "k_front<%s>" % {0: "double", 8: "signed char", 16: "short", 32: "int"}[args.score_bits]
    roofline = {"bound": "hbm", "kernel": kernel, "achieved": achieved,
                "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(kernel, args.config, batch.n_lnc, args.fitter), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": avg_score_ms,
                "stage_ms_per_step": {k: v / args.steps for k, v in stage_acc.items()}}
    # the same figure for the whole step: every algorithmic byte of the path (fit: the background read once, 4 B per
    # score here; gather + chain: 28 B per retained HSP) over the step time
    step_bytes = alg_bytes + 4.0 * (batch.n_bg + more_bg) + 28.0 * c.n_retained
    roofline["whole_step"] = {"algorithmic_bytes": step_bytes, "achieved": step_bytes / (ms_step * 1e-3) / 1e9,
                              "frac": step_bytes / (ms_step * 1e-3) / 1e9 / peak}

    # ---------------- end-to-end arm (host buffers through the C-ABI) ----------------
    # The host batch is the layout `ingest.load_files` returns (hostbatch.HostBatch: pinned, lossless narrow scores +
    # exception lists, narrow background, AoS coordinates), produced here from the synthetic columnar batch by the
    # same native encoder + allocator the ingest uses (HostBatch.from_columnar -> o2a_ingest_pack_narrow); the time
    # that packing takes is reported as `pack_ms` (a caller that parses JSON gets the layout from the parser itself).
    from ortho2align_b200.hostbatch import HostBatch
    e2e_steps = args.e2e_steps or max(2, min(args.steps, 5))
    # KDE: the reference evaluates ndtr once per (surviving HSP x background score) (fitting.py:295); the library
    # evaluates it per (score class x distinct background value). Both counts, against the device's measured ndtr rate.
    kde = None
    if with_kde:
        r_k = eng.fetch(survivors=False)                 # last batch of the last step
        last = more[-1][0] if more else dbatch
        aln_off = last.t["aln_off"].cpu().numpy(); bg_off = last.t["bg_off"].cpu().numpy()
        surv_l = np.add.reduceat(r_k.n_surv.astype(np.int64), aln_off[:-1][aln_off[:-1] < len(r_k.n_surv)]) if len(r_k.n_surv) else np.zeros(0)
        n_bg_l = np.diff(bg_off)[:len(surv_l)]
        ref_evals = float((surv_l * n_bg_l).sum()) * args.batches
        ndtr_rate = eng.fp64_probe("ndtr", 4000)
        dfma_rate = eng.fp64_probe("dfma", 20000)
        kde = {"reference_ndtr_evals_per_step": ref_evals, "reference_equivalent_evals_per_s": ref_evals / (ms_step * 1e-3),
               "device_ndtr_evals_per_s_probe": ndtr_rate, "device_dfma_per_s_probe": dfma_rate,
               "note": "value compression (integer backgrounds: ~50 distinct scores) replaces survivors x n_bg evaluations "
                       "by classes x distinct values; the probe is the library's own ndtr (0.5 erfc) at full occupancy"}
        roofline["kde"] = kde
    del dbatch, dstruct, more
    torch.cuda.empty_cache()
    if args.no_e2e:
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                              "warmup": args.warmup, "ms_per_step": ms_step, "roofline": roofline, "e2e": None,
                              "gpu_launches": int(launches), "graph": GRAPH_STATES[graph_state], "clocks": clocks, "note": "--no-e2e profiling run"}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        eng.close()
        return
    t0 = time.perf_counter()
    hb = HostBatch.from_columnar(batch, threads=procs, pinned=True, score_bits=args.e2e_score_bits or None,
                                 pin_coords="zerocopy" in args.e2e_coords)
    pack_ms = (time.perf_counter() - t0) * 1e3
    if with_kde:
        hb.ensure_kde_factor()
    hstruct = eng.hostbatch_struct(hb, with_kde)
    h2d = hb.h2d_bytes()

    def time_e2e(engines, structs, steps):
        """`steps` batches through `engines` (one host thread each, alternating batches): wall seconds, last results."""
        results, errors = [None] * len(engines), []

        def worker(i):
            try:
                for _ in range(steps // len(engines)):
                    engines[i].run(structs[i], params)
                    results[i] = engines[i].fetch(survivors=False, pinned=True)   # per-lncRNA thr/status/best + chains
            except Exception as ex:
                errors.append(f"{type(ex).__name__}: {ex}")

        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        if len(engines) == 1:
            worker(0)
        else:
            ths = [threading.Thread(target=worker, args=(i,)) for i in range(len(engines))]
            for th in ths:
                th.start()
            for th in ths:
                th.join()
        torch.cuda.synchronize()
        return time.perf_counter() - t0, results, errors

    # how the retained HSPs' coordinate records reach the device: every requested mode is measured with one context,
    # the fastest (max over ranks) is the one the headline uses
    modes = [m for m in args.e2e_coords.split(",") if m]
    mode_ms = {}
    for m in modes:
        eng.set_coords_mode(m)
        time_e2e([eng], [hstruct], 1)                                  # warm-up: buffers of this mode
        dt_m, res_m, err_m = time_e2e([eng], [hstruct], e2e_steps)
        tm = torch.tensor([dt_m / e2e_steps * 1e3 if not err_m else 1e30], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        mode_ms[m] = float(tm.item())
    best_mode = min(mode_ms, key=mode_ms.get)
    eng.set_coords_mode(best_mode)
    time_e2e([eng], [hstruct], 1)
    dt, results, errors = time_e2e([eng], [hstruct], e2e_steps)
    if errors:
        raise RuntimeError(errors[0])
    res = results[0]
    e2e_stage_ms = eng.stage_ms()
    single_engine_ms = dt / e2e_steps * 1e3
    dual_engine_ms = None
    dual_engine_error = None
    if args.e2e_engines == 2:
        # Two contexts on the same GPU, one host thread each (different ctxs may be driven from different threads,
        # include/o2a.h): every batch still crosses PCIe in full and its result is read back in full, but the link
        # no longer idles while a batch's last chunk computes and its result is downloaded.
        # Nothing in here may take a rank out of the collectives below: failures are recorded, not raised.
        eng2 = None
        try:
            eng2 = Engine(local_rank)
            eng2.set_coords_mode(best_mode)
            hstruct2 = eng2.hostbatch_struct(hb, with_kde)
            eng2.run(hstruct2, params)
            eng2.fetch(survivors=False, pinned=True)
        except Exception as ex:                                  # e.g. out of device memory for the second workspace
            dual_engine_error = f"{type(ex).__name__}: {ex}"
        dual_steps = 2 * ((e2e_steps + 1) // 2)
        dt2 = float("inf")
        if dual_engine_error is None:
            # the product's scheduler for many host batches (ortho2align_b200.stream.BatchStream): one worker thread
            # per context, batches handed out in order, results collected in order
            from ortho2align_b200.stream import BatchStream
            bs = BatchStream(engines=[eng, eng2])
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            try:
                t0 = time.perf_counter()
                results2 = list(bs.map((hb for _ in range(dual_steps)), args.fitter, True, 0.05, 5, 2,
                                       survivors=False, pinned=True))
                torch.cuda.synchronize()
                dt2 = time.perf_counter() - t0
                if not (np.array_equal(results2[-1].chain_idx, res.chain_idx) and np.array_equal(results2[-2].chain_idx, res.chain_idx)):
                    dual_engine_error = "the two contexts returned different chains"
            except Exception as ex:
                dual_engine_error = f"{type(ex).__name__}: {ex}"
        elif world > 1:
            torch.cuda.synchronize()
            dist.barrier()
        dual_engine_ms = None if dual_engine_error else dt2 / dual_steps * 1e3
        # the slowest rank decides, on every rank alike (8 ranks x 2 contexts can be host-bound: then one context is kept)
        both = torch.tensor([single_engine_ms, dual_engine_ms if dual_engine_ms is not None else 1e30],
                            dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(both, op=dist.ReduceOp.MAX)
        if float(both[1]) <= float(both[0]):
            dt, e2e_steps = dt2, dual_steps
        else:
            args.e2e_engines = 1
        if eng2 is not None:
            eng2.close()
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item()) / e2e_steps
    d2h = int(res.thr.nbytes + res.status.nbytes + res.best_aln.nbytes + res.n_surv.nbytes + res.n_keep.nbytes
              + res.chain_len.nbytes + res.chain_orient.nbytes + res.chain_score.nbytes + res.orient_scores.nbytes
              + res.chain_off.nbytes + res.chain_idx.nbytes + res.chain_pq.nbytes + 40)
    chunks = int(eng.lib.o2a_pipeline_chunks(eng.h))
    n_ret = int(res.n_keep.sum())
    if best_mode == "zerocopy":
        # coordinates stay in pinned host memory; only the retained HSPs are read over PCIe (one 32 B sector each)
        h2d_total = int(h2d + 32 * n_ret)
    elif best_mode == "gather":
        # the GPU writes the retained HSPs' indices to host memory (8 B each, counted as D2H), the host sends back
        # their 16-byte records
        h2d_total, d2h = int(h2d + 16 * n_ret), int(d2h + 8 * n_ret)
    else:
        h2d_total = int(h2d + hb.coords.nbytes)
    e2e = {"value": hsps_all / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d_total, "d2h_bytes_per_step": d2h,
           "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
           "host_layout": "hostbatch.HostBatch = what ingest.load_files returns (pinned; native encoder o2a_ingest_pack_narrow)",
           "score_layout": LAYOUTS[hb.score_bits], "bg_bits": hb.bg_bits, "pinned": bool(hb.pinned), "pack_ms": pack_ms,
           "coords_mode": best_mode, "coords_mode_ms_single_engine": mode_ms, "coords_zero_copy": best_mode == "zerocopy",
           "pipelined_chunks": chunks, "engines_per_gpu": args.e2e_engines, "single_engine_ms_per_step": single_engine_ms,
           "dual_engine_ms_per_step": dual_engine_ms, "dual_engine_error": dual_engine_error,
           ("stage_ms_last_step" if chunks <= 1 else "stage_ms_last_chunk"): e2e_stage_ms}

    # the ceiling this number runs against: all ranks copying pinned memory to their GPUs at once (plain cudaMemcpyAsync)
    try:
        ceil = pcie_ceiling(dev, world, 256, 10)
        mine = ceil["h2d"]["per_rank_gbs"]
        e2e["pcie_ceiling"] = ceil
        e2e["h2d_gbs_per_rank_achieved"] = round(h2d_total / e2e_s / 1e9, 2)
        e2e["h2d_fraction_of_ceiling"] = round(h2d_total / e2e_s / 1e9 / (min(mine) if min(mine) > 0 else 1.0), 3)
    except Exception as ex:
        e2e["pcie_ceiling"] = {"error": f"{type(ex).__name__}: {ex}"}

    # second end-to-end figure: the reference's own dtypes in host memory (float64 scores, int32 background, AoS
    # int32 coordinates, all pinned) through the same call, one context
    if args.e2e_f64 and world == 1:
        del hstruct
        hb.close()

        def pin(arr):
            tp = torch.empty(arr.shape, dtype=torch.from_numpy(arr[:0]).dtype, pin_memory=True)
            tp.numpy()[...] = arr
            return tp
        keep_alive = {k: pin(getattr(batch, k)) for k in ("bg_off", "aln_off", "hsp_off", "qlen", "slen", "bg", "score")}
        keep_alive["coords"] = pin(batch.coords4())
        batch._coords4 = None
        fields = {k: keep_alive[k].numpy() for k in ("bg_off", "aln_off", "hsp_off", "qlen", "slen", "bg", "score")}
        for k in ("qstart", "qend", "sstart", "send"):
            fields[k] = getattr(batch, k)            # SoA coordinates are ignored when `coords` is set
        fb = type(batch)(**fields, kde_factor=batch.kde_factor, names=None)
        fstruct = eng.host_struct(fb, with_kde, coords=keep_alive["coords"].numpy())
        eng.set_coords_mode(best_mode)
        time_e2e([eng], [fstruct], 1)
        f_steps = max(2, min(e2e_steps, 3))
        dt_f, res_f, err_f = time_e2e([eng], [fstruct], f_steps)
        tf = torch.tensor([dt_f / f_steps if not err_f else 1e30], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tf, op=dist.ReduceOp.MAX)
        f_h2d = sum(int(keep_alive[k].numpy().nbytes) for k in keep_alive if k != "coords") + 16 * n_ret
        e2e["float64_host_input"] = {
            "value": hsps_all / float(tf.item()), "unit": UNIT, "ms_per_step": float(tf.item()) * 1e3, "steps": f_steps,
            "h2d_bytes_per_step": int(f_h2d), "engines_per_gpu": 1, "coords_mode": best_mode,
            "score_layout": LAYOUTS[0], "error": err_f[0] if err_f else None,
            "same_chains": bool(res_f[0] is not None and np.array_equal(res_f[0].chain_idx, res.chain_idx))}
        del fstruct, keep_alive

    # ---------------- CPU baseline (rank 0, N=1 only) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as O
        threads = O.host_threads()
        k = min(args.cpu_lnc, batch.n_lnc)
        sample = batch.slice_lnc(0, k)
        rate, secs = cpu_port_rate(sample, args.fitter, threads, repeats=3)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"first {k} lncRNAs ({sample.n_hsp} HSPs) of the workload, {secs:.2f} s per pass, best of 3"}
        if not args.no_python_reference:
            # beside the C port (the harder baseline): the reference itself, through its own process pool, on all cores
            try:
                cpu["python_reference"] = python_reference_rate(args.config if args.config in (1, 2, 5) else 2, args.fitter
                                                                if args.fitter in ("hist", "kde") else "hist",
                                                                os.cpu_count() or 1, args.pyref_lnc)
            except Exception as ex:
                cpu["python_reference"] = {"unavailable": f"{type(ex).__name__}: {ex}"}

    strong = None
    if world > 1 and not args.no_strong:
        # beside the weak-scaling figure: ONE fixed workload (BASELINE configs[4]) sharded over the ranks
        strong = strong_scaling(args, eng, params, dev, rank, world, procs, args.strong_config, args.strong_lnc,
                                max(args.steps, 10), max(args.warmup, 3))
    files = None
    if rank == 0 and world == 1 and not args.no_files:
        # directories of the reference's JSON files -> the 7 + 4 output files (drop-in build_orthologs +
        # get_best_orthologs), BASELINE configs[4] shape; the JSON is written to /dev/shm by a fork pool first
        try:
            n_files_lnc = args.files_lnc
            if os.path.isdir("/dev/shm"):
                free = os.statvfs("/dev/shm"); free_gb = free.f_bavail * free.f_frsize / 1e9
                if free_gb < 12:
                    n_files_lnc = min(n_files_lnc, 10_000)
            # in a process of its own, like a user's `build_orthologs` run: this one holds gigabytes of synthetic batches
            # and pinned buffers, which makes every fork of the drop-in's worker pools (and every page fault) dearer
            import subprocess
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--stage", "files", "--config", "5",
                                "--lnc", str(n_files_lnc)], capture_output=True, text=True, timeout=900)
            if r.returncode != 0:
                raise RuntimeError(r.stderr.strip().splitlines()[-1] if r.stderr.strip() else f"exit code {r.returncode}")
            files = json.loads(r.stdout.strip().splitlines()[-1])
        except Exception as ex:                                # never lose the headline line to the secondary figure
            files = {"error": f"{type(ex).__name__}: {ex}"}
    if world > 1:
        g = gathered[0].view(world, 5).cpu().numpy()          # slot 0: first timed step (same batch every step)
        assert int(g[rank, 0]) == int(c.n_survivors)
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": CONFIG_NAMES[args.config], "batches_resident": args.batches, "lncRNAs_per_gpu": batch.n_lnc * args.batches,
                           "alignments_per_gpu": batch.n_aln * args.batches, "hsps_per_gpu": batch.n_hsp + more_hsp,
                           "bg_scores_per_gpu": batch.n_bg + more_bg, "fitter": args.fitter, "fdr": True, "alpha": 0.05,
                           "score_layout": LAYOUTS[args.score_bits],
                           "survivors_per_gpu": int(c.n_survivors), "orthologs_per_gpu": int(c.n_chains),
                           "parallelism": f"lncRNA shards x{world}, no data-path collective",
                           "l2": "inputs (%.2f GB/GPU) larger than L2; no flush needed" % (batch.nbytes() / 1e9),
                           "gen_seconds": round(t_gen, 1)},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "graph": GRAPH_STATES[graph_state],
                "clocks": clocks}
        if per_rank:
            line["per_rank"] = per_rank
        if strong:
            line["strong_scaling"] = strong
        if files:
            line["files"] = files
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    eng.close()


if __name__ == "__main__":
    main()
