"""Ingest throughput: JSON directories -> columnar batch.

    python tools/bench_ingest.py [--lnc 48] [--threads 0]

Times, on the same generated `align_files/` + `bg_files/` directories (BASELINE config 2 shape):
  python : json.load + ortho2align_b200.columnar.pack_lncrna            (1 process)
  native : libo2a_ingest.so with 1 thread and with all host threads
  sidecar: the columnar sidecar (ingest.save_sidecar / open_sidecar) written once and memory-mapped
  reference (only where /root/reference exists): json.load + GenomicRangesAlignment.from_dict
"""
import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from ortho2align_b200 import ingest, synth  # noqa: E402
from ortho2align_b200.pipeline import _load_lncrna  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lnc", type=int, default=48)
    ap.add_argument("--threads", type=int, default=0)
    args = ap.parse_args()
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


if __name__ == "__main__":
    main()
