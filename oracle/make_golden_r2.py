"""Round-2 golden vectors from the UNMODIFIED reference (container-only; needs /root/reference):

    python oracle/make_golden_r2.py         # rewrites tests/golden/r2_*.json

  r2_fitters_wide.json  HistogramFitter / KernelFitter on backgrounds OUTSIDE the envelope the round-1 library had
                        (> 512 distinct values, value range > 2^18, negative and near-int32-limit values)
  r2_build_wide.json    `_build` step by step (ref_harness.run_reference_lncrna) on lncRNAs with such backgrounds
  r2_best_weight.json   build_orthologs + get_best_orthologs for every -value x -function on inputs with fractional
                        (cut) scores, plus a 'relative' alignment and the files it yields
  r2_ties.json          per alignment: the ranks / keep set / chain under numpy's DEFAULT (unstable) argsort next to the
                        stable ones, so the tie-policy deviation can be counted on the GPU box too
TEST INFRASTRUCTURE: nothing under ortho2align_b200/ imports this.
"""
from __future__ import annotations

import json
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh  # noqa: E402
from oracle.make_golden import GOLDEN, batch_to_lists, versions  # noqa: E402
from ortho2align_b200 import synth  # noqa: E402


def wide_backgrounds(rng):
    yield "12000 values in [0, 60000]: > 512 distinct, range < 2^18", rng.integers(0, 60001, size=12000)
    yield "6000 values in [-400000, 900000]: range > 2^18", rng.integers(-400000, 900001, size=6000)
    yield "3000 values over almost all of int32", rng.integers(-2**31 + 5, 2**31 - 5, size=3000)
    yield "700 values, 600 distinct, range 2^19", rng.choice(rng.integers(-1000, 2**19, size=600), size=700)
    yield "5000 negative values", -rng.integers(1, 40000, size=5000)
    yield "2 far apart values", np.array([-2_000_000_000, 2_000_000_000])
    yield "1500 values in [1000, 1600): the shared-memory window does not start at 0", rng.integers(1000, 1600, size=1500)


def golden_fitters_wide():
    ref = rh.import_reference()
    rng = np.random.default_rng(2024)
    cases = []
    for what, bg in wide_backgrounds(rng):
        bg = [int(x) for x in bg]
        lo, hi = min(bg), max(bg)
        xs = sorted(set([float(lo - 1), float(lo), float(hi), float(hi + 1)]
                        + [float(x) for x in rng.uniform(lo - 2, hi + 2, size=60)]
                        + [float(x) for x in rng.choice(bg, size=60)]))
        alphas = [0.05, 0.01, 0.5, 0.001, 0.9]
        h = ref.fitting.HistogramFitter(list(bg))
        case = {"what": what, "bg": bg, "x": xs, "alphas": alphas,
                "hist_sf": [float(v) for v in h.sf(xs)], "hist_isf": [float(h.isf(a)) for a in alphas]}
        if len(bg) <= 3000:
            k = ref.fitting.KernelFitter(list(bg))
            case["kde_factor"] = float(k.estimator.factor)
            case["kde_sf"] = [float(v) for v in k.sf(xs)]
            case["kde_isf"] = [float(k.isf(a)) for a in alphas[:2]]
        cases.append(case)
    with open(os.path.join(GOLDEN, "r2_fitters_wide.json"), "w") as f:
        json.dump({"versions": versions(), "cases": cases}, f)
    print("r2_fitters_wide.json", os.path.getsize(os.path.join(GOLDEN, "r2_fitters_wide.json")))


def golden_build_wide():
    """lncRNAs whose backgrounds and HSP scores live on a wide, partly negative scale (a non-BLAST scoring scheme)."""
    rng = np.random.default_rng(77)
    b, m = synth.workload(1, n_lnc=6, with_meta=True, hsps_per_region=140, n_bg=4000, seed=8)
    scale = [37, 211, 1000, 5000, 30011, 1]
    for l in range(b.n_lnc):
        b0, b1 = int(b.bg_off[l]), int(b.bg_off[l + 1])
        a0, a1 = int(b.aln_off[l]), int(b.aln_off[l + 1])
        h0, h1 = int(b.hsp_off[a0]), int(b.hsp_off[a1])
        shift = int(rng.integers(-200000, 200000))
        jitter = rng.integers(0, scale[l], size=b1 - b0)
        b.bg[b0:b1] = (b.bg[b0:b1] * scale[l] + jitter + shift).astype(np.int32)
        hj = rng.integers(0, scale[l], size=h1 - h0)
        b.score[h0:h1] = np.floor(b.score[h0:h1]) * scale[l] + hj + shift + (b.score[h0:h1] - np.floor(b.score[h0:h1]))
    js = synth.to_json_dicts(b, m)
    out = {"versions": versions(), "fitting": "hist", "fdr": True, "alpha": 0.05, "gapopen": 5, "gapextend": 2,
           "batch": batch_to_lists(b), "meta": m, "lncrnas": [], "distinct_bg": []}
    for (nm, alns, bg) in js:
        out["lncrnas"].append(rh.run_reference_lncrna(alns, bg, "hist", True, stable_ties=True))
        out["distinct_bg"].append(len(set(bg)))
    with open(os.path.join(GOLDEN, "r2_build_wide.json"), "w") as f:
        json.dump(out, f)
    print("r2_build_wide.json", os.path.getsize(os.path.join(GOLDEN, "r2_build_wide.json")), out["distinct_bg"])


def golden_best_weight():
    ref = rh.import_reference("serial")
    b, m = synth.workload(1, n_lnc=8, with_meta=True, hsps_per_region=150, n_bg=500, seed=12, regions=4)
    js = synth.to_json_dicts(b, m)
    rng = np.random.default_rng(5)
    # fractional scores on strong HSPs (what cut_coordinates leaves, alignment.py:330), so that the BED score column
    # is a float sum and near-ties between orthologs of one lncRNA are decided by its last digits
    for nm, alns, bg in js:
        for a in alns:
            for h in a["HSPs"]:
                if h["score"] > 30 and rng.random() < 0.6:
                    h["score"] = float(h["score"]) * float(rng.uniform(0.8, 1.0))
            a["filtered_HSPs"] = [dict(h) for h in a["HSPs"]]
    js[2][1][1]["relative"] = True            # to_record raises ValueError -> that alignment counts as dropped
    for a in js[6][1]:
        a["relative"] = True                  # a lncRNA whose every alignment is relative -> insignificant.*
    out = {"versions": versions(), "inputs": [{"name": nm, "alignments": alns, "bg": bg} for nm, alns, bg in js], "runs": {}}
    tmp = tempfile.mkdtemp(prefix="o2a_golden_r2_")
    try:
        ad, bd = os.path.join(tmp, "align"), os.path.join(tmp, "bg")
        os.makedirs(ad), os.makedirs(bd)
        for nm, alns, bg in js:
            with open(os.path.join(ad, nm + ".json"), "w") as f:
                json.dump(alns, f)
            with open(os.path.join(bd, nm + ".json"), "w") as f:
                json.dump(bg, f)
        out["listdir_order"] = [fn for fn in os.listdir(ad) if fn.endswith("json")]
        od = os.path.join(tmp, "out")
        ref.pipeline.build_orthologs(ad, bd, "hist", od, threshold=0.05, gapopen=5, gapextend=2, fdr=True, cores=1,
                                     timeout=None, silent=True, stats_filename=os.path.join(tmp, "stats.txt"))
        files = {fn: open(os.path.join(od, fn)).read() for fn in sorted(os.listdir(od))}
        files["stats.txt"] = open(os.path.join(tmp, "stats.txt")).read()
        out["build"] = files
        j = lambda n: os.path.join(od, n)
        for value in ("total_length", "block_count", "block_length", "weight"):
            for function in ("max", "min"):
                names = [f"best.{value}.{function}.{k}" for k in ("q.bed", "q.tsv", "s.bed", "s.tsv")]
                ref.pipeline.get_best_orthologs(j("significant.query_orthologs.bed"), j("significant.query_orthologs.tsv"),
                                                j("significant.subject_orthologs.bed"), j("significant.subject_orthologs.tsv"),
                                                value, function, *[j(n) for n in names])
                out["runs"][f"{value}.{function}"] = [open(j(n)).read() for n in names]
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    with open(os.path.join(GOLDEN, "r2_best_weight.json"), "w") as f:
        json.dump(out, f)
    print("r2_best_weight.json", os.path.getsize(os.path.join(GOLDEN, "r2_best_weight.json")))


def golden_ties():
    """The tie-policy deviation, counted: the reference ranks with numpy's default (unstable) argsort; the build ranks
    ties in HSP order. Per alignment both outcomes (ranks, keep set, chain) from the unmodified reference."""
    b, m = synth.workload(2, n_lnc=10, with_meta=True, seed=31, hsps_per_region=700, n_bg=3000)
    js = synth.to_json_dicts(b, m)
    out = {"versions": versions(), "batch": batch_to_lists(b), "alignments": []}
    for (nm, alns, bg) in js:
        st = rh.run_reference_lncrna(alns, bg, "hist", True, stable_ties=True)
        un = rh.run_reference_lncrna(alns, bg, "hist", True, stable_ties=False)
        for ra, ru in zip(st["alignments"], un["alignments"]):
            out["alignments"].append({"p": ra["p"], "rank_stable": ra["rank"], "rank_default": ru["rank"],
                                      "keep_stable": ra["keep"], "keep_default": ru["keep"],
                                      "chain_stable": ra["chain"], "chain_default": ru["chain"],
                                      "records_differ": ra["records"] != ru["records"]})
    with open(os.path.join(GOLDEN, "r2_ties.json"), "w") as f:
        json.dump(out, f)
    a = out["alignments"]
    print("r2_ties.json", os.path.getsize(os.path.join(GOLDEN, "r2_ties.json")), "alignments", len(a),
          "rank differs", sum(x["rank_stable"] != x["rank_default"] for x in a),
          "keep set differs", sum(x["keep_stable"] != x["keep_default"] for x in a),
          "chain differs", sum(x["chain_stable"] != x["chain_default"] for x in a),
          "records differ", sum(x["records_differ"] for x in a))


if __name__ == "__main__":
    os.makedirs(GOLDEN, exist_ok=True)
    golden_fitters_wide()
    golden_build_wide()
    golden_best_weight()
    golden_ties()
