"""Record golden vectors from the UNMODIFIED reference (container-only; needs /root/reference).

    python oracle/make_golden.py            # rewrites tests/golden/*.json

The reference has no tests/fixtures of its own, so these files — its outputs as
executed here with the recorded numpy/scipy versions — are what pins the oracle and the CUDA path
on the GPU box, where /root/reference does not exist.
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
from ortho2align_b200 import synth  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def versions():
    import scipy
    return {"numpy": np.__version__, "scipy": scipy.__version__,
            "reference": "dmitrymyl/ortho2align v1.0.5 (/root/reference)",
            "note": "pebble replaced by the in-process stand-in of oracle/ref_harness.py"}


def batch_to_lists(b):
    return {k: getattr(b, k).tolist() for k in
            ("bg_off", "bg", "aln_off", "hsp_off", "qlen", "slen", "qstart", "qend", "sstart", "send", "score")} \
        | {"names": b.names}


def golden_build(name, cfg, n_lnc, hsps, n_bg, fitting, fdr, seed, regions=None):
    b, m = synth.workload(cfg, n_lnc=n_lnc, with_meta=True, hsps_per_region=hsps, n_bg=n_bg, seed=seed,
                          regions=regions)
    js = synth.to_json_dicts(b, m)
    out = {"versions": versions(), "fitting": fitting, "fdr": fdr, "alpha": 0.05, "gapopen": 5, "gapextend": 2,
           "batch": batch_to_lists(b), "meta": m, "lncrnas": []}
    for (nm, alns, bg) in js:
        stable = rh.run_reference_lncrna(alns, bg, fitting, fdr, stable_ties=True)
        if fdr:
            unstable = rh.run_reference_lncrna(alns, bg, fitting, fdr, stable_ties=False)
            for ra, ru in zip(stable["alignments"], unstable["alignments"]):
                ra["rank_default_argsort"] = ru["rank"]
                ra["keep_default_argsort"] = ru["keep"]
                ra["chain_default_argsort"] = ru["chain"]
        out["lncrnas"].append(stable)
    with open(os.path.join(GOLDEN, name), "w") as f:
        json.dump(out, f)
    print(name, os.path.getsize(os.path.join(GOLDEN, name)))


def golden_fitters():
    ref = rh.import_reference()
    rng = np.random.default_rng(11)
    cases = []
    specs = [(2, None), (3, None), (7, None), (64, None), (257, None), (1000, None), (1001, None), (4096, None),
             (10000, None), (50, "wide"), (333, "wide"), (1000, "narrow"), (20001, None)]
    for n, kind in specs:
        if kind == "wide":
            bg = rng.integers(-5000, 100000, size=n)
        elif kind == "narrow":
            bg = rng.integers(15, 18, size=n)
        else:
            bg = synth.null_scores(rng, n).astype(np.int64)
        bg = [int(x) for x in bg]
        if len(set(bg)) <= 1:
            bg += [1]
        if bg[0] == 1:
            bg.append(0)
        lo, hi = min(bg), max(bg)
        xs = sorted(set([float(lo - 1), float(lo), float(hi), float(hi + 1)]
                        + [float(x) for x in rng.uniform(lo - 2, hi + 2, size=40)]
                        + [float(x) for x in rng.integers(lo, hi + 1, size=40)]))
        alphas = [0.05, 0.01, 0.5, 0.001, 0.9]
        h = ref.fitting.HistogramFitter(list(bg))
        case = {"bg": bg, "x": xs, "alphas": alphas,
                "hist_sf": [float(v) for v in h.sf(xs)],
                "hist_isf": [float(h.isf(a)) for a in alphas],
                "hist_hbins_head": [float(v) for v in h.estimator._hbins[:8]],
                "hist_hcdf_tail": [float(v) for v in h.estimator._hcdf[-8:]],
                "hist_hcdf_sum": float(np.sum(h.estimator._hcdf))}
        if n <= 4096:
            k = ref.fitting.KernelFitter(list(bg))
            case["kde_factor"] = float(k.estimator.factor)
            case["kde_sf"] = [float(v) for v in k.sf(xs)]
            case["kde_isf"] = [float(k.isf(a)) for a in alphas[:2]]
        cases.append(case)
    # ranking (fitting.py:359-364) with stable argsort + the (m*p)/rank step (pipeline.py:770-774)
    rank_cases = []
    for n in (0, 1, 5, 33, 200):
        p = rng.choice([0.0, 1e-9, 0.001, 0.01, 0.02, 0.5], size=n) * rng.choice([1.0, 1.0, 0.7], size=n)
        order = np.argsort(p, kind="stable")
        ranks = np.empty_like(p)
        ranks[order] = np.arange(n) + 1
        m = n * 13 + 7
        q = m * p / ranks
        rank_cases.append({"p": p.tolist(), "rank": ranks.tolist(), "m": m, "q": q.tolist()})
    with open(os.path.join(GOLDEN, "fitters.json"), "w") as f:
        json.dump({"versions": versions(), "cases": cases, "ranking": rank_cases}, f)
    print("fitters.json", os.path.getsize(os.path.join(GOLDEN, "fitters.json")))


def golden_chains():
    ref = rh.import_reference()
    A = ref.alignment
    rng = np.random.default_rng(5)
    cases = []
    for it in range(400):
        n = int(rng.integers(0, 60)) if it < 380 else int(rng.integers(150, 400))
        span = int(rng.choice([30, 100, 1000, 100000]))
        qs = rng.integers(1, span, size=n)
        qe = qs + rng.integers(1, max(2, span // 5), size=n)
        ss = rng.integers(1, span, size=n)
        sl = rng.integers(0, max(2, span // 5), size=n)
        rev = rng.random(n) < 0.5
        se = np.where(rev, ss - sl, ss + np.maximum(sl, 1))
        sc = rng.integers(1, 60, size=n).astype(float)
        if it % 3 == 0:
            sc = sc * rng.uniform(0.3, 1, size=n)
        if it % 4 == 0 or n == 0:
            qlen = slen = None
        else:
            qlen, slen = int(rng.integers(1, span * 2)), int(rng.integers(1, span * 2))
        hs = [A.HSPVertex(int(qs[i]), int(qe[i]), int(ss[i]), int(se[i]),
                          (int(sc[i]) if float(sc[i]).is_integer() else float(sc[i])), kw=i) for i in range(n)]
        al = A.Alignment(hs, qlen, slen)
        G, E = int(rng.integers(0, 8)), int(rng.integers(0, 4))
        ch = al.get_best_chains(gapopen=G, gapextend=E)
        best = ch["direct"] if ch["direct"].score > ch["reverse"].score else ch["reverse"]
        cases.append({"qstart": qs.tolist(), "qend": qe.tolist(), "sstart": ss.tolist(), "send": se.tolist(),
                      "score": sc.tolist(), "qlen": int(al.qlen), "slen": int(al.slen), "gapopen": G, "gapextend": E,
                      "chain": [h.kwargs["kw"] for h in best.HSPs],
                      "orient": 0 if best is ch["direct"] else 1,
                      "score_direct": float(ch["direct"].score), "score_reverse": float(ch["reverse"].score)})
    with open(os.path.join(GOLDEN, "chains.json"), "w") as f:
        json.dump({"versions": versions(), "cases": cases}, f)
    print("chains.json", os.path.getsize(os.path.join(GOLDEN, "chains.json")))


def golden_e2e():
    """Reference build_orthologs + get_best_orthologs on directories (in-process pool stand-in, see
    ref_harness._SerialPool for why the reference's process pool cannot be used): the 7+4 output files
    and the stats block, plus the input JSON needed to regenerate the directories."""
    b, m = synth.workload(1, n_lnc=10, with_meta=True, hsps_per_region=120, n_bg=400, seed=3)
    js = synth.to_json_dicts(b, m)
    # make the run cover every output file: one lncRNA with no significant HSP at all ("dropped"),
    # one with a missing background file, one malformed (orientation None -> exception)
    for a in js[3][1]:
        for h in a["HSPs"]:
            h["score"] = 12
        a["filtered_HSPs"] = [dict(h) for h in a["HSPs"]]
    missing_bg = js[5][0]
    for h in js[7][1][0]["HSPs"]:
        h["score"] = 500
        h["qend"] = h["qstart"]
    out = {"versions": versions(), "inputs": [], "missing_bg": missing_bg, "runs": {}}
    for nm, alns, bg in js:
        out["inputs"].append({"name": nm, "alignments": alns, "bg": bg})
    tmp = tempfile.mkdtemp(prefix="o2a_golden_")
    try:
        ad, bd = os.path.join(tmp, "align"), os.path.join(tmp, "bg")
        os.makedirs(ad), os.makedirs(bd)
        for nm, alns, bg in js:
            with open(os.path.join(ad, nm + ".json"), "w") as f:
                json.dump(alns, f)
            if nm != missing_bg:
                with open(os.path.join(bd, nm + ".json"), "w") as f:
                    json.dump(bg, f)
        out["listdir_order"] = [fn for fn in os.listdir(ad) if fn.endswith("json")]
        for fitting, fdr in (("hist", True), ("hist", False), ("kde", True)):
            od = os.path.join(tmp, f"out_{fitting}_{int(fdr)}")
            rh.run_reference_build(ad, bd, od, fitting=fitting, fdr=fdr, cores=1, pool="serial")
            files = {}
            for fn in sorted(os.listdir(od)):
                with open(os.path.join(od, fn)) as f:
                    files[fn] = f.read()
            with open(od + ".stats.txt") as f:
                files["stats.txt"] = f.read()
            out["runs"][f"{fitting}_{int(fdr)}"] = files
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    with open(os.path.join(GOLDEN, "e2e.json"), "w") as f:
        json.dump(out, f)
    print("e2e.json", os.path.getsize(os.path.join(GOLDEN, "e2e.json")))


if __name__ == "__main__":
    os.makedirs(GOLDEN, exist_ok=True)
    golden_build("build_hist_fdr.json", 1, 8, 160, 1000, "hist", True, seed=1)
    golden_build("build_hist_nofdr.json", 1, 4, 120, 600, "hist", False, seed=2)
    golden_build("build_kde_fdr.json", 1, 4, 120, 500, "kde", True, seed=4)
    golden_build("build_c5_hist_fdr.json", 5, 40, 60, 200, "hist", True, seed=6)
    golden_fitters()
    golden_chains()
    golden_e2e()
