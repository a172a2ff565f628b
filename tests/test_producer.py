"""Producer side of the alignment JSON: BLAST table -> to_genomic -> cut_coordinates.

CPU tests pin the oracle restatement (oracle/producer.py, oracle/o2a_oracle_producer.c) and the native
BLAST reader to tests/golden/producer.json, recorded from the unmodified reference
(oracle/make_golden_producer.py). GPU tests compare `o2a_cut_batch` with the golden vectors and, on
seeded synthetic batches, with the oracle — bit-exact: coordinates, order, scores, qlen/slen, flags.
"""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_golden

KEYS = ("hsp_off", "qstart", "qend", "sstart", "send", "score", "cut", "qlen", "slen", "status")


def _case_inputs(c, parsed):
    from oracle import producer as OP
    h = np.array([r[:4] for r in parsed["hsps"]], dtype=np.int64).reshape(-1, 4)
    rel = dict(hsp_off=np.array([0, len(h)], dtype=np.int64), qstart=h[:, 0].astype(np.int32), qend=h[:, 1].astype(np.int32),
               sstart=h[:, 2].astype(np.int32), send=h[:, 3].astype(np.int32),
               score=np.array([float(r[4]) for r in parsed["hsps"]], dtype=np.float64),
               score_is_int=np.array([isinstance(r[4], int) for r in parsed["hsps"]], dtype=np.uint8))
    q, s, b = c["qrange"], c["srange"], c["bounds"]
    one = lambda v: np.array([v], dtype=np.int64)
    rg = dict(q_start=one(q[1]), q_end=one(q[2]), q_minus=np.array([q[3] == "-"], dtype=np.int8),
              s_start=one(s[1]), s_end=one(s[2]), s_minus=np.array([s[3] == "-"], dtype=np.int8))
    for k in ("qleft", "qright", "sleft", "sright"):
        rg[k] = one(b[k] if b.get(k) is not None else OP.NO_BOUND)
    return rel, rg


def _assert_cut_equals_golden(r, want, what):
    got = np.stack([r["qstart"], r["qend"], r["sstart"], r["send"]], 1).tolist() if len(r["score"]) else []
    assert got == [x[:4] for x in want["hsps"]], f"{what}: coordinates / membership / order"
    assert r["score"].tolist() == [float(x[4]) for x in want["hsps"]], f"{what}: scores (bit-exact)"
    assert [c != 0 for c in r["cut"].tolist()] == [bool(x[5]) for x in want["hsps"]], f"{what}: int/float type of the score"
    assert (int(r["qlen"][0]), int(r["slen"][0])) == (want["qlen"], want["slen"]), f"{what}: qlen/slen"
    assert int(r["status"][0]) == 0


def _same_cut(got, exp, what):
    for k in KEYS:
        assert got[k].shape == exp[k].shape, f"{what} {k}: shape {got[k].shape} vs {exp[k].shape}"
        g, e = got[k], exp[k]
        if k in ("qlen", "slen"):       # undefined where a coordinate left int32 (status bit 0): the columnar layout cannot hold it
            ok = (exp["status"] & 1) == 0
            g, e = g[ok], e[ok]
        assert np.array_equal(g, e), f"{what} {k}: first diff at {np.nonzero(g != e)[0][:5]}"


# ------------------------------------------------------------------------------------------------ CPU
def test_oracle_parse_and_cut_match_the_reference_goldens():
    from oracle import producer as OP
    g = load_golden("producer.json")
    n_cut = 0
    for i, c in enumerate(g["cases"]):
        try:
            p = OP.parse_blast(c["text"])
        except Exception as e:
            name = "ValueError" if isinstance(e, OP.BlastFormatError) else type(e).__name__
            assert c.get("parse_error") == name, f"case {i}"
            continue
        assert "parse_error" not in c, f"case {i}: the reference raised {c.get('parse_error')}"
        assert p["hsps"] == [h[:5] for h in c["parsed"]["hsps"]]
        assert [type(h[4]) for h in p["hsps"]] == [float if h[5] else int for h in c["parsed"]["hsps"]]
        assert (p["qlen"], p["slen"]) == (c["parsed"]["qlen"], c["parsed"]["slen"])
        rel, rg = _case_inputs(c, p)
        _assert_cut_equals_golden(OP.cut_batch(rel, rg), c["cut"], f"case {i} {c['bounds']}")
        n_cut += 1
    assert n_cut >= 50


def test_native_blast_reader_matches_the_reference_goldens():
    from ortho2align_b200 import build, producer
    build.build_ingest()
    g = load_golden("producer.json")
    r = producer.parse_blast_tables([c["text"] for c in g["cases"]], threads=3)
    for i, c in enumerate(g["cases"]):
        h0, h1 = int(r["hsp_off"][i]), int(r["hsp_off"][i + 1])
        if "parse_error" in c:
            assert r["status"][i] == 1 and r["errors"][i] == c["parse_error"] and h1 == h0, f"case {i}"
            continue
        assert r["status"][i] == 0, f"case {i}: {r['errors'].get(i)}"
        want = c["parsed"]
        got = np.stack([r[k][h0:h1] for k in ("qstart", "qend", "sstart", "send")], 1).tolist() if h1 > h0 else []
        assert got == [h[:4] for h in want["hsps"]], f"case {i}"
        assert r["score"][h0:h1].tolist() == [float(h[4]) for h in want["hsps"]]
        assert r["score_is_int"][h0:h1].tolist() == [0 if h[5] else 1 for h in want["hsps"]]
        assert (int(r["qlen"][i]), int(r["slen"][i])) == (want["qlen"], want["slen"]), f"case {i}"


def test_native_blast_reader_edge_cases():
    from ortho2align_b200 import producer, synth
    f = synth.BLAST_FIELDS
    head = "# BLASTN\n# Query: q\n"
    row = "q\ts\t99.0\t10\t0\t0\t5\t14\t100\t109\t1e-3\t20.1\t"
    tables = [
        "",                                                             # not a table -> ValueError
        head + "".join(f"# filler {k}\n" for k in range(120)) + f"# Fields: {f}\n# 1 hits found\n{row}12\n",   # header too far
        head + f"# Fields: {f}\n# 1 hits found\n" + row + "12",         # no trailing newline
        head + f"# Fields: {f}\n# 2 hits found\n" + row + "12\n",       # fewer rows than announced -> TypeError
        head + f"# Fields: {f}\n#1hits\n" + row + "12\n",               # no second token -> IndexError
        head + f"# Fields: {f}\n# many hits found\n" + row + "12\n",    # int('many') -> ValueError
        head + f"# Fields: {f}\n# 1 hits found\n" + row + "12.5\n",     # float score stays float
        head + f"# Fields: {f}\n# 1 hits found\n" + row.replace("\t5\t", "\t5.5\t") + "12\n",   # non-integral coordinate
        head + f"# Fields: {f}\n# 1 hits found\n" + row.replace("\t5\t", "\tx\t") + "12\n",     # str coordinate -> TypeError
        head + f"# Fields: {f}\n# 1 hits found\n  \t" + row + "12 \t\n",                        # strip() eats outer tabs
        head + f"# Fields: {f}\n# 1 hits found\n" + row + "1e1\n",       # numberize('1e1') = 10 (int)
    ]
    r = producer.parse_blast_tables(tables, threads=2)
    assert r["status"].tolist() == [1, 0, 0, 1, 1, 1, 0, 2, 1, 0, 0]
    assert [r["errors"].get(i) for i in (0, 3, 4, 5, 8)] == ["ValueError", "TypeError", "IndexError", "ValueError", "TypeError"]
    n = np.diff(r["hsp_off"]).tolist()
    assert n == [0, 0, 1, 0, 0, 0, 1, 0, 0, 1, 1]
    assert (r["qlen"][1], r["slen"][1]) == (1, 1)                        # empty alignment defaults (alignment.py:477-487)
    k = int(r["hsp_off"][2])
    assert (r["qstart"][k], r["qend"][k], r["sstart"][k], r["send"][k]) == (4, 14, 99, 109)
    assert r["score"].tolist() == [12.0, 12.5, 12.0, 10.0] and r["score_is_int"].tolist() == [1, 0, 1, 1]


def test_blast_reader_symbols_are_exported():
    from ortho2align_b200 import build, producer
    build.build_ingest()
    lib = producer.load_lib()
    with open(os.path.join(ROOT, "include", "o2a_blast.h")) as f:
        src = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    names = sorted(set(re.findall(r"\b(o2a_(?:blast|bg)_[a-z_0-9]+)\s*\(", src)))
    assert set(names) == set(producer.BLAST_EXPORTS)
    for n in names:
        assert hasattr(lib, n)


def test_background_sampling_is_cpythons_random_sample():
    """`random.Random(seed).sample(scores, observations)` (pipeline.py:228-231, 264-267) restated natively: the
    interpreter's own `random` module is the reference here (both branches of `sample`, seeds beyond 32 bits)."""
    import random
    from ortho2align_b200 import build, producer
    build.build_ingest()
    rng = np.random.default_rng(0)
    cases = [(5000, 1000, 0), (1001, 1000, 0), (4117, 1000, 3), (4118, 1000, 3), (27, 6, 2 ** 40 + 5), (100000, 1000, 123456789),
             (10, 10, 0), (7, 3, 0), (21, 5, 9), (22, 5, 0), (2, 1, 0), (300, 64, 2 ** 32), (1, 5, 0), (0, 5, 0)]
    cases += [(int(rng.integers(1, 3000)), int(rng.integers(1, 400)), int(rng.integers(0, 2 ** 62))) for _ in range(150)]
    for n, k, seed in cases:
        sc = rng.integers(12, 90, size=n).astype(np.int32)
        if n > k:
            r = random.Random()
            r.seed(seed)
            want = r.sample(sc.tolist(), k)
        else:
            want = sc.tolist()
        assert producer.sample_background(sc, k, seed).tolist() == want, (n, k, seed)


def test_backgrounds_from_blast_tables(tmp_path):
    """BLAST tables of the background alignments -> per-query background arrays, as
    `_estimate_bg_for_single_query_blast` would dump them (scores of the tables concatenated in order, then sampled)."""
    import random
    from oracle import producer as OP
    from ortho2align_b200 import producer, synth
    rng = np.random.default_rng(6)
    tables, owner = [], []
    for q in range(5):
        for _ in range(int(rng.integers(1, 4))):
            tables.append(synth.blast_tabular(*synth.relative_hits(rng, int(rng.integers(0, 90)), 3000, 50_000)))
            owner.append(q)
    bg_off, bg, found = producer.backgrounds_from_tables(tables, owner, 6, observations=70, seed=3, threads=2)
    assert len(bg_off) == 7 and found[5] == 0 and bg_off[6] == bg_off[5]
    for q in range(5):
        scores = [h[4] for t, o in zip(tables, owner) if o == q for h in OP.parse_blast(t)["hsps"]]
        if len(scores) > 70:
            r = random.Random()
            r.seed(3)
            scores = r.sample(scores, 70)
        assert bg[bg_off[q]:bg_off[q + 1]].tolist() == scores and found[q] >= len(scores)


def test_oracle_cut_is_thread_count_independent_and_ordered():
    from oracle import producer as OP
    from ortho2align_b200 import synth
    rel, rg = synth.producer_workload(40, seed=3, hits_per_region=300)
    a = OP.cut_batch(rel, rg, n_threads=1)
    b = OP.cut_batch(rel, rg, n_threads=4)
    _same_cut(a, b, "threads")
    kept = a["hsp_off"][-1]
    assert 0 < kept < rel["hsp_off"][-1]
    assert (a["cut"] & 1).sum() > 0                                     # some HSPs straddle a gene boundary
    # every kept HSP lies inside the gene on the query side
    aln = np.repeat(np.arange(len(rg["qleft"])), np.diff(a["hsp_off"]))
    lo, hi = np.minimum(a["qstart"], a["qend"]), np.maximum(a["qstart"], a["qend"])
    plus = rg["q_minus"][aln] == 0
    assert np.all(lo[plus] >= rg["qleft"][aln][plus]) and np.all(hi[plus] <= rg["qright"][aln][plus])


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_gpu_cut_matches_the_reference_goldens(engine):
    from oracle import producer as OP
    g = load_golden("producer.json")
    n = 0
    for i, c in enumerate(g["cases"]):
        if "cut" not in c:
            continue
        rel, rg = _case_inputs(c, OP.parse_blast(c["text"]))
        _assert_cut_equals_golden(engine.cut_batch(rel, rg), c["cut"], f"case {i} {c['bounds']}")
        n += 1
    assert n >= 50


def _with_subject_bounds(rg, rng):
    rg = dict(rg)
    n = len(rg["qleft"])
    from ortho2align_b200.producer import NO_BOUND
    lo = rg["s_start"] + rng.integers(0, 900_000, size=n)
    hi = rg["s_end"] - rng.integers(0, 900_000, size=n)
    rg["sleft"] = np.where(rng.random(n) < 0.7, lo, NO_BOUND)
    rg["sright"] = np.where(rng.random(n) < 0.7, hi, NO_BOUND)
    rg["qleft"] = np.where(rng.random(n) < 0.9, rg["qleft"], NO_BOUND)
    rg["qright"] = np.where(rng.random(n) < 0.9, rg["qright"], NO_BOUND)
    rg["relative"] = (rng.random(n) < 0.9).astype(np.uint8)
    return rg


@pytest.mark.gpu
@pytest.mark.parametrize("n_lnc,hits,regions", [(60, 300, 5), (7, 5000, 3), (400, 9, 2), (3, 0, 4)])
def test_gpu_cut_vs_oracle_synthetic(engine, n_lnc, hits, regions):
    from oracle import producer as OP
    from ortho2align_b200 import synth
    rel, rg = synth.producer_workload(n_lnc, seed=11, hits_per_region=hits, regions=regions)
    rng = np.random.default_rng(5)
    rel["score_is_int"] = (rng.random(len(rel["score"])) < 0.95).astype(np.uint8)
    rel["score"] = np.where(rel["score_is_int"] == 1, rel["score"], rel["score"] * 0.37)
    exp = OP.cut_batch(rel, rg)
    _same_cut(engine.cut_batch(rel, rg), exp, "gene bounds")
    assert engine.cut_n_kept == int(exp["hsp_off"][-1])
    rg4 = _with_subject_bounds(rg, rng)
    _same_cut(engine.cut_batch(rel, rg4), OP.cut_batch(rel, rg4), "all four bounds, mixed None, mixed relative")
    none = {k: v for k, v in rg.items() if k not in ("qleft", "qright")}
    r = engine.cut_batch(rel, none)                                     # no boundary at all: to_genomic only
    _same_cut(r, OP.cut_batch(rel, none), "no bounds")
    assert r["hsp_off"][-1] == rel["hsp_off"][-1] and not (r["cut"] & 1).any()


@pytest.mark.gpu
def test_gpu_cut_tile_boundaries_and_empty_inputs(engine):
    """Alignment sizes around the 256-HSP warp tile, empty alignments between them, an empty batch."""
    from oracle import producer as OP
    from ortho2align_b200 import synth
    sizes = [0, 1, 255, 256, 257, 0, 0, 511, 512, 513, 1024, 31, 32, 33, 0]
    rng = np.random.default_rng(3)
    n_aln = len(sizes)
    hsp_off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    n = int(hsp_off[-1])
    qs = rng.integers(0, 3000, size=n); ln = rng.integers(12, 200, size=n)
    ss = rng.integers(0, 50_000, size=n); rev = rng.random(n) < 0.5
    rel = dict(hsp_off=hsp_off, qstart=qs.astype(np.int32), qend=(qs + ln).astype(np.int32),
               sstart=np.where(rev, ss + ln, ss).astype(np.int32), send=np.where(rev, ss, ss + ln).astype(np.int32),
               score=synth.null_scores(rng, n).astype(np.float64))
    one = lambda v: np.full(n_aln, v, dtype=np.int64)
    rg = dict(q_start=one(1_000_000), q_end=one(1_003_200), q_minus=(np.arange(n_aln) % 2).astype(np.int8),
              s_start=one(5_000_000), s_end=one(5_050_200), s_minus=(np.arange(n_aln) % 3 == 0).astype(np.int8),
              qleft=one(1_001_000), qright=one(1_002_000))
    _same_cut(engine.cut_batch(rel, rg), OP.cut_batch(rel, rg), "tile edges")
    # nothing survives: every alignment ends up with the reference's default qlen = slen = 1
    rg_none = dict(rg, qleft=one(9_000_000), qright=one(9_000_001))
    r = engine.cut_batch(rel, rg_none)
    _same_cut(r, OP.cut_batch(rel, rg_none), "all dropped")
    assert r["hsp_off"][-1] == 0 and set(r["qlen"].tolist()) == {1} and set(r["slen"].tolist()) == {1}
    # empty batch
    empty = {k: v[:0] for k, v in rel.items()}
    empty["hsp_off"] = np.zeros(1, dtype=np.int64)
    r = engine.cut_batch(empty, {k: v[:0] for k, v in rg.items()})
    assert r["hsp_off"].tolist() == [0] and len(r["score"]) == 0


@pytest.mark.gpu
def test_gpu_cut_flags_what_python_integers_would_do_differently(engine):
    from oracle import producer as OP
    big = 2 ** 30
    rel = dict(hsp_off=np.array([0, 1, 2], dtype=np.int64), qstart=np.array([0, 0], dtype=np.int32),
               qend=np.array([big, 100], dtype=np.int32), sstart=np.array([0, 0], dtype=np.int32),
               send=np.array([big, 100], dtype=np.int32), score=np.array([50.0, 50.0]))
    one = lambda a, b: np.array([a, b], dtype=np.int64)
    rg = dict(q_start=one(0, 2 ** 31 - 50), q_end=one(big, 2 ** 31 + 50), q_minus=np.zeros(2, np.int8),
              s_start=one(0, 0), s_end=one(big, 100), s_minus=np.zeros(2, np.int8),
              qleft=one(big - 7, OP.NO_BOUND), qright=one(OP.NO_BOUND, OP.NO_BOUND))
    got, exp = engine.cut_batch(rel, rg), OP.cut_batch(rel, rg)
    assert got["status"].tolist() == exp["status"].tolist() == [2, 1]      # inexact product; coordinate beyond int32


@pytest.mark.gpu
def test_gpu_cut_kernel_class_boundaries(engine):
    """Alignments on both sides of the fast / general kernel split, and HSPs beyond the 30-bit guarantee of the
    32-bit classification inside fast alignments (dropped, kept, straddling) — all bit-exact with the oracle."""
    from oracle import producer as OP
    rng = np.random.default_rng(12)
    big = 2 ** 30
    n_per, n_aln = 700, 8
    hsp_off = np.arange(0, n_per * n_aln + 1, n_per, dtype=np.int64)
    n = n_per * n_aln
    qs = rng.integers(0, 4000, size=n); ln = rng.integers(12, 300, size=n)
    ss = rng.integers(0, 60_000, size=n); rev = rng.random(n) < 0.5
    qe = qs + ln
    s1, s2 = np.where(rev, ss + ln, ss), np.where(rev, ss, ss + ln)
    aln = np.repeat(np.arange(n_aln), n_per)
    # alignment 2: query coordinates beyond 2^30 (q_start small, so genomic values still fit int32)
    qs = np.where(aln == 2, qs + big + 5, qs); qe = np.where(aln == 2, qe + big + 5, qe)
    # alignment 3: subject coordinates beyond 2^30 on a few HSPs only
    far = (aln == 3) & (rng.random(n) < 0.1)
    s1 = np.where(far, s1 + 2 ** 31 - 61_000, s1); s2 = np.where(far, s2 + 2 ** 31 - 61_000, s2)
    rel = dict(hsp_off=hsp_off, qstart=qs.astype(np.int32), qend=qe.astype(np.int32), sstart=s1.astype(np.int32),
               send=s2.astype(np.int32), score=rng.integers(12, 90, size=n).astype(np.float64))
    i64 = lambda xs: np.array(xs, dtype=np.int64)
    N = OP.NO_BOUND
    rg = dict(q_start=i64([1000, 1000, 7, 1000, 1000, big + 1, 1000, 1000]), q_end=i64([9000] * 8),
              q_minus=np.array([0, 1, 0, 0, 0, 0, 0, 1], np.int8),
              s_start=i64([5000] * 8), s_end=i64([90_000] * 8), s_minus=np.array([0, 0, 1, 0, 1, 0, 0, 0], np.int8),
              qleft=i64([2000, 6000, big + 1500, 2000, 3000, big + 2000, N, 2 ** 31 + 5]),
              qright=i64([4000, 8000, big + 3000, 4000, 2500, big + 3500, 3000, N]),      # alignment 4: qleft > qright
              relative=np.array([1, 1, 1, 1, 1, 1, 0, 1], np.uint8))
    got, exp = engine.cut_batch(rel, rg), OP.cut_batch(rel, rg)
    _same_cut(got, exp, "kernel classes")
    kept = np.diff(exp["hsp_off"])
    assert kept[0] > 0 and kept[2] > 0 and kept[3] > 0 and kept[6] > 0 and (exp["cut"] & 1).sum() > 20
    assert exp["status"][3] == 1                                  # subject coordinates beyond int32 after to_genomic


@pytest.mark.gpu
def test_cut_then_build_stays_on_the_device(engine):
    """cut -> o2a_cut_view -> o2a_build_batch with device pointers == fetch the cut, build from host arrays."""
    import torch
    from oracle import oracle as O
    from oracle import producer as OP
    from ortho2align_b200 import producer, synth
    n_lnc, regions = 50, 4
    rel, rg = synth.producer_workload(n_lnc, seed=21, hits_per_region=400, regions=regions)
    rng = np.random.default_rng(2)
    rel["score"] = rel["score"] + np.where(rng.random(len(rel["score"])) < 0.1, rng.exponential(18, len(rel["score"])).round(), 0)
    aln_off = np.arange(0, n_lnc * regions + 1, regions, dtype=np.int64)
    bg = synth.null_scores(rng, n_lnc * 200).astype(np.int32)
    bg_off = np.arange(0, n_lnc * 200 + 1, 200, dtype=np.int64)
    cut = producer.cut_alignments(engine, rel, rg)
    host_batch = producer.to_columnar(cut, aln_off, bg_off, bg)
    exp = O.build_batch(producer.to_columnar(OP.cut_batch(rel, rg), aln_off, bg_off, bg), "hist", True)
    from helpers import compare_results
    compare_results(engine.build_batch(host_batch, "hist", True), exp, exact_p=True, what="cut (fetched) -> build")
    # the same without leaving the device
    dev = torch.device("cuda:0")
    engine.run_cut(engine.cut_struct(rel, rg))
    s = engine.cut_view()
    t_aln, t_bgo, t_bg = (torch.from_numpy(x).to(dev) for x in (aln_off, bg_off, bg))
    torch.cuda.synchronize()
    s.n_lnc, s.n_bg = n_lnc, len(bg)
    s.aln_off, s.bg_off, s.bg = t_aln.data_ptr(), t_bgo.data_ptr(), t_bg.data_ptr()
    engine.run(s, engine.params("hist", True))
    compare_results(engine.fetch(), exp, exact_p=True, what="cut -> build on the device")
    assert exp.chain_len.sum() > 0


@pytest.mark.gpu
def test_cut_accepts_device_resident_inputs(engine):
    import torch
    from ortho2align_b200 import synth
    rel, rg = synth.producer_workload(30, seed=8, hits_per_region=200)
    host = engine.cut_batch(rel, rg)
    dev = torch.device("cuda:0")
    d_rel = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in rel.items()}
    d_rg = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in rg.items()}
    torch.cuda.synchronize()
    engine.run_cut(engine.cut_struct(d_rel, d_rg, device=True))
    _same_cut(engine.fetch_cut(), host, "device-resident inputs")
    _same_cut(engine.fetch_cut(pinned=True), host, "pinned result buffers")


@pytest.mark.gpu
def test_blast_tables_to_orthologs_without_json(engine):
    """BLAST text -> native reader -> GPU cut -> GPU build == the oracle over the reference's own sequence of steps."""
    from oracle import oracle as O
    from oracle import producer as OP
    from ortho2align_b200 import producer, synth
    rng = np.random.default_rng(9)
    n_lnc, regions, flank, gene_len = 12, 3, 500, 900
    tables, q, s, genes = [], [], [], []
    for l in range(n_lnc):
        g0 = int(rng.integers(10_000, 5_000_000))
        for _ in range(regions):
            hits = list(synth.relative_hits(rng, 150, gene_len + 2 * flank, 20_000))
            hits[4] = hits[4] + np.where(rng.random(150) < 0.15, 40, 0)
            tables.append(synth.blast_tabular(*hits))
            s0 = int(rng.integers(10_000, 5_000_000))
            q.append((g0 - flank, g0 + gene_len + flank, "+-."[l % 3])); s.append((s0, s0 + 20_000, "+-."[(l // 3) % 3]))
            genes.append((g0, g0 + gene_len))
    rel = producer.parse_blast_tables(tables, threads=2)
    assert not rel["status"].any()
    want = [OP.parse_blast(t) for t in tables]
    assert np.diff(rel["hsp_off"]).tolist() == [len(w["hsps"]) for w in want]
    one = lambda xs: np.array(xs, dtype=np.int64)
    rg = dict(q_start=one([x[0] for x in q]), q_end=one([x[1] for x in q]), q_minus=np.array([x[2] == "-" for x in q], np.int8),
              s_start=one([x[0] for x in s]), s_end=one([x[1] for x in s]), s_minus=np.array([x[2] == "-" for x in s], np.int8),
              qleft=one([g[0] for g in genes]), qright=one([g[1] for g in genes]))
    cut = producer.cut_alignments(engine, rel, rg)
    exp_cut = OP.cut_batch(rel, rg)
    _same_cut(cut, exp_cut, "tables -> cut")
    assert np.array_equal(cut["score_is_int"], (exp_cut["cut"] == 0).astype(np.uint8))
    aln_off = np.arange(0, n_lnc * regions + 1, regions, dtype=np.int64)
    bg = synth.null_scores(rng, n_lnc * 300).astype(np.int32)
    bg_off = np.arange(0, n_lnc * 300 + 1, 300, dtype=np.int64)
    from helpers import compare_results
    got = engine.build_batch(producer.to_columnar(cut, aln_off, bg_off, bg), "hist", True)
    compare_results(got, O.build_batch(producer.to_columnar(exp_cut, aln_off, bg_off, bg), "hist", True), exact_p=True,
                    what="tables -> orthologs")
