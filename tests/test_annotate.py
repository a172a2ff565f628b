"""annotate_orthologs' interval join (pipeline.py:1027-1094, genomicranges.py:635-685,
497-547): the oracle against the goldens recorded from the reference (CPU), the host module with the oracle
injected (CPU), and the CUDA path through the C-ABI against the oracle and the goldens (gpu)."""
import os

import numpy as np
import pytest

from conftest import load_golden

STR = {"+": 1, "-": -1, ".": 0}


def _ranges(rows):
    from ortho2align_b200.annotate import Ranges
    return Ranges([d["chrom"] for d in rows], [d["start"] for d in rows], [d["end"] for d in rows],
                  [STR[d["strand"]] for d in rows], [d["name"] for d in rows], "bed6")


def _expected(case):
    nb = case["nb"]
    off = np.zeros(len(nb) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(x) for x in nb])
    idx = np.array([k for x in nb for k in x], dtype=np.int32)
    ji = oc = None
    if "ji" in case:
        ji = np.array([v for x in case["ji"] for v in x], dtype=np.float64)
        oc = np.array([v for x in case["oc"] for v in x], dtype=np.float64)
    return off, idx, ji, oc


def _check_join(case, run):
    from ortho2align_b200 import annotate as A
    qr, ar = _ranges(case["q"]), _ranges(case["a"])
    q, a, q_order, a_order = A.columnar(qr, ar)
    # the columnar sort is the reference's list order
    assert [[qr.chrom[i], int(qr.start[i]), int(qr.end[i]), qr.name[i]] for i in q_order] == case["q_sorted"]
    assert [[ar.chrom[i], int(ar.start[i]), int(ar.end[i]), ar.name[i]] for i in a_order] == case["a_sorted"]
    got = run(q, a, case["distance"], case["strandness"])
    off, idx, ji, oc = _expected(case)
    what = f"seed {case['seed']} d={case['distance']} s={case['strandness']}"
    assert np.array_equal(got["nb_off"], off), what
    assert np.array_equal(got["nb_idx"], idx), what
    if ji is not None:
        assert np.array_equal(got["ji"], ji) and np.array_equal(got["oc"], oc), what     # float64, bit for bit


def _random_columns(seed, n_q, n_a, n_chrom=24, span=2_000_000, inverted=0.0, long_frac=0.01):
    from ortho2align_b200 import synth
    return synth.annotate_workload(n_q, n_a, seed=seed, n_chrom=n_chrom, span=span, inverted=inverted, long_frac=long_frac)


# ---------------------------------------------------------------------------------------------------------
# CPU: oracle pinned to the goldens; host module (parsers, sort, text) with the oracle injected
# ---------------------------------------------------------------------------------------------------------
def test_oracle_matches_the_reference_goldens():
    from oracle import annotate as OA
    g = load_golden("annotate.json")
    assert len(g["joins"]) >= 20
    for case in g["joins"]:
        _check_join(case, lambda q, a, d, s: OA.annotate_batch(q, a, d, s))


class _OracleEngine:
    def annotate_batch(self, q, a, distance=0, strandness=True):
        from oracle import annotate as OA
        return OA.annotate_batch(q, a, distance, strandness)

    def close(self):
        pass


def _check_files(tmp_path, engine):
    from ortho2align_b200 import annotate as A
    g = load_golden("annotate.json")
    for i, case in enumerate(g["files"]):
        po, pa = tmp_path / f"orthologs{i}.bed", tmp_path / f"annotation{i}.{case['fmt']}"
        po.write_text(case["orthologs"]); pa.write_text(case["annotation"])
        out, stats = tmp_path / f"out{i}.tsv", tmp_path / f"stats{i}.txt"
        A.annotate_orthologs(str(po), str(pa), str(out), subject_name_regex=case["regex"], stats_filename=str(stats),
                             float_precision=case["precision"], engine=engine)
        assert out.read_text() == case["out"], f"file case {i} ({case['fmt']})"
        assert stats.read_text() == case["stats"].replace("{orthologs}", str(po)), f"file case {i} stats"


def test_host_module_reproduces_the_reference_files(tmp_path):
    _check_files(tmp_path, _OracleEngine())


def test_parser_errors_follow_the_reference():
    from ortho2align_b200 import annotate as A
    with pytest.raises(A.EmptyAnnotation):
        A.parse_annotation([])
    with pytest.raises(A.EmptyAnnotation):
        A.parse_annotation(["# only\n", "# comments\n"])
    with pytest.raises(A.UnrecognizedFormat):
        A.parse_annotation(["chr1\t1\n"])
    with pytest.raises(A.UnrecognizedFormat):
        A.parse_annotation(["track name=x\n", "chr1\t1\t2\n"])            # the sniffer does not skip track lines
    with pytest.raises(A.UnrecognizedFormat):
        A.parse_annotation(["chr1\t1\t2\tn\t0\t*\n"])                     # bad strand
    with pytest.raises(A.IncorrectLine) as e:
        A.parse_annotation(["chr1\t1\t2\n", "\n"])
    assert e.value.line_no == 2
    with pytest.raises(A.IncorrectLine):
        A.parse_annotation(["chr1\t1\t9\tn\t0\t+\t1\t9\t0\t2\t4,\t0,\n"])   # blockCount != len(blockSizes)
    with pytest.raises(ValueError):
        A.parse_annotation(['chr1\ts\te\t1\t2\t.\t+\t.\tgene_id "a";\n'], name_regex="no_group")
    r = A.parse_annotation(["#c\n", "chr1\t1\t2\n", "track x\n", "chr1\t3\t9\n"])
    assert r.fmt == "bed3" and r.name == ["chr1:1-2.", "chr1:3-9."] and list(r.strand) == [0, 0]
    r = A.parse_annotation(['chr1\ts\te\t5\t9\t.\t-\t.\tgene_id "a"; x "y";\n'], name_regex='gene_id "(.*?)"')
    assert (r.fmt, r.name, int(r.start[0]), int(r.end[0]), int(r.strand[0])) == ("gtf", ["a"], 4, 9, -1)
    r = A.parse_annotation(["chr1\ts\te\t5\t9\t.\t-\t.\tID=a;P=b\n"], name_regex="Name=(.*)")
    assert r.fmt == "gff" and r.name == ["chr1:4-9-"]
    with pytest.raises(OverflowError):
        A.columnar(A.Ranges(["c"], [0], [2 ** 31], [0], ["n"], "bed3"), A.Ranges(["c"], [0], [1], [0], ["m"], "bed3"))


def test_oracle_on_synthetic_columns_is_consistent():
    """The literal O(chromosome) left loop against a brute-force all-pairs restatement (well-formed ranges)."""
    from oracle import annotate as OA
    q, a = _random_columns(5, 300, 2000, n_chrom=3, span=50_000)
    got = OA.annotate_batch(q, a, 0, True)
    for i in range(300):
        same = (a["chrom"] == q["chrom"][i]) & (a["start"] <= q["end"][i]) & (a["end"] >= q["start"][i])
        same &= (q["strand"][i] == 0) | (a["strand"] == q["strand"][i]) | (a["strand"] == 0)
        assert np.array_equal(np.nonzero(same)[0], got["nb_idx"][got["nb_off"][i]:got["nb_off"][i + 1]])


# ---------------------------------------------------------------------------------------------------------
# GPU: the CUDA path through the C-ABI
# ---------------------------------------------------------------------------------------------------------
def _same(got, exp, what):
    for k in ("nb_off", "nb_idx", "ji", "oc"):
        assert np.array_equal(got[k], exp[k]), f"{what}: {k}"


@pytest.mark.gpu
def test_gpu_join_matches_the_reference_goldens(engine):
    for case in load_golden("annotate.json")["joins"]:
        _check_join(case, lambda q, a, d, s: engine.annotate_batch(q, a, d, s))


@pytest.mark.gpu
def test_gpu_files_match_the_reference(engine, tmp_path):
    _check_files(tmp_path, engine)


@pytest.mark.gpu
@pytest.mark.parametrize("n_q,n_a,n_chrom,span,inverted,distance,strandness", [
    (0, 0, 1, 1000, 0.0, 0, True), (0, 100, 2, 10_000, 0.0, 0, True), (100, 0, 2, 10_000, 0.0, 0, True),
    (1, 1, 1, 100, 0.0, 0, False), (31, 1023, 1, 30_000, 0.0, 0, True), (33, 1025, 2, 30_000, 0.0, 5, False),
    (5000, 20_000, 24, 2_000_000, 0.0, 0, True), (5000, 20_000, 24, 2_000_000, 0.0, 10_000, True),
    (4000, 9000, 3, 200_000, 0.1, 0, True), (4000, 9000, 3, 200_000, 0.3, 250, False),
    (3000, 40_000, 1, 1_000_000, 0.0, 0, True),            # one dense chromosome: long neighbour lists
    (60_000, 250_000, 40, 5_000_000, 0.0, 0, True),
])
def test_gpu_join_matches_the_oracle(engine, n_q, n_a, n_chrom, span, inverted, distance, strandness):
    from oracle import annotate as OA
    q, a = _random_columns(n_q * 7 + n_a, n_q, n_a, n_chrom, span, inverted)
    got = engine.annotate_batch(q, a, distance, strandness)
    _same(got, OA.annotate_batch(q, a, distance, strandness), f"{n_q}x{n_a} d={distance}")
    assert engine.launch_count() > 0


@pytest.mark.gpu
def test_gpu_join_extreme_coordinates(engine):
    """int32 limits on both sides, a distance that overflows int32 arithmetic, and identical keys."""
    from oracle import annotate as OA
    lo, hi = -(2 ** 31), 2 ** 31 - 1
    a = dict(chrom=np.array([0, 0, 0, 0, 1, 1], np.int32), start=np.array([lo, lo, -5, hi - 1, lo, 7], np.int32),
             end=np.array([lo, hi, 5, hi, -1, 7], np.int32), name=np.array([0, 1, 2, 3, 0, 0], np.int32),
             strand=np.array([1, -1, 0, 1, -1, 1], np.int8))
    q = dict(chrom=np.array([0, 0, 0, 1, 1, 2], np.int32), start=np.array([lo, -5, hi, 7, 7, 0], np.int32),
             end=np.array([lo, 5, hi, 7, 7, 9], np.int32), name=np.array([0, 2, 9, 0, 1, 0], np.int32),
             strand=np.array([0, 1, -1, 1, -1, 0], np.int8))
    for distance in (0, 1, 2 ** 31, 2 ** 33):
        for strandness in (True, False):
            _same(engine.annotate_batch(q, a, distance, strandness), OA.annotate_batch(q, a, distance, strandness),
                  f"extreme d={distance}")


@pytest.mark.gpu
def test_gpu_join_inverted_query_left_neighbour(engine):
    """An inverted ortholog (end < start) reaches a LEFT range through distance()'s first branch (self.end < start)
    although that range ends far below self.start - d: the candidate bound must not cut it off."""
    from oracle import annotate as OA
    i32, i8 = (lambda *v: np.array(v, np.int32)), (lambda *v: np.array(v, np.int8))
    a = dict(chrom=i32(0, 0, 0), start=i32(10, 55, 58), end=i32(20, 70, 59), name=i32(0, 1, 2), strand=i8(0, 0, 0))
    q = dict(chrom=i32(0, 0), start=i32(100, 100), end=i32(50, 100), name=i32(0, 0), strand=i8(0, 0))
    for d in (0, 10, 30, 100):
        exp = OA.annotate_batch(q, a, d, True)
        _same(engine.annotate_batch(q, a, d, True), exp, f"inverted d={d}")
    assert OA.annotate_batch(q, a, 10, True)["nb_idx"].tolist() == [1, 2]


@pytest.mark.gpu
def test_gpu_join_device_resident_inputs(engine):
    import torch
    from oracle import annotate as OA
    q, a = _random_columns(77, 20_000, 50_000)
    dq = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in q.items()}
    da = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in a.items()}
    engine.run_annotate(engine.ann_struct(dq, da, 0, True, device=True))
    _same(engine.fetch_annotation(pinned=True), OA.annotate_batch(q, a, 0, True), "device inputs")
    engine.set_timing(True)
    engine.run_annotate(engine.ann_struct(dq, da, 0, True, device=True))
    ms = engine.annotate_ms()
    engine.set_timing(False)
    assert set(ms) == {"h2d", "prefmax", "count", "write"} and all(v >= 0 for v in ms.values())


@pytest.mark.gpu
@pytest.mark.parametrize("block", range(4))
def test_gpu_join_random_sweep(engine, block):
    """Randomised shapes: tiny to mid-size lists, 1-6 chromosomes, dense and sparse spans, inverted ranges, all
    distance / strandness settings — every batch bit-exact against the oracle."""
    from oracle import annotate as OA
    rng = np.random.default_rng(9000 + block)
    for it in range(60):
        n_q = int(rng.choice([0, 1, 2, 31, 32, 33, 100, 257, 1500]))
        n_a = int(rng.choice([0, 1, 2, 31, 32, 33, 64, 65, 1023, 1024, 1025, 5000]))
        n_chrom = int(rng.integers(1, 7))
        span = int(rng.choice([50, 1000, 30_000, 1_000_000]))
        inverted = float(rng.choice([0.0, 0.0, 0.05, 0.5]))
        distance = int(rng.choice([0, 0, 1, 10, 500, 100_000]))
        strandness = bool(rng.integers(0, 2))
        q, a = _random_columns(int(rng.integers(1 << 30)), n_q, n_a, n_chrom, span, inverted,
                               long_frac=float(rng.choice([0.0, 0.01, 0.2])))
        _same(engine.annotate_batch(q, a, distance, strandness), OA.annotate_batch(q, a, distance, strandness),
              f"block {block} it {it}: {n_q}x{n_a} chrom={n_chrom} span={span} inv={inverted} d={distance} s={strandness}")


def test_columnar_order_is_pythons_tuple_order():
    """Ranks + lexsort must order exactly like the reference's SortedKeyList key (chrom, start, end, name) with
    Python's str comparison (code points: 'Z' < 'a', 'chr10' < 'chr2', non-ASCII last) and be stable for equal keys."""
    import random
    from ortho2align_b200 import annotate as A
    r = random.Random(3)
    chroms = ["chr1", "chr10", "chr2", "chrX", "Chr1", "1", "10", "2", "chrUn_é", "", "chr1 "]
    names = ["a", "B", "b", "a1", "a10", "a2", "", "é", "Z", "_"]
    rows = [(r.choice(chroms), r.randrange(5), r.randrange(5), r.choice(names)) for _ in range(400)]
    rg = A.Ranges([x[0] for x in rows], [x[1] for x in rows], [x[2] for x in rows], [0] * len(rows),
                  [x[3] for x in rows], "bed6")
    other = A.Ranges(["chr3", "chr1"], [1, 2], [3, 4], [0, 0], ["zz", "a"], "bed6")   # widens both dictionaries
    q, a, q_order, a_order = A.columnar(rg, other)
    assert q_order.tolist() == sorted(range(len(rows)), key=lambda i: rows[i])          # sorted() is stable too
    assert a_order.tolist() == [1, 0]
    # ranks are order-isomorphic to the strings across BOTH lists
    assert (q["chrom"][0] <= q["chrom"][-1]) and set(a["chrom"].tolist()) <= set(range(len(set(chroms) | {"chr3"})))
    k = {c: int(v) for c, v in zip([rows[i][0] for i in q_order], q["chrom"])}
    k.update({c: int(v) for c, v in zip([other.chrom[i] for i in a_order], a["chrom"])})
    ordered = sorted(k)
    assert [k[c] for c in ordered] == sorted(k.values())
