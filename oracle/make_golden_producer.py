"""Golden vectors for the producer side of the alignment JSON, recorded from the
UNMODIFIED reference in this container:

    GenomicRangesAlignment.from_file_blast(text, qrange, srange)      alignment.py:515-586, genomicranges.py:1160-1171
      -> .to_genomic()                                                genomicranges.py:1108-1129
      -> .cut_coordinates(qleft, qright, sleft, sright)               alignment.py:734-803, 267-333; genomicranges.py:1153-1158
      -> .to_dict()                                                   alignment.py:588-601

TEST INFRASTRUCTURE ONLY (runs here, where /root/reference exists). Output: tests/golden/producer.json.

    python oracle/make_golden_producer.py
"""
import io
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh  # noqa: E402
from ortho2align_b200 import synth  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def hsp_rows(hsps):
    """[qstart, qend, sstart, send, score, score_is_float] — JSON keeps int/float apart, so does this."""
    return [[h.qstart, h.qend, h.sstart, h.send, h.score, isinstance(h.score, float)] for h in hsps]


def run_case(ref, text, q, s, bounds):
    GR = ref.genomicranges.GenomicRange
    GRA = ref.genomicranges.GenomicRangesAlignment
    qr = GR(q[0], q[1], q[2], q[3], name="gene")
    sr = GR(s[0], s[1], s[2], s[3])
    out = {"text": text, "qrange": list(q), "srange": list(s), "bounds": bounds}
    try:
        aln = GRA.from_file_blast(io.StringIO(text), qr, sr)
    except Exception as e:  # what _align_syntenies turns into an ExceptionLogger (pipeline.py:545-546)
        out["parse_error"] = type(e).__name__
        return out
    out["parsed"] = {"hsps": hsp_rows(aln._all_HSPs), "qlen": aln.qlen, "slen": aln.slen}
    aln.to_genomic()
    out["genomic"] = hsp_rows(aln._all_HSPs)
    try:
        cut = aln.cut_coordinates(**bounds)
    except Exception as e:
        out["cut_error"] = type(e).__name__
        return out
    d = cut.to_dict()
    assert d["relative"] is False and d["filtered"] is False and d["HSPs"] == d["filtered_HSPs"]
    out["cut"] = {"hsps": hsp_rows(cut._all_HSPs), "qlen": cut.qlen, "slen": cut.slen}
    return out


def main():
    ref = rh.import_reference()
    rng = np.random.default_rng(20240611)
    cases = []
    flank, gene_len = 400, 700
    q_len = gene_len + 2 * flank
    for qstrand in "+-.":
        for sstrand in "+-.":
            gene_start = int(rng.integers(10_000, 1_000_000))
            q = ("chr1", gene_start - flank, gene_start + gene_len + flank, qstrand)
            s0 = int(rng.integers(10_000, 1_000_000))
            s = ("chr2", s0, s0 + 5000, sstrand)
            hits = synth.relative_hits(rng, 60, q_len, 5000)
            text = synth.blast_tabular(*hits)
            # the pipeline's cut (gene bounds on the query side only, pipeline.py:530-531)
            cases.append(run_case(ref, text, q, s, {"qleft": gene_start, "qright": gene_start + gene_len}))
            # all four boundaries, subject window narrowed so that subject-side cuts of both orientations occur
            cases.append(run_case(ref, text, q, s, {"qleft": gene_start + 37, "qright": gene_start + gene_len - 41,
                                                    "sleft": s0 + 900, "sright": s0 + 4100}))
            cases.append(run_case(ref, text, q, s, {"sleft": s0 + 2000}))
            cases.append(run_case(ref, text, q, s, {"sright": s0 + 2500, "qright": gene_start + 300}))
    # boundaries that coincide with HSP ends (the <, <= pairs of alignment.py:760-800), one HSP spanning the
    # whole gene (cut on both sides: the second cut starts from the first cut's rounded coordinates and float score)
    qs = np.array([1, 101, 151, 201, 50, 300, 120, 1]); qe = np.array([100, 150, 200, 260, 400, 340, 180, 500])
    ss = np.array([1001, 1101, 1300, 1201, 2050, 2340, 3120, 4500]); se = np.array([1100, 1150, 1251, 1260, 2400, 2300, 3183, 4001])
    sc = np.array([100, 50, 50, 60, 351, 41, 61, 500])
    text = synth.blast_tabular(qs, qe, ss, se, sc)
    for bounds in ({"qleft": 1100, "qright": 1200}, {"qleft": 1150, "qright": 1150}, {"qleft": 1101, "qright": 1259},
                   {"qleft": 1000, "qright": 1500}, {"qleft": 1260}, {"qright": 1000}, {},
                   {"sleft": 11100, "sright": 12300}, {"sleft": 11250, "sright": 11251}, {"sleft": 14001, "sright": 14500}):
        cases.append(run_case(ref, text, ("chrA", 1000, 1500, "+"), ("chrB", 10000, 15000, "+"), bounds))
        cases.append(run_case(ref, text, ("chrA", 1000, 1500, "-"), ("chrB", 10000, 15000, "-"), bounds))
    # float scores from the table (numberize keeps non-integral floats, utils.py:21-41), a 'query length' column
    text = synth.blast_tabular(qs, qe, ss, se, sc * 1.25,
                               fields=synth.BLAST_FIELDS).replace("\t125.0\n", "\t125\n")
    cases.append(run_case(ref, text, ("chrA", 1000, 1500, "+"), ("chrB", 10000, 15000, "+"), {"qleft": 1100, "qright": 1200}))
    lines = synth.blast_tabular(qs, qe, ss, se, sc, fields=synth.BLAST_FIELDS + ", query length, subject length").split("\n")
    lines = [ln + "\t777\t8888" if ln and not ln.startswith("#") else ln for ln in lines]
    cases.append(run_case(ref, "\n".join(lines), ("chrA", 1000, 1500, "+"), ("chrB", 10000, 15000, "+"), {"qleft": 1100}))
    # no hits; hits table announced but 'score' missing; not a BLAST table; Fields line with zero hits
    cases.append(run_case(ref, synth.blast_tabular(qs[:0], qe[:0], ss[:0], se[:0], sc[:0]),
                          ("chrA", 1000, 1500, "+"), ("chrB", 10000, 15000, "+"), {"qleft": 1100, "qright": 1200}))
    cases.append(run_case(ref, synth.blast_tabular(qs, qe, ss, se, sc, fields=synth.BLAST_FIELDS.replace(", score", "")),
                          ("chrA", 1000, 1500, "+"), ("chrB", 10000, 15000, "+"), {}))
    cases.append(run_case(ref, "query\tsubject\t1\t2\n", ("chrA", 1000, 1500, "+"), ("chrB", 10000, 15000, "+"), {}))
    cases.append(run_case(ref, "\n".join(["# BLASTN", f"# Fields: {synth.BLAST_FIELDS}", "# 0 hits found", ""]),
                          ("chrA", 1000, 1500, "+"), ("chrB", 10000, 15000, "+"), {}))
    out = {"versions": {"python": sys.version.split()[0], "numpy": np.__version__}, "cases": cases}
    path = os.path.join(GOLDEN, "producer.json")
    with open(path, "w") as f:
        json.dump(out, f)
    n_cut = sum(1 for c in cases if "cut" in c)
    n_changed = sum(sum(1 for h in c["cut"]["hsps"] if h[5]) for c in cases if "cut" in c)
    print("producer.json", os.path.getsize(path), "cases", len(cases), "with cut", n_cut, "cut HSPs", n_changed,
          "errors", [c.get("parse_error") or c.get("cut_error") for c in cases if "cut" not in c])


if __name__ == "__main__":
    main()
