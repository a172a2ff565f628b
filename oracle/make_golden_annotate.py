"""Record tests/golden/annotate.json from the UNMODIFIED reference (container-only; TEST INFRASTRUCTURE).

Two kinds of cases:
* "join": random range lists through the reference's own objects — `GenomicRangesList.get_neighbours`
  (genomicranges.py:2040-2067 -> find_neighbours :635-685) with several distances / strandness settings, then
  `calc_JI` / `calc_OC` (:497-547) per neighbour. Neighbour indices are positions in the reference's sorted
  annotation list, stored ascending (identical keys are interchangeable, see oracle/o2a_oracle_annotate.c).
  Some cases hold inverted ranges (end < start): the reference's loops still run on them (they pin the
  "right loop stops at the first miss" behaviour) but `len()` raises, so those carry no JI / OC.
* "files": `pipeline.annotate_orthologs` (:1027-1094) on BED12 orthologs against BED3/6/12, GTF and GFF
  annotations, with and without a name regex, plus the empty / malformed ortholog files that take the
  ParserException branch. Output file and stats block are stored verbatim.

Run: python -m oracle.make_golden_annotate
"""
from __future__ import annotations

import json
import os
import random
import tempfile

from . import ref_harness

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "annotate.json")
CHROMS = ["chr1", "chr2", "chr10", "chrX", "1", "chrUn_gl000220"]


def random_ranges(r, n, prefix, n_chrom, span, inverted=0.0, dup=0.15):
    out, last = [], None
    for i in range(n):
        s = r.randrange(0, span)
        e = s + r.choice([0, 1, 5, 50, 500, r.randrange(1, 3000), r.randrange(1, 40000)])
        if last is not None and r.random() < dup:
            s, e = last
        last = (s, e)
        if r.random() < inverted:
            s, e = e, s
        out.append(dict(chrom=r.choice(CHROMS[:n_chrom]), start=s, end=e, strand=r.choice("+-."),
                        name=f"{prefix}{r.randrange(max(1, n // 3))}"))
    return out


def join_case(ref, seed, n_q, n_a, distance, strandness, inverted=0.0, n_chrom=4, span=60000):
    r = random.Random(seed)
    GR, GRL = ref.genomicranges.GenomicRange, ref.genomicranges.GenomicRangesList
    q = random_ranges(r, n_q, "lnc", n_chrom, span, inverted)
    a = random_ranges(r, n_a, "gene", n_chrom, span, inverted)
    ql = GRL([GR(**d) for d in q], None)
    al = GRL([GR(**d) for d in a], None)
    pos = {id(g): k for k, g in enumerate(al)}
    ql.get_neighbours(al, distance=distance, strandness=strandness, relation="nb")
    nb, ji, oc = [], [], []
    for g in ql:                                         # sorted order, like annotate_orthologs iterates
        found = list(g.relations["nb"])
        idx = sorted(pos[id(x)] for x in found)
        nb.append(idx)
        if inverted == 0.0:
            by = {pos[id(x)]: x for x in found}
            ji.append([float(g.calc_JI(by[k])) for k in idx])
            oc.append([float(g.calc_OC(by[k])) for k in idx])
    case = dict(seed=seed, distance=distance, strandness=strandness, q=q, a=a, nb=nb,
                q_sorted=[[g.chrom, g.start, g.end, g.name] for g in ql],
                a_sorted=[[g.chrom, g.start, g.end, g.name] for g in al])
    if inverted == 0.0:
        case["ji"], case["oc"] = ji, oc
    return case


def file_texts(seed, n_q, n_a, fmt):
    r = random.Random(seed)
    q = random_ranges(r, n_q, "lnc", 5, 60000)
    a = random_ranges(r, n_a, "gene", 6, 60000)
    orth = "".join(f"{d['chrom']}\t{d['start']}\t{d['end']}\t{d['name']}\t{r.randrange(1000)}\t{d['strand']}\t"
                   f"{d['start']}\t{d['end']}\t0\t1\t{d['end'] - d['start']}\t0\n" for d in q)
    lines = ["# a comment line\n"] if fmt in ("gtf", "gff") else []
    for d in a:
        c, s, e, st, nm = d["chrom"], d["start"], d["end"], d["strand"], d["name"]
        if fmt == "bed3":
            lines.append(f"{c}\t{s}\t{e}\n")
        elif fmt == "bed6":
            lines.append(f"{c}\t{s}\t{e}\t{nm}\t0\t{st}\n")
        elif fmt == "bed12":
            lines.append(f"{c}\t{s}\t{e}\t{nm}\t0\t{st}\t{s}\t{e}\t0\t1\t{e - s}\t0\n")
        elif fmt == "gtf":
            lines.append(f'{c}\tsrc\texon\t{s + 1}\t{e}\t.\t{st}\t.\tgene_id "{nm}"; transcript_id "{nm}.1";\n')
        else:
            lines.append(f"{c}\tsrc\texon\t{s + 1}\t{e}\t.\t{st}\t.\tID={nm};Parent=p{nm}\n")
    return orth, "".join(lines)


def file_case(ref, orth, ann, fmt, regex=None, precision=8):
    with tempfile.TemporaryDirectory() as tmp:
        po, pa = os.path.join(tmp, "orthologs.bed"), os.path.join(tmp, "annotation." + fmt)
        out, stats = os.path.join(tmp, "out.tsv"), os.path.join(tmp, "stats.txt")
        open(po, "w").write(orth)
        open(pa, "w").write(ann)
        ref.pipeline.annotate_orthologs(po, pa, out, subject_name_regex=regex, stats_filename=stats,
                                        float_precision=precision)
        return dict(fmt=fmt, regex=regex, precision=precision, orthologs=orth, annotation=ann,
                    out=open(out).read(), stats=open(stats).read().replace(po, "{orthologs}"))


def main():
    ref = ref_harness.import_reference()
    joins = []
    k = 0
    for n_q, n_a in ((0, 5), (5, 0), (1, 1), (40, 60), (150, 400), (300, 120)):
        for distance, strandness in ((0, True), (0, False), (700, True)):
            joins.append(join_case(ref, 100 + k, n_q, n_a, distance, strandness))
            k += 1
    joins.append(join_case(ref, 200, 120, 300, 0, True, n_chrom=1, span=8000))          # one dense chromosome
    joins.append(join_case(ref, 201, 120, 300, 2500, False, n_chrom=2, span=20000))
    joins.append(join_case(ref, 202, 150, 250, 0, True, inverted=0.1))                  # inverted ranges
    joins.append(join_case(ref, 203, 150, 250, 300, False, inverted=0.25, n_chrom=2, span=15000))
    files = []
    for i, fmt in enumerate(("bed3", "bed6", "bed12", "gtf", "gff")):
        orth, ann = file_texts(300 + i, 80, 200, fmt)
        files.append(file_case(ref, orth, ann, fmt))
    orth, ann = file_texts(310, 60, 150, "gtf")
    files.append(file_case(ref, orth, ann, "gtf", regex=r'gene_id "(.*?)";', precision=3))
    orth, ann = file_texts(311, 60, 150, "gff")
    files.append(file_case(ref, orth, ann, "gff", regex=r"ID=(gene1.*?);"))            # some names fall back to the header
    files.append(file_case(ref, "", ann, "gff"))                                        # no orthologs: EmptyAnnotation branch
    files.append(file_case(ref, "chr1\t10\tnot_a_number\n", ann, "gff"))                # UnrecognizedFormat branch
    files.append(file_case(ref, orth + "chr1\t5\n", ann, "gff"))                        # IncorrectLine branch
    with open(OUT, "w") as f:
        json.dump(dict(made_by="oracle/make_golden_annotate.py", joins=joins, files=files), f, separators=(",", ":"))
    print(OUT, os.path.getsize(OUT), "bytes;", len(joins), "join cases,", len(files), "file cases")


if __name__ == "__main__":
    main()
