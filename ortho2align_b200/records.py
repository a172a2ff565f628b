"""Host-side record types and text formatters of the build_orthologs outputs.

Python restatement (no code shared with the reference) of the few pieces of
`ortho2align/genomicranges.py` the hot path's OUTPUT side needs:
  GenomicRange.from_dict / name / to_bed6        genomicranges.py:386-437, 796-855
  BaseGenomicRange.distance / find_neighbours    genomicranges.py:439-458, 635-685, 865-881
  sorted list key, to_bed6, drop_duplicates      genomicranges.py:1566-1572, 1839-1865
  GenomicRangesAlignmentChain._prepare_side / to_record / to_bed12   genomicranges.py:1192-1301
  nxor_strands                                   alignment.py:48-64
Text must be byte-identical to the reference's, so values are formatted with the same `str()`
calls on the same Python / numpy types (ints stay ints, q-values are numpy float64 scalars).
"""
from __future__ import annotations

import bisect

import numpy as np


class GenomicException(Exception):
    pass


class Range:
    """Minimal mirror of the reference's GenomicRange (only what the output side touches)."""
    __slots__ = ("chrom", "start", "end", "strand", "_name", "genome", "lifted")

    def __init__(self, chrom, start, end, strand=".", name=None, genome=None, lifted=None):
        self.chrom, self.start, self.end, self.strand = chrom, start, end, strand
        self._name, self.genome, self.lifted = name, genome, lifted

    @classmethod
    def from_dict(cls, d):
        # genomicranges.py:816-855: all eight keys are mandatory
        for key in ("chrom", "start", "end", "strand", "genome", "name", "sequence_file_path", "relations"):
            if key not in d:
                raise ValueError(f"Provided dict_ does not contain {key} key.")
        lifted = None
        rel = d.get("relations") or {}
        if "lifted" in rel:
            coll = rel["lifted"].get("collection")
            if coll is None:
                raise ValueError("There is no key `{key}` in the provided dictionary.")
            lifted = sorted((cls.from_dict(g) for g in coll), key=sort_key)
        return cls(d["chrom"], d["start"], d["end"], d["strand"], d["name"], d["genome"], lifted)

    @property
    def name(self):
        # genomicranges.py:427-433: falls back to the sequence header
        if self._name is None:
            return f"{self.chrom}:{self.start}-{self.end}{self.strand}"
        return self._name

    @name.setter
    def name(self, value):
        self._name = value

    def same(self, other):
        return (self.chrom == other.chrom and self.start == other.start and self.end == other.end
                and self.strand == other.strand and self.name == other.name and self.genome == other.genome)

    def distance(self, other):
        # genomicranges.py:865-881 + 439-458
        if self.genome != other.genome:
            raise GenomicException("inconsistent genomes")
        if self.chrom != other.chrom:
            raise GenomicException("inconsistent chromosomes")
        if self.end < other.start:
            return other.start - self.end
        if other.end < self.start:
            return other.end - self.start
        return 0

    def __str__(self):
        # genomicranges.py:772-781
        return "\t".join(str(v) for v in (self.chrom, self.start, self.end, self.name, self.strand, self.genome))


def sort_key(r):
    """BaseGenomicRangesList._genomic_key_func (genomicranges.py:1566-1572)."""
    return (r.chrom, r.start, r.end, r.name)


def find_neighbours(grange, sorted_list, distance=0):
    """BaseGenomicRange.find_neighbours (genomicranges.py:635-685) on a list sorted by `sort_key`."""
    keys = [sort_key(r) for r in sorted_list]
    index = bisect.bisect_left(keys, sort_key(grange))
    out = []
    left = index - 1
    try:
        while left >= 0:
            if abs(grange.distance(sorted_list[left])) <= distance:
                out.append(sorted_list[left])
            left -= 1
    except GenomicException:
        pass
    right = index
    try:
        while abs(grange.distance(sorted_list[right])) <= distance:
            out.append(sorted_list[right])
            right += 1
    except (GenomicException, IndexError):
        pass
    return sorted(out, key=sort_key)


def drop_duplicates(sorted_ranges):
    """BaseGenomicRangesList.drop_duplicates (genomicranges.py:1853-1865)."""
    if len(sorted_ranges) <= 1:
        return list(sorted_ranges)
    out = [sorted_ranges[0]]
    for r in sorted_ranges[1:]:
        if r.same(out[-1]) and r.lifted == out[-1].lifted:
            continue
        out.append(r)
    return out


def bed6_lines(ranges):
    """BaseGenomicRangesList.to_bed6 (genomicranges.py:1839-1851)."""
    return ["\t".join([str(r.chrom), str(r.start), str(r.end), str(r.name), ".", str(r.strand)]) + "\n"
            for r in ranges]


def nxor_strands(hsp_strand, alignment_strand):
    """alignment.py:48-64"""
    return "+" if hsp_strand == (alignment_strand in ["+", "."]) else "-"


def chain_records(hsps, pvals, qrange, srange_chrom, name):
    """The four strings `_build` keeps per non-empty chain (pipeline.py:784-787):
    (query tsv, query bed12, subject tsv, subject bed12).

    hsps: the chain's HSP dicts (reference JSON objects: ints stay ints) in chain order;
    pvals: hsp.pval of each (numpy float64 scalars, as `_build` assigns them, pipeline.py:772,779).
    Restates `_prepare_side` (genomicranges.py:1192-1243): both sides take their strand from
    HSPs[0] XNOR the QUERY range's strand; the subject side on '-' uses `send` and reverses blocks.
    """
    qs = [h["qstart"] for h in hsps]; qe = [h["qend"] for h in hsps]
    ss = [h["sstart"] for h in hsps]; se = [h["send"] for h in hsps]
    score = sum([h["score"] for h in hsps])
    h0 = hsps[0]
    q_strand_flag = h0["qend"] - h0["qstart"] > 0
    s_strand_flag = h0["send"] - h0["sstart"] > 0
    sides = []
    # query side
    c0, c1 = min(min(qs), min(qe)), max(max(qs), max(qe))
    sides.append((qrange.chrom, c0, c1, name, score, nxor_strands(q_strand_flag, qrange.strand), c0, c1, "0",
                  len(hsps), [abs(e - s) for s, e in zip(qs, qe)], [s - c0 for s in qs], list(pvals)))
    # subject side
    c0, c1 = min(min(ss), min(se)), max(max(ss), max(se))
    strand = nxor_strands(s_strand_flag, qrange.strand)
    sizes = [abs(e - s) for s, e in zip(ss, se)]
    pv = list(pvals)
    if strand == "+":
        starts = [s - c0 for s in ss]
    else:
        starts = [e - c0 for e in se]
        sizes.reverse(); starts.reverse(); pv.reverse()
    sides.append((srange_chrom, c0, c1, name, score, strand, c0, c1, "0", len(hsps), sizes, starts, pv))

    def join(side):
        """(13-column TSV line, 12-column BED12 line): the TSV is the BED12 plus the p/q-value list"""
        bed12 = "\t".join((",".join([str(v) for v in item]) if isinstance(item, list) else str(item)) for item in side[:12])
        return bed12 + "\t" + ",".join([str(v) for v in side[12]]), bed12
    q, s = join(sides[0]), join(sides[1])
    return q[0], q[1], s[0], s[1]


def as_pvals(values):
    """q/p values as the numpy float64 scalars the reference stores in hsp.pval."""
    return [np.float64(v) for v in values]
