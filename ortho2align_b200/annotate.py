"""`annotate_orthologs` (pipeline.py:1027-1094) with the interval join on the GPU.

Same signature, output file and stats block as the reference. What moves to the device is the work per
(ortholog, annotation range): `GenomicRangesList.get_neighbours` -> `find_neighbours`
(genomicranges.py:2040-2067, 635-685) and `calc_JI` / `calc_OC` (:497-547) — one `o2a_annotate_batch` call
(include/o2a.h) for the whole ortholog list. Text parsing and formatting stay on the host, like the
reference's, restated from parsing.py (sniffer :944-1007; BED3/6/12, GTF, GFF parsers :415-530, 648-841; only
the five fields the join and the report use are kept: chrom, start, end, strand, name).
There is no CPU fallback: without the CUDA library / a GPU, `Engine` raises.
"""
from __future__ import annotations

import re
from collections import Counter, defaultdict

import numpy as np

STRANDS = {"+": 1, "-": -1, ".": 0}
_INT32 = (-(2 ** 31), 2 ** 31 - 1)


# ---- the reference's parser exceptions (parsing.py:6-266), by name ------------------------------------
class ParserException(Exception):
    pass


class EmptyAnnotation(ParserException):
    def __str__(self):
        return "The annotation is empty."


class UnrecognizedFormat(ParserException):
    def __init__(self, fields):
        self.fields = fields
        super().__init__(f"Can not recognize annotation format with {fields} fields.")


class IncorrectLine(ParserException):
    def __init__(self, line, line_no):
        self.line, self.line_no = line, line_no
        super().__init__(f"Line {line_no} is incorrect:\n{line}")


# ---- field parsers (parsing.py:270-412, 555-596, 599-620, 735-757) ---------------------------------------
def _strand(f):
    if f in ("+", "-", "."):
        return f
    raise ValueError(f)


def _score(f):
    return f                                             # parse_score never raises; the value is not used here


def _item_rgb(f):
    if f in ("0", "."):
        return f
    rec = tuple(int(x) for x in f.split(","))
    if len(rec) != 3:
        raise ValueError(f)
    return rec


def _blocks(f):
    return tuple(int(i) for i in f.split(",") if i != "")


def _phase(f):
    if f == "." or f in ("0", "1", "2"):
        return f
    raise ValueError(f)


def _gtf_attrs(f):
    data = {tag: value.strip('"') for tag, value in (item.strip().split(" ", 1) for item in f.split(";")[:-1])}
    if len(data) == 0:
        raise ValueError(f)
    return data


def _gff_attrs(f):
    data = {tag: value for tag, value in (item.split("=") for item in f.split(";"))}
    if len(data) == 0:
        raise ValueError(f)
    return data


_BED = [str, int, int, str, _score, _strand, int, int, _item_rgb, int, _blocks, _blocks]
_GTF_FAST = [str, str, str, int, int, _score, _strand, _phase, str]        # gtf_fields_fast == gff_fields_fast here
_CHECK_NAME_REGEX = re.compile(r".*\(.*\).*")                             # parsing.py:552


def _try(funcs, record):
    try:
        for f, item in zip(funcs, record):
            f(item)
        return True
    except Exception:
        return False


def sniff(lines):
    """annotation_sniffer (parsing.py:944-1007) on the list of lines (with their newlines)."""
    line = ""
    for line in lines:
        if not line.startswith("#"):
            break
    else:
        line = ""
    if line == "":
        raise EmptyAnnotation
    record = line.strip().split("\t")
    n = len(record)
    if n in (3, 6, 12):
        if _try(_BED[:n], record):
            return f"bed{n}"
        raise UnrecognizedFormat(n)
    if n == 9:
        if _try(_GTF_FAST[:8] + [_gtf_attrs], record):
            return "gtf"
        if _try(_GTF_FAST[:8] + [_gff_attrs], record):
            return "gff"
        raise UnrecognizedFormat(9)
    raise UnrecognizedFormat(n)


class Ranges:
    """Columnar genomic ranges in file order: chrom / name (lists of str), start / end (int64), strand (int8)."""

    def __init__(self, chrom, start, end, strand, name, fmt):
        self.chrom, self.name, self.fmt = chrom, name, fmt
        self.start = np.asarray(start, dtype=np.int64)
        self.end = np.asarray(end, dtype=np.int64)
        self.strand = np.asarray(strand, dtype=np.int8)

    def __len__(self):
        return len(self.chrom)


def parse_annotation(path_or_lines, name_regex=None):
    """parse_annotation (parsing.py:1010-1044) -> Ranges. Raises the reference's ParserException family
    (EmptyAnnotation, UnrecognizedFormat, IncorrectLine) and ValueError for a name_regex without a group."""
    if isinstance(path_or_lines, str):
        with open(path_or_lines, "r") as f:
            lines = f.readlines()
    else:
        lines = list(path_or_lines)
    fmt = sniff(lines)
    chrom, start, end, strand, name = [], [], [], [], []
    if fmt.startswith("bed"):
        n = int(fmt[3:])
        funcs = _BED[:n]
        for line_no, line in enumerate(lines):
            if line.startswith("#") or line.startswith("track"):
                continue
            try:
                data = line.strip().split("\t")
                if len(data) != n:
                    raise ValueError(line)
                rec = [f(item) for f, item in zip(funcs, data)]
                if n == 12 and (len(rec[10]) != rec[9] or len(rec[11]) != rec[9]):
                    raise ValueError(line)
            except Exception:
                raise IncorrectLine(line, line_no + 1)
            chrom.append(rec[0]); start.append(rec[1]); end.append(rec[2])
            st = rec[5] if n >= 6 else "."
            strand.append(STRANDS[st])
            # GenomicRange.name (genomicranges.py:427-432): the sequence header stands in for a missing name
            name.append(rec[3] if n >= 6 else f"{rec[0]}:{rec[1]}-{rec[2]}{st}")
    else:
        if name_regex is None:
            get_name = lambda i: i
        elif _CHECK_NAME_REGEX.match(name_regex) is None:
            raise ValueError
        else:
            pattern = re.compile(name_regex)

            def get_name(i):
                m = pattern.search(i)
                return None if m is None else m[1]
        for line_no, line in enumerate(lines):
            if line.startswith("#"):
                continue
            try:
                data = line.strip().split("\t")
                if len(data) != 9:
                    raise ValueError(line)
                rec = [f(item) for f, item in zip(_GTF_FAST, data)]
                nm = get_name(rec[8])
            except Exception:
                raise IncorrectLine(line, line_no + 1)
            s, e = rec[3] - 1, rec[4]                    # parse_gtf_start / parse_gff_start: 1-based inclusive -> 0-based
            chrom.append(rec[0]); start.append(s); end.append(e); strand.append(STRANDS[rec[6]])
            name.append(nm if nm is not None else f"{rec[0]}:{s}-{e}{rec[6]}")
    return Ranges(chrom, start, end, strand, name, fmt)


def simple_hist(data):
    """utils.py:89-103."""
    counts = Counter(data)
    return "\n".join(f"{k}: {v}" for k, v in sorted(counts.items(), key=lambda i: (i[0], i[1])))


def _ranks(*lists):
    """One dictionary over several lists of str: rank order == Python's str order."""
    table = {s: i for i, s in enumerate(sorted(set().union(*lists)))}
    return [np.fromiter((table[s] for s in lst), dtype=np.int32, count=len(lst)) for lst in lists]


def columnar(orthologs: Ranges, annotation: Ranges):
    """-> (q, a, q_order, a_order): the two lists as the int32 / int8 columns of o2a_ann_input, each sorted by the
    reference's list key (chrom, start, end, name) (genomicranges.py:1566-1572; stable like SortedKeyList's sort),
    and the permutations back into file order."""
    for r in (orthologs, annotation):
        if len(r) and (min(r.start.min(), r.end.min()) < _INT32[0] or max(r.start.max(), r.end.max()) > _INT32[1]):
            raise OverflowError("coordinate outside int32: library limit of o2a_annotate_batch")
    qc, ac = _ranks(orthologs.chrom, annotation.chrom)
    qn, an = _ranks(orthologs.name, annotation.name)
    out = []
    for r, c, n in ((orthologs, qc, qn), (annotation, ac, an)):
        order = np.lexsort((n, r.end, r.start, c))       # last key is the primary one; lexsort is stable
        out.append((dict(chrom=c[order], start=r.start[order].astype(np.int32), end=r.end[order].astype(np.int32),
                         name=n[order], strand=r.strand[order]), order))
    (q, q_order), (a, a_order) = out
    return q, a, q_order, a_order


def annotate_ranges(engine, orthologs: Ranges, annotation: Ranges, distance=0, strandness=True):
    """get_neighbours + JI/OC on the device. -> (q_order, a_order, result dict of o2a_fetch_annotation)."""
    q, a, q_order, a_order = columnar(orthologs, annotation)
    res = engine.annotate_batch(q, a, distance=distance, strandness=strandness)
    return q_order, a_order, res


def annotate_orthologs(subject_orthologs, subject_annotation, output, subject_name_regex=None,
                       stats_filename=None, float_precision=8, device=0, engine=None):
    """Drop-in for pipeline.annotate_orthologs (:1027-1094); `device` / `engine` are the only additions."""
    subject_orthologs_filename = subject_orthologs
    try:
        orthologs = parse_annotation(subject_orthologs_filename)
    except ParserException:
        with open(output, "w") as outfile:
            outfile.write("Query\tOrthologs\n")
        stats_msg = "-----------------------\n" \
                    "annotate_orthologs stats:\n" \
                    f"Recieved 0 orthologs from {subject_orthologs_filename}.\n" \
                    "-----------------------"
    else:
        annotation = parse_annotation(subject_annotation, name_regex=subject_name_regex)
        own = engine is None
        if own:
            from .engine import Engine
            engine = Engine(device)
        try:
            q_order, a_order, res = annotate_ranges(engine, orthologs, annotation, distance=0, strandness=True)
        finally:
            if own:
                engine.close()
        nb_off, nb_idx, ji, oc = res["nb_off"], a_order[res["nb_idx"]], res["ji"], res["oc"]
        a_len = annotation.end - annotation.start
        q_len = orthologs.end - orthologs.start
        if (q_len < 0).any() or (len(nb_idx) and (a_len[nb_idx] < 0).any()):
            raise ValueError("__len__() should return >= 0")      # what len(grange) raises in the reference
        fmt = f".{float_precision}g"
        ji_s = [format(x, fmt) for x in ji.tolist()]
        oc_s = [format(x, fmt) for x in oc.tolist()]
        name_map = defaultdict(list)
        for k, qi in enumerate(q_order.tolist()):                 # the reference iterates the SORTED ortholog list
            lo, hi = int(nb_off[k]), int(nb_off[k + 1])
            idx = nb_idx[lo:hi].tolist()
            name_map[orthologs.name[qi]].append(([annotation.name[j] for j in idx], str(int(q_len[qi])),
                                                 [str(int(a_len[j])) for j in idx], ji_s[lo:hi], oc_s[lo:hi]))
        dist_annot_amounts = []
        with open(output, "w") as outfile:
            outfile.write("Query\tOrthologs\tQuery_length\tOrthologs_lengths\tJI\tOC\n")
            for ortholog_name, stats in name_map.items():
                for names, qlen, lens, jis, ocs in stats:
                    if names:
                        outfile.write(f"{ortholog_name}\t{','.join(names)}\t{qlen}\t{','.join(lens)}\t"
                                      f"{','.join(jis)}\t{','.join(ocs)}\n")
                        dist_annot_amounts.append(len(names))
                    else:
                        outfile.write(f"{ortholog_name}\tNotAnnotated\t{qlen}\t0\t0\t0\n")
                        dist_annot_amounts.append(0)
        stats_msg = "-----------------------\n" \
                    "annotate_orthologs stats:\n" \
                    f"Recieved {len(orthologs)} orthologs from {subject_orthologs_filename}.\n" \
                    f"Distribution of amount of annotations:\n{simple_hist(dist_annot_amounts)}\n" \
                    "Reported all annotations for each ortholog.\n" \
                    "-----------------------"
    if isinstance(stats_filename, str):
        with open(stats_filename, "a") as stats_file:
            stats_file.write(stats_msg + "\n")
    else:
        print(stats_msg)
