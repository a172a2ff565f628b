"""Producer side of the alignment JSON, columnar from the BLAST stream on.

The reference's `_align_syntenies` (`pipeline.py:507-546`) does, per query gene and per syntenic region:

    alignment = neighbourhood.align_blast(synteny, word_size=…)      # Popen blastn -outfmt "7 std score",
                                                                     # Alignment.from_file_blast (alignment.py:515-586)
    alignment.to_genomic()                                           # genomicranges.py:1108-1129
    alignment = alignment.cut_coordinates(qleft=grange.start, qright=grange.end)   # alignment.py:734-803
    alignment.qrange = grange
    alignments.append(alignment.to_dict())                           # -> align_files/<name>.json

and `build_orthologs` later parses that JSON back. Here the BLAST tables are read by the native parser
(`libo2a_ingest.so`, `include/o2a_blast.h`) straight into HSP arrays, and `to_genomic` + `cut_coordinates` +
the qlen/slen rule run as one batched GPU pass (`o2a_cut_batch`, `csrc/o2a_cut.cuh`) whose output is the
columnar batch `o2a_build_batch` consumes — no Python objects, no JSON round trip. BLAST itself (the
subprocess) is outside this path: the tables are inputs.

There is no CPU fallback: `cut_alignments` needs the CUDA library and a GPU.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .columnar import ColumnarBatch
from .ingest import LIB_PATH

NO_BOUND = np.iinfo(np.int64).min          # O2A_NO_BOUND: a boundary that is None
BLAST_EXPORTS = ["o2a_blast_create", "o2a_blast_destroy", "o2a_blast_parse", "o2a_blast_sizes", "o2a_blast_export",
                 "o2a_blast_error", "o2a_bg_sample"]
_lib = None


class BlastParseError(Exception):
    """What the reference raises inside `from_file_blast`; `.name` is the Python exception name."""

    def __init__(self, table, name):
        super().__init__(f"BLAST table {table}: {name}")
        self.table, self.name = table, name


def load_lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} not built: run `python -m ortho2align_b200.build`")
        lib = C.CDLL(LIB_PATH)
        P, i64 = C.c_void_p, C.c_int64
        lib.o2a_blast_create.restype = P
        lib.o2a_blast_create.argtypes = [C.c_int]
        lib.o2a_blast_destroy.argtypes = [P]
        lib.o2a_blast_parse.restype = C.c_int
        lib.o2a_blast_parse.argtypes = [P, C.POINTER(C.c_char_p), C.POINTER(i64), i64, C.c_int, C.c_int]
        lib.o2a_blast_sizes.restype = C.c_int
        lib.o2a_blast_sizes.argtypes = [P, C.POINTER(i64), C.POINTER(i64)]
        lib.o2a_blast_export.restype = C.c_int
        lib.o2a_blast_export.argtypes = [P] + [P] * 10
        lib.o2a_blast_error.restype = C.c_char_p
        lib.o2a_blast_error.argtypes = [P, i64]
        lib.o2a_bg_sample.restype = i64
        lib.o2a_bg_sample.argtypes = [P, i64, i64, C.c_uint64, P]
        _lib = lib
    return _lib


def parse_blast_tables(tables, threads=0, start_type=1, end_type="inclusive"):
    """`from_file_blast` over a list of table texts (str or bytes), in parallel.

    Returns dict(hsp_off[n+1], qstart, qend, sstart, send (int32, 0-based half-open relative), score (float64),
    score_is_int (uint8), qlen, slen (int64 per table), status (int32 per table: 0 ok, 1 the reference raises,
    2 not representable), errors {table: exception name}). Tables with status != 0 hold no HSPs."""
    lib = load_lib()
    bufs = [t.encode() if isinstance(t, str) else bytes(t) for t in tables]
    n = len(bufs)
    arr = (C.c_char_p * max(n, 1))(*bufs)
    lens = (C.c_int64 * max(n, 1))(*[len(b) for b in bufs])
    h = C.c_void_p(lib.o2a_blast_create(int(threads)))
    try:
        if lib.o2a_blast_parse(h, arr, lens, n, int(start_type), 1 if end_type == "inclusive" else 0) != 0:
            raise RuntimeError("o2a_blast_parse failed")
        nt, nh = C.c_int64(0), C.c_int64(0)
        lib.o2a_blast_sizes(h, C.byref(nt), C.byref(nh))
        nh = int(nh.value)
        out = dict(hsp_off=np.zeros(n + 1, np.int64), qstart=np.zeros(nh, np.int32), qend=np.zeros(nh, np.int32),
                   sstart=np.zeros(nh, np.int32), send=np.zeros(nh, np.int32), score=np.zeros(nh),
                   score_is_int=np.zeros(nh, np.uint8), qlen=np.zeros(n, np.int64), slen=np.zeros(n, np.int64),
                   status=np.zeros(n, np.int32))
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        lib.o2a_blast_export(h, p(out["status"]), p(out["hsp_off"]), p(out["qstart"]), p(out["qend"]), p(out["sstart"]),
                             p(out["send"]), p(out["score"]), p(out["score_is_int"]), p(out["qlen"]), p(out["slen"]))
        out["errors"] = {int(i): lib.o2a_blast_error(h, int(i)).decode() for i in np.nonzero(out["status"])[0]}
    finally:
        lib.o2a_blast_destroy(h)
    return out


def sample_background(scores, observations=1000, seed=0):
    """The scores `_estimate_bg_for_single_query_*` would dump for a query (`pipeline.py:223-233, 259-269`):
    all of them when there are at most `observations`, else `random.Random(seed).sample(scores, observations)`
    — restated natively (CPython's MT19937 + `Random.sample`), bit-identical to the interpreter's."""
    lib = load_lib()
    sc = np.ascontiguousarray(scores, dtype=np.int32)
    out = np.zeros(min(len(sc), int(observations)), dtype=np.int32)
    n = lib.o2a_bg_sample(sc.ctypes.data_as(C.c_void_p), len(sc), int(observations), abs(int(seed)),
                          out.ctypes.data_as(C.c_void_p))
    if n < 0:
        raise ValueError("o2a_bg_sample: invalid arguments")
    return out[:n]


def backgrounds_from_tables(tables, table_query, n_queries, observations=1000, seed=0, threads=0):
    """Background score arrays straight from BLAST tables (no bg_files/*.json): `tables[t]` belongs to query
    `table_query[t]` (ascending); the scores of a query's tables are concatenated in table order
    (`pipeline.py:223-226`) and sampled like the reference does. Returns (bg_off, bg, n_found) — `bg` still
    needs `columnar.patch_background` semantics, which `pack`/ingest apply at build time (`pipeline.py:746-754`)."""
    r = parse_blast_tables(tables, threads=threads)
    table_query = np.asarray(table_query, dtype=np.int64)
    parts, found = [], np.zeros(n_queries, dtype=np.int64)
    for q in range(n_queries):
        ts = np.nonzero(table_query == q)[0]
        sc = np.concatenate([r["score"][r["hsp_off"][t]:r["hsp_off"][t + 1]] for t in ts]) if len(ts) else np.zeros(0)
        found[q] = len(sc)
        parts.append(sample_background(sc.astype(np.int32), observations, seed))
    bg_off = np.zeros(n_queries + 1, dtype=np.int64)
    np.cumsum([len(p) for p in parts], out=bg_off[1:])
    return bg_off, (np.concatenate(parts) if parts else np.zeros(0, np.int32)).astype(np.int32), found


def ranges_from_granges(qranges, sranges, genes=None, bounds=None):
    """Per-alignment range arrays for `cut_alignments` from objects with .start/.end/.strand (the reference's
    GenomicRange or `records.Range`): `qranges[a]` the aligned query range (the flanked neighbourhood),
    `sranges[a]` the syntenic range, `genes[a]` the gene whose start/end are the query cut
    (`pipeline.py:530-531`). `bounds`: optional dict of explicit qleft/qright/sleft/sright lists (None entries
    allowed) for the general `cut_coordinates(...)` call."""
    n = len(qranges)
    i64 = lambda xs: np.array(list(xs), dtype=np.int64).reshape(n)
    rg = dict(q_start=i64(r.start for r in qranges), q_end=i64(r.end for r in qranges),
              q_minus=np.array([r.strand == "-" for r in qranges], dtype=np.int8).reshape(n),
              s_start=i64(r.start for r in sranges), s_end=i64(r.end for r in sranges),
              s_minus=np.array([r.strand == "-" for r in sranges], dtype=np.int8).reshape(n))
    if genes is not None:
        rg["qleft"], rg["qright"] = i64(g.start for g in genes), i64(g.end for g in genes)
    for k, v in (bounds or {}).items():
        rg[k] = i64(NO_BOUND if x is None else x for x in v)
    return rg


def cut_alignments(engine, rel, ranges):
    """`to_genomic()` + `cut_coordinates(...)` for every alignment of `rel` (the dict `parse_blast_tables`
    returns, or any dict with hsp_off/qstart/qend/sstart/send/score) on the GPU.

    Returns dict(hsp_off, qstart, qend, sstart, send, score, cut, score_is_int, qlen, slen, status) on the host;
    `cut[h]` bit 0 = the score was scaled by a cut, bit 1 = the table already held a float (`rel["score_is_int"]`
    == 0); `score_is_int = (cut == 0)` is the JSON number type the reference would write."""
    out = engine.cut_batch(rel, ranges)
    out["score_is_int"] = (out["cut"] == 0).astype(np.uint8)
    return out


def to_columnar(cut, aln_off, bg_off, bg):
    """The cut result grouped into lncRNAs (`aln_off[l]..aln_off[l+1]` = the alignments of gene l) plus the
    backgrounds = the ColumnarBatch `Engine.build_batch` takes."""
    return ColumnarBatch(np.asarray(bg_off, np.int64), np.asarray(bg, np.int32), np.asarray(aln_off, np.int64),
                         cut["hsp_off"], cut["qlen"], cut["slen"], cut["qstart"], cut["qend"], cut["sstart"],
                         cut["send"], cut["score"])


def alignment_dicts(cut, a, qrange_dict, srange_dict):
    """Alignment `a` of a cut result in the reference's JSON schema (`to_dict`, genomicranges.py:1077-1082,
    alignment.py:588-601; HSP `kwargs` — the BLAST columns the build step never reads — left empty)."""
    h0, h1 = int(cut["hsp_off"][a]), int(cut["hsp_off"][a + 1])
    is_int = cut.get("score_is_int")
    hsps = []
    for h in range(h0, h1):
        sc = float(cut["score"][h])
        if is_int is not None and is_int[h]:
            sc = int(sc)
        hsps.append({"qstart": int(cut["qstart"][h]), "qend": int(cut["qend"][h]), "sstart": int(cut["sstart"][h]),
                     "send": int(cut["send"][h]), "score": sc, "pval": None, "kwargs": {}})
    return {"HSPs": hsps, "qlen": int(cut["qlen"][a]), "slen": int(cut["slen"][a]),
            "filtered_HSPs": [dict(h) for h in hsps], "filtered": False,
            "qrange": qrange_dict, "srange": srange_dict, "relative": False}
