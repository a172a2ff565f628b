"""Oracle for the producer side of the alignment JSON. TEST INFRASTRUCTURE ONLY.

* `parse_blast` — plain-Python restatement of `Alignment.from_file_blast` (alignment.py:515-586) with
  `numberize` (utils.py:6-41); small inputs only (tests, smoke).
* `cut_batch` — ctypes wrapper of `o2a_oracle_cut_batch` (oracle/o2a_oracle_producer.c): `to_genomic` +
  `cut_coordinates` over a whole batch of alignments.

Pinned against the reference run in the build container and tests/golden/producer.json
(oracle/make_golden_producer.py); see the header of o2a_oracle_producer.c.
"""
from __future__ import annotations

import ctypes as C
import time

import numpy as np

from . import oracle as _O

NO_BOUND = np.iinfo(np.int64).min
REPLACE = {"q. start": "qstart", "q. end": "qend", "s. start": "sstart", "s. end": "send",
           "query length": "qlen", "subject length": "slen"}          # alignment.py:440-445
LAST_CALL_SECONDS = 0.0


class BlastFormatError(ValueError):
    """from_file_blast's ValueError (alignment.py:551-552, 563-565)."""


def _numberize(s):
    """utils.py:21-41: float if it parses, int if that float is integral, else the string."""
    try:
        v = float(s)
    except ValueError:
        return s
    return int(v) if v.is_integer() else v


def parse_blast(text, start_type=1, end_type="inclusive"):
    """-> dict(hsps=[[qstart, qend, sstart, send, score], ...], qlen, slen) or raises what the reference raises
    (BlastFormatError for its ValueErrors, UnboundLocalError for a Fields line followed by zero hits)."""
    lines = text.split("\n")
    # file_object.readline() keeps the newline; emulate on the split list (the last piece has none)
    pos = 0

    def readline():
        nonlocal pos
        if pos >= len(lines):
            return ""
        ln = lines[pos]
        pos += 1
        return ln + "\n" if pos < len(lines) else ln

    line = readline()
    if not line.startswith("#"):
        raise BlastFormatError("Provided file_object is not in accepted file format.")
    for _ in range(100):
        line = readline()
        if line.startswith("# Fields:"):
            break
    else:
        return {"hsps": [], "qlen": 1, "slen": 1}                     # cls([], ...) + alignment.py:477-487 defaults
    fields = line.lstrip("# Fields: ").rstrip("\n").split(", ")     # lstrip strips a character SET
    if "score" not in fields:
        raise BlastFormatError("'score' is not found in alignment fields.")
    fields = [REPLACE.get(f, f) for f in fields]
    hit_number = int(readline().split(" ")[1])
    hsps, last = [], None
    for _ in range(hit_number):
        vals = [_numberize(v) for v in readline().strip().split("\t")]
        d = dict(zip(fields, vals))
        row = [d["qstart"], d["qend"], d["sstart"], d["send"], d["score"]]
        if start_type == 1:
            row[0] -= 1; row[1] -= 1; row[2] -= 1; row[3] -= 1
        if end_type == "inclusive":
            row[1] += 1; row[3] += 1
        hsps.append(row)
        last = d
    if last is None:
        raise UnboundLocalError("hsp")                                  # alignment.py:583 with hit_number == 0
    qlen, slen = last.get("qlen"), last.get("slen")
    if qlen is None:
        qlen = max(h[1] for h in hsps) + 1
    if slen is None:
        slen = max(max(h[2] for h in hsps) + 1, max(h[3] for h in hsps) + 1)
    return {"hsps": hsps, "qlen": qlen, "slen": slen}


def _bind():
    L = _O.lib()
    if not getattr(L, "_producer_bound", False):
        P, i64 = C.c_void_p, C.c_int64
        L.o2a_oracle_cut_batch.restype = C.c_int32
        L.o2a_oracle_cut_batch.argtypes = [i64] + [P] * 18 + [C.c_int32] + [P] * 10
        L._producer_bound = True
    return L


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def cut_batch(rel, ranges, n_threads=0):
    """rel: hsp_off, qstart, qend, sstart, send (int32), score (float64); ranges: q_start, q_end, q_minus, s_start,
    s_end, s_minus (+ optional relative, qleft, qright, sleft, sright; NO_BOUND entries = None).
    -> dict(hsp_off, qstart, qend, sstart, send, score, cut, qlen, slen, status); cut: bit 0 = scaled by a cut,
    bit 1 = the input score was a float already (rel["score_is_int"] == 0)."""
    global LAST_CALL_SECONDS
    L = _bind()
    hsp_off = np.ascontiguousarray(rel["hsp_off"], dtype=np.int64)
    n_aln, n = len(hsp_off) - 1, int(hsp_off[-1])
    i32 = lambda k: np.ascontiguousarray(rel[k], dtype=np.int32)
    i64 = lambda k: None if ranges.get(k) is None else np.ascontiguousarray(ranges[k], dtype=np.int64)
    ins = [i32("qstart"), i32("qend"), i32("sstart"), i32("send"), np.ascontiguousarray(rel["score"], dtype=np.float64),
           None if rel.get("score_is_int") is None else np.ascontiguousarray(rel["score_is_int"], dtype=np.uint8)]
    rg = [i64("q_start"), i64("q_end"), np.ascontiguousarray(ranges["q_minus"], dtype=np.int8),
          i64("s_start"), i64("s_end"), np.ascontiguousarray(ranges["s_minus"], dtype=np.int8),
          None if ranges.get("relative") is None else np.ascontiguousarray(ranges["relative"], dtype=np.uint8),
          i64("qleft"), i64("qright"), i64("sleft"), i64("sright")]
    cap = max(n, 1)
    out_off = np.zeros(n_aln + 1, dtype=np.int64)
    o = [np.zeros(cap, dtype=np.int32) for _ in range(4)] + [np.zeros(cap), np.zeros(cap, dtype=np.uint8)]
    qlen, slen, status = np.zeros(n_aln, np.int64), np.zeros(n_aln, np.int64), np.zeros(n_aln, np.int32)
    t0 = time.perf_counter()
    rc = L.o2a_oracle_cut_batch(n_aln, _p(hsp_off), *[_p(a) for a in ins], *[_p(a) for a in rg], int(n_threads),
                                _p(out_off), *[_p(a) for a in o], _p(qlen), _p(slen), _p(status))
    LAST_CALL_SECONDS = time.perf_counter() - t0
    if rc != 0:
        raise RuntimeError(f"oracle cut_batch failed ({rc})")
    m = int(out_off[-1])
    return dict(hsp_off=out_off, qstart=o[0][:m], qend=o[1][:m], sstart=o[2][:m], send=o[3][:m], score=o[4][:m],
                cut=o[5][:m], qlen=qlen, slen=slen, status=status)
