"""Native ingest (libo2a_ingest.so, C++) of the reference's JSON inputs into a ColumnarBatch.

Replaces `json.load` + `GenomicRangesAlignment.from_dict` per lncRNA (`pipeline.py:743-754`) — the
94 % of the reference's build_orthologs time. Only the small `qrange` /
`srange` objects are decoded by Python, lazily, when a record has to be written.
"""
from __future__ import annotations

import ctypes as C
import json
import os

import numpy as np

from .columnar import ColumnarBatch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libo2a_ingest.so")
EXPORTS = ["o2a_ingest_create", "o2a_ingest_destroy", "o2a_ingest_load", "o2a_ingest_sizes",
           "o2a_ingest_export", "o2a_ingest_span", "o2a_ingest_file_error"]
_lib = None


def load_lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} not built: run `python -m ortho2align_b200.build`")
        lib = C.CDLL(LIB_PATH)
        P, i64 = C.c_void_p, C.c_int64
        lib.o2a_ingest_create.restype = P
        lib.o2a_ingest_create.argtypes = [C.c_int]
        lib.o2a_ingest_destroy.argtypes = [P]
        lib.o2a_ingest_load.restype = C.c_int
        lib.o2a_ingest_load.argtypes = [P, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), i64]
        lib.o2a_ingest_sizes.restype = C.c_int
        lib.o2a_ingest_sizes.argtypes = [P] + [C.POINTER(i64)] * 4
        lib.o2a_ingest_export.restype = C.c_int
        lib.o2a_ingest_export.argtypes = [P] + [P] * 16
        lib.o2a_ingest_span.restype = i64
        lib.o2a_ingest_span.argtypes = [P, i64, i64, i64, P, i64]
        lib.o2a_ingest_file_error.restype = C.c_char_p
        lib.o2a_ingest_file_error.argtypes = [P, i64]
        _lib = lib
    return _lib


class IngestResult:
    """batch: ColumnarBatch over the files that parsed (input order); lnc_file[l]: file index of
    lncRNA l; file_status[i] / errors[i]; score_is_int per HSP; range dicts on demand."""

    def __init__(self, lib, handle, n_files):
        self._lib, self._h = lib, handle
        sizes = [C.c_int64(0) for _ in range(4)]
        lib.o2a_ingest_sizes(handle, *[C.byref(s) for s in sizes])
        n_lnc, n_aln, n_hsp, n_bg = [int(s.value) for s in sizes]
        z = lambda n, dt: np.zeros(n, dtype=dt)
        self.file_status = z(n_files, np.int32)
        self.lnc_file = z(n_lnc, np.int64)
        bg_off, aln_off, hsp_off = z(n_lnc + 1, np.int64), z(n_lnc + 1, np.int64), z(n_aln + 1, np.int64)
        bg, qlen, slen = z(n_bg, np.int32), z(n_aln, np.int64), z(n_aln, np.int64)
        coords = np.zeros((n_hsp, 4), dtype=np.int32)
        qs, qe, ss, se = (z(n_hsp, np.int32) for _ in range(4))
        score, self.score_is_int = z(n_hsp, np.float64), z(n_hsp, np.uint8)
        self.spans = np.zeros((n_aln, 4), dtype=np.int64)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        lib.o2a_ingest_export(handle, p(self.file_status), p(self.lnc_file), p(bg_off), p(bg), p(aln_off), p(hsp_off),
                              p(qlen), p(slen), p(coords), p(qs), p(qe), p(ss), p(se), p(score), p(self.score_is_int),
                              p(self.spans))
        self.batch = ColumnarBatch(bg_off, bg, aln_off, hsp_off, qlen, slen, qs, qe, ss, se, score)
        self.batch._coords4 = coords
        self.errors = {i: lib.o2a_ingest_file_error(handle, i).decode() for i in np.nonzero(self.file_status)[0]}

    def range_dict(self, a, which):
        """Decoded qrange ('q') or srange ('s') dict of global alignment index `a`."""
        l = int(np.searchsorted(self.batch.aln_off, a, side="right") - 1)
        b, e = (self.spans[a, 0], self.spans[a, 1]) if which == "q" else (self.spans[a, 2], self.spans[a, 3])
        n = int(e - b)
        buf = C.create_string_buffer(n)
        self._lib.o2a_ingest_span(self._h, int(self.lnc_file[l]), int(b), int(e), buf, n)
        return json.loads(buf.raw[:n])

    def hsp_dict(self, h):
        """HSP `h` (global index) as the reference's JSON object, with the JSON number types."""
        b = self.batch
        sc = float(b.score[h])
        return {"qstart": int(b.qstart[h]), "qend": int(b.qend[h]), "sstart": int(b.sstart[h]), "send": int(b.send[h]),
                "score": int(sc) if self.score_is_int[h] else sc}

    def close(self):
        if self._h:
            self._lib.o2a_ingest_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def load_files(align_files, bg_files, threads=0) -> IngestResult:
    lib = load_lib()
    n = len(align_files)
    h = C.c_void_p(lib.o2a_ingest_create(int(threads)))
    arr_a = (C.c_char_p * n)(*[os.fsencode(p) for p in align_files])
    arr_b = (C.c_char_p * n)(*[os.fsencode(p) for p in bg_files])
    if lib.o2a_ingest_load(h, arr_a, arr_b, n) != 0:
        raise RuntimeError("o2a_ingest_load failed")
    return IngestResult(lib, h, n)
