"""ctypes binding of the native record formatter (include/o2a_records.h, libo2a_ingest.so): the text of the four
`significant.*_orthologs.{bed,tsv}` files from columnar chains, byte-identical to `records.chain_records`
(= the reference's `to_record` / `to_bed12`, genomicranges.py:1192-1301) at a fraction of its cost."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import ingest as _ingest

EXPORTS = ["o2a_records_format", "o2a_records_free", "o2a_records_repr_double", "o2a_best_orthologs_files"]


class O2ARecordsInput(C.Structure):
    _fields_ = [("n_chain", C.c_int64), ("hsp_off", C.c_void_p),
                ("qstart", C.c_void_p), ("qend", C.c_void_p), ("sstart", C.c_void_p), ("send", C.c_void_p),
                ("score", C.c_void_p), ("score_is_int", C.c_void_p), ("pval", C.c_void_p),
                ("qstrand_plus", C.c_void_p), ("str_off", C.c_void_p), ("str_blob", C.c_void_p)]


_bound = None


def lib():
    global _bound
    if _bound is None:
        L = _ingest.load_lib()
        L.o2a_records_format.restype = C.c_int
        L.o2a_records_format.argtypes = [C.POINTER(O2ARecordsInput), C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
        L.o2a_records_free.restype = None
        L.o2a_records_free.argtypes = [C.c_void_p]
        L.o2a_records_repr_double.restype = C.c_int
        L.o2a_records_repr_double.argtypes = [C.c_double, C.c_char_p]
        L.o2a_best_orthologs_files.restype = C.c_int
        L.o2a_best_orthologs_files.argtypes = [C.c_char_p] * 4 + [C.c_int, C.c_int] + [C.c_char_p] * 4 + [C.POINTER(C.c_int64)] * 2
        _bound = L
    return _bound


def repr_double(x):
    buf = C.create_string_buffer(40)
    n = lib().o2a_records_repr_double(float(x), buf)
    return buf.raw[:n].decode()


def best_orthologs_files(paths_in, value, function, paths_out):
    """get_best_orthologs (pipeline.py:959-1024) natively. Returns (n_records, n_groups), or None when the input holds
    something this reader leaves to the text-level restatement (include/o2a_records.h)."""
    from .result import BEST_FUNCTIONS, BEST_VALUES
    import os
    nr, ng = C.c_int64(0), C.c_int64(0)
    args = [os.fsencode(str(p)) for p in paths_in] + [BEST_VALUES[value], BEST_FUNCTIONS[function]] + \
           [os.fsencode(str(p)) for p in paths_out] + [C.byref(nr), C.byref(ng)]
    rc = lib().o2a_best_orthologs_files(*args)
    if rc == 0:
        return int(nr.value), int(ng.value)
    if rc == 1:
        return None
    raise OSError("o2a_best_orthologs_files: cannot read or write one of the files")


def format_records(hsp_off, qstart, qend, sstart, send, score, score_is_int, pval, qstrand_plus, strings, threads=0):
    """Per chain c: HSPs hsp_off[c]..hsp_off[c+1] (chain order) of the per-HSP arrays; qstrand_plus[c];
    strings[3c..3c+2] = str(qrange.chrom), str(name), str(srange.chrom) — a list of str, or a ready
    (str_off[3n+1] int64, blob bytes/uint8 array) pair.
    -> four `bytes`: query TSV, query BED12, subject TSV, subject BED12 (one line per chain, input order)."""
    n = len(hsp_off) - 1
    if isinstance(strings, tuple):
        str_off = np.ascontiguousarray(strings[0], dtype=np.int64)
        assert len(str_off) == 3 * n + 1
        blob = bytes(memoryview(np.ascontiguousarray(strings[1]))) + b"\0"
    else:
        assert len(strings) == 3 * n
        enc = [s.encode("utf-8") for s in strings]
        str_off = np.zeros(3 * n + 1, dtype=np.int64)
        if n:
            np.cumsum([len(b) for b in enc], out=str_off[1:])
        blob = b"".join(enc) + b"\0"
    keep = [np.ascontiguousarray(hsp_off, dtype=np.int64)]
    keep += [np.ascontiguousarray(a, dtype=np.int32) for a in (qstart, qend, sstart, send)]
    keep += [np.ascontiguousarray(score, dtype=np.float64), np.ascontiguousarray(score_is_int, dtype=np.uint8),
             np.ascontiguousarray(pval, dtype=np.float64), np.ascontiguousarray(qstrand_plus, dtype=np.uint8), str_off]
    s = O2ARecordsInput()
    s.n_chain = n
    for name, arr in zip(("hsp_off", "qstart", "qend", "sstart", "send", "score", "score_is_int", "pval",
                          "qstrand_plus", "str_off"), keep):
        setattr(s, name, arr.ctypes.data)
    s.str_blob = C.cast(C.c_char_p(blob), C.c_void_p).value
    out = (C.c_void_p * 4)()
    out_len = (C.c_int64 * 4)()
    if lib().o2a_records_format(C.byref(s), int(threads), out, out_len) != 0:
        raise RuntimeError("o2a_records_format failed (empty chain or invalid arguments)")
    try:
        return tuple(C.string_at(out[k], out_len[k]) for k in range(4))
    finally:
        for k in range(4):
            lib().o2a_records_free(out[k])
