"""Native ingest (libo2a_ingest.so, C++) of the reference's JSON inputs into a ColumnarBatch.

Replaces `json.load` + `GenomicRangesAlignment.from_dict` per lncRNA (`pipeline.py:743-754`) — the
94 % of the reference's build_orthologs time. Only the small `qrange` /
`srange` objects are decoded by Python, lazily, when a record has to be written.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import sys

import numpy as np

from .columnar import ColumnarBatch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libo2a_ingest.so")
EXPORTS = ["o2a_ingest_create", "o2a_ingest_destroy", "o2a_ingest_load", "o2a_ingest_sizes",
           "o2a_ingest_export", "o2a_ingest_span", "o2a_ingest_file_error",
           "o2a_ingest_narrow_plan", "o2a_ingest_export_narrow", "o2a_ingest_pack_plan", "o2a_ingest_pack_narrow",
           "o2a_ingest_range_texts", "o2a_ingest_range_fields"]
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
        lib.o2a_stat_files.restype = C.c_int
        lib.o2a_stat_files.argtypes = [C.POINTER(C.c_char_p), i64, C.c_int, P, P]
        lib.o2a_ingest_sizes.restype = C.c_int
        lib.o2a_ingest_sizes.argtypes = [P] + [C.POINTER(i64)] * 4
        lib.o2a_ingest_export.restype = C.c_int
        lib.o2a_ingest_export.argtypes = [P] + [P] * 17
        lib.o2a_ingest_narrow_plan.restype = C.c_int
        lib.o2a_ingest_narrow_plan.argtypes = [P, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(i64)]
        lib.o2a_ingest_export_narrow.restype = C.c_int
        lib.o2a_ingest_export_narrow.argtypes = [P, C.c_int32, C.c_int32, P, P, P, P, P]
        lib.o2a_ingest_pack_plan.restype = C.c_int
        lib.o2a_ingest_pack_plan.argtypes = [P, i64, P, i64, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(i64)]
        lib.o2a_ingest_pack_narrow.restype = C.c_int
        lib.o2a_ingest_pack_narrow.argtypes = [i64, P, P, C.c_int32, P, P, P, P, P, i64, C.c_int32, P, P, P, P, P, P, C.c_int]
        lib.o2a_ingest_range_texts.restype = i64
        lib.o2a_ingest_range_texts.argtypes = [P, P, P, i64]
        lib.o2a_ingest_range_fields.restype = i64
        lib.o2a_ingest_range_fields.argtypes = [P, P, P, i64, P, P]
        lib.o2a_ingest_span.restype = i64
        lib.o2a_ingest_span.argtypes = [P, i64, i64, i64, P, i64]
        lib.o2a_ingest_file_error.restype = C.c_char_p
        lib.o2a_ingest_file_error.argtypes = [P, i64]
        _lib = lib
    return _lib


class IngestResult:
    """host: HostBatch over the files that parsed (input order) — the pinned, narrow layout the CUDA library takes;
    batch: the same lncRNAs as a plain ColumnarBatch (built on first use); lnc_file[l]: file index of lncRNA l;
    file_status[i] / errors[i]; score (float64) and score_is_int per HSP; relative per alignment; range dicts on demand."""

    def __init__(self, lib, handle, n_files, pinned=True, pin_coords=False):
        from .hostbatch import HostBatch
        self._lib, self._h = lib, handle
        sizes = [C.c_int64(0) for _ in range(4)]
        lib.o2a_ingest_sizes(handle, *[C.byref(s) for s in sizes])
        n_lnc, n_aln, n_hsp, n_bg = [int(s.value) for s in sizes]
        sb, bb, nx = C.c_int32(0), C.c_int32(0), C.c_int64(0)
        lib.o2a_ingest_narrow_plan(handle, C.byref(sb), C.byref(bb), C.byref(nx))
        hb = HostBatch(n_lnc, n_aln, n_hsp, n_bg, sb.value, bb.value, nx.value, pinned=pinned, pin_coords=pin_coords)
        z = lambda n, dt: np.zeros(n, dtype=dt)
        self.file_status = z(n_files, np.int32)
        self.lnc_file = z(n_lnc, np.int64)
        self.score, self.score_is_int = np.empty(n_hsp, dtype=np.float64), np.empty(n_hsp, dtype=np.uint8)
        self.spans = np.zeros((n_aln, 4), dtype=np.int64)
        self.relative = z(n_aln, np.uint8)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        lib.o2a_ingest_export(handle, p(self.file_status), p(self.lnc_file), p(hb.bg_off), None, p(hb.aln_off), p(hb.hsp_off),
                              p(hb.qlen), p(hb.slen), p(hb.coords), None, None, None, None, p(self.score),
                              p(self.score_is_int), p(self.spans), p(self.relative))
        if lib.o2a_ingest_export_narrow(handle, sb.value, bb.value, p(hb.score_q), p(hb.exc_off), p(hb.exc_pos),
                                        p(hb.exc_val), p(hb.bg_q)) != 0:
            raise RuntimeError("o2a_ingest_export_narrow failed")
        self.host = hb.finish()
        self._batch = None
        self._fields = None
        self._range_off = self._range_blob = None
        self.errors = {i: lib.o2a_ingest_file_error(handle, i).decode() for i in np.nonzero(self.file_status)[0]}

    @property
    def batch(self):
        """Plain ColumnarBatch view of the same lncRNAs (int32 background, SoA coordinates, float64 scores)."""
        if self._batch is None:
            hb = self.host
            c = hb.coords
            b = ColumnarBatch(hb.bg_off, hb.bg_q.astype(np.int32), hb.aln_off, hb.hsp_off, hb.qlen, hb.slen,
                              np.ascontiguousarray(c[:, 0]), np.ascontiguousarray(c[:, 1]),
                              np.ascontiguousarray(c[:, 2]), np.ascontiguousarray(c[:, 3]), self.score)
            b._coords4 = c
            self._batch = b
        return self._batch

    @property
    def hsp_off(self):
        return self.host.hsp_off

    def gather_hsps(self, gidx):
        """qstart, qend, sstart, send, score of the HSPs `gidx` (global indices): one gather from the AoS records."""
        c = self.host.coords[gidx]
        return c[:, 0], c[:, 1], c[:, 2], c[:, 3], self.score[gidx]

    def _ranges(self):
        if self._range_off is None:
            n_aln = self.host.n_aln
            off = np.zeros(2 * n_aln + 1, dtype=np.int64)
            total = int(self._lib.o2a_ingest_range_texts(self._h, off.ctypes.data_as(C.c_void_p), None, 0))
            blob = np.zeros(max(total, 1), dtype=np.uint8)
            self._lib.o2a_ingest_range_texts(self._h, off.ctypes.data_as(C.c_void_p), blob.ctypes.data_as(C.c_void_p), total)
            self._range_off, self._range_blob = off, blob[:total]
        return self._range_off, self._range_blob

    def range_fields(self):
        """(name_off[3*n_aln+1], blob, q_plus[n_aln], simple[n_aln]): per alignment str(qrange.chrom), str(qrange.name),
        str(srange.chrom) as UTF-8 slices of `blob`, qrange.strand in ('+', '.'), and whether the native reader could
        render all three (o2a_ingest_range_fields) — what the record formatter needs, without a JSON decode in Python."""
        if self._fields is None:
            n_aln = self.host.n_aln
            off = np.zeros(3 * n_aln + 1, dtype=np.int64)
            qp, simple = np.zeros(n_aln, dtype=np.uint8), np.zeros(n_aln, dtype=np.uint8)
            p = lambda a: a.ctypes.data_as(C.c_void_p)
            total = int(self._lib.o2a_ingest_range_fields(self._h, p(off), None, 0, p(qp), p(simple)))
            blob = np.zeros(max(total, 1), dtype=np.uint8)
            self._lib.o2a_ingest_range_fields(self._h, p(off), p(blob), total, p(qp), p(simple))
            self._fields = (off, blob[:total], qp, simple)
        return self._fields

    def range_dict(self, a, which):
        """Decoded qrange ('q') or srange ('s') dict of global alignment index `a`. Consecutive alignments of
        a lncRNA carry the same qrange text: the same dict object is returned for identical text."""
        return _decode_cached(self, self.range_text(a, which), which)

    def range_text(self, a, which):
        """Raw JSON text of the qrange / srange object of alignment `a` (bytes)."""
        off, blob = self._ranges()
        k = 2 * int(a) + (0 if which == "q" else 1)
        return blob[off[k]:off[k + 1]].tobytes()

    def hsp_dict(self, h):
        return _hsp_dict(self.batch, self.score_is_int, h)

    def hsp_dicts(self, idx):
        return _hsp_dicts(self.batch, self.score_is_int, idx)

    def close(self):
        if self._h:
            self._ranges()                 # the texts outlive the native handle
            self.range_fields()
            self._lib.o2a_ingest_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            if self._h:
                self._lib.o2a_ingest_destroy(self._h)
                self._h = None
        except Exception:
            pass


def _decode_cached(owner, text, which="q"):
    """One-entry cache per side: the alignments of a lncRNA repeat the same qrange text (callers ask q, s, q, s, ...,
    so a shared entry would never hit), and the caller recognises a repeated qrange by object identity."""
    key = "_last_range_" + which
    last = getattr(owner, key, None)
    if last is not None and last[0] == text:
        return last[1]
    d = json.loads(text)
    setattr(owner, key, (text, d))
    return d


def _hsp_dicts(b, score_is_int, idx):
    """HSPs `idx` (global indices, array) as the reference's JSON objects: one gather per column."""
    idx = np.asarray(idx, dtype=np.int64)
    cols = [np.asarray(getattr(b, k))[idx].tolist() for k in ("qstart", "qend", "sstart", "send")]
    sc = np.asarray(b.score)[idx].tolist()
    isint = np.asarray(score_is_int)[idx].tolist()
    return [{"qstart": a, "qend": e, "sstart": c, "send": d, "score": int(v) if i else v}
            for a, e, c, d, v, i in zip(*cols, sc, isint)]


def _hsp_dict(b, score_is_int, h):
    """HSP `h` (global index) as the reference's JSON object, with the JSON number types."""
    sc = float(b.score[h])
    return {"qstart": int(b.qstart[h]), "qend": int(b.qend[h]), "sstart": int(b.sstart[h]), "send": int(b.send[h]),
            "score": int(sc) if score_is_int[h] else sc}


# --------------------------------------------------------------------------------------------------
# Columnar sidecar: the parsed batch written once next to the JSON directories (SURVEY.md §8f rank 1,
# "optional binary sidecar written once"). One file: 16-byte preamble, a JSON directory, then the
# arrays as raw little-endian sections aligned to 4 KiB so that they can be memory-mapped (or read
# straight into pinned buffers) without any parsing.
# --------------------------------------------------------------------------------------------------
SIDECAR_MAGIC = b"O2ACOL01"
SIDECAR_VERSION = 4
_SIDECAR_ALIGN = 4096
_SIDECAR_ARRAYS = ("bg_off", "bg", "aln_off", "hsp_off", "qlen", "slen", "qstart", "qend", "sstart", "send", "score")


def _encode_paths(paths):
    """os.fsencode of every path, without its per-call type dispatch for the usual list of str."""
    enc, err = sys.getfilesystemencoding(), sys.getfilesystemencodeerrors()
    return [p.encode(enc, err) if type(p) is str else os.fsencode(p) for p in paths]


def _manifest(align_files, bg_files):
    """What makes a sidecar stale: the ordered file list with sizes and mtimes (missing bg files: size -1), as three
    arrays — names (newline-joined basenames, uint8), sizes and mtimes (int64, alignment files then background files).
    The files are stat'ed by the native library on a pool of threads (`o2a_stat_files`); at 10^5 lncRNAs a JSON manifest
    cost more to decode and compare than the mapped arrays it guards."""
    lib = load_lib()
    n = len(align_files)
    paths = (C.c_char_p * (2 * n))(*_encode_paths(align_files), *_encode_paths(bg_files))
    size, mtime = np.empty(2 * n, dtype=np.int64), np.empty(2 * n, dtype=np.int64)
    if lib.o2a_stat_files(paths, 2 * n, 0, size.ctypes.data_as(C.c_void_p), mtime.ctypes.data_as(C.c_void_p)) != 0:
        raise RuntimeError("o2a_stat_files failed")
    names = "\n".join(a.rsplit(os.sep, 1)[-1] for a in align_files).encode("utf-8", "surrogateescape")
    return np.frombuffer(names, dtype=np.uint8), size, mtime


def save_sidecar(path, r, align_files, bg_files):
    """Write the IngestResult `r` (all files parsed) as a columnar sidecar; atomic via rename."""
    b = r.batch
    range_off, range_blob = r._ranges()
    sections = {k: np.ascontiguousarray(getattr(b, k)) for k in _SIDECAR_ARRAYS}
    sections["score_is_int"] = np.ascontiguousarray(r.score_is_int)
    sections["relative"] = np.ascontiguousarray(r.relative)
    sections["lnc_file"] = np.ascontiguousarray(r.lnc_file)
    sections["range_off"] = np.ascontiguousarray(range_off)
    sections["range_text"] = np.ascontiguousarray(range_blob)
    name_off, name_blob, q_plus, simple = r.range_fields()
    sections["name_off"], sections["name_blob"] = np.ascontiguousarray(name_off), np.ascontiguousarray(name_blob)
    sections["q_plus"], sections["simple"] = np.ascontiguousarray(q_plus), np.ascontiguousarray(simple)
    man_names, man_size, man_mtime = _manifest(align_files, bg_files)
    sections["man_names"], sections["man_size"], sections["man_mtime"] = np.ascontiguousarray(man_names), man_size, man_mtime
    directory = {"version": SIDECAR_VERSION, "n_files": len(align_files), "sections": {}}
    # two passes: the directory's own length decides where the first section starts
    def layout(base):
        off = base
        for k, arr in sections.items():
            off = (off + _SIDECAR_ALIGN - 1) // _SIDECAR_ALIGN * _SIDECAR_ALIGN
            directory["sections"][k] = {"dtype": arr.dtype.str, "shape": list(arr.shape), "offset": off}
            off += arr.nbytes
        return off
    layout(1 << 40)                                            # widest offsets -> upper bound of the JSON length
    head_len = len(json.dumps(directory).encode()) + 64
    layout(16 + head_len)
    head = json.dumps(directory).encode().ljust(head_len, b" ")
    tmp = f"{path}.tmp{os.getpid()}"
    with open(tmp, "wb") as f:
        f.write(SIDECAR_MAGIC + np.int64(head_len).tobytes())
        f.write(head)
        for k, arr in sections.items():
            f.seek(directory["sections"][k]["offset"])
            arr.tofile(f)
    os.replace(tmp, path)


class SidecarResult:
    """A sidecar opened for reading: same surface as IngestResult (batch, lnc_file, file_status, errors,
    score_is_int, range_dict, hsp_dict), arrays memory-mapped from the file."""

    def __init__(self, path, directory):
        sec = directory["sections"]
        def arr(k):
            d = sec[k]
            n = int(np.prod(d["shape"])) if d["shape"] else 1
            if n == 0:
                return np.zeros(d["shape"], dtype=np.dtype(d["dtype"]))
            m = np.memmap(path, dtype=np.dtype(d["dtype"]), mode="r", offset=d["offset"], shape=tuple(d["shape"]))
            return np.asarray(m)           # plain ndarray view of the mapping: np.memmap's own indexing is slow
        self.batch = ColumnarBatch(*[arr(k) for k in _SIDECAR_ARRAYS])
        self.score_is_int = arr("score_is_int")
        self.relative = arr("relative")
        self.score = self.batch.score
        self.lnc_file = arr("lnc_file")
        self._range_off, self._range_text = arr("range_off"), arr("range_text")
        self._fields = (arr("name_off"), arr("name_blob"), arr("q_plus"), arr("simple"))
        self._manifest = (arr("man_names"), arr("man_size"), arr("man_mtime"))
        self._host = None
        self.file_status = np.zeros(directory["n_files"], dtype=np.int32)
        self.errors = {}

    @property
    def host(self):
        """The pinned narrow layout of the mapped arrays (packed once, on first use)."""
        if self._host is None:
            from .hostbatch import HostBatch
            self._host = HostBatch.from_columnar(self.batch)
        return self._host

    @property
    def hsp_off(self):
        return self.batch.hsp_off

    def gather_hsps(self, gidx):
        b = self.batch
        return tuple(np.asarray(getattr(b, k))[gidx] for k in ("qstart", "qend", "sstart", "send", "score"))

    def _ranges(self):
        return self._range_off, self._range_text

    def range_fields(self):
        return self._fields

    def range_text(self, a, which):
        k = 2 * a + (0 if which == "q" else 1)
        return bytes(self._range_text[self._range_off[k]:self._range_off[k + 1]])

    def range_dict(self, a, which):
        return _decode_cached(self, self.range_text(a, which), which)

    def hsp_dict(self, h):
        return _hsp_dict(self.batch, self.score_is_int, h)

    def hsp_dicts(self, idx):
        return _hsp_dicts(self.batch, self.score_is_int, idx)

    def close(self):
        pass


def open_sidecar(path, align_files, bg_files):
    """SidecarResult if `path` holds a sidecar written for exactly these files (names, sizes, mtimes), else None."""
    try:
        with open(path, "rb") as f:
            pre = f.read(16)
            if len(pre) != 16 or pre[:8] != SIDECAR_MAGIC:
                return None
            directory = json.loads(f.read(int(np.frombuffer(pre[8:], dtype=np.int64)[0])))
    except (OSError, ValueError):
        return None
    if directory.get("version") != SIDECAR_VERSION or directory.get("n_files") != len(align_files):
        return None
    r = SidecarResult(path, directory)
    names, size, mtime = _manifest(align_files, bg_files)
    if not (np.array_equal(r._manifest[0], names) and np.array_equal(r._manifest[1], size) and np.array_equal(r._manifest[2], mtime)):
        return None
    return r


def load_files(align_files, bg_files, threads=0, pinned=True, pin_coords=False) -> IngestResult:
    """Parse the file pairs on `threads` native threads. The result's `.host` is the pinned narrow layout
    (`hostbatch.HostBatch`) the CUDA library takes as is; `pinned=False` keeps it in ordinary memory."""
    lib = load_lib()
    n = len(align_files)
    h = C.c_void_p(lib.o2a_ingest_create(int(threads)))
    arr_a = (C.c_char_p * n)(*_encode_paths(align_files))
    arr_b = (C.c_char_p * n)(*_encode_paths(bg_files))
    if lib.o2a_ingest_load(h, arr_a, arr_b, n) != 0:
        raise RuntimeError("o2a_ingest_load failed")
    return IngestResult(lib, h, n, pinned=pinned, pin_coords=pin_coords)
