"""HostBatch — a batch of lncRNAs in host memory in the layout `o2a_build_batch` moves over PCIe.

The reference hands every worker two file names and the worker `json.load`s them (`pipeline.py:743-754`).
Here the ingest (`ingest.load_files`) — or `HostBatch.from_columnar` for callers that already hold columnar
arrays — produces ONE set of buffers per batch, allocated with `o2a_host_alloc` (page-locked, device-mapped),
in the lossless narrow encoding of `include/o2a.h`:

    score_q   int8 / int16 / int32 per HSP   (BLAST raw scores are integers; the most negative value marks an
    exc_*     exception lists                  HSP whose exact float64 score — a cut HSP, alignment.py:330 — is listed)
    bg_q      int8 / int16 / int32 per background score (after the reference's patches, pipeline.py:746-754)
    coords    int32 [n_hsp, 4] AoS (qstart, qend, sstart, send): only the retained ~1 % is ever read — by host threads
              of the library (host-gather mode), so this array stays in ordinary memory unless pin_coords=True
    offsets   bg_off, aln_off, hsp_off, qlen, slen (int64)

`Engine.build_host(hb, ...)` passes exactly these pointers to the library: what the ingest returns is what
crosses PCIe. Without a CUDA device (`o2a_host_alloc` fails) the same layout lives in ordinary memory — the
layout is host logic and is tested on CPU; the compute still has no CPU path.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .columnar import ColumnarBatch

_WIDTH_DTYPE = {8: np.int8, 16: np.int16, 32: np.int32}


class _Block:
    """One allocation: page-locked through the CUDA library when possible, else a numpy buffer."""

    def __init__(self, nbytes, pinned):
        self.nbytes = max(int(nbytes), 16)
        self.addr = 0
        self.lib = None
        self.fallback = None
        if pinned:
            try:
                from . import _lib
                lib = _lib.load()
                addr = lib.o2a_host_alloc(self.nbytes)
                if addr:
                    self.addr, self.lib = int(addr), lib
            except Exception:
                pass
        if not self.addr:
            self.fallback = np.empty(self.nbytes, dtype=np.uint8)
            self.addr = self.fallback.ctypes.data

    @property
    def pinned(self):
        return self.lib is not None

    def array(self, shape, dtype):
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) if np.ndim(shape) else int(shape)
        assert n * dtype.itemsize <= self.nbytes
        buf = (C.c_char * (n * dtype.itemsize)).from_address(self.addr)
        a = np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)
        return a

    def free(self):
        if self.lib is not None and self.addr:
            self.lib.o2a_host_free(C.c_void_p(self.addr))
        self.addr, self.lib, self.fallback = 0, None, None


def default_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


class HostBatch:
    """See the module docstring. Arrays are views of blocks owned by this object: keep it alive while they are used."""

    ARRAYS = ("bg_off", "aln_off", "hsp_off", "qlen", "slen", "bg_q", "score_q", "exc_off", "exc_pos", "exc_val", "coords")

    def __init__(self, n_lnc, n_aln, n_hsp, n_bg, score_bits, bg_bits, n_exc, pinned=True, pin_coords=False):
        self.n_lnc, self.n_aln, self.n_hsp, self.n_bg = int(n_lnc), int(n_aln), int(n_hsp), int(n_bg)
        self.score_bits, self.bg_bits, self.n_exc = int(score_bits), int(bg_bits), int(n_exc)
        self._blocks = []
        self.kde_factor = None
        self.names = None
        self.max_aln_hsps = 0
        self.max_lnc_bg = 0
        i64 = np.int64
        self.bg_off = self._new(self.n_lnc + 1, i64, pinned)
        self.aln_off = self._new(self.n_lnc + 1, i64, pinned)
        self.hsp_off = self._new(self.n_aln + 1, i64, pinned)
        self.qlen = self._new(self.n_aln, i64, pinned)
        self.slen = self._new(self.n_aln, i64, pinned)
        self.bg_q = self._new(self.n_bg, _WIDTH_DTYPE[self.bg_bits], pinned)
        self.score_q = self._new(self.n_hsp, _WIDTH_DTYPE[self.score_bits], pinned)
        self.exc_off = self._new(self.n_aln + 1, i64, pinned)
        self.exc_pos = self._new(self.n_exc, np.int32, pinned)
        self.exc_val = self._new(self.n_exc, np.float64, pinned)
        self.coords = self._new((self.n_hsp, 4), np.int32, pinned and pin_coords)
        self.pinned = all(b.pinned for b in self._blocks[:-1]) if pinned else False      # everything that is uploaded

    def _new(self, shape, dtype, pinned):
        n = int(np.prod(shape)) if np.ndim(shape) else int(shape)
        blk = _Block(n * np.dtype(dtype).itemsize, pinned)
        self._blocks.append(blk)
        return blk.array(shape, dtype)

    def finish(self):
        """Derived scalars, once the offset arrays are filled."""
        self.max_aln_hsps = int(np.diff(self.hsp_off).max()) if self.n_aln else 0
        self.max_lnc_bg = int(np.diff(self.bg_off).max()) if self.n_lnc else 0
        return self

    def ensure_kde_factor(self):
        if self.kde_factor is None:
            from .columnar import kde_factor
            self.kde_factor = np.array([kde_factor(int(k)) for k in np.diff(self.bg_off)], dtype=np.float64)
        return self.kde_factor

    def h2d_bytes(self):
        """Bytes the library uploads per call (everything but the coordinates, which are read in place / gathered)."""
        return int(sum(getattr(self, k).nbytes for k in self.ARRAYS if k != "coords"))

    def scores(self):
        """The exact float64 scores (decoded from score_q + exception lists): what `ColumnarBatch.score` holds."""
        sc = self.score_q.astype(np.float64)
        if self.n_exc:
            aln = np.repeat(np.arange(self.n_aln), np.diff(self.exc_off))
            sc[self.hsp_off[aln] + self.exc_pos] = self.exc_val
        return sc

    def to_columnar(self) -> ColumnarBatch:
        c = self.coords
        b = ColumnarBatch(self.bg_off.copy(), self.bg_q.astype(np.int32), self.aln_off.copy(), self.hsp_off.copy(),
                          self.qlen.copy(), self.slen.copy(), np.ascontiguousarray(c[:, 0]), np.ascontiguousarray(c[:, 1]),
                          np.ascontiguousarray(c[:, 2]), np.ascontiguousarray(c[:, 3]), self.scores(),
                          kde_factor=self.kde_factor, names=self.names)
        return b

    @classmethod
    def from_columnar(cls, batch: ColumnarBatch, threads=0, pinned=True, score_bits=None, bg_bits=None, pin_coords=False):
        """Native multi-threaded packer (`o2a_ingest_pack_narrow`, libo2a_ingest.so). `score_bits` / `bg_bits`: force a
        width (8, 16, 32) instead of the planner's choice; a background that does not fit `bg_bits` is an error."""
        from . import ingest
        lib = ingest.load_lib()
        threads = threads or default_threads()
        sb, bb = C.c_int32(0), C.c_int32(0)
        nx = (C.c_int64 * 3)()
        score = np.ascontiguousarray(batch.score, dtype=np.float64)
        bg = np.ascontiguousarray(batch.bg, dtype=np.int32)
        p = lambda a: C.c_void_p(a.ctypes.data)
        if lib.o2a_ingest_pack_plan(p(score), batch.n_hsp, p(bg), batch.n_bg, threads, C.byref(sb), C.byref(bb), nx) != 0:
            raise RuntimeError("o2a_ingest_pack_plan failed")
        if score_bits is None:
            score_bits = sb.value
        if bg_bits is None:
            bg_bits = bb.value
        elif bg_bits < bb.value:
            from .columnar import PackError
            raise PackError(f"background scores do not fit int{bg_bits}")
        n_exc = int(nx[{8: 0, 16: 1, 32: 2}[score_bits]])
        hb = cls(batch.n_lnc, batch.n_aln, batch.n_hsp, batch.n_bg, score_bits, bg_bits, n_exc, pinned=pinned,
                 pin_coords=pin_coords)
        for k in ("bg_off", "aln_off", "hsp_off", "qlen", "slen"):
            getattr(hb, k)[...] = getattr(batch, k)
        soa = [np.ascontiguousarray(getattr(batch, k), dtype=np.int32) for k in ("qstart", "qend", "sstart", "send")]
        rc = lib.o2a_ingest_pack_narrow(batch.n_aln, p(hb.hsp_off), p(score), score_bits, p(hb.score_q), p(hb.exc_off),
                                        p(hb.exc_pos), p(hb.exc_val), p(bg), batch.n_bg, bg_bits, p(hb.bg_q),
                                        p(soa[0]), p(soa[1]), p(soa[2]), p(soa[3]), p(hb.coords), threads)
        if rc != 0:
            raise RuntimeError("o2a_ingest_pack_narrow failed")
        hb.kde_factor = batch.kde_factor
        hb.names = batch.names
        return hb.finish()

    def close(self):
        for b in self._blocks:
            b.free()
        self._blocks = []
        for k in self.ARRAYS:
            setattr(self, k, None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
