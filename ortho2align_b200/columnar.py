"""Columnar (CSR) batch of many lncRNAs: the layout every kernel and the C-ABI consume.

The reference keeps one Python object per HSP (`alignment.py:67-110`) inside one
`GenomicRangesAlignment` per syntenic region (`genomicranges.py:1024-1054`) inside one JSON
file per lncRNA (`pipeline.py:743-745`). Here the same information is three CSR levels:

    lncRNA l      : background scores  bg[bg_off[l]:bg_off[l+1]]        (int32)
                    alignments         aln_off[l] .. aln_off[l+1]
    alignment a   : HSPs               hsp_off[a] .. hsp_off[a+1],  qlen[a], slen[a]
    HSP h         : qstart,qend,sstart,send (int32, genomic), score (float64)

`qlen`/`slen` are *inputs* taken from the JSON (`alignment.py:477-487`), never recomputed.
The background is stored after the reference's degenerate-set patches (`pipeline.py:746-754`),
see `patch_background`.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

INT32_MAX = 2**31 - 1


def patch_background(query_bg):
    """The reference's background patches, in its order (`pipeline.py:746-754`).

    `None` stands for "file does not exist" (-> `[0, 1]`). Returns a new list.
    """
    bg = [0, 1] if query_bg is None else list(query_bg)
    if len(set(bg)) <= 1:
        bg += [1]
    if bg[0] == 1:
        bg.append(0)
    return bg


_KDE_FACTOR_CACHE = {}


def kde_factor(n):
    """Scott's factor exactly as `scipy.stats.gaussian_kde` computes it for 1-D data of size n.

    The reference divides by `estimator.factor` only (`fitting.py:295`); scipy computes it as
    `power(neff, -1/(d+4))` with `neff = 1/vecdot(w, w)`, `w = ones(n)/n`, which is *not*
    bit-identical to `n**-0.2` (e.g. n=1000 -> neff=1000.0000000000005). This is host logic on
    the reference side too, so it is computed here with numpy and handed to the device.
    """
    f = _KDE_FACTOR_CACHE.get(n)
    if f is None:
        w = np.ones(n) / n
        neff = 1.0 / np.vecdot(w, w)
        f = float(np.power(neff, -1.0 / 5.0))
        _KDE_FACTOR_CACHE[n] = f
    return f


@dataclass
class ColumnarBatch:
    """CSR batch. All arrays are C-contiguous numpy arrays on the host."""
    bg_off: np.ndarray      # int64 [n_lnc+1]
    bg: np.ndarray          # int32 [n_bg]
    aln_off: np.ndarray     # int64 [n_lnc+1]
    hsp_off: np.ndarray     # int64 [n_aln+1]
    qlen: np.ndarray        # int64 [n_aln]
    slen: np.ndarray        # int64 [n_aln]
    qstart: np.ndarray      # int32 [n_hsp]
    qend: np.ndarray        # int32 [n_hsp]
    sstart: np.ndarray      # int32 [n_hsp]
    send: np.ndarray        # int32 [n_hsp]
    score: np.ndarray       # float64 [n_hsp]
    kde_factor: np.ndarray = field(default=None)  # float64 [n_lnc], filled lazily
    names: list = field(default=None)             # optional lncRNA names

    @property
    def n_lnc(self):
        return len(self.aln_off) - 1

    @property
    def n_aln(self):
        return len(self.hsp_off) - 1

    @property
    def n_hsp(self):
        return int(self.hsp_off[-1])

    @property
    def n_bg(self):
        return int(self.bg_off[-1])

    def coords4(self):
        """AoS view (qstart,qend,sstart,send) per HSP for `o2a_batch.coords` (built once, cached)."""
        c = getattr(self, "_coords4", None)
        if c is None or len(c) != self.n_hsp:
            c = np.ascontiguousarray(np.stack([self.qstart, self.qend, self.sstart, self.send], axis=1))
            self._coords4 = c
        return c

    def bg_narrow(self, bits):
        """int16 / int8 copy of the background for `o2a_batch.bg_bits = 16 | 8`; PackError if a score does not fit
        (BLAST raw scores of background alignments are small: int8 holds them in practice, int16 always)."""
        dt = {8: np.int8, 16: np.int16}[bits]
        cache = getattr(self, "_bg_narrow", None)
        if cache is not None and cache[0] == bits and len(cache[1]) == self.n_bg:
            return cache[1]
        info = np.iinfo(dt)
        if self.n_bg and (self.bg.max() > info.max or self.bg.min() < info.min):
            raise PackError(f"background scores do not fit int{bits}")
        b = self.bg.astype(dt)
        self._bg_narrow = (bits, b)
        return b

    def bg16(self):
        return self.bg_narrow(16)

    def quantized_scores(self, bits=16):
        """Lossless narrow encoding of `score` for `o2a_batch.score_q` (score_bits = 8 | 16 | 32):
        integral scores inside the integer type as-is; every other HSP (fractional score of a cut
        HSP, out-of-range value) gets the sentinel (most negative value) and an entry in its
        alignment's exception list. Returns (score_q, exc_off, exc_pos, exc_val)."""
        dt = {8: np.int8, 16: np.int16, 32: np.int32}[bits]
        info = np.iinfo(dt)
        sc = self.score
        with np.errstate(invalid="ignore"):
            ok = (sc == np.floor(sc)) & (sc > info.min) & (sc <= info.max) & ~((sc == 0) & np.signbit(sc))
        q = np.where(ok, sc, info.min).astype(dt)
        exc = np.nonzero(~ok)[0]
        aln_of = np.searchsorted(self.hsp_off, exc, side="right") - 1
        exc_off = np.zeros(self.n_aln + 1, dtype=np.int64)
        np.cumsum(np.bincount(aln_of, minlength=self.n_aln), out=exc_off[1:])
        exc_pos = (exc - self.hsp_off[aln_of]).astype(np.int32)
        return np.ascontiguousarray(q), exc_off, np.ascontiguousarray(exc_pos), np.ascontiguousarray(sc[exc])

    def ensure_kde_factor(self):
        if self.kde_factor is None:
            n = np.diff(self.bg_off)
            self.kde_factor = np.array([kde_factor(int(k)) for k in n], dtype=np.float64)
        return self.kde_factor

    def validate(self):
        assert self.bg_off.dtype == np.int64 and self.aln_off.dtype == np.int64
        assert self.hsp_off.dtype == np.int64
        assert self.bg.dtype == np.int32
        for a in (self.qstart, self.qend, self.sstart, self.send):
            assert a.dtype == np.int32 and len(a) == self.n_hsp
        assert self.score.dtype == np.float64 and len(self.score) == self.n_hsp
        assert len(self.bg_off) == self.n_lnc + 1
        assert int(self.aln_off[-1]) == self.n_aln
        assert len(self.qlen) == self.n_aln and len(self.slen) == self.n_aln
        assert np.all(np.diff(self.bg_off) >= 2), "background must be patched (>= 2 scores)"
        return self

    def slice_lnc(self, lo, hi):
        """Sub-batch of lncRNAs [lo, hi) (contiguous shard)."""
        a0, a1 = int(self.aln_off[lo]), int(self.aln_off[hi])
        h0, h1 = int(self.hsp_off[a0]), int(self.hsp_off[a1])
        b0, b1 = int(self.bg_off[lo]), int(self.bg_off[hi])
        return ColumnarBatch(
            bg_off=(self.bg_off[lo:hi + 1] - b0).copy(), bg=self.bg[b0:b1].copy(),
            aln_off=(self.aln_off[lo:hi + 1] - a0).copy(),
            hsp_off=(self.hsp_off[a0:a1 + 1] - h0).copy(),
            qlen=self.qlen[a0:a1].copy(), slen=self.slen[a0:a1].copy(),
            qstart=self.qstart[h0:h1].copy(), qend=self.qend[h0:h1].copy(),
            sstart=self.sstart[h0:h1].copy(), send=self.send[h0:h1].copy(),
            score=self.score[h0:h1].copy(),
            kde_factor=None if self.kde_factor is None else self.kde_factor[lo:hi].copy(),
            names=None if self.names is None else self.names[lo:hi])

    def take_lnc(self, idx):
        """Sub-batch made of the listed lncRNAs, in that order (LPT shards). Vectorised: one gather per array."""
        idx = np.asarray(idx, dtype=np.int64)
        if len(idx) == self.n_lnc and np.array_equal(idx, np.arange(self.n_lnc)):
            return self

        def ranges(off, sel):
            """new offsets + flat source indices of the segments off[i]..off[i+1], i in sel"""
            n = (off[sel + 1] - off[sel]).astype(np.int64)
            new = np.zeros(len(sel) + 1, dtype=np.int64)
            np.cumsum(n, out=new[1:])
            src = np.repeat(off[sel] - new[:-1], n) + np.arange(int(new[-1]), dtype=np.int64)
            return new, src

        bg_off, bsrc = ranges(self.bg_off, idx)
        aln_off, asrc = ranges(self.aln_off, idx)          # asrc: source alignment indices, in order
        hsp_off, hsrc = ranges(self.hsp_off, asrc)
        return ColumnarBatch(bg_off, self.bg[bsrc], aln_off, hsp_off, self.qlen[asrc], self.slen[asrc],
                             self.qstart[hsrc], self.qend[hsrc], self.sstart[hsrc], self.send[hsrc], self.score[hsrc],
                             kde_factor=None if self.kde_factor is None else self.kde_factor[idx],
                             names=None if self.names is None else [self.names[int(i)] for i in idx])

    def nbytes(self):
        return sum(getattr(self, k).nbytes for k in
                   ("bg_off", "bg", "aln_off", "hsp_off", "qlen", "slen",
                    "qstart", "qend", "sstart", "send", "score"))


def _cat_off(offs):
    out = [np.zeros(1, dtype=np.int64)]
    base = 0
    for o in offs:
        out.append(o[1:] + base)
        base += int(o[-1])
    return np.concatenate(out)


def concat_batches(parts):
    if len(parts) == 0:
        z = np.zeros(1, dtype=np.int64)
        e32 = np.zeros(0, dtype=np.int32)
        return ColumnarBatch(z, e32, z.copy(), z.copy(), np.zeros(0, np.int64), np.zeros(0, np.int64),
                             e32, e32.copy(), e32.copy(), e32.copy(), np.zeros(0, np.float64),
                             names=[])
    kf = None
    if all(p.kde_factor is not None for p in parts):
        kf = np.concatenate([p.kde_factor for p in parts])
    names = None
    if all(p.names is not None for p in parts):
        names = [n for p in parts for n in p.names]
    cat = np.concatenate
    return ColumnarBatch(
        bg_off=_cat_off([p.bg_off for p in parts]), bg=cat([p.bg for p in parts]),
        aln_off=_cat_off([p.aln_off for p in parts]),
        hsp_off=_cat_off([p.hsp_off for p in parts]),
        qlen=cat([p.qlen for p in parts]), slen=cat([p.slen for p in parts]),
        qstart=cat([p.qstart for p in parts]), qend=cat([p.qend for p in parts]),
        sstart=cat([p.sstart for p in parts]), send=cat([p.send for p in parts]),
        score=cat([p.score for p in parts]), kde_factor=kf, names=names)


class PackError(ValueError):
    """Raised when a lncRNA cannot be represented (the reference would raise too, or the value
    range exceeds the int32/float64 columnar types)."""


def pack_lncrna(alignment_dicts, query_bg):
    """One lncRNA: list of alignment dicts (reference JSON schema, `genomicranges.py:1077-1106`)
    + raw background list (or None if the file is missing) -> single-lncRNA ColumnarBatch.

    Mirrors the reference quirk that the working HSP list is rebuilt from key 'HSPs'
    (`genomicranges.py:1094-1100`): 'filtered_HSPs' content is ignored.
    """
    bg = patch_background(query_bg)
    bg_arr = np.asarray(bg)
    if bg_arr.dtype.kind not in "iu" or (len(bg_arr) and (bg_arr.max() > INT32_MAX or bg_arr.min() < -INT32_MAX)):
        raise PackError("background scores must be integers within int32")
    n_aln = len(alignment_dicts)
    counts = np.fromiter((len(d["HSPs"]) for d in alignment_dicts), dtype=np.int64, count=n_aln)
    hsp_off = np.zeros(n_aln + 1, dtype=np.int64)
    np.cumsum(counts, out=hsp_off[1:])
    n = int(hsp_off[-1])
    coords = np.empty((4, n), dtype=np.int64)
    score = np.empty(n, dtype=np.float64)
    qlen = np.empty(n_aln, dtype=np.int64)
    slen = np.empty(n_aln, dtype=np.int64)
    k = 0
    for a, d in enumerate(alignment_dicts):
        hs = d["HSPs"]
        m = len(hs)
        if m:
            coords[0, k:k + m] = [h["qstart"] for h in hs]
            coords[1, k:k + m] = [h["qend"] for h in hs]
            coords[2, k:k + m] = [h["sstart"] for h in hs]
            coords[3, k:k + m] = [h["send"] for h in hs]
            score[k:k + m] = [h["score"] for h in hs]
        ql, sl = d.get("qlen"), d.get("slen")
        if ql is None:  # alignment.py:477-480
            ql = (int(coords[1, k:k + m].max()) if m else 0) + 1
        if sl is None:  # alignment.py:481-487
            sl = (max(int(coords[2, k:k + m].max()), int(coords[3, k:k + m].max())) if m else 0) + 1
        qlen[a], slen[a] = ql, sl
        k += m
    if n and (coords.max() > INT32_MAX or coords.min() < -INT32_MAX):
        raise PackError("HSP coordinates exceed int32")
    c32 = coords.astype(np.int32)
    return ColumnarBatch(
        bg_off=np.array([0, len(bg)], dtype=np.int64), bg=bg_arr.astype(np.int32),
        aln_off=np.array([0, n_aln], dtype=np.int64), hsp_off=hsp_off, qlen=qlen, slen=slen,
        qstart=np.ascontiguousarray(c32[0]), qend=np.ascontiguousarray(c32[1]),
        sstart=np.ascontiguousarray(c32[2]), send=np.ascontiguousarray(c32[3]), score=score)
