"""Host-side driver of the CUDA library: one `Engine` per GPU (one process per GPU under torchrun).

`Engine.build_batch(batch, ...)` replaces the process-pool fan-out of `_build` over lncRNAs
(`pipeline.py:846-860`): the whole columnar batch goes through the C-ABI in one call.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .columnar import ColumnarBatch
from .result import BEST_FUNCTIONS, BEST_VALUES, FITTERS, BatchResult


class O2AError(RuntimeError):
    pass


def _p(a):
    return None if a is None else C.c_void_p(a.__array_interface__["data"][0])


class DeviceBatch:
    """A ColumnarBatch resident in HBM (torch tensors own the memory; the library gets raw pointers)."""

    def __init__(self, batch: ColumnarBatch, device, aos=True, score_bits=0, bg_bits=32):
        import torch
        self.n_lnc, self.n_aln, self.n_hsp, self.n_bg = batch.n_lnc, batch.n_aln, batch.n_hsp, batch.n_bg
        self.max_aln_hsps = int(np.diff(batch.hsp_off).max()) if batch.n_aln else 0
        self.max_lnc_bg = int(np.diff(batch.bg_off).max()) if batch.n_lnc else 0
        self.t = {}
        for k in ("bg_off", "bg", "aln_off", "hsp_off", "qlen", "slen", "qstart", "qend", "sstart", "send", "score"):
            self.t[k] = torch.from_numpy(getattr(batch, k)).to(device, non_blocking=False)
        self.t["kde_factor"] = torch.from_numpy(batch.ensure_kde_factor()).to(device)
        self.aos = aos
        if aos:
            self.t["coords"] = torch.from_numpy(batch.coords4()).to(device)
        self.bg_bits = bg_bits
        if bg_bits in (8, 16):
            self.t["bg"] = torch.from_numpy(batch.bg_narrow(bg_bits)).to(device)
        self.score_bits = score_bits
        if score_bits:
            q, eo, ep, ev = batch.quantized_scores(score_bits)
            for k, v in (("score_q", q), ("exc_off", eo), ("exc_pos", ep), ("exc_val", ev)):
                self.t[k] = torch.from_numpy(v).to(device)
            del self.t["score"]
        self.nbytes = sum(v.numel() * v.element_size() for v in self.t.values())

    def struct(self, with_kde):
        s = _lib.O2ABatch()
        s.n_lnc, s.n_aln, s.n_hsp, s.n_bg = self.n_lnc, self.n_aln, self.n_hsp, self.n_bg
        for k in ("bg_off", "bg", "aln_off", "hsp_off", "qlen", "slen", "qstart", "qend", "sstart", "send"):
            setattr(s, k, self.t[k].data_ptr())
        s.kde_factor = self.t["kde_factor"].data_ptr() if with_kde else None
        s.coords = self.t["coords"].data_ptr() if self.aos else None
        s.bg_bits = self.bg_bits
        s.score_bits = self.score_bits
        if self.score_bits:
            for k in ("score_q", "exc_off", "exc_pos", "exc_val"):
                setattr(s, k, self.t[k].data_ptr())
        else:
            s.score = self.t["score"].data_ptr()
        s.mem = 1
        s.max_aln_hsps, s.max_lnc_bg = self.max_aln_hsps, self.max_lnc_bg
        return s


_ENGINES = {}
_ENGINES_LOCK = None


def get_engine(device=0):
    """The process-wide persistent Engine of `device` (created on first use, closed at interpreter exit): CUDA
    context, streams and the grow-only device workspace are paid once per process, not once per call."""
    global _ENGINES_LOCK
    import atexit
    import threading
    if _ENGINES_LOCK is None:
        _ENGINES_LOCK = threading.Lock()
        atexit.register(close_engines)
    with _ENGINES_LOCK:
        eng = _ENGINES.get(int(device))
        if eng is None or not eng.h:
            eng = Engine(int(device))
            eng.lock = threading.Lock()          # a ctx is single-threaded: callers serialise on it
            _ENGINES[int(device)] = eng
        return eng


def close_engines():
    for eng in list(_ENGINES.values()):
        try:
            eng.close()
        except Exception:
            pass
    _ENGINES.clear()


class Engine:
    def __init__(self, device=0):
        self.lib = _lib.load()
        self.device = int(device)
        h = C.c_void_p()
        rc = self.lib.o2a_create(C.byref(h), self.device)
        if rc != 0:
            raise O2AError(f"o2a_create(device={device}) failed ({rc}): a CUDA device is required, "
                           "there is no CPU fallback")
        self.h = h
        self.counts = None
        self._geom = None
        self._pinned = {}        # name -> (capacity in bytes, address): reusable page-locked result buffers

    def close(self):
        if getattr(self, "h", None):
            for _, addr in self._pinned.values():
                self.lib.o2a_host_free(C.c_void_p(addr))
            self._pinned = {}
            self.lib.o2a_destroy(self.h)
            self.h = None

    def _out(self, name, n, dtype, pinned):
        """Result array: a fresh numpy array, or (pinned=True) a view of a reusable page-locked
        buffer — valid until the next fetch on this engine, D2H at full PCIe rate."""
        dtype = np.dtype(dtype)
        if not pinned:
            return np.zeros(n, dtype=dtype)
        need = max(int(n) * dtype.itemsize, 16)
        cap, addr = self._pinned.get(name, (0, 0))
        if cap < need:
            if addr:
                self.lib.o2a_host_free(C.c_void_p(addr))
            cap = need + need // 4
            addr = self.lib.o2a_host_alloc(cap)
            if not addr:
                raise O2AError("o2a_host_alloc failed")
            self._pinned[name] = (cap, addr)
        buf = (C.c_char * need).from_address(addr)
        return np.frombuffer(buf, dtype=dtype, count=int(n))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise O2AError(f"{what} failed ({rc}): {self.lib.o2a_last_error(self.h).decode()}")

    def set_stream(self, cuda_stream_ptr):
        self._check(self.lib.o2a_set_stream(self.h, C.c_void_p(cuda_stream_ptr)), "o2a_set_stream")

    def set_timing(self, on=True):
        self._check(self.lib.o2a_set_timing(self.h, 1 if on else 0), "o2a_set_timing")

    def stage_ms(self):
        ms = (C.c_float * len(_lib.STAGES))()
        self._check(self.lib.o2a_get_stage_ms(self.h, ms), "o2a_get_stage_ms")
        return dict(zip(_lib.STAGES, [float(x) for x in ms]))

    def launch_count(self):
        return int(self.lib.o2a_launch_count(self.h))

    def graph_state(self):
        """How the last `run` was issued: 0 plain launches, 1 captured into a CUDA graph and launched, 2 graph replay."""
        return int(self.lib.o2a_graph_state(self.h))

    def fp64_probe(self, kind="dfma", iters=20000):
        """Sustained FP64 rate of the device: 'dfma' -> fused multiply-adds per second, 'ndtr' -> ndtr evaluations per second."""
        v = C.c_double(0.0)
        self._check(self.lib.o2a_fp64_probe(self.h, {"dfma": 0, "ndtr": 1}[kind], int(iters), C.byref(v)), "o2a_fp64_probe")
        return float(v.value)

    COORDS_MODES = {"auto": 0, "zerocopy": 1, "gather": 2, "upload": 3}

    def set_coords_mode(self, mode="auto"):
        """How the retained HSPs' coordinate records of HOST batches reach the device (include/o2a.h, O2A_COORDS_*)."""
        self._check(self.lib.o2a_set_coords_mode(self.h, self.COORDS_MODES[mode]), "o2a_set_coords_mode")

    def coords_mode(self):
        """Mode the last host batch used."""
        k = int(self.lib.o2a_coords_mode(self.h))
        return {v: n for n, v in self.COORDS_MODES.items()}[k]

    @staticmethod
    def params(fitter="hist", fdr=True, alpha=0.05, gapopen=5, gapextend=2,
               best_value="block_length", best_function="max"):
        p = _lib.O2AParams()
        p.fitter, p.fdr, p.alpha = FITTERS[fitter], int(bool(fdr)), float(alpha)
        p.gapopen, p.gapextend = int(gapopen), int(gapextend)
        p.best_value, p.best_function = BEST_VALUES[best_value], BEST_FUNCTIONS[best_function]
        return p

    @staticmethod
    def host_struct(batch: ColumnarBatch, with_kde, coords=None, narrow=None, bg16=None):
        """`coords`: optional AoS int32 [n_hsp, 4] array (e.g. pinned) used instead of the SoA arrays.
        `narrow`: optional (bits, score_q, exc_off, exc_pos, exc_val) from `quantized_scores` (the arrays
        must stay alive until the call returns) used instead of the float64 `score`.
        `bg16`: optional int16 copy of `batch.bg` (`ColumnarBatch.bg16`) used instead of the int32 array."""
        s = _lib.O2ABatch()
        s.n_lnc, s.n_aln, s.n_hsp, s.n_bg = batch.n_lnc, batch.n_aln, batch.n_hsp, batch.n_bg
        for k in ("bg_off", "bg", "aln_off", "hsp_off", "qlen", "slen", "qstart", "qend", "sstart", "send", "score"):
            a = getattr(batch, k)
            assert a.flags["C_CONTIGUOUS"]
            setattr(s, k, a.ctypes.data)
        s.kde_factor = batch.ensure_kde_factor().ctypes.data if with_kde else None
        s.coords = coords.ctypes.data if coords is not None else None
        if bg16 is not None:           # int16 or int8 copy of the background (ColumnarBatch.bg_narrow)
            assert bg16.dtype in (np.int16, np.int8) and len(bg16) == batch.n_bg
            s.bg, s.bg_bits = bg16.ctypes.data, 8 * bg16.dtype.itemsize
        if narrow is not None:
            s.score_bits = int(narrow[0])
            s.score_q, s.exc_off = narrow[1].ctypes.data, narrow[2].ctypes.data
            s.exc_pos, s.exc_val = narrow[3].ctypes.data, narrow[4].ctypes.data
        s.mem = 0
        s._keepalive = (batch, coords, narrow, bg16)     # the struct holds raw addresses only
        s.max_aln_hsps = int(np.diff(batch.hsp_off).max()) if batch.n_aln else 0
        s.max_lnc_bg = int(np.diff(batch.bg_off).max()) if batch.n_lnc else 0
        return s

    @staticmethod
    def hostbatch_struct(hb, with_kde):
        """`o2a_batch` over a `hostbatch.HostBatch` (the pinned narrow layout the ingest returns): exactly these
        buffers cross PCIe. The struct holds raw addresses; it keeps `hb` alive."""
        s = _lib.O2ABatch()
        s.n_lnc, s.n_aln, s.n_hsp, s.n_bg = hb.n_lnc, hb.n_aln, hb.n_hsp, hb.n_bg
        for k in ("bg_off", "aln_off", "hsp_off", "qlen", "slen", "exc_off", "exc_pos", "exc_val", "score_q", "coords"):
            setattr(s, k, getattr(hb, k).ctypes.data)
        s.bg, s.bg_bits, s.score_bits = hb.bg_q.ctypes.data, hb.bg_bits, hb.score_bits
        s.kde_factor = hb.ensure_kde_factor().ctypes.data if with_kde else None
        s.mem = 0
        s.max_aln_hsps, s.max_lnc_bg = hb.max_aln_hsps, hb.max_lnc_bg
        s._keepalive = hb
        return s

    def build_host(self, hb, fitter="hist", fdr=True, alpha=0.05, gapopen=5, gapextend=2,
                   best_value="block_length", best_function="max", survivors=True, pinned=False) -> BatchResult:
        """`build_batch` for a `HostBatch`: what `pipeline.build_orthologs` calls with the ingest's output."""
        p = self.params(fitter, fdr, alpha, gapopen, gapextend, best_value, best_function)
        self.run(self.hostbatch_struct(hb, with_kde=(fitter == "kde")), p)
        return self.fetch(survivors=survivors, pinned=pinned)

    def run(self, batch_struct, params):
        """Raw call: `o2a_build_batch` on a prepared struct. Returns the counts."""
        counts = _lib.O2ACounts()
        self._check(self.lib.o2a_build_batch(self.h, C.byref(batch_struct), C.byref(params), C.byref(counts)),
                    "o2a_build_batch")
        self.counts = counts
        self._geom = (batch_struct.n_lnc, batch_struct.n_aln)
        return counts

    def fetch_chains(self, pinned=False):
        n_lnc, n_aln = self._geom
        chain_off = self._out("chain_off", n_aln + 1, np.int64, pinned)
        n = int(self.counts.n_chain_hsps)
        chain_idx = self._out("chain_idx", n, np.int32, pinned)
        chain_pq = self._out("chain_pq", n, np.float64, pinned)
        self._check(self.lib.o2a_fetch_chains(self.h, _p(chain_off), _p(chain_idx), _p(chain_pq)), "o2a_fetch_chains")
        return chain_off, chain_idx, chain_pq

    def fetch(self, survivors=True, pinned=False) -> BatchResult:
        """Results of the last run. pinned=True returns views of reusable page-locked buffers (valid
        until the next fetch on this engine) instead of fresh arrays."""
        n_lnc, n_aln = self._geom
        o = lambda name, n, dt: self._out(name, n, dt, pinned)
        thr, status, best = o("thr", n_lnc, np.float64), o("status", n_lnc, np.int32), o("best", n_lnc, np.int64)
        self._check(self.lib.o2a_fetch_lnc(self.h, _p(thr), _p(status), _p(best)), "o2a_fetch_lnc")
        n_surv, n_keep = o("n_surv", n_aln, np.int32), o("n_keep", n_aln, np.int32)
        chain_len, chain_orient = o("chain_len", n_aln, np.int32), o("chain_orient", n_aln, np.uint8)
        chain_score = o("chain_score", n_aln, np.float64)
        orient_scores = o("orient_scores", 2 * n_aln, np.float64).reshape(n_aln, 2)
        self._check(self.lib.o2a_fetch_aln(self.h, _p(n_surv), _p(n_keep), _p(chain_len), _p(chain_orient),
                                           _p(chain_score), _p(orient_scores)), "o2a_fetch_aln")
        ns = int(self.counts.n_survivors) if survivors else 0
        surv_off = o("surv_off", n_aln + 1, np.int64)
        surv_idx, surv_p, surv_pq = o("surv_idx", ns, np.int32), o("surv_p", ns, np.float64), o("surv_pq", ns, np.float64)
        surv_keep = o("surv_keep", ns, np.uint8)
        if survivors:
            self._check(self.lib.o2a_fetch_survivors(self.h, _p(surv_off), _p(surv_idx), _p(surv_p), _p(surv_pq),
                                                     _p(surv_keep)), "o2a_fetch_survivors")
        elif pinned:
            surv_off[:] = 0
        chain_off, chain_idx, chain_pq = self.fetch_chains(pinned)
        return BatchResult(thr, status, best, n_surv, n_keep, chain_len, chain_orient, chain_score, orient_scores,
                           surv_off, surv_idx, surv_p, surv_pq, surv_keep, chain_off, chain_idx, chain_pq)

    def build_batch(self, batch: ColumnarBatch, fitter="hist", fdr=True, alpha=0.05, gapopen=5, gapextend=2,
                    best_value="block_length", best_function="max", survivors=True, aos=False,
                    score_bits=0, bg_bits=32) -> BatchResult:
        """Host buffers in, host results out (the e2e boundary)."""
        p = self.params(fitter, fdr, alpha, gapopen, gapextend, best_value, best_function)
        narrow = (score_bits,) + batch.quantized_scores(score_bits) if score_bits else None
        s = self.host_struct(batch, with_kde=(fitter == "kde"), coords=batch.coords4() if aos else None, narrow=narrow,
                             bg16=batch.bg_narrow(bg_bits) if bg_bits in (8, 16) else None)
        self.run(s, p)
        return self.fetch(survivors=survivors)

    # ---- producer side: to_genomic + cut_coordinates over a batch -----------------
    CUT_STAGES = ["h2d", "count", "scan", "write"]

    @staticmethod
    def cut_struct(rel, ranges, device=False):
        """o2a_cut_input over numpy arrays (host) or torch CUDA tensors (device=True). `rel`: hsp_off,
        qstart, qend, sstart, send, score; `ranges`: q_start, q_end, q_minus, s_start, s_end, s_minus and
        optionally relative, qleft, qright, sleft, sright (per alignment; NO_BOUND entries = None)."""
        from ._lib import O2ACutInput
        s = O2ACutInput()
        keep = []

        def ptr(v, dtype):
            if v is None:
                return None
            if device:
                assert v.is_cuda and v.is_contiguous() and str(v.dtype).endswith(np.dtype(dtype).name), (v.dtype, dtype)
                keep.append(v)
                return v.data_ptr()
            a = np.ascontiguousarray(v, dtype=dtype)
            keep.append(a)
            return a.ctypes.data

        n_aln = len(rel["hsp_off"]) - 1
        s.n_aln, s.n_hsp = n_aln, len(rel["score"])
        s.hsp_off = ptr(rel["hsp_off"], np.int64)
        for k in ("qstart", "qend", "sstart", "send"):
            setattr(s, k, ptr(rel[k], np.int32))
        s.score = ptr(rel["score"], np.float64)
        s.score_is_int = ptr(rel.get("score_is_int"), np.uint8)
        for k in ("q_start", "q_end", "s_start", "s_end"):
            setattr(s, k, ptr(ranges[k], np.int64))
        s.q_minus, s.s_minus = ptr(ranges["q_minus"], np.int8), ptr(ranges["s_minus"], np.int8)
        s.relative = ptr(ranges.get("relative"), np.uint8)
        for k in ("qleft", "qright", "sleft", "sright"):
            setattr(s, k, ptr(ranges.get(k), np.int64))
        s.mem = 1 if device else 0
        s._keepalive = keep
        return s

    def run_cut(self, cut_struct):
        n = C.c_int64(0)
        self._check(self.lib.o2a_cut_batch(self.h, C.byref(cut_struct), C.byref(n)), "o2a_cut_batch")
        self.cut_n_aln, self.cut_n_kept = int(cut_struct.n_aln), int(n.value)
        return self.cut_n_kept

    def fetch_cut(self, pinned=False):
        """pinned=True: views of reusable page-locked buffers (valid until the next fetch_cut on this engine)."""
        na, nk = self.cut_n_aln, self.cut_n_kept
        o = lambda k, n, dt: self._out("cut_" + k, n, dt, pinned)
        out = dict(hsp_off=o("hsp_off", na + 1, np.int64), qstart=o("qstart", nk, np.int32), qend=o("qend", nk, np.int32),
                   sstart=o("sstart", nk, np.int32), send=o("send", nk, np.int32), score=o("score", nk, np.float64),
                   cut=o("cut", nk, np.uint8), qlen=o("qlen", na, np.int64), slen=o("slen", na, np.int64),
                   status=o("status", na, np.int32))
        self._check(self.lib.o2a_fetch_cut(self.h, *[_p(out[k]) for k in ("hsp_off", "qstart", "qend", "sstart", "send",
                                                                         "score", "cut", "qlen", "slen", "status")]),
                    "o2a_fetch_cut")
        return out

    def cut_batch(self, rel, ranges):
        """Host arrays in, host arrays out: dict(hsp_off, qstart, qend, sstart, send, score, cut, qlen, slen, status)."""
        self.run_cut(self.cut_struct(rel, ranges))
        return self.fetch_cut()

    def cut_view(self):
        """O2ABatch with the device pointers of the last cut (mem = DEVICE): fill aln_off / bg and pass to run()."""
        s = _lib.O2ABatch()
        self._check(self.lib.o2a_cut_view(self.h, C.byref(s)), "o2a_cut_view")
        return s

    def cut_ms(self):
        ms = (C.c_float * 4)()
        self._check(self.lib.o2a_get_cut_ms(self.h, ms), "o2a_get_cut_ms")
        return dict(zip(self.CUT_STAGES, [float(x) for x in ms]))

    # ---- annotate_orthologs' interval join ---------------------------------
    ANN_STAGES = ["h2d", "prefmax", "count", "write"]

    def ann_struct(self, q, a, distance=0, strandness=True, device=False):
        """q / a: dicts of chrom, start, end, name (int32) and strand (int8: +1, -1, 0); `a` sorted by
        (chrom, start, end, name). numpy arrays, or CUDA tensors when device=True."""
        s = _lib.O2AAnnInput()
        keep = []

        def ptr(v, dtype):
            if device:
                assert v.is_cuda and v.is_contiguous() and str(v.dtype).endswith(np.dtype(dtype).name), (v.dtype, dtype)
                keep.append(v)
                return v.data_ptr()
            arr = np.ascontiguousarray(v, dtype=dtype)
            keep.append(arr)
            return arr.ctypes.data

        s.n_q, s.n_a = len(q["start"]), len(a["start"])
        for side, cols in (("q", q), ("a", a)):
            for k in ("chrom", "start", "end", "name"):
                setattr(s, f"{side}_{k}", ptr(cols[k], np.int32))
            setattr(s, f"{side}_strand", ptr(cols["strand"], np.int8))
        s.distance, s.strandness, s.mem = int(distance), int(bool(strandness)), 1 if device else 0
        s._keepalive = keep
        return s

    def run_annotate(self, ann_struct):
        n = C.c_int64(0)
        self._check(self.lib.o2a_annotate_batch(self.h, C.byref(ann_struct), C.byref(n)), "o2a_annotate_batch")
        self.ann_n_q, self.ann_n_pairs = int(ann_struct.n_q), int(n.value)
        return self.ann_n_pairs

    def fetch_annotation(self, pinned=False):
        nq, npairs = self.ann_n_q, self.ann_n_pairs
        o = lambda k, n, dt: self._out("ann_" + k, n, dt, pinned)
        out = dict(nb_off=o("nb_off", nq + 1, np.int64), nb_idx=o("nb_idx", npairs, np.int32),
                   ji=o("ji", npairs, np.float64), oc=o("oc", npairs, np.float64))
        self._check(self.lib.o2a_fetch_annotation(self.h, *[_p(out[k]) for k in ("nb_off", "nb_idx", "ji", "oc")]),
                    "o2a_fetch_annotation")
        return out

    def annotate_batch(self, q, a, distance=0, strandness=True):
        """find_neighbours + calc_JI + calc_OC for every ortholog: dict(nb_off, nb_idx, ji, oc)."""
        self.run_annotate(self.ann_struct(q, a, distance, strandness))
        return self.fetch_annotation()

    def annotate_ms(self):
        ms = (C.c_float * 4)()
        self._check(self.lib.o2a_get_annotate_ms(self.h, ms), "o2a_get_annotate_ms")
        return dict(zip(self.ANN_STAGES, [float(x) for x in ms]))

    # ---- operator-level seams -------------------------------------------------------------------
    def fitter_eval(self, fitter, bg, x=None, q=None, kde_factor=1.0):
        bg = np.ascontiguousarray(bg, dtype=np.int32)
        x = np.zeros(0) if x is None else np.ascontiguousarray(x, dtype=np.float64)
        q = np.zeros(0) if q is None else np.ascontiguousarray(q, dtype=np.float64)
        sf = np.zeros(len(x)); isf = np.zeros(len(q))
        st = C.c_int32(0)
        rc = self.lib.o2a_fitter_eval(self.h, FITTERS[fitter], _p(bg), len(bg), float(kde_factor),
                                      _p(x), len(x), _p(sf), _p(q), len(q), _p(isf), C.byref(st))
        if rc != 0:
            raise O2AError(f"o2a_fitter_eval failed ({rc}, status {st.value}): {self.lib.o2a_last_error(self.h).decode()}")
        return sf, isf

    def best_chain(self, qstart, qend, sstart, send, score, qlen, slen, gapopen=5, gapextend=2):
        a32 = lambda v: np.ascontiguousarray(v, dtype=np.int32)
        qstart, qend, sstart, send = a32(qstart), a32(qend), a32(sstart), a32(send)
        score = np.ascontiguousarray(score, dtype=np.float64)
        n = len(score)
        chain = np.zeros(max(n, 1), dtype=np.int32)
        clen = C.c_int64(0); orient = C.c_int32(0); cs = C.c_double(0); st = C.c_int32(0)
        s2 = np.zeros(2)
        self._check(self.lib.o2a_best_chain(self.h, n, _p(qstart), _p(qend), _p(sstart), _p(send), _p(score),
                                            int(qlen), int(slen), int(gapopen), int(gapextend), _p(chain),
                                            C.byref(clen), C.byref(orient), C.byref(cs), _p(s2), C.byref(st)),
                    "o2a_best_chain")
        if st.value == 1:
            raise KeyError("HSP with orientation None (alignment.py:865-872)")
        return chain[:clen.value].copy(), orient.value, cs.value, (float(s2[0]), float(s2[1]))
