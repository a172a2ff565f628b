"""ctypes front-end of the C restatement (oracle/o2a_oracle.c). TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
"""
from __future__ import annotations

import ctypes as C
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))

from ortho2align_b200.result import (BEST_FUNCTIONS, BEST_VALUES, FITTERS, BatchResult,  # noqa: E402
                                     compact_slots)

_lib = None


def lib():
    global _lib
    if _lib is None:
        from oracle import build as _b
        path = _b.LIB if os.path.exists(_b.LIB) and os.path.getmtime(_b.LIB) >= os.path.getmtime(_b.SRC) \
            else _b.build()
        _lib = C.CDLL(path)
        d, i64, i32 = C.c_double, C.c_int64, C.c_int
        P = C.c_void_p
        _lib.o2a_oracle_pairwise_sum.restype = d
        _lib.o2a_oracle_pairwise_sum.argtypes = [P, i64]
        _lib.o2a_oracle_interp.restype = d
        _lib.o2a_oracle_interp.argtypes = [d, P, P, i64]
        _lib.o2a_oracle_hist_fit.restype = i64
        _lib.o2a_oracle_hist_fit.argtypes = [P, i64, P, P]
        _lib.o2a_oracle_hist_sf.restype = d
        _lib.o2a_oracle_hist_sf.argtypes = [d, P, P, i64]
        _lib.o2a_oracle_hist_isf.restype = d
        _lib.o2a_oracle_hist_isf.argtypes = [d, P, P, i64]
        _lib.o2a_oracle_kde_cdf.restype = d
        _lib.o2a_oracle_kde_cdf.argtypes = [d, P, i64, d, P]
        _lib.o2a_oracle_kde_isf.restype = d
        _lib.o2a_oracle_kde_isf.argtypes = [d, P, i64, d, P, P]
        _lib.o2a_oracle_emp_sf.restype = d
        _lib.o2a_oracle_emp_sf.argtypes = [d, P, i64]
        _lib.o2a_oracle_emp_isf.restype = d
        _lib.o2a_oracle_emp_isf.argtypes = [d, P, i64]
        _lib.o2a_oracle_ranking.restype = None
        _lib.o2a_oracle_ranking.argtypes = [P, i64, P]
        _lib.o2a_oracle_best_chain.restype = i64
        _lib.o2a_oracle_best_chain.argtypes = [i64, P, P, P, P, P, P, i64, i64, d, d, P, P, P, P]
        _lib.o2a_oracle_build_batch.restype = i32
        _lib.o2a_oracle_build_batch.argtypes = [i64] + [P] * 12 + [i32, i32, d, i32, i32, i32, i32, i32] + [P] * 14
        _lib.o2a_oracle_max_threads.restype = i32
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def pairwise_sum(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return lib().o2a_oracle_pairwise_sum(_p(a), len(a))


def hist_fit(bg):
    bg = np.ascontiguousarray(bg, dtype=np.int32)
    B = len(bg) // 2
    hbins = np.empty(B + 1, dtype=np.float64)
    hcdf = np.empty(B + 1, dtype=np.float64)
    r = lib().o2a_oracle_hist_fit(_p(bg), len(bg), _p(hbins), _p(hcdf))
    if r < 0:
        raise ValueError(f"hist_fit failed ({r})")
    return hbins, hcdf


def hist_sf(x, hbins, hcdf):
    L = lib()
    return np.array([L.o2a_oracle_hist_sf(float(v), _p(hbins), _p(hcdf), len(hbins) - 1) for v in np.atleast_1d(x)])


def hist_isf(q, hbins, hcdf):
    return lib().o2a_oracle_hist_isf(float(q), _p(hbins), _p(hcdf), len(hbins) - 1)


def kde_sf(x, bg, factor):
    bg = np.ascontiguousarray(bg, dtype=np.int32)
    scratch = np.empty(len(bg), dtype=np.float64)
    L = lib()
    return np.array([1.0 - L.o2a_oracle_kde_cdf(float(v), _p(bg), len(bg), float(factor), _p(scratch))
                     for v in np.atleast_1d(x)])


def kde_isf(q, bg, factor):
    bg = np.ascontiguousarray(bg, dtype=np.int32)
    scratch = np.empty(len(bg), dtype=np.float64)
    err = C.c_int(0)
    r = lib().o2a_oracle_kde_isf(float(q), _p(bg), len(bg), float(factor), _p(scratch), C.byref(err))
    if err.value:
        raise RuntimeError("brentq failed to converge")
    return r


def emp_sf(x, bg):
    s = np.sort(np.ascontiguousarray(bg, dtype=np.int32))
    L = lib()
    return np.array([L.o2a_oracle_emp_sf(float(v), _p(s), len(s)) for v in np.atleast_1d(x)])


def emp_isf(q, bg):
    s = np.sort(np.ascontiguousarray(bg, dtype=np.int32))
    return lib().o2a_oracle_emp_isf(float(q), _p(s), len(s))


def ranking(p):
    p = np.ascontiguousarray(p, dtype=np.float64)
    r = np.empty(len(p), dtype=np.float64)
    lib().o2a_oracle_ranking(_p(p), len(p), _p(r))
    return r


def best_chain(qstart, qend, sstart, send, score, qlen, slen, gapopen=5, gapextend=2):
    """Alignment.best_chain (alignment.py:1032) on plain arrays -> (chain idx, orient, score, (sd, sr))."""
    a32 = lambda v: np.ascontiguousarray(v, dtype=np.int32)
    qstart, qend, sstart, send = a32(qstart), a32(qend), a32(sstart), a32(send)
    score = np.ascontiguousarray(score, dtype=np.float64)
    n = len(score)
    chain = np.empty(n + 2, dtype=np.int64)
    orient = C.c_int(0)
    cs = C.c_double(0)
    s2 = np.empty(2, dtype=np.float64)
    r = lib().o2a_oracle_best_chain(n, None, _p(qstart), _p(qend), _p(sstart), _p(send), _p(score),
                                    int(qlen), int(slen), float(gapopen), float(gapextend),
                                    _p(chain), C.byref(orient), C.byref(cs), _p(s2))
    if r < 0:
        raise KeyError("HSP with orientation None")
    return chain[:r].copy(), orient.value, cs.value, (float(s2[0]), float(s2[1]))


LAST_CALL_SECONDS = 0.0   # wall time of the last o2a_oracle_build_batch C call (bench.py's cpu_baseline)


def host_threads():
    """Threads the CPU baseline may use: the process's CPU affinity (torchrun exports
    OMP_NUM_THREADS=1, which must not throttle the baseline)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def build_batch(batch, fitter="hist", fdr=True, alpha=0.05, gapopen=5, gapextend=2,
                best_value="block_length", best_function="max", n_threads=0, compact=True):
    """`_build` over the whole columnar batch + best-ortholog selection. Returns BatchResult
    (or None with compact=False: only the C call runs, for timing)."""
    global LAST_CALL_SECONDS
    import time
    b = batch
    n_lnc, n_aln, n_hsp = b.n_lnc, b.n_aln, b.n_hsp
    kf = b.ensure_kde_factor() if FITTERS[fitter] == 1 else np.zeros(max(n_lnc, 1))
    thr = np.full(n_lnc, np.nan)
    status = np.zeros(n_lnc, dtype=np.int32)
    best_aln = np.full(n_lnc, -1, dtype=np.int64)
    n_surv = np.zeros(n_aln, dtype=np.int32)
    n_keep = np.zeros(n_aln, dtype=np.int32)
    chain_len = np.zeros(n_aln, dtype=np.int32)
    chain_orient = np.ones(n_aln, dtype=np.uint8)
    chain_score = np.full(n_aln, -np.inf)
    orient_scores = np.full((n_aln, 2), -np.inf)
    cap = max(n_hsp, 1)
    surv_idx = np.zeros(cap, dtype=np.int32)
    surv_p = np.zeros(cap)
    surv_pq = np.zeros(cap)
    surv_keep = np.zeros(cap, dtype=np.uint8)
    chain_idx = np.zeros(cap, dtype=np.int32)
    L = lib()
    t0 = time.perf_counter()
    rc = L.o2a_oracle_build_batch(
        n_lnc, _p(b.bg_off), _p(b.bg), _p(kf), _p(b.aln_off), _p(b.hsp_off), _p(b.qlen), _p(b.slen),
        _p(b.qstart), _p(b.qend), _p(b.sstart), _p(b.send), _p(b.score),
        FITTERS[fitter], int(bool(fdr)), float(alpha), int(gapopen), int(gapextend),
        BEST_VALUES[best_value], BEST_FUNCTIONS[best_function], int(n_threads),
        _p(thr), _p(status), _p(best_aln), _p(n_surv), _p(n_keep), _p(chain_len), _p(chain_orient),
        _p(chain_score), _p(orient_scores), _p(surv_idx), _p(surv_p), _p(surv_pq), _p(surv_keep),
        _p(chain_idx))
    LAST_CALL_SECONDS = time.perf_counter() - t0
    if rc != 0:
        raise RuntimeError(f"oracle build_batch failed ({rc})")
    if not compact:
        return None
    surv_off, (s_idx, s_p, s_pq, s_keep) = compact_slots(b.hsp_off, n_surv, surv_idx, surv_p, surv_pq, surv_keep)
    chain_off, (c_idx,) = compact_slots(b.hsp_off, chain_len, chain_idx)
    # hsp.pval of the chain HSPs: look the chain HSP up among the survivors of its alignment
    c_pq = np.empty(len(c_idx))
    for a in np.nonzero(chain_len)[0]:
        s0, s1 = surv_off[a], surv_off[a + 1]
        lut = dict(zip(s_idx[s0:s1].tolist(), s_pq[s0:s1].tolist()))
        c0, c1 = chain_off[a], chain_off[a + 1]
        c_pq[c0:c1] = [lut[i] for i in c_idx[c0:c1].tolist()]
    return BatchResult(thr, status, best_aln, n_surv, n_keep, chain_len, chain_orient, chain_score,
                       orient_scores, surv_off, s_idx, s_p, s_pq, s_keep, chain_off, c_idx, c_pq)
