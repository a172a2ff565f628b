"""Oracle for annotate_orthologs' interval join. TEST INFRASTRUCTURE ONLY.

`annotate_batch` is the ctypes wrapper of `o2a_oracle_annotate` (oracle/o2a_oracle_annotate.c): the literal
restatement of `find_neighbours` + `calc_JI` + `calc_OC` (genomicranges.py:635-685, 497-547) over columnar
lists. Pinned against the reference run in the build container (tests/test_oracle_vs_reference.py) and
tests/golden/annotate/ (oracle/make_golden_annotate.py).
"""
from __future__ import annotations

import ctypes as C
import time

import numpy as np

from . import oracle as _O

LAST_CALL_SECONDS = 0.0


def _bind():
    L = _O.lib()
    if not getattr(L, "_annotate_bound", False):
        P, i64, i32 = C.c_void_p, C.c_int64, C.c_int32
        L.o2a_oracle_annotate.restype = i32
        L.o2a_oracle_annotate.argtypes = [i64] + [P] * 5 + [i64] + [P] * 5 + [i64, i32, i32] + [P] * 4
        L._annotate_bound = True
    return L


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def annotate_batch(q, a, distance=0, strandness=True, n_threads=0):
    """q / a: dicts of chrom, start, end, name (int32 ranks / coordinates) and strand (int8: +1, -1, 0); `a` sorted
    by (chrom, start, end, name). -> dict(nb_off, nb_idx, ji, oc)."""
    global LAST_CALL_SECONDS
    L = _bind()
    cols = []
    for side in (q, a):
        cols.append([np.ascontiguousarray(side[k], dtype=np.int32) for k in ("chrom", "start", "end", "name")]
                    + [np.ascontiguousarray(side["strand"], dtype=np.int8)])
    n_q, n_a = len(cols[0][1]), len(cols[1][1])
    nb_off = np.zeros(n_q + 1, dtype=np.int64)
    args = [n_q, *[_p(x) for x in cols[0]], n_a, *[_p(x) for x in cols[1]], int(distance), int(bool(strandness)),
            int(n_threads)]
    t0 = time.perf_counter()
    rc = L.o2a_oracle_annotate(*args, _p(nb_off), None, None, None)
    if rc != 0:
        raise RuntimeError(f"oracle annotate failed ({rc})")
    m = int(nb_off[-1])
    nb_idx, ji, oc = np.zeros(max(m, 1), dtype=np.int32), np.zeros(max(m, 1)), np.zeros(max(m, 1))
    rc = L.o2a_oracle_annotate(*args, _p(nb_off), _p(nb_idx), _p(ji), _p(oc))
    LAST_CALL_SECONDS = time.perf_counter() - t0
    if rc != 0:
        raise RuntimeError(f"oracle annotate failed ({rc})")
    return dict(nb_off=nb_off, nb_idx=nb_idx[:m], ji=ji[:m], oc=oc[:m])
