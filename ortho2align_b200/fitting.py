"""GPU-backed fitters with the reference's plugin interface (`fitting.py:8-85`): a class built
from the background scores with `.sf`, `.isf` (plus `.cdf`, `.ppf`, `.pdf`, `.estimator` so that
`AbstractFitter.__subclasshook__` accepts it). Every evaluation runs in the CUDA library through
`o2a_fitter_eval`; there is no CPU fallback.

  HistogramFitter  ~ fitting.py:133-215 (bit-exact)
  KernelFitter     ~ fitting.py:245-356 (|dp| <= max(1e-12 p, 16*2^-53))
  EmpiricalFitter  — extension named by BASELINE.json; no reference counterpart
"""
from __future__ import annotations

import numpy as np

from .columnar import kde_factor

_engine = None


def _eng():
    global _engine
    if _engine is None:
        from .engine import Engine
        _engine = Engine(0)
    return _engine


def ranking(data):
    """`fitting.ranking` (`fitting.py:359-364`) with the build's tie policy: stable."""
    data = np.asarray(data)
    order = np.argsort(data, kind="stable")
    ranks = np.empty_like(data)
    ranks[order] = np.arange(len(data)) + 1
    return ranks


class _GpuFitter:
    kind = None

    def __init__(self, data):
        self.data = data
        arr = np.asarray(list(data))
        if arr.dtype.kind not in "iu":
            raise TypeError("the CUDA fitters take integer background scores (BLAST raw scores)")
        self._bg = arr.astype(np.int32)
        self._factor = kde_factor(len(self._bg)) if self.kind == "kde" else 1.0

    @property
    def estimator(self):
        return self

    def _call(self, x=None, q=None):
        scalar = np.isscalar(x if x is not None else q)
        arr = np.atleast_1d(np.asarray(x if x is not None else q, dtype=np.float64))
        sf, isf = _eng().fitter_eval(self.kind, self._bg, x=arr if x is not None else None,
                                     q=arr if q is not None else None, kde_factor=self._factor)
        out = sf if x is not None else isf
        return out[0] if scalar else out

    def sf(self, data):
        return self._call(x=data)

    def isf(self, data):
        return self._call(q=data)

    def cdf(self, data):
        return 1.0 - self.sf(data)

    def ppf(self, data):
        q = np.asarray(data, dtype=np.float64)
        return self.isf(1.0 - q)

    def pdf(self, data):
        raise NotImplementedError("pdf is not on the build_orthologs path (pipeline.py:755-796)")


class HistogramFitter(_GpuFitter):
    kind = "hist"


class KernelFitter(_GpuFitter):
    kind = "kde"


class EmpiricalFitter(_GpuFitter):
    kind = "empirical"


def best_chain(qstart, qend, sstart, send, score, qlen=None, slen=None, gapopen=5, gapextend=2):
    """`Alignment.best_chain` (`alignment.py:1032-1046`) on plain coordinate arrays.
    Returns (chain HSP indices in chain order, 'direct'|'reverse', chain score)."""
    qend_a, ss_a, se_a = np.asarray(qend), np.asarray(sstart), np.asarray(send)
    if qlen is None:            # alignment.py:477-480
        qlen = (int(qend_a.max()) if len(qend_a) else 0) + 1
    if slen is None:            # alignment.py:481-487
        slen = (max(int(ss_a.max()), int(se_a.max())) if len(ss_a) else 0) + 1
    chain, orient, sc, _ = _eng().best_chain(qstart, qend, sstart, send, score, qlen, slen, gapopen, gapextend)
    return chain, ("direct" if orient == 0 else "reverse"), sc
