"""Result of one `build_batch` call in CSR form (host numpy arrays)."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

FITTERS = {"hist": 0, "kde": 1, "empirical": 2}
BEST_VALUES = {"block_length": 0, "total_length": 1, "block_count": 2, "weight": 3}
BEST_FUNCTIONS = {"max": 0, "min": 1}

# per-lncRNA status: what the reference would have raised for this lncRNA (-> query_exceptions.bed)
STATUS_OK = 0
STATUS_ORIENTATION_NONE = 1     # KeyError/TypeError in _split_orientations (alignment.py:865-872)
STATUS_BRENTQ_NOT_CONVERGED = 2  # RuntimeError from scipy.optimize.brentq (fitting.py:239)
STATUS_TOO_MANY_BINS = 3        # ValueError from np.histogram (fitting.py:152)
STATUS_UNSUPPORTED = 4          # reserved (round 1's fitter envelope; no longer produced)


@dataclass
class BatchResult:
    thr: np.ndarray            # float64 [n_lnc]   score threshold isf(alpha)      (pipeline.py:757)
    status: np.ndarray         # int32   [n_lnc]
    best_aln: np.ndarray       # int64   [n_lnc]   global alignment index of the best ortholog, -1 none
    n_surv: np.ndarray         # int32   [n_aln]   HSPs with score > thr           (pipeline.py:765)
    n_keep: np.ndarray         # int32   [n_aln]   HSPs entering the chaining      (pipeline.py:774-776)
    chain_len: np.ndarray      # int32   [n_aln]
    chain_orient: np.ndarray   # uint8   [n_aln]   0 direct, 1 reverse
    chain_score: np.ndarray    # float64 [n_aln]   -inf when the chain is empty
    orient_scores: np.ndarray  # float64 [n_aln,2] (direct, reverse) best-chain scores
    surv_off: np.ndarray       # int64   [n_aln+1]
    surv_idx: np.ndarray       # int32   [n_surv_total]  HSP position inside its alignment
    surv_p: np.ndarray         # float64 p-value                                   (pipeline.py:766)
    surv_pq: np.ndarray        # float64 value stored in hsp.pval: q if fdr else p  (pipeline.py:772,779)
    surv_keep: np.ndarray      # uint8
    chain_off: np.ndarray      # int64   [n_aln+1]
    chain_idx: np.ndarray      # int32   [n_chain_total] HSP position inside its alignment, chain order
    chain_pq: np.ndarray       # float64 [n_chain_total] hsp.pval of the chain HSPs


def compact_slots(hsp_off, counts, *slot_arrays):
    """Slot layout (segment a at [hsp_off[a], hsp_off[a]+counts[a])) -> CSR offsets + packed arrays."""
    counts = counts.astype(np.int64)
    off = np.zeros(len(counts) + 1, dtype=np.int64)
    np.cumsum(counts, out=off[1:])
    total = int(off[-1])
    if total == 0:
        return off, [a[:0].copy() for a in slot_arrays]
    seg = np.repeat(np.arange(len(counts)), counts)
    pos = np.arange(total, dtype=np.int64) - off[:-1][seg] + hsp_off[:-1][seg]
    return off, [a[pos] for a in slot_arrays]
