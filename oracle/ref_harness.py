"""Drive the UNMODIFIED reference (imported from /root/reference) — container-only tooling.

TEST INFRASTRUCTURE. Nothing here ships or runs on the GPU box: `/root/reference` does not exist
there. This module is used (a) by `oracle/make_golden.py` to record golden input/output vectors
under `tests/golden/`, and (b) by the `-m "not gpu"` tests that can see `/root/reference` to pin
the C restatement (`oracle/o2a_oracle.c`) against the real thing.

`pebble` is not installable here (no network); the reference only uses
`ProcessPool(max_workers).map(f, iterable, timeout).result()` + `.close()` + `.join()`
(`parallel.py:352, 380-381, 414-420`), which the stand-in below provides on top of
`multiprocessing` (fork). Per-task exceptions surface from `next(iterator)` like pebble's.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("O2A_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "ortho2align"))


class _ResultIterator:
    def __init__(self, async_results, timeout):
        self._it = iter(async_results)
        self._timeout = timeout

    def __iter__(self):
        return self

    def __next__(self):
        r = next(self._it)           # StopIteration ends the stream
        return r.get(self._timeout)  # re-raises the task's exception (or mp.TimeoutError)


class _Future:
    def __init__(self, iterator):
        self._iterator = iterator

    def result(self):
        return self._iterator


class _StandInProcessPool:
    """Minimal pebble.ProcessPool look-alike (fork pool)."""

    def __init__(self, max_workers=None):
        self._pool = mp.get_context("fork").Pool(max_workers)

    def map(self, function, iterable, timeout=None):
        results = [self._pool.apply_async(function, (item,)) for item in iterable]
        return _Future(_ResultIterator(results, timeout))

    def close(self):
        self._pool.close()

    def join(self):
        self._pool.join()


class _SerialIterator:
    def __init__(self, function, iterable):
        self._function = function
        self._it = iter(iterable)

    def __iter__(self):
        return self

    def __next__(self):
        item = next(self._it)
        return self._function(item)   # a task's exception surfaces here, like pebble's


class _SerialPool:
    """In-process stand-in: same API, no fork, no pickling.

    Needed because results that carry a `GenomicRangesList` (every "dropped" lncRNA: its qrange
    holds relations['lifted'], `pipeline.py:789-795`) cannot be UNpickled in the parent:
    `SortedKeyList.__reduce__` re-calls the class as `cls(values, key)` while the reference
    overrides `__init__(self, collection[, sequence_file_path])` (`genomicranges.py:1581, 1908`),
    so the key function lands in `sequence_file_path` and `Path(function)` raises TypeError inside
    the pool's result thread — the parent then waits forever. The reference's own process
    boundary is therefore unusable for such inputs; parity for them is defined in-process.
    """

    def __init__(self, max_workers=None):
        self.max_workers = max_workers

    def map(self, function, iterable, timeout=None):
        return _Future(_SerialIterator(function, iterable))

    def close(self):
        pass

    def join(self):
        pass


def install_pebble_standin(mode=None):
    """mode: 'serial' (default; safe for every input) or 'fork' (real processes; hangs on inputs
    with dropped lncRNAs, see _SerialPool)."""
    mode = mode or os.environ.get("O2A_PEBBLE_STANDIN", "serial")
    pool = {"serial": _SerialPool, "fork": _StandInProcessPool}[mode]
    if "pebble" not in sys.modules:
        mod = types.ModuleType("pebble")
        mod.__doc__ = "stand-in for pebble (not installable offline); see oracle/ref_harness.py"
        sys.modules["pebble"] = mod
    sys.modules["pebble"].ProcessPool = pool
    par = sys.modules.get("ortho2align.parallel")
    if par is not None:
        par.ProcessPool = pool


def import_reference(pool=None):
    """Returns the reference package modules (pipeline, fitting, alignment, genomicranges)."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    install_pebble_standin(pool)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import ortho2align.alignment as alignment
    import ortho2align.fitting as fitting
    import ortho2align.genomicranges as genomicranges
    import ortho2align.pipeline as pipeline
    return types.SimpleNamespace(pipeline=pipeline, fitting=fitting, alignment=alignment,
                                 genomicranges=genomicranges)


def run_reference_lncrna(alignment_dicts, query_bg, fitting="hist", fdr=True, alpha=0.05,
                         gapopen=5, gapextend=2, stable_ties=True):
    """One lncRNA through the reference's own classes, step by step as `_build` does
    (`pipeline.py:746-796`), recording every intermediate the parity tests compare.

    HSPs are tagged with their position in the alignment's JSON list through `kwargs`
    (kept by `HSP.copy`, `alignment.py:251-265`) so chains come back as index lists.
    `stable_ties=True` patches `np.argsort` used by `ranking` (`fitting.py:361`) to
    `kind='stable'` — the tie policy of SURVEY.md §8a row 7; False runs numpy's default.
    """
    import copy

    import numpy as np
    ref = import_reference()
    dicts = copy.deepcopy(alignment_dicts)
    for d in dicts:
        for i, h in enumerate(d["HSPs"]):
            h["kwargs"] = {"i": i}
    alns = [ref.genomicranges.GenomicRangesAlignment.from_dict(d) for d in dicts]
    bg = [0, 1] if query_bg is None else list(query_bg)
    if len(set(bg)) <= 1:
        bg += [1]
    if bg[0] == 1:
        bg.append(0)
    fitter = {"hist": ref.fitting.HistogramFitter, "kde": ref.fitting.KernelFitter}[fitting]
    background = fitter(bg)
    thr = float(background.isf(alpha))
    out = {"thr": thr, "alignments": []}
    orig_argsort = np.argsort
    for aln in alns:
        aln.filter_by_score(thr)
        surv = [h.kwargs["i"] for h in aln.HSPs]
        pvals = np.asarray(background.sf([h.score for h in aln.HSPs]), dtype=float).reshape(-1)
        rec = {"surv": surv, "p": [float(x) for x in pvals]}
        if fdr:
            total = len(aln._all_HSPs)
            if stable_ties:
                ref.fitting.np.argsort = lambda a, *args, **kw: orig_argsort(a, kind="stable")
            try:
                ranks = ref.fitting.ranking(pvals)
            finally:
                ref.fitting.np.argsort = orig_argsort
            qvals = total * pvals / ranks
            for h, q in zip(aln.HSPs, qvals):
                h.pval = q
            retain = qvals <= alpha
            rec["rank"] = [float(x) for x in ranks]
            rec["pq"] = [float(x) for x in qvals]
            rec["keep"] = [bool(x) for x in retain]
            aln.filter_by_bool_array(retain)
        else:
            for h, p in zip(aln.HSPs, pvals):
                h.pval = p
            rec["pq"] = rec["p"]
            rec["keep"] = [True] * len(surv)
        aln.srange.name = aln.qrange.name
        chains = aln.get_best_chains(gapopen=gapopen, gapextend=gapextend)
        best = chains["direct"] if chains["direct"].score > chains["reverse"].score else chains["reverse"]
        rec["chain"] = [h.kwargs["i"] for h in best.HSPs]
        rec["orient"] = 0 if best is chains["direct"] else 1
        rec["chain_score"] = float(best.score)
        rec["score_direct"] = float(chains["direct"].score)
        rec["score_reverse"] = float(chains["reverse"].score)
        try:
            tot = best.to_record(mode="str")
            bed = best.to_bed12(mode="str")
            rec["records"] = [tot[0], bed[0], tot[1], bed[1]]
        except ValueError:
            rec["records"] = None
        out["alignments"].append(rec)
    return out


def run_reference_build(align_dir, bg_dir, outdir, fitting="hist", fdr=True, threshold=0.05,
                        gapopen=5, gapextend=2, cores=1, best_value="block_length", best_function="max",
                        pool="serial"):
    """The reference's `build_orthologs` + `get_best_orthologs` (`pipeline.py:801, 959`) on
    directories, through its own TimeoutProcessPool (pebble stand-in). Returns outdir."""
    ref = import_reference(pool)
    ref.pipeline.build_orthologs(align_dir, bg_dir, fitting, outdir, threshold=threshold,
                                 gapopen=gapopen, gapextend=gapextend, fdr=fdr, cores=cores,
                                 timeout=None, silent=True,
                                 stats_filename=os.path.join(outdir, "..", os.path.basename(outdir) + ".stats.txt"))
    j = lambda n: os.path.join(outdir, n)
    ref.pipeline.get_best_orthologs(
        j("significant.query_orthologs.bed"), j("significant.query_orthologs.tsv"),
        j("significant.subject_orthologs.bed"), j("significant.subject_orthologs.tsv"),
        best_value, best_function,
        j("bestSignificant.query_orthologs.bed"), j("bestSignificant.query_orthologs.tsv"),
        j("bestSignificant.subject_orthologs.bed"), j("bestSignificant.subject_orthologs.tsv"))
    return outdir
