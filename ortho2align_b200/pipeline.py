"""Drop-in `build_orthologs` / `get_best_orthologs` (same names, arguments, files and stats block as
`ortho2align/pipeline.py:801-956, 959-1024`) with the per-lncRNA `_build` workers
(`pipeline.py:722-798`) replaced by ONE batched call into the CUDA library per GPU.

What stays on the host, in Python, like in the reference: directory listing, JSON decoding, text
assembly of the BED/TSV records, the "dropped" / "exception" bookkeeping and the stats block.
There is no CPU fallback for the compute: `Engine` raises if the CUDA library or a GPU is missing.
"""
from __future__ import annotations

import json
import os
import re
import threading
import time
from collections import Counter
from pathlib import Path

import numpy as np

from .columnar import PackError, concat_batches, pack_lncrna
from .records import (Range, as_pvals, bed6_lines, chain_records, drop_duplicates, find_neighbours,
                      sort_key)
from .result import STATUS_OK

_STATUS_TEXT = {1: "None", 2: "Failed to converge after 100 iterations.",
                3: "Too many bins for data range. Cannot create finite-sized bins.",
                4: "background outside the range supported by the CUDA fitter"}


class ExceptionLogger(Exception):
    """Same shape as the reference's `parallel.ExceptionLogger` (`parallel.py:491-520`)."""

    def __init__(self, exception, variable, message=""):
        self.exception, self.variable, self.message = exception, variable, message

    def __str__(self):
        return f'The exception "{self.exception}" has been raised. ' \
               f'Important variable is {self.variable}. {self.message}'


def simple_hist(data):
    """`utils.simple_hist` (`utils.py:89-101`)."""
    counts = Counter(data)
    return "\n".join(f"{k}: {v}" for k, v in sorted(counts.items(), key=lambda i: (i[0], i[1])))


def lpt_shards(weights, n_shards):
    """Longest-processing-time greedy partition of lncRNAs by HSP count.
    Returns a list of index lists; lncRNAs are independent, so any partition is exact."""
    order = np.argsort(-np.asarray(weights), kind="stable")
    loads = [0] * n_shards
    shards = [[] for _ in range(n_shards)]
    for i in order:
        k = loads.index(min(loads))
        shards[k].append(int(i))
        loads[k] += int(weights[i])
    return [sorted(s) for s in shards]


def _load_lncrna(align_file, bg_file):
    """JSON -> (alignment dicts, ColumnarBatch) — the ingest half of `_build` (pipeline.py:743-754)."""
    with open(align_file, "r") as f:
        alignment_dicts = json.load(f)
    if os.path.exists(bg_file):
        with open(bg_file, "r") as f:
            bg = json.load(f)
    else:
        bg = None
    return alignment_dicts, pack_lncrna(alignment_dicts, bg)


def _ingest_json(alignments_files, bg_files):
    out = []
    for af, bf in zip(alignments_files, bg_files):
        alignment_dicts = None
        try:
            alignment_dicts, batch = _load_lncrna(af, bf)
        except PackError as e:
            qr = Range.from_dict(alignment_dicts[0]["qrange"]) if alignment_dicts else None
            out.append(ExceptionLogger(e, qr))
            continue
        d = alignment_dicts
        out.append((batch, 0, len(d), (lambda k, d=d: d[k]["qrange"]), (lambda k, d=d: d[k]["srange"]),
                    (lambda k, idx, d=d: [d[k]["HSPs"][int(j)] for j in idx]), None,
                    (lambda k, d=d: bool(d[k].get("relative")))))
    return out


PHASES = {}      # wall seconds of the phases of the last build_orthologs call (tools/bench_e2e_files.py)


def _ingest_native(alignments_files, bg_files, threads, cache=None):
    from . import ingest as native
    r = native.open_sidecar(cache, alignments_files, bg_files) if cache else None
    if r is None:
        r = native.load_files(alignments_files, bg_files, threads=threads or 1)
        if cache and not r.file_status.any():
            t0 = time.perf_counter()
            native.save_sidecar(cache, r, alignments_files, bg_files)
            PHASES["sidecar_write"] = time.perf_counter() - t0
    out = [None] * len(alignments_files)
    for i in np.nonzero(r.file_status)[0]:
        if r.file_status[i] == 1:
            raise ValueError(f"{alignments_files[i]}: {r.errors[int(i)]}")
        # values outside the columnar types: report like the Python packer does
        with open(alignments_files[i]) as f:
            d = json.load(f)
        out[i] = ExceptionLogger(PackError(r.errors[int(i)]), Range.from_dict(d[0]["qrange"]) if d else None)
    return _NativeLoaded(r, out)


class _NativeLoaded:
    """The per-file items of a native ingest, built on demand: with 10^5 lncRNAs the tuples and accessor closures of every
    file cost more than parsing them, and the columnar assembly (`_assemble_columnar`) never looks at one. Indexing and
    iteration give what the list of `_ingest_json` holds: an ExceptionLogger, or (source batch, lncRNA index in it,
    n_alignments, qrange_of, srange_of, hsps_of, (columnar source, first alignment), relative_of)."""

    def __init__(self, r, errors):
        self.r, self.b = r, r.host           # the pinned narrow layout: handed to the library as is
        self.errors = {i: x for i, x in enumerate(errors) if x is not None}
        self.n = len(errors)
        lnc_file = np.asarray(r.lnc_file)
        self.lnc_of_file = np.full(self.n, -1, dtype=np.int64)
        self.lnc_of_file[lnc_file] = np.arange(len(lnc_file))

    def __len__(self):
        return self.n

    def ok_indices(self):
        return np.nonzero(self.lnc_of_file >= 0)[0].tolist()

    def __getitem__(self, i):
        if i in self.errors:
            return self.errors[i]
        r, b, l = self.r, self.b, int(self.lnc_of_file[i])
        if l < 0:
            return None
        a0, a1 = int(b.aln_off[l]), int(b.aln_off[l + 1])
        hsp_off, rel = b.hsp_off, r.relative
        return (b, l, a1 - a0,
                (lambda k: r.range_dict(a0 + k, "q")), (lambda k: r.range_dict(a0 + k, "s")),
                (lambda k, idx: r.hsp_dicts(int(hsp_off[a0 + k]) + idx)),
                (r, a0),                     # columnar source of the lncRNA's HSPs, for the native record formatter
                (lambda k: bool(rel[a0 + k])))

    def __iter__(self):
        return (self[i] for i in range(self.n))


def _run_shards(batches, devices, fitter, fdr, alpha, gapopen, gapextend):
    """One persistent Engine (= one o2a_ctx, `engine.get_engine`) per GPU, each driven from its own thread; ctypes
    releases the GIL. A `HostBatch` (the ingest's pinned narrow layout) goes to the library as is; a plain
    `ColumnarBatch` (the Python JSON reader, multi-GPU shards) is packed into that layout first."""
    from .engine import get_engine
    from .hostbatch import HostBatch
    results = [None] * len(batches)
    errors = []

    def work(k):
        try:
            eng = get_engine(devices[k])
            hb = batches[k] if isinstance(batches[k], HostBatch) else HostBatch.from_columnar(batches[k])
            with eng.lock:
                results[k] = eng.build_host(hb, fitter, fdr, alpha, gapopen, gapextend, survivors=False)
        except Exception as e:  # surfaced below, on the caller's thread
            errors.append(e)

    if len(batches) == 1:
        work(0)
    else:
        threads = [threading.Thread(target=work, args=(k,)) for k in range(len(batches))]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
    if errors:
        raise errors[0]
    return results


def _format_chains_native(chain_desc, strings, qplus, src, threads):
    """The four ortholog files as bytes (query TSV, query BED12, subject TSV, subject BED12) for the chains collected by
    build_orthologs, in their order: one gather per column from the ingest's columnar batch, one library call."""
    from . import records_native
    n = len(chain_desc)
    ids, results = {}, []
    res_id = np.empty(n, dtype=np.int64)
    for k, d in enumerate(chain_desc):
        r = d[0]
        if id(r) not in ids:
            ids[id(r)] = len(results)
            results.append(r)
        res_id[k] = ids[id(r)]
    c0 = np.fromiter((d[1] for d in chain_desc), dtype=np.int64, count=n)
    c1 = np.fromiter((d[2] for d in chain_desc), dtype=np.int64, count=n)
    base = np.fromiter((d[3] for d in chain_desc), dtype=np.int64, count=n)
    lens = c1 - c0
    hsp_off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens, out=hsp_off[1:])
    gidx = np.empty(int(hsp_off[-1]), dtype=np.int64)
    pval = np.empty(int(hsp_off[-1]), dtype=np.float64)
    for rid, res in enumerate(results):
        sel = np.nonzero(res_id == rid)[0]
        l = lens[sel]
        within = np.arange(int(l.sum()), dtype=np.int64) - np.repeat(np.cumsum(l) - l, l)   # 0..len-1 inside each chain
        src_pos = np.repeat(c0[sel], l) + within
        dst_pos = np.repeat(hsp_off[sel], l) + within
        gidx[dst_pos] = np.asarray(res.chain_idx)[src_pos].astype(np.int64) + np.repeat(base[sel], l)
        pval[dst_pos] = np.asarray(res.chain_pq)[src_pos]
    *cols, score = src.gather_hsps(gidx)
    return records_native.format_records(hsp_off, *cols, score, np.asarray(src.score_is_int)[gidx],
                                         pval, np.asarray(qplus, dtype=np.uint8), strings, threads=threads or 1)


# A range object exactly as `json.dump(GenomicRange.to_dict())` writes it when it has no relations (every subject range of
# an alignment file): fixed key order, plain ASCII strings. Anything else (escapes, other key order, nested relations)
# does not match and goes through json.loads + Range.from_dict.
_PLAIN_RANGE = re.compile(
    rb'\{"chrom": "([^"\\]*)", "start": (-?\d+), "end": (-?\d+), "strand": "([^"\\]*)", "genome": (null|"[^"\\]*"), '
    rb'"sequence_file_path": \{"path": (?:null|"[^"\\]*")\}, "name": (null|"[^"\\]*"), "relations": \{\}, '
    rb'"relations_class": null\}\Z')


def _plain_range(text):
    """`Range.from_dict(json.loads(text))` for the common flat form, or None."""
    m = _PLAIN_RANGE.match(text)
    if m is None:
        return None
    chrom, start, end, strand, genome, name = m.groups()
    return Range(chrom.decode(), int(start), int(end), strand.decode(),
                 None if name == b"null" else name[1:-1].decode(), None if genome == b"null" else genome[1:-1].decode(), None)


_DROPPED_PER_WORKER = 2000     # dropped lncRNAs per forked worker below which the pool is not worth its start-up
_DROP_CTX = None       # (ingest result, aln_off) of the run whose dropped lncRNAs are being rendered; inherited by fork


def _dropped_chunk(lncs):
    """What build_orthologs writes for the lncRNAs `lncs` that ended without an ortholog (pipeline.py:789-791, 895-908), as
    plain tuples: per lncRNA (sort key, BED6 line) of its query range; per neighbour found by
    `find_neighbours(srange, qrange.lifted)` (sort key, the fields `same()` compares, a token that is equal exactly when
    the `lifted` lists compare equal, BED6 line). Runs in the calling process or in forked workers (`_render_dropped`)."""
    r, aln_off = _DROP_CTX
    q_items, s_items = [], []
    for l in lncs:
        a0, a1 = int(aln_off[l]), int(aln_off[l + 1])
        q_dropped, s_dropped, last_qd, qrange = None, [], None, None
        for a in range(a0, a1):                              # pipeline.py:789-791, per alignment
            qd = r.range_dict(a, "q")
            if qd is not last_qd:
                last_qd, qrange = qd, Range.from_dict(qd)
            srange = _plain_range(r.range_text(a, "s")) or Range.from_dict(r.range_dict(a, "s"))
            srange.name = qrange.name
            q_dropped = qrange
            s_dropped.append(srange)
        q_items.append((sort_key(q_dropped), bed6_lines([q_dropped])[0]))
        ids = None
        for grange in s_dropped:
            for nb in find_neighbours(grange, q_dropped.lifted):
                if nb.lifted is None:
                    token = None
                elif not nb.lifted:
                    token = ()                               # empty lists compare equal whoever owns them
                else:                                        # Range has no __eq__: lists are equal iff it is the same object
                    if ids is None:
                        ids = {id(x): k for k, x in enumerate(q_dropped.lifted)}
                    token = (l, ids[id(nb)])
                s_items.append((sort_key(nb), (nb.chrom, nb.start, nb.end, nb.strand, nb.name, nb.genome), token,
                                bed6_lines([nb])[0]))
    return q_items, s_items


def _render_dropped(r, aln_off, lncs, cores):
    """The lines of insignificant.query_orthologs.bed / insignificant.subject_orthologs.bed for the dropped lncRNAs `lncs`
    (file order): sorted by `_genomic_key_func`, the subject side without consecutive duplicates
    (`drop_duplicates`, genomicranges.py:1853-1865). The per-lncRNA work (JSON decode of the range objects,
    `find_neighbours`) is Python; from a few thousand lncRNAs on it is spread over forked workers that return tuples."""
    global _DROP_CTX
    _DROP_CTX = (r, aln_off)
    if lncs and hasattr(r, "_ranges"):
        r._ranges()                                           # the range texts are exported once, here, not once per worker
    try:
        n_proc = min(int(cores or 1), 8, len(lncs) // _DROPPED_PER_WORKER)      # a fork of a genome-scale process costs ~25 ms
        if n_proc > 1:
            import multiprocessing as mp
            step = -(-len(lncs) // (4 * n_proc))
            chunks = [lncs[i:i + step] for i in range(0, len(lncs), step)]
            with mp.get_context("fork").Pool(n_proc) as pool:
                parts = pool.map(_dropped_chunk, chunks)
        else:
            parts = [_dropped_chunk(lncs)]
    finally:
        _DROP_CTX = None
    q_items = [x for part in parts for x in part[0]]
    s_items = [x for part in parts for x in part[1]]
    q_items.sort(key=lambda x: x[0])                          # list.sort is stable, like sorted()
    s_items.sort(key=lambda x: x[0])
    s_lines, last = [], None
    for key, same, token, line in s_items:
        if last is not None and same == last[0] and token == last[1]:
            continue
        s_lines.append(line)
        last = (same, token)
    return [x[1] for x in q_items], s_lines


def _assemble_columnar(r, res, loaded_errors, n_files, threads):
    """The records of a whole run from columnar data, without a Python object per alignment: the native ingest's range
    fields (`IngestResult.range_fields`) + the batch result -> (native_text, dist_found, dropped = (query lines, subject lines)
    of the two insignificant.* files, exceptions in file order), or None when some range object needs a full JSON decode (the caller then
    takes the per-alignment path). `r`: IngestResult / SidecarResult whose `.host` batch is what `res` was computed on;
    `loaded_errors`: {file index: ExceptionLogger} of the files that did not load."""
    from . import records_native
    if not hasattr(r, "range_fields"):
        return None
    name_off, blob, q_plus, simple = r.range_fields()
    b = r.host
    n_lnc, n_aln = b.n_lnc, b.n_aln
    aln_off = np.asarray(b.aln_off)
    lnc_of_aln = np.repeat(np.arange(n_lnc), np.diff(aln_off))
    ok_lnc = np.asarray(res.status) == STATUS_OK
    chain_off = np.asarray(res.chain_off)
    has_chain = (np.diff(chain_off) > 0) & (np.asarray(r.relative) == 0) & ok_lnc[lnc_of_aln]     # relative: to_record raises
    csum = np.concatenate([[0], np.cumsum(has_chain)])
    n_orth = csum[aln_off[1:]] - csum[aln_off[:-1]]
    A = np.nonzero(has_chain)[0]
    if len(A) and not simple[A].all():
        return None
    native_text = None
    if len(A):
        c0, c1 = chain_off[A], chain_off[A + 1]
        lens = c1 - c0
        hsp_off = np.zeros(len(A) + 1, dtype=np.int64)
        np.cumsum(lens, out=hsp_off[1:])
        within = np.arange(int(hsp_off[-1]), dtype=np.int64) - np.repeat(hsp_off[:-1], lens)
        src_pos = np.repeat(c0, lens) + within
        gidx = np.asarray(res.chain_idx)[src_pos].astype(np.int64) + np.repeat(np.asarray(b.hsp_off)[A], lens)
        pval = np.asarray(res.chain_pq)[src_pos]
        # the three strings of every chain, re-packed back to back
        k3 = (3 * A[:, None] + np.arange(3)[None, :]).reshape(-1)
        s_len = name_off[k3 + 1] - name_off[k3]
        str_off = np.zeros(len(k3) + 1, dtype=np.int64)
        np.cumsum(s_len, out=str_off[1:])
        sblob = blob[np.repeat(name_off[k3] - str_off[:-1], s_len) + np.arange(int(str_off[-1]), dtype=np.int64)] \
            if int(str_off[-1]) else np.zeros(0, dtype=np.uint8)
        *cols, score = r.gather_hsps(gidx)
        native_text = records_native.format_records(hsp_off, *cols, score, np.asarray(r.score_is_int)[gidx], pval,
                                                    q_plus[A], (str_off, sblob), threads=threads or 1)
    # what needs Python objects: lncRNAs without any ortholog (dropped) and lncRNAs the reference would have aborted
    exceptions = dict(loaded_errors)
    lnc_file = np.asarray(r.lnc_file)
    for l in np.nonzero(~ok_lnc)[0].tolist():
        st = int(res.status[l])
        err = KeyError(None) if st == 1 else (RuntimeError if st == 2 else ValueError)(_STATUS_TEXT.get(st, st))
        a0, a1 = int(aln_off[l]), int(aln_off[l + 1])
        exceptions[int(lnc_file[l])] = ExceptionLogger(err, Range.from_dict(r.range_dict(a0, "q")) if a1 > a0 else None)
    dropped = _render_dropped(r, aln_off, np.nonzero(ok_lnc & (n_orth == 0))[0].tolist(), threads)
    dist_found = n_orth[ok_lnc].tolist()
    return native_text, dist_found, dropped, [exceptions[i] for i in sorted(exceptions)]


def build_orthologs(alignments, background, fitting, outdir, threshold=0.05, gapopen=5, gapextend=2,
                    fdr=False, cores=1, timeout=None, silent=False, stats_filename=None, devices=None,
                    ingest="native", cache=None, records="native"):
    """A build_orthologs step of the pipeline (signature of `pipeline.py:801-812`).

    `cores` = threads of the native JSON ingest (the per-lncRNA process pool it configured in the
    reference does not exist here); `timeout` is accepted for CLI compatibility. `devices`: CUDA
    device indices (default: [0]); with several, lncRNAs are LPT-sharded by HSP count and every
    GPU runs its shard independently. `ingest`: 'native' (libo2a_ingest.so, C++) or 'json'
    (json.load + Python packer; same results, used to cross-check the native reader).
    `fitting`: 'hist' | 'kde' (reference) or 'empirical' (extension). `cache`: path of a columnar
    sidecar (ingest.save_sidecar): written after the first native parse of these directories, and read
    back instead of the JSON on later runs as long as file names, sizes and mtimes are unchanged.
    `records`: 'native' (with the native ingest: the four ortholog files are formatted from the columnar arrays by
    libo2a_ingest.so, include/o2a_records.h) or 'python' (records.chain_records per chain; same bytes).
    """
    alignments_dir = Path(alignments)
    background_dir = Path(background)
    if fitting not in ("hist", "kde", "empirical"):
        # the reference leaves `fitter` unbound and dies with UnboundLocalError (pipeline.py:839-847)
        raise UnboundLocalError("cannot access local variable 'fitter' where it is not associated with a value")
    devices = [0] if not devices else list(devices)

    # (the same strings os.path.join / basename give, without 2 x 10^5 calls at genome scale)
    names = [fn for fn in os.listdir(alignments_dir) if fn.endswith("json")]
    a_prefix, b_prefix = os.path.join(str(alignments_dir), ""), os.path.join(str(background_dir), "")
    alignments_files = [a_prefix + fn for fn in names]
    total_alignments = len(alignments_files)
    bg_files = [b_prefix + fn for fn in names]
    total_bgs = len(bg_files)

    # ---- ingest -------------------------------------------------------------------------------
    # per file: an ExceptionLogger, or (source batch, lncRNA index in it, n_alignments, accessors)
    t_phase = time.perf_counter()
    loaded = _ingest_native(alignments_files, bg_files, cores, cache) if ingest == "native" \
        else _ingest_json(alignments_files, bg_files)
    native_loaded = isinstance(loaded, _NativeLoaded)
    ok_idx = loaded.ok_indices() if native_loaded else [i for i, x in enumerate(loaded) if not isinstance(x, ExceptionLogger)]

    PHASES["ingest"] = time.perf_counter() - t_phase - PHASES.get("sidecar_write", 0.0)
    t_phase = time.perf_counter()
    # ---- compute: shard by lncRNA across GPUs, one batched call each ---------------------------------
    per_lnc = {}
    if ok_idx:
        n_dev = min(len(devices), len(ok_idx))
        if native_loaded:                                       # every lncRNA lives in one batch
            src, src_l, one_source = [loaded.b], loaded.lnc_of_file[ok_idx].tolist(), True
        else:
            src, src_l = [loaded[i][0] for i in ok_idx], [loaded[i][1] for i in ok_idx]
            one_source = all(x is src[0] for x in src)
        if n_dev == 1:
            shards = [list(range(len(ok_idx)))]
        elif one_source:
            per_lnc_hsps = src[0].hsp_off[src[0].aln_off[1:]] - src[0].hsp_off[src[0].aln_off[:-1]]
            shards = lpt_shards((per_lnc_hsps[src_l] + 1).tolist(), n_dev)
        else:
            shards = lpt_shards([x.n_hsp + 1 for x in src], n_dev)
        if one_source and n_dev == 1 and src_l == list(range(src[0].n_lnc)):
            batches = [src[0]]                                  # the ingest's buffers, untouched
        elif one_source:
            columnar = src[0].to_columnar() if hasattr(src[0], "to_columnar") else src[0]
            batches = [columnar.take_lnc([src_l[j] for j in shard]) for shard in shards]
        else:
            batches = [concat_batches([src[j] for j in shard]) for shard in shards]
        results = _run_shards(batches, devices[:n_dev], fitting, fdr, threshold, gapopen, gapextend)

    PHASES["compute"] = time.perf_counter() - t_phase
    t_phase = time.perf_counter()
    # ---- assemble exactly what the pool would have returned ---------------------------------------------
    fast = None
    if native_loaded and records == "native" and ok_idx and len(batches) == 1 and batches[0] is loaded.b:
        fast = _assemble_columnar(loaded.r, results[0], loaded.errors, len(loaded), cores)
    orthologs, exceptions = [], []
    if fast is not None:
        native_text_fast, dist_found_fast, dropped_fast, exceptions = fast
        loaded = []                                                   # nothing left for the per-alignment path
    elif ok_idx:
        for shard, batch, res in zip(shards, batches, results):
            for local, j in enumerate(shard):
                per_lnc[ok_idx[j]] = (batch, res, local)
    # native record formatting: per chain (result object, chain slice, first HSP of its alignment in the source batch)
    chain_desc, chain_strings, chain_qplus, chain_src = [], [], [], None
    for i, item in enumerate(loaded):
        if isinstance(item, ExceptionLogger):
            exceptions.append(item)
            continue
        _, _, n_alns, qrange_of, srange_of, hsps_of, columnar_src, relative_of = item
        if records != "native":
            columnar_src = None
        batch, res, l = per_lnc[i]
        if res.status[l] != STATUS_OK:
            st = int(res.status[l])
            err = KeyError(None) if st == 1 else (RuntimeError if st == 2 else ValueError)(_STATUS_TEXT.get(st, st))
            exceptions.append(ExceptionLogger(err, Range.from_dict(qrange_of(0)) if n_alns else None))
            continue
        q_orth, s_orth, q_dropped, s_dropped, aligned = [], [], None, [], False
        a0 = int(batch.aln_off[l])
        last_qd, qrange = None, None
        for k in range(n_alns):
            a = a0 + k
            qd = qrange_of(k)
            if qd is not last_qd:                       # the alignments of a lncRNA share their qrange (never mutated here)
                last_qd, qrange = qd, Range.from_dict(qd)
            srange = Range.from_dict(srange_of(k))
            srange.name = qrange.name                                   # pipeline.py:781
            c0, c1 = int(res.chain_off[a]), int(res.chain_off[a + 1])
            if c1 > c0 and relative_of(k):
                c1 = c0               # to_record raises ValueError for a relative alignment (genomicranges.py:1259): dropped
            if c1 > c0 and columnar_src is not None:
                chain_src, a0_src = columnar_src
                chain_desc.append((res, c0, c1, int(chain_src.hsp_off[a0_src + k])))
                chain_strings += [str(qrange.chrom), str(qrange.name), str(srange.chrom)]
                chain_qplus.append(qrange.strand in ["+", "."])
                q_orth.append(None)                                     # the text is produced for all chains at once below
                s_orth.append(None)
                aligned = True
            elif c1 > c0:
                hsps = hsps_of(k, res.chain_idx[c0:c1])
                rec = chain_records(hsps, as_pvals(res.chain_pq[c0:c1]), qrange, srange.chrom, qrange.name)
                q_orth.append((rec[0], rec[1]))
                s_orth.append((rec[2], rec[3]))
                aligned = True
            else:                                                       # ValueError branch, pipeline.py:789-791
                q_dropped = qrange
                s_dropped.append(srange)
        orthologs.append((q_orth, s_orth, [] if aligned else [q_dropped, s_dropped]))

    total_exceptions = len(exceptions)
    query_exception_ranges = []
    if total_exceptions > 0:
        if not silent:
            print(f"{len(exceptions)} exceptions occured:")
        for exception in exceptions:
            if not silent:
                print(str(exception))
            query_exception_ranges.append(exception.variable)

    dist_found = [len(group[0]) for group in orthologs]
    native_text = _format_chains_native(chain_desc, chain_strings, chain_qplus, chain_src, cores) if chain_desc else None
    if fast is not None:
        dist_found = dist_found_fast
        native_text = native_text_fast if native_text_fast is not None else (b"", b"", b"", b"")
        orthologs = []                                                # the dropped lncRNAs come as finished lines
    if native_text is None:
        query_orthologs_bed12 = [o[1] for g in orthologs for o in g[0]]
        subject_orthologs_bed12 = [o[1] for g in orthologs for o in g[1]]
        query_orthologs_total = [o[0] for g in orthologs for o in g[0]]
        subject_orthologs_total = [o[0] for g in orthologs for o in g[1]]

    query_dropped, subject_dropped = [], []
    for g in orthologs:                                                 # pipeline.py:895-908
        item = g[2]
        if len(item) != 2:
            continue
        qd, sds = item
        query_dropped.append(qd)
        for grange in sds:
            subject_dropped.extend(find_neighbours(grange, qd.lifted))
    query_dropped_lines = bed6_lines(sorted(query_dropped, key=sort_key))
    subject_dropped_lines = bed6_lines(drop_duplicates(sorted(subject_dropped, key=sort_key)))
    if fast is not None:
        query_dropped_lines, subject_dropped_lines = dropped_fast
    total_dropped = len(query_dropped_lines)
    query_exception_list = sorted(query_exception_ranges, key=sort_key)

    if not os.path.exists(outdir):
        os.mkdir(outdir)
    j = lambda n: os.path.join(outdir, n)
    if native_text is not None:
        for name, text in (("significant.query_orthologs.tsv", native_text[0]), ("significant.query_orthologs.bed", native_text[1]),
                           ("significant.subject_orthologs.tsv", native_text[2]), ("significant.subject_orthologs.bed", native_text[3])):
            with open(j(name), "wb") as f:
                f.write(text)
    else:
        for name, lines in (("significant.query_orthologs.bed", query_orthologs_bed12),
                            ("significant.subject_orthologs.bed", subject_orthologs_bed12),
                            ("significant.query_orthologs.tsv", query_orthologs_total),
                            ("significant.subject_orthologs.tsv", subject_orthologs_total)):
            with open(j(name), "w") as f:
                for record in lines:
                    f.write(record + "\n")
    for name, lines in (("insignificant.query_orthologs.bed", query_dropped_lines),
                        ("insignificant.subject_orthologs.bed", subject_dropped_lines),
                        ("query_exceptions.bed", bed6_lines(query_exception_list))):
        with open(j(name), "w") as f:
            f.writelines(lines)

    stats_msg = "-----------------------\n" \
                f"build_orthologs stats:\n" \
                f"Recieved {total_alignments} transcript alignments and {total_bgs} backgrounds.\n" \
                f"Built orthologs for {total_alignments - total_dropped - total_exceptions} of transcripts.\n" \
                f"Found only unsignificant orthologs for {total_dropped} transcripts.\n" \
                f"Caught {total_exceptions} exceptions.\n" \
                f"Distribution of amount of orthologs:\n{simple_hist(dist_found)}\n" \
                f"Reported all significant orthologs for each transcript.\n" \
                "-----------------------"
    if isinstance(stats_filename, str):
        with open(stats_filename, "a") as f:
            f.write(stats_msg + "\n")
    elif not silent:
        print(stats_msg)
    PHASES["records_and_files"] = time.perf_counter() - t_phase


_EXTRACT = {
    "total_length": lambda line: int(line.strip().split("\t")[2]) - int(line.strip().split("\t")[1]),
    "block_count": lambda line: int(line.strip().split("\t")[9]),
    "block_length": lambda line: sum([int(i) for i in line.strip().split("\t")[10].split(",")]),
    "weight": lambda line: float(line.strip().split("\t")[4]),
    "name": lambda line: line.strip().split("\t")[3],
}


def get_best_orthologs(query_orthologs, query_total, subject_orthologs, subject_total, value, function,
                       outfile_query, outfile_query_total, outfile_subject, outfile_subject_total):
    """A get_best_orthologs step of the pipeline (`pipeline.py:959-1024`): streams the four
    `significant.*` files in lock-step, groups CONSECUTIVE records of the same query name and keeps
    the first extremal one by `value` computed on the query BED12 line. (The batched C-ABI call
    also returns this selection per lncRNA — `BatchResult.best_aln` — for in-memory callers; the
    file-based step costs < 1 ms per 100 records and stays text-driven like the reference's.)"""
    func = {"min": min, "max": max}[function]
    key = _EXTRACT[value]
    try:
        # native reader (libo2a_ingest.so, include/o2a_records.h); it declines anything it would not decide exactly
        # like the text-level restatement below (odd number literals, short records, carriage returns)
        from . import records_native
        if records_native.best_orthologs_files((query_orthologs, query_total, subject_orthologs, subject_total), value, function,
                                               (outfile_query, outfile_query_total, outfile_subject, outfile_subject_total)) is not None:
            return
    except (RuntimeError, OSError):
        pass                                  # library not built / unreadable file: the text path reports it
    with open(Path(query_orthologs)) as q_in, open(Path(query_total)) as qt_in, \
            open(Path(subject_orthologs)) as s_in, open(Path(subject_total)) as st_in, \
            open(Path(outfile_query), "w") as q_out, open(Path(outfile_query_total), "w") as qt_out, \
            open(Path(outfile_subject), "w") as s_out, open(Path(outfile_subject_total), "w") as st_out:
        outs = (q_out, qt_out, s_out, st_out)

        def flush(group):
            best = func(group, key=lambda rec: key(rec[0]))
            for o, line in zip(outs, best):
                o.write(line)

        purge = []
        for rec in zip(q_in, qt_in, s_in, st_in):
            if purge and _EXTRACT["name"](purge[-1][0]) != _EXTRACT["name"](rec[0]):
                flush(purge)
                purge = []
            purge.append(rec)
        if purge:
            flush(purge)
