"""Host logic of the drop-in pipeline (ingest, record text, dropped/exception bookkeeping, stats,
get_best_orthologs) against the files the unmodified reference wrote (tests/golden/e2e.json).
CPU tests inject the oracle in place of the GPU call; the gpu test runs the real thing."""
import json
import os

import numpy as np
import pytest

from conftest import load_golden


def _materialise(tmp_path, g):
    ad, bd = tmp_path / "align", tmp_path / "bg"
    ad.mkdir(); bd.mkdir()
    for item in g["inputs"]:
        (ad / (item["name"] + ".json")).write_text(json.dumps(item["alignments"]))
        if item["name"] != g["missing_bg"]:
            (bd / (item["name"] + ".json")).write_text(json.dumps(item["bg"]))
    return str(ad), str(bd)


def _oracle_runner(batches, devices, fitter, fdr, alpha, gapopen, gapextend):
    from oracle import oracle as O
    # the native ingest hands over its pinned narrow layout (HostBatch); the oracle takes the plain columnar form
    plain = [b.to_columnar() if hasattr(b, "to_columnar") else b for b in batches]
    return [O.build_batch(b, fitter, fdr, alpha, gapopen, gapextend) for b in plain]


def _run_and_compare(tmp_path, monkeypatch, g, key, fitting, fdr, exact, ingest="native", cache=None, dirs=None,
                     devices=None, cores=1):
    from ortho2align_b200 import pipeline
    ad, bd = dirs or _materialise(tmp_path, g)
    real_listdir = os.listdir
    monkeypatch.setattr(os, "listdir", lambda p: list(g["listdir_order"]) if str(p) == ad else real_listdir(p))
    od = str(tmp_path / f"out_{key}")
    stats = str(tmp_path / f"{key}.stats.txt")
    pipeline.build_orthologs(ad, bd, fitting, od, threshold=0.05, gapopen=5, gapextend=2, fdr=fdr, cores=cores,
                             timeout=None, silent=True, stats_filename=stats, ingest=ingest, cache=cache, devices=devices)
    j = lambda n: os.path.join(od, n)
    pipeline.get_best_orthologs(j("significant.query_orthologs.bed"), j("significant.query_orthologs.tsv"),
                                j("significant.subject_orthologs.bed"), j("significant.subject_orthologs.tsv"),
                                "block_length", "max",
                                j("bestSignificant.query_orthologs.bed"), j("bestSignificant.query_orthologs.tsv"),
                                j("bestSignificant.subject_orthologs.bed"), j("bestSignificant.subject_orthologs.tsv"))
    want = g["runs"][key]
    for fn, text in want.items():
        got = open(stats).read() if fn == "stats.txt" else open(j(fn)).read()
        if exact or not fn.endswith(".tsv"):
            assert got == text, f"{key}: {fn} differs"
        else:
            # KDE: the 13th TSV column (p/q values) is compared to tolerance, everything else verbatim
            gl, wl = got.splitlines(), text.splitlines()
            assert len(gl) == len(wl)
            for a, b in zip(gl, wl):
                ca, cb = a.split("\t"), b.split("\t")
                assert ca[:12] == cb[:12]
                pa = np.array([float(x) for x in ca[12].split(",")]); pb = np.array([float(x) for x in cb[12].split(",")])
                assert np.all(np.abs(pa - pb) <= np.maximum(1e-12 * np.abs(pb), 16 * 2.0 ** -53 * 200))


@pytest.mark.parametrize("key,fitting,fdr,exact", [("hist_1", "hist", True, True), ("hist_0", "hist", False, True),
                                                   ("kde_1", "kde", True, False)])
@pytest.mark.parametrize("ingest", ["native", "json"])
def test_host_logic_with_oracle_injected(tmp_path, monkeypatch, key, fitting, fdr, exact, ingest):
    from ortho2align_b200 import pipeline
    monkeypatch.setattr(pipeline, "_run_shards", _oracle_runner)
    _run_and_compare(tmp_path, monkeypatch, load_golden("e2e.json"), key, fitting, fdr, exact, ingest=ingest)


@pytest.mark.parametrize("ingest", ["native", "json"])
def test_lncrna_shards_over_several_devices_give_the_same_files(tmp_path, monkeypatch, ingest):
    """devices=[...] shards the lncRNAs (LPT by HSP count) into one batched call per device; the files do not change."""
    from ortho2align_b200 import pipeline
    seen = []

    def runner(batches, devices, *a):
        seen.append((len(batches), list(devices)))
        return _oracle_runner(batches, devices, *a)
    monkeypatch.setattr(pipeline, "_run_shards", runner)
    _run_and_compare(tmp_path, monkeypatch, load_golden("e2e.json"), "hist_1", "hist", True, True, ingest=ingest,
                     devices=[0, 1, 2])
    assert seen and seen[0][0] == 3 and seen[0][1] == [0, 1, 2]


def test_columnar_sidecar_replaces_the_json_parse(tmp_path, monkeypatch):
    """First run parses the JSON and writes the sidecar; the second reads only the sidecar; a touched input
    makes it stale. Output files are the reference's in all three runs."""
    from ortho2align_b200 import ingest, pipeline
    monkeypatch.setattr(pipeline, "_run_shards", _oracle_runner)
    g = load_golden("e2e.json")
    dirs = _materialise(tmp_path, g)
    cache = str(tmp_path / "columnar.o2ac")
    calls = []
    real = ingest.load_files
    monkeypatch.setattr(ingest, "load_files", lambda *a, **k: calls.append(1) or real(*a, **k))
    _run_and_compare(tmp_path, monkeypatch, g, "hist_1", "hist", True, True, cache=cache, dirs=dirs)
    assert os.path.exists(cache) and len(calls) == 1
    for sub in ("out_hist_1", "hist_1.stats.txt"):
        os.rename(tmp_path / sub, tmp_path / (sub + ".first"))
    _run_and_compare(tmp_path, monkeypatch, g, "hist_1", "hist", True, True, cache=cache, dirs=dirs)
    assert len(calls) == 1                                        # no JSON was parsed
    for sub in ("out_hist_1", "hist_1.stats.txt"):
        os.rename(tmp_path / sub, tmp_path / (sub + ".second"))
    first = sorted(os.listdir(dirs[0]))[0]
    st = os.stat(os.path.join(dirs[0], first))
    os.utime(os.path.join(dirs[0], first), ns=(st.st_atime_ns, st.st_mtime_ns + 10 ** 9))
    _run_and_compare(tmp_path, monkeypatch, g, "hist_1", "hist", True, True, cache=cache, dirs=dirs)
    assert len(calls) == 2                                        # stale -> parsed again and rewritten


@pytest.mark.gpu
@pytest.mark.parametrize("key,fitting,fdr,exact", [("hist_1", "hist", True, True), ("hist_0", "hist", False, True),
                                                   ("kde_1", "kde", True, False)])
def test_drop_in_pipeline_on_gpu(tmp_path, monkeypatch, key, fitting, fdr, exact):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    _run_and_compare(tmp_path, monkeypatch, load_golden("e2e.json"), key, fitting, fdr, exact)


@pytest.mark.gpu
def test_drop_in_pipeline_with_two_engines_on_one_gpu(tmp_path, monkeypatch):
    """devices=[0, 0]: two shards, two o2a_ctx on their own threads (what devices=[0, 1] does on a 2-GPU box)."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    _run_and_compare(tmp_path, monkeypatch, load_golden("e2e.json"), "hist_1", "hist", True, True, devices=[0, 0])


def test_cli_parses_reference_flags():
    from ortho2align_b200.cli import make_parser
    from ortho2align_b200.pipeline import build_orthologs, get_best_orthologs
    a = make_parser().parse_args("build_orthologs -alignments A -background B -fitting hist -threshold 0.01 "
                                 "--fdr -outdir O -cores 4 --silent".split())
    assert a.func is build_orthologs and a.fdr and a.fitting == "hist" and a.threshold == 0.01 and a.timeout is None
    b = make_parser().parse_args("get_best_orthologs -query_orthologs q -query_total qt -subject_orthologs s "
                                 "-subject_total st -value block_length -function max -outfile_query a "
                                 "-outfile_query_total b -outfile_subject c -outfile_subject_total d".split())
    assert b.func is get_best_orthologs and b.value == "block_length"


def test_lpt_shards_cover_everything_once():
    from ortho2align_b200.pipeline import lpt_shards
    rng = np.random.default_rng(0)
    w = rng.integers(1, 1000, size=57)
    for n in (1, 2, 4, 8):
        sh = lpt_shards(w, n)
        assert sorted(i for s in sh for i in s) == list(range(57))
        loads = [int(w[s].sum()) for s in sh]
        assert max(loads) - min(loads) <= int(w.max())


def test_background_patches_follow_the_reference_order():
    from ortho2align_b200.columnar import patch_background
    assert patch_background(None) == [0, 1]
    assert patch_background([]) == [1, 0]
    assert patch_background([1]) == [1, 1, 0]
    assert patch_background([5, 5]) == [5, 5, 1]
    assert patch_background([1, 7]) == [1, 7, 0]
    assert patch_background([3, 4]) == [3, 4]


class _OracleEngineForStream:
    """Stands in for Engine in the CPU test of BatchStream: the oracle computes, with a per-batch delay that makes
    completion order differ from input order."""
    created, closed = 0, 0

    def __init__(self, device):
        type(self).created += 1
        self.device = device

    def build_batch(self, batch, fitter, fdr, alpha, gapopen, gapextend, **kw):
        import time
        from oracle import oracle as O
        time.sleep(0.02 * (batch.n_lnc % 3))
        if getattr(batch, "_poison", False):
            raise RuntimeError("poisoned batch")
        return O.build_batch(batch, fitter, fdr, alpha, gapopen, gapextend)

    def close(self):
        type(self).closed += 1


def test_batch_stream_keeps_order_and_surfaces_errors():
    from oracle import oracle as O
    from ortho2align_b200 import synth
    from ortho2align_b200.stream import BatchStream
    batches = [synth.workload(1, n_lnc=n, seed=s, hsps_per_region=120, n_bg=300) for s, n in enumerate([3, 1, 2, 4, 2, 1, 3])]
    E = _OracleEngineForStream
    E.created = E.closed = 0
    got = list(BatchStream(devices=(0,), contexts=2, engine_factory=E).map(iter(batches), "hist", True))
    assert len(got) == len(batches) and E.created == 2 and E.closed == 2
    for b, r in zip(batches, got):
        exp = O.build_batch(b, "hist", True)
        assert np.array_equal(r.chain_idx, exp.chain_idx) and np.array_equal(r.thr, exp.thr)
        assert np.array_equal(r.best_aln, exp.best_aln)
    assert list(BatchStream(devices=(0, 1), contexts=1, engine_factory=E).map([], "hist", True)) == []
    batches[3]._poison = True
    with pytest.raises(RuntimeError, match="poisoned"):
        list(BatchStream(devices=(0,), contexts=2, engine_factory=E).map(batches, "hist", True))
    assert E.created == E.closed


def test_columnar_assembly_is_the_path_taken_and_matches_the_per_alignment_one(tmp_path, monkeypatch):
    """With the native ingest + native records the run is assembled from columnar data (pipeline._assemble_columnar: no
    Python object per alignment); records='python' forces the per-alignment path. Same bytes, and the reference's."""
    from ortho2align_b200 import pipeline
    calls = []
    real = pipeline._assemble_columnar

    def spy(*a, **k):
        out = real(*a, **k)
        calls.append(out is not None)
        return out
    monkeypatch.setattr(pipeline, "_assemble_columnar", spy)
    monkeypatch.setattr(pipeline, "_run_shards", _oracle_runner)
    g = load_golden("e2e.json")
    dirs = _materialise(tmp_path, g)
    _run_and_compare(tmp_path, monkeypatch, g, "hist_1", "hist", True, True, ingest="native", dirs=dirs)
    assert calls == [True]
    od_fast = {fn: open(os.path.join(tmp_path, "out_hist_1", fn)).read() for fn in os.listdir(tmp_path / "out_hist_1")}
    real_listdir = os.listdir
    monkeypatch.setattr(os, "listdir", lambda p: list(g["listdir_order"]) if str(p) == dirs[0] else real_listdir(p))
    pipeline.build_orthologs(dirs[0], dirs[1], "hist", str(tmp_path / "slow"), fdr=True, silent=True,
                             stats_filename=str(tmp_path / "slow.stats"), records="python")
    assert calls == [True]                                   # not consulted for records='python'
    for fn in ("significant.query_orthologs.bed", "significant.query_orthologs.tsv", "significant.subject_orthologs.bed",
               "significant.subject_orthologs.tsv", "insignificant.query_orthologs.bed", "insignificant.subject_orthologs.bed",
               "query_exceptions.bed"):
        assert open(tmp_path / "slow" / fn).read() == od_fast[fn], fn


def test_range_objects_with_null_names_escapes_and_odd_types(tmp_path, monkeypatch):
    """Names the native range reader renders itself (null name -> 'chrom:start-endstrand', \\u escapes, quotes) and
    values it leaves to Python's json (a numeric chrom): both give the files of the json.load path."""
    from ortho2align_b200 import pipeline, synth
    monkeypatch.setattr(pipeline, "_run_shards", _oracle_runner)
    b, m = synth.workload(1, n_lnc=5, with_meta=True, hsps_per_region=150, n_bg=400, seed=19)
    items = synth.to_json_dicts(b, m)
    for a in items[0][1]:
        a["qrange"]["name"] = None
    for a in items[1][1]:
        a["qrange"]["name"] = 'lnc "β" \\ tab\tend'
        a["srange"]["chrom"] = "chré/1"
    for a in items[2][1]:
        a["qrange"]["chrom"] = 7                               # str(7): not a JSON string -> the Python path for this run
    outs = {}
    for tag, its in (("simple", [it for k, it in enumerate(items) if k != 2]), ("odd", items)):
        d = tmp_path / tag
        d.mkdir()
        ad, bd = d / "align", d / "bg"
        ad.mkdir(); bd.mkdir()
        for name, alns, bg in its:
            (ad / (name + ".json")).write_text(json.dumps(alns))
            (bd / (name + ".json")).write_text(json.dumps(bg))
        for ing, rec in (("native", "native"), ("json", "python")):
            od = str(d / f"out_{ing}")
            pipeline.build_orthologs(str(ad), str(bd), "hist", od, fdr=True, silent=True, stats_filename=str(d / "s.txt"),
                                     ingest=ing, records=rec)
            outs[(tag, ing)] = {fn: open(os.path.join(od, fn), encoding="utf-8").read() for fn in sorted(os.listdir(od))}
        assert outs[(tag, "native")] == outs[(tag, "json")], tag
    q = outs[("simple", "native")]["significant.query_orthologs.bed"]
    assert 'lnc "β" \\ tab' in q and any(":" in ln.split("\t")[3] and ln.split("\t")[3][-1] in "+-." for ln in q.splitlines())


@pytest.mark.parametrize("value", ["total_length", "block_count", "block_length", "weight"])
@pytest.mark.parametrize("function", ["max", "min"])
def test_native_get_best_orthologs_equals_the_text_statement(tmp_path, monkeypatch, value, function):
    from ortho2align_b200 import pipeline, records_native
    g = load_golden("r2_best_weight.json")
    ins = []
    for k, fn in enumerate(("significant.query_orthologs.bed", "significant.query_orthologs.tsv",
                            "significant.subject_orthologs.bed", "significant.subject_orthologs.tsv")):
        p = tmp_path / fn
        p.write_text(g["build"][fn])
        ins.append(str(p))
    outs_n = [str(tmp_path / f"n{k}") for k in range(4)]
    outs_p = [str(tmp_path / f"p{k}") for k in range(4)]
    assert records_native.best_orthologs_files(ins, value, function, outs_n) is not None
    monkeypatch.setattr(records_native, "best_orthologs_files", lambda *a, **k: None)       # decline -> text path
    pipeline.get_best_orthologs(*ins, value, function, *outs_p)
    want = g["runs"][f"{value}.{function}"]
    assert [open(p).read() for p in outs_n] == want and [open(p).read() for p in outs_p] == want


def test_native_get_best_orthologs_declines_what_it_would_not_decide_like_python(tmp_path):
    from ortho2align_b200 import pipeline, records_native
    line = "chr1\t10\t50\tlncA\t12\t+\t10\t50\t0\t2\t10,20\t0,30\n"
    for bad in (line.replace("\n", "\r\n"), line.replace("10,20", "10, 20"), line.replace("\t12\t", "\tinf\t"), "chr1\t10\t50\n"):
        paths = []
        for k in range(4):
            p = tmp_path / f"in{k}"
            p.write_text(line + bad, newline="")
            paths.append(str(p))
        outs = [str(tmp_path / f"o{k}") for k in range(4)]
        value = "weight" if "inf" in bad else "block_length"
        assert records_native.best_orthologs_files(paths, value, "max", outs) is None
        if bad.count("\t") < 11:
            with pytest.raises(IndexError):
                pipeline.get_best_orthologs(*paths, value, "max", *outs)
        else:
            pipeline.get_best_orthologs(*paths, value, "max", *outs)
            assert open(outs[0]).read().count("\n") == 1


def test_plain_range_fast_path_equals_json_decode():
    """pipeline._plain_range (one anchored regular expression for the flat subject ranges of dropped lncRNAs) returns
    what Range.from_dict(json.loads(text)) returns, or None for anything it does not cover (escapes, relations, other
    key orders, non-string chrom) — then the caller falls back to the JSON decode."""
    import random
    from ortho2align_b200.pipeline import _plain_range
    from ortho2align_b200.records import Range
    rng = random.Random(5)
    names = [None, "lnc1", "gene (copy)", "a\\b", 'q"uote', "naïve", ""]
    covered = 0
    for _ in range(400):
        d = {"chrom": rng.choice(["chr1", "chrX", "scaffold_12", 'c"hr', 7]), "start": rng.randint(-10, 10 ** 9),
             "end": rng.randint(0, 10 ** 9), "strand": rng.choice(["+", "-", "."]), "genome": rng.choice([None, "hg38", "mm10"]),
             "sequence_file_path": {"path": rng.choice([None, "/data/genome.fa"])}, "name": rng.choice(names),
             "relations": rng.choice([{}, {}, {}, {"lifted": {"collection": [], "sequence_file_path": {"path": None},
                                                            "grange_class": "GenomicRange"}}]),
             "relations_class": None}
        if d["relations"]:
            d["relations_class"] = "GenomicRangesList"
        if rng.random() < 0.1:                                    # another key order
            d = dict(reversed(list(d.items())))
        text = json.dumps(d).encode()
        want = Range.from_dict(json.loads(text))
        got = _plain_range(text)
        if got is None:
            continue
        covered += 1
        assert (got.chrom, got.start, got.end, got.strand, got._name, got.genome, got.lifted) == \
               (want.chrom, want.start, want.end, want.strand, want._name, want.genome, want.lifted)
        assert type(got.start) is int and type(got.chrom) is str
    assert covered > 100


def test_dropped_lncrnas_rendered_by_forked_workers_give_the_same_files(tmp_path, monkeypatch):
    """pipeline._render_dropped spreads the Python-side bookkeeping of lncRNAs without orthologs over forked workers
    (from a few thousand on); forced here on a genome-scale-shaped sample: every output file must be byte-identical to
    the single-process run."""
    import filecmp
    from ortho2align_b200 import pipeline, synth
    monkeypatch.setattr(pipeline, "_run_shards", _oracle_runner)
    b, m = synth.workload(5, n_lnc=400, with_meta=True, seed=9)
    ad, bd = tmp_path / "align", tmp_path / "bg"
    ad.mkdir(), bd.mkdir()
    for name, alns, bg in synth.to_json_dicts(b, m):
        (ad / f"{name}.json").write_text(json.dumps(alns))
        (bd / f"{name}.json").write_text(json.dumps(bg))
    pools = []
    import multiprocessing as mp
    real_ctx = mp.get_context

    def spy(method=None):
        pools.append(method)
        return real_ctx(method)
    monkeypatch.setattr(mp, "get_context", spy)
    outs = {}
    for tag, cores, per_worker in (("serial", 1, 2000), ("forked", 3, 1)):
        monkeypatch.setattr(pipeline, "_DROPPED_PER_WORKER", per_worker)
        outs[tag] = str(tmp_path / tag)
        pipeline.build_orthologs(str(ad), str(bd), "hist", outs[tag], fdr=True, cores=cores, silent=True)
    assert pools == ["fork"]                                 # the second run went through the pool
    files = sorted(os.listdir(outs["serial"]))
    assert "insignificant.subject_orthologs.bed" in files
    assert os.path.getsize(os.path.join(outs["serial"], "insignificant.query_orthologs.bed")) > 0
    for fn in files:
        assert filecmp.cmp(os.path.join(outs["serial"], fn), os.path.join(outs["forked"], fn), shallow=False), fn
