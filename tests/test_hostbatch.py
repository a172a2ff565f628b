"""The pinned narrow host layout (hostbatch.HostBatch): what the native ingest returns and what crosses PCIe.
Host logic only — runs without a GPU (the buffers are then ordinary memory)."""
import json
import os

import numpy as np
import pytest

from test_ingest import _write_dirs

ARRAYS = ("bg_off", "aln_off", "hsp_off", "qlen", "slen", "bg_q", "score_q", "exc_off", "exc_pos", "exc_val", "coords")


def _same_columnar(a, b):
    for k in ("bg_off", "bg", "aln_off", "hsp_off", "qlen", "slen", "qstart", "qend", "sstart", "send", "score"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k


@pytest.mark.parametrize("cfg,kw", [(2, dict(n_lnc=12)), (5, dict(n_lnc=200)), (1, dict(n_lnc=8))])
def test_packer_matches_the_numpy_encoding_and_round_trips(cfg, kw):
    from ortho2align_b200 import synth
    from ortho2align_b200.hostbatch import HostBatch
    b = synth.workload(cfg, seed=3, **kw)
    for bits in (8, 16, 32):
        hb = HostBatch.from_columnar(b, threads=3, pinned=False, score_bits=bits, bg_bits=bits)
        q, eo, ep, ev = b.quantized_scores(bits)
        assert np.array_equal(hb.score_q, q) and np.array_equal(hb.exc_off, eo)
        assert np.array_equal(hb.exc_pos, ep) and np.array_equal(hb.exc_val, ev)
        assert np.array_equal(hb.bg_q, b.bg.astype(hb.bg_q.dtype)) and hb.bg_q.dtype.itemsize * 8 == bits
        assert np.array_equal(hb.coords, b.coords4())
        assert hb.max_aln_hsps == int(np.diff(b.hsp_off).max()) and hb.max_lnc_bg == int(np.diff(b.bg_off).max())
        _same_columnar(hb.to_columnar(), b)
        hb.close()


def test_planner_widens_when_too_many_scores_would_be_exceptions():
    from ortho2align_b200 import synth
    from ortho2align_b200.columnar import PackError
    from ortho2align_b200.hostbatch import HostBatch
    b = synth.workload(2, n_lnc=4, seed=1)
    assert HostBatch.from_columnar(b, pinned=False).score_bits == 8          # BLAST-like integer scores
    b.score[::3] += 1000.0                                                    # a third beyond int8
    hb = HostBatch.from_columnar(b, pinned=False)
    assert hb.score_bits == 16 and hb.n_exc == int((b.score != np.floor(b.score)).sum())
    b.score[::2] += 0.5                                                       # half fractional: 32 is the last resort
    hb = HostBatch.from_columnar(b, pinned=False)
    assert hb.score_bits == 32 and np.array_equal(hb.scores(), b.score)
    b.bg[5] = 40000
    assert HostBatch.from_columnar(b, pinned=False).bg_bits == 32
    with pytest.raises(PackError):
        HostBatch.from_columnar(b, pinned=False, bg_bits=16)


def test_extreme_scores_survive_every_width():
    from ortho2align_b200 import synth
    from ortho2align_b200.hostbatch import HostBatch
    b = synth.workload(5, n_lnc=40, seed=9)
    vals = [-128.0, -129.0, 127.0, 128.0, -32768.0, 32767.0, 32768.0, -2147483648.0, 2147483647.0, 2147483648.0,
            1e300, -1e300, 0.5, -0.0, 5e-324]
    b.score[:len(vals)] = vals
    for bits in (8, 16, 32):
        hb = HostBatch.from_columnar(b, pinned=False, score_bits=bits)
        got = hb.scores()
        assert np.array_equal(got, b.score) and np.array_equal(np.signbit(got), np.signbit(b.score))    # -0.0 included
        q, eo, ep, ev = b.quantized_scores(bits)
        assert np.array_equal(hb.score_q, q) and np.array_equal(hb.exc_pos, ep)


def test_ingest_returns_the_layout_the_packer_builds(tmp_path):
    from ortho2align_b200 import ingest, synth
    from ortho2align_b200.hostbatch import HostBatch
    b, m = synth.workload(2, n_lnc=10, with_meta=True, hsps_per_region=120, n_bg=500, seed=11)
    items = synth.to_json_dicts(b, m)
    af, bf = _write_dirs(tmp_path, items)
    r = ingest.load_files(af, bf, threads=4, pinned=False)
    want = HostBatch.from_columnar(b, pinned=False)
    assert (r.host.score_bits, r.host.bg_bits, r.host.n_exc) == (want.score_bits, want.bg_bits, want.n_exc)
    for k in ARRAYS:
        assert np.array_equal(getattr(r.host, k), getattr(want, k)), k
    _same_columnar(r.batch, b)
    assert np.array_equal(r.score, b.score) and not r.relative.any()
    # single-threaded and multi-threaded exports agree
    r1 = ingest.load_files(af, bf, threads=1, pinned=False)
    for k in ARRAYS:
        assert np.array_equal(getattr(r.host, k), getattr(r1.host, k)), k


def test_non_default_json_separators_and_key_orders_take_the_general_path(tmp_path):
    """The fast HSP path only matches json.dump's default text; anything else must parse to the same arrays."""
    from ortho2align_b200 import ingest, synth
    b, m = synth.workload(1, n_lnc=6, with_meta=True, hsps_per_region=60, n_bg=100, seed=2)
    items = synth.to_json_dicts(b, m)
    ad, bd = tmp_path / "align", tmp_path / "bg"
    ad.mkdir(); bd.mkdir()
    af, bf = [], []
    styles = [dict(), dict(separators=(",", ":")), dict(indent=1), dict(sort_keys=True), dict(separators=(" , ", " : "))]
    for k, (name, alns, bg) in enumerate(items):
        if k == 5:                                   # filtered_HSPs different from HSPs, floats, extra keys
            for a in alns:
                a["filtered_HSPs"] = a["filtered_HSPs"][:3]
                a["HSPs"][0]["score"] = 17.25
                a["HSPs"][1]["kwargs"] = {"note": 'a } tricky " ] string', "x": [1, 2, {"y": None}]}
        (ad / (name + ".json")).write_text(json.dumps(alns, **styles[k % len(styles)]))
        (bd / (name + ".json")).write_text(json.dumps(bg, **styles[(k + 1) % len(styles)]))
        af.append(str(ad / (name + ".json"))); bf.append(str(bd / (name + ".json")))
    r = ingest.load_files(af, bf, threads=2, pinned=False)
    assert (r.file_status == 0).all(), r.errors
    from ortho2align_b200.columnar import concat_batches, pack_lncrna
    want = concat_batches([pack_lncrna(json.loads(open(a).read()), json.loads(open(c).read())) for a, c in zip(af, bf)])
    _same_columnar(r.batch, want)


@pytest.mark.parametrize("value", [True, 1, "yes", [0]])
def test_relative_alignments_are_reported_as_dropped(tmp_path, monkeypatch, value):
    """to_record raises ValueError for an alignment whose 'relative' flag is truthy (genomicranges.py:1259); _build
    catches it and the alignment counts as dropped (pipeline.py:789-791)."""
    from test_pipeline_host import _oracle_runner
    from ortho2align_b200 import pipeline, synth
    b, m = synth.workload(2, n_lnc=4, with_meta=True, hsps_per_region=400, n_bg=2000, seed=4)
    items = synth.to_json_dicts(b, m)
    monkeypatch.setattr(pipeline, "_run_shards", _oracle_runner)
    outs = {}
    for tag, flag in (("plain", False), ("relative", value)):
        d = tmp_path / tag
        d.mkdir()
        for a in items[1][1]:
            a["relative"] = flag                      # every alignment of the second lncRNA
        items[2][1][0]["relative"] = flag            # one alignment of the third
        af, bf = _write_dirs(d, items)
        for ing in ("native", "json"):
            od = str(d / ("out_" + ing))
            pipeline.build_orthologs(os.path.dirname(af[0]), os.path.dirname(bf[0]), "hist", od, fdr=True, silent=True,
                                     stats_filename=str(d / "stats.txt"), ingest=ing)
            outs[(tag, ing)] = {fn: open(os.path.join(od, fn)).read() for fn in sorted(os.listdir(od))}
        assert outs[(tag, "native")] == outs[(tag, "json")]
    plain, rel = outs[("plain", "native")], outs[("relative", "native")]
    name1, name2 = items[1][0], items[2][0]
    assert name1 in plain["significant.query_orthologs.bed"]
    assert name1 not in rel["significant.query_orthologs.bed"] and name1 in rel["insignificant.query_orthologs.bed"]
    n_plain = sum(1 for ln in plain["significant.query_orthologs.bed"].splitlines() if ln.split("\t")[3] == name2)
    n_rel = sum(1 for ln in rel["significant.query_orthologs.bed"].splitlines() if ln.split("\t")[3] == name2)
    assert n_rel <= n_plain and name2 not in rel["insignificant.query_orthologs.bed"] or n_rel == 0
