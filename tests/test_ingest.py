"""Native JSON ingest (libo2a_ingest.so) against json.load + the Python packer on the same files."""
import json
import os

import numpy as np
import pytest

from conftest import load_golden


def _write_dirs(tmp_path, items, missing=()):
    ad, bd = tmp_path / "align", tmp_path / "bg"
    ad.mkdir(); bd.mkdir()
    af, bf = [], []
    for name, alns, bg in items:
        (ad / (name + ".json")).write_text(json.dumps(alns))
        if name not in missing:
            (bd / (name + ".json")).write_text(json.dumps(bg))
        af.append(str(ad / (name + ".json"))); bf.append(str(bd / (name + ".json")))
    return af, bf


def _python_batch(af, bf):
    from ortho2align_b200.columnar import concat_batches
    from ortho2align_b200.pipeline import _load_lncrna
    return concat_batches([_load_lncrna(a, b)[1] for a, b in zip(af, bf)])


def _same(a, b):
    for k in ("bg_off", "bg", "aln_off", "hsp_off", "qlen", "slen", "qstart", "qend", "sstart", "send", "score"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k


def test_native_ingest_matches_python_packer(tmp_path):
    from ortho2align_b200 import ingest, synth
    b, m = synth.workload(1, n_lnc=12, with_meta=True, hsps_per_region=150, n_bg=300, seed=5)
    items = synth.to_json_dicts(b, m)
    items[2][1][0]["qlen"] = None                    # null qlen/slen -> max+1 (alignment.py:477-487)
    items[2][1][0]["slen"] = None
    items[4] = (items[4][0], items[4][1], [5, 5])    # degenerate backgrounds (pipeline.py:751-754)
    items[5] = (items[5][0], items[5][1], [1])
    items[6] = (items[6][0], items[6][1], [])
    af, bf = _write_dirs(tmp_path, items, missing={items[7][0]})
    r = ingest.load_files(af, bf, threads=3)
    assert (r.file_status == 0).all()
    _same(r.batch, _python_batch(af, bf))
    assert np.array_equal(r.batch.coords4(), np.stack([r.batch.qstart, r.batch.qend, r.batch.sstart, r.batch.send], 1))
    # JSON number types survive (ints stay ints in the record text)
    h = 0
    for name, alns, bg in items:
        for a in alns:
            for hs in a["HSPs"]:
                assert bool(r.score_is_int[h]) == isinstance(hs["score"], int)
                assert r.hsp_dict(h)["score"] == hs["score"] and type(r.hsp_dict(h)["score"]) is type(hs["score"])
                h += 1
    a = 0
    for name, alns, bg in items:
        for al in alns:
            assert r.range_dict(a, "q") == al["qrange"] and r.range_dict(a, "s") == al["srange"]
            a += 1


def test_native_ingest_whitespace_escapes_and_errors(tmp_path):
    from ortho2align_b200 import ingest
    good = [{"relative": False, "qrange": {"name": 'a"b{[\\', "chrom": "c"}, "filtered": False,
             "HSPs": [{"kwargs": {"x": [1, {"y": "}"}]}, "score": 1.5e1, "send": 9, "sstart": 5, "qend": 20, "pval": None,
                       "qstart": 10}],
             "filtered_HSPs": None, "slen": 30, "srange": {"chrom": "d"}, "qlen": 40}]
    (tmp_path / "g.json").write_text(json.dumps(good, indent=2))
    (tmp_path / "bad.json").write_text('[{"HSPs": [{"qstart": 1, "qend": 2, "sstart": 3, "send": 4, "score": }]')
    (tmp_path / "big.json").write_text(json.dumps([{"HSPs": [{"qstart": 1, "qend": 2 ** 40, "sstart": 3, "send": 4, "score": 5}],
                                                    "qlen": 1, "slen": 1, "qrange": {}, "srange": {}}]))
    (tmp_path / "bg.json").write_text("[ 12 ,13,\n 14 ]")
    files = [str(tmp_path / n) for n in ("g.json", "bad.json", "big.json", "nope.json")]
    r = ingest.load_files(files, [str(tmp_path / "bg.json")] * 4, threads=2)
    assert r.file_status.tolist() == [0, 1, 2, 1]
    assert r.batch.n_lnc == 1 and r.batch.n_hsp == 1
    assert r.batch.score[0] == 15.0 and r.score_is_int[0] == 0
    assert r.batch.bg.tolist() == [12, 13, 14]
    assert r.range_dict(0, "q") == good[0]["qrange"]


def test_native_ingest_on_the_reference_golden_inputs(tmp_path):
    from ortho2align_b200 import ingest
    g = load_golden("e2e.json")
    items = [(i["name"], i["alignments"], i["bg"]) for i in g["inputs"]]
    af, bf = _write_dirs(tmp_path, items, missing={g["missing_bg"]})
    r = ingest.load_files(af, bf)
    _same(r.batch, _python_batch(af, bf))


def test_sidecar_round_trip(tmp_path):
    """The columnar sidecar holds exactly what the native parse produced, range texts included."""
    from ortho2align_b200 import ingest, synth
    b, m = synth.workload(1, n_lnc=9, with_meta=True, hsps_per_region=80, n_bg=100, seed=8)
    items = synth.to_json_dicts(b, m)
    items[3] = (items[3][0], items[3][1], [])
    af, bf = _write_dirs(tmp_path, items, missing={items[1][0]})
    r = ingest.load_files(af, bf, threads=2)
    path = str(tmp_path / "c.o2ac")
    ingest.save_sidecar(path, r, af, bf)
    s = ingest.open_sidecar(path, af, bf)
    assert s is not None
    _same(s.batch, r.batch)
    assert np.array_equal(s.lnc_file, r.lnc_file) and np.array_equal(s.score_is_int, r.score_is_int)
    for a in range(r.batch.n_aln):
        assert s.range_dict(a, "q") == r.range_dict(a, "q") and s.range_dict(a, "s") == r.range_dict(a, "s")
    for h in range(0, r.batch.n_hsp, 37):
        assert s.hsp_dict(h) == r.hsp_dict(h) and type(s.hsp_dict(h)["score"]) is type(r.hsp_dict(h)["score"])
    # stale: different order, a rewritten file, a background that appeared, garbage in the file
    assert ingest.open_sidecar(path, af[::-1], bf[::-1]) is None
    assert ingest.open_sidecar(str(tmp_path / "missing.o2ac"), af, bf) is None
    with open(bf[1], "w") as f:
        f.write("[3, 4, 5]")
    assert ingest.open_sidecar(path, af, bf) is None
    with open(path, "r+b") as f:
        f.write(b"garbage!")
    assert ingest.open_sidecar(path, af, bf) is None
