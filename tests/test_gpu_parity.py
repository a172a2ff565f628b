"""Parity tests proper: the CUDA path, called through the C-ABI, against (a) the golden vectors
recorded from the unmodified reference and (b) the oracle on the same seeded inputs."""
import numpy as np
import pytest

from conftest import batch_from_lists, load_golden
from helpers import KDE_THR_ABS, assert_pq_close, compare_results, compare_with_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,exact", [("build_hist_fdr.json", True), ("build_hist_nofdr.json", True),
                                        ("build_c5_hist_fdr.json", True), ("build_kde_fdr.json", False)])
def test_golden_build(engine, name, exact):
    g = load_golden(name)
    b = batch_from_lists(g["batch"])
    res = engine.build_batch(b, g["fitting"], g["fdr"], g["alpha"], g["gapopen"], g["gapextend"])
    assert (res.status == 0).all()
    compare_with_golden(res, b, g, exact_p=exact, what=name)


def test_golden_fitters(engine):
    from ortho2align_b200.columnar import kde_factor
    g = load_golden("fitters.json")
    for case in g["cases"]:
        bg = np.asarray(case["bg"], dtype=np.int32)
        sf, isf = engine.fitter_eval("hist", bg, case["x"], case["alphas"])
        assert sf.tolist() == case["hist_sf"]
        assert isf.tolist() == case["hist_isf"]
        if "kde_sf" in case:
            sf, isf = engine.fitter_eval("kde", bg, case["x"], case["alphas"][:2], kde_factor=kde_factor(len(bg)))
            assert_pq_close(sf, case["kde_sf"], "kde sf")
            assert np.allclose(isf, case["kde_isf"], rtol=0, atol=KDE_THR_ABS)


def test_golden_chains(engine):
    g = load_golden("chains.json")
    for i, c in enumerate(g["cases"]):
        chain, orient, score, s2 = engine.best_chain(c["qstart"], c["qend"], c["sstart"], c["send"], c["score"],
                                                     c["qlen"], c["slen"], c["gapopen"], c["gapextend"])
        assert chain.tolist() == c["chain"], f"case {i}"
        assert s2 == (c["score_direct"], c["score_reverse"]), f"case {i}"
        if c["chain"]:
            assert orient == c["orient"]


def test_orientation_none_raises(engine):
    with pytest.raises(KeyError):
        engine.best_chain([10, 30], [20, 30], [5, 50], [9, 60], [30.0, 40.0], 100, 100)


@pytest.mark.parametrize("cfg,kw", [
    (1, dict()),                                                   # BASELINE config 1 (96 strRNA genes)
    (2, dict(n_lnc=48)),                                           # config 2 shape, fewer lncRNAs
    (5, dict(n_lnc=1500)),                                         # ragged: many tiny alignments
    (2, dict(n_lnc=6, hsps_per_region=7, n_bg=2)),                 # minimum background size
    (2, dict(n_lnc=3, n_bg=40000, hsps_per_region=300)),           # B > shared-memory bin budget
])
@pytest.mark.parametrize("fitter,fdr", [("hist", True), ("hist", False), ("empirical", True)])
def test_synthetic_vs_oracle(engine, cfg, kw, fitter, fdr):
    from oracle import oracle as O
    from ortho2align_b200 import synth
    b = synth.workload(cfg, seed=5, **kw)
    exp = O.build_batch(b, fitter, fdr)
    got = engine.build_batch(b, fitter, fdr)
    compare_results(got, exp, exact_p=True, what=f"cfg{cfg} {fitter} fdr={fdr}")
    got = engine.build_batch(b, fitter, fdr, aos=True)          # AoS coordinates (o2a_batch.coords)
    compare_results(got, exp, exact_p=True, what=f"cfg{cfg} {fitter} fdr={fdr} aos")
    for bits in (8, 16, 32):                                    # narrow scores + exception lists
        got = engine.build_batch(b, fitter, fdr, aos=True, score_bits=bits, bg_bits=bits)
        compare_results(got, exp, exact_p=True, what=f"cfg{cfg} {fitter} fdr={fdr} score_bits={bits}")
    assert engine.counts.n_survivors == int(exp.n_surv.sum())
    assert engine.counts.n_chains == int((exp.chain_len > 0).sum())


def test_narrow_scores_with_many_exceptions(engine):
    """Scores outside int16, fractional scores everywhere, NaN-free extremes: the narrow encodings stay lossless."""
    from oracle import oracle as O
    from ortho2align_b200 import synth
    b = synth.workload(5, n_lnc=300, seed=17)
    rng = np.random.default_rng(1)
    b.score[rng.random(b.n_hsp) < 0.3] *= 0.77          # 30 % fractional
    b.score[rng.random(b.n_hsp) < 0.05] += 40000.0      # beyond int16
    b.score[rng.random(b.n_hsp) < 0.01] = -32768.0      # collides with the int16 sentinel value
    b.score[rng.random(b.n_hsp) < 0.01] = -128.0        # ... and with the int8 one
    b.score[rng.random(b.n_hsp) < 0.02] = 127.0         # largest int8 value: stays inline
    b.score[rng.random(b.n_hsp) < 0.02] = 128.0         # first value beyond int8
    exp = O.build_batch(b, "hist", True)
    for bits in (8, 16, 32):
        q, eo, ep, ev = b.quantized_scores(bits)
        assert eo[-1] == len(ep) > 0
        got = engine.build_batch(b, "hist", True, score_bits=bits)
        compare_results(got, exp, exact_p=True, what=f"narrow {bits}")
    import torch
    from ortho2align_b200.engine import DeviceBatch
    for bits in (8, 16):
        db = DeviceBatch(b, torch.device("cuda:0"), aos=True, score_bits=bits, bg_bits=bits)
        engine.run(db.struct(False), engine.params("hist", True))
        compare_results(engine.fetch(), exp, exact_p=True, what=f"narrow device-resident {bits}")


@pytest.mark.parametrize("value", ["block_length", "total_length", "block_count", "weight"])
@pytest.mark.parametrize("function", ["max", "min"])
def test_best_ortholog_selection_variants(engine, value, function):
    """get_best_orthologs -value/-function (pipeline.py:981-989): device selection == oracle == the
    reference's text-streaming step applied to records formatted from the chains."""
    from oracle import oracle as O
    from ortho2align_b200 import synth
    b = synth.workload(1, n_lnc=40, seed=31, hsps_per_region=400, n_bg=500)
    exp = O.build_batch(b, "hist", True, best_value=value, best_function=function)
    got = engine.build_batch(b, "hist", True, best_value=value, best_function=function)
    compare_results(got, exp, exact_p=True, what=f"best {value}/{function}")
    # independent check: recompute the key per alignment from the chain like the reference's extract lambdas
    for l in range(b.n_lnc):
        keys = []
        for a in range(b.aln_off[l], b.aln_off[l + 1]):
            c = got.chain_idx[got.chain_off[a]:got.chain_off[a + 1]] + b.hsp_off[a]
            if len(c) == 0:
                continue
            qs, qe = b.qstart[c].astype(np.int64), b.qend[c].astype(np.int64)
            k = {"block_length": int(np.abs(qe - qs).sum()),
                 "total_length": int(max(qs.max(), qe.max()) - min(qs.min(), qe.min())),
                 "block_count": len(c), "weight": float(sum(b.score[c].tolist()))}[value]
            keys.append((a, k))
        if not keys:
            assert got.best_aln[l] == -1
            continue
        pick = (max if function == "max" else min)(keys, key=lambda t: t[1])      # first extremal element
        assert got.best_aln[l] == pick[0]


@pytest.mark.parametrize("alpha,gapopen,gapextend", [(0.01, 5, 2), (0.2, 0, 0), (0.05, 11, 1)])
def test_parameter_variants(engine, alpha, gapopen, gapextend):
    from oracle import oracle as O
    from ortho2align_b200 import synth
    b = synth.workload(2, n_lnc=24, seed=41, hsps_per_region=800, n_bg=2000)
    for fitter in ("hist", "empirical"):
        exp = O.build_batch(b, fitter, True, alpha, gapopen, gapextend)
        got = engine.build_batch(b, fitter, True, alpha, gapopen, gapextend)
        compare_results(got, exp, exact_p=True, what=f"{fitter} a={alpha} G={gapopen} E={gapextend}")


def test_kde_vs_oracle(engine):
    from oracle import oracle as O
    from ortho2align_b200 import synth
    b = synth.workload(2, n_lnc=12, seed=9, hsps_per_region=400, n_bg=3000)
    exp = O.build_batch(b, "kde", True)
    got = engine.build_batch(b, "kde", True)
    compare_results(got, exp, exact_p=False, what="kde", m_max=float(np.diff(b.hsp_off).max()))


def test_dense_chaining_and_big_rank_vs_oracle(engine):
    """BASELINE config 4 shape at a size the O(n^2) oracle finishes: every HSP survives (p = 0) and
    enters the DP -> exercises k_rank_big (radix sort) and k_chain_big (global scratch)."""
    from oracle import oracle as O
    from ortho2align_b200 import synth
    b = synth.workload(4, n_lnc=6, seed=2, hsps_per_region=3000)
    for fdr in (True, False):
        exp = O.build_batch(b, "hist", fdr)
        got = engine.build_batch(b, "hist", fdr)
        assert int(exp.n_keep.min()) == 3000
        compare_results(got, exp, exact_p=True, what=f"dense fdr={fdr}")


def test_medium_survivor_sets_vs_oracle(engine):
    """Survivor counts around the shared-memory limits of the rank (1024) and chain (192) paths."""
    from oracle import oracle as O
    from ortho2align_b200 import synth
    b = synth.workload(2, n_lnc=8, seed=4, hsps_per_region=6000, n_bg=500)
    b.score[::3] += 40.0          # a third of the HSPs become clearly significant
    exp = O.build_batch(b, "hist", True)
    got = engine.build_batch(b, "hist", True)
    assert exp.n_surv.max() > 1024 and exp.n_keep.max() > 192
    compare_results(got, exp, exact_p=True, what="medium")


def test_edge_cases(engine):
    from oracle import oracle as O
    from ortho2align_b200 import synth
    from ortho2align_b200.columnar import ColumnarBatch, concat_batches
    b = synth.workload(2, n_lnc=5, seed=8, hsps_per_region=50, n_bg=100)
    # an alignment with zero HSPs, one lncRNA with zero alignments, one HSP with orientation None
    empty_aln = ColumnarBatch(
        bg_off=np.array([0, 3], dtype=np.int64), bg=np.array([5, 5, 1], dtype=np.int32),
        aln_off=np.array([0, 1], dtype=np.int64), hsp_off=np.array([0, 0], dtype=np.int64),
        qlen=np.array([1], dtype=np.int64), slen=np.array([1], dtype=np.int64),
        qstart=np.zeros(0, np.int32), qend=np.zeros(0, np.int32), sstart=np.zeros(0, np.int32),
        send=np.zeros(0, np.int32), score=np.zeros(0, np.float64))
    no_aln = ColumnarBatch(
        bg_off=np.array([0, 2], dtype=np.int64), bg=np.array([1, 0], dtype=np.int32),
        aln_off=np.array([0, 0], dtype=np.int64), hsp_off=np.array([0], dtype=np.int64),
        qlen=np.zeros(0, np.int64), slen=np.zeros(0, np.int64),
        qstart=np.zeros(0, np.int32), qend=np.zeros(0, np.int32), sstart=np.zeros(0, np.int32),
        send=np.zeros(0, np.int32), score=np.zeros(0, np.float64))
    bad = b.slice_lnc(0, 1)
    bad.score[:] = 500.0
    bad.qend[3] = bad.qstart[3]
    full = concat_batches([b, empty_aln, no_aln, bad, b.slice_lnc(2, 4)])
    exp = O.build_batch(full, "hist", True)
    got = engine.build_batch(full, "hist", True)
    assert exp.status.tolist().count(1) == 1
    compare_results(got, exp, exact_p=True, what="edge")
    assert engine.counts.n_failed_lnc == 1
    # empty batch
    empty = concat_batches([])
    got = engine.build_batch(empty, "hist", True)
    assert got.thr.shape == (0,) and got.chain_idx.shape == (0,)


def test_device_resident_inputs_match_host_inputs(engine):
    import torch
    from ortho2align_b200 import synth
    from ortho2align_b200.engine import DeviceBatch
    b = synth.workload(2, n_lnc=16, seed=12)
    host = engine.build_batch(b, "hist", True)
    for aos in (True, False):
        db = DeviceBatch(b, torch.device("cuda:0"), aos=aos)
        torch.cuda.synchronize()
        engine.run(db.struct(with_kde=False), engine.params("hist", True))
        dev = engine.fetch()
        compare_results(dev, host, exact_p=True, what=f"device-vs-host aos={aos}")


def test_pinned_host_inputs_use_zero_copy_coordinates(engine):
    import torch
    from ortho2align_b200 import synth
    from ortho2align_b200.columnar import ColumnarBatch
    b = synth.workload(2, n_lnc=16, seed=13)
    ref = engine.build_batch(b, "hist", True)
    pinned = {}
    for k in ("bg_off", "bg", "aln_off", "hsp_off", "qlen", "slen", "qstart", "qend", "sstart", "send", "score"):
        src = getattr(b, k)
        t = torch.empty(src.shape, dtype=torch.from_numpy(src[:0]).dtype, pin_memory=True)
        t.numpy()[...] = src
        pinned[k] = t
    hb = ColumnarBatch(**{k: v.numpy() for k, v in pinned.items()})
    c4 = torch.empty((b.n_hsp, 4), dtype=torch.int32, pin_memory=True)
    c4.numpy()[...] = b.coords4()
    engine.set_coords_mode("zerocopy")          # AUTO is the host-gather mode since round 2; zero-copy is opt-in
    try:
        for coords in (None, c4.numpy()):
            engine.run(engine.host_struct(hb, False, coords=coords), engine.params("hist", True))
            assert engine.lib.o2a_coords_zero_copy(engine.h) == 1 and engine.coords_mode() == "zerocopy"
            compare_results(engine.fetch(), ref, exact_p=True, what="pinned zero-copy")
    finally:
        engine.set_coords_mode("auto")
    engine.run(engine.host_struct(hb, False, coords=c4.numpy()), engine.params("hist", True))
    assert engine.lib.o2a_coords_zero_copy(engine.h) == 0 and engine.coords_mode() == "gather"
    compare_results(engine.fetch(), ref, exact_p=True, what="pinned, default mode")


@pytest.mark.parametrize("chunks", [2, 3, 16])
@pytest.mark.parametrize("fitter", ["hist", "kde", "empirical"])
def test_chunked_host_pipeline_matches_single_pass(engine, monkeypatch, chunks, fitter):
    """Large host batches are cut into lncRNA chunks whose uploads overlap the previous chunk's kernels;
    the result must not depend on the cut (lncRNAs are independent, pipeline.py:846-860)."""
    from oracle import oracle as O
    from ortho2align_b200 import synth
    kw = dict(n_lnc=40, n_bg=300, hsps_per_region=150) if fitter == "kde" else dict(n_lnc=70)
    b = synth.workload(2, seed=41, **kw)
    exp = O.build_batch(b, fitter, True)
    monkeypatch.setenv("O2A_NO_PIPELINE", "1")
    single = engine.build_batch(b, fitter, True)
    monkeypatch.delenv("O2A_NO_PIPELINE")
    monkeypatch.setenv("O2A_PIPELINE_MIN_HSPS", "0")
    monkeypatch.setenv("O2A_CHUNKS", str(chunks))
    for kwargs in (dict(), dict(aos=True), dict(aos=True, score_bits=16, bg_bits=16), dict(aos=True, score_bits=8, bg_bits=8)):
        got = engine.build_batch(b, fitter, True, **kwargs)
        assert engine.lib.o2a_pipeline_chunks(engine.h) == chunks
        compare_results(got, single, exact_p=True, what=f"chunks={chunks} {kwargs} vs single pass")
        compare_results(got, exp, exact_p=(fitter != "kde"), what=f"chunks={chunks} {kwargs} vs oracle",
                        m_max=float(np.diff(b.hsp_off).max()))
    # ragged cut: first lncRNAs empty, one lncRNA holding most HSPs
    e = synth.workload(5, n_lnc=200, seed=42)
    monkeypatch.setenv("O2A_NO_PIPELINE", "1")
    single = engine.build_batch(e, fitter, True)
    monkeypatch.delenv("O2A_NO_PIPELINE")
    compare_results(engine.build_batch(e, fitter, True), single, exact_p=True, what="ragged chunks")


def test_pinned_result_buffers_match_fresh_arrays(engine):
    from ortho2align_b200 import synth
    b = synth.workload(2, n_lnc=20, seed=21)
    fresh = engine.build_batch(b, "hist", True)
    engine.run(engine.host_struct(b, False), engine.params("hist", True))
    compare_results(engine.fetch(pinned=True), fresh, exact_p=True, what="pinned results")
    engine.run(engine.host_struct(b.slice_lnc(0, 5), False), engine.params("hist", True))
    small = engine.fetch(pinned=True)        # buffers are reused: shapes follow the new batch
    assert small.thr.shape == (5,) and np.array_equal(small.thr, fresh.thr[:5])


def test_round_trip_properties_at_scale(engine):
    """Size-independent properties on a batch too large for the oracle to be worth running."""
    from ortho2align_b200 import synth
    b = synth.workload(2, n_lnc=600, seed=33)
    r = engine.build_batch(b, "hist", True)
    assert r.surv_off[-1] == r.n_surv.sum() == engine.counts.n_survivors
    # survivors are exactly the HSPs with score > thr of their lncRNA, in order
    lnc_of_aln = np.repeat(np.arange(b.n_lnc), np.diff(b.aln_off))
    aln_of_hsp = np.repeat(np.arange(b.n_aln), np.diff(b.hsp_off))
    mask = b.score > r.thr[lnc_of_aln][aln_of_hsp]
    assert mask.sum() == r.surv_off[-1]
    aln_of_surv = np.repeat(np.arange(b.n_aln), r.n_surv)
    gidx = b.hsp_off[:-1][aln_of_surv] + r.surv_idx
    assert np.array_equal(np.nonzero(mask)[0], gidx)
    # q = m*p/rank with rank a permutation of 1..n per alignment; keep <=> q <= alpha
    m = np.diff(b.hsp_off)[aln_of_surv].astype(float)
    with np.errstate(divide="ignore", invalid="ignore"):
        rank = np.where(r.surv_pq > 0, m * r.surv_p / r.surv_pq, 0)
    assert np.array_equal(r.surv_keep.astype(bool), r.surv_pq <= 0.05)
    for a in range(0, b.n_aln, 97):
        s0, s1 = r.surv_off[a], r.surv_off[a + 1]
        p = r.surv_p[s0:s1]
        order = np.argsort(p, kind="stable")
        rk = np.empty(len(p)); rk[order] = np.arange(len(p)) + 1
        assert np.array_equal(r.surv_pq[s0:s1], (np.diff(b.hsp_off)[a] * p) / rk)
    # chains: members are retained HSPs, strictly colinear in the chain's orientation
    for a in np.nonzero(r.chain_len > 1)[0][:200]:
        h0 = b.hsp_off[a]
        ch = r.chain_idx[r.chain_off[a]:r.chain_off[a + 1]] + h0
        assert (b.qend[ch][:-1] < b.qstart[ch][1:]).all()
        if r.chain_orient[a] == 0:
            assert (b.send[ch][:-1] < b.sstart[ch][1:]).all()
        else:
            assert (b.send[ch][:-1] > b.sstart[ch][1:]).all()
    # idempotence
    r2 = engine.build_batch(b, "hist", True)
    compare_results(r2, r, exact_p=True, what="idempotence")


@pytest.mark.parametrize("mode", ["zerocopy", "gather", "upload"])
@pytest.mark.parametrize("cfg,kw,chunks", [(2, dict(n_lnc=40), 4), (5, dict(n_lnc=900), 3), (4, dict(n_lnc=6, hsps_per_region=3000), 2),
                                           (2, dict(n_lnc=12), 1)])
def test_host_batch_layouts_and_coordinate_modes(engine, monkeypatch, mode, cfg, kw, chunks):
    """The pinned narrow layout the ingest returns (HostBatch) through every way the library can fetch the retained
    HSPs' coordinates — in place over PCIe, listed and gathered by host threads, or uploaded whole — pipelined over
    lncRNA chunks or not: identical results. Config 4 retains everything, so the gather mode takes its dense fallback."""
    from oracle import oracle as O
    from ortho2align_b200 import synth
    from ortho2align_b200.hostbatch import HostBatch
    b = synth.workload(cfg, seed=21, **kw)
    exp = O.build_batch(b, "hist", True)
    monkeypatch.setenv("O2A_PIPELINE_MIN_HSPS", "0" if chunks > 1 else str(1 << 40))
    monkeypatch.setenv("O2A_CHUNKS", str(chunks))
    engine.set_coords_mode(mode)
    try:
        hb = HostBatch.from_columnar(b, pin_coords=True)
        got = engine.build_host(hb, "hist", True)
        assert engine.coords_mode() == (mode if (mode != "zerocopy" or hb._blocks[-1].pinned) else "gather")
        engine.set_coords_mode("auto")
        compare_results(engine.build_host(hb, "hist", True), exp, exact_p=True, what=f"cfg{cfg} auto chunks={chunks}")
        assert engine.coords_mode() == "gather"
        engine.set_coords_mode(mode)
        compare_results(got, exp, exact_p=True, what=f"cfg{cfg} {mode} chunks={chunks}")
        # pageable SoA caller arrays (what a plain ColumnarBatch hands over): zero-copy is not possible there
        got = engine.build_batch(b, "hist", True)
        assert engine.coords_mode() == ("upload" if mode == "upload" else "gather")
        compare_results(got, exp, exact_p=True, what=f"cfg{cfg} {mode} chunks={chunks} pageable SoA")
        got = engine.build_batch(b, "hist", True, aos=True, score_bits=16, bg_bits=16)
        compare_results(got, exp, exact_p=True, what=f"cfg{cfg} {mode} chunks={chunks} pageable AoS int16")
    finally:
        engine.set_coords_mode("auto")


def test_understated_size_hints_fail_the_call(engine):
    """max_aln_hsps / max_lnc_bg are hints that size per-CTA scratch: an understated value must not be trusted."""
    import torch
    from ortho2align_b200 import synth
    from ortho2align_b200.engine import DeviceBatch, O2AError
    b = synth.workload(4, n_lnc=3, hsps_per_region=2500, seed=2)
    p = engine.params("hist", True)
    s = engine.host_struct(b, False)
    s.max_aln_hsps = 100
    with pytest.raises(O2AError):
        engine.run(s, p)
    db = DeviceBatch(b, torch.device("cuda", 0))
    s = db.struct(False)
    s.max_aln_hsps = 100
    with pytest.raises(O2AError):
        engine.run(s, p)
    s = db.struct(False)
    s.max_lnc_bg = 10
    with pytest.raises(O2AError):
        engine.run(s, p)
    s = db.struct(False)                     # and the engine is still usable afterwards
    engine.run(s, p)
    from oracle import oracle as O
    compare_results(engine.fetch(), O.build_batch(b, "hist", True), exact_p=True, what="after a rejected call")


def test_batch_stream_with_real_engines(engine):
    """stream.BatchStream (the scheduler bench.py's two-context e2e figure runs on): several host batches through two
    library contexts on one GPU, HostBatch and ColumnarBatch alike — results in input order, equal to a plain engine's."""
    from ortho2align_b200 import synth
    from ortho2align_b200.engine import Engine
    from ortho2align_b200.hostbatch import HostBatch
    from ortho2align_b200.stream import BatchStream
    batches = [synth.workload(2, n_lnc=6 + 3 * k, seed=50 + k, hsps_per_region=400, n_bg=900) for k in range(5)]
    want = [engine.build_batch(b, "hist", True) for b in batches]
    mixed = [HostBatch.from_columnar(b) if k % 2 == 0 else b for k, b in enumerate(batches)]
    got = list(BatchStream(devices=(0,), contexts=2).map(iter(mixed), "hist", True))
    assert len(got) == len(want)
    for k, (g, w) in enumerate(zip(got, want)):
        compare_results(g, w, exact_p=True, what=f"stream batch {k}")
    e1, e2 = Engine(0), Engine(0)
    try:
        got = list(BatchStream(engines=[e1, e2]).map(iter(mixed), "hist", True, survivors=False))
        for g, w in zip(got, want):
            assert np.array_equal(g.chain_idx, w.chain_idx) and np.array_equal(g.best_aln, w.best_aln)
        assert e1.h and e2.h                                   # caller-owned engines stay open
    finally:
        e1.close(); e2.close()


def test_cuda_graph_replay_of_a_device_batch(engine):
    """A device-resident batch issued again with identical arguments (stage timing off) is captured into a CUDA graph
    on the second call and replayed from the third: every call returns the oracle's result, a changed parameter or a
    different batch goes back to plain launches, and o2a_launch_count keeps counting the kernels of a replay."""
    import torch
    from ortho2align_b200 import synth
    from ortho2align_b200.engine import DeviceBatch, Engine
    from oracle import oracle as O
    eng = Engine(0)                                   # own context: the shared fixture has stage timing on
    try:
        eng.set_timing(False)
        dev = torch.device("cuda", 0)
        for fitter, cfg, kw in (("hist", 2, dict(n_lnc=40, seed=61)), ("kde", 2, dict(n_lnc=12, seed=62, hsps_per_region=500)),
                                ("hist", 4, dict(n_lnc=3, seed=63, hsps_per_region=2500)), ("hist", 5, dict(n_lnc=300, seed=64))):
            b = synth.workload(cfg, **kw)
            want = O.build_batch(b, fitter, True)
            db = DeviceBatch(b, dev, aos=True)
            s, p = db.struct(with_kde=(fitter == "kde")), eng.params(fitter, True)
            states, per_call = [], []
            for k in range(5):
                l0 = eng.launch_count()
                eng.run(s, p)
                states.append(eng.graph_state()); per_call.append(eng.launch_count() - l0)
                compare_results(eng.fetch(), want, exact_p=(fitter == "hist"), what=f"{fitter} C{cfg} call {k}")
            # (the first fetch allocates its staging buffers, which postpones the capture by one call)
            assert states[0] == 0 and states.count(1) == 1 and states[-2:] == [2, 2], states
            assert states.index(1) < states.index(2) and per_call[-1] == per_call[0] > 0, (states, per_call)
            # a changed parameter: plain launches again, and the result follows the parameter
            p2 = eng.params(fitter, True, alpha=0.01)
            eng.run(s, p2)
            assert eng.graph_state() == 0
            compare_results(eng.fetch(), O.build_batch(b, fitter, True, alpha=0.01), exact_p=(fitter == "hist"), what="alpha 0.01")
            eng.run(s, p)
            compare_results(eng.fetch(), want, exact_p=(fitter == "hist"), what="back to alpha 0.05")
        # stage timing on: never a graph
        eng.set_timing(True)
        for _ in range(3):
            eng.run(s, p)
            assert eng.graph_state() == 0
        assert sum(eng.stage_ms().values()) > 0
    finally:
        eng.close()
