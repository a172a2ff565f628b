"""Live pinning of the oracle against the UNMODIFIED reference imported from /root/reference.
Only runs where the reference tree exists (the build container); the committed golden vectors
(tests/test_oracle_golden.py) carry the same evidence to the GPU box."""
import numpy as np
import pytest

from oracle import ref_harness as rh

pytestmark = pytest.mark.skipif(not rh.reference_available(), reason="/root/reference not present")


def test_build_batch_against_live_reference():
    from oracle import oracle as O
    from ortho2align_b200 import synth
    b, m = synth.workload(1, n_lnc=6, with_meta=True, hsps_per_region=250, n_bg=700, seed=21)
    js = synth.to_json_dicts(b, m)
    for fitting, fdr in (("hist", True), ("hist", False)):
        res = O.build_batch(b, fitting, fdr)
        a = 0
        for l, (name, alns, bg) in enumerate(js):
            r = rh.run_reference_lncrna(alns, bg, fitting, fdr)
            assert r["thr"] == res.thr[l]
            for ra in r["alignments"]:
                s0, s1 = res.surv_off[a], res.surv_off[a + 1]
                assert ra["surv"] == res.surv_idx[s0:s1].tolist()
                assert ra["p"] == res.surv_p[s0:s1].tolist()
                assert ra["pq"] == res.surv_pq[s0:s1].tolist()
                assert ra["chain"] == res.chain_idx[res.chain_off[a]:res.chain_off[a + 1]].tolist()
                assert ra["chain_score"] == res.chain_score[a]
                a += 1


def test_random_chains_against_live_reference():
    from oracle import oracle as O
    ref = rh.import_reference()
    A = ref.alignment
    rng = np.random.default_rng(77)
    for it in range(300):
        n = int(rng.integers(0, 30))
        span = int(rng.choice([40, 1000]))
        qs = rng.integers(1, span, size=n); qe = qs + rng.integers(1, max(2, span // 5), size=n)
        ss = rng.integers(1, span, size=n); sl = rng.integers(0, max(2, span // 5), size=n)
        rev = rng.random(n) < 0.5
        se = np.where(rev, ss - sl, ss + np.maximum(sl, 1))
        sc = rng.integers(1, 60, size=n).astype(float) * (rng.uniform(0.3, 1, size=n) if it % 2 else 1.0)
        hs = [A.HSPVertex(int(qs[i]), int(qe[i]), int(ss[i]), int(se[i]),
                          (int(sc[i]) if float(sc[i]).is_integer() else float(sc[i])), kw=i) for i in range(n)]
        al = A.Alignment(hs)
        ch = al.get_best_chains(gapopen=5, gapextend=2)
        best = ch["direct"] if ch["direct"].score > ch["reverse"].score else ch["reverse"]
        chain, orient, score, s2 = O.best_chain(qs, qe, ss, se, sc, al.qlen, al.slen, 5, 2)
        assert chain.tolist() == [h.kwargs["kw"] for h in best.HSPs]
        assert s2 == (float(ch["direct"].score), float(ch["reverse"].score))


def test_random_cuts_against_live_reference():
    """to_genomic + cut_coordinates: 400 random alignments, every strand combination,
    random subsets of the four boundaries, float and int table scores — coordinates, order, scores (bit-exact),
    score types and qlen/slen identical to the reference's objects."""
    from oracle import producer as OP
    ref = rh.import_reference()
    GR, GRA, HV = ref.genomicranges.GenomicRange, ref.genomicranges.GenomicRangesAlignment, ref.alignment.HSPVertex
    rng = np.random.default_rng(77)
    n_cut = 0
    for it in range(400):
        n = int(rng.integers(0, 40))
        qs = rng.integers(0, 2000, size=n); ln = rng.integers(1, 300, size=n)
        ss = rng.integers(0, 5000, size=n); sl = ln + rng.integers(-3, 4, size=n).clip(-ln + 1, None)
        rev = rng.random(n) < 0.5
        s1, s2 = np.where(rev, ss + sl, ss), np.where(rev, ss, ss + sl)
        is_int = rng.random(n) < 0.8
        sc = np.where(is_int, rng.integers(12, 200, size=n), rng.integers(12, 200, size=n) * 0.37)
        q0, s0 = int(rng.integers(1, 10 ** 6)), int(rng.integers(1, 10 ** 6))
        qstrand, sstrand = "+-."[rng.integers(3)], "+-."[rng.integers(3)]
        relative = bool(rng.random() < 0.85)
        hsps = [HV(int(a), int(b), int(c), int(d), int(e) if i else float(e)) for a, b, c, d, e, i in zip(qs, qs + ln, s1, s2, sc, is_int)]
        aln = GRA(hsps, GR("c1", q0, q0 + 2400, qstrand), GR("c2", s0, s0 + 5400, sstrand), relative=relative)
        aln.to_genomic()
        lo_q, hi_q = (q0, q0 + 2400) if relative else (0, 2400)
        lo_s, hi_s = (s0, s0 + 5400) if relative else (0, 5400)
        bounds = {}
        for k, lo, hi in (("qleft", lo_q, hi_q), ("qright", lo_q, hi_q), ("sleft", lo_s, hi_s), ("sright", lo_s, hi_s)):
            if rng.random() < 0.6:
                bounds[k] = int(rng.integers(lo, hi))
        cut = aln.cut_coordinates(**bounds)
        rel = dict(hsp_off=np.array([0, n]), qstart=qs, qend=qs + ln, sstart=s1, send=s2, score=sc.astype(float),
                   score_is_int=is_int.astype(np.uint8))
        one = lambda v: np.array([v], dtype=np.int64)
        rg = dict(q_start=one(q0), q_end=one(q0 + 2400), q_minus=np.array([qstrand == "-"]), s_start=one(s0),
                  s_end=one(s0 + 5400), s_minus=np.array([sstrand == "-"]), relative=np.array([relative], dtype=np.uint8),
                  **{k: one(bounds.get(k, OP.NO_BOUND)) for k in ("qleft", "qright", "sleft", "sright")})
        r = OP.cut_batch(rel, rg)
        want = cut._all_HSPs
        got = np.stack([r["qstart"], r["qend"], r["sstart"], r["send"]], 1).tolist() if len(want) else []
        assert got == [[h.qstart, h.qend, h.sstart, h.send] for h in want], (it, bounds)
        assert r["score"].tolist() == [float(h.score) for h in want], (it, bounds)
        assert [c != 0 for c in r["cut"]] == [isinstance(h.score, float) for h in want]
        assert (int(r["qlen"][0]), int(r["slen"][0])) == (cut.qlen, cut.slen)
        n_cut += int((r["cut"] & 1).sum())
    assert n_cut > 200


def test_random_interval_joins_against_live_reference(tmp_path):
    """annotate_orthologs' join: fresh random lists through the reference's get_neighbours / calc_JI / calc_OC
    (genomicranges.py:2040-2067, 635-685, 497-547), and whole files through pipeline.annotate_orthologs."""
    from oracle import annotate as OA
    from oracle import make_golden_annotate as G
    from ortho2align_b200 import annotate as A
    from test_annotate import _OracleEngine, _check_join
    ref = rh.import_reference()
    for k, (n_q, n_a, distance, strandness, inverted) in enumerate(
            [(80, 200, 0, True, 0.0), (80, 200, 1200, False, 0.0), (200, 90, 0, True, 0.2), (60, 300, 50, True, 0.0)]):
        case = G.join_case(ref, 900 + k, n_q, n_a, distance, strandness, inverted=inverted, n_chrom=3, span=30000)
        _check_join(case, lambda q, a, d, s: OA.annotate_batch(q, a, d, s))
    for i, fmt in enumerate(("bed6", "gtf", "gff")):
        orth, ann = G.file_texts(950 + i, 50, 120, fmt)
        case = G.file_case(ref, orth, ann, fmt)
        po, pa, out, stats = (tmp_path / f"o{i}.bed", tmp_path / f"a{i}.{fmt}", tmp_path / f"out{i}", tmp_path / f"s{i}")
        po.write_text(orth); pa.write_text(ann)
        A.annotate_orthologs(str(po), str(pa), str(out), stats_filename=str(stats), engine=_OracleEngine())
        assert out.read_text() == case["out"]
        assert stats.read_text() == case["stats"].replace("{orthologs}", str(po))
