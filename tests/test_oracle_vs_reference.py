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
