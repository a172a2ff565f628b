"""Parity comparison helpers shared by the CPU and GPU test modules."""
import numpy as np

# Tolerances of the parity contract (BASELINE.json north_star + SURVEY.md §7 hard part 2):
#   hist / empirical: p, q bit-exact (well inside the 1e-12 relative bar)
#   kde: |dp| <= max(1e-12 * p, 16 * 2^-53)  — scipy's ndtr cannot be reproduced bit-for-bit
KDE_REL = 1e-12
KDE_ABS = 16 * 2.0 ** -53
KDE_THR_ABS = 1e-10     # brentq stops at xtol = 2e-12; its path depends on 1e-16 differences of f


def assert_pq_close(a, b, what, abs_scale=1.0):
    """abs_scale: q = m*p/rank carries p's absolute floor multiplied by m/rank <= m."""
    a = np.asarray(a, dtype=float); b = np.asarray(b, dtype=float)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    tol = np.maximum(KDE_REL * np.abs(b), KDE_ABS * abs_scale)
    bad = np.abs(a - b) > tol
    assert not bad.any(), f"{what}: {bad.sum()} values differ, max |d|={np.abs(a - b).max()}"


def compare_results(got, exp, exact_p=True, what="", m_max=1.0):
    """got/exp: BatchResult. Integer / index / decision outputs must be identical; p/q bit-exact when
    exact_p else within the KDE tolerance (m*p/rank inherits p's relative error)."""
    assert np.array_equal(got.status, exp.status), f"{what} status"
    ok = exp.status == 0
    if exact_p:
        assert np.array_equal(got.thr[ok], exp.thr[ok]), f"{what} thr: {got.thr[ok][:4]} vs {exp.thr[ok][:4]}"
    else:
        assert np.allclose(got.thr[ok], exp.thr[ok], rtol=0, atol=KDE_THR_ABS), f"{what} thr"
    for k in ("n_surv", "n_keep", "chain_len", "surv_off", "surv_idx", "surv_keep", "chain_off", "chain_idx", "best_aln"):
        g, e = getattr(got, k), getattr(exp, k)
        assert g.shape == e.shape, f"{what} {k}: shape {g.shape} vs {e.shape}"
        assert np.array_equal(g, e), f"{what} {k}: first diff at {np.nonzero(g != e)[0][:5]}"
    nz = exp.chain_len > 0
    assert np.array_equal(got.chain_orient[nz], exp.chain_orient[nz]), f"{what} chain_orient"
    assert np.array_equal(got.chain_score, exp.chain_score), f"{what} chain_score"
    assert np.array_equal(got.orient_scores, exp.orient_scores), f"{what} orient_scores"
    if exact_p:
        assert np.array_equal(got.surv_p, exp.surv_p), f"{what} surv_p max|d|={np.abs(got.surv_p - exp.surv_p).max() if len(exp.surv_p) else 0}"
        assert np.array_equal(got.surv_pq, exp.surv_pq), f"{what} surv_pq"
        assert np.array_equal(got.chain_pq, exp.chain_pq), f"{what} chain_pq"
    else:
        assert_pq_close(got.surv_p, exp.surv_p, f"{what} surv_p")
        assert_pq_close(got.surv_pq, exp.surv_pq, f"{what} surv_pq", abs_scale=m_max)
        assert_pq_close(got.chain_pq, exp.chain_pq, f"{what} chain_pq", abs_scale=m_max)


def compare_with_golden(res, batch, gold, exact_p=True, what=""):
    """res: BatchResult of the golden's batch; gold: a tests/golden/build_*.json dict."""
    fdr = gold["fdr"]
    a = 0
    for l, ln in enumerate(gold["lncrnas"]):
        if exact_p:
            assert res.thr[l] == ln["thr"], f"{what} thr[{l}] {res.thr[l]} vs {ln['thr']}"
        else:
            assert abs(res.thr[l] - ln["thr"]) <= KDE_THR_ABS
        for ra in ln["alignments"]:
            s0, s1 = res.surv_off[a], res.surv_off[a + 1]
            assert res.surv_idx[s0:s1].tolist() == ra["surv"], f"{what} survivors of alignment {a}"
            if exact_p:
                assert res.surv_p[s0:s1].tolist() == ra["p"], f"{what} p of alignment {a}"
                assert res.surv_pq[s0:s1].tolist() == ra["pq"], f"{what} q of alignment {a}"
            else:
                assert_pq_close(res.surv_p[s0:s1], ra["p"], f"{what} p of alignment {a}")
                m = float(batch.hsp_off[a + 1] - batch.hsp_off[a])
                assert_pq_close(res.surv_pq[s0:s1], ra["pq"], f"{what} q of alignment {a}", abs_scale=m if fdr else 1.0)
            assert [bool(x) for x in res.surv_keep[s0:s1]] == ra["keep"], f"{what} keep of alignment {a}"
            c0, c1 = res.chain_off[a], res.chain_off[a + 1]
            assert res.chain_idx[c0:c1].tolist() == ra["chain"], f"{what} chain of alignment {a}"
            if ra["chain"]:
                assert int(res.chain_orient[a]) == ra["orient"]
            assert res.chain_score[a] == ra["chain_score"], f"{what} chain score of alignment {a}"
            assert res.orient_scores[a, 0] == ra["score_direct"] and res.orient_scores[a, 1] == ra["score_reverse"]
            a += 1
    assert a == batch.n_aln
