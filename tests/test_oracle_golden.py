"""The C restatement (oracle/) against the golden vectors recorded from the unmodified reference.
CPU only. These tests are what pins the oracle on machines where /root/reference is absent."""
import numpy as np
import pytest

from conftest import batch_from_lists, load_golden
from helpers import KDE_THR_ABS, assert_pq_close, compare_with_golden
from oracle import oracle as O
from ortho2align_b200.columnar import kde_factor


@pytest.mark.parametrize("name,exact", [("build_hist_fdr.json", True), ("build_hist_nofdr.json", True),
                                        ("build_c5_hist_fdr.json", True), ("build_kde_fdr.json", False)])
def test_build_batch_matches_reference_golden(name, exact):
    g = load_golden(name)
    b = batch_from_lists(g["batch"])
    res = O.build_batch(b, g["fitting"], g["fdr"], g["alpha"], g["gapopen"], g["gapextend"])
    assert (res.status == 0).all()
    compare_with_golden(res, b, g, exact_p=exact, what=name)


def test_fitters_match_reference_golden():
    g = load_golden("fitters.json")
    for case in g["cases"]:
        bg = np.asarray(case["bg"], dtype=np.int32)
        hbins, hcdf = O.hist_fit(bg)
        assert hbins[:8].tolist() == case["hist_hbins_head"][:len(hbins[:8])]
        assert hcdf[-8:].tolist() == case["hist_hcdf_tail"][-len(hcdf[-8:]):]
        assert float(np.sum(hcdf)) == case["hist_hcdf_sum"]
        assert O.hist_sf(case["x"], hbins, hcdf).tolist() == case["hist_sf"]
        assert [O.hist_isf(a, hbins, hcdf) for a in case["alphas"]] == case["hist_isf"]
        if "kde_sf" in case:
            f = kde_factor(len(bg))
            assert f == case["kde_factor"]          # host-side numpy formula reproduces scipy's factor
            assert_pq_close(O.kde_sf(case["x"], bg, f), case["kde_sf"], "kde sf")
            for a, want in zip(case["alphas"][:2], case["kde_isf"]):
                assert abs(O.kde_isf(a, bg, f) - want) <= KDE_THR_ABS


def test_ranking_stable_ties():
    g = load_golden("fitters.json")
    for rc in g["ranking"]:
        p = np.asarray(rc["p"], dtype=np.float64)
        r = O.ranking(p)
        assert r.tolist() == rc["rank"]
        q = (rc["m"] * p) / r if len(p) else p
        assert q.tolist() == rc["q"]


def test_chains_match_reference_golden():
    g = load_golden("chains.json")
    for i, c in enumerate(g["cases"]):
        chain, orient, score, s2 = O.best_chain(c["qstart"], c["qend"], c["sstart"], c["send"], c["score"],
                                                c["qlen"], c["slen"], c["gapopen"], c["gapextend"])
        assert chain.tolist() == c["chain"], f"case {i}"
        assert s2 == (c["score_direct"], c["score_reverse"]), f"case {i}"
        if c["chain"]:
            assert orient == c["orient"]


def test_pairwise_sum_is_numpy_sum():
    rng = np.random.default_rng(0)
    for n in list(range(0, 300)) + [1000, 4097, 65537, 100001]:
        a = rng.random(n) * rng.choice([1e-3, 1, 1e3], size=n)
        assert O.pairwise_sum(a) == float(np.sum(a))


def test_orientation_none_is_an_exception():
    with pytest.raises(KeyError):
        O.best_chain([10, 30], [20, 30], [5, 50], [9, 60], [30.0, 40.0], 100, 100)


def test_empirical_extension_matches_scipy_ecdf():
    from scipy import stats
    rng = np.random.default_rng(3)
    bg = rng.integers(12, 60, size=777).astype(np.int32)
    x = np.concatenate([rng.uniform(10, 62, size=50), bg[:20].astype(float)])
    want = stats.ecdf(bg).sf.evaluate(x)
    got = O.emp_sf(x, bg)
    # scipy evaluates 1 - k/n; the extension is defined as the exact rational (n - k)/n
    assert np.allclose(got, want, rtol=0, atol=2e-16)
    assert np.array_equal(got, np.array([(bg > v).sum() / len(bg) for v in x]))
    thr = O.emp_isf(0.05, bg)
    assert O.emp_sf([thr], bg)[0] <= 0.05
    below = np.sort(bg)[np.sort(bg) < thr]
    assert len(below) == 0 or O.emp_sf([below[-1]], bg)[0] > 0.05
