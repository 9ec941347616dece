"""The CPU oracle against the committed fixtures and the reference's documented behaviour.

The reference has no tests or golden vectors; what can be pinned is
(i) the oracle's own regression outputs on BASELINE config 1 (tests/golden/demo_oracle_*.npz,
made by tests/golden/make_golden.py), (ii) the argument checks / quirks of Bayes.cpp, and
(iii) loose statistical ranges from README.md:162-167."""
import numpy as np
import pytest

from tests.util_demo import GOLDEN, load_demo, load_demo_T1, synth


def test_demo_fixture_matches_survey_facts():
    d = load_demo()
    g = d["geno"]
    assert g.shape == (600, 1000)
    assert [(g == k).sum() for k in (0, 1, 2)] == [355873, 196946, 47181]
    y, X = load_demo_T1()
    assert y.shape == (300,) and X.shape == (300, 1000)


@pytest.mark.parametrize("model,Pi,fold", [
    ("BayesCpi", [0.95, 0.05], None),
    ("BayesR", [0.95, 0.02, 0.02, 0.01], [0, 1e-4, 1e-3, 1e-2]),
])
def test_oracle_config1_regression(oracle, model, Pi, fold):
    y, X = load_demo_T1()
    r = oracle.bayes(y, X, model, Pi, fold=fold, niter=200, nburn=100, thin=5, seed=666666)
    gold = np.load("%s/demo_oracle_%s.npz" % (GOLDEN, model))
    assert np.array_equal(r["diag"]["tracker"], gold["tracker"])
    assert np.array_equal(r["diag"]["nnz_trace"], gold["nnz_trace"])
    assert np.array_equal(r["diag"]["nzrate_count"], gold["nzrate_count"])
    for k in ("Vg", "Ve", "h2", "mu"):
        assert np.isclose(r[k], gold[k], rtol=1e-9)
    assert np.allclose(r["alpha"], gold["alpha"], rtol=1e-7, atol=1e-12)
    assert np.allclose(r["pip"], gold["pip"])


def test_oracle_int8_and_fp64_genotypes_agree(oracle):
    y, X = load_demo_T1()
    a = oracle.bayes(y, X, "BayesCpi", [0.95, 0.05], niter=60, nburn=20, thin=5)
    b = oracle.bayes(y, X.astype(np.float64), "BayesCpi", [0.95, 0.05], niter=60, nburn=20, thin=5)
    assert np.array_equal(a["alpha"], b["alpha"]) and a["Ve"] == b["Ve"]


def test_oracle_statistical_range_on_demo(oracle):
    # README.md:162-167 reports h2 0.357 for BayesCpi on this trait (with extra model terms)
    y, X = load_demo_T1()
    r = oracle.bayes(y, X, "BayesCpi", [0.95, 0.05], niter=2000, nburn=1200, thin=5)
    assert 0.2 < r["h2"] < 0.55
    assert 150 < r["Vg"] + r["Ve"] < 280
    assert abs(r["mu"] - y.mean()) < 6


def test_oracle_argument_checks(oracle):
    y, X = load_demo_T1()
    with pytest.raises(RuntimeError, match="sum of Pi should be 1"):
        oracle.bayes(y, X, "BayesCpi", [0.9, 0.05])
    with pytest.raises(RuntimeError, match="all markers have no effect size"):
        oracle.bayes(y, X, "BayesCpi", [1.0, 0.0])
    with pytest.raises(RuntimeError, match="'fold' should be provided"):
        oracle.bayes(y, X, "BayesR", [0.95, 0.02, 0.02, 0.01])
    with pytest.raises(RuntimeError, match="length of Pi should be 2"):
        oracle.bayes(y, X, "BayesCpi", [0.9, 0.05, 0.05], fold=[0, 1, 2])
    with pytest.raises(RuntimeError, match="dfvg should not be less than 2"):
        oracle.bayes(y, X, "BayesCpi", [0.95, 0.05], dfvg=2.0)
    yy = y.copy()
    yy[3] = np.nan
    with pytest.raises(RuntimeError, match="NAs are not allowed in y"):
        oracle.bayes(yy, X, "BayesCpi", [0.95, 0.05])


def test_oracle_quirks(oracle):
    y, X = load_demo_T1()
    # unknown model strings fall through to BayesR (Bayes.cpp:97)
    # -- but only with a length-2 Pi, because of the check at :293
    a = oracle.bayes(y, X, "BayesR", [0.95, 0.05], fold=[0, 1e-2], niter=30, nburn=10)
    b = oracle.bayes(y, X, "Nonsense", [0.95, 0.05], fold=[0, 1e-2], niter=30, nburn=10)
    assert np.array_equal(a["alpha"], b["alpha"])
    with pytest.raises(RuntimeError, match="length of Pi should be 2"):
        oracle.bayes(y, X, "Nonsense", [0.95, 0.02, 0.02, 0.01], fold=[0, 1e-4, 1e-3, 1e-2], niter=30, nburn=10)
    # n_records integer division and early break (:124, :916): niter=33,nburn=10,thin=5 -> 4 records, stops at iter 30
    c = oracle.bayes(y, X, "BayesCpi", [0.95, 0.05], niter=33, nburn=10, thin=5)
    assert c["diag"]["n_records"] == 4 and c["diag"]["iters_done"] == 30
    # monomorphic SNPs are skipped but counted in m (:589, :806-813); BayesRR reports pip = 1
    d = oracle.bayes(y, X, "BayesRR", [0.95, 0.05], niter=30, nburn=10)
    mono = X.var(axis=0) == 0
    assert mono.sum() > 0 and np.all(d["alpha"][mono] == 0) and np.all(d["pip"] == 1)
    # PIP == 1 is replaced by (nzct-1)/nzct (:1030)
    e = oracle.bayes(y, X, "BayesCpi", [0.01, 0.99], niter=30, nburn=10)
    assert e["pip"].max() == (e["diag"]["nzct"] - 1) / e["diag"]["nzct"]


def test_oracle_covariates_random_effects_windows(oracle):
    d = load_demo()
    y, X = load_demo_T1()
    gid = {s: i for i, s in enumerate(d["geno_id"])}
    keep = [i for i, (p, t) in enumerate(zip(d["phe_id"], d["T1"])) if p in gid and not np.isnan(t)]
    sex = (d["sex"][keep] == "Male").astype(np.float64)[:, None]
    bwt = d["bwt"][keep][:, None]
    Cmat = np.hstack([sex, bwt])
    _, loc = np.unique(d["loc"][keep], return_inverse=True)
    _, dam = np.unique(d["dam"][keep], return_inverse=True)
    R = np.stack([loc, dam], axis=1)
    wind = (np.arange(X.shape[1]) // 50) + 1
    r = oracle.bayes(y, X, "BayesCpi", [0.95, 0.05], C_=Cmat, R=R, niter=400, nburn=200, thin=5, windindx=wind)
    assert r["beta"].shape == (2,) and r["Vr"].shape == (2,) and r["gwas"].shape == (20,)
    assert np.all(r["Vr"] >= 0) and 0 < r["h2"] < 1
    assert np.all((r["gwas"] >= 0) & (r["gwas"] < 1))
    # residual definition e = y - mu - C beta - X alpha - Z r  (Bayes.cpp:942-1011)
    e = y - r["mu"] - Cmat @ r["beta"] - X.astype(float) @ r["alpha"]
    e -= r["r"][loc] + r["r"][loc.max() + 1 + dam]
    assert np.allclose(e, r["e"], atol=1e-9)


def test_oracle_recovers_synthetic_effects(oracle):
    y, X = synth(600, 400, seed=3, n_causal=5, h2=0.6)
    r = oracle.bayes(y, X, "BayesCpi", [0.95, 0.05], niter=600, nburn=300, thin=5)
    assert 0.35 < r["h2"] < 0.8
    gebv = X.astype(float) @ r["alpha"]
    assert np.corrcoef(gebv, y)[0, 1] > 0.6
