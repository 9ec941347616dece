"""GPU parity tests: the CUDA path through the C ABI against the CPU oracle on the same inputs.
Bar: SNP class labels, per-iteration NnzSnp, PIP and window counts bit-exact; effects, variance
components, residuals within 1e-5 relative (BASELINE.json north_star) -- in practice ~1e-10."""
import numpy as np
import pytest

import hibayes_b200 as hb
from tests.util_demo import load_demo, load_demo_T1, synth

pytestmark = pytest.mark.gpu

RTOL = 1e-5

MODELS = [
    ("BayesCpi", [0.95, 0.05], None),
    ("BayesC", [0.95, 0.05], None),
    ("BayesB", [0.95, 0.05], None),
    ("BayesBpi", [0.95, 0.05], None),
    ("BayesR", [0.95, 0.02, 0.02, 0.01], [0, 1e-4, 1e-3, 1e-2]),
    ("BayesRR", [0.95, 0.05], None),
    ("BayesA", [0.95, 0.05], None),
    ("BayesL", [0.95, 0.05], None),
]


def _compare(got, ref, exact_classes=True):
    if exact_classes:
        assert np.array_equal(got["diag"]["tracker"], ref["diag"]["tracker"])
        assert np.array_equal(got["diag"]["nnz_trace"], ref["diag"]["nnz_trace"])
        assert np.array_equal(got["diag"]["nzrate_count"], ref["diag"]["nzrate_count"])
        assert np.array_equal(got["pip"], ref["pip"])
    assert got["diag"]["n_records"] == ref["diag"]["n_records"]
    assert got["diag"]["iters_done"] == ref["diag"]["iters_done"]
    for k in ("Vg", "Ve", "h2", "mu"):
        assert abs(got[k] / ref[k] - 1) < RTOL, (k, got[k], ref[k])
    scale = np.abs(ref["alpha"]).max() + 1e-300
    assert np.abs(got["alpha"] - ref["alpha"]).max() < RTOL * scale
    assert np.allclose(got["pi"], ref["pi"], rtol=RTOL)
    assert np.allclose(got["g"], ref["g"], rtol=RTOL, atol=RTOL * np.abs(ref["g"]).max())
    assert np.allclose(got["e"], ref["e"], rtol=RTOL, atol=RTOL * np.abs(ref["e"]).max())
    assert np.allclose(got["diag"]["vare_trace"], ref["diag"]["vare_trace"], rtol=RTOL)
    assert np.allclose(got["diag"]["vara_trace"], ref["diag"]["vara_trace"], rtol=RTOL, atol=1e-12)


def test_device_synth_matches_host_and_column_stats():
    n, m = 3000, 500
    X = hb.synth_geno_host(n, m, seed=20260101)
    e = hb.Engine(n, m, seed=1)
    e.synth_geno(20260101)
    xpx, sumx = e.col_stats()
    Xd = X.astype(np.float64)
    assert np.array_equal(xpx, (Xd * Xd).sum(axis=0))
    assert np.array_equal(sumx, Xd.sum(axis=0))
    alpha = np.random.default_rng(0).normal(size=m)
    alpha[::3] = 0
    assert np.allclose(e.predict(alpha), Xd @ alpha, rtol=1e-12, atol=1e-9)
    # loading the host copy gives the same device state
    e2 = hb.Engine(n, m, seed=1)
    e2.load_geno(X)
    x2, s2 = e2.col_stats()
    assert np.array_equal(x2, xpx) and np.array_equal(s2, sumx)
    e3 = hb.Engine(n, m, seed=1)
    e3.load_geno(Xd)
    assert np.array_equal(e3.col_stats()[0], xpx)
    e.close(); e2.close(); e3.close()


@pytest.mark.parametrize("n,m,tile,lag", [(700, 300, 64, 3), (2100, 200, 128, 2)])
def test_gram_band_is_exact(n, m, tile, lag):
    X = hb.synth_geno_host(n, m, seed=5)
    e = hb.Engine(n, m, tile_snps=tile, lag_tiles=lag)
    e.load_geno(X)
    e.build_gram()
    G = e.get_gram()
    T = G.shape[0]
    Xp = np.zeros((n, (T + lag) * tile), dtype=np.int64)
    Xp[:, :m] = X
    for t in range(T):
        for dt in range(lag):
            ref = Xp[:, t * tile:(t + 1) * tile].T @ Xp[:, (t + dt) * tile:(t + dt + 1) * tile]
            if t + dt >= T:
                ref[:] = 0
            assert np.array_equal(G[t, dt], ref), (t, dt)
    e.close()


@pytest.mark.parametrize("model,Pi,fold", MODELS)
def test_config1_demo_data_all_models(oracle, model, Pi, fold):
    """BASELINE config 1: ibrm() on inst/extdata/demo, T1 ~ 1, 200 iterations."""
    y, X = load_demo_T1()
    kw = dict(niter=200, nburn=100, thin=5, seed=666666)
    ref = oracle.bayes(y, X, model, Pi, fold=fold, **kw)
    got = hb.Bayes(y, X, model, Pi, fold=fold, **kw)
    _compare(got, ref)


def test_config1_matches_committed_golden():
    y, X = load_demo_T1()
    from tests.util_demo import GOLDEN
    gold = np.load("%s/demo_oracle_BayesCpi.npz" % GOLDEN)
    got = hb.Bayes(y, X, "BayesCpi", [0.95, 0.05], niter=200, nburn=100, thin=5, seed=666666)
    assert np.array_equal(got["diag"]["tracker"], gold["tracker"])
    assert np.array_equal(got["diag"]["nnz_trace"], gold["nnz_trace"])
    assert np.allclose(got["alpha"], gold["alpha"], rtol=1e-6, atol=1e-10)
    assert abs(got["Ve"] / gold["Ve"] - 1) < RTOL


@pytest.mark.parametrize("tile,lag,slabs", [(64, 1, 0), (64, 2, 0), (64, 8, 0), (128, 3, 0), (64, 4, 5), (64, 4, 2), (256, 4, 0), (256, 1, 3), (128, 8, 0)])
def test_result_does_not_depend_on_tiling(oracle, tile, lag, slabs):
    y, X = synth(1500, 1000, seed=21, n_causal=15)
    kw = dict(niter=12, nburn=4, thin=2, seed=31337)
    ref = oracle.bayes(y, X, "BayesR", [0.9, 0.05, 0.03, 0.02], fold=[0, 1e-4, 1e-3, 1e-2], **kw)
    got = hb.Bayes(y, X, "BayesR", [0.9, 0.05, 0.03, 0.02], fold=[0, 1e-4, 1e-3, 1e-2], tile_snps=tile, lag_tiles=lag,
                   n_slabs=slabs, **kw)
    _compare(got, ref)


def test_ragged_shapes(oracle):
    # n not a multiple of 16, m not a multiple of the tile, monomorphic columns at tile borders
    y, X = synth(333, 517, seed=8, n_causal=6)
    X[:, 0] = 0; X[:, 63] = 2; X[:, 64] = 1; X[:, 516] = 0
    kw = dict(niter=20, nburn=10, thin=5, seed=5)
    ref = oracle.bayes(y, X, "BayesCpi", [0.9, 0.1], **kw)
    got = hb.Bayes(y, X, "BayesCpi", [0.9, 0.1], **kw)
    _compare(got, ref)
    assert np.all(got["alpha"][[0, 63, 64, 516]] == 0)


def test_covariates_random_effects_windows(oracle):
    d = load_demo()
    y, X = load_demo_T1()
    gid = {s: i for i, s in enumerate(d["geno_id"])}
    keep = [i for i, (p, t) in enumerate(zip(d["phe_id"], d["T1"])) if p in gid and not np.isnan(t)]
    Cmat = np.hstack([(d["sex"][keep] == "Male").astype(np.float64)[:, None], d["bwt"][keep][:, None]])
    _, loc = np.unique(d["loc"][keep], return_inverse=True)
    _, dam = np.unique(d["dam"][keep], return_inverse=True)
    R = np.stack([loc, dam], axis=1)
    wind = (np.arange(X.shape[1]) // 50) + 1
    kw = dict(C_=Cmat, R=R, niter=120, nburn=60, thin=5, windindx=wind, seed=99)
    ref = oracle.bayes(y, X, "BayesCpi", [0.95, 0.05], **kw)
    got = hb.Bayes(y, X, "BayesCpi", [0.95, 0.05], **kw)
    _compare(got, ref)
    assert np.array_equal(got["diag"]["wppa_count"], ref["diag"]["wppa_count"])
    assert np.array_equal(got["gwas"], ref["gwas"])
    assert np.allclose(got["beta"], ref["beta"], rtol=RTOL)
    assert np.allclose(got["Vr"], ref["Vr"], rtol=RTOL)
    assert np.allclose(got["r"], ref["r"], rtol=RTOL, atol=1e-9)


def test_medium_synthetic_bayesr(oracle):
    y, X = synth(5000, 8192, seed=20260101, n_causal=80)
    kw = dict(niter=8, nburn=2, thin=2, seed=20260101)
    ref = oracle.bayes(y, X, "BayesR", [0.95, 0.02, 0.02, 0.01], fold=[0, 1e-4, 1e-3, 1e-2], **kw)
    got = hb.Bayes(y, X, "BayesR", [0.95, 0.02, 0.02, 0.01], fold=[0, 1e-4, 1e-3, 1e-2], **kw)
    _compare(got, ref)


def test_sweep_is_deterministic_run_to_run():
    y, X = synth(2000, 2048, seed=3, n_causal=20)
    kw = dict(niter=6, nburn=2, thin=2, seed=77)
    a = hb.Bayes(y, X, "BayesR", [0.95, 0.02, 0.02, 0.01], fold=[0, 1e-4, 1e-3, 1e-2], **kw)
    b = hb.Bayes(y, X, "BayesR", [0.95, 0.02, 0.02, 0.01], fold=[0, 1e-4, 1e-3, 1e-2], **kw)
    assert np.array_equal(a["alpha"], b["alpha"]) and np.array_equal(a["g"], b["g"]) and a["Ve"] == b["Ve"]


def test_large_shape_invariants():
    """Size-independent properties at a shape the oracle cannot run (n = 50 000 rows as in the metric, m = 32 768):
    the residual and genetic values the sweeps leave behind equal y - mu - X g and X g recomputed from the final
    effects, class counts add up, and two runs agree bit for bit."""
    n, m = 50000, 32768
    e = hb.Engine(n, m, seed=77)
    e.synth_geno(20260101)
    xpx, sumx = e.col_stats()
    active = (n * xpx != sumx * sumx)
    e.set_snp_info(xpx, active.astype(np.uint8))
    e.build_gram()
    rng = np.random.default_rng(1)
    beta = np.zeros(m)
    idx = rng.choice(m, 100, replace=False)
    beta[idx] = rng.standard_normal(100)
    gv = e.predict(beta)
    y = gv * np.sqrt(0.5 / gv.var()) + rng.normal(scale=np.sqrt(0.5), size=n)
    r0 = y - y.mean()
    vx = (xpx - sumx * sumx / n) / (n - 1)
    varg = 0.25 * y.var() / (0.05 * vx.sum())
    fold = [0.0, 1e-4, 1e-3, 1e-2]
    logpi = list(np.log([0.95, 0.02, 0.02, 0.01]))

    def run():
        e.set_residual(r0)
        e.set_u(np.zeros(n))
        e.set_effects(np.zeros(m))
        outs = []
        for it in range(3):
            outs.append(e.sweep(iter=it, model_index=6, vare=0.5 * y.var(), logpi=logpi, vara_fold=[varg * f for f in fold], fold=fold,
                                rnorm2_bound=float(r0 @ r0) * 2))
        return outs, e.get_residual(), e.get_u(), e.get_effects(), e.get_tracker()

    o1, r1, u1, g1, t1 = run()
    o2, r2, u2, g2, t2 = run()
    assert np.array_equal(t1, t2) and np.array_equal(g1, g2) and np.array_equal(r1, r2) and np.array_equal(u1, u2)
    xg = e.predict(g1)
    scale = np.abs(r0).max()
    assert np.abs(u1 - xg).max() < 1e-9 * scale                      # u = X g         (Bayes.cpp:789)
    assert np.abs(r1 - (r0 - xg)).max() < 1e-9 * scale               # yadj = y - mu - X g (:787)
    cnt = np.array(o1[-1]["count"][:4])
    assert cnt.sum() == active.sum() and np.array_equal(cnt, np.bincount(t1[active], minlength=4))
    assert abs(o1[-1]["sum_r2"] - r1 @ r1) < 1e-9 * (r1 @ r1)
    assert np.all((g1 != 0) == (t1 > 0))
    e.close()


@pytest.mark.gpu
def test_loader_chunks_and_range_checks(monkeypatch):
    """hb_engine_load_geno_i8 / _f64 (engine.cu, load_chunked): pinned double buffering over many small chunks gives the
    same device tiles as one chunk (same chain, bit for bit); values outside {0,1,2} are refused -- int8 by the check inside
    the pack kernel, fp64 by the converting host threads -- with the place named."""
    import hibayes_b200 as hb
    y, X = synth(700, 900, seed=31, n_causal=9)
    kw = dict(model="BayesR", Pi=[0.95, 0.02, 0.02, 0.01], fold=[0, 1e-4, 1e-3, 1e-2], niter=8, nburn=2, thin=2, seed=77)
    one = hb.Bayes(y, X, **kw)
    monkeypatch.setenv("HB_LOAD_CHUNK", str(700 * 37))       # 37 columns per chunk: 25 chunks, both bounce buffers reused
    many = hb.Bayes(y, X, **kw)
    many64 = hb.Bayes(y, X.astype(np.float64), **kw)
    for got in (many, many64):
        assert np.array_equal(got["diag"]["tracker"], one["diag"]["tracker"])
        assert np.array_equal(got["alpha"], one["alpha"]) and got["Ve"] == one["Ve"]
    for bad_value in (3, -1, 100):
        Xb = X.copy()
        Xb[17, 523] = bad_value
        with pytest.raises(RuntimeError, match=r"genotype value %d outside \{0,1,2\} \(column 523\)" % bad_value):
            hb.Bayes(y, Xb, **kw)
    Xf = X.astype(np.float64)
    Xf[5, 877] = 0.5
    with pytest.raises(RuntimeError, match=r"genotype \(5,877\) = 0\.5"):
        hb.Bayes(y, Xf, **kw)
