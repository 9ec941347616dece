"""SBayesD: the oracle on CPU, and the CUDA path against it."""
import os

import numpy as np
import pytest

from tests.util_demo import load_demo_T1, synth
from tests.util_sumstat import make_sumstat

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
RTOL = 1e-5


def test_oracle_sbayesd_runs_and_respects_the_quirks(oracle):
    y, X = synth(800, 300, seed=12, n_causal=10)
    ss, ld = make_sumstat(y, X, n_na=7, seed=3)
    wind = (np.arange(300) // 25) + 1
    r = oracle.sbayesd(ss, ld, "BayesCpi", [0.9, 0.1], niter=60, nburn=20, thin=5, windindx=wind, seed=5)
    na = np.isnan(ss[:, 1])
    assert r["diag"]["n_used"] == 800 and r["diag"]["n_records"] == 8
    assert np.all(r["alpha"][na] == 0) and np.all(r["pip"][na] == 0)          # SNPs with NA statistics are skipped (:100-103)
    assert 0 < r["h2"] < 1 and r["Vg"] > 0 and r["Ve"] > 0
    assert r["gwas"].shape == (12,) and np.all((r["gwas"] >= 0) & (r["gwas"] < 1))
    # effects explain the marginal statistics: r_hat = xy - n LD g  (SBayesD.cpp:105, 351-356)
    n = r["diag"]["n_used"]
    # dense models keep every SNP in
    r2 = oracle.sbayesd(ss, ld, "BayesRR", [0.9, 0.1], niter=30, nburn=10, thin=5, seed=5)
    assert np.all(r2["pip"] == 1.0)
    with pytest.raises(RuntimeError, match="fold"):
        oracle.sbayesd(ss, ld, "BayesR", [0.9, 0.05, 0.05], niter=10, nburn=5, thin=1)


GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _check_pin(r, pin):
    """oracle output against a committed pin (tests/golden/make_golden.py): discrete outputs exactly, the rest to 1e-9"""
    assert np.array_equal(r["diag"]["tracker"], pin["tracker"])
    assert np.array_equal(r["diag"]["nnz_trace"], pin["nnz_trace"])
    assert np.array_equal(r["diag"]["nzrate_count"], pin["nzrate_count"])
    assert r["diag"]["n_used"] == int(pin["n_used"])
    for k in ("Vg", "Ve", "h2"):
        assert abs(r[k] / float(pin[k]) - 1) < 1e-9, (k, r[k], float(pin[k]))
    assert np.allclose(r["alpha"], pin["alpha"], rtol=1e-9, atol=1e-12 * np.abs(pin["alpha"]).max())
    assert np.allclose(r["pi"], pin["pi"], rtol=1e-9)
    assert np.allclose(r["diag"]["vare_trace"], pin["vare_trace"], rtol=1e-9)
    assert np.allclose(r["diag"]["vara_trace"], pin["vara_trace"], rtol=1e-9)


PIN_MODELS = [("BayesCpi", [0.95, 0.05], None), ("BayesR", [0.95, 0.02, 0.02, 0.01], [0, 1e-4, 1e-3, 1e-2])]
PIN_KW = dict(niter=100, nburn=50, thin=5, seed=666666)


@pytest.mark.parametrize("model,Pi,fold", PIN_MODELS)
def test_oracle_sbayesd_pinned_on_the_reference_cojo_file(oracle, model, Pi, fold):
    """SBayesD oracle on the reference's own summary statistics (inst/extdata/demo.ma, columns MAF/BETA/SE/NMISS as
    R/sbayes.r:209 selects them) with the LD matrix of the bundled genotypes (centred X'X/n, tXXmat.cpp:174-179)."""
    d = np.load(os.path.join(GOLDEN, "demo.npz"))
    X = d["geno"].astype(np.float64)
    Xc = X - X.mean(axis=0)
    ld = np.asfortranarray(Xc.T @ Xc / X.shape[0])
    ss = np.asfortranarray(np.column_stack([d["ma_maf"], d["ma_beta"], d["ma_se"], d["ma_n"]]))
    r = oracle.sbayesd(ss, ld, model, Pi, fold=fold, **PIN_KW)
    _check_pin(r, np.load(os.path.join(GOLDEN, "demo_oracle_sbayesd_%s.npz" % model)))


@pytest.mark.parametrize("model,Pi,fold", PIN_MODELS)
def test_oracle_sbayess_pinned(oracle, model, Pi, fold):
    """SBayesS oracle on stored inputs (summary statistics + thresholded LD matrix of a synthetic data set)."""
    import scipy.sparse as sp
    inp = np.load(os.path.join(GOLDEN, "sbayess_inputs.npz"))
    r = oracle.sbayess(np.asfortranarray(inp["sumstat"]), sp.csc_matrix(inp["ld_thresholded"]), model, Pi, fold=fold, **PIN_KW)
    _check_pin(r, np.load(os.path.join(GOLDEN, "demo_oracle_sbayess_%s.npz" % model)))


def _compare(got, ref):
    assert np.array_equal(got["diag"]["tracker"], ref["diag"]["tracker"])
    assert np.array_equal(got["diag"]["nnz_trace"], ref["diag"]["nnz_trace"])
    assert np.array_equal(got["diag"]["nzrate_count"], ref["diag"]["nzrate_count"])
    assert np.array_equal(got["pip"], ref["pip"])
    assert got["diag"]["n_records"] == ref["diag"]["n_records"] and got["diag"]["iters_done"] == ref["diag"]["iters_done"]
    for k in ("Vg", "Ve", "h2"):
        assert abs(got[k] / ref[k] - 1) < RTOL, (k, got[k], ref[k])
    scale = np.abs(ref["alpha"]).max() + 1e-300
    assert np.abs(got["alpha"] - ref["alpha"]).max() < RTOL * scale
    assert np.allclose(got["pi"], ref["pi"], rtol=RTOL)
    assert np.allclose(got["diag"]["r_hat"], ref["diag"]["r_hat"], rtol=RTOL, atol=RTOL * np.abs(ref["diag"]["r_hat"]).max())
    assert np.allclose(got["diag"]["vare_trace"], ref["diag"]["vare_trace"], rtol=RTOL)
    assert np.allclose(got["diag"]["vara_trace"], ref["diag"]["vara_trace"], rtol=RTOL)


@pytest.mark.gpu
@pytest.mark.parametrize("model,Pi,fold", MODELS)
def test_sbayesd_demo_all_models(oracle, model, Pi, fold):
    """sbrm()-style inputs from the bundled demo genotypes (300 x 1000): every model string against the oracle."""
    import hibayes_b200 as hb
    y, X = load_demo_T1()
    ss, ld = make_sumstat(y, X)
    kw = dict(niter=60, nburn=30, thin=5, seed=666666)
    ref = oracle.sbayesd(ss, ld, model, Pi, fold=fold, **kw)
    got = hb.SBayesD(ss, ld, model, Pi, fold=fold, **kw)
    _compare(got, ref)


@pytest.mark.gpu
def test_sbayesd_na_rows_windows_ragged(oracle):
    import hibayes_b200 as hb
    y, X = synth(1200, 777, seed=31, n_causal=12)   # m not a multiple of the tile
    ss, ld = make_sumstat(y, X, n_na=20, seed=2)
    wind = (np.arange(777) // 40) + 1
    kw = dict(niter=40, nburn=10, thin=3, windindx=wind, seed=11)
    ref = oracle.sbayesd(ss, ld, "BayesR", [0.9, 0.05, 0.03, 0.02], fold=[0, 1e-4, 1e-3, 1e-2], **kw)
    got = hb.SBayesD(ss, ld, "BayesR", [0.9, 0.05, 0.03, 0.02], fold=[0, 1e-4, 1e-3, 1e-2], **kw)
    _compare(got, ref)
    assert np.array_equal(got["diag"]["wppa_count"], ref["diag"]["wppa_count"])
    assert np.array_equal(got["gwas"], ref["gwas"])


def _sparse_ld(ld, n, chisq=10.0):
    """ldmat(..., chisq=) style sparsification (tXXmat.cpp:146-153): keep r^2 * n > chisq and the diagonal."""
    import scipy.sparse as sp
    d = np.sqrt(np.clip(np.diag(ld), 1e-300, None))
    r2 = (ld / d[:, None] / d[None, :]) ** 2
    keep = (r2 * n > chisq) | np.eye(ld.shape[0], dtype=bool)
    return sp.csc_matrix(np.where(keep, ld, 0.0))


def test_oracle_sbayess_differs_from_dense_only_through_its_own_rules(oracle):
    y, X = synth(900, 250, seed=4, n_causal=8)
    ss, ld = make_sumstat(y, X)
    import scipy.sparse as sp
    kw = dict(niter=30, nburn=10, thin=5, seed=3)
    full = oracle.sbayess(ss, sp.csc_matrix(ld), "BayesCpi", [0.9, 0.1], **kw)
    dense = oracle.sbayesd(ss, ld, "BayesCpi", [0.9, 0.1], **kw)
    # a sparse matrix that stores every entry has varediff = 0: SBayesS then equals SBayesD unless a re-draw happens
    assert np.array_equal(full["diag"]["tracker"], dense["diag"]["tracker"])
    assert np.allclose(full["alpha"], dense["alpha"], rtol=1e-12, atol=1e-15)
    sparse = oracle.sbayess(ss, _sparse_ld(ld, 900), "BayesCpi", [0.9, 0.1], **kw)
    assert sparse["Ve"] > 0 and not np.allclose(sparse["alpha"], dense["alpha"])


@pytest.mark.gpu
@pytest.mark.parametrize("model,Pi,fold", MODELS)
def test_sbayess_all_models(oracle, model, Pi, fold):
    """Every model string with a sparsified LD matrix (n > m so that the thresholded matrix stays well conditioned)."""
    import hibayes_b200 as hb
    y, X = synth(1500, 400, seed=23, n_causal=10)
    ss, ld = make_sumstat(y, X)
    sld = _sparse_ld(ld, len(y))
    kw = dict(niter=60, nburn=30, thin=5, seed=4242)
    ref = oracle.sbayess(ss, sld, model, Pi, fold=fold, **kw)
    assert np.all(np.isfinite(ref["alpha"])) and np.isfinite(ref["Ve"])
    got = hb.SBayesS(ss, sld, model, Pi, fold=fold, **kw)
    _compare(got, ref)


@pytest.mark.gpu
@pytest.mark.parametrize("model,Pi,fold", [("BayesCpi", [0.5, 0.5], None), ("BayesR", [0.4, 0.2, 0.2, 0.2], [0, 1e-3, 1e-2, 1e-1])])
def test_sbayess_redraw_loop(oracle, model, Pi, fold):
    """Inflated marginal effects make g^2 vx exceed the phenotypic variance: the re-draw loop of SBayesS.cpp:388-398
    (including its give-up after 100 tries and the overwritten sum of squares) must agree with the oracle."""
    import hibayes_b200 as hb
    y, X = synth(600, 200, seed=9, n_causal=5)
    ss, ld = make_sumstat(y, X)
    ss[[17, 90, 150], 1] *= 40.0
    sld = _sparse_ld(ld, 600, chisq=5.0)
    kw = dict(niter=30, nburn=10, thin=4, seed=77)
    ref = oracle.sbayess(ss, sld, model, Pi, fold=fold, **kw)
    got = hb.SBayesS(ss, sld, model, Pi, fold=fold, **kw)
    _compare(got, ref)


@pytest.mark.gpu
@pytest.mark.parametrize("model,Pi,fold", [("BayesCpi", [0.9, 0.1], None), ("BayesR", [0.9, 0.05, 0.03, 0.02], [0, 1e-4, 1e-3, 1e-2]),
                                           ("BayesRR", [0.0, 1.0], None)])
def test_sbayess_csc_banded_many_tiles(oracle, model, Pi, fold):
    """The CSC data path of the device (SBayesS.cpp:292-296, 403-407): a banded LD matrix over 12 tiles (m not a multiple of
    256) whose stored pattern is NOT symmetric (entries dropped from one triangle only) and has an empty column; the device
    holds the compressed columns, not a dense copy."""
    import scipy.sparse as sp
    import hibayes_b200 as hb
    rng = np.random.default_rng(12)
    n, m = 2500, 2900
    # genotypes with local correlation: neighbouring SNPs share a latent block
    base = rng.integers(0, 3, size=(n, m // 4 + 2))
    X = np.empty((n, m), dtype=np.float64)
    for j in range(m):
        flip = rng.random(n) < 0.35
        X[:, j] = np.where(flip, rng.integers(0, 3, size=n), base[:, j // 4])
    y = X[:, rng.choice(m, 15, replace=False)] @ rng.normal(scale=0.4, size=15) + rng.normal(size=n)
    ss, ld = make_sumstat(y, X)
    i, j = np.indices((m, m))
    band = np.abs(i - j) <= 300                                  # reaches the neighbouring tile and one beyond
    d = np.sqrt(np.clip(np.diag(ld), 1e-300, None))
    r2n = (ld / d[:, None] / d[None, :]) ** 2 * n
    keep = (band & (r2n > 6.0)) | np.eye(m, dtype=bool)
    keep &= ~((i > j) & ((i + 3 * j) % 11 == 0))                 # lower-triangle entries without their mirror image
    keep[:, 1234] = False                                        # a column with nothing stored, not even the diagonal
    ss[1234, 1] = np.nan                                         # (xpx = 0 there: the SNP is not estimated)
    sld = sp.csc_matrix(np.where(keep, ld, 0.0))
    assert sld.nnz < 0.05 * m * m
    kw = dict(niter=25, nburn=10, thin=3, seed=515)
    ref = oracle.sbayess(ss, sld, model, Pi, fold=fold, **kw)
    assert np.all(np.isfinite(ref["alpha"])) and np.isfinite(ref["Ve"])
    got = hb.SBayesS(ss, sld, model, Pi, fold=fold, **kw)
    _compare(got, ref)
    assert got["diag"]["ld_bytes_device"] == sld.nnz * 12 + (m + 1) * 4     # compressed columns, no dense copy
    assert got["diag"]["columns_total"] > 0 and got["diag"]["ld_entries_total"] > 0


@pytest.mark.gpu
def test_sbayesd_many_tiles(oracle):
    """Dense LD over 10 tiles: the decisions of a tile overlap the previous tile's column updates (one grid barrier per
    tile), rows of the next tile are brought up to date by the deciding CTA itself."""
    import hibayes_b200 as hb
    y, X = synth(3000, 2500, seed=71, n_causal=30)
    ss, ld = make_sumstat(y, X, n_na=15, seed=4)
    kw = dict(niter=20, nburn=8, thin=3, seed=31)
    ref = oracle.sbayesd(ss, ld, "BayesR", [0.9, 0.05, 0.03, 0.02], fold=[0, 1e-4, 1e-3, 1e-2], **kw)
    got = hb.SBayesD(ss, ld, "BayesR", [0.9, 0.05, 0.03, 0.02], fold=[0, 1e-4, 1e-3, 1e-2], **kw)
    _compare(got, ref)
    assert got["diag"]["ld_bytes_device"] == 2500 * 2500 * 8
