"""The reference-side binding, compiled and run (INTEGRATION.md sections 2, 4, 5).

integration/rcpp/{Bayes,SBayes,tXXmat,read_bed}.cpp are the files a maintainer puts in place of the reference's: the SAME
exported C++ signatures (`Rcpp::List Bayes(arma::vec&, arma::mat&, std::string, ...)`, `SBayesD`, `SBayesS`, `BigStat`,
`tXXmat_Geno`, `tXXmat_Chr`, `read_bed`), the same named Rcpp::List coming back, the work done by libhibayes_b200.so through
its C ABI.  R / Rcpp / Armadillo are absent here, so they are built against the stand-in headers of oracle/ref_shim into
oracle/_ref/libhibayes_dropin.so with the same plain-buffer harness (ref_entry.cpp) that wraps the reference's own files in
oracle/_ref/libhibayes_ref.so.  These tests call both through that harness on the same arguments: what the reference's
Bayes() returns on the CPU and what the drop-in Bayes() returns from the GPU must agree (inclusion pattern of every recorded
iteration and PIP bit-exact, floats 1e-5; LD matrices and decoded genotypes bit-exact).
"""
import os

import numpy as np
import pytest

from oracle import hb_oracle
from tests.util_demo import GOLDEN, load_demo_T1, synth
from tests.util_sumstat import make_sumstat

pytestmark = pytest.mark.gpu
RTOL = 1e-5


@pytest.fixture(scope="module")
def libs():
    if hb_oracle.dropin_lib() is None:
        pytest.skip("oracle/_ref/libhibayes_dropin.so is not built")
    if hb_oracle.ref_lib() is None:
        pytest.skip("oracle/_ref/libhibayes_ref.so is not built (needs /root/reference or the prebuilt file)")
    return hb_oracle


def _three(fn, *args, seed, **kw):
    """(compiled reference on the CPU, drop-in bodies on the GPU) for the same call."""
    o = fn(*args, seed=seed, record_tape=True, store_alpha=True, **kw)
    r = fn(*args, seed=seed, replay_on_reference=o["tape"], store_alpha=True, **kw)
    d = fn(*args, seed=seed, replay_on_reference=hb_oracle.seed_tape(seed), library="dropin", store_alpha=True, **kw)
    assert d["replay"]["consumed"] == 2          # the two uniforms of seed_from_r(), nothing else
    return r, d


def _close(r, d, keys, scalars):
    sr, sd = r["MCMCsamples"], d["MCMCsamples"]
    assert np.array_equal(sr["alpha"] != 0, sd["alpha"] != 0), "inclusion pattern of a recorded iteration differs"
    assert np.array_equal(r["pip"], d["pip"])
    for k in sr:
        if sr[k].size:
            assert np.abs(sr[k] - sd[k]).max() <= RTOL * np.abs(sr[k]).max(), "MCMCsamples$" + k
    for k in keys:
        if np.asarray(r[k]).size:
            assert np.abs(r[k] - d[k]).max() <= RTOL * np.abs(r[k]).max(), k
    for k in scalars:
        assert abs(r[k] - d[k]) <= RTOL * abs(r[k]), (k, r[k], d[k])


@pytest.mark.parametrize("model,Pi,fold", [("BayesR", [0.95, 0.02, 0.02, 0.01], [0, 1e-4, 1e-3, 1e-2]), ("BayesCpi", [0.95, 0.05], None),
                                           ("BayesB", [0.95, 0.05], None), ("BayesA", [0.95, 0.05], None), ("BayesL", [0.95, 0.05], None),
                                           ("BayesRR", [0.95, 0.05], None)])
def test_bayes_signature_on_the_demo_data(libs, model, Pi, fold):
    y, X = load_demo_T1()
    r, d = _three(libs.bayes, y, X, model, Pi, fold=fold, niter=40, nburn=10, thin=3, seed=0x1234567890ABCDEF)
    _close(r, d, ("alpha", "pi", "g", "e"), ("Vg", "Ve", "h2", "mu"))


def test_bayes_signature_with_covariates_random_effects_windows(libs):
    rng = np.random.default_rng(5)
    n, m = 400, 700
    X = rng.integers(0, 3, size=(n, m)).astype(np.float64)
    X[:, 7] = 1.0
    y = X[:, :10] @ rng.normal(scale=0.4, size=10) + rng.normal(size=n) + 2
    Cm = np.column_stack([rng.normal(size=n), rng.integers(0, 2, n).astype(float)])
    R = np.column_stack([rng.integers(0, 5, size=n), rng.integers(0, 12, size=n)])
    wind = np.arange(m) // 50 + 1
    r, d = _three(libs.bayes, y, X, "BayesR", [0.9, 0.05, 0.03, 0.02], fold=[0, 1e-4, 1e-3, 1e-2], niter=30, nburn=10, thin=2, seed=5,
                  C_=Cm, R=R, windindx=wind, dfvr=4.0, s2vr=0.3)
    _close(r, d, ("alpha", "pi", "g", "e", "beta", "Vr", "r", "gwas"), ("Vg", "Ve", "h2", "mu"))


def test_bayes_signature_with_the_single_step_term(libs):
    import scipy.sparse as sp
    rng = np.random.default_rng(3)
    n, m, ne, qe = 500, 900, 160, 220
    X = rng.integers(0, 3, size=(n, m)).astype(np.float64)      # integer rows: the device tiles are int8
    J = np.concatenate([-np.ones(n - ne), rng.uniform(-1, 0, ne)])
    y = X[:, :30] @ rng.normal(scale=0.3, size=30) + rng.normal(size=n) + 1.5
    A = sp.random(qe, qe, density=0.015, random_state=5, format="csr", data_rvs=lambda k: rng.uniform(-0.15, 0.15, k))
    G = sp.csc_matrix(A + A.T + sp.diags(np.full(qe, 2.0)))
    index1 = rng.permutation(qe)[:ne] + 1
    r, d = _three(libs.bayes, y, X, "BayesCpi", [0.9, 0.1], niter=14, nburn=4, thin=2, seed=99, epsl_y_J=J, epsl_Gi=G, epsl_index=index1)
    _close(r, d, ("alpha", "g", "e", "epsilon"), ("Vg", "Ve", "h2", "mu", "Veps", "J"))


def test_bayes_signature_with_the_bslmm_term(libs):
    rng = np.random.default_rng(12)
    n, m = 320, 600
    X = rng.integers(0, 3, size=(n, m)).astype(np.float64)
    y = X[:, :25] @ rng.normal(scale=0.3, size=25) + rng.normal(size=n) + 0.5
    Xc = X - X.mean(axis=0)
    Kval, K = np.linalg.eigh(Xc @ Xc.T / m + 0.01 * np.eye(n))
    r, d = _three(libs.bayes, y, X, "BSLMM", [0.9, 0.1], niter=10, nburn=4, thin=2, seed=321, Kival=Kval, Ki=np.asfortranarray(K))
    _close(r, d, ("alpha", "g", "e"), ("Vg", "Ve", "h2", "mu"))


def test_bayes_signature_errors_are_the_references(libs):
    y, X = synth(50, 20, seed=2, n_causal=2)
    for kw, text in [(dict(Pi=[0.5, 0.4]), "sum of Pi should be 1."), (dict(Pi=[0.9, 0.1], dfvg=1.5), "dfvg should not be less than 2.")]:
        Pi = kw.pop("Pi")
        with pytest.raises(RuntimeError, match=text):
            libs.bayes(y, X, "BayesCpi", Pi, niter=4, nburn=2, thin=1, seed=1, replay_on_reference=hb_oracle.seed_tape(1), library="dropin", **kw)


@pytest.mark.parametrize("model,Pi,fold", [("BayesR", [0.95, 0.02, 0.02, 0.01], [0, 1e-4, 1e-3, 1e-2]), ("BayesCpi", [0.95, 0.05], None),
                                           ("BayesA", [0.95, 0.05], None)])
def test_sbayes_signatures(libs, model, Pi, fold):
    from tests.test_sbayes import _sparse_ld
    y, X = synth(900, 300, seed=23, n_causal=8)
    ss, ld = make_sumstat(y, X, n_na=3)
    wind = np.arange(300) // 20 + 1 if model != "BayesA" else None
    kw = dict(fold=fold, niter=30, nburn=10, thin=4, windindx=wind)
    for fn, L in ((libs.sbayesd, ld), (libs.sbayess, _sparse_ld(ld, len(y)))):
        r, d = _three(fn, ss, L, model, Pi, seed=4242, **kw)
        _close(r, d, ("alpha", "pi") + (("gwas",) if wind is not None else ()), ("Vg", "Ve", "h2"))


def test_ldmat_signatures(libs):
    """BigStat(), tXXmat_Geno(), tXXmat_Chr(): what the drop-in bodies return from the tcgen05 kernel = what the reference's
    own loops return, every bit, incl. which entries the arma::sp_mat stores."""
    y, X = synth(333, 300, seed=4)
    X[:, 5] = 1
    a, b = libs.ref_bigstat(X), libs.ref_bigstat(X, library="dropin")
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    chr_ = np.repeat([1, 2, 3], 100)
    for c, q in [(None, None), (None, 3.84), (None, 0.0), (chr_, None), (chr_, 3.84), (chr_, 0.0)]:
        r, sr = libs.ref_txxmat(X, chr=c, chisq=q)
        d, sd = libs.ref_txxmat(X, chr=c, chisq=q, library="dropin")
        assert np.array_equal(r, d, equal_nan=True), (c is not None, q)
        assert sr == sd


@pytest.mark.parametrize("impute,dominance", [(True, False), (False, False), (True, True)])
def test_read_bed_signature(libs, tmp_path, impute, dominance):
    from tests.util_bed import make_bed
    d = np.load(os.path.join(GOLDEN, "demo_bed.npz"))
    p = tmp_path / "demo.bed"
    p.write_bytes(d["bed"].tobytes())
    assert np.array_equal(libs.ref_read_bed(str(p), 600, 1000, impute=impute, dominance=dominance),
                          libs.ref_read_bed(str(p)[:-4], 600, 1000, impute=impute, dominance=dominance, library="dropin"))   # (suffix added, :99-102)
    img, _ = make_bed(203, 77, seed=5, p_missing=0.15, all_missing_cols=(9,))
    q = tmp_path / "ragged.bed"
    q.write_bytes(img.tobytes())
    assert np.array_equal(libs.ref_read_bed(str(q), 203, 77, impute=impute, dominance=dominance),
                          libs.ref_read_bed(str(q), 203, 77, impute=impute, dominance=dominance, library="dropin"))
