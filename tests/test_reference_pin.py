"""The oracle pinned against the REFERENCE ITSELF.

oracle/_ref/libhibayes_ref.so is /root/reference/src/{Bayes,SBayesD,SBayesS,stats,solver}.cpp compiled unmodified against
stand-in R / Rcpp / Armadillo headers (oracle/ref_shim, oracle/Makefile).  Its random draws replay the tape of variates
the oracle consumed, in the order of the reference's own sampler calls; every draw is checked for kind and (gamma,
chi-square) for a bit-identical shape, so a different call order or degree of freedom anywhere aborts the run.

Three layers:
  * live (CPU, needs the library; built here where /root/reference exists, prebuilt on the GPU box): oracle vs compiled
    reference on the demo data, synthetic data with covariates / random effects / windows / the single-step term / the
    BSLMM term, SBayesD, SBayesS incl. its re-draw loop -- effects of every recorded iteration BIT-IDENTICAL
    (BayesL: 1e-9, see below), posterior means to 1e-12;
  * fixtures (CPU, any box): tests/golden/ref_*.npz hold what the compiled reference returned
    (tests/golden/make_ref_golden.py); the oracle must reproduce them;
  * GPU: the CUDA path against the same reference fixtures directly (bit-exact inclusion pattern and PIP, 1e-5 on floats).

BayesL: oracle and device evaluate the smaller root of the inverse-Gaussian quadratic (stats.cpp:57-59) in its
cancellation-free form; the reference's literal expression loses digits to cancellation, and the chain carries that to
1e-12 (n > m) ... 4e-10 (demo data, m > n) relative.  Everything else is the same arithmetic in the same order.
"""
import os

import numpy as np
import pytest

from oracle import hb_oracle
from tests.util_demo import GOLDEN, load_demo_T1, synth
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
KEEP = slice(0, 1000, 7)
BAYESL_RTOL = 1e-7   # (the docstring's BayesL paragraph; 4e-10 observed on the demo data, where m > n makes the chain stiff)


@pytest.fixture(scope="module")
def ref():
    if hb_oracle.ref_lib() is None:
        pytest.skip("oracle/_ref/libhibayes_ref.so not built (needs /root/reference; the GPU box uses the prebuilt file)")
    return hb_oracle


def _pair(fn, *args, **kw):
    """(oracle result, compiled-reference result) of the same call, the reference fed with the oracle's tape."""
    o = fn(*args, record_tape=True, store_alpha=True, **kw)
    r = fn(*args, replay_on_reference=o["tape"], store_alpha=True, **kw)
    assert r["replay"]["consumed"] == len(o["tape"]), "the reference made fewer sampler calls than the oracle consumed"
    return o, r


def _same(o, r, exact=True, keys=("alpha", "pip", "pi"), scalars=("Vg", "Ve", "h2"), rtol=1e-9):
    so, sr = o["MCMCsamples"], r["MCMCsamples"]
    assert np.array_equal(so["alpha"] != 0, sr["alpha"] != 0), "inclusion pattern of a recorded iteration differs"
    assert np.array_equal(o["pip"], r["pip"])
    if exact:
        # every recorded effect, variance and pi: the same bits
        for k in ("alpha", "Vg", "Ve", "pi"):
            assert np.array_equal(so[k], sr[k]), "MCMCsamples$%s differs from the compiled reference" % k
    for k in so:
        if so[k].size == 0:
            continue
        assert np.abs(so[k] - sr[k]).max() <= rtol * np.abs(sr[k]).max(), k
    for k in keys:
        assert np.abs(o[k] - r[k]).max() <= rtol * np.abs(r[k]).max(), k
    for k in scalars:
        assert abs(o[k] - r[k]) <= rtol * abs(r[k]), (k, o[k], r[k])


# ---- live: oracle vs the compiled reference ---------------------------------------------------------------------------
@pytest.mark.parametrize("model,Pi,fold", MODELS)
def test_bayes_demo_data_all_models(ref, model, Pi, fold):
    """BASELINE config 1 (inst/extdata/demo, T1 ~ 1) through the reference's own Bayes() (Bayes.cpp:60-1094)."""
    y, X = load_demo_T1()
    o, r = _pair(ref.bayes, y, X, model, Pi, fold=fold, niter=40, nburn=10, thin=3, seed=11)
    _same(o, r, exact=model != "BayesL", keys=("alpha", "pip", "pi", "g", "e"), scalars=("Vg", "Ve", "h2", "mu"), rtol=BAYESL_RTOL if model == "BayesL" else 1e-12)


@pytest.mark.parametrize("model,Pi,fold", [("BayesR", [0.9, 0.05, 0.03, 0.02], [0, 1e-4, 1e-3, 1e-2]), ("BayesCpi", [0.9, 0.1], None),
                                           ("BayesBpi", [0.9, 0.1], None)])
def test_bayes_covariates_random_effects_windows(ref, model, Pi, fold):
    """Bayes.cpp:484-516 (covariates, environmental random effects with makeZ()'s sorted levels), :826-845 (PIP / WPPA),
    a monomorphic SNP (:589)."""
    rng = np.random.default_rng(5)
    n, m = 200, 120
    X = rng.integers(0, 3, size=(n, m)).astype(np.float64)
    X[:, 7] = 1.0
    y = X[:, :10] @ rng.normal(scale=0.4, size=10) + rng.normal(size=n) + 2
    Cm = np.column_stack([rng.normal(size=n), rng.integers(0, 2, n).astype(float)])
    R = np.column_stack([rng.integers(0, 5, size=n), rng.integers(0, 12, size=n)])
    wind = np.arange(m) // 10 + 1
    o, r = _pair(ref.bayes, y, X, model, Pi, fold=fold, niter=30, nburn=10, thin=2, seed=5, C_=Cm, R=R, windindx=wind, dfvr=4.0, s2vr=0.3)
    _same(o, r, keys=("alpha", "pip", "pi", "g", "e", "beta", "Vr", "r", "gwas"), scalars=("Vg", "Ve", "h2", "mu"))


@pytest.mark.parametrize("drop_diag", [(), (0, 5)])
@pytest.mark.parametrize("model,Pi,fold", [("BayesRR", [0.0, 1.0], None), ("BayesR", [0.9, 0.05, 0.03, 0.02], [0, 1e-4, 1e-3, 1e-2])])
def test_bayes_single_step_term(ref, model, Pi, fold, drop_diag):
    """Bayes.cpp:254-275, 554-584 with Gibbs(sp_mat) of solver.cpp:131-140 -- real-valued (imputed) rows, an entry of Gi
    without a stored diagonal.  The sparse sums are taken in another order here and there: 1e-9, not bits."""
    from tests.test_single_step import single_step_case
    y, X, J, G, index1 = single_step_case(7, drop_diag=drop_diag)
    o, r = _pair(ref.bayes, y, X, model, Pi, fold=fold, niter=20, nburn=6, thin=2, seed=31337, epsl_y_J=J, epsl_Gi=G, epsl_index=index1)
    _same(o, r, exact=False, keys=("alpha", "pip", "g", "e", "epsilon"), scalars=("Vg", "Ve", "h2", "mu", "Veps", "J"))


@pytest.mark.parametrize("model,Pi", [("BSLMM", [0.9, 0.1]), ("BayesRR", [0.0, 1.0])])
def test_bayes_bslmm_term(ref, model, Pi):
    """Bayes.cpp:203-233, 518-552, 955-964 (Ki / Kival; arma::randn draws through the tape as well)."""
    from tests.test_single_step import bslmm_case
    y, X, Kval, K = bslmm_case(3)
    o, r = _pair(ref.bayes, y, X, model, Pi, niter=16, nburn=4, thin=2, seed=2024, Kival=Kval, Ki=K)
    _same(o, r, exact=False, keys=("alpha", "pip", "g", "e"), scalars=("Vg", "Ve", "h2", "mu"))


@pytest.mark.parametrize("model,Pi,fold", MODELS)
def test_sbayesd_and_sbayess_all_models(ref, model, Pi, fold):
    """SBayesD.cpp:5-609 and SBayesS.cpp:21-679 on summary statistics with NA rows, windows for the mixture models."""
    from tests.test_sbayes import _sparse_ld
    y, X = synth(700, 150, seed=23, n_causal=8)
    ss, ld = make_sumstat(y, X, n_na=3)
    wind = np.arange(150) // 10 + 1 if model in ("BayesR", "BayesCpi") else None
    kw = dict(fold=fold, niter=30, nburn=10, thin=4, seed=4242, windindx=wind)
    for fn, L in ((ref.sbayesd, ld), (ref.sbayess, _sparse_ld(ld, len(y)))):
        o, r = _pair(fn, ss, L, model, Pi, **kw)
        _same(o, r, exact=model != "BayesL", keys=("alpha", "pip", "pi") + (("gwas",) if wind is not None else ()),
              rtol=BAYESL_RTOL if model == "BayesL" else 1e-12)


@pytest.mark.parametrize("model,Pi,fold", [("BayesCpi", [0.5, 0.5], None), ("BayesR", [0.4, 0.2, 0.2, 0.2], [0, 1e-3, 1e-2, 1e-1])])
def test_sbayess_redraw_loop(ref, model, Pi, fold):
    """SBayesS.cpp:388-398 / :489-499: the re-draw loop, its give-up after 100 tries, the overwritten sum of squares."""
    from tests.test_sbayes import _sparse_ld
    y, X = synth(600, 200, seed=9, n_causal=5)
    ss, ld = make_sumstat(y, X)
    ss[[17, 90, 150], 1] *= 40.0
    o, r = _pair(ref.sbayess, ss, _sparse_ld(ld, 600, chisq=5.0), model, Pi, fold=fold, niter=30, nburn=10, thin=4, seed=77)
    assert len(o["tape"]) > 30 * (200 + 8) * 1.5, "the case is meant to re-draw"
    _same(o, r)


def test_the_tape_check_catches_a_different_call_order(ref):
    """One draw removed from the tape: every later draw is off by one position and the kinds stop matching."""
    y, X = synth(120, 60, seed=1, n_causal=4)
    kw = dict(fold=None, niter=6, nburn=2, thin=2, seed=3, store_alpha=True)
    o = ref.bayes(y, X, "BayesCpi", [0.9, 0.1], record_tape=True, **kw)
    tape = np.delete(o["tape"], 5)
    with pytest.raises(RuntimeError, match="tape (mismatch|exhausted)"):
        ref.bayes(y, X, "BayesCpi", [0.9, 0.1], replay_on_reference=tape, **kw)
    bad = o["tape"].copy()
    k = int(np.flatnonzero(bad["kind"] == 3)[2])
    bad["param"][k] += 1.0            # another degree of freedom for one chi-square draw
    with pytest.raises(RuntimeError, match="tape mismatch"):
        ref.bayes(y, X, "BayesCpi", [0.9, 0.1], replay_on_reference=bad, **kw)


@pytest.mark.parametrize("kwargs,text", [
    (dict(model="BayesCpi", Pi=[0.5, 0.4]), "sum of Pi should be 1."),
    (dict(model="BayesCpi", Pi=[1.0, 0.0]), "all markers have no effect size."),
    (dict(model="BayesR", Pi=[0.9, 0.1]), "'fold' should be provided for BayesR model."),
    (dict(model="BayesCpi", Pi=[0.9, 0.1], dfvg=1.5), "dfvg should not be less than 2."),
    (dict(model="BayesCpi", Pi=[0.9, 0.05, 0.05], fold=[0, 1, 2]), "length of Pi should be 2, the first value is the proportion of non-effect markers."),
])
def test_error_texts_are_the_references(ref, kwargs, text):
    """The checks of Bayes.cpp:92-117, 325, 356: the oracle refuses with the text the compiled reference throws."""
    y, X = synth(50, 20, seed=2, n_causal=2)
    kw = dict(niter=4, nburn=2, thin=1, seed=1)
    kw.update(kwargs)
    model, Pi = kw.pop("model"), kw.pop("Pi")
    with pytest.raises(RuntimeError) as eo:
        ref.bayes(y, X, model, Pi, **kw)
    with pytest.raises(RuntimeError) as er:
        ref.bayes(y, X, model, Pi, replay_on_reference=np.zeros(0, dtype=hb_oracle.TAPE_DTYPE), **kw)
    assert text in str(eo.value) and text in str(er.value)


def test_dense_models_with_windows_fail_in_the_reference(ref):
    """BayesRR / BayesA / BayesL keep no snptracker (Bayes.cpp:292-299), so `snptracker.elem(windxi)` (:839) is out of
    bounds and Armadillo throws: the reference cannot run WPPA for these models.  (The product returns zero counts.)"""
    y, X = synth(60, 30, seed=2, n_causal=2)
    kw = dict(niter=6, nburn=2, thin=2, seed=1, windindx=np.arange(30) // 10 + 1, store_alpha=True)
    o = ref.bayes(y, X, "BayesRR", [0.0, 1.0], record_tape=True, **kw)
    with pytest.raises(RuntimeError, match="index out of bounds"):
        ref.bayes(y, X, "BayesRR", [0.0, 1.0], replay_on_reference=o["tape"], **kw)


# ---- fixtures: what the compiled reference returned, stored ------------------------------------------------------------------
def _fixture_inputs(kind):
    import scipy.sparse as sp
    if kind == "bayes":
        return load_demo_T1()
    if kind == "sbayesd":
        d = np.load(os.path.join(GOLDEN, "demo.npz"))
        G = d["geno"].astype(np.float64)
        Gc = G - G.mean(axis=0)
        return (np.asfortranarray(np.column_stack([d["ma_maf"], d["ma_beta"], d["ma_se"], d["ma_n"]])),
                np.asfortranarray(Gc.T @ Gc / G.shape[0]))
    inp = np.load(os.path.join(GOLDEN, "sbayess_inputs.npz"))
    return np.asfortranarray(inp["sumstat"]), sp.csc_matrix(inp["ld_thresholded"])


def _fixture_kw(f):
    return dict(niter=int(f["niter"]), nburn=int(f["nburn"]), thin=int(f["thin"]), seed=int(f["seed"]),
                fold=list(f["fold"]) if f["fold"].size else None)


def _check_against_fixture(got, f, model, rtol):
    """got: a run of the oracle (rtol 1e-9: the same arithmetic) or of the CUDA path (rtol 1e-5, north_star's bar)."""
    sa = got["MCMCsamples"]["alpha"]
    assert np.array_equal((sa != 0).sum(axis=0), f["nnz_per_record"])
    assert np.array_equal(sa[KEEP, :] != 0, f["store_alpha_subset"] != 0)
    assert np.array_equal(got["pip"], f["pip"])
    sc = np.abs(f["store_alpha_subset"]).max()
    if model == "BayesL":
        rtol = max(rtol, BAYESL_RTOL)
    elif rtol <= 1e-9:
        assert np.array_equal(sa[KEEP, :], f["store_alpha_subset"])
        assert np.array_equal(got["MCMCsamples"]["Vg"], f["store_Vg"]) and np.array_equal(got["MCMCsamples"]["Ve"], f["store_Ve"])
    assert np.abs(sa[KEEP, :] - f["store_alpha_subset"]).max() <= rtol * sc
    assert np.abs(got["alpha"] - f["alpha"]).max() <= rtol * np.abs(f["alpha"]).max()
    assert np.allclose(got["MCMCsamples"]["Vg"], f["store_Vg"], rtol=rtol) and np.allclose(got["MCMCsamples"]["Ve"], f["store_Ve"], rtol=rtol)
    assert np.allclose(got["MCMCsamples"]["pi"], f["store_pi"], rtol=rtol, atol=1e-300)
    for k in ("Vg", "Ve", "h2"):
        assert abs(got[k] - float(f[k])) <= rtol * abs(float(f[k])), k
    if "mu" in f.files:
        assert abs(got["mu"] - float(f["mu"])) <= rtol * abs(float(f["mu"]))
        assert np.abs(got["g"] - f["g"]).max() <= rtol * np.abs(f["g"]).max()
        assert np.abs(got["e"] - f["e"]).max() <= rtol * np.abs(f["e"]).max()


@pytest.mark.parametrize("kind", ["bayes", "sbayesd", "sbayess"])
@pytest.mark.parametrize("model,Pi,fold", MODELS)
def test_oracle_reproduces_the_reference_fixtures(kind, model, Pi, fold):
    f = np.load(os.path.join(GOLDEN, "ref_%s_%s.npz" % (kind, model)))
    assert int(f["tape_consumed"]) == int(f["tape_len"])
    a, b = _fixture_inputs(kind)
    fn = {"bayes": hb_oracle.bayes, "sbayesd": hb_oracle.sbayesd, "sbayess": hb_oracle.sbayess}[kind]
    got = fn(a, b, model, list(f["Pi"]), store_alpha=True, **_fixture_kw(f))
    _check_against_fixture(got, f, model, 1e-9)


# ---- GPU: the CUDA path against the reference's outputs ------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("model,Pi,fold", MODELS)
def test_gpu_bayes_against_the_reference_fixtures(model, Pi, fold):
    """hb_bayes() on the demo data against what the compiled reference returned for the same arguments and variates."""
    import hibayes_b200 as hb
    f = np.load(os.path.join(GOLDEN, "ref_bayes_%s.npz" % model))
    y, X = _fixture_inputs("bayes")
    got = hb.Bayes(y, X, model, list(f["Pi"]), store_alpha=True, **_fixture_kw(f))
    _check_against_fixture(got, f, model, 1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["sbayesd", "sbayess"])
@pytest.mark.parametrize("model,Pi,fold", MODELS)
def test_gpu_sbayes_against_the_reference_fixtures(kind, model, Pi, fold):
    import hibayes_b200 as hb
    f = np.load(os.path.join(GOLDEN, "ref_%s_%s.npz" % (kind, model)))
    ss, ld = _fixture_inputs(kind)
    got = (hb.SBayesD if kind == "sbayesd" else hb.SBayesS)(ss, ld, model, list(f["Pi"]), store_alpha=True, **_fixture_kw(f))
    _check_against_fixture(got, f, model, 1e-5)


# ---- the stages in front of the sweeps: LD builder and .bed decoder (no random numbers: plain equality) -----------------------
def test_bigstat_and_txxmat_are_the_references(ref):
    """oracle/hb_oracle_ld.c against BigStat(), tXXmat_Geno(), tXXmat_Chr() of tXXmat.cpp as compiled: every entry the
    same bits, the number of stored entries of the returned arma::sp_mat = the non-zeros of the oracle's matrix."""
    y, X = synth(333, 90, seed=4)
    X[:, 5] = 1                                         # a monomorphic SNP: xx = 0, r = NaN in the sparse branch
    a, b = ref.bigstat(X), ref.ref_bigstat(X)
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    chr_ = np.repeat([1, 2, 3], 30)
    for c, q in [(None, None), (None, 3.84), (chr_, None), (chr_, 3.84), (chr_, 0.0)]:
        o = ref.txxmat(X, chr=c, chisq=q)
        r, stored = ref.ref_txxmat(X, chr=c, chisq=q)
        assert np.array_equal(o, r, equal_nan=True), (c is not None, q)
        if c is not None or q is not None:
            assert stored == np.count_nonzero(o)
    # tXXmat_Geno takes its sparse branch only for chisq > 0 (tXXmat.cpp:117-120); ldmat_plan() maps that before the ABI
    r0, stored0 = ref.ref_txxmat(X, chisq=0.0)
    assert stored0 == 90 * 90 and np.array_equal(r0, ref.txxmat(X))


@pytest.mark.parametrize("impute,dominance", [(True, False), (False, False), (True, True)])
def test_read_bed_is_the_references(ref, tmp_path, impute, dominance):
    """oracle read_bed against read_bed<char>() of read_bed.cpp as compiled, on the reference's own demo.bed and on a file
    with missing genotypes, a number of individuals that is not a multiple of 4, more SNPs than one buffer holds."""
    from tests.util_bed import make_bed
    d = np.load(os.path.join(GOLDEN, "demo_bed.npz"))
    p = tmp_path / "demo.bed"
    p.write_bytes(d["bed"].tobytes())
    want, _ = ref.read_bed(d["bed"], 600, 1000, impute=impute, dominance=dominance)
    assert np.array_equal(ref.ref_read_bed(str(p), 600, 1000, impute=impute, dominance=dominance), want)
    img, _ = make_bed(203, 77, seed=5, p_missing=0.15, all_missing_cols=(9,))
    q = tmp_path / "ragged.bed"
    q.write_bytes(img.tobytes())
    want, _ = ref.read_bed(img, 203, 77, impute=impute, dominance=dominance)
    for max_line in (10000, 16):
        assert np.array_equal(ref.ref_read_bed(str(q), 203, 77, impute=impute, dominance=dominance, max_line=max_line), want)


@pytest.mark.parametrize("model,Pi,fold", [("BayesR", [0.95, 0.02, 0.02, 0.01], [0, 1e-4, 1e-3, 1e-2]), ("BayesBpi", [0.9, 0.1], None),
                                           ("BayesA", [0.9, 0.1], None)])
def test_bayes_long_chain_stays_bit_identical(ref, model, Pi, fold):
    """150 iterations on n = 1 500 x m = 2 500 (about 4e5 draws replayed, 3.75e5 SNP updates): oracle and compiled reference do
    not drift apart -- every recorded effect still the same bits at the end."""
    y, X = synth(1500, 2500, seed=77, n_causal=25)
    o, r = _pair(ref.bayes, y, X, model, Pi, fold=fold, niter=150, nburn=50, thin=10, seed=909)
    assert len(o["tape"]) > 150 * 2500
    _same(o, r, keys=("alpha", "pip", "pi", "g", "e"), scalars=("Vg", "Ve", "h2", "mu"), rtol=1e-12)


@pytest.mark.parametrize("model,Pi,fold", [("BayesCpi", [0.9, 0.1], None), ("BayesBpi", [0.9, 0.1], None), ("BayesL", [0.9, 0.1], None)])
def test_bayes_single_step_real_valued_rows_other_models(ref, model, Pi, fold):
    """The single-step call as ssbrm() makes it (R/ssbayes.r:305-321: real-valued imputed rows, J covariate, sparse Gi) for
    more models, at a size where the imputed rows matter (a third of the records)."""
    from tests.test_single_step import single_step_case
    y, X, J, G, index1 = single_step_case(21, n=240, m=90, ne=80, qe=110)
    o, r = _pair(ref.bayes, y, X, model, Pi, fold=fold, niter=24, nburn=8, thin=2, seed=4711, epsl_y_J=J, epsl_Gi=G, epsl_index=index1)
    _same(o, r, exact=False, keys=("alpha", "pip", "g", "e", "epsilon"), scalars=("Vg", "Ve", "h2", "mu", "Veps", "J"),
          rtol=BAYESL_RTOL if model == "BayesL" else 1e-9)
