"""CPU tests of the stages in front of the sweeps: the oracle of the .bed decoder
(read_bed.cpp:97-232) against an independent numpy decode of the reference's bundled demo.bed and against a
pure-Python restatement on files with missing genotypes; the host build of the device decoder's byte code
against the oracle; the oracle of the LD builder (tXXmat.cpp) against numpy; the branch logic of ldmat()
(R/ldm.r:44-94)."""
import os

import numpy as np
import pytest

import hibayes_b200 as hb
from hibayes_b200 import _lib
from tests.util_bed import make_bed, py_read_bed
from tests.util_demo import GOLDEN, load_demo


def test_oracle_decodes_the_reference_demo_bed(oracle):
    img = np.load(os.path.join(GOLDEN, "demo_bed.npz"))["bed"]
    d = load_demo()
    nid, m = d["geno"].shape
    assert img.shape[0] == 3 + m * ((nid + 3) // 4)
    got, miss = oracle.read_bed(img, nid, m)
    assert np.array_equal(got, d["geno"])          # independent numpy decode (tests/golden/make_golden.py)
    assert not miss.any()
    # genotype counts quoted in SURVEY.md 8c for this file
    assert [(got == v).sum() for v in (0, 1, 2)] == [355873, 196946, 47181]
    dom, _ = oracle.read_bed(img, nid, m, dominance=True)
    assert np.array_equal(dom, (d["geno"] == 1).astype(np.int8))


@pytest.mark.parametrize("nid", [1, 4, 37, 130])
@pytest.mark.parametrize("dominance", [False, True])
@pytest.mark.parametrize("impute", [True, False])
def test_oracle_imputation_against_python_restatement(oracle, nid, dominance, impute):
    img, _ = make_bed(nid, 60, seed=100 + nid, p_missing=0.15)
    got, miss = oracle.read_bed(img, nid, 60, impute=impute, dominance=dominance)
    want, wmiss = py_read_bed(img, nid, 60, impute, dominance)
    assert np.array_equal(got, want)
    assert np.array_equal(miss, wmiss)
    if impute:
        assert got.min() >= 0 and got.max() <= (1 if dominance else 2)


def test_oracle_major_genotype_ties_and_all_missing(oracle):
    # fields: 3 -> genotype 0, 2 -> 1, 0 -> 2, 1 -> missing.  Six individuals per SNP.
    def snp(fields):
        b = np.zeros(2, dtype=np.uint8)
        for i, f in enumerate(fields):
            b[i // 4] |= f << (2 * (i % 4))
        return b
    rows = [snp([3, 3, 2, 2, 1, 1]),   # tie 0 vs 1 -> first strict maximum: 0
            snp([2, 2, 0, 0, 1, 1]),   # tie 1 vs 2 -> 1
            snp([0, 0, 0, 2, 1, 1]),   # 2 is major
            snp([1, 1, 1, 1, 1, 1])]   # all missing -> 0
    img = np.concatenate([np.array([0x6C, 0x1B, 0x01], dtype=np.uint8)] + rows)
    got, miss = oracle.read_bed(img, 6, 4)
    assert miss.tolist() == [1, 1, 1, 1]
    assert got[4:, :].tolist() == [[0, 1, 2, 0], [0, 1, 2, 0]]
    # dominance: both homozygotes count as 0 (read_bed.cpp:203-209) -> SNP 2: three 0s, one 1 -> 0
    dom, _ = oracle.read_bed(img, 6, 4, dominance=True)
    assert dom[4:, :].tolist() == [[0, 0, 0, 0], [0, 0, 0, 0]]


@pytest.mark.parametrize("nid,dominance,impute", [(37, False, True), (37, True, True), (130, False, False), (5, False, True)])
def test_device_decoder_byte_code_on_the_host(oracle, nid, dominance, impute):
    """hb_test_bed_decode_snp runs bed_count_byte / bed_major / bed_field / bed_code -- the functions the kernels
    k_bed_info, k_pack_bed, k_bed_to_rows, k_bed_to_colmajor call -- compiled for the host."""
    L = hb.load_library()
    m = 40
    img, _ = make_bed(nid, m, seed=7 + nid, p_missing=0.2)
    want, wmiss = oracle.read_bed(img, nid, m, impute=impute, dominance=dominance)
    bps = (nid + 3) // 4
    rng = np.random.default_rng(1)
    sel = np.ascontiguousarray(rng.permutation(nid)[: max(1, nid // 2)], dtype=np.int32)
    for j in range(m):
        row = np.ascontiguousarray(img[3 + j * bps: 3 + (j + 1) * bps])
        out = np.zeros(nid, dtype=np.int8)
        info = np.zeros(1, dtype=np.uint8)
        _lib.check(L.hb_test_bed_decode_snp(row.ctypes.data, nid, None, nid, int(impute), int(dominance), out.ctypes.data,
                                            info.ctypes.data))
        assert np.array_equal(out, want[:, j]), j
        assert (info[0] >> 7) == wmiss[j]
        # a row selection is the same decode followed by indexing (the major genotype is counted over the whole file)
        out2 = np.zeros(sel.shape[0], dtype=np.int8)
        _lib.check(L.hb_test_bed_decode_snp(row.ctypes.data, nid, sel.ctypes.data, sel.shape[0], int(impute), int(dominance),
                                            out2.ctypes.data, None))
        assert np.array_equal(out2, want[sel, j])


def _demo_slice(n=600, m=240):
    g = load_demo()["geno"]
    return np.asfortranarray(g[:n, :m])


def test_oracle_bigstat_and_dense_ld_against_numpy(oracle):
    X = _demo_slice()
    n, m = X.shape
    st = oracle.bigstat(X)
    Xd = X.astype(np.float64)
    assert np.array_equal(st["sum"], Xd.sum(axis=0))
    assert np.array_equal(st["mean"], Xd.sum(axis=0) / n)
    assert np.allclose(st["xx"], np.sqrt(((Xd - Xd.mean(axis=0)) ** 2).sum(axis=0)), rtol=1e-13)
    ld = oracle.txxmat(X)
    Xc = Xd - Xd.mean(axis=0)
    assert np.allclose(ld, Xc.T @ Xc / n, rtol=1e-9, atol=1e-12)
    assert np.array_equal(ld, ld.T)                               # both triangles get the same value (:168)
    assert np.array_equal(np.diag(ld), st["xx"] * st["xx"] / n)   # :157
    mono = np.where(X.max(axis=0) == X.min(axis=0))[0]
    assert mono.size > 0 and np.all(ld[mono, mono] == 0)          # the demo data has monomorphic SNPs


def test_oracle_sparse_and_chromosome_branches(oracle):
    X = _demo_slice(600, 200)
    n, m = X.shape
    full = oracle.txxmat(X)
    st = oracle.bigstat(X)
    chisq = 3.84
    sp = oracle.txxmat(X, chisq=chisq)
    with np.errstate(divide="ignore", invalid="ignore"):
        r = full * n / np.outer(st["xx"], st["xx"])   # full = p12 / n
        keep = ~(r * r * n <= chisq)
    off = ~np.eye(m, dtype=bool)
    # away from the diagonal the kept entries carry the dense values and the dropped ones are 0
    assert np.array_equal(sp[off & keep], full[off & keep])
    assert np.all(sp[off & ~keep] == 0)
    assert 0 < (sp != 0).sum() < m * m
    # the sparse branch recomputes the diagonal from the centring formula (:133-141): close, not identical
    dg = np.diag(sp)
    poly = st["xx"] > 0
    assert np.allclose(dg[poly], np.diag(full)[poly], rtol=1e-9)
    # chromosomes: nothing across, the Geno values within
    chr_ = np.repeat([1, 2, 3, 4], m // 4).astype(np.int32)
    same = chr_[:, None] == chr_[None, :]
    cd = oracle.txxmat(X, chr=chr_)
    assert np.array_equal(cd[same], full[same]) and np.all(cd[~same] == 0)
    cs = oracle.txxmat(X, chr=chr_, chisq=0.0)   # tXXmat_Chr takes the sparse branch for chisq = 0 (:520-523)
    assert np.all(cs[~same] == 0)
    assert np.array_equal(cs[same & off & (full != 0)], full[same & off & (full != 0)])


def test_ldmat_branch_logic_follows_the_reference():
    m = 12
    two = [1] * 6 + [2] * 6
    assert hb.ldmat_plan(m) == ("geno_dense", None)                        # R/ldm.r:79-84
    assert hb.ldmat_plan(m, chisq=0) == ("geno_dense", None)               # :80-82
    assert hb.ldmat_plan(m, chisq=-1) == ("geno_dense", None)              # :45-47
    assert hb.ldmat_plan(m, chisq=3.84) == ("geno_sparse", 3.84)
    assert hb.ldmat_plan(m, map_chr=[7] * m, chisq=0) == ("geno_dense", None)      # :52-55
    assert hb.ldmat_plan(m, map_chr=[7] * m, chisq=5.0) == ("geno_sparse", 5.0)
    assert hb.ldmat_plan(m, map_chr=two) == ("chr_dense", None)            # :90-91
    assert hb.ldmat_plan(m, map_chr=two, chisq=0) == ("chr_sparse", 0)     # tXXmat.cpp:520-523
    assert hb.ldmat_plan(m, map_chr=two, ldchr=True, chisq=2.0) == ("geno_sparse", 2.0)
    with pytest.raises(RuntimeError, match="0 is not allowed in chromosome"):
        hb.ldmat_plan(m, map_chr=[0] + [1] * (m - 1))


def test_bed_and_ldmat_entry_points_refuse_without_a_gpu():
    if hb.device_count() > 0:
        pytest.skip("a GPU is present")
    img, _ = make_bed(8, 4, seed=1)
    with pytest.raises(RuntimeError, match="no CPU path|CUDA"):
        hb.read_bed(img, 8, 4)
    with pytest.raises(RuntimeError, match="no CPU path|CUDA"):
        hb.LdMat(np.zeros((8, 4), dtype=np.int8))
    # argument checks come before the device: wrong magic, short image
    bad = img.copy()
    bad[0] = 0
    with pytest.raises(RuntimeError, match="magic"):
        hb.read_bed(bad, 8, 4)
    with pytest.raises(RuntimeError, match="needed for"):
        hb.read_bed(img[:-1], 8, 4)


@pytest.mark.parametrize("mode", ["dense", "sparse", "chr_dense", "chr_sparse0", "chr_sparse"])
def test_device_ld_epilogue_on_the_host_is_bit_identical_to_the_oracle(oracle, mode):
    """hb_test_ld_entries runs ld_entry() -- what k_ld_panel applies to every accumulator element -- on the host,
    fed with the exact Gram matrix: every entry must equal the oracle's bit for bit (operation order of
    tXXmat.cpp:141-148, smaller index in the role of the outer loop variable, diagonal rules of both branches)."""
    L = hb.load_library()
    X = _demo_slice(600, 160)          # includes monomorphic SNPs (0/0 -> NaN in r)
    n, m = X.shape
    st = oracle.bigstat(X)
    G = np.asfortranarray((X.astype(np.int64).T @ X.astype(np.int64)).astype(np.int32))
    chr_ = (np.arange(m) * 3 // m + 1).astype(np.int32)
    c, q = {"dense": (None, None), "sparse": (None, 3.84), "chr_dense": (chr_, None), "chr_sparse0": (chr_, 0.0),
            "chr_sparse": (chr_, 10.0)}[mode]
    out = np.zeros((m, m), order="F")
    _lib.check(L.hb_test_ld_entries(n, m, G.ctypes.data, st["sum"].ctypes.data, st["mean"].ctypes.data, st["xx"].ctypes.data,
                                    None if c is None else c.ctypes.data, int(q is not None), 0.0 if q is None else q,
                                    out.ctypes.data))
    want = oracle.txxmat(X, chr=c, chisq=q)
    assert np.array_equal(out, want)
    assert (out != 0).sum() > m


@pytest.mark.parametrize("n,m", [(600, 240), (130, 70), (1001, 33)])
def test_device_bigstat_code_on_the_host_is_bit_identical_to_the_oracle(oracle, n, m):
    """hb_test_ld_stats runs ld_stats_row() -- the body of k_ld_stats -- on the host, on the device layout (one
    zero-padded row of Kpad bytes per SNP)."""
    L = hb.load_library()
    from tests.util_demo import synth
    X = _demo_slice(n, m) if n == 600 else synth(n, m, seed=n)[1]
    Kpad = (n + 127) // 128 * 128
    raw = np.zeros(m * Kpad + 16, dtype=np.int8)
    off = (-raw.ctypes.data) % 16
    Xc = raw[off:off + m * Kpad].reshape(m, Kpad)
    Xc[:, :n] = X.T
    sm, mean, xx = np.zeros(m), np.zeros(m), np.zeros(m)
    _lib.check(L.hb_test_ld_stats(Xc.ctypes.data, Kpad, n, m, sm.ctypes.data, mean.ctypes.data, xx.ctypes.data))
    st = oracle.bigstat(X)
    assert np.array_equal(sm, st["sum"]) and np.array_equal(mean, st["mean"]) and np.array_equal(xx, st["xx"])


@pytest.mark.parametrize("model,Pi,fold", [("BayesCpi", [0.95, 0.05], None),
                                            ("BayesR", [0.95, 0.02, 0.02, 0.01], [0, 1e-4, 1e-3, 1e-2])])
def test_oracle_pipeline_ldmat_into_sbayesd_reproduces_the_golden_pin(oracle, model, Pi, fold):
    """ldmat() -> sbrm() as a user chains them (R/ldm.r -> R/sbayes.r:198-215): the LD oracle's matrix of the bundled
    genotypes, fed to the SBayesD oracle with the reference's COJO file, lands on the committed pin (which was made
    with a numpy X'X/n): same inclusion indicators, variance components to 1e-9."""
    d = load_demo()
    ld = oracle.txxmat(np.asfortranarray(d["geno"]))
    ss = np.asfortranarray(np.column_stack([d["ma_maf"], d["ma_beta"], d["ma_se"], d["ma_n"]]))
    r = oracle.sbayesd(ss, ld, model, Pi, fold=fold, niter=100, nburn=50, thin=5, seed=666666)
    g = np.load(os.path.join(GOLDEN, "demo_oracle_sbayesd_%s.npz" % model))
    assert np.array_equal(r["diag"]["tracker"], g["tracker"])
    assert np.array_equal(r["diag"]["nnz_trace"], g["nnz_trace"])
    for k in ("Vg", "Ve", "h2"):
        assert abs(r[k] / float(g[k]) - 1) < 1e-9, k
    assert np.allclose(r["alpha"], g["alpha"], rtol=1e-7, atol=1e-12)


def test_sbrm_argument_handling_follows_the_reference():
    import scipy.sparse as sp
    m = 6
    cojo = np.arange(m * 8, dtype=float).reshape(m, 8)
    dense = np.eye(m)
    a = hb.sbrm_plan(cojo, dense)
    assert not a["sparse"] and a["model"] == "BayesB" and a["Pi"] == [0.95, 0.05] and a["fold"] is None
    assert (a["niter"], a["nburn"]) == (20000, 12000)                       # R/sbayes.r:186-191
    assert np.array_equal(a["sumstat"], cojo[:, [3, 4, 5, 7]])              # :205
    b = hb.sbrm_plan(cojo, sp.csc_matrix(dense), method="BayesR")
    assert b["sparse"] and b["Pi"] == [0.95, 0.02, 0.02, 0.01] and b["fold"] == [0, 0.0001, 0.001, 0.01]
    assert (b["niter"], b["nburn"]) == (50000, 30000)
    with pytest.raises(RuntimeError, match="Unrecognized type of ldm"):
        hb.sbrm_plan(cojo, [[1.0]])
    with pytest.raises(RuntimeError, match="bad setting for collecting frequency"):
        hb.sbrm_plan(cojo, dense, niter=100, nburn=96, thin=5)
    with pytest.raises(RuntimeError, match="can not implement GWAS analysis for the method: BayesA"):
        hb.sbrm_plan(cojo, dense, method="BayesA", windindx=np.ones(m, dtype=np.int32))
    with pytest.raises(NotImplementedError):
        hb.sbrm_plan(cojo, dense, method="CG")
