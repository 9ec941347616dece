"""The Python wrappers of the entry points that have not run on hardware yet (LdMat, read_bed, BedGeno in Bayes()/Engine,
ibrm, sbrm, predict_samples) against a recording stand-in for the shared library: argument order, shapes and dtypes of
what crosses the C ABI.  No computation is faked into results -- the stand-in returns 0 and leaves outputs as they are."""
import ctypes as C

import numpy as np
import pytest

import hibayes_b200 as hb
from hibayes_b200 import _lib
from tests.util_bed import make_bed


class Recorder:
    def __init__(self):
        self.calls = []

    def __getattr__(self, name):
        if not name.startswith("hb_"):
            raise AttributeError(name)

        def fn(*args):
            self.calls.append((name, args))
            if name in ("hb_engine_create", "hb_ldmat_create"):
                args[-1]._obj.value = 0xBEEF          # a non-null handle
            if name == "hb_bayes" and args[0]._obj.x_type == 2:
                # the hb_bed_source lives for the duration of the call only: look at it now
                src = C.cast(args[0]._obj.X, C.POINTER(_lib.BedSource)).contents
                self.bed_seen = (src.nid, src.len, src.impt, src.dominance, bool(src.rows))
            if name == "hb_engine_describe":
                for a, v in zip(args[1:5], (8, 384, 256, 5)):
                    a._obj.value = v
            return 0
        return fn

    def names(self):
        return [c[0] for c in self.calls]


@pytest.fixture
def rec(monkeypatch):
    r = Recorder()
    monkeypatch.setattr(_lib, "_LIB", r)
    return r


def test_ldmat_wrapper_calls(rec):
    X = np.asfortranarray(np.random.default_rng(0).integers(0, 3, size=(50, 20)).astype(np.int8))
    h = hb.LdMat(X, panel_cols=64)
    assert rec.names()[:3] == ["hb_ldmat_create", "hb_ldmat_set_panel_cols", "hb_ldmat_load_i8"]
    assert rec.calls[0][1][1:3] == (50, 20) and rec.calls[2][1][2] == 50
    st = h.stats()
    assert set(st) == {"mean", "sum", "xx"} and st["xx"].shape == (20,)
    d = h.dense(chr=np.ones(20), chisq=3.84)
    name, args = rec.calls[-1]
    assert name == "hb_ldmat_dense" and args[2] == 1 and args[3] == 3.84 and args[5] == 20 and d.shape == (20, 20)
    s = h.sparse()
    assert [c[0] for c in rec.calls[-2:]] == ["hb_ldmat_sparse", "hb_ldmat_sparse_get"] and s.shape == (20, 20) and s.nnz == 0
    assert rec.calls[-2][1][1] is None and rec.calls[-2][1][2] == 0
    h.close()
    assert rec.names()[-1] == "hb_ldmat_destroy"
    with pytest.raises(TypeError):
        hb.LdMat(X.astype(np.float64))


def test_front_ends_route_to_the_right_kernels(rec):
    X = np.asfortranarray(np.random.default_rng(1).integers(0, 3, size=(40, 12)).astype(np.int8))
    assert isinstance(hb.ldmat(X), np.ndarray) and "hb_ldmat_dense" in rec.names()
    rec.calls.clear()
    hb.ldmat(X, chisq=5.0)
    assert "hb_ldmat_sparse" in rec.names() and "hb_ldmat_dense" not in rec.names()
    rec.calls.clear()
    hb.ldmat(X, map_chr=["1"] * 6 + ["X"] * 6)
    call = [c for c in rec.calls if c[0] == "hb_ldmat_sparse"][0]
    assert call[1][1] is not None and call[1][2] == 0          # chromosome codes, no threshold: tXXmat_Chr dense branch


def test_bed_inputs_cross_the_abi_as_declared(rec):
    img, _ = make_bed(10, 6, seed=3)
    out, miss = hb.read_bed(img, 10, 6, impute=False, mode="D")
    name, a = rec.calls[-1]
    assert name == "hb_bed_decode" and a[0] == 0 and a[2] == img.shape[0] and a[3:7] == (10, 6, 0, 1)
    assert out.shape == (10, 6) and out.dtype == np.int8 and miss.shape == (6,)
    g = hb.BedGeno(img, 10, 6, rows=[9, 0, 4])
    assert g.shape == (3, 6)
    e = hb.Engine(3, 6)
    e.load_geno(g)
    name, a = rec.calls[-1]
    assert name == "hb_engine_load_bed" and a[2] == img.shape[0] and a[3] == 10 and a[4] == g.rows.ctypes.data and a[5:] == (1, 0)
    A = np.zeros((6, 4))
    assert e.predict_samples(A).shape == (3, 4)
    name, a = rec.calls[-1]
    assert name == "hb_engine_predict_samples" and a[2] == 6 and a[3] == 4 and a[5] == 3
    with pytest.raises(ValueError):
        e.predict_samples(np.zeros((5, 4)))
    # Bayes() hands the .bed source over as x_type 2 with a pointer to an hb_bed_source
    y = np.arange(3, dtype=np.float64)
    hb.Bayes(y, g, "BayesCpi", [0.95, 0.05], niter=10, nburn=5)
    name, a = rec.calls[-1]
    assert name == "hb_bayes"
    args = a[0]._obj
    assert args.x_type == 2 and args.n == 3 and args.m == 6
    assert rec.bed_seen == (10, img.shape[0], 1, 0, True)
    with pytest.raises(RuntimeError, match="Number of individuals not equals"):
        hb.Bayes(np.arange(4, dtype=np.float64), g, "BayesCpi", [0.95, 0.05], niter=10, nburn=5)


def test_ibrm_and_sbrm_wrappers(rec):
    import scipy.sparse as sp
    X = np.asfortranarray(np.random.default_rng(2).integers(0, 3, size=(30, 8)).astype(np.int8))
    y = np.arange(30, dtype=np.float64)
    y[[3, 7]] = np.nan
    r = hb.ibrm(y, X, method="BayesR", niter=20, nburn=10, thin=2)
    names = rec.names()
    assert "hb_bayes" in names and names[-3:] == ["hb_engine_load_geno_i8", "hb_engine_predict", "hb_engine_destroy"]
    b = [c for c in rec.calls if c[0] == "hb_bayes"][0][1][0]._obj
    assert b.n == 28 and b.m == 8 and b.n_fold == 4 and b.niter == 20
    assert r["g"].shape == (30,)
    rec.calls.clear()
    cojo = np.ones((8, 8))
    hb.sbrm(cojo, np.eye(8), niter=20, nburn=10)
    assert rec.names() == ["hb_sbayesd"]
    rec.calls.clear()
    hb.sbrm(cojo, sp.csc_matrix(np.eye(8)), niter=20, nburn=10)
    assert rec.names() == ["hb_sbayess"]
