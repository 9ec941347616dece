"""GPU parity tests of the two scalar sides of the sweep -- the ring of workers (default) and the serial CTA + helpers
(csrc/hb_serial.cuh, HB_SERIAL=1; tiles of 256 SNPs) -- against the CPU oracle and against each other: every lag, every
mixture model, more candidates than a package holds, the flip-heavy regime where classes have to be re-decided tile
after tile, and the metric's number of rows (n = 50 000, the production tiling)."""
import os

import numpy as np
import pytest

import hibayes_b200 as hb
from tests.test_gpu_parity import _compare
from tests.util_demo import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _serial_mode():
    # the engine reads HB_SERIAL when it is created: every hb.Bayes() of this file runs the serial-CTA scalar side
    # unless a test switches it off for a comparison
    old = os.environ.get("HB_SERIAL")
    os.environ["HB_SERIAL"] = "1"
    yield
    if old is None:
        os.environ.pop("HB_SERIAL", None)
    else:
        os.environ["HB_SERIAL"] = old


PI_R = [0.95, 0.02, 0.02, 0.01]
FOLD_R = [0, 1e-4, 1e-3, 1e-2]


class _env:
    def __init__(self, **kw):
        self.kw, self.old = kw, {}

    def __enter__(self):
        for k, v in self.kw.items():
            self.old[k] = os.environ.get(k)
            os.environ[k] = str(v)

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("lag", [1, 2, 3, 4, 5, 8])
def test_serial_cta_matches_the_oracle_for_every_lag(oracle, lag):
    # 9 tiles of 256 SNPs (the last one ragged), every lag the far / near / local correction paths distinguish
    y, X = synth(3000, 2200, seed=101, n_causal=30)
    kw = dict(niter=10, nburn=4, thin=2, seed=4711)
    ref = oracle.bayes(y, X, "BayesR", [0.9, 0.05, 0.03, 0.02], fold=FOLD_R, **kw)
    got = hb.Bayes(y, X, "BayesR", [0.9, 0.05, 0.03, 0.02], fold=FOLD_R, tile_snps=256, lag_tiles=lag, **kw)
    _compare(got, ref)


@pytest.mark.parametrize("model,Pi,fold", [("BayesCpi", [0.95, 0.05], None), ("BayesB", [0.9, 0.1], None),
                                           ("BayesBpi", [0.95, 0.05], None), ("BayesC", [0.8, 0.2], None),
                                           ("BayesR", PI_R, FOLD_R)])
def test_serial_cta_mixture_models(oracle, model, Pi, fold):
    y, X = synth(4000, 4096, seed=7, n_causal=40)
    kw = dict(niter=8, nburn=2, thin=2, seed=99)
    ref = oracle.bayes(y, X, model, Pi, fold=fold, **kw)
    got = hb.Bayes(y, X, model, Pi, fold=fold, **kw)
    _compare(got, ref)


def test_serial_cta_and_ring_of_workers_agree():
    y, X = synth(3000, 4096, seed=5, n_causal=40)
    kw = dict(niter=8, nburn=2, thin=2, seed=2024)
    a = hb.Bayes(y, X, "BayesR", PI_R, fold=FOLD_R, **kw)
    with _env(HB_SERIAL=0):
        b = hb.Bayes(y, X, "BayesR", PI_R, fold=FOLD_R, **kw)
    # same classes; the effects agree to rounding (a repaired tile ends with the step-by-step chain, an unrepaired one with
    # the solved chain matrix, and the two modes do not repair the same tiles)
    assert np.array_equal(a["diag"]["tracker"], b["diag"]["tracker"])
    assert np.allclose(a["alpha"], b["alpha"], rtol=1e-9, atol=1e-14) and np.allclose(a["g"], b["g"], rtol=1e-9, atol=1e-12)
    assert abs(a["Ve"] / b["Ve"] - 1) < 1e-10


def test_many_candidates_per_tile(oracle):
    # Pi puts 40 % of the SNPs in the model: ~100 candidates per tile, more than a package holds -> the generic path
    y, X = synth(2500, 1536, seed=12, n_causal=200)
    kw = dict(niter=6, nburn=2, thin=2, seed=5)
    ref = oracle.bayes(y, X, "BayesCpi", [0.6, 0.4], **kw)
    got = hb.Bayes(y, X, "BayesCpi", [0.6, 0.4], **kw)
    _compare(got, ref)


@pytest.mark.parametrize("mode", ["ring", "serial"])
def test_flip_heavy_chain_against_the_oracle(oracle, mode):
    """Large effect variances (folds x 64) and h2 = 0.9 at n = 20 000: the changes inside a tile move the other SNPs of the
    tile across their class boundaries, so the speculation misses and tiles need repair rounds (rounds > tiles)."""
    y, X = synth(20000, 4096, seed=33, n_causal=400, h2=0.9)
    fold = [0.0] + [64 * f for f in FOLD_R[1:]]
    kw = dict(niter=30, nburn=10, thin=5, seed=808)
    ref = oracle.bayes(y, X, "BayesR", PI_R, fold=fold, **kw)
    with _env(HB_SERIAL=int(mode == "serial")):
        got = hb.Bayes(y, X, "BayesR", PI_R, fold=fold, **kw)
    _compare(got, ref)
    assert got["diag"]["rounds_total"] > got["diag"]["tiles_total"], (got["diag"]["rounds_total"], got["diag"]["tiles_total"])
    if mode == "ring":
        return
    with _env(HB_SERIAL=0):
        ring = hb.Bayes(y, X, "BayesR", PI_R, fold=fold, **kw)
    # (not bit for bit here: a repaired tile ends with the step-by-step chain, an unrepaired one with the solved chain
    # matrix, and the two modes do not repair the same tiles; the classes are the same, the effects agree to rounding)
    assert np.array_equal(ring["diag"]["tracker"], got["diag"]["tracker"])
    assert np.allclose(ring["alpha"], got["alpha"], rtol=1e-9, atol=1e-14)


@pytest.mark.parametrize("mode", ["ring", "serial"])
def test_metric_rows_against_the_oracle(oracle, mode):
    """The production configuration of the metric (n = 50 000 rows: 131 slabs of 384 rows, tiles of 256, default lag)
    on m = 65 536 SNPs for three sweeps: classes bit-exact, effects to 1e-5."""
    n, m = 50000, 65536
    X = hb.synth_geno_host(n, m, seed=20260101)
    rng = np.random.default_rng(3)
    idx = rng.choice(m, 60, replace=False)
    gv = X[:, idx].astype(np.float64) @ rng.standard_normal(60)
    y = gv * np.sqrt(0.5 / gv.var()) + rng.normal(scale=np.sqrt(0.5), size=n)
    kw = dict(niter=3, nburn=1, thin=1, seed=20260101)
    ref = oracle.bayes(y, X, "BayesR", PI_R, fold=FOLD_R, **kw)
    with _env(HB_SERIAL=int(mode == "serial")):
        got = hb.Bayes(y, X, "BayesR", PI_R, fold=FOLD_R, **kw)
    _compare(got, ref)
    assert got["diag"]["tiles_total"] == 3 * 256


@pytest.mark.parametrize("model,Pi,fold", [("BayesR", PI_R, FOLD_R), ("BayesCpi", [0.9, 0.1], None), ("BayesRR", [0.0, 1.0], None)])
def test_result_does_not_depend_on_the_row_slot_path(model, Pi, fold):
    """A tile whose candidates do not fit the row slots reads the Gram band from global memory instead of shared memory.
    Which path a tile takes depends on the row set, i.e. on timing -- so both must give the same bits, or the ranks of a
    row-sharded run drift apart (seen on 8 GPUs in the flip-heavy regime before the two paths summed in the same order)."""
    y, X = synth(6000, 3072, seed=17, n_causal=120, h2=0.8)
    fold_ = None if fold is None else [0.0] + [16 * f for f in fold[1:]]
    kw = dict(niter=10, nburn=2, thin=2, seed=606)
    with _env(HB_SERIAL=0):
        a = hb.Bayes(y, X, model, Pi, fold=fold_, **kw)
        with _env(HB_KROW=3):
            b = hb.Bayes(y, X, model, Pi, fold=fold_, **kw)
        with _env(HB_KROW=12):
            c = hb.Bayes(y, X, model, Pi, fold=fold_, **kw)
    for other in (b, c):
        assert np.array_equal(a["diag"]["tracker"], other["diag"]["tracker"])
        assert np.array_equal(a["alpha"], other["alpha"]) and np.array_equal(a["g"], other["g"]) and a["Ve"] == other["Ve"]
