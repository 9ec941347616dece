"""CPU-side checks of the product library: it loads, exports every symbol include/hibayes_b200.h
declares, refuses to run without a GPU (no CPU fallback), and the host driver reproduces the
reference's argument checks (Bayes.cpp:92-117, :293, :325, :356) before touching the device."""
import os
import re

import numpy as np
import pytest

import hibayes_b200 as hb
from hibayes_b200 import _lib
from tests.util_demo import load_demo_T1

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _no_gpu():
    return hb.device_count() <= 0


def test_library_exports_every_declared_symbol():
    L = hb.load_library()
    header = open(os.path.join(ROOT, "include", "hibayes_b200.h")).read()
    declared = set(re.findall(r"\b(hb_[A-Za-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(L, name), "libhibayes_b200.so does not export %s" % name
    assert declared == set(_lib.SYMBOLS)


def test_no_cpu_fallback():
    if not _no_gpu():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        hb.Engine(100, 100)
    y, X = load_demo_T1()
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        hb.Bayes(y, X, "BayesCpi", [0.95, 0.05], niter=10, nburn=5)


def test_product_does_not_reference_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "hibayes_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "hb_oracle" not in src and "oracle/" not in src, f


def test_host_driver_argument_checks_match_reference_messages():
    y, X = load_demo_T1()
    cases = [
        (dict(model="BayesCpi", Pi=[0.9, 0.05]), "sum of Pi should be 1"),
        (dict(model="BayesCpi", Pi=[1.0, 0.0]), "all markers have no effect size"),
        (dict(model="BayesCpi", Pi=[1.5, -0.5]), "elements of Pi should be at the range"),
        (dict(model="BayesR", Pi=[0.95, 0.02, 0.02, 0.01]), "'fold' should be provided"),
        (dict(model="BayesCpi", Pi=[0.9, 0.05, 0.05], fold=[0, 1, 2]), "length of Pi should be 2"),
        (dict(model="BayesCpi", Pi=[0.95, 0.05], dfvg=2.0), "dfvg should not be less than 2"),
        (dict(model="BayesCpi", Pi=[0.95, 0.05], niter=5, nburn=10), "shold be larger than burn-in"),
    ]
    for kw, msg in cases:
        kw.setdefault("niter", 20)
        kw.setdefault("nburn", 10)
        with pytest.raises(RuntimeError, match=msg):
            hb.Bayes(y, X, **kw)
    yy = y.copy()
    yy[0] = np.nan
    with pytest.raises(RuntimeError, match="NAs are not allowed in y"):
        hb.Bayes(yy, X, "BayesCpi", [0.95, 0.05], niter=20, nburn=10)
    with pytest.raises(RuntimeError, match="Number of individuals not equals"):
        hb.Bayes(y[:-1], X, "BayesCpi", [0.95, 0.05], niter=20, nburn=10)


def test_synthetic_genotypes_host_generator():
    X = hb.synth_geno_host(1000, 64, seed=20260101)
    assert X.dtype == np.int8 and X.min() >= 0 and X.max() <= 2
    p = X.mean(axis=0) / 2
    assert 0.02 < p.min() and p.max() < 0.56
    # rows are addressed globally: a shard starting at row 400 is the same data
    Y = hb.synth_geno_host(600, 64, seed=20260101, row_offset=400)
    assert np.array_equal(Y, X[400:, :])
    Z = hb.synth_geno_host(1000, 64, seed=7)
    assert not np.array_equal(Z, X)
