"""CPU tests of the window cutters (cutwind.cpp:13-65) against an independent numpy restatement."""
import numpy as np
import pytest

import hibayes_b200 as hb
from tests.util_demo import load_demo


def np_by_bp(chr_, pos, bp):
    out = np.zeros(len(chr_), dtype=np.int32)
    count = 1
    for c in np.unique(chr_):
        idx = np.flatnonzero(chr_ == c)
        bp0 = 1.0
        while bp0 <= pos[idx].max():
            sel = idx[(pos[idx] >= bp0) & (pos[idx] < bp0 + bp)]
            if sel.size:
                out[sel] = count
                count += 1
            bp0 += bp
    return out


def np_by_num(chr_, pos, k):
    out = np.zeros(len(chr_), dtype=np.int32)
    count = 1
    for c in np.unique(chr_):
        idx = np.flatnonzero(chr_ == c)
        if idx.size <= k:
            out[idx] = count
            count += 1
            continue
        o = idx[np.argsort(pos[idx], kind="stable")]
        for st in range(0, idx.size, k):
            out[o[st:st + k]] = count
            count += 1
    return out


def test_windows_on_the_demo_map():
    d = load_demo()
    chr_ = np.array([float(c) for c in d["chr"]])
    pos = d["pos"].astype(np.float64)
    for bp in (1e6, 2.5e6, 1e9):
        if pos.max() >= bp:
            w = hb.cutwind(chr_, pos, windsize=bp)
            assert np.array_equal(w, np_by_bp(chr_, pos, bp)) and w.min() >= 1
    for k in (1, 7, 50, 1000):
        w = hb.cutwind(chr_, pos, windnum=k)
        assert np.array_equal(w, np_by_num(chr_, pos, k))
        assert w.max() == len(np.unique(w)) and w.min() == 1


def test_windows_unsorted_chromosomes_and_errors():
    rng = np.random.default_rng(2)
    chr_ = rng.integers(1, 5, size=500).astype(np.float64)
    pos = rng.integers(1, 10_000, size=500).astype(np.float64)
    assert np.array_equal(hb.cutwind(chr_, pos, windsize=777.0), np_by_bp(chr_, pos, 777.0))
    assert np.array_equal(hb.cutwind(chr_, pos, windnum=13), np_by_num(chr_, pos, 13))
    with pytest.raises(RuntimeError, match="larger than the total number of markers"):
        hb.cutwind(chr_, pos, windnum=501)
    with pytest.raises(RuntimeError, match="smaller than wind size"):
        hb.cutwind(chr_, pos, windsize=1e6)


def test_ibrm_argument_handling_follows_the_reference():
    a = hb.ibrm_plan()
    assert a["model"] == "BayesCpi" and a["Pi"] == [0.95, 0.05] and (a["niter"], a["nburn"]) == (20000, 12000) and a["windindx"] is None
    b = hb.ibrm_plan("BayesR")
    assert b["Pi"] == [0.95, 0.02, 0.02, 0.01] and b["fold"] == [0, 0.0001, 0.001, 0.01] and (b["niter"], b["nburn"]) == (50000, 30000)
    d = load_demo()
    c = hb.ibrm_plan("BayesCpi", windnum=50, map_chr=d["chr"], map_pos=d["pos"])
    chr_ = np.array([float(v) for v in d["chr"]])
    assert np.array_equal(c["windindx"], np_by_num(chr_, d["pos"].astype(np.float64), 50))
    # a chromosome name that is not a number follows the largest number (R/bayes.r:234-243)
    names = ["1"] * 5 + ["X"] * 5 + ["2"] * 5
    e = hb.ibrm_plan("BayesC", windnum=5, map_chr=names, map_pos=np.arange(1, 16))
    assert e["windindx"].tolist() == [1] * 5 + [3] * 5 + [2] * 5
    with pytest.raises(RuntimeError, match="can not implement GWAS analysis"):
        hb.ibrm_plan("BayesL", windnum=5, map_chr=names, map_pos=np.arange(1, 16))
    with pytest.raises(RuntimeError, match="map information must be provided"):
        hb.ibrm_plan("BayesC", windnum=5)
    with pytest.raises(RuntimeError, match="bad setting for collecting frequency"):
        hb.ibrm_plan(niter=10, nburn=6)
    with pytest.raises(NotImplementedError):
        hb.ibrm_plan("BSLMM")
