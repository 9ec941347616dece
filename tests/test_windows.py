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
