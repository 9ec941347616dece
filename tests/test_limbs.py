"""CPU test of csrc/hb_limbs.h, the fixed-point limb dot prepared for the sweep's inner loop:
the six-limb split is lossless for |q| < 2^47, the dp4a sums and their merge reproduce x'q exactly (Python integers),
and the only error against the fp64 dot is the quantisation bound sum(x) / (2 scale)."""
import ctypes as C

import numpy as np
import pytest

import hibayes_b200 as hb
from hibayes_b200 import _lib


def _dot(x, r, scale):
    L = hb.load_library()
    dq, ok = C.c_longlong(0), C.c_int(0)
    x = np.ascontiguousarray(x, dtype=np.uint8)
    r = np.ascontiguousarray(r, dtype=np.float64)
    _lib.check(L.hb_test_limb_dot(x.ctypes.data, r.ctypes.data, x.shape[0], scale, C.byref(dq), C.byref(ok)))
    return dq.value, ok.value


@pytest.mark.parametrize("n,seed", [(384, 1), (24, 2), (4, 3), (50016, 4)])
def test_limb_dot_is_exact_in_fixed_point(n, seed):
    rng = np.random.default_rng(seed)
    x = rng.integers(0, 3, size=n).astype(np.uint8)
    r = np.clip(rng.normal(scale=3.0, size=n), -7.99, 7.99)
    r[: min(n, 4)] = [7.99, -7.99, 0.0, -1e-15][: min(n, 4)]
    scale = 2.0 ** 44                          # |r| < 8  ->  |q| < 2^47
    got, ok = _dot(x, r, scale)
    assert ok == 1
    q = [int(np.rint(v * scale)) for v in r]   # Python integers: exact
    assert got == sum(int(a) * b for a, b in zip(x, q))
    exact = float(np.dot(x.astype(np.float64), r))
    assert abs(got / scale - exact) <= x.sum() / (2 * scale) + 1e-12 * abs(exact)


def test_limb_split_limits():
    x = np.array([2, 2, 2, 2], dtype=np.uint8)
    lim = 2.0 ** 47
    got, ok = _dot(x, np.array([lim - 1, -(lim - 1), 1.0, -1.0]), 1.0)
    assert ok == 1 and got == 0
    got, ok = _dot(x, np.array([-(lim - 1), -(lim - 1), -(lim - 1), -(lim - 1)]), 1.0)
    assert ok == 1 and got == -8 * (int(lim) - 1)
    _, ok = _dot(x, np.array([lim, 0.0, 0.0, 0.0]), 1.0)
    assert ok == 0
    _, ok = _dot(x, np.array([np.nan, 0.0, 0.0, 0.0]), 1.0)
    assert ok == 0
