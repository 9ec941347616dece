"""The class decisions of the sweep: certified thresholds on rhs^2 (hb_sweep.cuh, solve_thresholds / thr_class)
against the exact inverse-CDF draw (class_cum / class_from_cum, Bayes.cpp:759-781).  Host build of the device code."""
import ctypes as C

import numpy as np

import hibayes_b200 as hb


def _params(rng, F, xx, vare, varg, fold, pi):
    a, c = np.zeros(F - 1), np.zeros(F - 1)
    for k in range(1, F):
        vf = varg * fold[k]
        v = xx + vare / vf
        a[k - 1] = -0.5 * np.log(vf * (xx / vare) + 1.0) + np.log(pi[k])
        c[k - 1] = 0.5 / (vare * v)
    return a, c, float(np.log(pi[0]))


def _check(L, F, u, a, c, logpi0, rrs):
    TL, TH = np.zeros(F - 1), np.zeros(F - 1)
    assert L.hb_test_class_thresholds(F, C.c_double(u), a.ctypes.data, c.ctypes.data, C.c_double(logpi0), TL.ctypes.data, TH.ctypes.data) == 0
    by, ex = C.c_int(), C.c_int()
    undecided = 0
    for rr in rrs:
        assert L.hb_test_class_of(F, C.c_double(rr), C.c_double(u), a.ctypes.data, c.ctypes.data, C.c_double(logpi0), TL.ctypes.data,
                                  TH.ctypes.data, C.byref(by), C.byref(ex)) == 0
        if by.value < 0:
            undecided += 1
        else:
            assert by.value == ex.value, (F, u, rr, by.value, ex.value, TL, TH)
    return TL, TH, undecided


def test_threshold_classes_equal_the_exact_draw():
    L = hb.load_library()
    for f in ("hb_test_class_thresholds", "hb_test_class_of"):
        getattr(L, f).restype = C.c_int
    L.hb_test_class_thresholds.argtypes = [C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
    L.hb_test_class_of.argtypes = [C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p,
                                   C.POINTER(C.c_int), C.POINTER(C.c_int)]
    rng = np.random.default_rng(7)
    total = und = certified = brackets = 0
    for trial in range(400):
        F = 4 if trial % 3 else 2
        xx = rng.uniform(200, 60000)
        vare = rng.uniform(0.2, 3.0)
        varg = 10 ** rng.uniform(-7, -3)
        fold = [0, 1e-4, 1e-3, 1e-2] if F == 4 else [0, 1.0]
        pi = rng.dirichlet(np.ones(F) * 0.7) * 0.98 + 0.02 / F
        a, c, logpi0 = _params(rng, F, xx, vare, varg, fold, pi)
        u = rng.uniform() if trial % 5 else rng.choice([1e-9, 1e-4, 0.5, 1 - 1e-4, 1 - 1e-9])
        sd = np.sqrt(xx * vare)
        rrs = np.concatenate([[0.0], (rng.normal(size=60) * sd * rng.choice([0.3, 1, 3, 30])) ** 2])
        TL, TH, n_und = _check(L, F, u, a, c, logpi0, rrs)
        # right at the brackets: just outside must agree with the exact draw, inside is left to the exact evaluation
        edge = []
        for b in range(F - 1):
            if np.isfinite(TH[b]) and TH[b] > 0:
                edge += [TH[b], TH[b] * (1 + 1e-12)]
            if TL[b] > 0:
                edge += [TL[b], TL[b] * (1 - 1e-12), 0.5 * (TL[b] + TH[b])]
            if trial % 5:
                certified += int(np.isfinite(TH[b]))
                brackets += 1
        _, _, n_und2 = _check(L, F, u, a, c, logpi0, np.array(edge)) if edge else (0, 0, 0)
        if trial % 5:   # (the extreme uniforms of every fifth trial are allowed to stay uncertified: absolute margin)
            total += len(rrs)
            und += n_und
    assert und <= 0.001 * total             # random right-hand sides almost never fall inside a bracket
    assert certified >= 0.999 * brackets    # and practically every boundary gets certified


def test_unordered_variances_are_left_to_the_exact_path():
    """thr_class is only valid for class-ordered slopes; the engine switches thresholds off otherwise (engine.cu),
    here: with ordered slopes the decision is monotone in rhs^2."""
    L = hb.load_library()
    L.hb_test_class_of.argtypes = [C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p,
                                   C.POINTER(C.c_int), C.POINTER(C.c_int)]
    a, c, logpi0 = _params(None, 4, 25000.0, 0.6, 1e-5, [0, 1e-4, 1e-3, 1e-2], [0.95, 0.02, 0.02, 0.01])
    TL, TH = np.full(3, -1.0), np.full(3, np.inf)
    by, ex = C.c_int(), C.c_int()
    last = 0
    for rr in np.linspace(0, 4e5, 400):
        L.hb_test_class_of(4, C.c_double(rr), C.c_double(0.97), a.ctypes.data, c.ctypes.data, C.c_double(logpi0), TL.ctypes.data,
                           TH.ctypes.data, C.byref(by), C.byref(ex))
        assert by.value == -1 and ex.value >= last
        last = ex.value
