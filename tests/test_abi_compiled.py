"""A compiled C++ caller of include/hibayes_b200.h (tests/abi/abi_check.cpp, g++ against libhibayes_b200.so): the struct
layouts the Python harness mirrors with ctypes are checked against the compiler's, and -- on a GPU -- hb_bayes() and
hb_sbayesd() called from C++ return what the ctypes wrappers return for the same data."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import hibayes_b200 as hb
from hibayes_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "abi", "abi_check")


def _build():
    src = os.path.join(ROOT, "tests", "abi", "abi_check.cpp")
    lib = _lib.library_path()
    hb.load_library()
    deps = [src, os.path.join(ROOT, "include", "hibayes_b200.h"), lib]
    if not os.path.exists(EXE) or any(os.path.getmtime(d) > os.path.getmtime(EXE) for d in deps):
        subprocess.check_call(["g++", "-O1", "-std=c++17", src, "-o", EXE, "-L" + os.path.dirname(lib), "-lhibayes_b200",
                               "-Wl,-rpath," + os.path.dirname(lib)])
    return EXE


def _kv(out):
    return dict(line.split(" ", 1) for line in out.strip().splitlines())


def test_struct_layouts_match_the_compiler():
    kv = _kv(subprocess.run([_build(), "layout"], capture_output=True, text=True, check=True).stdout)
    mirrors = {"hb_engine_config": _lib.EngineConfig, "hb_sweep_in": _lib.SweepIn, "hb_sweep_out": _lib.SweepOut,
               "hb_bayes_args": _lib.BayesArgs, "hb_bayes_out": _lib.BayesOut, "hb_sbayes_args": _lib.SBayesArgs,
               "hb_sbayes_out": _lib.SBayesOut, "hb_bed_source": _lib.BedSource}
    checked = 0
    for key, val in kv.items():
        kind, name = key.split(".", 1)
        if kind == "sizeof":
            if name in mirrors:
                assert C.sizeof(mirrors[name]) == int(val), (name, C.sizeof(mirrors[name]), val)
                checked += 1
        elif kind in mirrors:
            field = {"lambda": "lambda_"}.get(name, name)
            assert getattr(mirrors[kind], field).offset == int(val), (key, getattr(mirrors[kind], field).offset, val)
            checked += 1
    assert checked >= 60


@pytest.mark.gpu
def test_cpp_caller_gets_what_the_ctypes_wrappers_get():
    kv = _kv(subprocess.run([_build(), "run", "0"], capture_output=True, text=True, check=True).stdout)
    assert kv["bayes.rc"] == "0", kv.get("bayes.error")
    assert kv["sbayesd.rc"] == "0", kv.get("sbayesd.error")

    def lcg_stream(seed):
        s = seed
        while True:
            s = (s * 1664525 + 1013904223) & 0xffffffff
            yield s >> 8

    g = lcg_stream(12345)
    n, m = 500, 700
    X = np.array([next(g) % 3 for _ in range(n * m)], dtype=np.int8).reshape((n, m), order="F")
    y, cov = np.zeros(n), np.zeros(n)
    for i in range(n):
        cov[i] = (next(g) % 1000) / 500.0 - 1.0
        v = 0.7 * cov[i] + (next(g) % 2000) / 1000.0 - 1.0
        for j in range(10):
            v += 0.25 * float(X[i, j * 37])
        y[i] = v
    res = hb.Bayes(y, X, "BayesCpi", [0.9, 0.1], C_=cov[:, None], niter=10, nburn=4, thin=2, seed=2718)
    f = float.fromhex
    assert f(kv["bayes.Vg"]) == res["Vg"] and f(kv["bayes.Ve"]) == res["Ve"] and f(kv["bayes.mu"]) == res["mu"]
    assert f(kv["bayes.beta"]) == res["beta"][0] and f(kv["bayes.pi0"]) == res["pi"][0]
    assert int(kv["bayes.tracker_sum"]) == int((res["diag"]["tracker"].astype(np.int64) * (np.arange(m) + 1)).sum())
    asum = 0.0
    for j in range(m):
        asum += res["alpha"][j] * (j % 7 + 1)
    assert f(kv["bayes.alpha_sum"]) == asum
    ms = 300
    ld = np.zeros((ms, ms), order="F")
    ss = np.zeros((ms, 4), order="F")
    for j in range(ms):
        ld[j, j] = 0.4 + 0.001 * (j % 50)
        if j + 1 < ms:
            ld[j + 1, j] = 0.05
            ld[j, j + 1] = 0.05
        ss[j] = [0.2, ((next(g) % 2000) / 1000.0 - 1.0) * 0.05, 0.02, 5000.0]
    r2 = hb.SBayesD(ss, ld, "BayesR", [0.9, 0.05, 0.03, 0.02], fold=[0, 1e-4, 1e-3, 1e-2], niter=12, nburn=4, thin=2, seed=99)
    assert f(kv["sbayesd.Vg"]) == r2["Vg"] and f(kv["sbayesd.Ve"]) == r2["Ve"]
    sasum = 0.0
    for j in range(ms):
        sasum += r2["alpha"][j] * (j % 7 + 1)
    assert f(kv["sbayesd.alpha_sum"]) == sasum


def test_cpp_caller_fails_loudly_without_a_gpu():
    if hb.device_count() > 0:
        pytest.skip("a GPU is present")
    kv = _kv(subprocess.run([_build(), "run", "0"], capture_output=True, text=True, check=True).stdout)
    assert kv["bayes.rc"] != "0" and ("no CPU fallback" in kv["bayes.error"] or "CUDA" in kv["bayes.error"])
