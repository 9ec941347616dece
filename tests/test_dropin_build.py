"""CPU checks of the reference-side binding (integration/rcpp): the drop-in bodies compile against the
stand-in Rcpp / Armadillo / bigmemory headers, link against libhibayes_b200.so, export the harness entry points, refuse bad
arguments with the reference's texts before touching the device, and fail loudly -- no CPU fallback -- without a GPU.
The results they return are checked on the GPU in tests/test_dropin_rcpp.py."""
import os
import subprocess

import numpy as np
import pytest

from oracle import hb_oracle
from tests.util_demo import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    if hb_oracle.dropin_lib() is None:
        pytest.skip("oracle/_ref/libhibayes_dropin.so is not built (needs hibayes_b200/libhibayes_b200.so)")
    return hb_oracle


def test_dropin_library_exports_the_harness_and_links_the_product():
    _lib()
    path = os.path.join(ROOT, "oracle", "_ref", "libhibayes_dropin.so")
    syms = subprocess.check_output(["nm", "-D", "--defined-only", path], text=True)
    for name in ("hbref_bayes", "hbref_sbayesd", "hbref_sbayess", "hbref_bigstat", "hbref_txxmat", "hbref_read_bed"):
        assert " T " + name in syms, name
    und = subprocess.check_output(["nm", "-D", "--undefined-only", path], text=True)
    for name in ("hb_bayes", "hb_sbayesd", "hb_sbayess", "hb_ldmat_dense", "hb_ldmat_sparse", "hb_bed_decode", "hb_last_error"):
        assert " U " + name in und, name          # the work is the product library's, through its C ABI
    assert "hbo_" not in und                      # ... and nothing of the oracle


def test_dropin_sources_use_only_the_public_header():
    for f in os.listdir(os.path.join(ROOT, "integration", "rcpp")):
        src = open(os.path.join(ROOT, "integration", "rcpp", f)).read()
        assert "oracle" not in src.replace("position-addressed", "") and "csrc/" not in src.replace("(csrc/hb_rng.h)", ""), f


@pytest.mark.parametrize("kwargs,text", [
    (dict(Pi=[0.5, 0.4]), "sum of Pi should be 1."),
    (dict(Pi=[1.0, 0.0]), "all markers have no effect size."),
    (dict(Pi=[0.9, 0.1], dfvg=1.5), "dfvg should not be less than 2."),
])
def test_dropin_bayes_refuses_with_the_references_texts(kwargs, text):
    L = _lib()
    y, X = synth(50, 20, seed=2, n_causal=2)
    kw = dict(niter=4, nburn=2, thin=1, seed=1)
    kw.update(kwargs)
    Pi = kw.pop("Pi")
    with pytest.raises(RuntimeError, match=text):
        L.bayes(y, X, "BayesCpi", Pi, replay_on_reference=hb_oracle.seed_tape(1), library="dropin", **kw)


def test_dropin_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    L = _lib()
    y, X = synth(50, 20, seed=2, n_causal=2)
    with pytest.raises(RuntimeError, match="CUDA|sm_100|device"):
        L.bayes(y, X, "BayesCpi", [0.9, 0.1], niter=4, nburn=2, thin=1, seed=1, replay_on_reference=hb_oracle.seed_tape(1), library="dropin")
