"""hibayes_b200 -- B200-native single-site Gibbs engine behind hibayes' Bayes()/SBayesD()/SBayesS().

Python is only the test/bench harness around the C ABI of include/hibayes_b200.h; the product is
libhibayes_b200.so (CUDA sm_100a kernels + C++ host driver).  There is no CPU fallback: loading
fails loudly when the library or a CUDA device is missing.
"""
from ._lib import load_library, last_error, device_count  # noqa: F401
from .api import Bayes, BedGeno, cutwind, Engine, LdMat, SBayesD, SBayesS, ibrm, ibrm_plan, ldmat, ldmat_plan, read_bed, sbrm, sbrm_plan, synth_geno_host, synth_geno_host_into  # noqa: F401
