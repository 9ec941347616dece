"""debug: first SNP whose class differs from the oracle, per lag / Pi / iteration count"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import hibayes_b200 as hb
from oracle import hb_oracle
from tests.util_demo import synth

FOLD_R = [0, 1e-4, 1e-3, 1e-2]
y, X = synth(3000, 2200, seed=101, n_causal=30)
for Pi in ([0.9, 0.05, 0.03, 0.02], [0.97, 0.01, 0.01, 0.01]):
    for lag in (1, 2, 3, 8):
        for niter in (1, 2, 4):
            kw = dict(niter=niter, nburn=0, thin=1, seed=4711)
            ref = hb_oracle.bayes(y, X, "BayesR", Pi, fold=FOLD_R, **kw)
            got = hb.Bayes(y, X, "BayesR", Pi, fold=FOLD_R, tile_snps=256, lag_tiles=lag, **kw)
            d = np.nonzero(got["diag"]["tracker"] != ref["diag"]["tracker"])[0]
            da = np.abs(got["alpha"] - ref["alpha"]).max() / (np.abs(ref["alpha"]).max() + 1e-300)
            nz = np.bincount(np.nonzero(ref["diag"]["tracker"])[0] // 256, minlength=9)
            print("Pi0 %.2f lag %d niter %d: class diffs %d first %s (tile %s) alpha rel err %.2e rounds %d/%d nz/tile %s" % (
                Pi[0], lag, niter, d.size, d[:3], (d[:3] // 256), da, got["diag"]["rounds_total"], got["diag"]["tiles_total"], nz), flush=True)
