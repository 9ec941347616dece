"""Golden vectors from the REFERENCE ITSELF: outputs of /root/reference/src/Bayes.cpp, SBayesD.cpp and SBayesS.cpp as
compiled into oracle/_ref/libhibayes_ref.so (oracle/Makefile; stand-in R/Rcpp/Armadillo headers in oracle/ref_shim), run
here on reference-held inputs (inst/extdata/demo.*, stored in demo.npz) with the random variates of the oracle's tape.

Run in the build container only (needs /root/reference):
    python tests/golden/make_ref_golden.py

Writes ref_bayes_<model>.npz (BASELINE config 1: demo data, T1 ~ 1; 60 iterations, 20 burn-in, thin 4),
ref_sbayesd_<model>.npz (demo.ma + the LD matrix of the demo genotypes) and ref_sbayess_<model>.npz (the stored
synthetic inputs of sbayess_inputs.npz).  Each file holds what the reference returned (alpha, pip, pi, Vg, Ve, h2, mu and
the MCMCsamples of Vg, Ve, pi and, for a subset of SNPs, alpha) and the run's arguments.  tests/test_reference_pin.py
compares the oracle with these files on every box (no reference needed) and the CUDA path with them on the GPU.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import hb_oracle  # noqa: E402
from tests.util_demo import load_demo_T1  # noqa: E402

MODELS = [
    ("BayesCpi", [0.95, 0.05], None),
    ("BayesC", [0.95, 0.05], None),
    ("BayesB", [0.95, 0.05], None),
    ("BayesBpi", [0.95, 0.05], None),
    ("BayesR", [0.95, 0.02, 0.02, 0.01], [0, 1e-4, 1e-3, 1e-2]),
    ("BayesRR", [0.95, 0.05], None),
    ("BayesA", [0.95, 0.05], None),
    ("BayesL", [0.95, 0.05], None),
]
KW = dict(niter=60, nburn=20, thin=4, seed=20240229)
KEEP = slice(0, 1000, 7)   # SNPs whose full alpha store is kept (the posterior mean of every SNP is kept anyway)


def pack(r, Pi, fold):
    mc = r["MCMCsamples"]
    out = {"alpha": r["alpha"], "pip": r["pip"], "pi": r["pi"], "Vg": r["Vg"], "Ve": r["Ve"], "h2": r["h2"],
           "store_Vg": mc["Vg"], "store_Ve": mc["Ve"], "store_pi": mc["pi"], "store_alpha_subset": mc["alpha"][KEEP, :],
           "nnz_per_record": (mc["alpha"] != 0).sum(axis=0), "Pi": np.array(Pi), "fold": np.array([] if fold is None else fold),
           "niter": KW["niter"], "nburn": KW["nburn"], "thin": KW["thin"], "seed": KW["seed"],
           "tape_len": r["replay"]["tape_len"], "tape_consumed": r["replay"]["consumed"]}
    if "mu" in r:
        out.update(mu=r["mu"], g=r["g"], e=r["e"], store_mu=mc["mu"])
    return out


def main():
    assert hb_oracle.ref_lib() is not None, "oracle/_ref/libhibayes_ref.so is not built (needs /root/reference)"
    import scipy.sparse as sp
    y, X = load_demo_T1()
    d = np.load(os.path.join(HERE, "demo.npz"))
    G = d["geno"].astype(np.float64)
    Gc = G - G.mean(axis=0)
    ld = np.asfortranarray(Gc.T @ Gc / G.shape[0])
    ss = np.asfortranarray(np.column_stack([d["ma_maf"], d["ma_beta"], d["ma_se"], d["ma_n"]]))
    inp = np.load(os.path.join(HERE, "sbayess_inputs.npz"))
    ss_s, ld_s = np.asfortranarray(inp["sumstat"]), sp.csc_matrix(inp["ld_thresholded"])
    for model, Pi, fold in MODELS:
        o = hb_oracle.bayes(y, X, model, Pi, fold=fold, record_tape=True, store_alpha=True, **KW)
        r = hb_oracle.bayes(y, X, model, Pi, fold=fold, replay_on_reference=o["tape"], store_alpha=True, **KW)
        np.savez_compressed(os.path.join(HERE, "ref_bayes_%s.npz" % model), **pack(r, Pi, fold))
        o = hb_oracle.sbayesd(ss, ld, model, Pi, fold=fold, record_tape=True, store_alpha=True, **KW)
        r = hb_oracle.sbayesd(ss, ld, model, Pi, fold=fold, replay_on_reference=o["tape"], store_alpha=True, **KW)
        np.savez_compressed(os.path.join(HERE, "ref_sbayesd_%s.npz" % model), **pack(r, Pi, fold))
        o = hb_oracle.sbayess(ss_s, ld_s, model, Pi, fold=fold, record_tape=True, store_alpha=True, **KW)
        r = hb_oracle.sbayess(ss_s, ld_s, model, Pi, fold=fold, replay_on_reference=o["tape"], store_alpha=True, **KW)
        np.savez_compressed(os.path.join(HERE, "ref_sbayess_%s.npz" % model), **pack(r, Pi, fold))
        print(model, "done")


if __name__ == "__main__":
    main()
