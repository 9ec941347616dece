"""Helpers shared by the tests: the bundled demo data set and synthetic genotypes."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_demo():
    return np.load(os.path.join(GOLDEN, "demo.npz"), allow_pickle=False)


def load_demo_T1():
    """y, X for `T1 ~ 1` as ibrm() would pass them to Bayes(): the individuals of the
    phenotype file that have a genotype and a non-missing T1, in phenotype-file order
    (R/bayes.r:161-165, 281-291)."""
    d = load_demo()
    gid = {s: i for i, s in enumerate(d["geno_id"])}
    rows, ys = [], []
    for pid, t in zip(d["phe_id"], d["T1"]):
        if pid in gid and not np.isnan(t):
            rows.append(gid[pid])
            ys.append(t)
    X = np.asfortranarray(d["geno"][rows, :])
    return np.array(ys, dtype=np.float64), X


def synth(n, m, seed=20260101, n_causal=None, h2=0.5):
    """Synthetic genotypes/phenotypes per SURVEY.md section 8(d): p_j ~ U(0.05,0.5),
    x_ij ~ Binomial(2,p_j) int8, causal effects N(0,1) scaled to var(Xb)=h2, e ~ N(0,1-h2)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    p = rng.uniform(0.05, 0.5, size=m)
    X = np.empty((n, m), dtype=np.int8, order="F")
    for j in range(m):
        X[:, j] = rng.binomial(2, p[j], size=n)
    if n_causal is None:
        n_causal = max(1, m // 100)
    idx = rng.choice(m, size=n_causal, replace=False)
    b = rng.normal(size=n_causal)
    gv = X[:, idx].astype(np.float64) @ b
    gv *= np.sqrt(h2 / gv.var())
    y = gv + rng.normal(scale=np.sqrt(1 - h2), size=n)
    return y, X
