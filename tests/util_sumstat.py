"""Summary statistics and a dense LD matrix from genotypes, the way sbrm() would receive them: marginal regressions
(BETA, SE per SNP on the standardised scale of the COJO file) and LD = centred X'X / n (tXXmat.cpp:174-179)."""
import numpy as np


def make_sumstat(y, X, n_na=0, seed=0):
    X = np.asarray(X, dtype=np.float64)
    n, m = X.shape
    Xc = X - X.mean(axis=0)
    ld = Xc.T @ Xc / n
    yc = y - y.mean()
    xpx = (Xc * Xc).sum(axis=0)
    keep = xpx > 0
    beta = np.zeros(m)
    se = np.ones(m)
    beta[keep] = (Xc[:, keep].T @ yc) / xpx[keep]
    resid_var = ((yc[:, None] - Xc * beta) ** 2).sum(axis=0) / (n - 2)
    se[keep] = np.sqrt(resid_var[keep] / xpx[keep])
    maf = X.mean(axis=0) / 2
    ss = np.column_stack([maf, beta, se, np.full(m, float(n))])
    # monomorphic columns have no regression; the reference skips SNPs with NA BETA/SE/N (SBayesD.cpp:100-103)
    ss[~keep, 1:3] = np.nan
    if n_na:
        idx = np.random.default_rng(seed).choice(m, size=n_na, replace=False)
        ss[idx, 1] = np.nan
    return np.asfortranarray(ss), np.asfortranarray(ld)
