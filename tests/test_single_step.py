"""The single-step term of Bayes() (J + epsilon, /root/reference/src/Bayes.cpp:254-275, 554-584, and the sparse Gauss-Seidel
sampler Gibbs(sp_mat), src/solver.cpp:131-140): the oracle (oracle/hb_oracle.c) against an independent numpy restatement
written from the reference with DENSE matrices (Z, Z'Z, LHS = Z'Z + Gi * ve/veps, A.col(i) . x) on a tiny BayesRR chain, with and
without stored diagonal entries in Gi; the GPU test compares hb_bayes() -- whose host loop carries the same term -- with the
oracle."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import hb_oracle

# addresses of the draws (hibayes_b200/csrc/hb_rng.h)
DOM_ITER, DOM_SNP, DOM_EPS, DOM_K = 0, 1, 4, 5
IT_MU, IT_VARG, IT_VARE, IT_J, IT_VEPS, IT_VB = 0, 1, 2, 4, 5, 6
SL_MAIN = 1


def _draws(seed):
    L = hb_oracle.lib()
    L.hbo_draw_uz.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                              C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.hbo_draw_chisq.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_double]
    L.hbo_draw_chisq.restype = C.c_double
    L.hbo_var.argtypes = [C.c_void_p, C.c_int]
    L.hbo_var.restype = C.c_double

    def z(dom, it, idx, slot=0):
        u, zz = C.c_double(), C.c_double()
        L.hbo_draw_uz(seed, dom, it, idx, slot, 0, C.byref(u), C.byref(zz))
        return zz.value

    def chisq(it, idx, df):
        return L.hbo_draw_chisq(seed, DOM_ITER, it, idx, 0, float(df))

    def var(x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        return L.hbo_var(x.ctypes.data, len(x))

    return z, chisq, var


def bayesrr_single_step_numpy(y, X, J, Gi, index1, niter, nburn, thin, seed):
    """Bayes() for model "BayesRR" with the single-step term and nothing else, restated from Bayes.cpp with dense algebra."""
    z, chisq, var = _draws(seed)
    n, m = X.shape
    index = np.asarray(index1) - 1                                  # :255-256
    ne, qe = len(index), Gi.shape[0]
    Gi = np.asarray(Gi.todense()) if sp.issparse(Gi) else np.asarray(Gi)
    Z = np.zeros((ne, qe))                                         # :271-272
    Z[np.arange(ne), index] = 1.0
    ZZ = Z.T @ Z
    JtJ = float(J @ J)
    vary = var(y)
    dfvara, h2 = 4.0, 0.5
    vara = (dfvara - 2) / dfvara * vary * h2                        # :319-348
    vare = vary * (1 - h2)
    s2vara = vara * (dfvara - 2) / dfvara
    xpx = (X * X).sum(axis=0)
    vx = np.array([var(X[:, j]) for j in range(m)])
    sumvx = vx.sum()
    nvar0 = int((vx == 0).sum())
    varg = vara / ((1 - 0.0) * sumvx)                               # BayesRR: Pi = (0, 1)  (:288-292)
    s2varg = s2vara / sumvx
    dfvare, s2vare = -2.0, 0.0
    vepstmp = vara                                                  # :350
    mu = float(np.mean(y))
    yadj = y - mu
    u = np.zeros(n)
    g = np.zeros(m)
    estR, estR_tmp = np.zeros(qe), np.zeros(qe)
    Jb = 0.0
    rec = dict(mu=[], vara=[], vare=[], veps=[], J=[], eps=[], g=[])
    for it in range(niter):
        mu_ = -(yadj.sum() / n + np.sqrt(vare / n) * z(DOM_ITER, it, IT_MU))     # :480-482
        mu -= mu_
        yadj = yadj + mu_
        # ---- :554-584
        rhs = float(J @ yadj) + JtJ * Jb
        gi = rhs / JtJ + np.sqrt(vare / JtJ) * z(DOM_ITER, it, IT_J)
        yadj = yadj + (Jb - gi) * J
        u = u - (Jb - gi) * J
        Jb = gi
        LHS = ZZ + Gi * (vare / vepstmp)
        RHS = Z.T @ yadj[n - ne:] + ZZ @ estR_tmp
        for i in range(qe):                                         # solver.cpp:131-140
            invlhs = 1.0 / LHS[i, i]
            Ax = float(LHS[:, i] @ estR_tmp)
            uu = invlhs * (RHS[i] - Ax) + estR_tmp[i]
            estR_tmp[i] = uu + np.sqrt(invlhs * vare) * z(DOM_EPS, it, i)
        estR = estR - estR_tmp
        d = Z @ estR
        yadj[n - ne:] += d
        u[n - ne:] -= d
        vepstmp = float(estR_tmp @ Gi @ estR_tmp) + s2vara * dfvara
        vepstmp /= chisq(it, IT_VEPS, dfvara + qe)
        estR = estR_tmp.copy()
        veps = vepstmp
        # ---- BayesRR sweep :588-603
        for j in range(m):
            if vx[j] == 0:
                continue
            x = X[:, j]
            rhs = float(x @ yadj) + xpx[j] * g[j]
            v = xpx[j] + vare / varg
            gn = rhs / v + np.sqrt(vare / v) * z(DOM_SNP, it, j, SL_MAIN)
            yadj = yadj + (g[j] - gn) * x
            u = u - (g[j] - gn) * x
            g[j] = gn
        varg = (float(g @ g) + s2varg * dfvara) / chisq(it, IT_VARG, dfvara + m - nvar0)
        vara = var(u)                                               # :819
        vare = (float(yadj @ yadj) + s2vare * dfvare) / chisq(it, IT_VARE, n + dfvare)   # :823
        if it >= nburn and (it + 1 - nburn) % thin == 0:
            rec["mu"].append(mu); rec["vara"].append(vara); rec["vare"].append(vare); rec["veps"].append(veps)
            rec["J"].append(Jb); rec["eps"].append(estR.copy()); rec["g"].append(g.copy())
    return {"mu": np.mean(rec["mu"]), "Vg": np.mean(rec["vara"]), "Ve": np.mean(rec["vare"]), "Veps": np.mean(rec["veps"]),
            "J": np.mean(rec["J"]), "epsilon": np.mean(rec["eps"], axis=0), "alpha": np.mean(rec["g"], axis=0), "u": u}


def single_step_case(seed, n=60, m=24, ne=18, qe=25, drop_diag=()):
    """A tiny single-step data set: the last ne individuals carry an epsilon; Gi a sparse SPD matrix of the A^-1 kind (a
    few off-diagonals per row); some individuals of Gi share no record (their count in Z'Z is 0) and two records may
    point at the same entry."""
    rng = np.random.default_rng(seed)
    X = rng.integers(0, 3, size=(n, m)).astype(np.float64)
    X[:, 3] = 1.0                                                    # a monomorphic SNP (vx = 0 is skipped)
    X[n - ne:] += rng.normal(scale=0.3, size=(ne, m))                # imputed rows are real-valued (R/ssbayes.r:305)
    X[n - ne:, 3] = 1.0
    J = np.concatenate([-np.ones(n - ne), rng.uniform(-1, 0, ne)])   # (ssbayes.r:311-317 builds J this way)
    y = X @ rng.normal(scale=0.2, size=m) + rng.normal(size=n) + 3.0
    A = sp.random(qe, qe, density=0.12, random_state=seed, format="csr")
    G = (A @ A.T + sp.diags(np.full(qe, 2.0))).tolil()
    index1 = rng.choice(qe, size=ne, replace=True) + 1               # 1-based (Bayes.cpp:255-256)
    for k in drop_diag:                                              # entries of Gi without a stored diagonal value: A(i, i)
        G[index1[k] - 1, index1[k] - 1] = 0.0                        # is then the record count alone (> 0 for these)
    G = sp.csc_matrix(G)
    G.eliminate_zeros()
    return y, X, J, G, index1


@pytest.mark.parametrize("drop_diag", [(), (0, 5)])
def test_oracle_single_step_against_a_dense_numpy_restatement(drop_diag):
    y, X, J, G, index1 = single_step_case(7, drop_diag=drop_diag)
    kw = dict(niter=9, nburn=3, thin=2, seed=31337)
    ref = bayesrr_single_step_numpy(y, X, J, G, index1, **kw)
    got = hb_oracle.bayes(y, X, "BayesRR", [0.0, 1.0], epsl_y_J=J, epsl_Gi=G, epsl_index=index1, **kw)
    for key in ("mu", "Vg", "Ve", "Veps", "J"):
        assert abs(got[key] - ref[key]) <= 1e-9 * abs(ref[key]), (key, got[key], ref[key])
    assert np.allclose(got["epsilon"], ref["epsilon"], rtol=1e-9, atol=1e-12)
    assert np.allclose(got["alpha"], ref["alpha"], rtol=1e-9, atol=1e-12)
    assert np.allclose(got["g"], ref["u"], rtol=1e-9, atol=1e-11)     # the reference returns u as "g" (:1023)


def test_oracle_single_step_residual_identity():
    """e = y - mu - X alpha - J * Jhat - Z eps  (Bayes.cpp:942-1011) for a mixture model with the term."""
    y, X, J, G, index1 = single_step_case(11)
    got = hb_oracle.bayes(y, X, "BayesCpi", [0.9, 0.1], epsl_y_J=J, epsl_Gi=G, epsl_index=index1, niter=12, nburn=4, thin=2, seed=5)
    n, ne = len(y), len(index1)
    e = y - got["mu"] - X @ got["alpha"] - got["J"] * J
    e[n - ne:] -= got["epsilon"][index1 - 1]
    assert np.allclose(got["e"], e, rtol=1e-10, atol=1e-10)


@pytest.mark.gpu
@pytest.mark.parametrize("model,Pi,fold", [("BayesCpi", [0.9, 0.1], None), ("BayesR", [0.9, 0.05, 0.03, 0.02], [0, 1e-4, 1e-3, 1e-2]),
                                           ("BayesRR", [0.0, 1.0], None)])
def test_gpu_single_step_against_the_oracle(model, Pi, fold):
    """hb_bayes() with epsl_y_J / epsl_Gi / epsl_index (integer genotypes: the device tiles are int8) against the oracle,
    Gi with a missing diagonal entry, covariates and a random effect next to the term."""
    import hibayes_b200 as hb
    rng = np.random.default_rng(3)
    n, m, ne, qe = 700, 1500, 220, 300
    X = rng.integers(0, 3, size=(n, m)).astype(np.int8)
    J = np.concatenate([-np.ones(n - ne), rng.uniform(-1, 0, ne)])
    y = X[:, :30].astype(np.float64) @ rng.normal(scale=0.3, size=30) + rng.normal(size=n) + 1.5
    # an A^-1-like matrix: a few small off-diagonals per row, strictly diagonally dominant (the sampler's chain is stable)
    A = sp.random(qe, qe, density=0.012, random_state=5, format="csr", data_rvs=lambda k: rng.uniform(-0.15, 0.15, k))
    G = (A + A.T + sp.diags(np.full(qe, 2.0))).tolil()
    index1 = rng.permutation(qe)[:ne] + 1
    k0 = index1[3] - 1                          # an entry with nothing stored at all, not even its diagonal: A(i, i) is the
    G[k0, :] = 0.0                              # record count alone (Gibbs(sp_mat) reads A(i, i) of the sum, solver.cpp:134)
    G[:, k0] = 0.0
    G = sp.csc_matrix(G)
    G.eliminate_zeros()
    Cm = np.column_stack([rng.normal(size=n), rng.integers(0, 2, n).astype(float)])
    R = rng.integers(0, 5, size=(n, 1))
    kw = dict(niter=12, nburn=4, thin=2, seed=99, C_=Cm, R=R, epsl_y_J=J, epsl_Gi=G, epsl_index=index1)
    ref = hb_oracle.bayes(y, X.astype(np.float64), model, Pi, fold=fold, **kw)
    got = hb.Bayes(y, X, model, Pi, fold=fold, **kw)
    assert np.array_equal(got["diag"]["tracker"], ref["diag"]["tracker"])
    assert np.array_equal(got["diag"]["nnz_trace"], ref["diag"]["nnz_trace"])
    for key in ("mu", "Vg", "Ve", "Veps", "J", "h2"):
        assert abs(got[key] - ref[key]) <= 1e-5 * abs(ref[key]), (key, got[key], ref[key])
    sc = np.abs(ref["alpha"]).max()
    assert np.abs(got["alpha"] - ref["alpha"]).max() <= 1e-5 * sc
    assert np.abs(got["epsilon"] - ref["epsilon"]).max() <= 1e-5 * np.abs(ref["epsilon"]).max()
    assert np.abs(got["beta"] - ref["beta"]).max() <= 1e-5 * np.abs(ref["beta"]).max()
    assert np.abs(got["e"] - ref["e"]).max() <= 1e-5 * np.abs(ref["e"]).max()
    # MCMCsamples of the other terms (Bayes.cpp:867-876, 987-1020)
    for key in ("Veps", "J", "epsilon", "Vr", "r", "beta"):
        a_, b_ = got["MCMCsamples"][key], ref["MCMCsamples"][key]
        assert a_.shape == b_.shape and np.abs(a_ - b_).max() <= 1e-5 * np.abs(b_).max(), key
    assert np.abs(got["r"] - ref["r"]).max() <= 1e-5 * np.abs(ref["r"]).max()
    assert np.abs(got["Vr"] - ref["Vr"]).max() <= 1e-5 * np.abs(ref["Vr"]).max()


def bayesrr_bslmm_numpy(y, X, Kval, K, niter, nburn, thin, seed):
    """Bayes() for model "BayesRR" with the BSLMM polygenic term (Ki, Kival) and nothing else, restated from
    Bayes.cpp:203-233, 518-552, 854-858, 955-964 with numpy's dense algebra."""
    z, chisq, var = _draws(seed)
    n, m = X.shape
    nk = K.shape[1]
    vary = var(y)
    dfvara, h2 = 4.0, 0.5
    vara = (dfvara - 2) / dfvara * vary * h2
    vare = vary * (1 - h2)
    s2vara = vara * (dfvara - 2) / dfvara
    xpx = (X * X).sum(axis=0)
    vx = np.array([var(X[:, j]) for j in range(m)])
    sumvx = vx.sum()
    nvar0 = int((vx == 0).sum())
    varg = vara / sumvx
    s2varg = s2vara / sumvx
    dfvare, s2vare = -2.0, 0.0
    vbtmp = vara                                                    # :333
    mu = float(np.mean(y))
    yadj = y - mu
    u = np.zeros(n)
    g = np.zeros(m)
    k_estR, k_tmp, k_store = np.zeros(nk), np.zeros(nk), np.zeros(nk)
    rec = dict(mu=[], vara=[], vare=[], g=[])
    for it in range(niter):
        mu_ = -(yadj.sum() / n + np.sqrt(vare / n) * z(DOM_ITER, it, IT_MU))
        mu -= mu_
        yadj = yadj + mu_
        # ---- :518-552
        k_rhs = yadj + k_tmp
        ev = (Kval * vare) / (Kval + vare / vbtmp)
        k_tmp = K @ ((ev / vare) * (K.T @ k_rhs))
        assert np.all(ev >= -1e-6 * np.abs(ev).max())
        ev = np.where(ev < 0, 0.0, ev)
        rn = np.array([z(DOM_K, it, j) for j in range(nk)])
        k_tmp = k_tmp + K @ (np.sqrt(ev) * rn)
        k_estR = k_estR - k_tmp
        yadj = yadj + k_estR
        u = u - k_estR
        Kg = K.T @ k_tmp
        vbtmp = float(Kg @ ((1 / Kval) * Kg)) + s2vara * dfvara
        vbtmp /= chisq(it, IT_VB, dfvara + nk)
        k_estR = k_tmp.copy()
        # ---- BayesRR sweep :588-603
        for j in range(m):
            if vx[j] == 0:
                continue
            x = X[:, j]
            rhs = float(x @ yadj) + xpx[j] * g[j]
            v = xpx[j] + vare / varg
            gn = rhs / v + np.sqrt(vare / v) * z(DOM_SNP, it, j, SL_MAIN)
            yadj = yadj + (g[j] - gn) * x
            u = u - (g[j] - gn) * x
            g[j] = gn
        varg = (float(g @ g) + s2varg * dfvara) / chisq(it, IT_VARG, dfvara + m - nvar0)
        vara = var(u)
        vare = (float(yadj @ yadj) + s2vare * dfvare) / chisq(it, IT_VARE, n + dfvare)
        if it >= nburn and (it + 1 - nburn) % thin == 0:
            rec["mu"].append(mu); rec["vara"].append(vara); rec["vare"].append(vare); rec["g"].append(g.copy())
            k_store += k_estR
    k_store /= len(rec["mu"])
    Kg = (K.T @ k_store) / Kval / sumvx
    ghat = X.T @ (K @ Kg)
    ghat -= ghat.mean()
    alpha = np.mean(rec["g"], axis=0) + ghat
    Mu = np.mean(rec["mu"])
    return {"mu": Mu, "Vg": np.mean(rec["vara"]), "Ve": np.mean(rec["vare"]), "alpha": alpha, "u": u, "e": y - Mu - X @ alpha}


def bslmm_case(seed, n=50, m=30):
    rng = np.random.default_rng(seed)
    X = rng.integers(0, 3, size=(n, m)).astype(np.float64)
    X[:, 5] = 2.0
    y = X @ rng.normal(scale=0.2, size=m) + rng.normal(size=n) + 1.0
    Xc = X - X.mean(axis=0)
    Kmat = Xc @ Xc.T / m + 0.05 * np.eye(n)          # a GRM with the ridge R/bayes.r adds (lambda)
    Kval, K = np.linalg.eigh(Kmat)
    return y, X, Kval, np.asfortranarray(K)


def test_oracle_bslmm_against_a_dense_numpy_restatement():
    y, X, Kval, K = bslmm_case(3)
    kw = dict(niter=8, nburn=2, thin=2, seed=2024)
    ref = bayesrr_bslmm_numpy(y, X, Kval, K, **kw)
    got = hb_oracle.bayes(y, X, "BayesRR", [0.0, 1.0], Kival=Kval, Ki=K, **kw)
    for key in ("mu", "Vg", "Ve"):
        assert abs(got[key] - ref[key]) <= 1e-9 * abs(ref[key]), (key, got[key], ref[key])
    assert np.allclose(got["alpha"], ref["alpha"], rtol=1e-8, atol=1e-11)
    assert np.allclose(got["g"], ref["u"], rtol=1e-8, atol=1e-10)
    assert np.allclose(got["e"], ref["e"], rtol=1e-8, atol=1e-9)


@pytest.mark.gpu
@pytest.mark.parametrize("model,Pi,fold", [("BSLMM", [0.9, 0.1], None), ("BayesR", [0.9, 0.05, 0.03, 0.02], [0, 1e-4, 1e-3, 1e-2])])
def test_gpu_bslmm_against_the_oracle(model, Pi, fold):
    """hb_bayes() with Ki / Kival (the polygenic term of BSLMM, Bayes.cpp:518-552, on the device) against the oracle."""
    import hibayes_b200 as hb
    rng = np.random.default_rng(12)
    n, m = 640, 1200
    X = rng.integers(0, 3, size=(n, m)).astype(np.int8)
    y = X[:, :25].astype(np.float64) @ rng.normal(scale=0.3, size=25) + rng.normal(size=n) + 0.5
    Xc = X.astype(np.float64) - X.mean(axis=0)
    Kval, K = np.linalg.eigh(Xc @ Xc.T / m + 0.01 * np.eye(n))
    K = np.asfortranarray(K)
    kw = dict(niter=10, nburn=4, thin=2, seed=321, Kival=Kval, Ki=K, store_alpha=True)
    ref = hb_oracle.bayes(y, X.astype(np.float64), model, Pi, fold=fold, **kw)
    got = hb.Bayes(y, X, model, Pi, fold=fold, **kw)
    assert np.array_equal(got["diag"]["tracker"], ref["diag"]["tracker"])
    assert np.array_equal(got["diag"]["nnz_trace"], ref["diag"]["nnz_trace"])
    for key in ("mu", "Vg", "Ve", "h2"):
        assert abs(got[key] - ref[key]) <= 1e-5 * abs(ref[key]), (key, got[key], ref[key])
    assert np.abs(got["alpha"] - ref["alpha"]).max() <= 1e-5 * np.abs(ref["alpha"]).max()
    assert np.abs(got["g"] - ref["g"]).max() <= 1e-5 * np.abs(ref["g"]).max()
    assert np.abs(got["e"] - ref["e"]).max() <= 1e-5 * np.abs(ref["e"]).max()
    a_, b_ = got["MCMCsamples"]["alpha"], ref["MCMCsamples"]["alpha"]
    assert np.abs(a_ - b_).max() <= 1e-5 * np.abs(b_).max()
