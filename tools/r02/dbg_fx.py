import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, scipy.sparse as sp
import hibayes_b200 as hb
from oracle import hb_oracle
rng = np.random.default_rng(3)
n, m, ne, qe = 700, 1500, 220, 300
X = rng.integers(0, 3, size=(n, m)).astype(np.int8)
J = np.concatenate([-np.ones(n - ne), rng.uniform(-1, 0, ne)])
y = X[:, :30].astype(np.float64) @ rng.normal(scale=0.3, size=30) + rng.normal(size=n) + 1.5
A = sp.random(qe, qe, density=0.02, random_state=5, format="csr")
G = (A @ A.T + sp.diags(np.full(qe, 1.5))).tolil()
index1 = rng.permutation(qe)[:ne] + 1
G[index1[3] - 1, index1[3] - 1] = 0.0
G = sp.csc_matrix(G); G.eliminate_zeros()
Cm = np.column_stack([rng.normal(size=n), rng.integers(0, 2, n).astype(float)])
R = rng.integers(0, 5, size=(n, 1))
base = dict(niter=12, nburn=4, thin=2, seed=99)
variants = {"plain": {}, "C": dict(C_=Cm), "R": dict(R=R), "CR": dict(C_=Cm, R=R), "eps": dict(epsl_y_J=J, epsl_Gi=G, epsl_index=index1),
            "all": dict(C_=Cm, R=R, epsl_y_J=J, epsl_Gi=G, epsl_index=index1)}
for model, Pi, fold in [("BayesR", [0.9, 0.05, 0.03, 0.02], [0, 1e-4, 1e-3, 1e-2]), ("BayesRR", [0.0, 1.0], None), ("BayesCpi", [0.9, 0.1], None)]:
    for name, extra in variants.items():
        kw = dict(base, **extra)
        ref = hb_oracle.bayes(y, X.astype(np.float64), model, Pi, fold=fold, **kw)
        try:
            got = hb.Bayes(y, X, model, Pi, fold=fold, **kw)
            ok = np.array_equal(got["diag"]["tracker"], ref["diag"]["tracker"]) and abs(got["Ve"] / ref["Ve"] - 1) < 1e-5
            print(model, name, "ok" if ok else "MISMATCH", got["Ve"], ref["Ve"], flush=True)
        except Exception as e:
            print(model, name, "ERR", str(e)[:100], "oracle vare trace", ref["diag"]["vare_trace"][:4], "varg", ref["diag"]["varg_trace"][:3], flush=True)
