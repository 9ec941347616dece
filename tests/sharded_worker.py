"""Worker of the world_size-2 tests (launched with torch.distributed.run, one process per rank).

mode "gloo": CPU only -- the collectives of hibayes_b200.sharded over gloo and the host-side sharding arithmetic.
mode "gpu":  one GPU per rank -- hb_bayes() on row shards against the CPU oracle on the full data."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    mode = sys.argv[1]
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    if mode.startswith("gpu"):
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
        dist.init_process_group("nccl")
    else:
        dist.init_process_group("gloo")
    from hibayes_b200.sharded import Comm, shard_rows
    from tests.util_demo import synth
    comm = Comm()
    assert comm.rank == rank and comm.world == world
    if mode == "gloo":
        from oracle import hb_oracle
        L = hb_oracle.lib()
        n = 1001
        y = np.random.default_rng(5).normal(size=n) * 3 + 1
        lo, hi = shard_rows(n, rank, world)
        assert comm.total_rows(hi - lo) == n
        # two-pass variance over the shards == Armadillo's var() on the whole vector (host_bayes.cpp, world > 1)
        s0 = comm.allreduce_f64(np.array([y[lo:hi].sum()]))[0]
        mean = s0 / n
        acc = comm.allreduce_f64(np.array([((mean - y[lo:hi]) ** 2).sum(), (mean - y[lo:hi]).sum()]))
        var = (acc[0] - acc[1] ** 2 / n) / (n - 1)
        ref = L.hbo_var(np.ascontiguousarray(y).ctypes.data, n)
        assert abs(var / ref - 1) < 1e-12, (var, ref)
        # exact integer column sums add up over the shards
        X = np.random.default_rng(6).integers(0, 3, size=(n, 40)).astype(np.float64)
        xpx = comm.allreduce_f64((X[lo:hi] ** 2).sum(axis=0).copy())
        assert np.array_equal(xpx, (X ** 2).sum(axis=0))
        # rank-ordered all-gather of fixed-size byte strings (the IPC handles)
        got = comm.allgather_bytes(bytes([rank + 1]) * 64)
        assert got == b"".join(bytes([r + 1]) * 64 for r in range(world))
        # the C callbacks wrap the same functions
        ar64, _, agb = comm.callbacks()
        buf = np.array([1.0 + rank, 2.0])
        import ctypes as C
        assert ar64(None, buf.ctypes.data_as(C.POINTER(C.c_double)), 2) == 0
        assert np.allclose(buf, [sum(1.0 + r for r in range(world)), 2.0 * world])
        print("rank %d gloo ok" % rank, flush=True)
    elif mode == "gpu_fx":
        # covariates, an environmental random effect and the single-step term on row shards (csrc/effects.cu with the sums
        # all-reduced): the epsilon rows are the tail of y, so rank 1 holds them all and rank 0 a few (uneven split)
        import hibayes_b200 as hb
        import scipy.sparse as sp
        from oracle import hb_oracle
        rng = np.random.default_rng(8)
        n, m, ne, qe = 2400, 1800, 1500, 1700
        X = rng.integers(0, 3, size=(n, m)).astype(np.int8)
        J = np.concatenate([-np.ones(n - ne), rng.uniform(-1, 0, ne)])
        y = X[:, :40].astype(np.float64) @ rng.normal(scale=0.3, size=40) + rng.normal(size=n) + 1.5
        A = sp.random(qe, qe, density=0.004, random_state=5, format="csr")
        G = sp.csc_matrix(A @ A.T + sp.diags(np.full(qe, 1.5)))
        index1 = rng.permutation(qe)[:ne] + 1
        Cm = np.column_stack([rng.normal(size=n), rng.integers(0, 2, n).astype(float)])
        R = np.column_stack([rng.integers(0, 6, size=n), rng.integers(0, 40, size=n)])
        kw = dict(niter=12, nburn=4, thin=2, seed=99)
        lo, hi = shard_rows(n, rank, world)
        ne_lo = max(lo, n - ne)           # first epsilon row of this shard
        loc_idx = index1[ne_lo - (n - ne):hi - (n - ne)] if hi > ne_lo else index1[:0]
        got = hb.Bayes(y[lo:hi], X[lo:hi], "BayesCpi", [0.9, 0.1], C_=Cm[lo:hi], R=R[lo:hi], epsl_y_J=J[lo:hi], epsl_Gi=G,
                       epsl_index=loc_idx, device=torch.cuda.current_device(), comm=comm, **kw)
        if rank == 0:
            ref = hb_oracle.bayes(y, X.astype(np.float64), "BayesCpi", [0.9, 0.1], C_=Cm, R=R, epsl_y_J=J, epsl_Gi=G, epsl_index=index1, **kw)
            assert np.array_equal(got["diag"]["tracker"], ref["diag"]["tracker"]), "class labels differ from the oracle"
            assert np.array_equal(got["diag"]["nnz_trace"], ref["diag"]["nnz_trace"])
            for k in ("Vg", "Ve", "h2", "mu", "Veps", "J"):
                assert abs(got[k] / ref[k] - 1) < 1e-5, (k, got[k], ref[k])
            for k in ("alpha", "beta", "epsilon", "r", "Vr"):
                assert np.abs(got[k] - ref[k]).max() <= 1e-5 * np.abs(ref[k]).max(), k
            assert np.allclose(got["e"], ref["e"][lo:hi], rtol=1e-5, atol=1e-5 * np.abs(ref["e"]).max())
            assert np.allclose(got["MCMCsamples"]["epsilon"], ref["MCMCsamples"]["epsilon"], rtol=1e-5, atol=1e-8)
        print("rank %d gpu_fx ok" % rank, flush=True)
    else:
        import hibayes_b200 as hb
        model = sys.argv[2] if len(sys.argv) > 2 else "BayesR"
        Pi = [0.9, 0.05, 0.03, 0.02] if model == "BayesR" else [0.9, 0.1]
        fold = [0, 1e-4, 1e-3, 1e-2] if model == "BayesR" else None
        y, X = synth(3001, 2500, seed=44, n_causal=25)
        kw = dict(niter=12, nburn=4, thin=2, seed=909)
        lo, hi = shard_rows(len(y), rank, world)
        got = hb.Bayes(y[lo:hi], X[lo:hi], model, Pi, fold=fold, device=torch.cuda.current_device(), comm=comm, **kw)
        if rank == 0:
            from oracle import hb_oracle
            ref = hb_oracle.bayes(y, X, model, Pi, fold=fold, **kw)
            assert np.array_equal(got["diag"]["tracker"], ref["diag"]["tracker"]), "class labels differ from the oracle"
            assert np.array_equal(got["diag"]["nnz_trace"], ref["diag"]["nnz_trace"])
            scale = np.abs(ref["alpha"]).max()
            assert np.abs(got["alpha"] - ref["alpha"]).max() < 1e-5 * scale
            for k in ("Vg", "Ve", "h2", "mu"):
                assert abs(got[k] / ref[k] - 1) < 1e-5, (k, got[k], ref[k])
            assert np.allclose(got["g"], ref["g"][lo:hi], rtol=1e-5, atol=1e-5 * np.abs(ref["g"]).max())
            assert np.allclose(got["e"], ref["e"][lo:hi], rtol=1e-5, atol=1e-5 * np.abs(ref["e"]).max())
        # every rank took the same decisions
        t = torch.from_numpy(got["diag"]["tracker"].astype(np.int64)).cuda()
        t0 = t.clone()
        dist.broadcast(t0, 0)
        assert torch.equal(t, t0)
        print("rank %d gpu %s ok" % (rank, model), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
