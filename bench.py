#!/usr/bin/env python
"""bench.py -- Gibbs SNP-updates/s of the BayesR sweep (BASELINE.json metric) and the other BASELINE configs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config metric|c2|c3] [--scaling weak|strong]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one MCMC iteration's SNP sweep over the whole synthetic genotype matrix followed by the per-iteration
reductions -- what replaces Bayes.cpp:586-823 -- driven by the host-side scalar updates of the chain.

--config metric (default; what the driver runs)  BayesR, n = 50 000 x m = 1 000 000 int8 genotypes generated on the device.
    value   K steps timed with inputs resident in HBM (barrier + synchronize on both sides, max over ranks).
    e2e     N = 1: the drop-in call hb_bayes() (what _hibayes_Bayes would forward to) with HOST y and X for W + K iterations;
            the timed region is the whole call minus its set-up (X host->device + column statistics + Gram band, reported
            as e2e.x_load_s), i.e. y upload, every sweep with its host-side scalar updates, and the download of the results.
            N > 1 (and hosts too small for the 50 GB matrix): the engine steps with the residual crossing PCIe both ways.
    N > 1   --scaling weak (default): every rank holds 50 000 rows of an N x 50 000-row matrix; a unit of `value` is one
            SNP update over one rank's shard (config.snp_updates_per_s is the plain m x steps / time).
            --scaling strong: the 50 000 rows are split over the ranks (last row).
            `parity_check`: an untimed small row-sharded run against the CPU oracle on rank 0, and equality of the
            class labels over the ranks after the timed run.
--config c2   BASELINE configs[1]: ibrm() BayesR n = 50 000 x m = 500 000, 1000 iterations through hb_bayes() with host
              buffers; ms per sweep and rounds per tile around iterations 10 / 100 / 500 / 1000.
--config c3   BASELINE configs[2]: BayesB n = 200 000 x m = 1 000 000 row-sharded over 8 GPUs (25 000 rows per GPU).
--config c4   BASELINE configs[3]: sbrm() SBayesD on a dense fp64 LD matrix of the largest m that fits (--m, default 100 000).
--config c5   BASELINE configs[4] surrogate: the single-step call of Bayes() (J + epsilon on the device) with integer rows, 1 GPU.
`--impl reference`  the reference's own CPU data path (per-SNP ddot + 2 daxpy on a column-major fp64 matrix,
            Bayes.cpp:751-802, all host threads) on a bounded column sample of the same workload: Bayes() of the
            reference's own Bayes.cpp as compiled into oracle/_ref (stand-in R/Rcpp/Armadillo headers, OpenBLAS level 1),
            with the oracle port's two variants of the bare data path listed beside it.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "gibbs_snp_updates_per_sec_bayesr_n50k"
UNIT = "SNP-updates/s"
PI_R = [0.95, 0.02, 0.02, 0.01]
FOLD = [0.0, 1e-4, 1e-3, 1e-2]
if os.environ.get("HB_BENCH_FOLD_SCALE"):   # experiments only: stronger effects stress the speculation of the scalar chain
    FOLD = [f * float(os.environ["HB_BENCH_FOLD_SCALE"]) for f in FOLD]
KERNELS_PER_STEP = 6   # k_prep, k_absmax (scale of the integer dots), k_sweep, k_post1, k_post2, k_tail


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json, copy bandwidth)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def _traffic():
    """dram bytes per k_sweep launch at the metric shape, from the committed ncu capture (profiles/); N = 1 only."""
    for name in ("r02_sweep_traffic.json", "r01_sweep_traffic.json"):
        try:
            return json.load(open(os.path.join(ROOT, "profiles", name)))
        except Exception:
            continue
    return None


class ClockSampler:
    def __init__(self, gpu=0):
        self.gpu, self.samples, self.reasons, self.stop = gpu, [], set(), threading.Event()
        self.maxc = None
        self.th = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.maxc = float(out[1])
                for nm, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            self.stop.wait(0.1)

    def __enter__(self):
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.th.join(timeout=6)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.maxc,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_reference(n, m_cpu, sweeps, threads):
    """The reference's CPU data path on an n x m_cpu fp64 column-major sample: the oracle's OpenMP level-1 loops
    (hbo_time_sweep_fp64: one untimed warm-up sweep, then `sweeps` timed ones) and the same per-SNP ddot / 2 daxpy through
    the bundled OpenBLAS (scipy.linalg.blas, multi-threaded), the stand-in for MKL; the faster of the two is reported
    (BASELINE.md section 3)."""
    from oracle import hb_oracle
    hb_oracle.lib()
    v_omp, _ = hb_oracle.time_sweep_fp64(n, m_cpu, sweeps, threads)
    v_blas = None
    try:
        v_blas = _cpu_openblas(n, min(m_cpu, 4000), sweeps)
    except Exception:
        pass
    if v_blas is not None and v_blas > v_omp:
        return v_blas, "OpenBLAS ddot/daxpy (scipy.linalg.blas)", {"openmp": v_omp, "openblas": v_blas}
    return v_omp, "OpenMP ddot/daxpy", {"openmp": v_omp, "openblas": v_blas}


def _openblas_path():
    import glob
    import scipy
    hits = glob.glob(os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)), "scipy.libs", "libscipy_openblas-*.so"))
    return (hits[0], "scipy_") if hits else (None, None)


def cpu_reference_compiled(n, m_cpu, sweeps, blas=True):
    """The reference ITSELF on the CPU: Bayes() of /root/reference/src/Bayes.cpp as compiled into
    oracle/_ref/libhibayes_ref.so (oracle/Makefile, stand-in R/Rcpp/Armadillo headers), model BayesR on the first m_cpu
    columns of the workload as an fp64 arma::mat (what ibrm() hands over), its ddot_/daxpy_ forwarded to the bundled
    multi-threaded OpenBLAS (blas=True; an R linked against OpenBLAS) or run by reference-BLAS loops.  Its random draws
    replay the oracle's tape for the same call; the timed region is iterations 2..1+sweeps of the MCMC loop, from the first
    draw of iteration 2 to the last draw of the run (time stamps taken inside the library at those tape positions).
    Returns SNP-updates/s or None when the library is not there."""
    from oracle import hb_oracle
    import ctypes as C
    import hibayes_b200 as hb
    R = hb_oracle.ref_lib()
    if R is None:
        return None
    X = hb.synth_geno_host(n, m_cpu, 20260101)
    rng = np.random.default_rng(7)
    y = X[:, :50].astype(np.float64) @ rng.normal(scale=0.1, size=50) + rng.normal(size=n)
    kw = dict(fold=FOLD, thin=1, seed=12345)
    p1 = len(hb_oracle.bayes(y, X, "BayesR", PI_R, niter=1, nburn=0, record_tape=True, **kw)["tape"])
    tape = hb_oracle.bayes(y, X, "BayesR", PI_R, niter=1 + sweeps, nburn=sweeps, record_tape=True, **kw)["tape"]
    path, prefix = _openblas_path() if blas else (None, None)
    R.hbref_use_blas.argtypes = [C.c_char_p, C.c_char_p]
    if R.hbref_use_blas(path.encode() if path else None, prefix.encode() if prefix else None) != 0:
        raise RuntimeError(R.hbref_last_error().decode())
    marks = (C.c_uint64 * 2)(p1, len(tape) - 1)
    R.hbref_set_time_marks(marks, 2)
    try:
        hb_oracle.bayes(y, X, "BayesR", PI_R, niter=1 + sweeps, nburn=sweeps, replay_on_reference=tape, **kw)
    finally:
        R.hbref_use_blas(None, None)
    t = (C.c_double * 2)()
    R.hbref_get_time_marks(t, 2)
    R.hbref_set_time_marks(marks, 0)
    return m_cpu * sweeps / (t[1] - t[0])


def _cpu_openblas(n, m_cpu, sweeps):
    from scipy.linalg import blas
    rng = np.random.default_rng(1)
    X = np.asfortranarray(rng.integers(0, 3, size=(n, m_cpu)).astype(np.float64))
    r = rng.standard_normal(n)
    u = np.zeros(n)
    g = np.zeros(m_cpu)
    xpx = (X * X).sum(axis=0)
    un = rng.uniform(size=(sweeps + 1, m_cpu))
    zn = rng.standard_normal(size=(sweeps + 1, m_cpu))
    t_total = 0.0
    for sw in range(sweeps + 1):
        t0 = time.perf_counter()
        for j in range(m_cpu):
            x = X[:, j]
            rhs = blas.ddot(x, r) + xpx[j] * g[j]
            inc = un[sw, j] < 0.05                         # (the class decision is negligible next to the three passes)
            gn = (rhs / (xpx[j] + 1e3) + 1e-3 * zn[sw, j]) if inc else 0.0
            d = g[j] - gn
            if d != 0.0:
                blas.daxpy(x, r, a=d)
                blas.daxpy(x, u, a=-d)
            g[j] = gn
        if sw > 0:
            t_total += time.perf_counter() - t0
    return m_cpu * sweeps / t_total


# ------------------------------------------------------------------------------------------------ chains
class Chain:
    """Host-side scalar updates of the chain around hb_engine_sweep (Bayes.cpp:480-482, 664-670, 803-823); numpy's
    generator stands in for the host draws (same seed on every rank) -- only the workload matters here."""

    def __init__(self, model, n_total, vary, sumvx, seed):
        self.model = model
        self.n = n_total
        self.rng = np.random.default_rng(seed)
        self.df = 4.0
        self.pi = np.array(PI_R if model == "BayesR" else [0.95, 0.05])
        vara = (self.df - 2) / self.df * vary * 0.5
        self.vare = vary * 0.5
        self.varg = vara / ((1 - self.pi[0]) * sumvx)
        self.s2varg = vara * (self.df - 2) / self.df / ((1 - self.pi[0]) * sumvx)
        self.sum_r, self.sum_r2 = 0.0, vary * (n_total - 1)
        self.it = 0

    def sweep_args(self):
        mu_ = -(self.sum_r / self.n + math.sqrt(self.vare / self.n) * self.rng.standard_normal())
        rn2 = self.sum_r2 + 2 * mu_ * self.sum_r + self.n * mu_ * mu_
        if self.model == "BayesR":
            return dict(iter=self.it, model_index=6, vare=self.vare, logpi=list(np.log(self.pi)),
                        vara_fold=[self.varg * f for f in FOLD], fold=FOLD, dfvara=self.df, s2varg=self.s2varg,
                        mu_shift=mu_, rnorm2_bound=rn2)
        # BayesB: per-SNP variances are drawn on the device from dfvara / s2varg (Bayes.cpp:636)
        return dict(iter=self.it, model_index=3, vare=self.vare, logpi=list(np.log(self.pi)) + [0.0, 0.0],
                    vara_fold=[0.0, self.varg, 0.0, 0.0], fold=[0.0, 1.0, 0.0, 0.0], dfvara=self.df, s2varg=self.s2varg,
                    mu_shift=mu_, rnorm2_bound=rn2)

    def update(self, so, sum_r, sum_r2):
        F = len(self.pi)
        cnt = np.array(so["count"][:F])
        nnz = cnt[1:].sum()
        if self.model == "BayesR":
            self.varg = (so["varg_acc"] + self.s2varg * self.df) / self.rng.chisquare(self.df + nnz)
        self.pi = self.rng.dirichlet(cnt + 1)
        self.vare = sum_r2 / self.rng.chisquare(self.n - 2)
        self.sum_r, self.sum_r2 = sum_r, sum_r2
        self.it += 1


def cpu_arm(n, m_cpu, sweeps, threads):
    """(value, kind, sample, variants) of the CPU arm: the compiled reference (oracle/_ref, kind "reference") when the
    library is there -- its own Bayes() with OpenBLAS level 1 -- else the oracle port of its data path (kind "port").
    The port's two variants (the bare ddot / 2 daxpy data path, without the class probabilities) are measured in both
    cases and listed under `variants`."""
    v_port, how, both = cpu_reference(n, m_cpu, sweeps, threads)
    both = dict(both)
    v_ref = None
    try:
        m_ref = min(m_cpu, 8000)   # (the compiled reference holds X as fp64 and the oracle run that records its tape is serial)
        v_ref = cpu_reference_compiled(n, m_ref, sweeps, blas=True)
        both["compiled_reference_openblas"] = v_ref
        both["compiled_reference_netlib_loops_1_thread"] = cpu_reference_compiled(n, min(m_cpu, 2000), 2, blas=False)
    except Exception as e:   # noqa: BLE001 -- the arm must print a line in any case
        both["compiled_reference_error"] = str(e)[:200]
    if v_ref is not None:
        note = "" if v_ref >= v_port else ("; the oracle port of the bare data path (no class probabilities) runs at %.0f/s, see variants" % v_port)
        return v_ref, "reference", ("Bayes() of the reference's own Bayes.cpp compiled into oracle/_ref (BayesR, n=%d x m=%d fp64 "
                                    "arma::mat = the first columns of the workload, ddot_/daxpy_ -> bundled multi-threaded OpenBLAS), "
                                    "1 warm-up + %d timed iterations of its MCMC loop%s" % (n, m_ref, sweeps, note)), both
    sample = ("n=%d x m=%d fp64 column-major, 1 warm-up + %d timed sweeps, %s (oracle port of Bayes.cpp:751-802; oracle/_ref not "
              "built)" % (n, m_cpu, sweeps, how))
    return v_port, "port", sample, both


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n, m_cpu = args.n, args.m_cpu
    sweeps = max(3, min(args.steps, 5))   # each "step" = one sweep over the m_cpu-column sample; >= 3 timed sweeps
    t0 = time.time()
    val, kind, sample, both = cpu_arm(n, m_cpu, sweeps, threads)
    wall = time.time() - t0
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * m_cpu / val, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "ibrm() BayesR sweep, synthetic n=%d x m=%d (CPU sample: first %d columns)" % (n, args.m, m_cpu),
                   "n": n, "m": args.m, "m_sample": m_cpu, "variants": both},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": wall,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ host data
def _mem_available():
    try:
        for ln in open("/proc/meminfo"):
            if ln.startswith("MemAvailable:"):
                return int(ln.split()[1]) * 1024
    except Exception:
        pass
    return 0


def host_genotypes(n, m, seed, threads=None):
    """The device generator's matrix on the host (column-major int8), blocks of columns on several threads."""
    import hibayes_b200 as hb
    X = np.empty((n, m), dtype=np.int8, order="F")
    threads = threads or min(32, os.cpu_count() or 1)
    step = max(1, -(-m // (4 * threads)))
    blocks = [(c0, min(m, c0 + step)) for c0 in range(0, m, step)]

    def work(b):
        c0, c1 = b
        hb.synth_geno_host_into(X[:, c0:c1], seed, col_offset=c0)

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(work, blocks))
    return X


def product_call(n, m, niter, seed, device, comm=None):
    """hb_bayes() with host buffers: synthetic y from 1000 causal SNPs (h2 = 0.5), BayesR defaults of R/bayes.r."""
    import hibayes_b200 as hb
    t0 = time.time()
    X = host_genotypes(n, m, seed)
    rng = np.random.default_rng(seed)
    causal = rng.choice(m, size=min(1000, m), replace=False)
    gv = X[:, causal].astype(np.float64) @ rng.standard_normal(causal.size)
    y = gv * math.sqrt(0.5 / gv.var()) + np.random.default_rng(seed + 1).normal(scale=math.sqrt(0.5), size=n)
    t_gen = time.time() - t0
    t0 = time.perf_counter()
    res = hb.Bayes(y, X, "BayesR", PI_R, fold=FOLD, niter=niter, nburn=niter // 2, thin=5, seed=seed, device=device, comm=comm)
    wall = time.perf_counter() - t0
    return res, wall, t_gen, X.nbytes


# ------------------------------------------------------------------------------------------------ GPU arm
def parity_check_sharded(comm, local_rank):
    """Untimed: hb_bayes() on row shards of a small problem against the CPU oracle on the full data (rank 0), and the
    same class labels on every rank.  Returns a dict for the JSON line."""
    import torch
    import hibayes_b200 as hb
    from hibayes_b200.sharded import shard_rows
    from tests.util_demo import synth
    out = {}
    y, X = synth(3001, 2500, seed=44, n_causal=25)
    kw = dict(niter=12, nburn=4, thin=2, seed=909)
    Pi = [0.9, 0.05, 0.03, 0.02]
    lo, hi = shard_rows(len(y), comm.rank, comm.world)
    got = hb.Bayes(y[lo:hi], X[lo:hi], "BayesR", Pi, fold=FOLD, device=local_rank, comm=comm, **kw)
    ok = True
    if comm.rank == 0:
        from oracle import hb_oracle
        ref = hb_oracle.bayes(y, X, "BayesR", Pi, fold=FOLD, **kw)
        ok = bool(np.array_equal(got["diag"]["tracker"], ref["diag"]["tracker"])
                  and np.array_equal(got["diag"]["nnz_trace"], ref["diag"]["nnz_trace"])
                  and np.abs(got["alpha"] - ref["alpha"]).max() < 1e-5 * np.abs(ref["alpha"]).max()
                  and abs(got["Ve"] / ref["Ve"] - 1) < 1e-5 and abs(got["Vg"] / ref["Vg"] - 1) < 1e-5)
    t = torch.from_numpy(got["diag"]["tracker"].astype(np.int64)).cuda()
    t0 = t.clone()
    comm.dist.broadcast(t0, 0)
    same = bool(torch.equal(t, t0))
    flags = torch.tensor([int(ok), int(same)], device="cuda")
    comm.dist.all_reduce(flags, op=comm.dist.ReduceOp.MIN)
    out["small_sharded_vs_oracle"] = "ok" if int(flags[0]) else "MISMATCH"
    out["small_sharded_labels_equal_across_ranks"] = bool(int(flags[1]))
    out["shape"] = "BayesR n=3001 x m=2500, 12 iterations, rows over %d ranks" % comm.world
    return out


def run_gpu(args):
    import hibayes_b200 as hb
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    comm = None
    if world > 1:
        import torch
        import torch.distributed as dist
        from hibayes_b200.sharded import Comm
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl")
        comm = Comm()

    def allsum(a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        return comm.allreduce_f64(a) if comm else a

    def barrier():
        if comm:
            comm.dist.barrier()

    model = "BayesB" if args.config == "c3" else "BayesR"
    m = args.m
    if args.config == "c3":
        n_local = 200000 // max(world, 1) if world > 1 else 25000
        scaling = "strong"   # the configuration is fixed: 200 000 rows over the ranks
    elif args.scaling == "strong" and world > 1:
        n_local = -(-args.n // world)
        n_local = -(-n_local // 4) * 4
        scaling = "strong"
    else:
        n_local = args.n
        scaling = "weak"
    n_total = n_local * world
    parity = None
    if comm:
        parity = parity_check_sharded(comm, local_rank)
    t_setup = time.time()
    eng = hb.Engine(n_local, m, device=local_rank, tile_snps=args.tile, lag_tiles=args.lag, seed=args.seed, rank=rank, world=world)
    eng.synth_geno(args.seed, row_offset=rank * n_local)
    xpx, sumx = eng.col_stats()
    xpx, sumx = allsum(xpx), allsum(sumx)
    vx = (xpx - sumx * sumx / n_total) / (n_total - 1)
    active = (n_total * xpx != sumx * sumx)
    eng.set_snp_info(xpx, active.astype(np.uint8))
    t_gram = time.time()
    eng.build_gram()
    if comm:
        ptr, cnt = eng.gram_device()
        comm.allreduce_i32_device(ptr, cnt)
        eng.set_peers(comm.allgather_bytes(eng.ipc_handle()))
    t_gram = time.time() - t_gram
    rng = np.random.default_rng(args.seed)
    beta = np.zeros(m)
    causal = rng.choice(m, size=min(1000, m), replace=False)
    beta[causal] = rng.standard_normal(causal.size)
    gv = eng.predict(beta)
    s1 = allsum(np.array([gv.sum(), (gv * gv).sum()]))
    gvar = s1[1] / n_total - (s1[0] / n_total) ** 2
    gv *= math.sqrt(0.5 / gvar)
    y = gv + np.random.default_rng(args.seed + 1 + rank).normal(scale=math.sqrt(0.5), size=n_local)
    s2 = allsum(np.array([y.sum(), (y * y).sum()]))
    ymean = s2[0] / n_total
    vary = (s2[1] - n_total * ymean * ymean) / (n_total - 1)
    r = y - ymean
    eng.set_residual(r)
    chain = Chain(model, n_total, float(vary), float(vx.sum()), args.seed)
    s3 = allsum(np.array([r.sum(), r @ r]))
    chain.sum_r, chain.sum_r2 = float(s3[0]), float(s3[1])
    t_setup = time.time() - t_setup
    desc = eng.describe()

    def step():
        so = eng.sweep(**chain.sweep_args())
        s = allsum(np.array([so["sum_r"], so["sum_r2"]]))
        chain.update(so, float(s[0]), float(s[1]))
        return so

    import torch
    for _ in range(args.warmup):
        step()
    # ---- timed region: K steps, inputs resident in HBM
    sweep_ms, dev_ms, changed, rounds = [], [], [], []
    with ClockSampler(local_rank) as clk:
        torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            so = step()
            a, b, c = eng.last_sweep_ms()
            sweep_ms.append(b)
            dev_ms.append(a + b + c)
            changed.append(so["n_changed"])
            rounds.append(so["rounds"])
        torch.cuda.synchronize()
        barrier()
        wall = time.perf_counter() - t0
    wall = float(np.max(allsum_max(comm, wall)))
    m_active = int(active.sum())
    snp_updates = m_active * args.steps / wall                # plain SNP updates of the n_total-row model per second
    value = (world if scaling == "weak" else 1) * snp_updates  # weak scaling: one unit = one SNP update over one rank's shard
    if comm:
        # every rank took the same decisions in the timed run
        t = torch.from_numpy(eng.get_tracker().astype(np.int64)).cuda()
        t0_ = t.clone()
        comm.dist.broadcast(t0_, 0)
        flag = torch.tensor([int(torch.equal(t, t0_))], device="cuda")
        comm.dist.all_reduce(flag, op=comm.dist.ReduceOp.MIN)
        parity["timed_run_labels_equal_across_ranks"] = bool(int(flag[0]))
    # ---- end-to-end region through the engine: same steps, residual over PCIe in both directions
    r = eng.get_residual()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        eng.set_residual(r)                 # H2D: n doubles
        so = step()
        r = eng.get_residual()              # D2H: n doubles
        _ = eng.get_u()                     # D2H: n doubles
    torch.cuda.synchronize()
    barrier()
    e2e_s = float(np.max(allsum_max(comm, time.perf_counter() - t0)))
    e2e = {"value": (world if scaling == "weak" else 1) * m_active * args.steps / e2e_s, "unit": UNIT,
           "h2d_bytes_per_step": 8 * n_local, "d2h_bytes_per_step": 16 * n_local + 160,
           "path": "hb_engine_* steps with host residual buffers (the residual crosses PCIe both ways every step)"}
    eng.close()
    del eng
    # ---- end-to-end through the drop-in call (N = 1, metric shape, when the host can hold X)
    if world == 1 and args.config == "metric" and not args.no_product:
        need = int(1.7 * n_local * m)
        if _mem_available() > need:
            niter = args.warmup + args.steps
            res, call_s, gen_s, xbytes = product_call(n_local, m, niter, args.seed, local_rank)
            dg = res["diag"]
            run_s = call_s - dg["seconds_setup"]
            e2e = {"value": m_active * niter / run_s, "unit": UNIT,
                   "h2d_bytes_per_step": int((8 * n_local) / niter), "d2h_bytes_per_step": int((3 * 8 * m + 16 * n_local) / niter),
                   "path": "hb_bayes() with host y and X (int8), %d iterations; timed = the call minus its set-up" % niter,
                   "x_load_s": dg["seconds_setup"], "x_bytes": int(xbytes), "call_s": call_s, "host_generation_s": gen_s,
                   "value_incl_x_load": m_active * niter / call_s,
                   "device_ms_per_sweep": float(np.mean(dg["sweep_ms_trace"][args.warmup:])),
                   "rounds_per_tile": float(dg["rounds_total"]) / max(1, dg["tiles_total"])}
        else:
            e2e["note"] = "host memory too small for the %d GB int8 matrix: engine-level steps" % (n_local * m // 10**9)
    peak, peak_src = _peaks()
    kern_s = float(np.mean(sweep_ms)) * 1e-3
    achieved = n_local * m / kern_s / 1e9   # algorithmic bytes: n per SNP update (one read of the int8 column)
    tr = _traffic() if (world == 1 and args.config == "metric" and n_local == 50000 and m == 1000000) else None
    workload = {"metric": "ibrm() BayesR sweep", "c3": "ibrm() BayesB sweep (BASELINE configs[2])", "c2": ""}[args.config]
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%s, synthetic int8 genotypes, n=%d individuals per GPU x m=%d SNPs, %d GPU(s), n_total=%d"
                               % (workload, n_local, m, world, n_total),
                   "n_per_gpu": n_local, "n_total": n_total, "m": m, "m_active": m_active, "model": model,
                   "Pi": list(chain.pi) if False else (PI_R if model == "BayesR" else [0.95, 0.05]), "fold": FOLD,
                   "unit_of_value": ("one SNP update over one rank's %d-row shard" % n_local) if scaling == "weak"
                                    else "one SNP update of the %d-row model" % n_total,
                   "snp_updates_per_s": snp_updates,
                   "layout": desc, "l2": "inputs (%.1f GB per GPU) larger than L2" % (desc["geno_bytes"] / 1e9),
                   "changed_snps_per_sweep": float(np.mean(changed)), "scalar_rounds_per_sweep": float(np.mean(rounds)),
                   "rounds_per_tile": float(np.mean(rounds)) / max(1, -(-m // desc["tile_snps"])),
                   "device_ms_per_step": float(np.mean(dev_ms)), "setup_s": t_setup, "gram_s": t_gram,
                   "frac_of_8TBps": achieved / 8000.0},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": (tr or {}).get("dram_bytes_per_launch"), "traffic_source": (tr or {}).get("source"),
                     "peak_source": peak_src, "kernel": "k_sweep", "kernel_ms": kern_s * 1e3,
                     "algorithmic_bytes_per_launch": n_local * m},
        "e2e": e2e,
        "gpu_launches": KERNELS_PER_STEP * args.steps,
        "clocks": clk.summary(),
    }
    if parity is not None:
        line["parity_check"] = parity
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        val, kind, sample, both = cpu_arm(n_local, args.m_cpu, 3, threads)
        line["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample, "variants": both}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if comm:
        comm.dist.barrier()
        comm.dist.destroy_process_group()


def run_c2(args):
    """BASELINE configs[1]: the drop-in call on n = 50 000 x m = 500 000 for 1000 iterations."""
    n, m, niter = args.n, 500000 if args.m == 1000000 else args.m, args.niter
    res, call_s, gen_s, xbytes = product_call(n, m, niter, args.seed, 0)
    dg = res["diag"]
    T = dg["tiles_total"] // max(1, dg["iters_done"])
    run_s = call_s - dg["seconds_setup"]

    def around(it):
        lo, hi = max(0, it - 10), min(niter, it)
        return {"iteration": it, "device_ms_per_sweep": float(np.mean(dg["sweep_ms_trace"][lo:hi])),
                "rounds_per_tile": float(np.mean(dg["rounds_trace"][lo:hi])) / max(1, T),
                "nnz_snps": int(dg["nnz_trace"][hi - 1])}

    peak, peak_src = _peaks()
    ms_steady = float(np.mean(dg["sweep_ms_trace"][niter // 2:]))
    line = {
        "metric": METRIC, "value": m * niter / run_s, "unit": UNIT, "n_gpus": 1, "steps": niter, "warmup": 0,
        "ms_per_step": 1e3 * run_s / niter, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "BASELINE configs[1]: ibrm() BayesR, synthetic n=%d x m=%d int8, %d iterations through hb_bayes() "
                               "with host y and X" % (n, m, niter),
                   "n": n, "m": m, "niter": niter, "x_load_s": dg["seconds_setup"], "x_bytes": int(xbytes), "call_s": call_s,
                   "host_generation_s": gen_s, "value_incl_x_load": m * niter / call_s,
                   "at": [around(it) for it in (10, 100, 500, 1000) if it <= niter],
                   "Vg": res["Vg"], "Ve": res["Ve"], "h2": res["h2"], "pi": list(res["pi"])},
        "roofline": {"bound": "hbm", "achieved": n * m / (ms_steady * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": n * m / (ms_steady * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                     "kernel": "k_prep + k_sweep + reductions (device time of an iteration, second half of the chain)",
                     "kernel_ms": ms_steady, "algorithmic_bytes_per_launch": n * m},
        "e2e": {"value": m * niter / run_s, "unit": UNIT, "h2d_bytes_per_step": int(8 * n / niter),
                "d2h_bytes_per_step": int((3 * 8 * m + 16 * n) / niter), "path": "hb_bayes(), timed = the call minus its set-up"},
        "gpu_launches": KERNELS_PER_STEP * niter,
    }
    print(json.dumps(line), flush=True)


def run_c4(args):
    """BASELINE configs[3]: sbrm() SBayesD, dense LD.  The reference's m = 300 000 fp64 matrix is 720 GB; the largest m this
    run takes is --m (default 100 000: 80 GB of fp64 LD in HBM, and on the host while it is handed over).  LD = centred
    X'X / n of a synthetic reference panel (n_ref = 5 000) built by the device LD builder (hb_ldmat_*), summary statistics
    from marginal regressions at N = 50 000.  Reported: LD-column updates per second (one per changed SNP,
    SBayesD.cpp:351-356), LD bytes those updates streamed / device time against the measured HBM peak."""
    import hibayes_b200 as hb
    m = 100000 if args.m == 1000000 else args.m
    n_ref, N = 5000, 50000.0
    niter = 50 if args.niter == 1000 else args.niter
    t0 = time.time()
    X = host_genotypes(n_ref, m, args.seed)
    h = hb.LdMat(X)
    ld = h.dense()
    ld_ms = h.last_ms()
    h.close()
    rng = np.random.default_rng(args.seed)
    beta = np.zeros(m)
    causal = rng.choice(m, size=min(1000, m), replace=False)
    beta[causal] = rng.standard_normal(causal.size)
    vx = np.ascontiguousarray(np.diag(ld))
    gvar = float(beta @ (ld @ beta))
    beta *= math.sqrt(0.5 / gvar)
    keep = vx > 0
    bhat = np.zeros(m)
    bhat[keep] = (ld @ beta)[keep] / vx[keep] + rng.standard_normal(int(keep.sum())) * np.sqrt(1.0 / (N * vx[keep]))
    se = np.ones(m)
    se[keep] = np.sqrt((1.0 - 0.0) / (N * vx[keep]))
    ss = np.asfortranarray(np.column_stack([X.mean(axis=0) / 2, bhat, se, np.full(m, N)]))
    ss[~keep, 1:3] = np.nan
    del X
    t_gen = time.time() - t0
    t0 = time.perf_counter()
    res = hb.SBayesD(ss, ld, "BayesR", PI_R, fold=FOLD, niter=niter, nburn=niter // 2, thin=5, seed=args.seed)
    call_s = time.perf_counter() - t0
    dg = res["diag"]
    sweep_s = dg["seconds_sweep"]
    cols, ent = dg["columns_total"], dg["ld_entries_total"]
    peak, peak_src = _peaks()
    ach = ent * 8 / sweep_s / 1e9
    line = {
        "metric": "sbayesd_ld_column_updates_per_sec", "value": cols / sweep_s, "unit": "LD-column-updates/s", "n_gpus": 1,
        "steps": niter, "warmup": 0, "ms_per_step": 1e3 * sweep_s / niter, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "BASELINE configs[3]: sbrm() SBayesD BayesR, dense fp64 LD m=%d (%.0f GB; the reference's m=300000 "
                               "is 720 GB and fits no single GPU), %d iterations through hb_sbayesd()" % (m, m * m * 8 / 1e9, niter),
                   "m": m, "n_ref": n_ref, "N": N, "precision": "fp64", "ld_bytes_device": dg["ld_bytes_device"],
                   "columns_per_sweep": cols / niter, "snp_updates_per_s": m * niter / sweep_s,
                   "rounds_per_tile": dg["rounds_total"] / max(1, dg["tiles_total"]), "ld_builder_kernel_ms": ld_ms,
                   "setup_s": t_gen, "call_s": call_s, "Vg": res["Vg"], "Ve": res["Ve"], "nnz_last": int(dg["nnz_trace"][-1])},
        "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                     "peak_source": peak_src, "kernel": "k_ld_sweep<dense>", "kernel_ms": 1e3 * sweep_s / niter,
                     "algorithmic_bytes_per_launch": ent * 8 / niter,
                     "note": "algorithmic bytes = m * 8 per changed SNP (SURVEY.md 8d); the tile decisions (one CTA) pace the sweep, "
                             "the column updates of the other CTAs overlap them"},
        "e2e": {"value": cols / (call_s), "unit": "LD-column-updates/s", "h2d_bytes_per_step": int(m * m * 8 / niter),
                "d2h_bytes_per_step": int(4 * 8 * m / niter), "path": "hb_sbayesd() with host sumstat and LD (the LD upload is inside)"},
        "gpu_launches": niter,
    }
    print(json.dumps(line), flush=True)


def random_pedigree_ainv(n_ped, n_founders, rng):
    """A^-1 of a random pedigree by Henderson's rules without inbreeding (what make_Ainv builds, src/rm.cpp:173-206, up
    to its integer-division quirk): animals 0 .. n_founders-1 have no parents, every later animal two random earlier ones."""
    import scipy.sparse as sp
    kids = np.arange(n_founders, n_ped)
    sire = (rng.random(kids.size) * kids).astype(np.int64)
    dam = (rng.random(kids.size) * kids).astype(np.int64)
    dam = np.where(dam == sire, (dam + 1) % kids, dam)
    rows = [np.arange(n_founders)]
    cols = [np.arange(n_founders)]
    vals = [np.ones(n_founders)]
    for a, b, v in ((kids, kids, 2.0), (kids, sire, -1.0), (sire, kids, -1.0), (kids, dam, -1.0), (dam, kids, -1.0),
                    (sire, sire, 0.5), (dam, dam, 0.5), (sire, dam, 0.5), (dam, sire, 0.5)):
        rows.append(a); cols.append(b); vals.append(np.full(kids.size, v))
    A = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n_ped, n_ped))
    return sp.csc_matrix(A)


def run_c5(args):
    """BASELINE configs[4] as far as this build goes: ssbrm()'s call of Bayes() -- BayesCpi with the single-step term (J,
    sparse Gi = A^-1 block of the non-genotyped individuals, epsilon sampler) -- on ONE GPU with INTEGER rows for the
    non-genotyped individuals too (the real-valued imputed rows of R/ssbayes.r:305 are not loadable: DESIGN.md section 8).
    n = 30 000 genotyped + 20 000 non-genotyped individuals with records, pedigree of 200 000 (150 000 non-genotyped:
    qe), m = --m (default 200 000)."""
    import hibayes_b200 as hb
    n_g, n_n, n_ped, qe = 30000, 20000, 200000, 150000
    n = n_g + n_n
    m = 200000 if args.m == 1000000 else args.m
    niter = 60 if args.niter == 1000 else args.niter
    rng = np.random.default_rng(args.seed)
    t0 = time.time()
    X = host_genotypes(n, m, args.seed)
    causal = rng.choice(m, size=min(1000, m), replace=False)
    gv = X[:, causal].astype(np.float64) @ rng.standard_normal(causal.size)
    y = gv * math.sqrt(0.5 / gv.var()) + rng.normal(scale=math.sqrt(0.5), size=n)
    Ainv = random_pedigree_ainv(n_ped, 20000, rng)
    nongeno = np.sort(rng.choice(n_ped, size=qe, replace=False))      # the pedigree members without genotypes
    Gi = Ainv[nongeno][:, nongeno].tocsc()
    index1 = np.sort(rng.choice(qe, size=n_n, replace=False)) + 1      # the ones with a record: rows n_g .. n-1 of y
    J = np.concatenate([-np.ones(n_g), -rng.uniform(0.2, 1.0, n_n)])
    t_gen = time.time() - t0
    t0 = time.perf_counter()
    res = hb.Bayes(y, X, "BayesCpi", [0.95, 0.05], niter=niter, nburn=niter // 2, thin=5, seed=args.seed,
                   epsl_y_J=J, epsl_Gi=Gi, epsl_index=index1)
    call_s = time.perf_counter() - t0
    dg = res["diag"]
    run_s = call_s - dg["seconds_setup"]
    sweep_ms = float(np.mean(dg["sweep_ms_trace"][niter // 2:]))
    peak, peak_src = _peaks()
    line = {
        "metric": METRIC.replace("bayesr_n50k", "ssbayescpi_n50k"), "value": m * niter / run_s, "unit": UNIT, "n_gpus": 1, "steps": niter,
        "warmup": 0, "ms_per_step": 1e3 * run_s / niter, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "BASELINE configs[4] surrogate: Bayes() as ssbrm() calls it, BayesCpi + single-step term, n=%d genotyped + %d "
                               "non-genotyped records (INTEGER rows: the real-valued row class is not built), pedigree %d, qe=%d, m=%d, "
                               "%d iterations through hb_bayes() on 1 GPU" % (n_g, n_n, n_ped, qe, m, niter),
                   "n": n, "m": m, "qe": qe, "ne": n_n, "Gi_nnz": int(Gi.nnz), "device_sweep_ms": sweep_ms,
                   "other_ms_per_iteration": 1e3 * run_s / niter - sweep_ms,
                   "other_is": "J + epsilon step, host draws, PIP / record stores, and the end-of-call summaries (X * alpha, downloads) spread over the iterations", "x_load_s": dg["seconds_setup"], "setup_s": t_gen,
                   "call_s": call_s, "Vg": res["Vg"], "Ve": res["Ve"], "Veps": res["Veps"], "J": res["J"],
                   "rounds_per_tile": float(dg["rounds_total"]) / max(1, dg["tiles_total"])},
        "roofline": {"bound": "hbm", "achieved": n * m / (sweep_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": n * m / (sweep_ms * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                     "kernel": "k_prep + k_sweep + reductions (device time of an iteration's SNP sweep)", "kernel_ms": sweep_ms,
                     "algorithmic_bytes_per_launch": n * m},
        "e2e": {"value": m * niter / run_s, "unit": UNIT, "h2d_bytes_per_step": int(8 * n / niter),
                "d2h_bytes_per_step": int((3 * 8 * m + 16 * n + 8 * qe) / niter), "path": "hb_bayes(), timed = the call minus its set-up"},
        "gpu_launches": (KERNELS_PER_STEP + 9) * niter,
    }
    print(json.dumps(line), flush=True)


def allsum_max(comm, x):
    """max over ranks of a host scalar"""
    if not comm:
        return np.array([x])
    import torch
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    comm.dist.all_reduce(t, op=comm.dist.ReduceOp.MAX)
    return t.cpu().numpy()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--config", default="metric", choices=["metric", "c2", "c3", "c4", "c5"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--n", type=int, default=50000)
    ap.add_argument("--m", type=int, default=1000000)
    ap.add_argument("--niter", type=int, default=1000)
    ap.add_argument("--m-cpu", dest="m_cpu", type=int, default=20000)
    ap.add_argument("--tile", type=int, default=0)
    ap.add_argument("--lag", type=int, default=0)
    ap.add_argument("--seed", type=int, default=20260101)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-product", action="store_true", help="skip the hb_bayes() end-to-end leg (engine-level e2e only)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    elif args.config == "c2":
        run_c2(args)
    elif args.config == "c4":
        run_c4(args)
    elif args.config == "c5":
        run_c5(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
