#!/usr/bin/env python
"""bench.py -- Gibbs SNP-updates/s of the BayesR sweep (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one MCMC iteration's SNP sweep over the whole synthetic genotype matrix (n = 50 000
individuals x m = 1 000 000 SNPs, int8, generated on the device) followed by the per-iteration reductions -- what
replaces Bayes.cpp:586-823 -- driven by the host-side scalar updates of a BayesR chain.

`value`     K steps timed with inputs resident in HBM (barrier + synchronize on both sides, max over ranks).
`e2e`       the same K steps through the host-facing C ABI with HOST buffers: every step uploads the residual
            (what the host driver does after its non-SNP effects) and downloads residual and genetic values.
N > 1       weak scaling over individuals: every rank holds n = 50 000 rows of an N x 50 000-row matrix (the
            reference's C3 shape is 8 x 25 000); a unit of `value` is one SNP update over one rank's 50 000-row
            shard, so N ranks sweeping m SNPs do N x m units per step.  The dots of a tile are exchanged inside the
            sweep kernel (NVLink peer atomics), scalars of the iteration by one small all-reduce.
`--impl reference`  the reference's own CPU data path (per-SNP ddot + 2 daxpy on a column-major fp64 matrix,
            Bayes.cpp:751-802, all host threads) on a bounded column sample of the same workload (oracle port: the
            reference itself needs R/Rcpp/Armadillo and cannot be built here).
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
PI0 = [0.95, 0.02, 0.02, 0.01]
FOLD = [0.0, 1e-4, 1e-3, 1e-2]
if os.environ.get("HB_BENCH_FOLD_SCALE"):   # experiments only: stronger effects stress the speculation of the scalar chain
    FOLD = [f * float(os.environ["HB_BENCH_FOLD_SCALE"]) for f in FOLD]
KERNELS_PER_STEP = 5   # k_prep, k_sweep, k_post1, k_post2, k_tail


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json, copy bandwidth)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def _traffic():
    """dram bytes per k_sweep launch at the bench shape, from the committed ncu capture (profiles/)."""
    path = os.path.join(ROOT, "profiles", "r01_sweep_traffic.json")
    try:
        return json.load(open(path))
    except Exception:
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


def cpu_reference(n, m_cpu, sweeps, threads):
    from oracle import hb_oracle
    hb_oracle.lib()
    val, _ = hb_oracle.time_sweep_fp64(n, m_cpu, sweeps, threads)
    return val


class Chain:
    """Host-side scalar updates of the BayesR chain around hb_engine_sweep (Bayes.cpp:480-482, 803-823); numpy's
    generator stands in for the host draws (same seed on every rank) -- only the workload matters here."""

    def __init__(self, n_total, vary, sumvx, seed):
        self.n = n_total
        self.rng = np.random.default_rng(seed)
        self.df = 4.0
        vara = (self.df - 2) / self.df * vary * 0.5
        self.vare = vary * 0.5
        self.varg = vara / ((1 - PI0[0]) * sumvx)
        self.s2varg = vara * (self.df - 2) / self.df / ((1 - PI0[0]) * sumvx)
        self.pi = np.array(PI0)
        self.sum_r, self.sum_r2 = 0.0, vary * (n_total - 1)
        self.it = 0

    def sweep_args(self):
        mu_ = -(self.sum_r / self.n + math.sqrt(self.vare / self.n) * self.rng.standard_normal())
        rn2 = self.sum_r2 + 2 * mu_ * self.sum_r + self.n * mu_ * mu_
        return dict(iter=self.it, model_index=6, vare=self.vare, logpi=list(np.log(self.pi)),
                    vara_fold=[self.varg * f for f in FOLD], fold=FOLD, dfvara=self.df, s2varg=self.s2varg,
                    mu_shift=mu_, rnorm2_bound=rn2)

    def update(self, so, sum_r, sum_r2):
        cnt = np.array(so["count"][:4])
        nnz = cnt[1:].sum()
        self.varg = (so["varg_acc"] + self.s2varg * self.df) / self.rng.chisquare(self.df + nnz)
        self.pi = self.rng.dirichlet(cnt + 1)
        self.vare = sum_r2 / self.rng.chisquare(self.n - 2)
        self.sum_r, self.sum_r2 = sum_r, sum_r2
        self.it += 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n, m_cpu = args.n, args.m_cpu
    # each "step" = one sweep over the m_cpu-column sample; hbo_time_sweep_fp64 runs one untimed warm-up sweep
    # itself, then `steps` timed sweeps
    t0 = time.time()
    val = cpu_reference(n, m_cpu, max(1, args.steps), threads)
    wall = time.time() - t0
    sample = ("n=%d x m=%d fp64 column-major, %d timed sweeps, OpenMP ddot/daxpy (oracle port; the reference needs "
              "R/Rcpp/Armadillo and cannot be built here)" % (n, m_cpu, max(1, args.steps)))
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * m_cpu / val, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "ibrm() BayesR sweep, synthetic n=%d x m=%d (CPU sample: first %d columns)" % (n, args.m, m_cpu),
                   "n": n, "m": args.m, "m_sample": m_cpu},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": wall,
    }
    print(json.dumps(line), flush=True)


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

    n_local, m = args.n, args.m
    n_total = n_local * world
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
    chain = Chain(n_total, float(vary), float(vx.sum()), args.seed)
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
    value = world * m_active * args.steps / wall
    # ---- end-to-end region: same steps through the host-facing ABI, residual over PCIe in both directions
    r = eng.get_residual()                  # the chain's current residual (not the one from before the sweeps)
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
    e2e = world * m_active * args.steps / e2e_s
    peak, peak_src = _peaks()
    kern_s = float(np.mean(sweep_ms)) * 1e-3
    achieved = n_local * m / kern_s / 1e9   # algorithmic bytes: n per SNP update (one read of the int8 column)
    tr = _traffic()
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "ibrm() BayesR sweep, synthetic int8 genotypes, n=%d individuals per GPU x m=%d SNPs, %d GPU(s)"
                               % (n_local, m, world),
                   "n_per_gpu": n_local, "n_total": n_total, "m": m, "m_active": m_active, "Pi": PI0, "fold": FOLD,
                   "unit_of_value": "one SNP update over one rank's %d-row shard" % n_local,
                   "layout": desc, "l2": "inputs (%.1f GB per GPU) larger than L2" % (desc["geno_bytes"] / 1e9),
                   "changed_snps_per_sweep": float(np.mean(changed)), "scalar_rounds_per_sweep": float(np.mean(rounds)),
                   "device_ms_per_step": float(np.mean(dev_ms)), "setup_s": t_setup, "gram_s": t_gram,
                   "frac_of_8TBps": achieved / 8000.0},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": (tr or {}).get("dram_bytes_per_launch"), "traffic_source": (tr or {}).get("source"),
                     "peak_source": peak_src, "kernel": "k_sweep", "kernel_ms": kern_s * 1e3,
                     "algorithmic_bytes_per_launch": n_local * m},
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": 8 * n_local, "d2h_bytes_per_step": 16 * n_local + 160},
        "gpu_launches": KERNELS_PER_STEP * args.steps,
        "clocks": clk.summary(),
    }
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        val = cpu_reference(n_local, args.m_cpu, 1, threads)
        line["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": "n=%d x m=%d fp64 column-major (first columns of the workload), 1 warm + 1 timed sweep, "
                                          "OpenMP ddot/daxpy" % (n_local, args.m_cpu)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    eng.close()
    if comm:
        comm.dist.barrier()
        comm.dist.destroy_process_group()


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
    ap.add_argument("--n", type=int, default=50000)
    ap.add_argument("--m", type=int, default=1000000)
    ap.add_argument("--m-cpu", dest="m_cpu", type=int, default=20000)
    ap.add_argument("--tile", type=int, default=0)
    ap.add_argument("--lag", type=int, default=0)
    ap.add_argument("--seed", type=int, default=20260101)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
