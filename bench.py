#!/usr/bin/env python
"""bench.py -- Gibbs SNP-updates/s of the BayesR sweep (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one MCMC iteration's SNP sweep over the whole synthetic genotype matrix
(n=50 000 x m=1 000 000 int8 per SURVEY.md 8d, generated on the device) followed by the
per-iteration reductions -- what replaces Bayes.cpp:586-823.  `value` times K steps on the device
(CUDA events on the engine's stream, inputs resident in HBM); `e2e` times the same K steps through
the host-facing C ABI with the residual crossing PCIe in both directions every step.
`--impl reference` times the reference's own CPU data path (per-SNP ddot + 2 daxpy on a column-major
fp64 matrix, Bayes.cpp:751-802) with all host threads on a bounded column sample of the workload.
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


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


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
            self.stop.wait(0.2)

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
    """Host-side scalar updates of the BayesR chain around hb_engine_sweep (Bayes.cpp:480-482,
    803-823); numpy's generator stands in for the host draws -- only the workload matters here."""

    def __init__(self, n, m_active, vary, sumvx, seed):
        self.n, self.m_active = n, m_active
        self.rng = np.random.default_rng(seed)
        self.df = 4.0
        vara = (self.df - 2) / self.df * vary * 0.5
        self.vare = vary * 0.5
        self.varg = vara / ((1 - PI0[0]) * sumvx)
        self.s2varg = vara * (self.df - 2) / self.df / ((1 - PI0[0]) * sumvx)
        self.pi = np.array(PI0)
        self.sum_r, self.sum_r2 = 0.0, vary * (n - 1)
        self.it = 0

    def sweep_args(self):
        mu_ = -(self.sum_r / self.n + math.sqrt(self.vare / self.n) * self.rng.standard_normal())
        rn2 = self.sum_r2 + 2 * mu_ * self.sum_r + self.n * mu_ * mu_
        return dict(iter=self.it, model_index=6, vare=self.vare, logpi=list(np.log(self.pi)),
                    vara_fold=[self.varg * f for f in FOLD], fold=FOLD, dfvara=self.df, s2varg=self.s2varg,
                    mu_shift=mu_, rnorm2_bound=rn2)

    def update(self, so):
        cnt = np.array(so["count"][:4])
        nnz = cnt[1:].sum()
        self.varg = (so["varg_acc"] + self.s2varg * self.df) / self.rng.chisquare(self.df + nnz)
        self.pi = self.rng.dirichlet(cnt + 1)
        self.vare = so["sum_r2"] / self.rng.chisquare(self.n - 2)
        self.sum_r, self.sum_r2 = so["sum_r"], so["sum_r2"]
        self.it += 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n, m_cpu = args.n, args.m_cpu
    # each "step" = one sweep over the m_cpu-column sample; hbo_time_sweep_fp64 runs one untimed
    # warm-up sweep itself, then `steps` timed sweeps
    t0 = time.time()
    val = cpu_reference(n, m_cpu, max(1, args.steps), threads)
    wall = time.time() - t0
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * m_cpu / val, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "BayesR sweep n=%d x m=1000000 (CPU sample: first %d columns, fp64 column-major X)" % (n, m_cpu),
                   "n": n, "m_sample": m_cpu},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "n=%d x m=%d fp64, %d sweeps, OpenMP ddot/daxpy (reference not buildable: needs R/Rcpp)"
                                   % (n, m_cpu, max(1, args.steps))},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": wall,
    }
    print(json.dumps(line), flush=True)


def run_gpu(args):
    import hibayes_b200 as hb
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl")
    n_total, m = args.n, args.m
    # rows are sharded across ranks (weak scaling: n rows per GPU)
    n_local = n_total
    t_setup = time.time()
    eng = hb.Engine(n_local, m, device=local_rank, tile_snps=args.tile, lag_tiles=args.lag, seed=args.seed, rank=rank, world=world)
    eng.synth_geno(args.seed, row_offset=rank * n_local)
    xpx, sumx = eng.col_stats()
    if world > 1:
        raise NotImplementedError("multi-GPU row sharding lands with the NVLink exchange (DESIGN.md)")
    vx = (xpx - sumx * sumx / n_local) / (n_local - 1)
    active = (n_local * xpx != sumx * sumx)
    eng.set_snp_info(xpx, active.astype(np.uint8))
    t_gram = time.time()
    eng.build_gram()
    t_gram = time.time() - t_gram
    rng = np.random.default_rng(args.seed)
    beta = np.zeros(m)
    causal = rng.choice(m, size=min(1000, m), replace=False)
    beta[causal] = rng.standard_normal(causal.size)
    gv = eng.predict(beta)
    gv *= math.sqrt(0.5 / gv.var())
    y = gv + rng.normal(scale=math.sqrt(0.5), size=n_local)
    r = y - y.mean()
    eng.set_residual(r)
    chain = Chain(n_local, int(active.sum()), float(y.var(ddof=1)), float(vx.sum()), args.seed)
    chain.sum_r, chain.sum_r2 = float(r.sum()), float(r @ r)
    t_setup = time.time() - t_setup
    desc = eng.describe()

    def step():
        so = eng.sweep(**chain.sweep_args())
        chain.update(so)
        return so

    for _ in range(args.warmup):
        step()
    # ---- device-timed region: K steps, inputs resident in HBM
    dev_ms, sweep_ms, changed, rounds = [], [], [], []
    with ClockSampler(local_rank) as clk:
        for _ in range(args.steps):
            so = step()
            a, b, c = eng.last_sweep_ms()
            dev_ms.append(a + b + c)
            sweep_ms.append(b)
            changed.append(so["n_changed"])
            rounds.append(so["rounds"])
    total_ms = float(sum(dev_ms))
    m_active = int(active.sum())
    value = m_active * args.steps / (total_ms * 1e-3)
    # ---- end-to-end region: same steps through the host-facing ABI, residual over PCIe each step
    t0 = time.perf_counter()
    for _ in range(args.steps):
        eng.set_residual(r)                 # H2D: n doubles (what the host driver does after non-SNP effects)
        so = step()
        r = eng.get_residual()              # D2H: n doubles
        _ = eng.get_u()                     # D2H: n doubles
    e2e_s = time.perf_counter() - t0
    e2e = m_active * args.steps / e2e_s
    peak, peak_src = _peaks()
    kern_s = float(np.mean(sweep_ms)) * 1e-3
    achieved = n_local * m / kern_s / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "ibrm() BayesR sweep, synthetic n=%d x m=%d int8 genotypes, 1 GPU" % (n_total, m),
                   "n": n_total, "m": m, "m_active": m_active, "Pi": PI0, "fold": FOLD,
                   "layout": desc, "l2": "inputs (%.1f GB) larger than L2" % (desc["geno_bytes"] / 1e9),
                   "changed_snps_per_sweep": float(np.mean(changed)), "scalar_rounds_per_sweep": float(np.mean(rounds)), "setup_s": t_setup, "gram_s": t_gram,
                   "frac_of_8TBps": achieved / 8000.0},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "peak_source": peak_src, "kernel": "k_sweep", "kernel_ms": kern_s * 1e3},
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": 8 * n_local, "d2h_bytes_per_step": 16 * n_local + 128},
        "gpu_launches": 3 * args.steps,
        "clocks": clk.summary(),
    }
    if rank == 0 and not args.no_cpu:
        threads = os.cpu_count() or 1
        val = cpu_reference(n_total, args.m_cpu, 1, threads)
        line["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": "n=%d x m=%d fp64 column-major, 1 warm + 1 timed sweep" % (n_total, args.m_cpu)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    eng.close()


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
