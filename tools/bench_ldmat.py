"""Device timings of the stages in front of the sweeps: the LD builder's Gram/epilogue kernel and the
.bed -> tile decode.  Not part of bench.py's contract (the metric is the BayesR sweep); prints one JSON line per stage.

    python tools/bench_ldmat.py --n 5000 --m 30000            # synthetic reference panel, dense LD (7.2 GB on the host)
    python tools/bench_ldmat.py --n 5000 --m 30000 --chisq 3.84
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hibayes_b200 as hb  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=5000)
    ap.add_argument("--m", type=int, default=30000)
    ap.add_argument("--chisq", type=float, default=None)
    ap.add_argument("--seed", type=int, default=20260101)
    ap.add_argument("--cpu-m", type=int, default=2000,
                    help="SNPs of the CPU sample: the oracle's literal tXXmat loop (OpenMP) on the first columns, timed beside the device")
    ap.add_argument("--gebv", type=int, default=0, help="records of the GEBV sample product X %%*%% alpha (f4) to time instead (uses --n, --m)")
    a = ap.parse_args()
    if a.gebv:
        # `M %*% MCMCsamples$alpha` (R/bayes.r:303-304): batched fp64 kernel, X read once per 64 records
        e = hb.Engine(a.n, a.m)
        e.synth_geno(a.seed)
        rng = np.random.default_rng(1)
        A = np.asfortranarray(rng.normal(size=(a.m, a.gebv)))
        t0 = time.perf_counter()
        out = e.predict_samples(A)
        wall = time.perf_counter() - t0
        ms = e.last_predict_ms()
        t0 = time.perf_counter()
        for c in range(min(a.gebv, 4)):
            e.predict(A[:, c].copy())
        per_rec = (time.perf_counter() - t0) / min(a.gebv, 4)
        flops = 2.0 * a.n * a.m * a.gebv
        print(json.dumps({"stage": "gebv_samples", "n": a.n, "m": a.m, "records": a.gebv, "kernel_ms": ms, "call_s": wall,
                          "fp64_TFLOPs": flops / (ms * 1e-3) / 1e12, "x_GBs": a.n * a.m * -(-a.gebv // 64) / (ms * 1e-3) / 1e9,
                          "one_record_pass_s": per_rec, "speedup_vs_record_by_record": per_rec * a.gebv / wall,
                          "checksum": float(out.sum())}))
        e.close()
        return
    if a.cpu_m > 0:
        # the reference's CPU path (oracle port of tXXmat_Geno: pairs x individuals scalar loop, OpenMP over SNPs)
        from oracle import hb_oracle
        mc = min(a.cpu_m, a.m)
        Xc = hb.synth_geno_host(a.n, mc, a.seed)
        t0 = time.perf_counter()
        hb_oracle.txxmat(Xc)
        tc = time.perf_counter() - t0
        print(json.dumps({"stage": "ldmat_cpu_port", "n": a.n, "m_sample": mc, "seconds": tc, "cores": os.cpu_count(),
                          "pair_updates_per_s": mc * (mc + 1) / 2 * a.n / tc,
                          "note": "oracle/hb_oracle_ld.c, upper triangle only (as the reference); the device computes both triangles"}))
    X = hb.synth_geno_host(a.n, a.m, a.seed)
    t0 = time.perf_counter()
    h = hb.LdMat(X)
    t_load = time.perf_counter() - t0
    t0 = time.perf_counter()
    h.stats()
    t_stats = time.perf_counter() - t0
    t0 = time.perf_counter()
    if a.chisq is None:
        out = h.dense()
        nnz = a.m * a.m
    else:
        out = h.sparse(chisq=a.chisq)
        nnz = out.nnz
    t_call = time.perf_counter() - t0
    ms = h.last_ms()
    ops = 2.0 * a.n * a.m * a.m            # int8 multiply-adds of the full (both triangles) Gram, as computed
    print(json.dumps({
        "stage": "ldmat", "n": a.n, "m": a.m, "chisq": a.chisq, "kernel_ms": ms, "int8_TOPS": ops / (ms * 1e-3) / 1e12,
        "fp64_out_GBs": a.m * a.m * 8 / (ms * 1e-3) / 1e9, "call_s": t_call, "load_s": t_load, "stats_s": t_stats,
        "stored_entries": int(nnz), "note": "kernel_ms = sum of k_ld_panel launches (CUDA events on the handle's stream)"}))
    h.close()
    # .bed decode into engine tiles: build an image of the same genotypes (no missing values), time the load
    nid, m = a.n, min(a.m, 20000)
    code = np.array([3, 2, 0], dtype=np.uint8)
    f = code[X[:, :m]]
    bps = (nid + 3) // 4
    body = np.zeros((m, bps), dtype=np.uint8)
    for x in range(4):
        part = f[x::4, :].T
        body[:, :part.shape[1]] |= part << (2 * x)
    img = np.concatenate([np.array([0x6C, 0x1B, 0x01], dtype=np.uint8), body.reshape(-1)])
    e = hb.Engine(nid, m)
    t0 = time.perf_counter()
    e.load_geno(hb.BedGeno(img, nid, m))
    t_bed = time.perf_counter() - t0
    t0 = time.perf_counter()
    e.load_geno(X[:, :m])
    t_i8 = time.perf_counter() - t0
    print(json.dumps({"stage": "bed_to_tiles", "nid": nid, "m": m, "bed_load_s": t_bed, "int8_load_s": t_i8,
                      "bed_bytes": int(img.shape[0]), "int8_bytes": int(nid) * m,
                      "note": "host wall time incl. the PCIe copy of the source (image vs int8 matrix)"}))
    e.close()


if __name__ == "__main__":
    main()
