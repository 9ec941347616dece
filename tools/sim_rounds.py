"""CPU study of the scalar chain's speculation rounds -- no GPU, no product code.

Runs a BayesR Gibbs chain (Bayes.cpp:743-815) in numpy on synthetic genotypes with the bench's per-SNP regime
(allele frequencies U(0.05,0.5), 0.1 % causal SNPs, the remaining genetic variance of the 1M-SNP workload folded
into the noise) and, for every tile of B consecutive SNPs of every sweep, replays what phase S of the sweep kernel
does with the tile: classes speculated from the right-hand sides at tile entry, candidates chained under the
speculated classes, every SNP re-classified with its exact right-hand side, repeat while any class differs.

It reports, per sweep: candidates per tile, rounds per tile of the CURRENT scheme, and of two alternatives:
  cond      after a first round that missed, ONE in-order chain over the candidates in which every lane decides
            its class at its turn from its own exact right-hand side (no speculation for candidates), followed by
            the usual verification of the non-candidates;
  cond+near the same with the non-candidates whose rhs^2 is within `--near` (relative) of a class boundary added
            to the chain as conditional lanes.
and classifies the misses of the first round (candidate changed class / non-candidate became non-zero).

    python tools/sim_rounds.py --n 50000 --m 5120 --sweeps 40
    python tools/sim_rounds.py --n 50000 --m 5120 --sweeps 40 --fold-scale 64   # HB_BENCH_FOLD_SCALE=64 of bench.py
"""
import argparse
import time

import numpy as np

FOLD = np.array([0.0, 1e-4, 1e-3, 1e-2])


def classify(rhs, xpx, u, vare, vara_fold, logpi):
    """vectorised Bayes.cpp:759-781: class per SNP from rhs, with the SNP's uniform u"""
    F = len(logpi)
    s = np.empty((F,) + rhs.shape)
    s[0] = logpi[0]
    for k in range(1, F):
        s[k] = -0.5 * (np.log(vara_fold[k] * xpx / vare + 1) - rhs * (rhs / (xpx + vare / vara_fold[k])) / vare) + logpi[k]
    p = np.exp(s - s.max(axis=0))
    p /= p.sum(axis=0)
    cum = np.cumsum(p, axis=0)
    cls = (u[None, :] >= cum).sum(axis=0)
    cls[cls >= F] = 0  # falls to 0 if the cumulative sum rounds below u (:773-781)
    return cls


def boundary_distance(rhs, xpx, u, vare, vara_fold, logpi):
    """relative distance in rhs^2 to the nearest point where the class changes (bisection-free estimate: the class
    at rhs^2 (1 +- eps) for a ladder of eps)"""
    base = classify(rhs, xpx, u, vare, vara_fold, logpi)
    dist = np.full(rhs.shape, np.inf)
    for eps in (0.02, 0.05, 0.1, 0.2, 0.3, 0.5, 1.0, 2.0):
        for sgn in (+1, -1):
            f = 1 + sgn * eps
            if f <= 0:
                continue
            c2 = classify(rhs * np.sqrt(f), xpx, u, vare, vara_fold, logpi)
            hit = (c2 != base) & (dist == np.inf)
            dist[hit] = eps
    return dist


def draw(rhs, cls, xpx, z, vare, vara_fold):
    g = np.zeros_like(rhs)
    nz = cls > 0
    v = xpx[nz] + vare / vara_fold[cls[nz]]
    g[nz] = rhs[nz] / v + np.sqrt(vare / v) * z[nz]
    return g


def exact_tile(entry, Gl, gold, xpx, u, z, vare, vara_fold, logpi):
    """the literal one-SNP-at-a-time pass over a tile; entry = x'r at tile entry (without xpx*g)"""
    B = len(entry)
    delta = np.zeros(B)
    cls = np.zeros(B, dtype=int)
    gnew = np.zeros(B)
    for i in range(B):
        if xpx[i] == 0:
            gnew[i] = gold[i]
            continue
        rhs = entry[i] - Gl[i, :i] @ delta[:i] + xpx[i] * gold[i]
        c = classify(np.array([rhs]), xpx[i:i + 1], u[i:i + 1], vare, vara_fold, logpi)[0]
        cls[i] = c
        gnew[i] = draw(np.array([rhs]), np.array([c]), xpx[i:i + 1], z[i:i + 1], vare, vara_fold)[0]
        delta[i] = gnew[i] - gold[i]
    return cls, gnew, delta


def chain_fixed(entry, Gl, gold, xpx, z, vare, vara_fold, spec, cand):
    """candidates chained in order under the speculated classes (what chain_matvec / chain_candidates compute)"""
    delta = np.zeros(len(entry))
    idx = np.flatnonzero(cand)
    for i in idx:
        rhs = entry[i] - Gl[i, idx[idx < i]] @ delta[idx[idx < i]] + xpx[i] * gold[i]
        g = 0.0
        if spec[i] > 0:
            v = xpx[i] + vare / vara_fold[spec[i]]
            g = rhs / v + np.sqrt(vare / v) * z[i]
        delta[i] = g - gold[i]
    return delta


def chain_conditional(entry, Gl, gold, xpx, u, z, vare, vara_fold, logpi, lanes):
    """in-order chain in which every lane decides its own class from its exact right-hand side"""
    delta = np.zeros(len(entry))
    cls = np.zeros(len(entry), dtype=int)
    idx = np.flatnonzero(lanes)
    for i in idx:
        rhs = entry[i] - Gl[i, idx[idx < i]] @ delta[idx[idx < i]] + xpx[i] * gold[i]
        c = classify(np.array([rhs]), xpx[i:i + 1], u[i:i + 1], vare, vara_fold, logpi)[0]
        cls[i] = c
        g = 0.0
        if c > 0:
            v = xpx[i] + vare / vara_fold[c]
            g = rhs / v + np.sqrt(vare / v) * z[i]
        delta[i] = g - gold[i]
    return delta, cls


def verify(entry, Gl, gold, xpx, u, delta, vare, vara_fold, logpi):
    rhs = entry - Gl @ delta + xpx * gold   # Gl strictly lower: only earlier SNPs count
    cls2 = classify(rhs, np.where(xpx > 0, xpx, 1.0), u, vare, vara_fold, logpi)
    cls2[xpx == 0] = 0
    return cls2, rhs


def replay(entry, Gl, gold, xpx, u, z, vare, vara_fold, logpi, truth_cls, near):
    act = xpx > 0
    spec = classify(entry + xpx * gold, np.where(act, xpx, 1.0), u, vare, vara_fold, logpi)
    spec[~act] = 0
    out = {}
    # ---- current scheme
    s, rounds = spec.copy(), 0
    first_miss = None
    while True:
        rounds += 1
        cand = act & ((gold != 0) | (s != 0))
        delta = chain_fixed(entry, Gl, gold, xpx, z, vare, vara_fold, s, cand)
        cls2, rhs = verify(entry, Gl, gold, xpx, u, delta, vare, vara_fold, logpi)
        bad = act & (cls2 != s)
        if rounds == 1:
            first_miss = (bad & cand).sum(), (bad & ~cand).sum()
            k1 = int(cand.sum())
        if not bad.any() or rounds > 50:
            break
        s = cls2
    assert np.array_equal(s[act], truth_cls[act]), "replay of the current scheme does not reach the literal result"
    out["cur"] = rounds
    out["k"] = k1
    out["miss_cand"], out["miss_non"] = first_miss
    # ---- conditional chain after a missed first round
    for tag, use_near in (("cond", False), ("cond_near", True)):
        if out["cur"] == 1:
            out[tag] = 1
            out[tag + "_lanes"] = k1
            continue
        s, rounds = spec.copy(), 1
        cand = act & ((gold != 0) | (s != 0))
        delta = chain_fixed(entry, Gl, gold, xpx, z, vare, vara_fold, s, cand)
        cls2, rhs = verify(entry, Gl, gold, xpx, u, delta, vare, vara_fold, logpi)
        s = cls2
        lanes_max = 0
        while True:
            rounds += 1
            lanes = act & ((gold != 0) | (s != 0))
            if use_near:
                dist = boundary_distance(rhs, np.where(act, xpx, 1.0), u, vare, vara_fold, logpi)
                lanes |= act & (dist <= near)
            lanes_max = max(lanes_max, int(lanes.sum()))
            delta, cl = chain_conditional(entry, Gl, gold, xpx, u, z, vare, vara_fold, logpi, lanes)
            s = np.where(lanes, cl, s)
            cls2, rhs = verify(entry, Gl, gold, xpx, u, delta, vare, vara_fold, logpi)
            bad = act & (cls2 != s)
            if not bad.any() or rounds > 50:
                break
            s = cls2
        assert np.array_equal(s[act], truth_cls[act]), tag
        out[tag] = rounds
        out[tag + "_lanes"] = lanes_max
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=50000)
    ap.add_argument("--m", type=int, default=5120)
    ap.add_argument("--tile", type=int, default=256)
    ap.add_argument("--sweeps", type=int, default=40)
    ap.add_argument("--near", type=float, default=0.3)
    ap.add_argument("--m-full", type=int, default=1000000, help="SNP count of the workload whose regime is imitated")
    ap.add_argument("--seed", type=int, default=20260101)
    ap.add_argument("--report-every", type=int, default=5)
    ap.add_argument("--leads", type=str, default="",
                    help="comma list of lead times in tiles, e.g. 0,1,2,4: how complete must the gathered candidate rows be "
                         "when phase P runs that many tiles before phase S (DESIGN.md section 10, one serial CTA)")
    ap.add_argument("--nears", type=str, default="0,0.1,0.3,0.5,1.0")
    ap.add_argument("--fold-scale", type=float, default=1.0,
                    help="multiplies the variance folds like HB_BENCH_FOLD_SCALE of bench.py (64: the flip-heavy regime)")
    a = ap.parse_args()
    global FOLD
    FOLD = FOLD * a.fold_scale
    rng = np.random.default_rng(a.seed)
    n, m, B = a.n, a.m, a.tile
    T = m // B
    assert T * B == m
    t0 = time.time()
    p = rng.uniform(0.05, 0.5, size=m)
    X = np.empty((n, m), dtype=np.int8, order="F")
    for j in range(m):
        r_ = rng.random(n, dtype=np.float32)
        q2 = (1 - p[j]) ** 2
        q1 = q2 + 2 * p[j] * (1 - p[j])
        X[:, j] = (r_ >= q2).astype(np.int8) + (r_ >= q1).astype(np.int8)
    # causal effects with the bench's per-SNP size: 1000 causal SNPs of m_full carry var 0.5
    n_c = max(1, round(m * 1000 / a.m_full))
    idx = rng.choice(m, size=n_c, replace=False)
    vx = X[:, idx].astype(np.float64).var(axis=0)
    b = rng.normal(size=n_c) * np.sqrt(0.5 / 1000 / vx.mean())
    gv = X[:, idx].astype(np.float64) @ b
    y = gv + rng.normal(scale=np.sqrt(1.0 - gv.var()), size=n)
    print("data: n=%d m=%d causal=%d var(gv)=%.4g  (%.0f s)" % (n, m, n_c, gv.var(), time.time() - t0), flush=True)
    # exact Gram blocks of the tiles (integers; float32 sums of integers < 2^24 are exact)
    t0 = time.time()
    xpx = np.zeros(m)
    Gl = []
    for t in range(T):
        Xt = X[:, t * B:(t + 1) * B].astype(np.float32)
        G = (Xt.T @ Xt).astype(np.float64)
        xpx[t * B:(t + 1) * B] = np.diag(G)
        Gl.append(np.tril(G, -1))
    print("gram: %.0f s" % (time.time() - t0), flush=True)
    vxall = xpx / n - (X.astype(np.float32).sum(axis=0).astype(np.float64) / n) ** 2 if n * m < 4e8 else None
    # priors as Bayes.cpp:319-355 for the m_full workload
    pi = np.array([0.95, 0.02, 0.02, 0.01])
    vary = y.var(ddof=1)
    vara_ = 0.5 * vary * 0.5
    vare = vary * 0.5
    sumvx_full = 2 * (p * (1 - p)).mean() * a.m_full
    varg = vara_ / ((1 - pi[0]) * sumvx_full)
    dfg, s2g = 4.0, varg * 0.5
    g = np.zeros(m)
    r = y - y.mean()
    leads = [int(v) for v in a.leads.split(",") if v != ""]
    nears = [float(v) for v in a.nears.split(",") if v != ""]
    lead_stats = {(ld, nr): [0, 0, 0.0, 0] for ld in leads for nr in nears}   # tiles, tiles with a miss, rows, missed SNPs
    hist = []   # residual at the entry of the last max(leads) tiles
    for it in range(a.sweeps):
        ts = time.time()
        r -= rng.normal(r.mean(), np.sqrt(vare / n))
        vara_fold = varg * FOLD
        vara_fold[0] = 1.0   # unused
        logpi = np.log(pi)
        stats = []
        counts = np.zeros(4)
        varg_acc = 0.0
        for t in range(T):
            cols = slice(t * B, (t + 1) * B)
            Xt = X[:, cols].astype(np.float32)
            entry = (Xt.T @ r.astype(np.float32)).astype(np.float64)
            # float32 products of {0,1,2} with r are exact enough for statistics; make the right-hand sides exact
            # where it matters by recomputing in fp64 for this tile
            entry = Xt.T.astype(np.float64) @ r if n <= 60000 else entry
            u = rng.random(B)
            z = rng.normal(size=B)
            gold = g[cols].copy()
            cls, gnew, delta = exact_tile(entry, Gl[t], gold, xpx[cols], u, z, vare, vara_fold, logpi)
            if leads:
                hist.append(r.copy())
                if len(hist) > max(leads) + 1:
                    hist.pop(0)
                if it >= 10:
                    act = xpx[cols] > 0
                    xs = np.where(act, xpx[cols], 1.0)
                    needed = act & ((gold != 0) | (cls != 0))
                    Xt64 = Xt.T.astype(np.float64)
                    for ld in leads:
                        r_old = hist[max(0, len(hist) - 1 - ld)]
                        rhs_old = Xt64 @ r_old + xpx[cols] * gold
                        spec_old = classify(rhs_old, xs, u, vare, vara_fold, logpi)
                        dist = boundary_distance(rhs_old, xs, u, vare, vara_fold, logpi)
                        for nr in nears:
                            pkg = act & ((gold != 0) | (spec_old != 0) | (dist <= nr))
                            st = lead_stats[(ld, nr)]
                            miss = needed & ~pkg
                            st[0] += 1
                            st[1] += int(miss.any())
                            st[2] += int(pkg.sum())
                            st[3] += int(miss.sum())
            stats.append(replay(entry, Gl[t], gold, xpx[cols], u, z, vare, vara_fold, logpi, cls, a.near))
            ch = np.flatnonzero(delta != 0)
            if ch.size:
                r -= Xt[:, ch].astype(np.float64) @ delta[ch]
            g[cols] = gnew
            for k in range(4):
                counts[k] += (cls == k).sum()
            nzm = cls > 0
            varg_acc += (gnew[nzm] ** 2 / FOLD[cls[nzm]]).sum()
        nnz = m - counts[0]
        varg = (varg_acc + s2g * dfg) / rng.chisquare(dfg + nnz)
        pi = rng.dirichlet(counts + 1)
        vare = (r @ r) / rng.chisquare(n - 2)
        if (it + 1) % a.report_every == 0 or it == 0:
            def mean(key):
                return float(np.mean([s[key] for s in stats]))
            print("sweep %3d  nnz %5d (%.1f%%)  cand/tile %.1f  rounds/tile: current %.3f  cond %.3f (lanes<=%d)  cond+near %.3f (lanes<=%d)"
                  "  first-round misses/tile: candidates %.2f, non-candidates %.2f   vare %.4f varg %.3g  [%.0f s]"
                  % (it + 1, nnz, 100 * nnz / m, mean("k"), mean("cur"), mean("cond"),
                     max(s["cond_lanes"] for s in stats), mean("cond_near"), max(s["cond_near_lanes"] for s in stats),
                     mean("miss_cand"), mean("miss_non"), vare, varg, time.time() - ts), flush=True)
    if leads:
        print("package completeness when phase P runs `lead` tiles before phase S (tiles of sweeps > 10):")
        print("  lead  near   tiles  tiles with a missing row  missing rows/tile  rows gathered/tile")
        for (ld, nr), st in sorted(lead_stats.items()):
            if st[0]:
                print("  %4d  %4.2f  %6d  %8.2f %%               %8.4f           %6.1f"
                      % (ld, nr, st[0], 100.0 * st[1] / st[0], st[3] / st[0], st[2] / st[0]))


if __name__ == "__main__":
    main()
