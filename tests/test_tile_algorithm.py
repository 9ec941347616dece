"""CPU check of the algorithm the sweep kernel implements (DESIGN.md "Chaining the tiles"):
dots taken against a residual that is D tiles stale + exact Gram-band corrections + speculative
commit rounds must reproduce the literal one-SNP-at-a-time sweep of Bayes.cpp:751-802.
Pure numpy emulation -- no GPU, no product code; it guards the maths the CUDA code transcribes."""
import ctypes as C

import numpy as np
import pytest

from tests.util_demo import synth


def _draws(oracle, seed, it, m):
    L = oracle.lib()
    u, z = C.c_double(), C.c_double()
    us, zs = np.empty(m), np.empty(m)
    for j in range(m):
        L.hbo_draw_uz(seed, 1, it, j, 1, 0, C.byref(u), C.byref(z))
        us[j], zs[j] = u.value, z.value
    return us, zs


def _eval(rhs, xx, vare, vara_fold, logpi, u, z):
    F = len(logpi)
    s = np.empty(F)
    s[0] = logpi[0]
    for k in range(1, F):
        s[k] = -0.5 * (np.log(vara_fold[k] * xx / vare + 1) - rhs * (rhs / (xx + vare / vara_fold[k])) / vare) + logpi[k]
    p = np.exp(s - s.max())
    p /= p.sum()
    acc, cls = 0.0, 0
    for k in range(F):
        acc += p[k]
        if u < acc:
            cls = k
            break
    if cls == 0:
        return 0, 0.0
    v = xx + vare / vara_fold[cls]
    return cls, rhs / v + np.sqrt(vare / v) * z


def literal_sweep(X, r, g, xpx, vare, vara_fold, logpi, us, zs):
    r, g = r.copy(), g.copy()
    cls_out = np.zeros(len(g), dtype=int)
    for j in range(X.shape[1]):
        if xpx[j] == 0:
            continue
        rhs = X[:, j] @ r + (xpx[j] * g[j] if g[j] != 0 else 0.0)
        cls, gn = _eval(rhs, xpx[j], vare, vara_fold, logpi, us[j], zs[j])
        if gn != g[j]:
            r -= X[:, j] * (gn - g[j])
        g[j], cls_out[j] = gn, cls
    return r, g, cls_out


def tiled_sweep(X, r, g, xpx, vare, vara_fold, logpi, us, zs, B, D):
    n, m = X.shape
    T = (m + B - 1) // B
    Xp = np.zeros((n, T * B))
    Xp[:, :m] = X
    xp = np.zeros(T * B); xp[:m] = xpx
    gp = np.zeros(T * B); gp[:m] = g
    up = np.full(T * B, 0.5); up[:m] = us
    zp = np.zeros(T * B); zp[:m] = zs
    r_stream = r.copy()
    ring = np.zeros((D, B))
    queue, tile_qend, applied = [], [], 0
    cls_out = np.zeros(T * B, dtype=int)
    for t in range(T + D):
        if t >= D:  # streaming CTAs apply what the scalar CTA published for tiles <= t-D
            qend = tile_qend[t - D]
            for (j, dl) in queue[applied:qend]:
                r_stream -= Xp[:, j] * dl
            applied = qend
        if t >= T:
            continue
        cols = slice(t * B, (t + 1) * B)
        d = Xp[:, cols].T @ r_stream
        d -= ring[t % D]
        ring[t % D] = 0
        gold = gp[cols].copy()
        rhs = d + np.where(gold != 0, xp[cols] * gold, 0.0)
        gnew, cls = gold.copy(), np.zeros(B, dtype=int)
        done = xp[cols] == 0
        chg = []
        while True:
            changed = np.zeros(B, dtype=bool)
            for i in range(B):
                if not done[i]:
                    cls[i], gnew[i] = _eval(rhs[i], xp[t * B + i], vare, vara_fold, logpi, up[t * B + i], zp[t * B + i])
                    changed[i] = gnew[i] != gold[i]
            if not changed.any():
                break
            first = int(np.argmax(changed))
            delta = gnew[first] - gold[first]
            done[: first + 1] = True
            chg.append((first, delta))
            queue.append((t * B + first, delta))
            G = Xp[:, t * B + first] @ Xp[:, cols]
            rhs[~done] -= G[~done] * delta
            if first + 1 >= B:
                break
        for dt in range(1, D):
            if t + dt >= T:
                break
            cols2 = slice((t + dt) * B, (t + dt + 1) * B)
            for (a, dl) in chg:
                ring[(t + dt) % D] += (Xp[:, t * B + a] @ Xp[:, cols2]) * dl
        gp[cols], cls_out[cols] = gnew, cls
        tile_qend.append(len(queue))
    return r_stream, gp[:m], cls_out[:m]


@pytest.mark.parametrize("B,D", [(8, 1), (8, 3), (16, 4), (32, 2)])
def test_lagged_tiles_with_gram_corrections_equal_literal_sweep(oracle, B, D):
    y, X8 = synth(120, 200, seed=5, n_causal=10, h2=0.7)
    X = X8.astype(np.float64)
    X[:, 17] = 1.0  # a monomorphic SNP is skipped (Bayes.cpp:589)
    xpx = (X * X).sum(axis=0)
    xpx[17] = 0.0
    r0 = y - y.mean()
    vare, varg = 0.5 * y.var(), 0.02
    fold = np.array([0, 1e-2, 1e-1, 1.0])
    vara_fold = varg * fold
    logpi = np.log([0.7, 0.1, 0.1, 0.1])  # dense enough that most tiles hold several changes
    g = np.zeros(200)
    r_lit, r_til = r0.copy(), r0.copy()
    g_lit, g_til = g.copy(), g.copy()
    for it in range(3):
        us, zs = _draws(oracle, 99, it, 200)
        r_lit, g_lit, c_lit = literal_sweep(X, r_lit, g_lit, xpx, vare, vara_fold, logpi, us, zs)
        r_til, g_til, c_til = tiled_sweep(X, r_til, g_til, xpx, vare, vara_fold, logpi, us, zs, B, D)
        assert (c_lit != 0).sum() > 20
        assert np.array_equal(c_lit, c_til)
        assert np.allclose(g_lit, g_til, rtol=1e-9, atol=1e-12)
        assert np.allclose(r_lit, r_til, rtol=1e-9, atol=1e-10)
