"""Reads an HB_TRACE dump ([T][16] globaltimer ns per tile) and prints where a tile's time goes.
events: 0 dots seen by worker, 1 P done, 2 hand-over from t-1 arrived, 3 S done (hand-over posted), 4 updates published,
5 C done, 6 AXPY warp of slab 0 fetched the tile's updates, 7 compute warp 0 of slab 0 added its dots of the tile.
Serial mode (hb_serial.cuh): 0/1 = the helper's phase P (dots seen, package flagged), 2 = package in the serial CTA's shared
memory, 3 = tile final, 4 = published, 5 = serial CTA done with the tile, 8/9/10 = the helper's phase C (changes seen, first far
correction posted, done), 11 = the loader saw the package flag."""
import sys
import numpy as np
a = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(-1, 16).astype(np.int64)
D = int(sys.argv[2]) if len(sys.argv) > 2 else 4
T = a.shape[0]
lo, hi = T // 4, 3 * T // 4
s = a[lo:hi]
def d(x, y): return float(np.median(s[:, x] - s[:, y]))
print("tiles", T, "window", lo, hi)
print("period (S done t - S done t-1): %.0f ns" % float(np.median(np.diff(a[lo:hi, 3]))))
print("dots added by slab0 (7) -> dots seen by worker (0): %.0f" % d(0, 7))
print("dots seen (0) -> P done (1): %.0f" % d(1, 0))
print("P done (1) -> hand-over arrived (2): %.0f" % d(2, 1))
print("hand-over arrived (2) -> S done (3): %.0f" % d(3, 2))
print("S done (3) -> published (4): %.0f" % d(4, 3))
print("published (4) -> AXPY fetched (6): %.0f" % d(6, 4))
print("S done (3) -> C done (5): %.0f" % d(5, 3))
print("S done of t-1 (3) -> hand-over arrived at t (2): %.0f" % float(np.median(a[lo + 1:hi, 2] - a[lo:hi - 1, 3])))
print("AXPY fetched t (6) -> slab0 added dots of t+D (7): %.0f" % float(np.median(a[lo + D:hi, 7] - a[lo:hi - D, 6])))
print("published t (4) -> dots of t+D seen (0): %.0f" % float(np.median(a[lo + D:hi, 0] - a[lo:hi - D, 4])))
# who gates the start of phase S: the tile's own phase P (dots + gather) or the previous tile's hand-over?
late = (a[lo + 1:hi, 1] - a[lo:hi - 1, 3]).astype(np.float64)
print("P done of t minus S done of t-1: median %.0f ns, P later in %.1f %% of tiles" % (float(np.median(late)), 100.0 * float(np.mean(late > 0))))
gate = np.maximum(a[lo + 1:hi, 1], a[lo:hi - 1, 3])
det = (a[lo + 1:hi, 2] - gate).astype(np.float64)
print("hand-over seen after max(P done, S done of t-1): median %.0f ns, p10 %.0f, p90 %.0f" % (float(np.median(det)), float(np.percentile(det, 10)), float(np.percentile(det, 90))))
per = np.diff(a[lo:hi, 3]).astype(np.float64)
print("period: p10 %.0f p50 %.0f p90 %.0f mean %.0f" % (float(np.percentile(per, 10)), float(np.percentile(per, 50)), float(np.percentile(per, 90)), float(per.mean())))
dd = (a[lo + D:hi, 0] - a[lo:hi - D, 3]).astype(np.float64)
print("S done of t -> dots of t+D seen by its worker: median %.0f ns (budget: D periods minus P and S)" % float(np.median(dd)))
ho = (a[lo + 1:hi, 2] - a[lo:hi - 1, 3]).astype(np.float64)
tt = np.arange(lo + 1, hi)
for par, name in ((1, "odd tiles (cluster mode: local hand-over)"), (0, "even tiles")):
    x = ho[(tt & 1) == par]
    print("S done of t-1 -> hand-over arrived at t, %s: mean %.0f p10 %.0f p50 %.0f p90 %.0f" % (name, x.mean(), np.percentile(x, 10), np.percentile(x, 50), np.percentile(x, 90)))
sd = (a[lo:hi, 3] - a[lo:hi, 2]).astype(np.float64)
print("hand-over arrived -> S done: mean %.0f p10 %.0f p50 %.0f p90 %.0f" % (sd.mean(), np.percentile(sd, 10), np.percentile(sd, 50), np.percentile(sd, 90)))

if a[lo:hi, 8].any():
    print("-- serial mode")
    print("package flagged (1) -> loader saw the flag (11): %.0f" % d(11, 1))
    print("loader saw the flag (11) -> package in shared memory, tile started (2): %.0f" % d(2, 11))
    print("published (4) -> helper saw the changes (8): %.0f" % d(8, 4))
    print("helper saw the changes (8) -> first far correction posted (9): %.0f" % d(9, 8))
    print("helper saw the changes (8) -> phase C done (10): %.0f" % d(10, 8))
    print("tile final (3) -> serial CTA done with the tile (5): %.0f" % d(5, 3))
    print("serial CTA done with t-1 (5) -> tile t started (2): %.0f" % float(np.median(a[lo + 1:hi, 2] - a[lo:hi - 1, 5])))
    if a[lo:hi, 12].any():
        print("loader saw the flag (11) -> bulk copies issued (12): %.0f" % d(12, 11))
        print("bulk copies issued (12) -> far slots landed (13): %.0f" % d(13, 12))
        print("far slots landed (13) -> far sums done (14): %.0f" % d(14, 13))
        print("bulk copies issued (12) -> package landed, as seen by the chain (15): %.0f" % d(15, 12))
        print("package seen (15) -> tile started (2; = far sums seen): %.0f" % d(2, 15))
        print("serial CTA done with t-2 (5) -> loader saw the flag of t (11): %.0f" % float(np.median(a[lo + 2:hi, 11] - a[lo:hi - 2, 5])))
