"""Builds the committed fixtures under tests/golden/ from the reference's bundled demo data.

Run in the build container only (reads /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py

Outputs
  demo.npz                 genotypes (600 x 1000 int8, decoded with the reference's code map
                           read_bed.cpp:116-120: 00->2, 10->1, 11->0, 01->NA), ids, phenotypes,
                           map and COJO summary statistics of inst/extdata/demo.*
  demo_oracle_<model>.npz  outputs of the CPU oracle on BASELINE config 1 (demo data, T1 ~ 1,
                           200 iterations) -- regression pins for the oracle itself; the reference
                           publishes no golden vectors.
  demo_oracle_sbayesd_<model>.npz
                           the same for the SBayesD oracle on the reference's COJO file demo.ma with the LD
                           matrix of the demo genotypes
  demo_bed.npz             the raw bytes of inst/extdata/demo.bed (the decoder's input; demo.npz holds an independent
                           numpy decode of the same file) -- `python make_golden.py bed`
  sbayess_inputs.npz, demo_oracle_sbayess_<model>.npz
                           SBayesS oracle on a small synthetic data set with a thresholded LD matrix (inputs
                           stored); `python make_golden.py sbayes` rebuilds the SBayes pins only.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/inst/extdata"


def read_bed(prefix):
    fam = [l.split() for l in open(prefix + ".fam")]
    bim = [l.split() for l in open(prefix + ".bim")]
    n, m = len(fam), len(bim)
    raw = np.fromfile(prefix + ".bed", dtype=np.uint8)
    assert raw[0] == 0x6C and raw[1] == 0x1B and raw[2] == 0x01, "not a SNP-major .bed"
    bpl = (n + 3) // 4
    raw = raw[3:].reshape(m, bpl)
    code = np.array([2, -9, 1, 0], dtype=np.int8)  # index = 2-bit value; 1 -> NA
    g = np.empty((m, bpl * 4), dtype=np.int8)
    for x in range(4):
        g[:, x::4] = code[(raw >> (2 * x)) & 3]
    g = g[:, :n].T.copy()  # n x m
    return g, fam, bim


def main():
    geno, fam, bim = read_bed(os.path.join(REF, "demo"))
    assert (geno >= 0).all(), "demo data has no missing genotypes"
    phe_lines = [l.rstrip("\n").split("\t") for l in open(os.path.join(REF, "demo.phe"))]
    hdr, rows = phe_lines[0], phe_lines[1:]
    cols = {h: [r[i] for r in rows] for i, h in enumerate(hdr)}

    def num(v):
        try:
            return float(v)
        except ValueError:
            return np.nan

    ma = [l.split() for l in open(os.path.join(REF, "demo.ma"))][1:]
    ped = [l.split() for l in open(os.path.join(REF, "demo.ped"))][1:]
    np.savez_compressed(
        os.path.join(HERE, "demo.npz"),
        geno=geno,
        geno_id=np.array([f[1] for f in fam]),
        snp=np.array([b[1] for b in bim]), chr=np.array([b[0] for b in bim]),
        pos=np.array([int(b[3]) for b in bim], dtype=np.int64),
        phe_id=np.array(cols["id"]), sex=np.array(cols["sex"]), season=np.array(cols["season"]),
        day=np.array([num(v) for v in cols["day"]]), bwt=np.array([num(v) for v in cols["bwt"]]),
        loc=np.array(cols["loc"]), dam=np.array(cols["dam"]), T1=np.array([num(v) for v in cols["T1"]]),
        ma_maf=np.array([num(r[3]) for r in ma]), ma_beta=np.array([num(r[4]) for r in ma]),
        ma_se=np.array([num(r[5]) for r in ma]), ma_n=np.array([num(r[7]) for r in ma]),
        ped=np.array(ped),
    )
    from tests.util_demo import load_demo_T1
    from oracle import hb_oracle
    y, X = load_demo_T1()
    for model, Pi, fold in [("BayesCpi", [0.95, 0.05], None),
                            ("BayesR", [0.95, 0.02, 0.02, 0.01], [0, 1e-4, 1e-3, 1e-2])]:
        r = hb_oracle.bayes(y, X, model, Pi, fold=fold, niter=200, nburn=100, thin=5, seed=666666)
        np.savez_compressed(
            os.path.join(HERE, "demo_oracle_%s.npz" % model),
            Vg=r["Vg"], Ve=r["Ve"], h2=r["h2"], mu=r["mu"], alpha=r["alpha"], pi=r["pi"], pip=r["pip"],
            g=r["g"], tracker=r["diag"]["tracker"], nnz_trace=r["diag"]["nnz_trace"],
            nzrate_count=r["diag"]["nzrate_count"], vare_trace=r["diag"]["vare_trace"],
        )
        print(model, "Vg %.4f Ve %.4f h2 %.4f mu %.4f pi %s nnz_last %d" %
              (r["Vg"], r["Ve"], r["h2"], r["mu"], np.round(r["pi"], 4), r["diag"]["nnz_trace"][-1]))


def bed_main():
    raw = np.fromfile(os.path.join(REF, "demo.bed"), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "demo_bed.npz"), bed=raw)
    print("demo_bed.npz:", raw.shape[0], "bytes")


def demo_sumstat():
    """sbrm()-style inputs from the bundled files: the reference's own COJO file demo.ma (MAF, BETA, SE, NMISS:
    R/sbayes.r:209) and the LD matrix of the bundled genotypes, centred X'X / n (tXXmat.cpp:174-179)."""
    d = np.load(os.path.join(HERE, "demo.npz"))
    X = d["geno"].astype(np.float64)
    Xc = X - X.mean(axis=0)
    ld = Xc.T @ Xc / X.shape[0]
    ss = np.column_stack([d["ma_maf"], d["ma_beta"], d["ma_se"], d["ma_n"]])
    return np.asfortranarray(ss), np.asfortranarray(ld)


def sparse_ld(ld, n, chisq=3.84):
    """the chi-square sparsifier of tXXmat.cpp:146-153: an entry stays when r^2 n > chisq"""
    import scipy.sparse as sp
    dd = np.sqrt(np.clip(np.diag(ld), 1e-300, None))
    r2 = (ld / dd[:, None] / dd[None, :]) ** 2
    keep = (r2 * n > chisq) | np.eye(ld.shape[0], dtype=bool)
    return sp.csc_matrix(np.where(keep, ld, 0.0))


SBAYESS_M = 200


def sbayes_main():
    """regression pins for the SBayesD / SBayesS oracles (demo_oracle_sbayes*.npz)"""
    from oracle import hb_oracle
    ss, ld = demo_sumstat()
    kw = dict(niter=100, nburn=50, thin=5, seed=666666)
    # SBayesS: a synthetic data set (1500 individuals, 200 SNPs), inputs stored next to the outputs
    # (sbayess_inputs.npz).  With the bundled demo data a thresholded LD matrix is not positive definite (few animals,
    # long LD blocks) and the chain diverges -- in the reference as well, nothing in SBayesS.cpp guards against it.
    from tests.util_demo import synth
    from tests.util_sumstat import make_sumstat
    ys, Xs = synth(1500, SBAYESS_M, seed=23, n_causal=10)
    ss_s, ld_s = make_sumstat(ys, Xs)
    sld = sparse_ld(ld_s, len(ys))
    np.savez_compressed(os.path.join(HERE, "sbayess_inputs.npz"), sumstat=ss_s, ld_thresholded=sld.toarray())
    for tag, fn, ssx, ldm in (("sbayesd", hb_oracle.sbayesd, ss, ld), ("sbayess", hb_oracle.sbayess, ss_s, sld)):
        for model, Pi, fold in [("BayesCpi", [0.95, 0.05], None),
                                ("BayesR", [0.95, 0.02, 0.02, 0.01], [0, 1e-4, 1e-3, 1e-2])]:
            r = fn(ssx, ldm, model, Pi, fold=fold, **kw)
            assert np.all(np.isfinite(r["alpha"])) and np.isfinite(r["Ve"]), (tag, model)
            np.savez_compressed(
                os.path.join(HERE, "demo_oracle_%s_%s.npz" % (tag, model)),
                Vg=r["Vg"], Ve=r["Ve"], h2=r["h2"], alpha=r["alpha"], pi=r["pi"], pip=r["pip"],
                tracker=r["diag"]["tracker"], nnz_trace=r["diag"]["nnz_trace"], nzrate_count=r["diag"]["nzrate_count"],
                vare_trace=r["diag"]["vare_trace"], vara_trace=r["diag"]["vara_trace"], n_used=r["diag"]["n_used"],
            )
            print(tag, model, "Vg %.5f Ve %.5f h2 %.4f pi %s nnz_last %d n %d" %
                  (r["Vg"], r["Ve"], r["h2"], np.round(r["pi"], 4), r["diag"]["nnz_trace"][-1], r["diag"]["n_used"]))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "sbayes":
        sbayes_main()   # needs only tests/golden/demo.npz
    elif len(sys.argv) > 1 and sys.argv[1] == "bed":
        bed_main()
    else:
        main()
        bed_main()
        sbayes_main()
