"""Python mirror of the reference interface for the hot path, on top of the C ABI.

`Bayes()` has the argument list of the reference's C++ `Bayes()` (/root/reference/src/Bayes.cpp:60-88,
what R's ibrm() calls at R/bayes.r:294-296) and returns a dict named like its Rcpp::List
(:919-1040).  `Engine` exposes the device engine (hb_engine_*) used by the host driver, the
parity tests and bench.py.  Nothing here computes: every call goes to libhibayes_b200.so.
"""
import ctypes as C
import math

import numpy as np

from . import _lib

MODEL_INDEX = {"BayesRR": 1, "BayesA": 2, "BayesB": 3, "BayesBpi": 3, "BayesC": 4, "BayesCpi": 4, "BSLMM": 4, "BayesL": 5}


def _ptr(a):
    return None if a is None else a.ctypes.data


def _nan(v):
    return float("nan") if v is None else float(v)


def synth_geno_host(n, m, seed, row_offset=0):
    """The synthetic genotype matrix of hb_engine_synth_geno, generated on the host (int8, F order)."""
    L = _lib.load_library()
    X = np.empty((n, m), dtype=np.int8, order="F")
    _lib.check(L.hb_synth_geno_host(X.ctypes.data, n, m, seed, row_offset))
    return X


def synth_geno_host_into(Xblock, seed, col_offset=0, row_offset=0):
    """Fills a column block (F-ordered int8 view, n x ncols) of the synthetic matrix: columns col_offset .. of the matrix
    synth_geno_host() returns.  The call releases the GIL, so blocks can be filled from several threads."""
    L = _lib.load_library()
    n, nc = Xblock.shape
    assert Xblock.dtype == np.int8 and Xblock.strides == (1, n)
    _lib.check(L.hb_synth_geno_host_cols(Xblock.ctypes.data, n, col_offset, nc, seed, row_offset))


def _is_zero_chromosome(v):
    """"0" as a string, 0 / 0.0 as a number (R compares the values after as.character / as.numeric)."""
    if isinstance(v, (bytes, str)):
        t = v.decode() if isinstance(v, bytes) else v
        try:
            return float(t) == 0.0
        except ValueError:
            return False
    try:
        return float(v) == 0.0
    except (TypeError, ValueError):
        return False


class BedGeno:
    """Genotypes held as the image of a SNP-major PLINK .bed file; accepted by Bayes() and LdMat in place of
    a matrix, decoded on the device (hb_engine_load_bed / hb_ldmat_load_bed; read_bed<char>() of
    /root/reference/src/read_bed.cpp:97-232).  rows: 0-based file individuals in the order of y
    (the `M[index, ]` selection of R/bayes.r:281-291), default all in file order."""

    def __init__(self, image, nid, m, rows=None, impute=True, mode="A"):
        if mode not in ("A", "D"):
            raise ValueError("mode must be 'A' or 'D'")  # R/read_plink.r:28 match.arg
        self.image = np.ascontiguousarray(np.frombuffer(image, dtype=np.uint8) if isinstance(image, (bytes, bytearray)) else image,
                                          dtype=np.uint8)
        self.nid, self.m = int(nid), int(m)
        self.rows = None if rows is None else np.ascontiguousarray(rows, dtype=np.int32)
        self.impute, self.dominance = bool(impute), mode == "D"
        self.shape = (self.nid if self.rows is None else self.rows.shape[0], self.m)

    @classmethod
    def from_file(cls, bfile, nid, m, **kw):
        path = bfile if bfile.endswith(".bed") else bfile + ".bed"  # read_bed.cpp:99-102
        return cls(np.fromfile(path, dtype=np.uint8), nid, m, **kw)

    def c_struct(self):
        b = _lib.BedSource()
        b.file, b.len, b.nid = self.image.ctypes.data, self.image.shape[0], self.nid
        b.rows = _ptr(self.rows)
        b.impt, b.dominance = int(self.impute), int(self.dominance)
        return b


def read_bed(image, nid, m, impute=True, mode="A", device=0):
    """read_bed() of the reference (R/read_plink.r:60-67 -> src/read_bed.cpp:235) on the device: returns the
    nid x m int8 matrix a "char" big.matrix would hold (NA = -128) and the per-SNP missing flags."""
    L = _lib.load_library()
    g = BedGeno(image, nid, m, impute=impute, mode=mode)
    out = np.empty((nid, m), dtype=np.int8, order="F")
    miss = np.zeros(m, dtype=np.uint8)
    _lib.check(L.hb_bed_decode(device, g.image.ctypes.data, g.image.shape[0], nid, m, int(g.impute), int(g.dominance),
                               out.ctypes.data, miss.ctypes.data))
    return out, miss


class LdMat:
    """Device LD builder (hb_ldmat_*): tXXmat_Geno / tXXmat_Chr of /root/reference/src/tXXmat.cpp."""

    def __init__(self, geno, device=0, panel_cols=0):
        self.L = _lib.load_library()
        self.h = C.c_void_p()
        self.n, self.m = geno.shape
        _lib.check(self.L.hb_ldmat_create(device, self.n, self.m, C.byref(self.h)))
        if panel_cols:
            _lib.check(self.L.hb_ldmat_set_panel_cols(self.h, panel_cols))
        if isinstance(geno, BedGeno):
            _lib.check(self.L.hb_ldmat_load_bed(self.h, geno.image.ctypes.data, geno.image.shape[0], geno.nid, _ptr(geno.rows),
                                                int(geno.impute), int(geno.dominance)))
        else:
            Xf = np.asfortranarray(geno)
            if Xf.dtype != np.int8:
                raise TypeError("LdMat holds int8 genotypes (the reference's 'char' big.matrix)")
            _lib.check(self.L.hb_ldmat_load_i8(self.h, Xf.ctypes.data, Xf.shape[0]))

    def close(self):
        if self.h:
            self.L.hb_ldmat_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def stats(self):
        """BigStat(): dict(mean, sum, xx)."""
        mean, sm, xx = np.zeros(self.m), np.zeros(self.m), np.zeros(self.m)
        _lib.check(self.L.hb_ldmat_stats(self.h, mean.ctypes.data, sm.ctypes.data, xx.ctypes.data))
        return {"mean": mean, "sum": sm, "xx": xx}

    @staticmethod
    def _chr(chr_, m):
        if chr_ is None:
            return None
        c = np.ascontiguousarray(chr_, dtype=np.int32)
        if c.shape[0] != m:
            raise ValueError("chr needs one code per SNP")
        return c

    def dense(self, chr=None, chisq=None):
        c = self._chr(chr, self.m)
        out = np.zeros((self.m, self.m), order="F")
        _lib.check(self.L.hb_ldmat_dense(self.h, _ptr(c), int(chisq is not None), 0.0 if chisq is None else float(chisq),
                                         out.ctypes.data, self.m))
        return out

    def sparse(self, chr=None, chisq=None):
        import scipy.sparse as sp
        c = self._chr(chr, self.m)
        nnz = C.c_longlong(0)
        _lib.check(self.L.hb_ldmat_sparse(self.h, _ptr(c), int(chisq is not None), 0.0 if chisq is None else float(chisq),
                                          C.byref(nnz)))
        colptr = np.zeros(self.m + 1, dtype=np.int64)
        rowidx = np.zeros(max(nnz.value, 1), dtype=np.int32)
        val = np.zeros(max(nnz.value, 1))
        _lib.check(self.L.hb_ldmat_sparse_get(self.h, colptr.ctypes.data, rowidx.ctypes.data, val.ctypes.data))
        return sp.csc_matrix((val[:nnz.value], rowidx[:nnz.value], colptr), shape=(self.m, self.m))

    def last_ms(self):
        ms = C.c_float(0)
        _lib.check(self.L.hb_ldmat_last_ms(self.h, C.byref(ms)))
        return ms.value


def ldmat_plan(m, map_chr=None, chisq=None, ldchr=False):
    """The branch ldmat() of the reference takes (R/ldm.r:44-94) for its arguments: returns
    (kernel, chisq) with kernel in {"geno_dense", "geno_sparse", "chr_dense", "chr_sparse"}.
    map_chr: the chromosome column of `map` (None = no map given)."""
    if chisq is not None and chisq < 0:
        chisq = None  # :45-47
    if map_chr is not None:
        map_chr = np.asarray(map_chr)
        if map_chr.shape[0] != m:
            raise ValueError("map needs one row per SNP")
        one = len(set(map_chr.tolist())) == 1
        if one:
            ldchr = True  # :52
        if chisq is not None and chisq == 0 and one:
            chisq = None  # :53-55
        if any(v is None or (isinstance(v, float) and math.isnan(v)) for v in map_chr.tolist()):
            raise RuntimeError("NAs are not allowed in chromosome.")  # :60
        if any(_is_zero_chromosome(v) for v in map_chr.tolist()):
            raise RuntimeError("0 is not allowed in chromosome.")  # :63
    else:
        if chisq is not None and chisq == 0:
            chisq = None  # :80-82
        ldchr = True  # :83
    if ldchr:
        # tXXmat_Geno: sparse only for chisq > 0 (tXXmat.cpp:118-121)
        return ("geno_sparse" if (chisq is not None and chisq > 0) else "geno_dense"), chisq
    return ("chr_sparse" if chisq is not None else "chr_dense"), chisq  # tXXmat_Chr, :520-523


def ldmat(geno, map_chr=None, chisq=None, ldchr=False, device=0):
    """ldmat() of the reference (R/ldm.r:31-112) without the gwas.geno merge: a dense ndarray for the
    full matrix, a scipy CSC matrix where the reference returns a dgCMatrix."""
    m = geno.shape[1]
    kernel, chisq = ldmat_plan(m, map_chr, chisq, ldchr)
    chr_codes = None
    if kernel.startswith("chr"):
        # non-numeric chromosome names become max+1, max+2, ... (R/ldm.r:67-76); only equality matters
        names = [str(v) for v in np.asarray(map_chr).tolist()]
        lut = {s: i for i, s in enumerate(dict.fromkeys(names))}
        chr_codes = np.array([lut[s] for s in names], dtype=np.int32)
    h = LdMat(geno, device=device)
    try:
        if kernel == "geno_dense":
            return h.dense()
        if kernel == "geno_sparse":
            return h.sparse(chisq=chisq)
        if kernel == "chr_dense":
            return h.sparse(chr=chr_codes)
        return h.sparse(chr=chr_codes, chisq=chisq)
    finally:
        h.close()


def Bayes(y, X, model, Pi, Kival=None, Ki=None, C_=None, R=None, fold=None, niter=50000, nburn=20000, thin=5,
          epsl_y_J=None, epsl_Gi=None, epsl_index=None, dfvr=None, s2vr=None, vg=None, dfvg=None, s2vg=None,
          ve=None, dfve=None, s2ve=None, windindx=None, outfreq=100, threads=0, verbose=False,
          seed=666666, device=0, tile_snps=0, lag_tiles=0, n_slabs=0, store_alpha=False, comm=None):
    """GPU twin of hibayes' Bayes().  R: (n, nr) integer level codes (0-based) standing for the
    CharacterMatrix of environmental random effects; seed: the Philox run key the Rcpp shim
    derives from R's RNG state (the reference takes no seed argument, Bayes.cpp:60-88)."""
    if (Ki is None) != (Kival is None):
        raise RuntimeError("Ki and Kival should be provided together.")
    L = _lib.load_library()
    y = np.ascontiguousarray(y, dtype=np.float64)
    n = y.shape[0]
    bed_src = None
    if isinstance(X, BedGeno):
        if X.shape[0] != n:
            raise RuntimeError("Number of individuals not equals.")  # Bayes.cpp:96
        bed_src = X.c_struct()
        Xf, xt, m = bed_src, 2, X.shape[1]
    else:
        X = np.asarray(X)
        if X.shape[0] != n:
            raise RuntimeError("Number of individuals not equals.")  # Bayes.cpp:96
        if X.dtype == np.int8:
            Xf, xt = np.asfortranarray(X), 1
        else:
            Xf, xt = np.asfortranarray(X, dtype=np.float64), 0
        m = Xf.shape[1]
    Pi = np.ascontiguousarray(Pi, dtype=np.float64)
    F = Pi.shape[0]
    fold_a = None if fold is None else np.ascontiguousarray(fold, dtype=np.float64)
    if fold_a is not None and fold_a.shape[0] != F:
        raise RuntimeError("length of Pi and fold not equals.")  # :115-117
    a = _lib.BayesArgs()
    a.n, a.m, a.y, a.x_type = n, m, _ptr(y), xt
    a.X = C.addressof(bed_src) if bed_src is not None else _ptr(Xf)
    a.model = model.encode()
    a.n_fold, a.Pi, a.fold = F, _ptr(Pi), _ptr(fold_a)
    keep = [y, Xf, X, Pi, fold_a]
    nc = 0
    if C_ is not None:
        Cf = np.asfortranarray(C_, dtype=np.float64)
        if Cf.shape[0] != n:
            raise RuntimeError("Number of individuals does not match for covariates.")
        nc = Cf.shape[1]
        a.C = _ptr(Cf)
        keep.append(Cf)
    a.nc = nc
    nr, n_levels = 0, 0
    if R is not None:
        Rf = np.asfortranarray(R, dtype=np.int32)
        if Rf.shape[0] != n:
            raise RuntimeError("Number of individuals does not match for environmental random effects.")
        nr = Rf.shape[1]
        if Rf.size and Rf.min() < 0:
            raise RuntimeError("level codes of the environmental random effects must be >= 0")
        nlev = np.ascontiguousarray(Rf.max(axis=0) + 1, dtype=np.int32)
        if comm is not None and comm.world > 1:   # the levels of a term are the same set on every rank
            allmax = np.frombuffer(comm.allgather_bytes(nlev.astype(np.int64).tobytes()), dtype=np.int64).reshape(comm.world, nr)
            nlev = np.ascontiguousarray(allmax.max(axis=0), dtype=np.int32)
        n_levels = int(nlev.sum())
        a.Rlev, a.nlev = _ptr(Rf), _ptr(nlev)
        keep += [Rf, nlev]
    a.nr = nr
    a.niter, a.nburn, a.thin = niter, nburn, thin
    a.dfvr, a.s2vr, a.vg, a.dfvg = _nan(dfvr), _nan(s2vr), _nan(vg), _nan(dfvg)
    a.s2vg, a.ve, a.dfve, a.s2ve = _nan(s2vg), _nan(ve), _nan(dfve), _nan(s2ve)
    nw = 0
    if windindx is not None:
        w = np.ascontiguousarray(windindx, dtype=np.int32)
        if w.shape[0] != m:
            raise RuntimeError("windindx should have one entry per SNP.")
        nw = int(w.max())
        a.windindx = _ptr(w)
        keep.append(w)
    a.outfreq, a.verbose, a.seed = outfreq, int(bool(verbose)), seed
    ne = qe = 0
    if epsl_index is not None:
        import scipy.sparse as sp
        ei = np.ascontiguousarray(epsl_index, dtype=np.int32)
        G = sp.csc_matrix(epsl_Gi)
        G.sort_indices()
        cp = np.ascontiguousarray(G.indptr, dtype=np.int32)
        ri = np.ascontiguousarray(G.indices, dtype=np.int32)
        gv = np.ascontiguousarray(G.data, dtype=np.float64)
        yj = np.ascontiguousarray(epsl_y_J, dtype=np.float64)
        ne, qe = ei.shape[0], G.shape[0]
        if G.shape[0] != G.shape[1]:
            raise RuntimeError("variance-covariance matrix should be in square.")   # Bayes.cpp:263
        if G.indptr[-1] >= 2 ** 31:
            raise RuntimeError("epsl_Gi has too many stored entries for 32-bit column pointers")
        if ne and (ei.min() < 1 or ei.max() > qe):
            raise RuntimeError("epsl_index should be 1-based positions inside epsl_Gi")
        if ne > n or yj.shape[0] != n:
            raise RuntimeError("epsl_y_J / epsl_index do not match the number of individuals")
        a.epsl_y_J, a.epsl_index, a.Gi_colptr, a.Gi_rowidx, a.Gi_val = _ptr(yj), _ptr(ei), _ptr(cp), _ptr(ri), _ptr(gv)
        keep += [ei, cp, ri, gv, yj]
    a.ne, a.qe = ne, qe
    if Ki is not None:   # BSLMM: eigenvectors (n x n) and eigenvalues of the relationship matrix (Bayes.cpp:218-233)
        Kf = np.asfortranarray(Ki, dtype=np.float64)
        kv = np.ascontiguousarray(Kival, dtype=np.float64)
        if Kf.ndim != 2 or Kf.shape[0] != Kf.shape[1]:
            raise RuntimeError("variance-covariance matrix should be in square.")   # :221
        if Kf.shape[0] != n or kv.shape[0] != Kf.shape[1]:
            raise RuntimeError("Ki / Kival do not match the number of individuals")
        a.nk, a.Ki, a.Kival = Kf.shape[1], _ptr(Kf), _ptr(kv)
        keep += [Kf, kv]
    a.device, a.tile_snps, a.lag_tiles, a.n_slabs = device, tile_snps, lag_tiles, n_slabs
    if comm is not None and comm.world > 1:
        a.rank, a.world, a.n_total = comm.rank, comm.world, comm.total_rows(n)
        a.allreduce_sum_f64, a.allreduce_sum_i32_dev, a.allgather_bytes = comm.callbacks()
        keep.append(comm)
    nrec = max((niter - nburn) // thin, 0)
    o = _lib.BayesOut()
    res = {
        "beta": np.zeros(nc), "alpha": np.zeros(m), "pi": np.zeros(F), "pip": np.zeros(m),
        "gwas": np.zeros(nw), "g": np.zeros(n), "e": np.zeros(n), "Vr": np.zeros(nr),
        "r": np.zeros(n_levels), "epsilon": np.zeros(qe),
    }
    mc = {
        "mu": np.zeros(nrec), "Vg": np.zeros(nrec), "Ve": np.zeros(nrec), "h2": np.zeros(nrec),
        "pi": np.zeros((F, nrec), order="F"), "beta": np.zeros((nc, nrec), order="F"),
    }
    if store_alpha:
        mc["alpha"] = np.zeros((m, nrec), order="F")
    dg = {
        "tracker": np.zeros(m, dtype=np.int32), "nzrate_count": np.zeros(m), "wppa_count": np.zeros(nw),
        "nnz_trace": np.zeros(niter, dtype=np.int32), "vara_trace": np.zeros(niter),
        "vare_trace": np.zeros(niter), "varg_trace": np.zeros(niter),
        "rounds_trace": np.zeros(niter, dtype=np.int32), "sweep_ms_trace": np.zeros(niter, dtype=np.float32),
    }
    o.rounds_trace, o.sweep_ms_trace = _ptr(dg["rounds_trace"]), _ptr(dg["sweep_ms_trace"])
    if nr:   # MCMCsamples of the other terms (Bayes.cpp:987-1020)
        mc["Vr"], mc["r"] = np.zeros((nr, nrec), order="F"), np.zeros((n_levels, nrec), order="F")
        o.vr_store, o.estR_store = _ptr(mc["Vr"]), _ptr(mc["r"])
    if qe:
        mc["Veps"], mc["J"], mc["epsilon"] = np.zeros(nrec), np.zeros(nrec), np.zeros((qe, nrec), order="F")
        o.veps_store, o.J_store, o.epsilon_store = _ptr(mc["Veps"]), _ptr(mc["J"]), _ptr(mc["epsilon"])
    o.beta, o.alpha, o.pi, o.pip = _ptr(res["beta"]), _ptr(res["alpha"]), _ptr(res["pi"]), _ptr(res["pip"])
    o.gwas = _ptr(res["gwas"]) if nw else None
    o.g, o.e, o.vr, o.estR, o.epsilon = _ptr(res["g"]), _ptr(res["e"]), _ptr(res["Vr"]), _ptr(res["r"]), _ptr(res["epsilon"])
    o.mu_store, o.vara_store, o.vare_store, o.hsq_store = _ptr(mc["mu"]), _ptr(mc["Vg"]), _ptr(mc["Ve"]), _ptr(mc["h2"])
    o.pi_store, o.beta_store = _ptr(mc["pi"]), _ptr(mc["beta"])
    o.alpha_store = _ptr(mc["alpha"]) if store_alpha else None
    o.tracker_final, o.nzrate_count = _ptr(dg["tracker"]), _ptr(dg["nzrate_count"])
    o.wppa_count = _ptr(dg["wppa_count"]) if nw else None
    o.nnz_trace, o.vara_trace, o.vare_trace, o.varg_trace = (_ptr(dg["nnz_trace"]), _ptr(dg["vara_trace"]),
                                                             _ptr(dg["vare_trace"]), _ptr(dg["varg_trace"]))
    _lib.check(L.hb_bayes(C.byref(a), C.byref(o)))
    res.update({"Vg": o.Vg, "Ve": o.Ve, "h2": o.h2, "mu": o.mu, "Veps": o.Veps, "J": o.J})
    res["MCMCsamples"] = mc
    dg.update({"n_records": o.n_records_done, "nzct": o.nzct, "iters_done": o.iters_done,
               "seconds_sweep": o.seconds_sweep, "seconds_setup": o.seconds_setup,
               "rounds_total": o.rounds_total, "tiles_total": o.tiles_total})
    res["diag"] = dg
    return res


def _sbayes(sparse, sumstat, ldm, model, Pi, niter, nburn, thin, fold, windindx, vg, dfvg, s2vg, ve, dfve, s2ve, outfreq, verbose,
            seed, device, store_alpha=False):
    L = _lib.load_library()
    ss = np.asfortranarray(sumstat, dtype=np.float64)
    a = _lib.SBayesArgs()
    keep = [ss]
    if sparse:
        import scipy.sparse as sp
        G = sp.csc_matrix(ldm)
        G.sort_indices()
        G.eliminate_zeros()
        cp = np.ascontiguousarray(G.indptr, dtype=np.int32)
        ri = np.ascontiguousarray(G.indices, dtype=np.int32)
        gv = np.ascontiguousarray(G.data, dtype=np.float64)
        a.ld_colptr, a.ld_rowidx, a.ld_val = _ptr(cp), _ptr(ri), _ptr(gv)
        keep += [cp, ri, gv]
        m = G.shape[0]
    else:
        ld = np.asfortranarray(ldm, dtype=np.float64)
        a.ldm = _ptr(ld)
        keep.append(ld)
        m = ld.shape[0]
    if ss.shape[0] != m:
        raise RuntimeError("Number of SNPs not equals.")  # SBayesD.cpp:29-31
    Pi = np.ascontiguousarray(Pi, dtype=np.float64)
    F = Pi.shape[0]
    fo = None if fold is None else np.ascontiguousarray(fold, dtype=np.float64)
    if fo is not None and fo.shape[0] != F:
        raise RuntimeError("length of Pi and fold not equals.")
    a.m, a.sumstat, a.model, a.n_fold, a.Pi, a.fold = m, _ptr(ss), model.encode(), F, _ptr(Pi), _ptr(fo)
    a.niter, a.nburn, a.thin = niter, nburn, thin
    a.vg, a.dfvg, a.s2vg, a.ve, a.dfve, a.s2ve = _nan(vg), _nan(dfvg), _nan(s2vg), _nan(ve), _nan(dfve), _nan(s2ve)
    nw = 0
    if windindx is not None:
        w = np.ascontiguousarray(windindx, dtype=np.int32)
        nw = int(w.max())
        a.windindx = _ptr(w)
        keep.append(w)
    a.outfreq, a.verbose, a.seed, a.device = outfreq, int(bool(verbose)), seed, device
    o = _lib.SBayesOut()
    nrec = max((niter - nburn) // thin, 0)
    res = {"alpha": np.zeros(m), "pi": np.zeros(F), "pip": np.zeros(m), "gwas": np.zeros(nw)}
    mc = {"Vg": np.zeros(nrec), "Ve": np.zeros(nrec), "h2": np.zeros(nrec), "pi": np.zeros((F, nrec), order="F")}
    dg = {"tracker": np.zeros(m, dtype=np.int32), "nzrate_count": np.zeros(m), "wppa_count": np.zeros(nw),
          "nnz_trace": np.zeros(niter, dtype=np.int32), "vara_trace": np.zeros(niter), "vare_trace": np.zeros(niter),
          "varg_trace": np.zeros(niter), "r_hat": np.zeros(m)}
    o.alpha, o.pi, o.pip = _ptr(res["alpha"]), _ptr(res["pi"]), _ptr(res["pip"])
    o.gwas = _ptr(res["gwas"]) if nw else None
    o.vara_store, o.vare_store, o.hsq_store, o.pi_store = _ptr(mc["Vg"]), _ptr(mc["Ve"]), _ptr(mc["h2"]), _ptr(mc["pi"])
    o.tracker_final, o.nzrate_count = _ptr(dg["tracker"]), _ptr(dg["nzrate_count"])
    o.wppa_count = _ptr(dg["wppa_count"]) if nw else None
    o.nnz_trace, o.vara_trace, o.vare_trace, o.varg_trace = (_ptr(dg["nnz_trace"]), _ptr(dg["vara_trace"]),
                                                             _ptr(dg["vare_trace"]), _ptr(dg["varg_trace"]))
    o.r_hat_final = _ptr(dg["r_hat"])
    if store_alpha:   # MCMCsamples$alpha (SBayesD.cpp:566), m x records
        mc["alpha"] = np.zeros((m, nrec), order="F")
        o.alpha_store = _ptr(mc["alpha"])
    _lib.check((L.hb_sbayess if sparse else L.hb_sbayesd)(C.byref(a), C.byref(o)))
    res.update({"Vg": o.Vg, "Ve": o.Ve, "h2": o.h2, "MCMCsamples": mc})
    dg.update({"n_records": o.n_records_done, "nzct": o.nzct, "iters_done": o.iters_done, "n_used": o.n_used,
               "seconds_sweep": o.seconds_sweep, "columns_total": o.columns_total, "ld_entries_total": o.ld_entries_total,
               "ld_bytes_device": o.ld_bytes_device, "rounds_total": o.rounds_total, "tiles_total": o.tiles_total})
    res["diag"] = dg
    return res


def SBayesD(sumstat, ldm, model, Pi, niter=50000, nburn=20000, thin=5, fold=None, windindx=None, vg=None, dfvg=None,
            s2vg=None, ve=None, dfve=None, s2ve=None, outfreq=100, threads=0, verbose=False, seed=666666, device=0,
            store_alpha=False):
    """GPU twin of hibayes' SBayesD() (/root/reference/src/SBayesD.cpp:5-24; what sbrm() calls at R/sbayes.r:215 for a
    dense LD matrix).  sumstat: m x 4 (MAF, BETA, SE, N = columns 4,5,6,8 of the COJO file, R/sbayes.r:209), NaN = NA;
    ldm: m x m.  Returns a dict named like the Rcpp::List (:532-578)."""
    return _sbayes(False, sumstat, ldm, model, Pi, niter, nburn, thin, fold, windindx, vg, dfvg, s2vg, ve, dfve, s2ve, outfreq,
                   verbose, seed, device, store_alpha)


def SBayesS(sumstat, ldm, model, Pi, niter=50000, nburn=20000, thin=5, fold=None, windindx=None, vg=None, dfvg=None,
            s2vg=None, ve=None, dfve=None, s2ve=None, outfreq=100, threads=0, verbose=False, seed=666666, device=0,
            store_alpha=False):
    """GPU twin of hibayes' SBayesS() (/root/reference/src/SBayesS.cpp:21-40; sbrm() with a sparse LD matrix,
    R/sbayes.r:213): ldm is a scipy sparse matrix (dgCMatrix in R)."""
    return _sbayes(True, sumstat, ldm, model, Pi, niter, nburn, thin, fold, windindx, vg, dfvg, s2vg, ve, dfve, s2ve, outfreq,
                   verbose, seed, device, store_alpha)




def cutwind(chr, pos, windsize=None, windnum=None):
    """Window ids per SNP for the WPPA counters: cutwind_by_bp / cutwind_by_num of the reference
    (src/cutwind.cpp:13-65) with the checks of R/sbayes.r:176-182."""
    L = _lib.load_library()
    c = np.ascontiguousarray(chr, dtype=np.float64)
    p_ = np.ascontiguousarray(pos, dtype=np.float64)
    m = c.shape[0]
    out = np.zeros(m, dtype=np.int32)
    if windnum is not None:
        if m < windnum:
            raise RuntimeError("Number of markers specified in a window is larger than the total number of markers.")
        _lib.check(L.hb_cutwind_by_num(c.ctypes.data, p_.ctypes.data, m, int(windnum), out.ctypes.data))
    else:
        if windsize is None:
            raise ValueError("give windsize or windnum")
        if p_.max() < windsize:
            raise RuntimeError("Maximum of physical position is smaller than wind size.")
        _lib.check(L.hb_cutwind_by_bp(c.ctypes.data, p_.ctypes.data, m, float(windsize), out.ctypes.data))
    return out


IBRM_METHODS = ("BayesCpi", "BayesA", "BayesL", "BSLMM", "BayesR", "BayesB", "BayesC", "BayesBpi", "BayesRR")


def ibrm_plan(method="BayesCpi", Pi=None, fold=None, niter=None, nburn=None, thin=5, windsize=None, windnum=None,
              map_chr=None, map_pos=None):
    """The argument handling of ibrm() (R/bayes.r:151-276) that decides what Bayes() is called with: the method's
    default chain length and mixture, the thin check, and the windows (cutwind) when a window size or count is given."""
    if method not in IBRM_METHODS:
        raise ValueError("'arg' should be one of " + ", ".join(IBRM_METHODS))     # match.arg, :166
    if method == "BSLMM":
        raise NotImplementedError("ibrm(\"BSLMM\") builds the GRM and its eigen-decomposition in R (make_grm): pass Ki / Kival to Bayes() instead")
    windindx = None
    if windsize is not None or windnum is not None:
        if method in ("BayesA", "BayesRR", "BayesL"):
            raise RuntimeError("can not implement GWAS analysis for the method: " + method)   # :212-213
        if map_chr is None or map_pos is None:
            raise RuntimeError("map information must be provided.")              # :214-215
        pos = np.asarray(map_pos, dtype=np.float64)
        if np.isnan(pos).any():
            raise RuntimeError("NAs are not allowed in physical position.")
        if (pos == 0).any():
            raise RuntimeError("0 is not allowed in physical position.")
        names = [str(v) for v in np.asarray(map_chr).tolist()]
        if any(_is_zero_chromosome(v) for v in np.asarray(map_chr).tolist()):
            raise RuntimeError("0 is not allowed in chromosome.")
        # numeric chromosome codes; names that are not numbers follow the largest number (:234-243)
        def num(v):
            try:
                return float(v)
            except ValueError:
                return None
        vals = [num(v) for v in names]
        mx = max([v for v in vals if v is not None], default=0.0)
        extra = {}
        for v, nm in zip(vals, names):
            if v is None and nm not in extra:
                extra[nm] = mx + len(extra) + 1
        codes = np.array([extra[nm] if v is None else v for v, nm in zip(vals, names)], dtype=np.float64)
        windindx = cutwind(codes, pos, windsize=windsize, windnum=windnum)
    if niter is None:
        niter = 50000 if method == "BayesR" else 20000                           # :262-264
    if nburn is None:
        nburn = 30000 if method == "BayesR" else 12000                           # :265-267
    if thin >= niter - nburn:
        raise RuntimeError("bad setting for collecting frequency 'thin'.")       # :268
    if Pi is None:                                                               # :270-277
        if method == "BayesR":
            Pi = [0.95, 0.02, 0.02, 0.01]
            if fold is None:
                fold = [0, 0.0001, 0.001, 0.01]
        else:
            Pi = [0.95, 0.05]
    return dict(model=method, Pi=list(Pi), fold=None if fold is None else list(fold), niter=int(niter), nburn=int(nburn),
                thin=int(thin), windindx=windindx)


def ibrm(y, M, method="BayesCpi", X=None, R=None, map_chr=None, map_pos=None, Pi=None, fold=None, niter=None, nburn=None,
         thin=5, windsize=None, windnum=None, dfvr=None, s2vr=None, vg=None, dfvg=None, s2vg=None, ve=None, dfve=None,
         s2ve=None, printfreq=100, seed=666666, verbose=False, device=0):
    """ibrm() of the reference (R/bayes.r:121-320) after the formula has been resolved: y (NaN = no record) in the row
    order of M (matrix or BedGeno), X fixed-effect design columns, R level codes of the environmental random effects.
    Individuals without a record are predicted: g covers all rows of M (:303-308)."""
    a = ibrm_plan(method, Pi, fold, niter, nburn, thin, windsize, windnum, map_chr, map_pos)
    y = np.asarray(y, dtype=np.float64)
    n_all = M.shape[0]
    if y.shape[0] != n_all:
        raise RuntimeError("number of individuals mismatched in 'M' and 'M.id'.")   # :157
    has = ~np.isnan(y)
    rows = np.flatnonzero(has).astype(np.int32)

    def take(sel):
        if isinstance(M, BedGeno):
            base = np.arange(M.nid, dtype=np.int32) if M.rows is None else M.rows
            return BedGeno(M.image, M.nid, M.m, rows=base[sel], impute=M.impute, mode="D" if M.dominance else "A")
        return np.asfortranarray(np.asarray(M)[sel, :])

    Mfit = M if has.all() else take(rows)
    res = Bayes(y[has], Mfit, a["model"], a["Pi"], C_=None if X is None else np.asarray(X)[has], R=None if R is None else np.asarray(R)[has],
                fold=a["fold"], niter=a["niter"], nburn=a["nburn"], thin=a["thin"], dfvr=dfvr, s2vr=s2vr, vg=vg, dfvg=dfvg,
                s2vg=s2vg, ve=ve, dfve=dfve, s2ve=s2ve, windindx=a["windindx"], outfreq=printfreq,
                verbose=verbose and printfreq > 0, seed=seed, device=device)
    # gebv of every individual of M, with or without a record: the mean over the stored samples of M %*% alpha
    # (R/bayes.r:303-308) is M times the posterior mean of alpha
    e = Engine(n_all, M.shape[1], device=device)
    try:
        e.load_geno(M)
        g = e.predict(res["alpha"])
    finally:
        e.close()
    res["g"] = g
    return res

SBRM_METHODS = ("BayesB", "BayesA", "BayesL", "BayesRR", "BayesBpi", "BayesC", "BayesCpi", "BayesR", "CG")


def sbrm_plan(sumstat, ldm, method="BayesB", Pi=None, fold=None, niter=None, nburn=None, thin=5, windindx=None):
    """The argument handling of sbrm() (R/sbayes.r:126-215) up to the call of SBayesD()/SBayesS(): which kernel
    (dense or sparse LD), the method's default chain length and mixture, the COJO columns that become `sumstat`.
    sumstat: the 8-column COJO table (SNP, A1, A2, MAF, BETA, SE, P, NMISS) as a 2-D array of floats (the three name
    columns may hold anything numeric) or an m x 4 array already reduced to (MAF, BETA, SE, NMISS)."""
    import scipy.sparse as sp
    if isinstance(ldm, np.ndarray) and ldm.ndim == 2:
        sparse = False                                             # :126-127
    elif sp.issparse(ldm):
        sparse = True                                              # :128-129 (dgCMatrix)
    else:
        raise RuntimeError("Unrecognized type of ldm.")            # :131
    if method not in SBRM_METHODS:
        raise ValueError("'arg' should be one of " + ", ".join(SBRM_METHODS))   # match.arg, :134
    if windindx is not None and method in ("BayesA", "BayesRR", "BayesL"):
        raise RuntimeError("can not implement GWAS analysis for the method: " + method)   # :136-137
    if method == "CG":
        raise NotImplementedError("the conjugate-gradient solver (conjgt_den/conjgt_spa) is not part of this build")
    if niter is None:
        niter = 50000 if method == "BayesR" else 20000             # :186-188
    if nburn is None:
        nburn = 30000 if method == "BayesR" else 12000             # :189-191
    if thin >= niter - nburn:
        raise RuntimeError("bad setting for collecting frequency 'thin'.")   # :192
    if Pi is None:                                                 # :194-201
        if method == "BayesR":
            Pi = [0.95, 0.02, 0.02, 0.01]
            if fold is None:
                fold = [0, 0.0001, 0.001, 0.01]
        else:
            Pi = [0.95, 0.05]
    ss = np.asarray(sumstat, dtype=np.float64)
    if ss.ndim != 2 or ss.shape[1] not in (4, 8):
        raise RuntimeError("Inappropriate summary data format.")
    if ss.shape[1] == 8:
        ss = ss[:, [3, 4, 5, 7]]                                   # :205: columns 4, 5, 6, 8
    return dict(sparse=sparse, sumstat=np.asfortranarray(ss), model=method, Pi=list(Pi), fold=None if fold is None else list(fold),
                niter=int(niter), nburn=int(nburn), thin=int(thin), windindx=windindx)


def sbrm(sumstat, ldm, method="BayesB", Pi=None, fold=None, niter=None, nburn=None, thin=5, windindx=None, vg=None, dfvg=None,
         s2vg=None, ve=None, dfve=None, s2ve=None, printfreq=100, seed=666666, verbose=False, device=0):
    """sbrm() of the reference (R/sbayes.r:101-239) for the Bayesian methods, without the map/window bookkeeping
    (pass `windindx` directly): dispatches to SBayesD() for a dense LD matrix and SBayesS() for a sparse one."""
    a = sbrm_plan(sumstat, ldm, method, Pi, fold, niter, nburn, thin, windindx)
    fn = SBayesS if a["sparse"] else SBayesD
    return fn(a["sumstat"], ldm, a["model"], a["Pi"], niter=a["niter"], nburn=a["nburn"], thin=a["thin"], fold=a["fold"],
              windindx=a["windindx"], vg=vg, dfvg=dfvg, s2vg=s2vg, ve=ve, dfve=dfve, s2ve=s2ve, outfreq=printfreq,
              verbose=verbose and printfreq > 0, seed=seed, device=device)


class Engine:
    """Thin handle on hb_engine_* (one per GPU)."""

    def __init__(self, n, m, device=0, tile_snps=0, lag_tiles=0, n_slabs=0, seed=0, rank=0, world=1):
        self.L = _lib.load_library()
        cfg = _lib.EngineConfig(device, n, m, tile_snps, lag_tiles, n_slabs, seed, rank, world)
        self.h = C.c_void_p()
        _lib.check(self.L.hb_engine_create(C.byref(cfg), C.byref(self.h)))
        self.n, self.m = n, m

    def close(self):
        if self.h:
            self.L.hb_engine_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def describe(self):
        a, b, c, d = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        g, h = C.c_uint64(), C.c_uint64()
        _lib.check(self.L.hb_engine_describe(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(d), C.byref(g), C.byref(h)))
        return {"n_slabs": a.value, "rows_per_slab": b.value, "tile_snps": c.value, "lag_tiles": d.value,
                "geno_bytes": g.value, "gram_bytes": h.value}

    def load_geno(self, X):
        if isinstance(X, BedGeno):
            _lib.check(self.L.hb_engine_load_bed(self.h, X.image.ctypes.data, X.image.shape[0], X.nid, _ptr(X.rows),
                                                 int(X.impute), int(X.dominance)))
            return
        X = np.asarray(X)
        if X.dtype == np.int8:
            Xf = np.asfortranarray(X)
            _lib.check(self.L.hb_engine_load_geno_i8(self.h, Xf.ctypes.data, Xf.shape[0]))
        else:
            Xf = np.asfortranarray(X, dtype=np.float64)
            _lib.check(self.L.hb_engine_load_geno_f64(self.h, Xf.ctypes.data, Xf.shape[0]))

    def synth_geno(self, seed, row_offset=0):
        _lib.check(self.L.hb_engine_synth_geno(self.h, seed, row_offset))

    def col_stats(self):
        xpx, sumx = np.zeros(self.m), np.zeros(self.m)
        _lib.check(self.L.hb_engine_col_stats(self.h, xpx.ctypes.data, sumx.ctypes.data))
        return xpx, sumx

    def set_snp_info(self, xpx, active):
        xpx = np.ascontiguousarray(xpx, dtype=np.float64)
        active = np.ascontiguousarray(active, dtype=np.uint8)
        _lib.check(self.L.hb_engine_set_snp_info(self.h, xpx.ctypes.data, active.ctypes.data))

    def build_gram(self):
        _lib.check(self.L.hb_engine_build_gram(self.h))

    def ipc_handle(self):
        buf = C.create_string_buffer(64)
        _lib.check(self.L.hb_engine_ipc_handle(self.h, buf))
        return buf.raw

    def set_peers(self, handles):
        _lib.check(self.L.hb_engine_set_peers(self.h, C.c_char_p(handles)))

    def gram_device(self):
        ptr, cnt = C.c_void_p(), C.c_uint64()
        _lib.check(self.L.hb_engine_gram_device(self.h, C.byref(ptr), C.byref(cnt)))
        return ptr.value, cnt.value

    def u_centered_sums(self, mean):
        a, b = C.c_double(), C.c_double()
        _lib.check(self.L.hb_engine_u_centered_sums(self.h, mean, C.byref(a), C.byref(b)))
        return a.value, b.value

    def get_gram(self):
        d = self.describe()
        B, D = d["tile_snps"], d["lag_tiles"]
        T = (self.m + B - 1) // B
        out = np.zeros((T, D, B, B), dtype=np.int32)
        _lib.check(self.L.hb_engine_get_gram(self.h, out.ctypes.data))
        return out

    def _set(self, name, arr, dtype=np.float64):
        arr = np.ascontiguousarray(arr, dtype=dtype)
        _lib.check(getattr(self.L, "hb_engine_" + name)(self.h, arr.ctypes.data))

    def _get(self, name, count, dtype=np.float64):
        out = np.zeros(count, dtype=dtype)
        _lib.check(getattr(self.L, "hb_engine_" + name)(self.h, out.ctypes.data))
        return out

    def set_residual(self, r): self._set("set_residual", r)
    def get_residual(self): return self._get("get_residual", self.n)
    def set_u(self, u): self._set("set_u", u)
    def get_u(self): return self._get("get_u", self.n)
    def set_effects(self, g): self._set("set_effects", g)
    def get_effects(self): return self._get("get_effects", self.m)
    def get_tracker(self): return self._get("get_tracker", self.m, np.int32)
    def set_vargL(self, v): self._set("set_vargL", v)
    def get_effect_sums(self): return self._get("get_effect_sums", self.m)

    def predict(self, alpha):
        alpha = np.ascontiguousarray(alpha, dtype=np.float64)
        out = np.zeros(self.n)
        _lib.check(self.L.hb_engine_predict(self.h, alpha.ctypes.data, out.ctypes.data))
        return out

    def predict_samples(self, alpha_samples):
        """X @ alpha_samples (m x n_records) -> n x n_records: `M %*% res$MCMCsamples$alpha`, R/bayes.r:303-304."""
        A = np.asfortranarray(alpha_samples, dtype=np.float64)
        if A.ndim != 2 or A.shape[0] != self.m:
            raise ValueError("alpha_samples must be m x n_records")
        out = np.zeros((self.n, A.shape[1]), order="F")
        _lib.check(self.L.hb_engine_predict_samples(self.h, A.ctypes.data, self.m, A.shape[1], out.ctypes.data, self.n))
        return out

    def last_predict_ms(self):
        ms = C.c_float(0)
        _lib.check(self.L.hb_engine_last_predict_ms(self.h, C.byref(ms)))
        return ms.value

    def sweep(self, iter, model_index, vare, logpi, vara_fold, fold=None, dfvara=4.0, s2varg=0.0, lambda_=0.0, lambda2=0.0,
              mu_shift=0.0, rnorm2_bound=1.0):
        si = _lib.SweepIn()
        F = len(logpi)
        si.iter, si.model_index, si.n_fold = iter, model_index, F
        for k in range(F):
            si.logpi[k] = logpi[k]
            si.vara_fold[k] = vara_fold[k]
            si.fold[k] = 0.0 if fold is None else fold[k]
        si.vare, si.dfvara, si.s2varg, si.lambda_, si.lambda2 = vare, dfvara, s2varg, lambda_, lambda2
        si.mu_shift, si.rnorm2_bound = mu_shift, rnorm2_bound
        so = _lib.SweepOut()
        _lib.check(self.L.hb_engine_sweep(self.h, C.byref(si), C.byref(so)))
        return {"count": list(so.count), "varg_acc": so.varg_acc, "sum_vargL": so.sum_vargL, "sum_r": so.sum_r,
                "sum_r2": so.sum_r2, "sum_u": so.sum_u, "var_u": so.var_u, "n_changed": so.n_changed, "status": so.status, "rounds": so.rounds}

    def last_sweep_ms(self):
        a, b, c = C.c_float(), C.c_float(), C.c_float()
        _lib.check(self.L.hb_engine_last_sweep_ms(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def set_windows(self, windindx):
        if windindx is None:
            _lib.check(self.L.hb_engine_set_windows(self.h, None))
        else:
            self._set("set_windows", windindx, np.int32)

    def accumulate_pip(self): _lib.check(self.L.hb_engine_accumulate_pip(self.h))
    def accumulate_effects(self): _lib.check(self.L.hb_engine_accumulate_effects(self.h))

    def get_pip_counts(self, nw=0):
        nz = np.zeros(self.m)
        wp = np.zeros(nw)
        _lib.check(self.L.hb_engine_get_pip_counts(self.h, nz.ctypes.data, wp.ctypes.data if nw else None, nw))
        return nz, wp
