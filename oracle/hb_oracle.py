"""ctypes loader for the CPU oracle (oracle/libhb_oracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm
import this module; nothing under hibayes_b200/ does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libhb_oracle.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
        _LIB.hbo_bayes.restype = C.c_int
        _LIB.hbo_last_error.restype = C.c_char_p
        _LIB.hbo_var.restype = C.c_double
        _LIB.hbo_var.argtypes = [C.c_void_p, C.c_int]
        _LIB.hbo_qnorm.restype = C.c_double
        _LIB.hbo_qnorm.argtypes = [C.c_double]
        _LIB.hbo_draw_gamma.restype = C.c_double
        _LIB.hbo_draw_gamma.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_double]
        _LIB.hbo_draw_chisq.restype = C.c_double
        _LIB.hbo_draw_chisq.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_double]
        _LIB.hbo_draw_uz.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                     C.POINTER(C.c_double), C.POINTER(C.c_double)]
        _LIB.hbo_invgauss.restype = C.c_double
        _LIB.hbo_invgauss.argtypes = [C.c_double] * 4
        _LIB.hbo_invgauss_literal_root.restype = C.c_double
        _LIB.hbo_invgauss_literal_root.argtypes = [C.c_double] * 3
        _LIB.hbo_time_sweep_fp64.restype = C.c_double
        _LIB.hbo_time_sweep_fp64.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.POINTER(C.c_double)]
    return _LIB


class Args(C.Structure):
    _fields_ = [
        ("n", C.c_int), ("m", C.c_int), ("y", C.c_void_p), ("X", C.c_void_p), ("x_is_int8", C.c_int),
        ("model", C.c_char_p), ("n_fold", C.c_int), ("Pi", C.c_void_p), ("fold", C.c_void_p),
        ("nc", C.c_int), ("C", C.c_void_p), ("nr", C.c_int), ("Rlev", C.c_void_p), ("nlev", C.c_void_p),
        ("niter", C.c_int), ("nburn", C.c_int), ("thin", C.c_int),
        ("dfvr", C.c_double), ("s2vr", C.c_double), ("vg", C.c_double), ("dfvg", C.c_double),
        ("s2vg", C.c_double), ("ve", C.c_double), ("dfve", C.c_double), ("s2ve", C.c_double),
        ("windindx", C.c_void_p), ("seed", C.c_uint64),
        ("ne", C.c_int), ("qe", C.c_int), ("epsl_y_J", C.c_void_p), ("epsl_index", C.c_void_p),
        ("Gi_colptr", C.c_void_p), ("Gi_rowidx", C.c_void_p), ("Gi_val", C.c_void_p),
        ("nk", C.c_int), ("Kival", C.c_void_p), ("Ki", C.c_void_p),
    ]


class Out(C.Structure):
    _fields_ = [
        ("Vg", C.c_double), ("Ve", C.c_double), ("h2", C.c_double), ("mu", C.c_double),
        ("Veps", C.c_double), ("J", C.c_double),
        ("beta", C.c_void_p), ("alpha", C.c_void_p), ("pi", C.c_void_p), ("pip", C.c_void_p),
        ("gwas", C.c_void_p), ("g", C.c_void_p), ("e", C.c_void_p), ("vr", C.c_void_p),
        ("estR", C.c_void_p), ("epsilon", C.c_void_p),
        ("mu_store", C.c_void_p), ("vara_store", C.c_void_p), ("vare_store", C.c_void_p),
        ("hsq_store", C.c_void_p), ("pi_store", C.c_void_p), ("alpha_store", C.c_void_p),
        ("beta_store", C.c_void_p),
        ("tracker_final", C.c_void_p), ("nzrate_count", C.c_void_p), ("wppa_count", C.c_void_p),
        ("nnz_trace", C.c_void_p), ("vara_trace", C.c_void_p), ("vare_trace", C.c_void_p),
        ("varg_trace", C.c_void_p),
        ("n_records_done", C.c_int), ("nzct", C.c_int), ("iters_done", C.c_int),
        ("seconds_sweep", C.c_double),
        ("vr_store", C.c_void_p), ("estR_store", C.c_void_p), ("veps_store", C.c_void_p), ("J_store", C.c_void_p),
        ("epsilon_store", C.c_void_p),
    ]


_REF = None
_DROPIN = None
TAPE_DTYPE = np.dtype([("kind", np.int32), ("pad", np.int32), ("value", np.float64), ("param", np.float64)])


def ref_lib():
    """oracle/_ref/libhibayes_ref.so: the reference's own Bayes.cpp / SBayesD.cpp / SBayesS.cpp / stats.cpp / solver.cpp
    compiled unmodified against the stand-in headers of oracle/ref_shim (oracle/Makefile).  Built here, where
    /root/reference exists; on the GPU box the prebuilt file is used.  None when it is not there."""
    global _REF
    if _REF is None:
        path = os.path.join(_HERE, "_ref", "libhibayes_ref.so")
        if not os.path.exists(path) and os.path.exists("/root/reference/src/Bayes.cpp"):
            build()
        if not os.path.exists(path):
            return None
        _REF = C.CDLL(path)
        _REF.hbref_last_error.restype = C.c_char_p
        for f in (_REF.hbref_bayes, _REF.hbref_sbayesd, _REF.hbref_sbayess):
            f.restype = C.c_int
            f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    return _REF


def dropin_lib():
    """oracle/_ref/libhibayes_dropin.so: the same plain-buffer harness as ref_lib(), linked with the drop-in Rcpp bodies of
    integration/rcpp/ (the reference's C++ signatures forwarding to libhibayes_b200.so) instead of the reference's files.
    Needs a GPU at run time.  Its `tape` is two uniforms: what seed_from_r() draws to make the run key."""
    global _DROPIN
    if _DROPIN is None:
        path = os.path.join(_HERE, "_ref", "libhibayes_dropin.so")
        if not os.path.exists(path):
            build()
        if not os.path.exists(path):
            return None
        _DROPIN = C.CDLL(path)
        _DROPIN.hbref_last_error.restype = C.c_char_p
        for f in (_DROPIN.hbref_bayes, _DROPIN.hbref_sbayesd, _DROPIN.hbref_sbayess):
            f.restype = C.c_int
            f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    return _DROPIN


def seed_tape(seed):
    """The two unif_rand() values from which the drop-in bodies rebuild `seed` (integration/rcpp/hb_dropin.h)."""
    t = np.zeros(2, dtype=TAPE_DTYPE)
    t["value"][0] = float((seed >> 32) & 0xFFFFFFFF) / 4294967296.0
    t["value"][1] = float(seed & 0xFFFFFFFF) / 4294967296.0
    return t


def _tape_begin(cap):
    L = lib()
    L.hbo_tape_begin.argtypes = [C.c_void_p, C.c_uint64]
    L.hbo_tape_end.restype = C.c_uint64
    buf = np.zeros(int(cap), dtype=TAPE_DTYPE)
    L.hbo_tape_begin(buf.ctypes.data, buf.shape[0])
    return buf


def _tape_end(buf):
    n = lib().hbo_tape_end()
    if n > buf.shape[0]:
        raise RuntimeError("tape buffer too small: %d draws, room for %d" % (n, buf.shape[0]))
    return buf[:n].copy()


def _run(oracle_fn, ref_fn, a, o, replay, record_cap, library="reference"):
    """One call: the oracle (optionally recording its tape) or, with replay = a tape, the compiled reference
    (library "reference") or the drop-in Rcpp bodies over the GPU library (library "dropin")."""
    if replay is not None:
        R = ref_lib() if library == "reference" else dropin_lib()
        if R is None:
            raise RuntimeError("oracle/_ref/libhibayes_%s.so is not built" % ("ref" if library == "reference" else "dropin"))
        tp = np.ascontiguousarray(replay, dtype=TAPE_DTYPE)
        used = C.c_size_t(0)
        rc = getattr(R, ref_fn)(C.addressof(a), C.addressof(o), tp.ctypes.data, tp.shape[0], C.byref(used))
        if rc != 0:
            raise RuntimeError(("reference: " if library == "reference" else "drop-in: ") + R.hbref_last_error().decode())
        return {"consumed": used.value, "tape_len": int(tp.shape[0])}
    buf = _tape_begin(record_cap) if record_cap else None
    rc = oracle_fn(C.byref(a), C.byref(o))
    tape = _tape_end(buf) if record_cap else None
    if rc != 0:
        raise RuntimeError(lib().hbo_last_error().decode())
    return {"tape": tape}


def _ptr(a):
    return None if a is None else a.ctypes.data


def _nan(v):
    return float("nan") if v is None else float(v)


def bayes(y, X, model, Pi, fold=None, C_=None, R=None, niter=200, nburn=100, thin=5,
          dfvr=None, s2vr=None, vg=None, dfvg=None, s2vg=None, ve=None, dfve=None, s2ve=None,
          windindx=None, seed=666666, epsl_y_J=None, epsl_Gi=None, epsl_index=None,
          store_alpha=False, Kival=None, Ki=None, record_tape=False, replay_on_reference=None, library="reference"):
    """Oracle twin of hibayes' C++ Bayes() (Bayes.cpp:60-88 argument list).

    X: (n, m) array, float64 or int8 (Fortran order is used internally).
    R: (n, nr) integer level codes (0-based) for environmental random effects.
    epsl_Gi: scipy.sparse matrix (qe x qe); epsl_index 1-based.
    Returns a dict named like the reference's Rcpp::List plus diagnostics.
    record_tape: also return res["tape"], the variates consumed in the order of the reference's sampler calls.
    replay_on_reference: a tape -> the SAME arguments go to the compiled reference (ref_lib()) instead of the oracle;
    the diagnostics the reference does not return stay zero, res["replay"] says how much of the tape it consumed.
    """
    L = lib()
    y = np.ascontiguousarray(y, dtype=np.float64)
    n = y.shape[0]
    X = np.asarray(X)
    if X.dtype == np.int8:
        Xf = np.asfortranarray(X)
        is8 = 1
    else:
        Xf = np.asfortranarray(X, dtype=np.float64)
        is8 = 0
    m = Xf.shape[1]
    Pi = np.ascontiguousarray(Pi, dtype=np.float64)
    F = Pi.shape[0]
    fold_a = None if fold is None else np.ascontiguousarray(fold, dtype=np.float64)
    a = Args()
    a.n, a.m, a.y, a.X, a.x_is_int8 = n, m, _ptr(y), _ptr(Xf), is8
    a.model = model.encode()
    a.n_fold, a.Pi, a.fold = F, _ptr(Pi), _ptr(fold_a)
    keep = [y, Xf, Pi, fold_a]
    nc = 0
    if C_ is not None:
        Cf = np.asfortranarray(C_, dtype=np.float64)
        nc = Cf.shape[1]
        a.C = _ptr(Cf)
        keep.append(Cf)
    a.nc = nc
    nr, n_levels = 0, 0
    if R is not None:
        Rf = np.asfortranarray(R, dtype=np.int32)
        nr = Rf.shape[1]
        nlev = np.ascontiguousarray(Rf.max(axis=0) + 1, dtype=np.int32)
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
        nw = int(w.max())
        a.windindx = _ptr(w)
        keep.append(w)
    a.seed = seed
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
        a.epsl_y_J, a.epsl_index, a.Gi_colptr, a.Gi_rowidx, a.Gi_val = _ptr(yj), _ptr(ei), _ptr(cp), _ptr(ri), _ptr(gv)
        keep += [ei, cp, ri, gv, yj]
    a.ne, a.qe = ne, qe
    if Ki is not None:   # BSLMM: eigenvectors (n x n) and eigenvalues of the relationship matrix
        Kf = np.asfortranarray(Ki, dtype=np.float64)
        kv = np.ascontiguousarray(Kival, dtype=np.float64)
        a.nk, a.Ki, a.Kival = Kf.shape[1], _ptr(Kf), _ptr(kv)
        keep += [Kf, kv]
    nrec = max((niter - nburn) // thin, 0)
    o = Out()
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
    if nr:
        mc["Vr"], mc["r"] = np.zeros((nr, nrec), order="F"), np.zeros((n_levels, nrec), order="F")
    if ne:
        mc["Veps"], mc["J"], mc["epsilon"] = np.zeros(nrec), np.zeros(nrec), np.zeros((qe, nrec), order="F")
    dg = {
        "tracker": np.zeros(m, dtype=np.int32), "nzrate_count": np.zeros(m), "wppa_count": np.zeros(nw),
        "nnz_trace": np.zeros(niter, dtype=np.int32), "vara_trace": np.zeros(niter),
        "vare_trace": np.zeros(niter), "varg_trace": np.zeros(niter),
    }
    if nr:
        o.vr_store, o.estR_store = _ptr(mc["Vr"]), _ptr(mc["r"])
    if ne:
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
    cap = niter * (3 * m + nc + n_levels + nr + qe + a.nk + F + 16) if record_tape else 0
    info = _run(L.hbo_bayes, "hbref_bayes", a, o, replay_on_reference, cap, library)
    if record_tape:
        res["tape"] = info["tape"]
    if replay_on_reference is not None:
        res["replay"] = info
    res.update({"Vg": o.Vg, "Ve": o.Ve, "h2": o.h2, "mu": o.mu, "Veps": o.Veps, "J": o.J})
    res["MCMCsamples"] = mc
    dg.update({"n_records": o.n_records_done, "nzct": o.nzct, "iters_done": o.iters_done,
               "seconds_sweep": o.seconds_sweep})
    res["diag"] = dg
    return res


def time_sweep_fp64(n, m_cpu, sweeps, threads, seed=1):
    cs = C.c_double(0)
    v = lib().hbo_time_sweep_fp64(n, m_cpu, sweeps, threads, seed, C.byref(cs))
    return v, cs.value


class _SBayesArgs(C.Structure):
    _fields_ = [("m", C.c_int), ("sumstat", C.c_void_p), ("ldm", C.c_void_p), ("model", C.c_char_p), ("n_fold", C.c_int),
                ("Pi", C.c_void_p), ("fold", C.c_void_p), ("niter", C.c_int), ("nburn", C.c_int), ("thin", C.c_int),
                ("vg", C.c_double), ("dfvg", C.c_double), ("s2vg", C.c_double), ("ve", C.c_double), ("dfve", C.c_double),
                ("s2ve", C.c_double), ("windindx", C.c_void_p), ("seed", C.c_uint64),
                ("ld_colptr", C.c_void_p), ("ld_rowidx", C.c_void_p), ("ld_val", C.c_void_p)]


class _SBayesOut(C.Structure):
    _fields_ = [("Vg", C.c_double), ("Ve", C.c_double), ("h2", C.c_double), ("alpha", C.c_void_p), ("pi", C.c_void_p),
                ("pip", C.c_void_p), ("gwas", C.c_void_p), ("vara_store", C.c_void_p), ("vare_store", C.c_void_p),
                ("hsq_store", C.c_void_p), ("pi_store", C.c_void_p), ("alpha_store", C.c_void_p),
                ("tracker_final", C.c_void_p), ("nzrate_count", C.c_void_p), ("wppa_count", C.c_void_p),
                ("nnz_trace", C.c_void_p), ("vara_trace", C.c_void_p), ("vare_trace", C.c_void_p), ("varg_trace", C.c_void_p),
                ("r_hat_final", C.c_void_p), ("n_records_done", C.c_int), ("nzct", C.c_int), ("iters_done", C.c_int),
                ("n_used", C.c_int)]


def sbayes_buffers(m, F, niter, nburn, thin, nw, out_struct):
    """Output arrays shared by the oracle wrapper and the GPU wrapper (same field names in both structs)."""
    nrec = max((niter - nburn) // thin, 0)
    res = {"alpha": np.zeros(m), "pi": np.zeros(F), "pip": np.zeros(m), "gwas": np.zeros(nw)}
    mc = {"Vg": np.zeros(nrec), "Ve": np.zeros(nrec), "h2": np.zeros(nrec), "pi": np.zeros((F, nrec), order="F")}
    dg = {"tracker": np.zeros(m, dtype=np.int32), "nzrate_count": np.zeros(m), "wppa_count": np.zeros(nw),
          "nnz_trace": np.zeros(niter, dtype=np.int32), "vara_trace": np.zeros(niter), "vare_trace": np.zeros(niter),
          "varg_trace": np.zeros(niter), "r_hat": np.zeros(m)}
    o = out_struct
    o.alpha, o.pi, o.pip = res["alpha"].ctypes.data, res["pi"].ctypes.data, res["pip"].ctypes.data
    o.gwas = res["gwas"].ctypes.data if nw else None
    o.vara_store, o.vare_store, o.hsq_store, o.pi_store = (mc["Vg"].ctypes.data, mc["Ve"].ctypes.data, mc["h2"].ctypes.data,
                                                           mc["pi"].ctypes.data)
    o.tracker_final, o.nzrate_count = dg["tracker"].ctypes.data, dg["nzrate_count"].ctypes.data
    o.wppa_count = dg["wppa_count"].ctypes.data if nw else None
    o.nnz_trace, o.vara_trace, o.vare_trace, o.varg_trace = (dg["nnz_trace"].ctypes.data, dg["vara_trace"].ctypes.data,
                                                             dg["vare_trace"].ctypes.data, dg["varg_trace"].ctypes.data)
    o.r_hat_final = dg["r_hat"].ctypes.data
    return res, mc, dg


def _sbayes(sumstat, ldm, sparse, model, Pi, fold, niter, nburn, thin, windindx, vg, dfvg, s2vg, ve, dfve, s2ve, seed,
            record_tape=False, replay_on_reference=None, store_alpha=False, library="reference"):
    L = lib()
    ss = np.asfortranarray(sumstat, dtype=np.float64)
    keep = [ss]
    a = _SBayesArgs()
    if sparse:
        import scipy.sparse as sp
        G = sp.csc_matrix(ldm)
        G.sort_indices()
        G.eliminate_zeros()
        cp = np.ascontiguousarray(G.indptr, dtype=np.int32)
        ri = np.ascontiguousarray(G.indices, dtype=np.int32)
        gv = np.ascontiguousarray(G.data, dtype=np.float64)
        a.ld_colptr, a.ld_rowidx, a.ld_val = cp.ctypes.data, ri.ctypes.data, gv.ctypes.data
        keep += [cp, ri, gv]
        m = G.shape[0]
    else:
        ld = np.asfortranarray(ldm, dtype=np.float64)
        a.ldm = ld.ctypes.data
        keep.append(ld)
        m = ld.shape[0]
    Pi = np.ascontiguousarray(Pi, dtype=np.float64)
    F = Pi.shape[0]
    fo = None if fold is None else np.ascontiguousarray(fold, dtype=np.float64)
    a.m, a.sumstat, a.model, a.n_fold, a.Pi, a.fold = m, ss.ctypes.data, model.encode(), F, Pi.ctypes.data, _ptr(fo)
    a.niter, a.nburn, a.thin = niter, nburn, thin
    a.vg, a.dfvg, a.s2vg, a.ve, a.dfve, a.s2ve = _nan(vg), _nan(dfvg), _nan(s2vg), _nan(ve), _nan(dfve), _nan(s2ve)
    nw = 0
    if windindx is not None:
        w = np.ascontiguousarray(windindx, dtype=np.int32)
        nw = int(w.max())
        a.windindx = w.ctypes.data
        keep.append(w)
    a.seed = seed
    o = _SBayesOut()
    res, mc, dg = sbayes_buffers(m, F, niter, nburn, thin, nw, o)
    fn = L.hbo_sbayess if sparse else L.hbo_sbayesd
    fn.restype = C.c_int
    if store_alpha:
        mc["alpha"] = np.zeros((m, max((niter - nburn) // thin, 0)), order="F")
        o.alpha_store = mc["alpha"].ctypes.data
    cap = niter * (103 * m + F + 16) if record_tape else 0   # (SBayesS may re-draw a SNP up to 101 times)
    cap = min(cap, 50_000_000)
    info = _run(fn, "hbref_sbayess" if sparse else "hbref_sbayesd", a, o, replay_on_reference, cap, library)
    if record_tape:
        res["tape"] = info["tape"]
    if replay_on_reference is not None:
        res["replay"] = info
    res.update({"Vg": o.Vg, "Ve": o.Ve, "h2": o.h2, "MCMCsamples": mc})
    dg.update({"n_records": o.n_records_done, "nzct": o.nzct, "iters_done": o.iters_done, "n_used": o.n_used})
    res["diag"] = dg
    return res


def sbayesd(sumstat, ldm, model, Pi, fold=None, niter=200, nburn=100, thin=5, windindx=None, vg=None, dfvg=None, s2vg=None,
            ve=None, dfve=None, s2ve=None, seed=666666, **kw):
    """CPU oracle of SBayesD(): sumstat m x 4 (MAF, BETA, SE, N), ldm m x m dense."""
    return _sbayes(sumstat, ldm, False, model, Pi, fold, niter, nburn, thin, windindx, vg, dfvg, s2vg, ve, dfve, s2ve, seed, **kw)


def sbayess(sumstat, ldm, model, Pi, fold=None, niter=200, nburn=100, thin=5, windindx=None, vg=None, dfvg=None, s2vg=None,
            ve=None, dfve=None, s2ve=None, seed=666666, **kw):
    """CPU oracle of SBayesS(): ldm a scipy sparse matrix (or anything csc_matrix() accepts)."""
    return _sbayes(sumstat, ldm, True, model, Pi, fold, niter, nburn, thin, windindx, vg, dfvg, s2vg, ve, dfve, s2ve, seed, **kw)


# ---- .bed decoder and LD builder (oracle/hb_oracle_ld.c) ----
def read_bed(image, nid, m, impute=True, dominance=False, na_code=-128):
    """CPU oracle of read_bed<char>() (/root/reference/src/read_bed.cpp:97-232): (nid x m int8 F-order, miss flags)."""
    L = lib()
    img = np.ascontiguousarray(image, dtype=np.uint8)
    out = np.zeros((nid, m), dtype=np.int8, order="F")
    miss = np.zeros(m, dtype=np.uint8)
    L.hbo_read_bed.restype = C.c_int
    L.hbo_read_bed.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    if L.hbo_read_bed(img.ctypes.data, img.shape[0], nid, m, int(impute), int(dominance), na_code, out.ctypes.data,
                      miss.ctypes.data) != 0:
        raise RuntimeError("hbo_read_bed: file image too short")
    return out, miss


def bigstat(X):
    """CPU oracle of BigStat<char>() (/root/reference/src/tXXmat.cpp:43-77)."""
    L = lib()
    Xf = np.asfortranarray(X, dtype=np.int8)
    n, m = Xf.shape
    mean, sm, xx = np.zeros(m), np.zeros(m), np.zeros(m)
    L.hbo_bigstat.restype = None
    L.hbo_bigstat.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.hbo_bigstat(Xf.ctypes.data, n, n, m, mean.ctypes.data, sm.ctypes.data, xx.ctypes.data)
    return {"mean": mean, "sum": sm, "xx": xx}


def txxmat(X, chr=None, chisq=None):
    """CPU oracle of tXXmat_Geno<char>() (chr None; /root/reference/src/tXXmat.cpp:100-185) and tXXmat_Chr<char>()
    (:504-605): the full m x m matrix (F order); where the reference returns an arma::sp_mat its stored entries are
    the non-zeros of this matrix.  chisq None = R_NilValue (dense branch)."""
    L = lib()
    Xf = np.asfortranarray(X, dtype=np.int8)
    n, m = Xf.shape
    c = None if chr is None else np.ascontiguousarray(chr, dtype=np.int32)
    out = np.zeros((m, m), order="F")
    L.hbo_txxmat.restype = None
    L.hbo_txxmat.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_void_p]
    L.hbo_txxmat(Xf.ctypes.data, n, n, m, _ptr(c), int(chisq is not None), 0.0 if chisq is None else float(chisq),
                 out.ctypes.data)
    return out


# ---- the same three through the compiled reference (oracle/_ref/libhibayes_ref.so: tXXmat.cpp, read_bed.cpp) ----
def _ldlib(library):
    R = ref_lib() if library == "reference" else dropin_lib()
    if R is None:
        raise RuntimeError("oracle/_ref library '%s' is not built" % library)
    return R


def ref_bigstat(X, library="reference"):
    """BigStat() of the reference itself (tXXmat.cpp:43-98) on a big.matrix of type char."""
    R = _ldlib(library)
    Xf = np.asfortranarray(X, dtype=np.int8)
    n, m = Xf.shape
    mean, sm, xx = np.zeros(m), np.zeros(m), np.zeros(m)
    R.hbref_bigstat.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    if R.hbref_bigstat(Xf.ctypes.data, n, m, mean.ctypes.data, sm.ctypes.data, xx.ctypes.data) != 0:
        raise RuntimeError("reference: " + R.hbref_last_error().decode())
    return {"mean": mean, "sum": sm, "xx": xx}


def ref_txxmat(X, chr=None, chisq=None, library="reference"):
    """tXXmat_Geno() / tXXmat_Chr() of the reference itself (tXXmat.cpp:100-206, 504-626): (m x m matrix with zeros where
    the returned arma::sp_mat stores nothing, number of stored entries)."""
    R = _ldlib(library)
    Xf = np.asfortranarray(X, dtype=np.int8)
    n, m = Xf.shape
    c = None if chr is None else np.ascontiguousarray(chr, dtype=np.int32)
    out = np.zeros((m, m), order="F")
    stored = C.c_longlong(0)
    R.hbref_txxmat.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.POINTER(C.c_longlong)]
    if R.hbref_txxmat(Xf.ctypes.data, n, m, _ptr(c), int(chisq is not None), 0.0 if chisq is None else float(chisq),
                      out.ctypes.data, C.byref(stored)) != 0:
        raise RuntimeError("reference: " + R.hbref_last_error().decode())
    return out, stored.value


def ref_read_bed(path, nid, m, impute=True, dominance=False, max_line=10000, library="reference"):
    """read_bed<char>() of the reference itself (read_bed.cpp:97-247) on a .bed file: nid x m int8 (F order)."""
    R = _ldlib(library)
    out = np.zeros((nid, m), dtype=np.int8, order="F")
    R.hbref_read_bed.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_long, C.c_int, C.c_int, C.c_void_p]
    if R.hbref_read_bed(path.encode(), nid, m, max_line, int(impute), int(dominance), out.ctypes.data) != 0:
        raise RuntimeError("reference: " + R.hbref_last_error().decode())
    return out
