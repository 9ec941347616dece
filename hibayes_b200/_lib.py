"""ctypes binding of libhibayes_b200.so (the C ABI declared in include/hibayes_b200.h)."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
ALLREDUCE_F64 = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_double), C.c_size_t)
ALLREDUCE_I32_DEV = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t)
ALLGATHER_BYTES = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t)
_LIB = None
MAX_FOLD = 8

# every symbol include/hibayes_b200.h declares
SYMBOLS = [
    "hb_last_error", "hb_device_count", "hb_engine_create", "hb_engine_destroy", "hb_engine_load_geno_i8",
    "hb_engine_load_geno_f64", "hb_engine_synth_geno", "hb_synth_geno_host", "hb_synth_geno_host_cols", "hb_engine_col_stats",
    "hb_engine_set_snp_info", "hb_engine_build_gram", "hb_engine_get_gram", "hb_engine_set_residual", "hb_engine_get_residual",
    "hb_engine_set_u", "hb_engine_get_u", "hb_engine_set_effects", "hb_engine_get_effects", "hb_engine_get_tracker",
    "hb_engine_set_vargL", "hb_engine_sweep", "hb_engine_set_windows", "hb_engine_accumulate_pip",
    "hb_engine_get_pip_counts", "hb_engine_accumulate_effects", "hb_engine_get_effect_sums", "hb_engine_predict",
    "hb_engine_last_sweep_ms", "hb_engine_describe", "hb_bayes",
    "hb_engine_ipc_handle", "hb_engine_set_peers", "hb_engine_gram_device", "hb_engine_u_centered_sums",
    "hb_test_class_thresholds", "hb_test_class_of", "hb_test_class_batch_device", "hb_ld_engine_create", "hb_ld_engine_destroy", "hb_ld_engine_load_dense", "hb_ld_engine_load_csc", "hb_ld_engine_describe", "hb_ld_engine_set_state",
    "hb_ld_engine_set_vargL", "hb_ld_engine_set_sparse_info", "hb_ld_engine_get", "hb_ld_engine_sweep", "hb_sbayesd", "hb_sbayess",
    "hb_engine_load_bed", "hb_ldmat_create", "hb_ldmat_destroy", "hb_ldmat_load_i8", "hb_ldmat_load_bed", "hb_ldmat_stats",
    "hb_ldmat_dense", "hb_ldmat_sparse", "hb_ldmat_sparse_get", "hb_ldmat_set_panel_cols", "hb_ldmat_last_ms", "hb_bed_decode",
    "hb_test_bed_decode_snp", "hb_engine_predict_samples", "hb_test_ld_entries", "hb_test_ld_stats", "hb_test_limb_dot", "hb_cutwind_by_bp", "hb_cutwind_by_num",
    "hb_engine_last_predict_ms", "hb_engine_device_state", "hb_fx_create", "hb_fx_destroy", "hb_fx_dot", "hb_fx_self_dot", "hb_fx_axpy", "hb_fx_level_sums",
    "hb_fx_level_apply", "hb_fx_eps_set_counts", "hb_fx_eps_rhs", "hb_fx_eps_set_rhs", "hb_fx_eps_sample", "hb_fx_eps_accumulate",
    "hb_fx_eps_get", "hb_fx_describe", "hb_fx_k_step", "hb_fx_k_accumulate", "hb_fx_k_ghat_vec", "hb_engine_xt_vec",
]


class BedSource(C.Structure):
    _fields_ = [("file", C.c_void_p), ("len", C.c_size_t), ("nid", C.c_int), ("rows", C.c_void_p), ("impt", C.c_int),
                ("dominance", C.c_int)]


class SBayesArgs(C.Structure):
    _fields_ = [("m", C.c_int), ("sumstat", C.c_void_p), ("ldm", C.c_void_p), ("model", C.c_char_p), ("n_fold", C.c_int),
                ("Pi", C.c_void_p), ("fold", C.c_void_p), ("niter", C.c_int), ("nburn", C.c_int), ("thin", C.c_int),
                ("vg", C.c_double), ("dfvg", C.c_double), ("s2vg", C.c_double), ("ve", C.c_double), ("dfve", C.c_double),
                ("s2ve", C.c_double), ("windindx", C.c_void_p), ("outfreq", C.c_int), ("verbose", C.c_int),
                ("seed", C.c_uint64), ("device", C.c_int),
                ("ld_colptr", C.c_void_p), ("ld_rowidx", C.c_void_p), ("ld_val", C.c_void_p)]


class SBayesOut(C.Structure):
    _fields_ = [("Vg", C.c_double), ("Ve", C.c_double), ("h2", C.c_double), ("alpha", C.c_void_p), ("pi", C.c_void_p),
                ("pip", C.c_void_p), ("gwas", C.c_void_p), ("vara_store", C.c_void_p), ("vare_store", C.c_void_p),
                ("hsq_store", C.c_void_p), ("pi_store", C.c_void_p), ("alpha_store", C.c_void_p),
                ("tracker_final", C.c_void_p), ("nzrate_count", C.c_void_p), ("wppa_count", C.c_void_p),
                ("nnz_trace", C.c_void_p), ("vara_trace", C.c_void_p), ("vare_trace", C.c_void_p), ("varg_trace", C.c_void_p),
                ("r_hat_final", C.c_void_p), ("n_records_done", C.c_int), ("nzct", C.c_int), ("iters_done", C.c_int),
                ("n_used", C.c_int), ("seconds_sweep", C.c_double),
                ("columns_total", C.c_longlong), ("ld_entries_total", C.c_longlong), ("ld_bytes_device", C.c_longlong),
                ("rounds_total", C.c_longlong), ("tiles_total", C.c_longlong)]

ALLREDUCE_F64 = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_double), C.c_size_t)
ALLREDUCE_I32_DEV = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t)
ALLGATHER_BYTES = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t)


class EngineConfig(C.Structure):
    _fields_ = [("device", C.c_int), ("n", C.c_int), ("m", C.c_int), ("tile_snps", C.c_int), ("lag_tiles", C.c_int),
                ("n_slabs", C.c_int), ("seed", C.c_uint64), ("rank", C.c_int), ("world", C.c_int)]


class SweepIn(C.Structure):
    _fields_ = [("iter", C.c_int), ("model_index", C.c_int), ("n_fold", C.c_int),
                ("fold", C.c_double * MAX_FOLD), ("logpi", C.c_double * MAX_FOLD), ("vara_fold", C.c_double * MAX_FOLD),
                ("vare", C.c_double), ("dfvara", C.c_double), ("s2varg", C.c_double),
                ("lambda_", C.c_double), ("lambda2", C.c_double), ("mu_shift", C.c_double), ("rnorm2_bound", C.c_double)]


class SweepOut(C.Structure):
    _fields_ = [("count", C.c_double * MAX_FOLD), ("varg_acc", C.c_double), ("sum_vargL", C.c_double),
                ("sum_r", C.c_double), ("sum_r2", C.c_double), ("sum_u", C.c_double), ("var_u", C.c_double),
                ("n_changed", C.c_int), ("status", C.c_int), ("rounds", C.c_int), ("reserved", C.c_int)]


class BayesArgs(C.Structure):
    _fields_ = [
        ("n", C.c_int), ("m", C.c_int), ("y", C.c_void_p), ("X", C.c_void_p), ("x_type", C.c_int),
        ("model", C.c_char_p), ("n_fold", C.c_int), ("Pi", C.c_void_p), ("fold", C.c_void_p),
        ("nc", C.c_int), ("C", C.c_void_p), ("nr", C.c_int), ("Rlev", C.c_void_p), ("nlev", C.c_void_p),
        ("niter", C.c_int), ("nburn", C.c_int), ("thin", C.c_int),
        ("dfvr", C.c_double), ("s2vr", C.c_double), ("vg", C.c_double), ("dfvg", C.c_double),
        ("s2vg", C.c_double), ("ve", C.c_double), ("dfve", C.c_double), ("s2ve", C.c_double),
        ("windindx", C.c_void_p), ("outfreq", C.c_int), ("verbose", C.c_int), ("seed", C.c_uint64),
        ("ne", C.c_int), ("qe", C.c_int), ("epsl_y_J", C.c_void_p), ("epsl_index", C.c_void_p),
        ("Gi_colptr", C.c_void_p), ("Gi_rowidx", C.c_void_p), ("Gi_val", C.c_void_p),
        ("device", C.c_int), ("tile_snps", C.c_int), ("lag_tiles", C.c_int), ("n_slabs", C.c_int),
        ("rank", C.c_int), ("world", C.c_int), ("n_total", C.c_longlong), ("comm_ctx", C.c_void_p),
        ("allreduce_sum_f64", ALLREDUCE_F64), ("allreduce_sum_i32_dev", ALLREDUCE_I32_DEV),
        ("allgather_bytes", ALLGATHER_BYTES),
        ("nk", C.c_int), ("Kival", C.c_void_p), ("Ki", C.c_void_p),
    ]


class BayesOut(C.Structure):
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
        ("seconds_sweep", C.c_double), ("seconds_setup", C.c_double),
        ("rounds_total", C.c_longlong), ("tiles_total", C.c_longlong),
        ("rounds_trace", C.c_void_p), ("sweep_ms_trace", C.c_void_p),
        ("vr_store", C.c_void_p), ("estR_store", C.c_void_p), ("veps_store", C.c_void_p), ("J_store", C.c_void_p),
        ("epsilon_store", C.c_void_p),
    ]


def library_path():
    return os.path.join(_HERE, "libhibayes_b200.so")


def load_library():
    """Loads libhibayes_b200.so; raises if it is missing (build with `python __graft_entry__.py`)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise RuntimeError("hibayes_b200: %s not built -- run `python __graft_entry__.py` (nvcc, sm_100a); "
                           "there is no CPU fallback" % path)
    L = C.CDLL(path)
    L.hb_last_error.restype = C.c_char_p
    for name in SYMBOLS:
        getattr(L, name)  # raises AttributeError if the ABI is incomplete
    L.hb_engine_create.argtypes = [C.POINTER(EngineConfig), C.POINTER(C.c_void_p)]
    L.hb_engine_destroy.argtypes = [C.c_void_p]
    L.hb_engine_destroy.restype = None
    L.hb_engine_load_geno_i8.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    L.hb_engine_load_geno_f64.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    L.hb_engine_synth_geno.argtypes = [C.c_void_p, C.c_uint64, C.c_int64]
    L.hb_synth_geno_host.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_int64]
    L.hb_synth_geno_host_cols.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int64]
    L.hb_engine_col_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.hb_engine_set_snp_info.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.hb_engine_build_gram.argtypes = [C.c_void_p]
    L.hb_engine_get_gram.argtypes = [C.c_void_p, C.c_void_p]
    for f in ("set_residual", "get_residual", "set_u", "get_u", "set_effects", "get_effects", "get_tracker", "set_vargL",
              "get_effect_sums", "set_windows"):
        getattr(L, "hb_engine_" + f).argtypes = [C.c_void_p, C.c_void_p]
    L.hb_engine_sweep.argtypes = [C.c_void_p, C.POINTER(SweepIn), C.POINTER(SweepOut)]
    L.hb_engine_accumulate_pip.argtypes = [C.c_void_p]
    L.hb_engine_accumulate_effects.argtypes = [C.c_void_p]
    L.hb_engine_get_pip_counts.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    L.hb_engine_predict.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.hb_engine_last_sweep_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.hb_engine_describe.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                     C.POINTER(C.c_int), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.hb_bayes.argtypes = [C.POINTER(BayesArgs), C.POINTER(BayesOut)]
    L.hb_sbayesd.argtypes = [C.POINTER(SBayesArgs), C.POINTER(SBayesOut)]
    L.hb_sbayess.argtypes = [C.POINTER(SBayesArgs), C.POINTER(SBayesOut)]
    L.hb_engine_ipc_handle.argtypes = [C.c_void_p, C.c_void_p]
    L.hb_engine_set_peers.argtypes = [C.c_void_p, C.c_void_p]
    L.hb_engine_gram_device.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    L.hb_engine_u_centered_sums.argtypes = [C.c_void_p, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.hb_engine_predict_samples.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t]
    L.hb_test_ld_entries.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double,
                                     C.c_void_p]
    L.hb_test_ld_stats.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.hb_test_limb_dot.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.POINTER(C.c_longlong), C.POINTER(C.c_int)]
    L.hb_cutwind_by_bp.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_void_p]
    L.hb_cutwind_by_num.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.hb_engine_load_bed.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_int, C.c_int]
    L.hb_ldmat_create.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    L.hb_ldmat_destroy.argtypes = [C.c_void_p]
    L.hb_ldmat_destroy.restype = None
    L.hb_ldmat_load_i8.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    L.hb_ldmat_load_bed.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_int, C.c_int]
    L.hb_ldmat_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.hb_ldmat_dense.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_size_t]
    L.hb_ldmat_sparse.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.POINTER(C.c_longlong)]
    L.hb_ldmat_sparse_get.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.hb_ldmat_set_panel_cols.argtypes = [C.c_void_p, C.c_int]
    L.hb_ldmat_last_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    L.hb_bed_decode.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.hb_test_bed_decode_snp.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    _LIB = L
    return L


def last_error():
    return load_library().hb_last_error().decode()


def device_count():
    return load_library().hb_device_count()


def check(rc):
    if rc != 0:
        raise RuntimeError(last_error())
