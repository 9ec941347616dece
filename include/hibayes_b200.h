/*
 * hibayes_b200.h -- C ABI of the B200-native single-site Gibbs engine that replaces the
 * per-SNP sweep of hibayes' Bayes() / SBayesD() / SBayesS().
 *
 * Two layers, both plain C (pointers + sizes, caller-owned memory, int status codes,
 * nothing thrown across the boundary; hb_last_error() gives the message):
 *
 *   1. hb_bayes()        -- drop-in for the body of  Rcpp::List Bayes(...)
 *                           (/root/reference/src/Bayes.cpp:60-88 signature, :919-1040 return
 *                           list).  The Rcpp shim _hibayes_Bayes (src/RcppExports.cpp:16-50)
 *                           unpacks its 27 SEXPs exactly as today and forwards plain pointers;
 *                           see INTEGRATION.md for the stub.
 *   2. hb_engine_*()     -- the device engine the host driver is written against: load the
 *                           genotype matrix once, then one hb_engine_sweep() per MCMC iteration
 *                           replaces the switch(model_index) block Bayes.cpp:586-816 and the
 *                           reductions at :480,:819,:823; non-SNP effects stay on the host and
 *                           exchange the residual through hb_engine_{get,set}_residual().
 *
 * All functions return 0 on success.  A handle is not thread-safe; calls are synchronous.
 */
#ifndef HIBAYES_B200_H
#define HIBAYES_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HB_MAX_FOLD 8

/* model_index as decoded at Bayes.cpp:97 / SBayesD.cpp:28 */
enum {
  HB_MODEL_RR = 1, HB_MODEL_A = 2, HB_MODEL_B = 3, HB_MODEL_C = 4, HB_MODEL_L = 5, HB_MODEL_R = 6
};

const char* hb_last_error(void);
/* number of CUDA devices visible; <0 on driver error */
int hb_device_count(void);

/* ------------------------------------------------------------------ engine layer */
typedef struct hb_engine hb_engine;

typedef struct {
  int device;        /* CUDA ordinal */
  int n;             /* individuals held by this engine (rows of X, y) */
  int m;             /* SNPs (columns of X) */
  int tile_snps;     /* B: SNPs per tile step (64, 128 or 256); 0 = default (256) */
  int lag_tiles;     /* D: tiles in flight between a dot and its residual update (1..8); 0 = default (5) */
  int n_slabs;       /* row slabs = streaming CTAs; 0 = default (SM count - 1, fewer for small n) */
  uint64_t seed;     /* Philox run key (hb_rng.h) */
  /* row sharding across ranks (one engine per GPU); world = 1 for a single GPU */
  int rank, world;
} hb_engine_config;

int hb_engine_create(const hb_engine_config* cfg, hb_engine** out);
void hb_engine_destroy(hb_engine* e);

/* Genotypes: column-major n x m host matrix, leading dimension ld (>= n), values {0,1,2}
 * (what read_bed.cpp:116-120 produces after imputation).  Borrowed for the call only.
 * Replaces the `arma::mat& X` argument, Bayes.cpp:62.  _f64 accepts the R numeric matrix. */
int hb_engine_load_geno_i8(hb_engine* e, const int8_t* X, size_t ld);
int hb_engine_load_geno_f64(hb_engine* e, const double* X, size_t ld);
/* Genotypes straight from the image of a SNP-major PLINK .bed file (all `len` bytes incl. the three magic bytes;
 * nid individuals per SNP, m = the engine's SNP count): decoded on the device into the tile layout with the
 * reference's code map and major-genotype imputation (read_bed<char>(), /root/reference/src/read_bed.cpp:97-232),
 * which removes the big.matrix -> R numeric matrix detour of R/bayes.r:284.  rows: NULL (nid == n, file order) or n
 * 0-based file individuals, output row i = individual rows[i] (the `M[index, ]` selection of R/bayes.r:281-291);
 * the major genotype is counted over all nid individuals of the file, as the reference does at read time.
 * impt = 0 refuses SNPs with missing genotypes (the engine holds {0,1,2} only); dominance = the reader's `d`. */
int hb_engine_load_bed(hb_engine* e, const uint8_t* file, size_t len, int nid, const int32_t* rows, int impt,
                       int dominance);
/* Synthetic genotypes generated on the device: p_j ~ U(0.05,0.5),
 * x_ij ~ Binomial(2,p_j), addressed by (seed, global row, column) so that any row sharding
 * yields the same matrix; row_offset is this rank's first global row.
 * hb_synth_geno_host() writes the same matrix on the host (column-major int8). */
int hb_engine_synth_geno(hb_engine* e, uint64_t seed, int64_t row_offset);
int hb_synth_geno_host(int8_t* X, int n, int m, uint64_t seed, int64_t row_offset);
/* columns col0 .. col0 + ncols - 1 of the same matrix (X: n x ncols); lets a caller fill a large matrix from several threads */
int hb_synth_geno_host_cols(int8_t* X, int n, int col0, int ncols, uint64_t seed, int64_t row_offset);

/* Column statistics, Bayes.cpp:310-317: xpx_j = sum x^2, sumx_j = sum x (exact integers
 * returned as doubles; the caller forms var(x_j)).  Local rows only when world > 1. */
int hb_engine_col_stats(hb_engine* e, double* xpx, double* sumx);
/* Marks SNPs the sweep skips (vx == 0, Bayes.cpp:589) and sets the global xpx. */
int hb_engine_set_snp_info(hb_engine* e, const double* xpx_global, const uint8_t* active);
/* One-off band Gram blocks X_t' [X_t .. X_{t+D-1}] (exact int32) used to chain the tiles. */
int hb_engine_build_gram(hb_engine* e);
/* copies the band out: int32 [tiles][lag][tile_snps][tile_snps] (tests, diagnostics) */
int hb_engine_get_gram(hb_engine* e, int32_t* out);

int hb_engine_set_residual(hb_engine* e, const double* yadj); /* n doubles */
int hb_engine_get_residual(hb_engine* e, double* yadj);
int hb_engine_set_u(hb_engine* e, const double* u);
int hb_engine_get_u(hb_engine* e, double* u);
int hb_engine_set_effects(hb_engine* e, const double* g);      /* m doubles */
int hb_engine_get_effects(hb_engine* e, double* g);
int hb_engine_get_tracker(hb_engine* e, int32_t* tracker);     /* m, class label per SNP */
int hb_engine_set_vargL(hb_engine* e, const double* vargL);    /* BayesL per-SNP variances */

typedef struct {
  int iter;                    /* MCMC iteration (draw address) */
  int model_index;             /* HB_MODEL_* */
  int n_fold;                  /* mixture components (2 for B/C, F for R) */
  double fold[HB_MAX_FOLD];    /* BayesR fold_ (Bayes.cpp:108-114) */
  double logpi[HB_MAX_FOLD];   /* log(Pi) */
  double vara_fold[HB_MAX_FOLD]; /* BayesR: varg*fold_k; C/RR: [1] = varg */
  double vare;                 /* residual variance */
  double dfvara, s2varg;       /* per-SNP inv-chi^2 prior (BayesA/B, :613,:636) */
  double lambda, lambda2;      /* BayesL (:729) */
  double mu_shift;             /* added to every residual entry before the sweep (:482) */
  double rnorm2_bound;         /* >= yadj'yadj at sweep start; sizes the fixed-point dot scale */
} hb_sweep_in;

typedef struct {
  double count[HB_MAX_FOLD];   /* SNPs per class after the sweep (fold_snp_num, :803-805) */
  double varg_acc;             /* model 4: sum g^2 (:698); 6: sum g^2/fold (:791); 1: g'g (:603) */
  double sum_vargL;            /* model 5: sum(vargL) (:739) */
  double sum_r, sum_r2;        /* sum(yadj), yadj'yadj (:480,:823) */
  double sum_u, var_u;         /* var(u) with n-1 (:819) */
  int n_changed;               /* SNPs whose effect changed (= residual updates applied) */
  int status;                  /* 0 ok; else device-side abort code */
  int rounds;                  /* speculation rounds of the scalar chain summed over tiles (>= tiles) */
  int reserved;
} hb_sweep_out;

int hb_engine_sweep(hb_engine* e, const hb_sweep_in* in, hb_sweep_out* out);

/* PIP / WPPA counters on the device, Bayes.cpp:826-845 (windindx 1-based, 0/NULL = none) */
int hb_engine_set_windows(hb_engine* e, const int32_t* windindx);
int hb_engine_accumulate_pip(hb_engine* e);
int hb_engine_get_pip_counts(hb_engine* e, double* nzrate, double* wppa, int nw);
/* running sum of g over recorded iterations (posterior mean of alpha, :970) */
int hb_engine_accumulate_effects(hb_engine* e);
int hb_engine_get_effect_sums(hb_engine* e, double* gsum);

/* out[i] = sum_j X[i][j] * alpha[j] over the resident genotypes (the X*g of Bayes.cpp:971) */
int hb_engine_predict(hb_engine* e, const double* alpha, double* out);
/* `M %*% MCMCsamples$alpha` of R/bayes.r:303-304: out (n x n_records, ld_out) = X * alpha (m x n_records, ld_alpha) */
int hb_engine_predict_samples(hb_engine* e, const double* alpha, size_t ld_alpha, int n_records, double* out, size_t ld_out);
/* device time of the batched kernel(s) of the last hb_engine_predict_samples() call, ms */
int hb_engine_last_predict_ms(hb_engine* e, float* ms);

/* Row sharding over the GPUs of one node: one process and one engine per GPU, created with
 * (rank, world) and this rank's rows.  The x_j'r of Bayes.cpp:593 becomes a sum over ranks: inside the sweep
 * kernel every rank adds its exact fixed-point part of a tile's dots to every rank's accumulators with NVLink
 * peer atomics, so all ranks see identical dots and take identical decisions (no broadcast, no host round trip).
 * hb_engine_ipc_handle() exports this rank's accumulators (64-byte CUDA IPC handle); the caller exchanges the
 * handles (e.g. torch.distributed.all_gather) and passes all `world` of them, rank-ordered, to
 * hb_engine_set_peers().  Set-up sums (xpx, sumx, the Gram band) and the per-iteration scalars of hb_sweep_out
 * are all-reduced by the caller (NCCL): hb_engine_gram_device() exposes the int32 band on the device for that,
 * hb_engine_u_centered_sums() gives the two accumulators of var(u) (Bayes.cpp:819) about a global mean. */
int hb_engine_ipc_handle(hb_engine* e, void* handle64);
int hb_engine_set_peers(hb_engine* e, const void* handles);
int hb_engine_gram_device(hb_engine* e, void** device_ptr, uint64_t* n_int32);
int hb_engine_u_centered_sums(hb_engine* e, double mean_u, double* sum_sq, double* sum_dev);

/* timing of the last sweep's kernels in ms (prep, sweep, tail) measured with CUDA events */
int hb_engine_last_sweep_ms(hb_engine* e, float* prep_ms, float* sweep_ms, float* tail_ms);
/* description of the layout chosen: slabs, rows per slab, tile, lag, bytes of X on device */
int hb_engine_describe(hb_engine* e, int* n_slabs, int* rows_per_slab, int* tile_snps, int* lag_tiles,
                       uint64_t* geno_bytes, uint64_t* gram_bytes);

/* ------------------------------------------------------------------ non-SNP effects on the device (csrc/effects.cu)
 * The steps of an MCMC iteration of Bayes() around the SNP sweep, on the engine's own residual (yadj) and genetic values
 * (u), so that neither vector leaves the device inside the loop:
 *   covariates            /root/reference/src/Bayes.cpp:484-494
 *   env. random effects   Bayes.cpp:496-516
 *   single-step J         Bayes.cpp:555-562
 *   single-step epsilon   Bayes.cpp:563-582 with the sparse Gauss-Seidel sampler Gibbs(sp_mat), src/solver.cpp:131-140
 * The host driver (hb_bayes) owns the scalar draws and, when rows are sharded over ranks, all-reduces the returned sums. */
int hb_engine_device_state(hb_engine* e, double** r_dev, double** u_dev, void** cuda_stream, int* device, int* n);

typedef struct hb_fx hb_fx;
typedef struct {
  int n;                                             /* this rank's rows (= the engine's n) */
  int nc; const double* C;                           /* n x nc column-major */
  int nr; const int32_t* Rlev; const int32_t* nlev;  /* n x nr 0-based level codes, levels per term */
  const double* J;                                   /* NULL or n (epsl_y_J) */
  int ne, qe;                                        /* the LAST ne rows carry an epsilon; epsl_Gi is qe x qe (CSC) */
  const int32_t* epsl_index;                         /* ne entries, 1-based */
  const int32_t* Gi_colptr; const int32_t* Gi_rowidx; const double* Gi_val;
  uint64_t seed;
  int nk; const double* Ki;                          /* BSLMM: n x nk eigenvectors (nk = n), or 0 / NULL */
} hb_fx_desc;
enum { HB_FX_COV = 0, HB_FX_J = 1, HB_FX_RESID = 2, HB_FX_ONES = 3 };
int hb_fx_create(hb_engine* e, const hb_fx_desc* d, hb_fx** out);
void hb_fx_destroy(hb_fx* f);
/* x'yadj over this rank's rows, x = covariate idx / J / yadj itself; x'x */
int hb_fx_dot(hb_fx* f, int kind, int idx, double* out);
int hb_fx_self_dot(hb_fx* f, int kind, int idx, double* out);
/* yadj += a_r x, u += a_u x  (HB_FX_ONES: x = 1, the intercept shift of Bayes.cpp:482) */
int hb_fx_axpy(hb_fx* f, int kind, int idx, double a_r, double a_u);
/* Z'yadj of random term `term` (nlev[term] sums, this rank's rows); yadj[k] += diff[level of row k] */
int hb_fx_level_sums(hb_fx* f, int term, double* sums);
int hb_fx_level_apply(hb_fx* f, int term, const double* diff);
/* epsilon: records per entry (Z'Z, all ranks); RHS part Z'yadj.tail(ne) (host copy optional; set_rhs after an all-reduce);
 * one Gauss-Seidel sampling pass + the update of yadj/u + eps' Gi eps (Bayes.cpp:565-577); record sums */
int hb_fx_eps_set_counts(hb_fx* f, const double* cnt);
int hb_fx_eps_rhs(hb_fx* f, double* rhs_host);
int hb_fx_eps_set_rhs(hb_fx* f, const double* rhs_host);
int hb_fx_eps_sample(hb_fx* f, int iter, double vare, double ratio, double* quad);
int hb_fx_eps_accumulate(hb_fx* f);
int hb_fx_eps_get(hb_fx* f, double* est, double* sum);
int hb_fx_describe(hb_fx* f, int* eps_levels);
/* BSLMM block Gibbs step, Bayes.cpp:519-544: k_new = K ((c1 % K'(yadj + k_old)) + c2 % z) with the host's c1 = eval / ve,
 * c2 = sqrt(max(eval, 0)) (nk each) and z the position-addressed normals; yadj += k_old - k_new, u -= k_old - k_new;
 * *quad = (K'k_new)' diag(1 / Kival) (K'k_new).  accumulate: the running sum of the records (:858).  ghat_vec: the n-vector
 * K ((K' k_mean) / Kival / sumvx) whose X' product is added to the stored effects (:956-962). */
int hb_fx_k_step(hb_fx* f, int iter, const double* c1, const double* c2, const double* kival, double* quad);
int hb_fx_k_accumulate(hb_fx* f);
int hb_fx_k_ghat_vec(hb_fx* f, const double* kival, double sumvx, double count, double* v_host);
/* out[j] = sum_i X[i][j] v[i] over the resident genotypes (X.t() * v, Bayes.cpp:961) */
int hb_engine_xt_vec(hb_engine* e, const double* v, double* out);

/* ------------------------------------------------------------------ host driver layer */
#define HB_NA (__builtin_nan(""))

/* x_type == 2: the genotypes come from a PLINK .bed image (hb_engine_load_bed) */
typedef struct {
  const uint8_t* file; size_t len;  /* whole file image */
  int nid;                          /* individuals in the file (.fam lines) */
  const int32_t* rows;              /* NULL or n 0-based file individuals, in the order of y */
  int impt, dominance;              /* read_bed()'s impt and d */
} hb_bed_source;

/* Argument list of Rcpp::List Bayes(...) (/root/reference/src/Bayes.cpp:60-88): every one of its 27 arguments has a field
 * here (`threads` is ignored, `seed` replaces R's RNG state).  Model "BSLMM" without Ki is the BayesCpi sweep, as in the
 * reference (Bayes.cpp:97-106). */
typedef struct {
  int n, m;
  const double* y;          /* n                                   (arma::vec& y) */
  const void* X;            /* n x m column-major                  (arma::mat& X) */
  int x_type;               /* 0: double, 1: int8, 2: X points to an hb_bed_source, a .bed image decoded on the device */
  const char* model;        /* std::string model */
  int n_fold;
  const double* Pi;         /* arma::vec Pi */
  const double* fold;       /* Nullable<arma::vec> fold, NULL = R_NilValue */
  int nc; const double* C;  /* Nullable<arma::mat> C, n x nc column-major */
  int nr; const int32_t* Rlev; const int32_t* nlev; /* Nullable<CharacterMatrix> R as 0-based level codes */
  int niter, nburn, thin;
  double dfvr, s2vr, vg, dfvg, s2vg, ve, dfve, s2ve; /* Nullable<double>: NaN = R_NilValue */
  const int32_t* windindx;  /* Nullable<arma::uvec>, 1-based, m entries */
  int outfreq; int verbose;
  uint64_t seed;            /* derived by the Rcpp shim from R's RNG state */
  /* single-step term: epsl_y_J, epsl_Gi (CSC), epsl_index (1-based) */
  int ne, qe;
  const double* epsl_y_J; const int32_t* epsl_index;
  const int32_t* Gi_colptr; const int32_t* Gi_rowidx; const double* Gi_val;
  /* engine knobs (0 = defaults) */
  int device, tile_snps, lag_tiles, n_slabs;
  /* Row sharding (world > 1): y, X hold this rank's n individuals; n_total is the sum over ranks.  The caller
   * supplies the collectives (torch.distributed / NCCL / MPI): all three return 0 on success and are called by
   * every rank in the same order.  world <= 1: leave everything 0 / NULL. */
  int rank, world;
  long long n_total;
  void* comm_ctx;
  int (*allreduce_sum_f64)(void* ctx, double* host_buf, size_t count);        /* in place, host memory */
  int (*allreduce_sum_i32_dev)(void* ctx, void* device_buf, size_t count);    /* in place, device memory */
  int (*allgather_bytes)(void* ctx, const void* mine, void* all, size_t bytes_per_rank); /* rank-ordered */
  /* BSLMM polygenic term (Nullable<arma::vec> Kival, Nullable<arma::mat> Ki; Bayes.cpp:64-65, 203-233, 518-552, 955-964):
   * Ki = eigenvectors of the relationship matrix, n x nk column-major with nk = n, Kival = its nk eigenvalues; nk = 0 /
   * NULL = R_NilValue.  Runs on the device (csrc/effects.cu); not available with world > 1. */
  int nk; const double* Kival; const double* Ki;
} hb_bayes_args;

typedef struct {
  double Vg, Ve, h2, mu, Veps, J;
  double* beta; double* alpha; double* pi; double* pip; double* gwas;
  double* g; double* e; double* vr; double* estR; double* epsilon;
  double* mu_store; double* vara_store; double* vare_store; double* hsq_store;
  double* pi_store; double* alpha_store; double* beta_store;
  int32_t* tracker_final; double* nzrate_count; double* wppa_count;
  int32_t* nnz_trace; double* vara_trace; double* vare_trace; double* varg_trace;
  int n_records_done, nzct, iters_done;
  double seconds_sweep;     /* device time inside hb_engine_sweep, seconds */
  double seconds_setup;     /* load + stats + gram */
  /* diagnostics of the scalar chain: speculation rounds and tiles summed over the sweeps (rounds > tiles means that
   * classes had to be re-decided after a first look), and the sweeps' device time per kernel */
  long long rounds_total, tiles_total;
  /* optional per-iteration traces (niter entries each, NULL = not wanted): speculation rounds of the sweep and the
   * device time of the iteration's kernels in milliseconds */
  int32_t* rounds_trace; float* sweep_ms_trace;
  /* MCMCsamples of the other terms (Bayes.cpp:867-876, 987-1020; optional, NULL = not wanted): Vr nr x records, r
   * (all terms' levels) x records, Veps and J one per record, epsilon qe x records -- column-major like the reference's */
  double* vr_store; double* estR_store; double* veps_store; double* J_store; double* epsilon_store;
} hb_bayes_out;

int hb_bayes(const hb_bayes_args* a, hb_bayes_out* o);

/* Test hooks (host code, no GPU needed): the class-decision code of the sweep compiled for the host.
 * a[k-1], c[k-1] for k = 1..n_fold-1: class score s_k = a_k + c_k * rhs^2 (Bayes.cpp:759-763); TL/TH: n_fold-1 each. */
int hb_test_class_thresholds(int n_fold, double u, const double* a, const double* c, double logpi0, double* TL, double* TH);
int hb_test_class_of(int n_fold, double rr, double u, const double* a, const double* c, double logpi0, const double* TL,
                     const double* TH, int* by_threshold, int* exact);
/* the same two decisions taken on the device for `count` (rr, u) pairs (needs a GPU): int8 classes, -1 = inside a bracket */
int hb_test_class_batch_device(int device, int n_fold, long long count, const double* rr, const double* u, const double* a,
                               const double* c, double logpi0, int8_t* by_threshold, int8_t* exact);

/* ------------------------------------------------------------------ SBayesD (dense LD, summary statistics) */
/* Device engine for the LD-column sweep of SBayesD(), /root/reference/src/SBayesD.cpp:253-456: r_hat lives on the
 * device, one hb_ld_engine_sweep() per MCMC iteration replaces the switch(model_index) block and returns the sums
 * the host needs for the variance draws (:460-467). */
typedef struct hb_ld_engine hb_ld_engine;
int hb_ld_engine_create(int device, int m, uint64_t seed, hb_ld_engine** out);
void hb_ld_engine_destroy(hb_ld_engine* e);
int hb_ld_engine_load_dense(hb_ld_engine* e, const double* ldm);   /* m x m column-major (arma::mat ldm, :7) */
/* arma::sp_mat ldm of SBayesS() (/root/reference/src/SBayesS.cpp:21-40; walked by column at :292-296, :403-407) as compressed
 * sparse columns: colptr m + 1, rowidx / val nnz, row indices strictly ascending inside a column.  The device keeps the
 * CSC arrays (12 bytes per stored entry), not a dense copy. */
int hb_ld_engine_load_csc(hb_ld_engine* e, const int32_t* colptr, const int32_t* rowidx, const double* val);
int hb_ld_engine_describe(hb_ld_engine* e, int* csc, uint64_t* ld_bytes, int* grid);
/* xpx_j = n * LD_jj (:93-96), ifest (:100-103), xy (:104), initial r_hat (:105) */
int hb_ld_engine_set_state(hb_ld_engine* e, const double* xpx, const uint8_t* ifest, const double* xy, const double* r_hat);
int hb_ld_engine_set_vargL(hb_ld_engine* e, const double* vargL);
int hb_ld_engine_set_sparse_info(hb_ld_engine* e, const double* varediff, const double* vx);
int hb_ld_engine_get(hb_ld_engine* e, double* g, int32_t* tracker, double* r_hat);   /* any of them may be NULL */

typedef struct {
  int iter, model_index, n_fold;
  double fold[HB_MAX_FOLD], logpi[HB_MAX_FOLD], vara_fold[HB_MAX_FOLD];   /* as hb_sweep_in */
  double vare, dfvara, s2varg, lambda, lambda2;
  double nscale;               /* n = int(mean(N)), :34 */
  /* SBayesS (SBayesS.cpp:131-141, 285, 388-398): per-SNP residual variance varediff_j * vara + vare and the re-draw
   * of effects with g^2 vx_j > vary; needs hb_ld_engine_set_sparse_info() */
  int sparse_mode;
  double vara, vary;
} hb_ld_sweep_in;
typedef struct {
  double count[HB_MAX_FOLD];   /* estimated SNPs per class */
  double varg_acc;             /* model 1: g'g (:269); 4: sum g^2 (:345); 6: sum g^2/fold (:432) */
  double sum_vargL;            /* model 5 (:386) */
  double g_xy_minus_rhat;      /* g'(xy - r_hat), :460 */
  double g_xy_plus_rhat;       /* g'(xy + r_hat), :466 */
  int n_changed, status, rounds, reserved;
  float sweep_ms;              /* device time of the sweep kernel */
  unsigned long long ld_entries;  /* LD entries the column updates streamed (dense: rows x changed columns; CSC: stored entries) */
} hb_ld_sweep_out;
int hb_ld_engine_sweep(hb_ld_engine* e, const hb_ld_sweep_in* in, hb_ld_sweep_out* out);

/* Host driver: drop-in for the body of  Rcpp::List SBayesD(...)  (SBayesD.cpp:5-24 signature, :532-578 return list). */
typedef struct {
  int m;
  const double* sumstat;    /* m x 4 column-major: MAF, BETA, SE, N (R/sbayes.r:209); NaN = NA */
  const double* ldm;        /* m x m column-major */
  const char* model;
  int n_fold;
  const double* Pi;
  const double* fold;       /* NULL = R_NilValue */
  int niter, nburn, thin;
  double vg, dfvg, s2vg, ve, dfve, s2ve;   /* NaN = R_NilValue */
  const int32_t* windindx;  /* m, 1-based, or NULL */
  int outfreq, verbose;
  uint64_t seed;
  int device;
  /* hb_sbayess(): the LD matrix as arma::sp_mat / dgCMatrix (CSC, m columns); ldm = NULL */
  const int32_t* ld_colptr; const int32_t* ld_rowidx; const double* ld_val;
} hb_sbayes_args;
typedef struct {
  double Vg, Ve, h2;
  double* alpha; double* pi; double* pip; double* gwas;
  double* vara_store; double* vare_store; double* hsq_store; double* pi_store; double* alpha_store;
  int32_t* tracker_final; double* nzrate_count; double* wppa_count;
  int32_t* nnz_trace; double* vara_trace; double* vare_trace; double* varg_trace;
  double* r_hat_final;
  int n_records_done, nzct, iters_done, n_used;
  double seconds_sweep;
  /* diagnostics of the sweeps (bench): changed SNPs (= LD columns walked), LD entries streamed by their updates, bytes of
   * LD held on the device, speculation rounds and tiles */
  long long columns_total, ld_entries_total, ld_bytes_device, rounds_total, tiles_total;
} hb_sbayes_out;
int hb_sbayesd(const hb_sbayes_args* a, hb_sbayes_out* o);
/* drop-in for the body of  Rcpp::List SBayesS(...)  (/root/reference/src/SBayesS.cpp:21-40): sparse LD matrix.  The
 * device copy of the LD matrix is dense in this build (the column updates then add zeros where the sparse matrix
 * has no entry: same results); m is therefore limited by m*m*8 bytes of HBM. */
int hb_sbayess(const hb_sbayes_args* a, hb_sbayes_out* o);

/* ------------------------------------------------------------------ LD builder and .bed reader
 * hb_ldmat_*: tXXmat_Geno() / tXXmat_Chr() of /root/reference/src/tXXmat.cpp:100-185, 504-605 (R: ldmat(),
 * R/ldm.r:31-112) on the device.  The genotype matrix is loaded once (int8, one row per SNP); the m x m matrix
 *   LD_ij = (x_i'x_j - sum_i mean_j - sum_j mean_i + n mean_i mean_j) / n          (tXXmat.cpp:141,148)
 * comes from an exact int8 tensor-core Gram with the centring fused, in the reference's operation order.
 *   chr        NULL -> tXXmat_Geno; m chromosome codes -> tXXmat_Chr (pairs across chromosomes are 0 / not stored)
 *   has_chisq  0 -> the reference's dense branch (diagonal xx^2/n, :157,:584);
 *              1 -> its sparse branch: entries with r^2 n <= chisq are dropped (:142-145), diagonal included.
 * The tXXmat_*_gwas variants (merging a second genotype set) are not built. */
typedef struct hb_ldmat hb_ldmat;
int hb_ldmat_create(int device, int n, int m, hb_ldmat** out);
void hb_ldmat_destroy(hb_ldmat* h);
int hb_ldmat_load_i8(hb_ldmat* h, const int8_t* X, size_t ld);           /* n x m column-major (the char big.matrix) */
int hb_ldmat_load_bed(hb_ldmat* h, const uint8_t* file, size_t len, int nid, const int32_t* rows, int impt, int dominance);
/* BigStat(), tXXmat.cpp:43-77: mean, sum, xx = sqrt(sum (x - mean)^2); any pointer may be NULL */
int hb_ldmat_stats(hb_ldmat* h, double* mean, double* sum, double* xx);
/* full m x m matrix, column-major with leading dimension ldo (zeros where the reference stores nothing) */
int hb_ldmat_dense(hb_ldmat* h, const int32_t* chr, int has_chisq, double chisq, double* out, size_t ldo);
/* the arma::sp_mat results (wrap() -> dgCMatrix): compute, then fetch colptr[m+1], rowidx[nnz] (ascending per
 * column), val[nnz].  Stored entries = assigned values != 0, as arma::sp_mat keeps them. */
int hb_ldmat_sparse(hb_ldmat* h, const int32_t* chr, int has_chisq, double chisq, long long* nnz);
int hb_ldmat_sparse_get(hb_ldmat* h, long long* colptr, int32_t* rowidx, double* val);
int hb_ldmat_set_panel_cols(hb_ldmat* h, int cols);   /* columns per device panel (multiple of 64; default <= 1 GiB) */
int hb_ldmat_last_ms(hb_ldmat* h, float* gram_ms);    /* device time of the Gram/epilogue kernel in the last call */

/* read_bed<char>() of /root/reference/src/read_bed.cpp:97-232 into the caller's nid x m column-major int8 matrix (the
 * "char" big.matrix of R/read_plink.r:51-59): code map 00->2 (0 with dominance), 10->1, 11->0, 01->NA (-128), missing
 * replaced by the SNP's major genotype when impt; miss (m flags, may be NULL) = the reader's miss[]. */
int hb_bed_decode(int device, const uint8_t* file, size_t len, int nid, int m, int impt, int dominance, int8_t* out,
                  uint8_t* miss);
/* host build of the decoder's byte-level code for one SNP (CPU tests; no device needed) */
/* Window cutters behind `windindx` (/root/reference/src/cutwind.cpp:13-65; R/sbayes.r:176-182): chr and pos per SNP (numeric
 * chromosome codes), 1-based window ids out; 0 = in no window (positions < 1, which the reference leaves undefined). */
int hb_cutwind_by_bp(const double* chr, const double* pos, int m, double bp, int32_t* windindx);
int hb_cutwind_by_num(const double* chr, const double* pos, int m, int fixN, int32_t* windindx);

/* host build of the LD builder's epilogue: out (m x m) from an exact int32 Gram matrix and BigStat's vectors */
int hb_test_ld_entries(int n, int m, const int32_t* gram, const double* sum, const double* mean, const double* xx,
                       const int32_t* chr, int has_chisq, double chisq, double* out);
/* host build of the LD builder's BigStat code: Xc = m rows of Kpad bytes, 16-byte aligned, zeros beyond n */
int hb_test_ld_stats(const int8_t* Xc, int Kpad, int n, int m, double* sum, double* mean, double* xx);
/* host build of csrc/hb_limbs.h (fixed-point limb dot, a building block not yet used by the sweep): x'q, q = rint(r scale) */
int hb_test_limb_dot(const uint8_t* x, const double* r, int n, double scale, long long* dot_q, int* ok);
int hb_test_bed_decode_snp(const uint8_t* snp_bytes, int nid, const int32_t* rows, int n, int impt, int dominance,
                           int8_t* out, uint8_t* info_out);

#ifdef __cplusplus
}
#endif
#endif /* HIBAYES_B200_H */
