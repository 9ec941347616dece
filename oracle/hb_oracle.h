/*
 * hb_oracle.h -- CPU oracle for the hibayes Gibbs hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
 * arm may load this library; the product (hibayes_b200/) never links or calls it.
 *
 * PARITY PINNED AGAINST THE REFERENCE ITSELF (YinLiLin/hibayes @ 98328fe): the reference ships no tests and no golden
 * vectors, and R / Rcpp / RcppArmadillo / bigmemory are absent from this image -- but its C++ sources compile
 * UNMODIFIED, from where they lie, against the stand-in headers of oracle/ref_shim/ (a small eager Armadillo, the Rcpp
 * types the files touch, reference-BLAS ddot_/daxpy_) into oracle/_ref/libhibayes_ref.so (oracle/Makefile).  The one
 * substitution is the random stream: libR's sequential generator is replaced by a REPLAY of the variates this oracle
 * consumed (hbo_tape_*), every draw checked for kind and (gamma, chi-square) for a bit-identical shape.  On the same
 * inputs and variates the compiled Bayes(), SBayesD(), SBayesS() return the same bits as this oracle for every recorded
 * effect, variance and pi of 7 of the 8 models (BayesL: 1e-12 ... 4e-10, its inverse-Gaussian root is evaluated without
 * the reference's cancellation), and BigStat / tXXmat_Geno / tXXmat_Chr / read_bed<char> the same bits throughout
 * (tests/test_reference_pin.py; golden vectors of the compiled reference in tests/golden/ref_*.npz).  What stays a
 * restatement: Armadillo's own reductions (sum / mean / var / dot: its published two-accumulator algorithms, restated in
 * ref_shim/mini_arma.h as they are in this file) and the sampler algorithms behind the variates (hb_rng.h: Philox,
 * AS241, Marsaglia-Tsang, Michael-Schucany-Haas; pinned to R's documented distributions by tests/test_samplers.py).
 */
#ifndef HB_ORACLE_H
#define HB_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HBO_NA (__builtin_nan(""))

typedef struct {
  /* data */
  int n, m;
  const double* y;       /* n */
  const void* X;         /* n x m column-major */
  int x_is_int8;         /* 0: double, 1: int8 */
  const char* model;     /* "BayesRR","BayesA","BayesB","BayesBpi","BayesC","BayesCpi","BayesL","BayesR" */
  int n_fold;
  const double* Pi;      /* n_fold */
  const double* fold;    /* n_fold or NULL */
  int nc;
  const double* C;       /* n x nc column-major or NULL */
  int nr;
  const int32_t* Rlev;   /* n x nr column-major 0-based level codes or NULL */
  const int32_t* nlev;   /* nr */
  int niter, nburn, thin;
  double dfvr, s2vr, vg, dfvg, s2vg, ve, dfve, s2ve; /* NaN = not given */
  const int32_t* windindx; /* m, 1-based window ids, or NULL */
  uint64_t seed;
  /* single-step epsilon term (Bayes.cpp:254-275,554-584); ne = 0 disables */
  int ne, qe;
  const double* epsl_y_J;     /* n */
  const int32_t* epsl_index;  /* ne, 1-based */
  const int32_t* Gi_colptr;   /* qe+1, CSC of epsl_Gi */
  const int32_t* Gi_rowidx;
  const double* Gi_val;
  /* BSLMM polygenic term (Bayes.cpp:203-233, 518-552, 955-964); nk = 0 disables.  Ki: n x nk column-major eigenvectors
   * of the relationship matrix (nk must equal n: Bayes.cpp:519 adds an nk-vector to yadj), Kival: nk eigenvalues */
  int nk;
  const double* Kival;
  const double* Ki;
} hbo_bayes_args;

typedef struct {
  double Vg, Ve, h2, mu, Veps, J;
  double* beta;        /* nc */
  double* alpha;       /* m */
  double* pi;          /* n_fold */
  double* pip;         /* m */
  double* gwas;        /* nw (max window id) or NULL */
  double* g;           /* n   (the reference returns u here, Bayes.cpp:1023) */
  double* e;           /* n */
  double* vr;          /* nr */
  double* estR;        /* sum(nlev) */
  double* epsilon;     /* qe */
  /* MCMC stores, n_records columns each (NULL to skip) */
  double* mu_store; double* vara_store; double* vare_store; double* hsq_store;
  double* pi_store;    /* n_fold x n_records col-major */
  double* alpha_store; /* m x n_records col-major */
  double* beta_store;  /* nc x n_records */
  /* diagnostics for parity tests */
  int32_t* tracker_final; /* m */
  double* nzrate_count;   /* m: raw counts before /nzct */
  double* wppa_count;     /* nw */
  int32_t* nnz_trace;     /* niter */
  double* vara_trace;     /* niter */
  double* vare_trace;     /* niter */
  double* varg_trace;     /* niter */
  int n_records_done;
  int nzct;
  int iters_done;
  double seconds_sweep;   /* wall seconds inside the SNP sweeps */
  /* MCMCsamples of the other terms, Bayes.cpp:867-876 (NULL to skip): Vr nr x records, r levels x records, Veps, J per
   * record, epsilon qe x records */
  double* vr_store; double* estR_store; double* veps_store; double* J_store; double* epsilon_store;
} hbo_bayes_out;

/* returns 0 on success, else nonzero and hbo_last_error() holds the message
 * (same texts as the Rcpp::exception()s in Bayes.cpp:92-117,325,356). */
int hbo_bayes(const hbo_bayes_args* a, hbo_bayes_out* o);
const char* hbo_last_error(void);

/* ---- tape of the random variates a run consumes, in the order of the REFERENCE's own sampler calls (stats.cpp) ----
 * Recorded by hbo_bayes / hbo_sbayesd / hbo_sbayess between hbo_tape_begin() and hbo_tape_end(); replayed to the compiled
 * reference (oracle/_ref/libhibayes_ref.so, oracle/ref_shim/) in place of libR's generator, so that both see the same
 * numbers.  kind: 0 unif_rand(), 1 norm_rand() (a standard normal), 2 R::rgamma(shape = param, 1), 3 R::rchisq(df = param). */
typedef struct { int32_t kind; int32_t pad; double value; double param; } hbo_tape_entry;
void hbo_tape_begin(hbo_tape_entry* buf, uint64_t cap);
uint64_t hbo_tape_end(void);   /* entries the run produced (may exceed cap: then the tape is incomplete) */

/* the reference's literal class draw (Bayes.cpp:757-781) for `count` (rhs, uniform) pairs of one SNP */
void hbo_class_literal_batch(int n_fold, long long count, const double* rhs, const double* rval, double xx, double vare_,
                             const double* vara_fold, const double* logpi, int8_t* out);

/* helpers exposed for unit tests */
double hbo_var(const double* x, int n);             /* Armadillo var(), norm_type 0 */
double hbo_qnorm(double p);
void hbo_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
double hbo_draw_gamma(uint64_t seed, uint32_t dom, uint32_t iter, uint32_t idx, uint32_t slot, double shape);
double hbo_draw_chisq(uint64_t seed, uint32_t dom, uint32_t iter, uint32_t idx, uint32_t slot, double df);
void hbo_draw_uz(uint64_t seed, uint32_t dom, uint32_t iter, uint32_t idx, uint32_t slot, uint32_t attempt, double* u, double* z);
double hbo_invgauss(double mu, double lambda, double u, double z);
/* the reference's literal (cancelling) expression for the smaller root, stats.cpp:57-59 */
double hbo_invgauss_literal_root(double mu, double lambda, double z);

/* CPU timing of the reference's level-1 path: per-SNP ddot + 2 daxpy on a
 * column-major fp64 X (Bayes.cpp:756,787-789), `threads` OpenMP threads split
 * each vector op.  Returns SNP-updates per second. */
double hbo_time_sweep_fp64(int n, int m_cpu, int sweeps, int threads, uint64_t seed, double* checksum);


/* ---- SBayesD: dense-LD summary-statistics Gibbs sampler (/root/reference/src/SBayesD.cpp:5-609) ---- */
typedef struct {
  int m;
  const double* sumstat;  /* m x 4 column-major: MAF, BETA, SE, N (R/sbayes.r:209); NaN = NA */
  const double* ldm;      /* m x m column-major (SBayesD) or NULL */
  const char* model;
  int n_fold;
  const double* Pi;
  const double* fold;     /* n_fold or NULL */
  int niter, nburn, thin;
  double vg, dfvg, s2vg, ve, dfve, s2ve;   /* NaN = not given */
  const int32_t* windindx;                 /* m, 1-based, or NULL */
  uint64_t seed;
  /* SBayesS: the LD matrix as arma::sp_mat / dgCMatrix (CSC, row indices ascending); ldm = NULL */
  const int32_t* ld_colptr; const int32_t* ld_rowidx; const double* ld_val;
} hbo_sbayes_args;

typedef struct {
  double Vg, Ve, h2;
  double* alpha;       /* m */
  double* pi;          /* n_fold */
  double* pip;         /* m */
  double* gwas;        /* nw or NULL */
  double* vara_store; double* vare_store; double* hsq_store; double* pi_store; double* alpha_store; /* NULL to skip */
  int32_t* tracker_final; double* nzrate_count; double* wppa_count;
  int32_t* nnz_trace; double* vara_trace; double* vare_trace; double* varg_trace;   /* niter each */
  double* r_hat_final;  /* m */
  int n_records_done, nzct, iters_done, n_used;   /* n_used = int(mean(N)) */
} hbo_sbayes_out;

int hbo_sbayesd(const hbo_sbayes_args* a, hbo_sbayes_out* o);
/* SBayesS: sparse-LD variant (/root/reference/src/SBayesS.cpp:21-679) */
int hbo_sbayess(const hbo_sbayes_args* a, hbo_sbayes_out* o);

#ifdef __cplusplus
}
#endif
#endif
