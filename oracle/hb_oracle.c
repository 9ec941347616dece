/*
 * hb_oracle.c -- literal CPU restatement of hibayes' individual-level Gibbs
 * sampler Bayes() (/root/reference/src/Bayes.cpp:60-1094) with its samplers
 * (/root/reference/src/stats.cpp:3-28,55-76) and the sparse Gauss-Seidel
 * sampler used by the single-step model (/root/reference/src/solver.cpp:131-140).
 *
 * TEST INFRASTRUCTURE ONLY -- see hb_oracle.h (parity pinned against the compiled reference, oracle/_ref).
 *
 * Loop structure, operation order and quirks follow the reference line by line;
 * the only substitution is the random stream (hb_rng.h addresses instead of
 * libR's sequential generator).  Third-party arithmetic restated here:
 *   - BLAS level-1 ddot_/daxpy_ (hibayes.h:21-29): unit-stride sequential loops;
 *   - Armadillo var()/mean()/sum() (op_var::direct_var, arrayops::accumulate:
 *     two interleaved accumulators), version unpinned (DESCRIPTION:36).
 * Compile with -ffp-contract=off so no fused multiply-adds are introduced.
 */
#include "hb_oracle.h"
#include "../hibayes_b200/csrc/hb_rng.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static char g_err[512];
const char* hbo_last_error(void) { return g_err; }
static int fail(const char* msg) {
  snprintf(g_err, sizeof g_err, "%s", msg);
  return 1;
}

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* ---- third-party arithmetic, restated -------------------------------------- */
/* arrayops::accumulate */
static double acc_sum(const double* x, int n) {
  double a1 = 0.0, a2 = 0.0;
  int i, j;
  for (i = 0, j = 1; j < n; i += 2, j += 2) { a1 += x[i]; a2 += x[j]; }
  if (i < n) a1 += x[i];
  return a1 + a2;
}
static double acc_mean(const double* x, int n) { return acc_sum(x, n) / (double)n; }

/* op_var::direct_var, norm_type 0 */
double hbo_var(const double* x, int n) {
  if (n < 2) return 0.0;
  double mean = acc_mean(x, n);
  double acc2 = 0.0, acc3 = 0.0;
  int i, j;
  for (i = 0, j = 1; j < n; i += 2, j += 2) {
    double ti = mean - x[i], tj = mean - x[j];
    acc2 += ti * ti + tj * tj;
    acc3 += ti + tj;
  }
  if (i < n) { double ti = mean - x[i]; acc2 += ti * ti; acc3 += ti; }
  return (acc2 - acc3 * acc3 / (double)n) / (double)(n - 1);
}

static double ddot(int n, const double* x, const double* y) {
  double s = 0.0;
  for (int i = 0; i < n; ++i) s += x[i] * y[i];
  return s;
}
static void daxpy(int n, double a, const double* x, double* y) {
  for (int i = 0; i < n; ++i) y[i] += a * x[i];
}

/* ---- sampler wrappers (stats.cpp) ------------------------------------------ */
static hb_key_t KEY;
/* tape of consumed variates (hb_oracle.h): every TAPE() sits where the reference makes the sampler call */
static hbo_tape_entry* g_tape;
static uint64_t g_tape_cap, g_tape_n;
void hbo_tape_begin(hbo_tape_entry* buf, uint64_t cap) { g_tape = buf; g_tape_cap = cap; g_tape_n = 0; }
uint64_t hbo_tape_end(void) { g_tape = 0; return g_tape_n; }
static inline void TAPE(int kind, double value, double param) {
  if (!g_tape) return;
  if (g_tape_n < g_tape_cap) { g_tape[g_tape_n].kind = kind; g_tape[g_tape_n].pad = 0; g_tape[g_tape_n].value = value; g_tape[g_tape_n].param = param; }
  g_tape_n++;
}
enum { TP_U = 0, TP_Z = 1, TP_GAMMA = 2, TP_CHISQ = 3 };
static double z_at(uint32_t dom, uint32_t iter, uint32_t idx, uint32_t slot, uint32_t attempt) {
  const double z = hb_draw_z(KEY, dom, iter, idx, slot, attempt);
  TAPE(TP_Z, z, 0.0);
  return z;
}
static double norm_at(uint32_t dom, uint32_t iter, uint32_t idx, uint32_t slot, double mean, double sd) {
  return mean + sd * z_at(dom, iter, idx, slot, 0); /* stats.cpp:8-11 */
}
static double chisq_at(uint32_t dom, uint32_t iter, uint32_t idx, uint32_t slot, double df) {
  const double c = hb_draw_chisq(KEY, dom, iter, idx, slot, df); /* stats.cpp:22-24 */
  TAPE(TP_CHISQ, c, df);
  return c;
}
static double gamma_at(uint32_t dom, uint32_t iter, uint32_t idx, uint32_t slot, double shape) {
  const double v = hb_draw_gamma(KEY, dom, iter, idx, slot, shape); /* stats.cpp:13-15, scale applied by the caller */
  TAPE(TP_GAMMA, v, shape);
  return v;
}

/* exposed helpers */
double hbo_qnorm(double p) { return hb_qnorm(p); }
void hbo_philox(const uint32_t c[4], const uint32_t k[2], uint32_t out[4]) {
  hb_philox4x32_10(c[0], c[1], c[2], c[3], k[0], k[1], out);
}
double hbo_draw_gamma(uint64_t seed, uint32_t dom, uint32_t iter, uint32_t idx, uint32_t slot, double shape) {
  return hb_draw_gamma(hb_make_key(seed), dom, iter, idx, slot, shape);
}
double hbo_draw_chisq(uint64_t seed, uint32_t dom, uint32_t iter, uint32_t idx, uint32_t slot, double df) {
  return hb_draw_chisq(hb_make_key(seed), dom, iter, idx, slot, df);
}
void hbo_draw_uz(uint64_t seed, uint32_t dom, uint32_t iter, uint32_t idx, uint32_t slot, uint32_t attempt, double* u, double* z) {
  hb_draw_uz(hb_make_key(seed), dom, iter, idx, slot, attempt, u, z);
}
double hbo_invgauss(double mu, double lambda, double u, double z) { return hb_invgauss_from_uz(mu, lambda, u, z); }
double hbo_invgauss_literal_root(double mu, double lambda, double z) { return hb_invgauss_literal_root(mu, lambda, z); }

/* The class draw of BayesR exactly as the reference writes it (Bayes.cpp:757-781) for `count` right-hand sides of one SNP:
 * rhs already holds x'r (+ xx * oldgi), rval the uniform.  n_fold = 2 gives the BayesB/C form of :641-645 / :685-689
 * (acceptProb = 1 / sum(exp(s - s[0])), class 1 unless rval < acceptProb) -- the same numbers in the same order. */
void hbo_class_literal_batch(int n_fold, long long count, const double* rhs_v, const double* rval_v, double xx, double vare_,
                             const double* vara_fold, const double* logpi, int8_t* out) {
#pragma omp parallel for schedule(static)
  for (long long q = 0; q < count; ++q) {
    double s[8], stemp[8];
    const double rhs = rhs_v[q], rval = rval_v[q];
    const double lhs = xx / vare_;
    s[0] = logpi[0];
    for (int j = 1; j < n_fold; ++j) {
      const double vare_vara = vare_ / vara_fold[j];
      const double logdetV = log(vara_fold[j] * lhs + 1);
      const double uhat = rhs / (xx + vare_vara);
      s[j] = -0.5 * (logdetV - (rhs * uhat / vare_)) + logpi[j];
    }
    for (int j = 0; j < n_fold; ++j) {
      double temp = 0.0;
      for (int k = 0; k < n_fold; ++k) temp += exp(s[k] - s[j]);
      stemp[j] = 1 / temp;
    }
    double acceptProb = 0;
    int indistflag = 0;
    for (int j = 0; j < n_fold; ++j) {
      acceptProb += stemp[j];
      if (rval < acceptProb) { indistflag = j; break; }
    }
    out[q] = (int8_t)indistflag;
  }
}

/* column accessor: the reference holds X as arma::mat (fp64); an int8 source is
 * widened into a scratch column so the arithmetic is identical. */
static const double* xcol(const hbo_bayes_args* a, int j, double* scratch) {
  if (!a->x_is_int8) return (const double*)a->X + (size_t)j * a->n;
  const int8_t* p = (const int8_t*)a->X + (size_t)j * a->n;
  for (int i = 0; i < a->n; ++i) scratch[i] = (double)p[i];
  return scratch;
}

static int isna(double v) { return v != v; }

#define FREE_ALL() do { \
  free(cpc); free(beta); free(vrtmp); free(vrv); free(estR); free(estR_tmp); free(r_rhs); free(r_cnt); \
  free(R_off); free(snptracker); free(nzrate); free(g); free(u); free(xpx); free(vx); free(yadj); \
  free(scratch); free(vargL); free(Pi); free(fold_); free(fold_snp_num); free(logpi); free(s); free(stemp); \
  free(vara_fold); free(vare_vara_fold); free(wppai); free(wstart); free(wmembers); free(gsum); free(pisum); \
  free(k_estR); free(k_tmp); free(k_sum); free(k_rhs); free(k_eval); free(k_t); free(k_w); \
  free(betasum); free(estRsum); free(e_estR); free(e_tmp); free(e_rhs); free(e_lhsdiag); free(e_cnt); free(e_sum); \
  free(diff); } while (0)

int hbo_bayes(const hbo_bayes_args* a, hbo_bayes_out* o) {
  g_err[0] = 0;
  const int n = a->n, m = a->m;
  const char* model = a->model;
  KEY = hb_make_key(a->seed);
  double *cpc = 0, *beta = 0, *vrtmp = 0, *vrv = 0, *estR = 0, *estR_tmp = 0, *r_rhs = 0, *r_cnt = 0;
  int* R_off = 0;
  double *snptracker = 0, *nzrate = 0, *g = 0, *u = 0, *xpx = 0, *vx = 0, *yadj = 0, *scratch = 0, *vargL = 0;
  double *Pi = 0, *fold_ = 0, *fold_snp_num = 0, *logpi = 0, *s = 0, *stemp = 0, *vara_fold = 0, *vare_vara_fold = 0;
  double *wppai = 0; int *wstart = 0, *wmembers = 0;
  double *gsum = 0, *pisum = 0, *betasum = 0, *estRsum = 0;
  double *e_estR = 0, *e_tmp = 0, *e_rhs = 0, *e_lhsdiag = 0, *e_cnt = 0, *e_sum = 0, *diff = 0;
  double *k_estR = 0, *k_tmp = 0, *k_sum = 0, *k_rhs = 0, *k_eval = 0, *k_t = 0, *k_w = 0;   /* BSLMM */

  /* Bayes.cpp:92-117 argument checks */
  for (int i = 0; i < n; ++i) if (isna(a->y[i])) return fail("NAs are not allowed in y.");
  int is = !strcmp(model, "BayesRR") ? 1 : !strcmp(model, "BayesA") ? 2 :
           (!strcmp(model, "BayesB") || !strcmp(model, "BayesBpi")) ? 3 :
           (!strcmp(model, "BayesC") || !strcmp(model, "BayesCpi") || !strcmp(model, "BSLMM")) ? 4 :
           !strcmp(model, "BayesL") ? 5 : 6;
  const int model_index = is;
  int fixpi = 0;
  if (!strcmp(model, "BayesB") || !strcmp(model, "BayesC")) fixpi = 1;
  const int n_fold = a->n_fold;
  if (n_fold < 2) return fail("Pi should be a vector.");
  { double sp = 0.0; /* arma::sum -> accumulate */
    sp = acc_sum(a->Pi, n_fold);
    if (sp != 1) return fail("sum of Pi should be 1."); }
  if (a->Pi[0] == 1) return fail("all markers have no effect size.");
  for (int i = 0; i < n_fold; ++i)
    if (a->Pi[i] < 0 || a->Pi[i] > 1) return fail("elements of Pi should be at the range of [0, 1]");
  Pi = (double*)malloc(sizeof(double) * n_fold);
  memcpy(Pi, a->Pi, sizeof(double) * n_fold);
  fold_ = (double*)calloc(n_fold > 2 ? n_fold : 2, sizeof(double));
  if (a->fold) memcpy(fold_, a->fold, sizeof(double) * n_fold);
  else {
    if (!strcmp(model, "BayesR")) { FREE_ALL(); return fail("'fold' should be provided for BayesR model."); }
    if (n_fold != 2) { FREE_ALL(); return fail("length of Pi and fold not equals."); }
  }

  const double vary = hbo_var(a->y, n); /* :121 */
  const double h2 = 0.5;
  const int niter = a->niter, nburn = a->nburn, thin = a->thin;
  const int n_records = (niter - nburn) / thin; /* :124 integer division */

  /* covariates :126-147 */
  const int nc = a->nc;
  if (nc) {
    cpc = (double*)malloc(sizeof(double) * nc);
    beta = (double*)calloc(nc, sizeof(double));
    for (int i = 0; i < nc; ++i) cpc[i] = ddot(n, a->C + (size_t)i * n, a->C + (size_t)i * n);
  }
  /* env. random effects :149-201 */
  const int nr = a->nr;
  double dfr = isna(a->dfvr) ? -1 : a->dfvr;
  double s2r = isna(a->s2vr) ? 0 : a->s2vr;
  int n_levels = 0;
  if (nr) {
    vrtmp = (double*)malloc(sizeof(double) * nr);
    vrv = (double*)calloc(nr, sizeof(double));
    R_off = (int*)malloc(sizeof(int) * (nr + 1));
    R_off[0] = 0;
    for (int i = 0; i < nr; ++i) {
      vrtmp[i] = vary * (1 - h2) / (nr + 1);
      n_levels += a->nlev[i];
      R_off[i + 1] = n_levels;
    }
    estR = (double*)calloc(n_levels, sizeof(double));
    estR_tmp = (double*)calloc(n_levels, sizeof(double));
    r_rhs = (double*)calloc(n_levels, sizeof(double));
    r_cnt = (double*)calloc(n_levels, sizeof(double));
    for (int i = 0; i < nr; ++i)
      for (int k = 0; k < n; ++k) r_cnt[R_off[i] + a->Rlev[(size_t)i * n + k]] += 1.0; /* diag(Z'Z) */
    diff = (double*)malloc(sizeof(double) * n);
  }
  /* single-step epsilon term :235-275 */
  const int ne = a->ne, qe = ne ? a->qe : 0;
  double veps = 0, vepstmp = 0, JtJ = 0, epsl_J_beta = 0;
  if (ne) {
    if (!a->Gi_colptr) { FREE_ALL(); return fail("variance-covariance matrix should be provided for epsilon term."); }
    JtJ = ddot(n, a->epsl_y_J, a->epsl_y_J);
    e_estR = (double*)calloc(qe, sizeof(double));
    e_tmp = (double*)calloc(qe, sizeof(double));
    e_rhs = (double*)calloc(qe, sizeof(double));
    e_lhsdiag = (double*)calloc(qe, sizeof(double));
    e_cnt = (double*)calloc(qe, sizeof(double)); /* diag of epsl_ZZ */
    e_sum = (double*)calloc(qe, sizeof(double));
    for (int i = 0; i < ne; ++i) e_cnt[a->epsl_index[i] - 1] += 1.0;
  }

  /* BSLMM polygenic term :203-233 */
  const int nk = a->nk;
  double vbtmp = 0;
  if (nk) {
    if (!a->Ki || !a->Kival) { FREE_ALL(); return fail("Ki and Kival should be provided together."); }
    if (nk != n) { FREE_ALL(); return fail("variance-covariance matrix should be in square."); }   /* :221, and :519 needs nk == n */
    k_estR = (double*)calloc(nk, sizeof(double));
    k_tmp = (double*)calloc(nk, sizeof(double));
    k_sum = (double*)calloc(nk, sizeof(double));
    k_rhs = (double*)calloc(n, sizeof(double));
    k_eval = (double*)calloc(nk, sizeof(double));
    k_t = (double*)calloc(nk, sizeof(double));
    k_w = (double*)calloc(nk, sizeof(double));
  }

  int count = 0, nzct = 0, NnzSnp = 0, indistflag;
  double xx, oldgi, gi, gi_, rhs, lhs, logdetV, acceptProb, uhat, v;
  double vara_, dfvara_, s2vara_, vare_, dfvare_, s2vare_, vargi, s2varg_;
  int have_tracker = 0;
  if (!strcmp(model, "BayesRR") || !strcmp(model, "BayesA") || !strcmp(model, "BayesL")) { /* :288-292 */
    NnzSnp = m;
    Pi[0] = 0; Pi[1] = 1;
    fixpi = 1;
  } else {
    if (strcmp(model, "BayesR") && n_fold != 2) {
      FREE_ALL();
      return fail("length of Pi should be 2, the first value is the proportion of non-effect markers.");
    }
    have_tracker = 1;
  }
  nzrate = (double*)calloc(m, sizeof(double));
  snptracker = (double*)calloc(m, sizeof(double));
  g = (double*)calloc(m, sizeof(double));
  u = (double*)calloc(n, sizeof(double));
  xpx = (double*)calloc(m, sizeof(double));
  vx = (double*)calloc(m, sizeof(double));
  scratch = (double*)malloc(sizeof(double) * n);
  yadj = (double*)malloc(sizeof(double) * n);

  /* :310-317 column statistics */
  for (int i = 0; i < m; ++i) {
    const double* xi = xcol(a, i, scratch);
    double ss = 0.0; /* sum(square(Xi)) -> accumulate over the squared vector */
    { double a1 = 0, a2 = 0; int p, q;
      for (p = 0, q = 1; q < n; p += 2, q += 2) { a1 += xi[p] * xi[p]; a2 += xi[q] * xi[q]; }
      if (p < n) a1 += xi[p] * xi[p];
      ss = a1 + a2; }
    xpx[i] = ss;
    vx[i] = hbo_var(xi, n);
  }
  double sumvx = acc_sum(vx, m);
  int nvar0 = 0;
  for (int i = 0; i < m; ++i) nvar0 += (vx[i] == 0);

  /* :319-375 priors */
  dfvara_ = isna(a->dfvg) ? 4 : a->dfvg;
  if (dfvara_ <= 2) { FREE_ALL(); return fail("dfvg should not be less than 2."); }
  vara_ = isna(a->vg) ? ((dfvara_ - 2) / dfvara_) * vary * h2 : a->vg;
  vepstmp = vara_;
  vbtmp = vara_;   /* :333 */
  vare_ = isna(a->ve) ? vary * (1 - h2) / (nr + 1) : a->ve;
  dfvare_ = isna(a->dfve) ? -2 : a->dfve;
  s2vara_ = isna(a->s2vg) ? vara_ * (dfvara_ - 2) / dfvara_ : a->s2vg;
  double varg = vara_ / ((1 - Pi[0]) * sumvx);
  s2varg_ = s2vara_ / ((1 - Pi[0]) * sumvx);
  s2vare_ = isna(a->s2ve) ? 0 : a->s2ve;
  if (niter < nburn) { FREE_ALL(); return fail("Number of total iteration ('niter') shold be larger than burn-in ('nburn')."); }
  double R2 = (dfvara_ - 2) / dfvara_;
  double lambda2 = 2 * (1 - R2) / (R2) * sumvx;
  double lambda = sqrt(lambda2);
  double shape, shape0 = 1.1;
  double rate, rate0 = (shape0 - 1) / lambda2;
  if (!strcmp(model, "BayesL")) {
    vargL = (double*)malloc(sizeof(double) * m);
    for (int i = 0; i < m; ++i) vargL[i] = varg;
  }
  stemp = (double*)calloc(n_fold, sizeof(double));
  fold_snp_num = (double*)calloc(n_fold, sizeof(double));
  logpi = (double*)calloc(n_fold, sizeof(double));
  s = (double*)calloc(n_fold, sizeof(double));
  vara_fold = (double*)calloc(n_fold, sizeof(double));
  vare_vara_fold = (double*)calloc(n_fold, sizeof(double));
  for (int j = 0; j < n_fold; ++j) vara_fold[j] = (vara_ / ((1 - Pi[0]) * sumvx)) * fold_[j];

  /* :376-391 windows */
  int nw = 0, WPPA = 0;
  if (a->windindx) {
    WPPA = 1;
    for (int i = 0; i < m; ++i) if (a->windindx[i] > nw) nw = a->windindx[i];
    wppai = (double*)calloc(nw, sizeof(double));
    wstart = (int*)calloc(nw + 1, sizeof(int));
    wmembers = (int*)malloc(sizeof(int) * m);
    for (int i = 0; i < m; ++i) if (a->windindx[i] >= 1) wstart[a->windindx[i]]++;
    for (int w = 0; w < nw; ++w) wstart[w + 1] += wstart[w];
    int* fill = (int*)calloc(nw, sizeof(int));
    for (int i = 0; i < m; ++i) if (a->windindx[i] >= 1) {
      int w = a->windindx[i] - 1;
      wmembers[wstart[w] + fill[w]++] = i;
    }
    free(fill);
  }

  /* accumulators for posterior means (the reference stores every record and
   * averages at the end, :919-985; optional full stores are filled as well) */
  gsum = (double*)calloc(m, sizeof(double));
  pisum = (double*)calloc(n_fold, sizeof(double));
  if (nc) betasum = (double*)calloc(nc, sizeof(double));
  if (nr) estRsum = (double*)calloc(n_levels, sizeof(double));
  double musum = 0, varasum = 0, varesum = 0, hsqsum = 0, vepssum = 0, Jsum = 0;
  double* vrsum = nr ? (double*)calloc(nr, sizeof(double)) : 0;

  /* :469-471 */
  double mu_, mu = acc_mean(a->y, n);
  for (int i = 0; i < n; ++i) yadj[i] = a->y[i] - mu;
  double t_sweep = 0.0;
  int iter;

  for (iter = 0; iter < niter; ++iter) {
    const uint32_t it = (uint32_t)iter;
    /* intercept :480-482 */
    mu_ = -norm_at(HB_DOM_ITER, it, HB_IT_MU, 0, acc_sum(yadj, n) / n, sqrt(vare_ / n));
    mu -= mu_;
    for (int i = 0; i < n; ++i) yadj[i] += mu_ * 1.0;

    /* covariates :484-494 */
    for (int i = 0; i < nc; ++i) {
      const double* dci = a->C + (size_t)i * n;
      oldgi = beta[i];
      v = cpc[i];
      rhs = ddot(n, dci, yadj);
      rhs += v * oldgi;
      gi = norm_at(HB_DOM_COV, it, (uint32_t)i, 0, rhs / v, sqrt(vare_ / v));
      gi_ = oldgi - gi;
      daxpy(n, gi_, dci, yadj);
      beta[i] = gi;
    }

    /* env. random effects :496-516 */
    for (int i = 0; i < nr; ++i) {
      const int off = R_off[i], qr = a->nlev[i];
      const int32_t* lev = a->Rlev + (size_t)i * n;
      for (int q = 0; q < qr; ++q) r_rhs[off + q] = 0.0;
      for (int k = 0; k < n; ++k) r_rhs[off + lev[k]] += yadj[k];      /* Z' yadj */
      for (int q = 0; q < qr; ++q) r_rhs[off + q] += r_cnt[off + q] * estR[off + q]; /* + ZZ estR */
      for (int q = 0; q < qr; ++q) {
        double l = r_cnt[off + q] + vare_ / vrtmp[i];
        estR_tmp[off + q] = norm_at(HB_DOM_RAND, it, (uint32_t)(off + q), 0, r_rhs[off + q] / l, sqrt(vare_ / l));
      }
      for (int k = 0; k < n; ++k) diff[k] = estR[off + lev[k]] - estR_tmp[off + lev[k]];
      daxpy(n, 1.0, diff, yadj);
      vrtmp[i] = (ddot(qr, estR_tmp + off, estR_tmp + off) + s2r * dfr) /
                 chisq_at(HB_DOM_ITER, it, HB_IT_VR0 + (uint32_t)i, 0, qr + dfr);
      vrv[i] = hbo_var(estR_tmp + off, qr);
      for (int q = 0; q < qr; ++q) estR[off + q] = estR_tmp[off + q];
    }

    /* BSLMM polygenic term :518-552 (block Gibbs sampler on the eigen-decomposition K diag(Kval) K') */
    if (nk) {
      const double* K = a->Ki;
      for (int i = 0; i < n; ++i) k_rhs[i] = yadj[i] + k_tmp[i];                                  /* :519 */
      double emax = 0.0;
      for (int j = 0; j < nk; ++j) {                                                              /* :531 */
        k_eval[j] = (a->Kival[j] * vare_) / (a->Kival[j] + vare_ / vbtmp);
        if (fabs(k_eval[j]) > emax) emax = fabs(k_eval[j]);
      }
      for (int j = 0; j < nk; ++j) k_t[j] = ddot(n, K + (size_t)j * n, k_rhs);                    /* K.t() * k_RHS */
      for (int j = 0; j < nk; ++j) k_w[j] = (k_eval[j] / vare_) * k_t[j];
      for (int i = 0; i < n; ++i) k_tmp[i] = 0.0;
      for (int j = 0; j < nk; ++j) daxpy(n, k_w[j], K + (size_t)j * n, k_tmp);                    /* :532 */
      for (int j = 0; j < nk; ++j)
        if (!(k_eval[j] >= -1e-06 * emax)) {                                                      /* :533 */
          FREE_ALL();
          return fail("matrix is not positive definite, try to specify parameter 'lambda' with a small value, eg: 0.001 or bigger");
        }
      for (int j = 0; j < nk; ++j) {                                                              /* :534-535 */
        if (k_eval[j] < 0) k_eval[j] = 0.0;
        k_w[j] = sqrt(k_eval[j]) * z_at(HB_DOM_K, it, (uint32_t)j, 0, 0);
      }
      for (int i = 0; i < n; ++i) k_rhs[i] = 0.0;
      for (int j = 0; j < nk; ++j) daxpy(n, k_w[j], K + (size_t)j * n, k_rhs);
      for (int i = 0; i < n; ++i) k_tmp[i] += k_rhs[i];
      daxpy(nk, -1.0, k_tmp, k_estR);                                                             /* :537-538 */
      daxpy(n, 1.0, k_estR, yadj);                                                                /* :539 */
      daxpy(n, -1.0, k_estR, u);                                                                  /* :540 */
      vbtmp = 0.0;
      for (int j = 0; j < nk; ++j) {                                                              /* :543-544 */
        const double kg = ddot(n, K + (size_t)j * n, k_tmp);
        vbtmp += kg * ((1 / a->Kival[j]) * kg);
      }
      vbtmp += s2vara_ * dfvara_;
      vbtmp /= chisq_at(HB_DOM_ITER, it, HB_IT_VB, 0, dfvara_ + nk);                              /* :546-547 */
      for (int j = 0; j < nk; ++j) k_estR[j] = k_tmp[j];                                          /* :551 */
    }

    /* single-step J + epsilon :554-584, solver.cpp:131-140 */
    if (ne) {
      oldgi = epsl_J_beta;
      v = JtJ;
      rhs = ddot(n, a->epsl_y_J, yadj);
      rhs += v * oldgi;
      gi = norm_at(HB_DOM_ITER, it, HB_IT_J, 0, rhs / v, sqrt(vare_ / v));
      gi_ = oldgi - gi;
      daxpy(n, gi_, a->epsl_y_J, yadj);
      gi_ *= -1;
      daxpy(n, gi_, a->epsl_y_J, u);
      epsl_J_beta = gi;
      const double ratio = vare_ / vepstmp;
      for (int q = 0; q < qe; ++q) e_rhs[q] = 0.0;
      for (int i = 0; i < ne; ++i) e_rhs[a->epsl_index[i] - 1] += yadj[n - ne + i];
      for (int q = 0; q < qe; ++q) e_rhs[q] += e_cnt[q] * e_tmp[q];
      for (int i = 0; i < qe; ++i) { /* Gibbs(sp_mat) */
        double aii = e_cnt[i], Ax = 0.0;
        int have_diag = 0;
        for (int p = a->Gi_colptr[i]; p < a->Gi_colptr[i + 1]; ++p) {
          int rix = a->Gi_rowidx[p];
          double aval = a->Gi_val[p] * ratio + (rix == i ? e_cnt[i] : 0.0);
          if (rix == i) { aii = aval; have_diag = 1; }
          Ax += aval * e_tmp[rix];
        }
        if (!have_diag) Ax += e_cnt[i] * e_tmp[i];
        double invlhs = 1.0 / aii;
        double uu = invlhs * (e_rhs[i] - Ax) + e_tmp[i];
        e_tmp[i] = norm_at(HB_DOM_EPS, it, (uint32_t)i, 0, uu, sqrt(invlhs * vare_));
      }
      for (int q = 0; q < qe; ++q) e_estR[q] -= e_tmp[q];
      for (int i = 0; i < ne; ++i) {
        double d = e_estR[a->epsl_index[i] - 1];
        yadj[n - ne + i] += d;
        u[n - ne + i] -= d;
      }
      vepstmp = 0.0;
      for (int c = 0; c < qe; ++c) {
        double colsum = 0.0;
        for (int p = a->Gi_colptr[c]; p < a->Gi_colptr[c + 1]; ++p) colsum += a->Gi_val[p] * e_tmp[a->Gi_rowidx[p]];
        vepstmp += colsum * e_tmp[c];
      }
      vepstmp += s2vara_ * dfvara_;
      vepstmp /= chisq_at(HB_DOM_ITER, it, HB_IT_VEPS, 0, dfvara_ + qe);
      for (int q = 0; q < qe; ++q) e_estR[q] = e_tmp[q];
      veps = vepstmp;
    }

    double t0 = now_s();
    switch (model_index) {
      case 1: /* BayesRR :587-606 */
        for (int i = 0; i < m; ++i) {
          if (!vx[i]) continue;
          const double* dxi = xcol(a, i, scratch);
          xx = xpx[i];
          oldgi = g[i];
          rhs = ddot(n, dxi, yadj);
          rhs += xx * oldgi;
          v = xx + vare_ / varg;
          gi = norm_at(HB_DOM_SNP, it, (uint32_t)i, HB_SL_MAIN, rhs / v, sqrt(vare_ / v));
          gi_ = oldgi - gi;
          daxpy(n, gi_, dxi, yadj);
          gi_ *= -1;
          daxpy(n, gi_, dxi, u);
          g[i] = gi;
        }
        varg = (ddot(m, g, g) + s2varg_ * dfvara_) / chisq_at(HB_DOM_ITER, it, HB_IT_VARG, 0, dfvara_ + m - nvar0);
        break;
      case 2: /* BayesA :607-626 */
        for (int i = 0; i < m; ++i) {
          if (!vx[i]) continue;
          const double* dxi = xcol(a, i, scratch);
          xx = xpx[i];
          oldgi = g[i];
          varg = (oldgi * oldgi + s2varg_ * dfvara_) / chisq_at(HB_DOM_SNP, it, (uint32_t)i, HB_SL_CHI, dfvara_ + 1);
          rhs = ddot(n, dxi, yadj);
          rhs += xx * oldgi;
          v = xx + vare_ / varg;
          gi = norm_at(HB_DOM_SNP, it, (uint32_t)i, HB_SL_MAIN, rhs / v, sqrt(vare_ / v));
          gi_ = oldgi - gi;
          daxpy(n, gi_, dxi, yadj);
          gi_ *= -1;
          daxpy(n, gi_, dxi, u);
          g[i] = gi;
        }
        break;
      case 3: /* BayesB / BayesBpi :627-670 */
      case 4: /* BayesC / BayesCpi / BSLMM :671-717 */
        for (int j = 0; j < n_fold; ++j) logpi[j] = log(Pi[j]);
        s[0] = logpi[0];
        vargi = 0;
        for (int i = 0; i < m; ++i) {
          if (!vx[i]) continue;
          const double* dxi = xcol(a, i, scratch);
          xx = xpx[i];
          oldgi = g[i];
          if (model_index == 3)
            varg = (oldgi * oldgi + s2varg_ * dfvara_) / chisq_at(HB_DOM_SNP, it, (uint32_t)i, HB_SL_CHI, dfvara_ + 1);
          rhs = ddot(n, dxi, yadj);
          if (oldgi) rhs += xx * oldgi;
          lhs = xx / vare_;
          logdetV = log(varg * lhs + 1);
          uhat = rhs / (xx + vare_ / varg);
          s[1] = -0.5 * (logdetV - (rhs * uhat / vare_)) + logpi[1];
          { double t = 0.0; /* sum(exp(s - s[0])) */
            double e0 = exp(s[0] - s[0]), e1 = exp(s[1] - s[0]);
            t = e0 + e1;
            acceptProb = 1 / t; }
          double rval, zval;
          hb_draw_uz(KEY, HB_DOM_SNP, it, (uint32_t)i, HB_SL_MAIN, 0, &rval, &zval);
          TAPE(TP_U, rval, 0.0);
          indistflag = rval < acceptProb ? 0 : 1;
          snptracker[i] = indistflag;
          if (indistflag) {
            v = xx + vare_ / varg;
            TAPE(TP_Z, zval, 0.0);
            gi = rhs / v + sqrt(vare_ / v) * zval;
            gi_ = oldgi - gi;
            daxpy(n, gi_, dxi, yadj);
            gi_ *= -1;
            daxpy(n, gi_, dxi, u);
            if (model_index == 4) vargi += (gi * gi);
          } else {
            gi = 0;
            if (oldgi) {
              gi_ = oldgi;
              daxpy(n, gi_, dxi, yadj);
              gi_ *= -1;
              daxpy(n, gi_, dxi, u);
            }
          }
          g[i] = gi;
        }
        fold_snp_num[1] = acc_sum(snptracker, m);
        fold_snp_num[0] = m - nvar0 - fold_snp_num[1];
        NnzSnp = (int)fold_snp_num[1];
        if (model_index == 4)
          varg = (vargi + s2varg_ * dfvara_) / chisq_at(HB_DOM_ITER, it, HB_IT_VARG, 0, dfvara_ + NnzSnp);
        if (!fixpi) { /* rdirichlet_sample stats.cpp:69-76 */
          double tot = 0.0;
          for (int j = 0; j < n_fold; ++j) {
            Pi[j] = gamma_at(HB_DOM_ITER, it, HB_IT_PI0 + (uint32_t)j, 0, fold_snp_num[j] + 1);
          }
          tot = acc_sum(Pi, n_fold);
          for (int j = 0; j < n_fold; ++j) Pi[j] /= tot;
        }
        break;
      case 5: /* BayesL :718-742 */
        for (int i = 0; i < m; ++i) {
          if (!vx[i]) continue;
          const double* dxi = xcol(a, i, scratch);
          xx = xpx[i];
          oldgi = g[i];
          rhs = ddot(n, dxi, yadj);
          rhs += xx * oldgi;
          v = xx + 1 / vargL[i];
          gi = norm_at(HB_DOM_SNP, it, (uint32_t)i, HB_SL_MAIN, rhs / v, sqrt(vare_ / v));
          if (fabs(gi) < 1e-6) gi = 1e-6;
          { double uu, zz;
            hb_draw_uz(KEY, HB_DOM_SNP, it, (uint32_t)i, HB_SL_IG, 0, &uu, &zz);
            TAPE(TP_Z, zz, 0.0); TAPE(TP_U, uu, 0.0);
            vargi = 1 / hb_invgauss_from_uz(sqrt(vare_) * lambda / fabs(gi), lambda2, uu, zz); }
          if (vargi >= 0) vargL[i] = vargi;
          gi_ = oldgi - gi;
          daxpy(n, gi_, dxi, yadj);
          gi_ *= -1;
          daxpy(n, gi_, dxi, u);
          g[i] = gi;
        }
        shape = shape0 + m - nvar0;
        rate = rate0 + acc_sum(vargL, m) / 2;
        lambda2 = gamma_at(HB_DOM_ITER, it, HB_IT_LAMBDA, 0, shape) * (1 / rate);
        lambda = sqrt(lambda2);
        break;
      case 6: /* BayesR :743-815 */
        for (int j = 0; j < n_fold; ++j) logpi[j] = log(Pi[j]);
        s[0] = logpi[0];
        varg = 0;
        for (int j = 1; j < n_fold; ++j) vare_vara_fold[j] = vare_ / vara_fold[j];
        for (int i = 0; i < m; ++i) {
          if (!vx[i]) continue;
          const double* dxi = xcol(a, i, scratch);
          xx = xpx[i];
          oldgi = g[i];
          rhs = ddot(n, dxi, yadj);
          if (oldgi) { rhs += xx * oldgi; }
          lhs = xx / vare_;
          for (int j = 1; j < n_fold; ++j) {
            logdetV = log(vara_fold[j] * lhs + 1);
            uhat = rhs / (xx + vare_vara_fold[j]);
            s[j] = -0.5 * (logdetV - (rhs * uhat / vare_)) + logpi[j];
          }
          for (int j = 0; j < n_fold; ++j) {
            double temp = 0.0;
            for (int k = 0; k < n_fold; ++k) temp += exp(s[k] - s[j]);
            stemp[j] = 1 / temp;
          }
          acceptProb = 0;
          indistflag = 0;
          double rval, zval;
          hb_draw_uz(KEY, HB_DOM_SNP, it, (uint32_t)i, HB_SL_MAIN, 0, &rval, &zval);
          TAPE(TP_U, rval, 0.0);
          for (int j = 0; j < n_fold; ++j) {
            acceptProb += stemp[j];
            if (rval < acceptProb) { indistflag = j; break; }
          }
          snptracker[i] = indistflag;
          if (indistflag) {
            v = xx + vare_vara_fold[indistflag];
            TAPE(TP_Z, zval, 0.0);
            gi = rhs / v + sqrt(vare_ / v) * zval;
            gi_ = oldgi - gi;
            daxpy(n, gi_, dxi, yadj);
            gi_ *= -1;
            daxpy(n, gi_, dxi, u);
            varg += (gi * gi / fold_[indistflag]);
          } else {
            gi = 0;
            if (oldgi) {
              gi_ = oldgi;
              daxpy(n, gi_, dxi, yadj);
              gi_ *= -1;
              daxpy(n, gi_, dxi, u);
            }
          }
          g[i] = gi;
        }
        for (int j = 0; j < n_fold; ++j) {
          double c = 0;
          for (int i = 0; i < m; ++i) c += (snptracker[i] == j);
          fold_snp_num[j] = c;
        }
        NnzSnp = m - (int)fold_snp_num[0];
        varg = (varg + s2varg_ * dfvara_) / chisq_at(HB_DOM_ITER, it, HB_IT_VARG, 0, dfvara_ + NnzSnp);
        for (int j = 0; j < n_fold; ++j) vara_fold[j] = varg * fold_[j];
        fold_snp_num[0] -= nvar0;
        if (!fixpi) {
          for (int j = 0; j < n_fold; ++j)
            Pi[j] = gamma_at(HB_DOM_ITER, it, HB_IT_PI0 + (uint32_t)j, 0, fold_snp_num[j] + 1);
          double tot = acc_sum(Pi, n_fold);
          for (int j = 0; j < n_fold; ++j) Pi[j] /= tot;
        }
        break;
    }
    t_sweep += now_s() - t0;

    /* :819 genetic variance, :823 residual variance */
    vara_ = hbo_var(u, n);
    vare_ = (ddot(n, yadj, yadj) + s2vare_ * dfvare_) / chisq_at(HB_DOM_ITER, it, HB_IT_VARE, 0, n + dfvare_);

    if (o->nnz_trace) o->nnz_trace[iter] = NnzSnp;
    if (o->vara_trace) o->vara_trace[iter] = vara_;
    if (o->vare_trace) o->vare_trace[iter] = vare_;
    if (o->varg_trace) o->varg_trace[iter] = varg;

    /* :826-845 PIP / WPPA counters */
    if (iter >= nburn) {
      if (have_tracker)
        for (int i = 0; i < m; ++i) if (snptracker[i]) nzrate[i] += 1;
      if (WPPA) {
        for (int w = 0; w < nw; ++w) {
          int any = 0;
          for (int p = wstart[w]; p < wstart[w + 1]; ++p) if (snptracker[wmembers[p]]) { any = 1; break; }
          if (any) wppai[w] += 1;
        }
      }
      nzct++;
    }

    /* :848-882 record */
    if (iter >= nburn && (iter + 1 - nburn) % thin == 0) {
      musum += mu;
      if (o->mu_store) o->mu_store[count] = mu;
      if (!fixpi) {
        for (int j = 0; j < n_fold; ++j) pisum[j] += Pi[j];
        if (o->pi_store) for (int j = 0; j < n_fold; ++j) o->pi_store[(size_t)count * n_fold + j] = Pi[j];
      }
      varasum += vara_; varesum += vare_;
      if (o->vara_store) o->vara_store[count] = vara_;
      if (o->vare_store) o->vare_store[count] = vare_;
      if (nk) for (int j = 0; j < nk; ++j) k_sum[j] += k_estR[j];                                 /* :858 */
      for (int i = 0; i < m; ++i) gsum[i] += g[i];
      if (o->alpha_store) memcpy(o->alpha_store + (size_t)count * m, g, sizeof(double) * m);
      if (nc) {
        for (int i = 0; i < nc; ++i) betasum[i] += beta[i];
        if (o->beta_store) memcpy(o->beta_store + (size_t)count * nc, beta, sizeof(double) * nc);
      }
      double vt = vara_ + vare_;
      if (nr) {
        for (int i = 0; i < nr; ++i) { vt += vrv[i]; vrsum[i] += vrv[i]; }
        for (int q = 0; q < n_levels; ++q) estRsum[q] += estR[q];
        if (o->vr_store) memcpy(o->vr_store + (size_t)count * nr, vrv, sizeof(double) * nr);
        if (o->estR_store) memcpy(o->estR_store + (size_t)count * n_levels, estR, sizeof(double) * n_levels);
      }
      if (ne) {
        vepssum += veps; Jsum += epsl_J_beta;
        for (int q = 0; q < qe; ++q) e_sum[q] += e_estR[q];
        if (o->veps_store) o->veps_store[count] = veps;
        if (o->J_store) o->J_store[count] = epsl_J_beta;
        if (o->epsilon_store) memcpy(o->epsilon_store + (size_t)count * qe, e_estR, sizeof(double) * qe);
      }
      hsqsum += vara_ / vt;
      if (o->hsq_store) o->hsq_store[count] = vara_ / vt;
      count++;
    }
    if (count == n_records) { ++iter; break; } /* :916 */
  }
  o->iters_done = iter;
  o->n_records_done = count;
  o->nzct = nzct;
  o->seconds_sweep = t_sweep;

  /* :919-1040 posterior summaries (means of the stored records) */
  const double rc = (double)count;
  o->Vg = varasum / rc; o->Ve = varesum / rc; o->h2 = hsqsum / rc;
  const double Mu = musum / rc;
  o->mu = Mu;
  if (o->e) for (int i = 0; i < n; ++i) o->e[i] = a->y[i] - Mu * 1.0;
  if (nc) {
    for (int i = 0; i < nc; ++i) {
      double b = betasum[i] / rc;
      if (o->beta) o->beta[i] = b;
      if (o->e) for (int k = 0; k < n; ++k) o->e[k] -= a->C[(size_t)i * n + k] * b;
    }
  }
  for (int i = 0; i < m; ++i) gsum[i] /= rc;
  if (nk) { /* :955-964: the polygenic values expressed as SNP effects and added to every stored sample */
    const double* K = a->Ki;
    for (int j = 0; j < nk; ++j) k_sum[j] /= rc;
    for (int j = 0; j < nk; ++j) k_t[j] = (ddot(n, K + (size_t)j * n, k_sum) / a->Kival[j]) / sumvx;   /* Kg */
    for (int i = 0; i < n; ++i) k_rhs[i] = 0.0;
    for (int j = 0; j < nk; ++j) daxpy(n, k_t[j], K + (size_t)j * n, k_rhs);                          /* K * Kg */
    double* ghat = (double*)malloc(sizeof(double) * m);
    for (int i = 0; i < m; ++i) ghat[i] = ddot(n, xcol(a, i, scratch), k_rhs);                        /* X.t() * (K * Kg) */
    const double gm = acc_mean(ghat, m);
    for (int i = 0; i < m; ++i) ghat[i] -= gm;
    for (int i = 0; i < m; ++i) gsum[i] += ghat[i];
    if (o->alpha_store)
      for (int c = 0; c < count; ++c)
        for (int i = 0; i < m; ++i) o->alpha_store[(size_t)c * m + i] += ghat[i];
    free(ghat);
  }
  if (o->alpha) memcpy(o->alpha, gsum, sizeof(double) * m);
  if (o->e) {
    for (int i = 0; i < m; ++i) {
      if (gsum[i] == 0.0) continue;
      const double* dxi = xcol(a, i, scratch);
      daxpy(n, -gsum[i], dxi, o->e);
    }
  }
  if (o->pi) {
    if (!fixpi) for (int j = 0; j < n_fold; ++j) o->pi[j] = pisum[j] / rc;
    else for (int j = 0; j < n_fold; ++j) o->pi[j] = Pi[j];
  }
  if (fixpi && o->pi_store)
    for (int c = 0; c < count; ++c) { o->pi_store[(size_t)c * n_fold] = Pi[0]; o->pi_store[(size_t)c * n_fold + 1] = Pi[1]; }
  if (ne) {
    o->Veps = vepssum / rc; o->J = Jsum / rc;
    if (o->e) {
      for (int k = 0; k < n; ++k) o->e[k] -= o->J * a->epsl_y_J[k];
      for (int i = 0; i < ne; ++i) o->e[n - ne + i] -= e_sum[a->epsl_index[i] - 1] / rc;
    }
    if (o->epsilon) for (int q = 0; q < qe; ++q) o->epsilon[q] = e_sum[q] / rc;
  }
  if (nr) {
    for (int i = 0; i < nr; ++i) if (o->vr) o->vr[i] = vrsum[i] / rc;
    for (int q = 0; q < n_levels; ++q) estRsum[q] /= rc;
    if (o->estR) memcpy(o->estR, estRsum, sizeof(double) * n_levels);
    if (o->e)
      for (int i = 0; i < nr; ++i)
        for (int k = 0; k < n; ++k) o->e[k] -= estRsum[R_off[i] + a->Rlev[(size_t)i * n + k]];
  }
  if (o->g) memcpy(o->g, u, sizeof(double) * n);
  if (o->nzrate_count) memcpy(o->nzrate_count, nzrate, sizeof(double) * m);
  if (o->tracker_final) for (int i = 0; i < m; ++i) o->tracker_final[i] = (int32_t)snptracker[i];
  if (o->pip) { /* :1026-1032 */
    if (!have_tracker) for (int i = 0; i < m; ++i) o->pip[i] = 1.0;
    else for (int i = 0; i < m; ++i) {
      double r = nzrate[i] / nzct;
      if (r == 1) r = (nzct - 1) / (double)nzct;
      o->pip[i] = r;
    }
  }
  if (WPPA) {
    if (o->wppa_count) memcpy(o->wppa_count, wppai, sizeof(double) * nw);
    if (o->gwas) for (int w = 0; w < nw; ++w) {
      double r = wppai[w] / nzct;
      if (r == 1) r = (nzct - 1) / (double)nzct;
      o->gwas[w] = r;
    }
  }
  free(vrsum);
  FREE_ALL();
  return 0;
}

/* ---- CPU timing of the reference's level-1 data path ------------------------ */
/* BayesR sweep on a column-major fp64 X exactly as Bayes.cpp:751-802 drives it:
 * one ddot per SNP, two daxpy per changed SNP; vector ops split over OpenMP
 * threads (the stand-in for a threaded BLAS).  X ~ Binomial(2, p_j). */
static double par_ddot(int n, const double* x, const double* y, int threads) {
  double s = 0.0;
  (void)threads;
#pragma omp parallel for reduction(+ : s) num_threads(threads) schedule(static)
  for (int i = 0; i < n; ++i) s += x[i] * y[i];
  return s;
}
static void par_daxpy2(int n, double a, const double* x, double* y, double* u, int threads) {
  (void)threads;
#pragma omp parallel for num_threads(threads) schedule(static)
  for (int i = 0; i < n; ++i) { y[i] += a * x[i]; u[i] -= a * x[i]; }
}

double hbo_time_sweep_fp64(int n, int m_cpu, int sweeps, int threads, uint64_t seed, double* checksum) {
  hb_key_t key = hb_make_key(seed);
  double* X = (double*)malloc(sizeof(double) * (size_t)n * m_cpu);
  double* yadj = (double*)malloc(sizeof(double) * n);
  double* u = (double*)calloc(n, sizeof(double));
  double* g = (double*)calloc(m_cpu, sizeof(double));
  double* xpx = (double*)malloc(sizeof(double) * m_cpu);
  if (!X || !yadj || !u || !g || !xpx) { free(X); free(yadj); free(u); free(g); free(xpx); return -1.0; }
#pragma omp parallel for num_threads(threads) schedule(static)
  for (int j = 0; j < m_cpu; ++j) {
    uint32_t w[4];
    hb_philox4x32_10((uint32_t)j, 0, 0, 77u, key.k0, key.k1, w);
    double p = 0.05 + 0.45 * hb_u01(w[0], w[1]);
    double ss = 0;
    double* col = X + (size_t)j * n;
    for (int i = 0; i < n; i += 2) {
      hb_philox4x32_10((uint32_t)j, (uint32_t)i, 1, 77u, key.k0, key.k1, w);
      double a0 = hb_u01(w[0], w[1]), a1 = hb_u01(w[2], w[3]);
      double q2 = (1 - p) * (1 - p), q1 = q2 + 2 * p * (1 - p);
      col[i] = a0 < q2 ? 0 : a0 < q1 ? 1 : 2;
      if (i + 1 < n) col[i + 1] = a1 < q2 ? 0 : a1 < q1 ? 1 : 2;
    }
    for (int i = 0; i < n; ++i) ss += col[i] * col[i];
    xpx[j] = ss;
  }
  for (int i = 0; i < n; ++i) yadj[i] = hb_draw_z(key, 9u, 0, (uint32_t)i, 0, 0);
  const double fold[4] = {0, 1e-4, 1e-3, 1e-2};
  const double logpi[4] = {log(0.95), log(0.02), log(0.02), log(0.01)};
  double vare = 0.5, varg = 0.5 / (0.05 * 0.3 * m_cpu);
  double t_total = 0.0;
  long updates = 0;
  for (int sw = -1; sw < sweeps; ++sw) { /* sw = -1 is the warm-up sweep */
    double t0 = now_s();
    for (int i = 0; i < m_cpu; ++i) {
      const double* dxi = X + (size_t)i * n;
      double xx = xpx[i], oldgi = g[i];
      double rhs = par_ddot(n, dxi, yadj, threads);
      if (oldgi) rhs += xx * oldgi;
      double lhs = xx / vare, s[4], stemp[4];
      s[0] = logpi[0];
      for (int j = 1; j < 4; ++j) {
        double logdetV = log(varg * fold[j] * lhs + 1);
        double uhat = rhs / (xx + vare / (varg * fold[j]));
        s[j] = -0.5 * (logdetV - (rhs * uhat / vare)) + logpi[j];
      }
      for (int j = 0; j < 4; ++j) {
        double t = 0;
        for (int k = 0; k < 4; ++k) t += exp(s[k] - s[j]);
        stemp[j] = 1 / t;
      }
      double rval, zval, acc = 0;
      int cls = 0;
      hb_draw_uz(key, HB_DOM_SNP, (uint32_t)(sw + 1), (uint32_t)i, HB_SL_MAIN, 0, &rval, &zval);
      for (int j = 0; j < 4; ++j) { acc += stemp[j]; if (rval < acc) { cls = j; break; } }
      double gi = 0;
      if (cls) {
        double v = xx + vare / (varg * fold[cls]);
        gi = rhs / v + sqrt(vare / v) * zval;
        par_daxpy2(n, oldgi - gi, dxi, yadj, u, threads);
      } else if (oldgi) {
        par_daxpy2(n, oldgi, dxi, yadj, u, threads);
      }
      g[i] = gi;
    }
    if (sw >= 0) { t_total += now_s() - t0; updates += m_cpu; }
  }
  double cs = 0;
  for (int i = 0; i < n; ++i) cs += yadj[i];
  if (checksum) *checksum = cs;
  free(X); free(yadj); free(u); free(g); free(xpx);
  return (double)updates / t_total;
}

/* ============================================================================================
 * SBayesD -- literal restatement of /root/reference/src/SBayesD.cpp:5-609 (dense LD matrix).
 * Same random-number addresses as hbo_bayes (hb_rng.h); per-iteration extras: HB_IT_VARA for the
 * genetic variance (:461) and HB_IT_VARE for the residual variance (:467).
 * ============================================================================================ */
#define HB_ORACLE_MAX_FOLD 16
/* sparse = 1: SBayesS.cpp -- the LD matrix is a sparse column-compressed matrix, the residual variance of SNP i is
 * inflated to varei = varediff_i * vara + vare (:131-141, :285), and BayesC/Cpi and BayesR re-draw effects that would
 * explain more than the phenotypic variance (:388-398, :489-499; note `vargi = gi * gi` inside that loop, kept). */
static int sbayes_impl(const hbo_sbayes_args* a, hbo_sbayes_out* o, int sparse) {
  if (!a || !o) return fail("null argument");
  if (sparse ? !(a->ld_colptr && a->ld_rowidx && a->ld_val) : !a->ldm) return fail("LD matrix missing");
  KEY = hb_make_key(a->seed);
  const int m = a->m;
  const char* model = a->model;
  /* :28 (no BSLMM here) */
  const int model_index = !strcmp(model, "BayesRR") ? 1 : !strcmp(model, "BayesA") ? 2 :
                          (!strcmp(model, "BayesB") || !strcmp(model, "BayesBpi")) ? 3 :
                          (!strcmp(model, "BayesC") || !strcmp(model, "BayesCpi")) ? 4 : !strcmp(model, "BayesL") ? 5 : 6;
  const double* ss = a->sumstat;
#define SS(k, c) ss[(size_t)(c) * m + (k)]
  /* :33-34  int n = mean(finite N) */
  int n;
  { double acc = 0.0; int cnt = 0;
    for (int k = 0; k < m; ++k) if (isfinite(SS(k, 3))) { acc += SS(k, 3); cnt++; }
    if (!cnt) return fail("Lack of SE.");
    n = (int)(acc / cnt); }
  int fixpi = (!strcmp(model, "BayesB") || !strcmp(model, "BayesC"));
  const int n_fold = a->n_fold;
  if (n_fold < 2) return fail("Pi should be a vector.");
  if (acc_sum(a->Pi, n_fold) != 1) return fail("sum of Pi should be 1.");
  if (a->Pi[0] == 1) return fail("all markers have no effect size.");
  for (int i = 0; i < n_fold; ++i)
    if (a->Pi[i] < 0 || a->Pi[i] > 1) return fail("elements of Pi should be at the range of [0, 1]");
  double Pi[HB_ORACLE_MAX_FOLD], fold_[HB_ORACLE_MAX_FOLD];
  if (n_fold > HB_ORACLE_MAX_FOLD) return fail("too many mixture components");
  for (int i = 0; i < n_fold; ++i) { Pi[i] = a->Pi[i]; fold_[i] = a->fold ? a->fold[i] : 0.0; }
  if (!a->fold) {
    if (!strcmp(model, "BayesR")) return fail("'fold' should be provided for BayesR model.");
    if (n_fold != 2) return fail("length of Pi and fold not equals.");
  }
  const int niter = a->niter, nburn = a->nburn, thin = a->thin;
  const int n_records = (niter - nburn) / thin;
  int count = 0, nzct = 0, NnzSnp = 0, have_tracker = 0;
  if (!strcmp(model, "BayesRR") || !strcmp(model, "BayesA") || !strcmp(model, "BayesL")) {
    NnzSnp = m; Pi[0] = 0; Pi[1] = 1; fixpi = 1;
  } else {
    if (strcmp(model, "BayesR") && n_fold != 2)
      return fail("length of Pi should be 2, the first value is the proportion of non-effect markers.");
    have_tracker = 1;
  }
  double* xy = calloc(m, 8); double* r_hat = calloc(m, 8); double* tmp = calloc(m, 8); double* yyi = calloc(m, 8);
  double* g = calloc(m, 8); double* xpx = calloc(m, 8); double* vx = calloc(m, 8); double* snptracker = calloc(m, 8);
  double* nzrate = calloc(m, 8); double* gsum = calloc(m, 8); double* vargL = calloc(m, 8);
  uint8_t* ifest = malloc(m);
  const double* ldm = a->ldm;
  double* varediff = calloc(m, 8);
  for (int i = 0; i < m; ++i) {   /* :92-96 (SBayesS :109-113, :131-141) */
    if (sparse) {
      vx[i] = 0.0;
      for (int q = a->ld_colptr[i]; q < a->ld_colptr[i + 1]; ++q) if (a->ld_rowidx[q] == i) vx[i] = a->ld_val[q];
      varediff[i] = (m - (double)(a->ld_colptr[i + 1] - a->ld_colptr[i])) / m;
    } else vx[i] = ldm[(size_t)i * m + i];
    xpx[i] = vx[i] * n;
  }
#define LD_UPDATE(i, gi_) do { if (sparse) { for (int q_ = a->ld_colptr[i]; q_ < a->ld_colptr[(i) + 1]; ++q_) r_hat[a->ld_rowidx[q_]] += (gi_) * a->ld_val[q_]; } \
                               else daxpy(m, (gi_), ldm + (size_t)(i) * m, r_hat); } while (0)
#define VAREI(i) (sparse ? varediff[i] * vara_ + vare_ : vare_)
/* SBayesS.cpp:388-398 / :489-499 */
#define REDRAW_LOOP(i) do { if (sparse && (gi * gi * vx[i]) > vary) { int ii = 0; \
    while ((gi * gi * vx[i]) > vary) { gi = rhs / v + sqrt(varei / v) * z_at(HB_DOM_SNP, it, (uint32_t)(i), HB_SL_RETRY, (uint32_t)(ii + 1)); \
      vargi = gi * gi; ii++; if (ii > 100) gi = 0; } } } while (0)
  int count_y = 0, nvar0 = 0;
  for (int k = 0; k < m; ++k) {   /* :100-112 */
    ifest[k] = 1;
    if (isnan(SS(k, 1)) || isnan(SS(k, 2)) || isnan(SS(k, 3))) { ifest[k] = 0; nvar0++; }
    else {
      xy[k] = xpx[k] * SS(k, 1);
      r_hat[k] = xy[k];
      yyi[k] = xpx[k] * (SS(k, 1) * SS(k, 1) + (SS(k, 3) - 2) * SS(k, 2) * SS(k, 2));
      count_y++;
    }
  }
  if (count_y == 0) return fail("Lack of SE.");
  const double yy = acc_sum(yyi, m) / count_y;
  const double vary = yy / (n - 1);
  const double h2 = 0.5;
  const double dfvara_ = isnan(a->dfvg) ? 4 : a->dfvg;
  if (dfvara_ <= 2) return fail("dfvg should not be less than 2.");
  double vara_ = isnan(a->vg) ? ((dfvara_ - 2) / dfvara_) * vary * h2 : a->vg;
  double vare_ = isnan(a->ve) ? vary * (1 - h2) : a->ve;
  const double dfvare_ = isnan(a->dfve) ? -2 : a->dfve;
  const double s2vara_ = isnan(a->s2vg) ? vara_ * (dfvara_ - 2) / dfvara_ : a->s2vg;
  const double sumvx = acc_sum(vx, m);
  double varg = vara_ / ((1 - Pi[0]) * sumvx);
  const double s2varg_ = s2vara_ / ((1 - Pi[0]) * sumvx);
  const double s2vare_ = isnan(a->s2ve) ? 0 : a->s2ve;
  if (niter < nburn) return fail("Number of total iteration ('niter') shold be larger than burn-in ('nburn').");
  const double R2 = (dfvara_ - 2) / dfvara_;
  double lambda2 = 2 * (1 - R2) / (R2)*sumvx;
  double lambda = sqrt(lambda2);
  const double shape0 = 1.1, rate0 = (shape0 - 1) / lambda2;
  if (model_index == 5) for (int i = 0; i < m; ++i) vargL[i] = varg;
  double fold_snp_num[HB_ORACLE_MAX_FOLD] = {0}, logpi[HB_ORACLE_MAX_FOLD], s[HB_ORACLE_MAX_FOLD], stemp[HB_ORACLE_MAX_FOLD];
  double vara_fold[HB_ORACLE_MAX_FOLD], vare_vara_fold[HB_ORACLE_MAX_FOLD] = {0}, pisum[HB_ORACLE_MAX_FOLD] = {0};
  for (int j = 0; j < n_fold; ++j) vara_fold[j] = (vara_ / ((1 - Pi[0]) * sumvx)) * fold_[j];
  int nw = 0;
  double* wppai = NULL;
  if (a->windindx) { for (int i = 0; i < m; ++i) if (a->windindx[i] > nw) nw = a->windindx[i]; wppai = calloc(nw > 0 ? nw : 1, 8); }
  double varasum = 0, varesum = 0, hsqsum = 0;
  int iter;
  for (iter = 0; iter < niter; ++iter) {
    const uint32_t it = (uint32_t)iter;
    double xx, gi, gi_, rhs, lhs, logdetV, acceptProb, uhat, v, vargi;
    int indistflag;
    switch (model_index) {
      case 1: /* :254-270 */
        for (int i = 0; i < m; ++i) {
          if (!ifest[i]) continue;
          const double varei = VAREI(i);
          xx = xpx[i]; gi = g[i]; rhs = r_hat[i];
          if (gi) rhs += xx * gi;
          v = xx + varei / varg;
          gi = norm_at(HB_DOM_SNP, it, (uint32_t)i, HB_SL_MAIN, rhs / v, sqrt(varei / v));
          gi_ = (g[i] - gi) * n;
          LD_UPDATE(i, gi_);
          g[i] = gi;
        }
        varg = (ddot(m, g, g) + s2varg_ * dfvara_) / chisq_at(HB_DOM_ITER, it, HB_IT_VARG, 0, dfvara_ + count_y);
        break;
      case 2: /* :273-289 */
        for (int i = 0; i < m; ++i) {
          if (!ifest[i]) continue;
          const double varei = VAREI(i);
          xx = xpx[i]; gi = g[i];
          varg = (gi * gi + s2varg_ * dfvara_) / chisq_at(HB_DOM_SNP, it, (uint32_t)i, HB_SL_CHI, dfvara_ + 1);
          rhs = r_hat[i];
          if (gi) rhs += xx * gi;
          v = xx + varei / varg;
          gi = norm_at(HB_DOM_SNP, it, (uint32_t)i, HB_SL_MAIN, rhs / v, sqrt(varei / v));
          gi_ = (g[i] - gi) * n;
          LD_UPDATE(i, gi_);
          g[i] = gi;
        }
        break;
      case 3: case 4: /* :290-364 */
        for (int j = 0; j < n_fold; ++j) logpi[j] = log(Pi[j]);
        s[0] = logpi[0];
        vargi = 0;
        for (int i = 0; i < m; ++i) {
          if (!ifest[i]) continue;
          const double varei = VAREI(i);
          xx = xpx[i]; gi = g[i];
          if (model_index == 3)
            varg = (gi * gi + s2varg_ * dfvara_) / chisq_at(HB_DOM_SNP, it, (uint32_t)i, HB_SL_CHI, dfvara_ + 1);
          rhs = r_hat[i];
          if (gi) rhs += xx * gi;
          lhs = xx / varei;
          logdetV = log(varg * lhs + 1);
          uhat = rhs / (xx + varei / varg);
          s[1] = -0.5 * (logdetV - (rhs * uhat / varei)) + logpi[1];
          acceptProb = 1 / (exp(s[0] - s[0]) + exp(s[1] - s[0]));
          double rval, zval;
          hb_draw_uz(KEY, HB_DOM_SNP, it, (uint32_t)i, HB_SL_MAIN, 0, &rval, &zval);
          TAPE(TP_U, rval, 0.0);
          indistflag = rval < acceptProb ? 0 : 1;
          snptracker[i] = indistflag;
          if (indistflag == 0) gi = 0;
          else {
            v = xx + varei / varg;
            TAPE(TP_Z, zval, 0.0);
            gi = rhs / v + sqrt(varei / v) * zval;
            if (model_index == 4) { REDRAW_LOOP(i); vargi += gi * gi; }
          }
          if (gi != g[i]) {
            gi_ = (g[i] - gi) * n;
            LD_UPDATE(i, gi_);
            g[i] = gi;
          }
        }
        fold_snp_num[1] = acc_sum(snptracker, m);
        fold_snp_num[0] = m - nvar0 - fold_snp_num[1];
        NnzSnp = (int)fold_snp_num[1];
        if (model_index == 4)
          varg = (vargi + s2varg_ * dfvara_) / chisq_at(HB_DOM_ITER, it, HB_IT_VARG, 0, dfvara_ + NnzSnp);
        if (!fixpi) {
          for (int j = 0; j < n_fold; ++j) Pi[j] = gamma_at(HB_DOM_ITER, it, HB_IT_PI0 + (uint32_t)j, 0, fold_snp_num[j] + 1);
          const double tot = acc_sum(Pi, n_fold);
          for (int j = 0; j < n_fold; ++j) Pi[j] /= tot;
        }
        break;
      case 5: /* :365-389 */
        for (int i = 0; i < m; ++i) {
          if (!ifest[i]) continue;
          const double varei = VAREI(i);
          xx = xpx[i]; gi = g[i]; rhs = r_hat[i];
          if (gi) rhs += xx * gi;
          v = xx + 1 / vargL[i];
          gi = norm_at(HB_DOM_SNP, it, (uint32_t)i, HB_SL_MAIN, rhs / v, sqrt(varei / v));
          if (fabs(gi) < 1e-6) gi = 1e-6;
          { double uu, zz;
            hb_draw_uz(KEY, HB_DOM_SNP, it, (uint32_t)i, HB_SL_IG, 0, &uu, &zz);
            TAPE(TP_Z, zz, 0.0); TAPE(TP_U, uu, 0.0);
            vargi = 1 / hb_invgauss_from_uz(sqrt(varei) * lambda / fabs(gi), lambda2, uu, zz); }
          if (vargi > 0) vargL[i] = vargi;
          if (gi != g[i]) {
            gi_ = (g[i] - gi) * n;
            LD_UPDATE(i, gi_);
            g[i] = gi;
          }
        }
        { const double shape = shape0 + count_y, rate = rate0 + acc_sum(vargL, m) / 2;
          lambda2 = gamma_at(HB_DOM_ITER, it, HB_IT_LAMBDA, 0, shape) * (1 / rate);
          lambda = sqrt(lambda2); }
        break;
      default: /* 6: BayesR :390-455 */
        for (int j = 0; j < n_fold; ++j) logpi[j] = log(Pi[j]);
        s[0] = logpi[0];
        varg = 0;
        for (int j = 1; j < n_fold; ++j) vare_vara_fold[j] = vare_ / vara_fold[j];
        for (int i = 0; i < m; ++i) {
          if (!ifest[i]) continue;
          const double varei = VAREI(i);
          xx = xpx[i]; gi = g[i]; rhs = r_hat[i];
          if (gi) rhs += xx * gi;
          lhs = xx / varei;
          for (int j = 1; j < n_fold; ++j) {
            logdetV = log(vara_fold[j] * lhs + 1);
            uhat = rhs / (xx + varei / vara_fold[j]);
            s[j] = -0.5 * (logdetV - (rhs * uhat / varei)) + logpi[j];
          }
          for (int j = 0; j < n_fold; ++j) {
            double temp = 0.0;
            for (int k = 0; k < n_fold; ++k) temp += exp(s[k] - s[j]);
            stemp[j] = 1 / temp;
          }
          acceptProb = 0; indistflag = 0;
          double rval, zval;
          hb_draw_uz(KEY, HB_DOM_SNP, it, (uint32_t)i, HB_SL_MAIN, 0, &rval, &zval);
          TAPE(TP_U, rval, 0.0);
          for (int j = 0; j < n_fold; ++j) { acceptProb += stemp[j]; if (rval < acceptProb) { indistflag = j; break; } }
          snptracker[i] = indistflag;
          if (indistflag == 0) gi = 0;
          else {
            v = xx + varei / vara_fold[indistflag];
            TAPE(TP_Z, zval, 0.0);
            gi = rhs / v + sqrt(varei / v) * zval;
            REDRAW_LOOP(i);
            varg += (gi * gi / fold_[indistflag]);
          }
          if (gi != g[i]) {
            gi_ = (g[i] - gi) * n;
            LD_UPDATE(i, gi_);
            g[i] = gi;
          }
        }
        for (int j = 0; j < n_fold; ++j) { double c = 0; for (int i = 0; i < m; ++i) c += (snptracker[i] == j); fold_snp_num[j] = c; }
        NnzSnp = m - (int)fold_snp_num[0];
        varg = (varg + s2varg_ * dfvara_) / chisq_at(HB_DOM_ITER, it, HB_IT_VARG, 0, dfvara_ + NnzSnp);
        for (int j = 0; j < n_fold; ++j) vara_fold[j] = varg * fold_[j];
        fold_snp_num[0] -= nvar0;
        if (!fixpi) {
          for (int j = 0; j < n_fold; ++j) Pi[j] = gamma_at(HB_DOM_ITER, it, HB_IT_PI0 + (uint32_t)j, 0, fold_snp_num[j] + 1);
          const double tot = acc_sum(Pi, n_fold);
          for (int j = 0; j < n_fold; ++j) Pi[j] /= tot;
        }
        break;
    }
    /* :458-468 */
    for (int i = 0; i < m; ++i) tmp[i] = xy[i] - r_hat[i];
    vara_ = (ddot(m, g, tmp) + s2vara_ * dfvara_) / chisq_at(HB_DOM_ITER, it, HB_IT_VARA, 0, n + dfvara_);
    for (int i = 0; i < m; ++i) tmp[i] = xy[i] + r_hat[i];
    vare_ = (yy - ddot(m, g, tmp) + s2vare_ * dfvare_) / chisq_at(HB_DOM_ITER, it, HB_IT_VARE, 0, n + dfvare_);
    if (vare_ < 0) vare_ = vara_ * 0.5;
    if (o->nnz_trace) o->nnz_trace[iter] = NnzSnp;
    if (o->vara_trace) o->vara_trace[iter] = vara_;
    if (o->vare_trace) o->vare_trace[iter] = vare_;
    if (o->varg_trace) o->varg_trace[iter] = varg;
    if (iter >= nburn) {   /* :470-491 */
      if (have_tracker) for (int i = 0; i < m; ++i) if (snptracker[i]) nzrate[i] += 1;
      if (wppai)
        for (int w = 0; w < nw; ++w) {
          int any = 0;
          for (int i = 0; i < m && !any; ++i) if (a->windindx[i] == w + 1 && snptracker[i]) any = 1;
          if (any) wppai[w] += 1;
        }
      nzct++;
    }
    if (iter >= nburn && (iter + 1 - nburn) % thin == 0) {   /* :493-505 */
      if (!fixpi) {
        for (int j = 0; j < n_fold; ++j) pisum[j] += Pi[j];
        if (o->pi_store) for (int j = 0; j < n_fold; ++j) o->pi_store[(size_t)count * n_fold + j] = Pi[j];
      }
      varasum += vara_; varesum += vare_;
      if (o->vara_store) o->vara_store[count] = vara_;
      if (o->vare_store) o->vare_store[count] = vare_;
      for (int i = 0; i < m; ++i) gsum[i] += g[i];
      if (o->alpha_store) memcpy(o->alpha_store + (size_t)count * m, g, 8 * (size_t)m);
      hsqsum += vara_ / (vara_ + vare_);
      if (o->hsq_store) o->hsq_store[count] = vara_ / (vara_ + vare_);
      count++;
    }
    if (count == n_records) { ++iter; break; }
  }
  o->iters_done = iter; o->n_records_done = count; o->nzct = nzct; o->n_used = n;
  o->Vg = varasum / count; o->Ve = varesum / count; o->h2 = hsqsum / count;
  if (o->alpha) for (int i = 0; i < m; ++i) o->alpha[i] = gsum[i] / count;
  if (o->pi) for (int j = 0; j < n_fold; ++j) o->pi[j] = fixpi ? Pi[j] : pisum[j] / count;
  if (fixpi && o->pi_store) for (int c = 0; c < count; ++c) { o->pi_store[(size_t)c * n_fold] = Pi[0]; o->pi_store[(size_t)c * n_fold + 1] = Pi[1]; }
  if (o->nzrate_count) memcpy(o->nzrate_count, nzrate, 8 * (size_t)m);
  if (o->tracker_final) for (int i = 0; i < m; ++i) o->tracker_final[i] = have_tracker ? (int32_t)snptracker[i] : 0;
  if (o->pip)
    for (int i = 0; i < m; ++i) {
      if (!have_tracker) { o->pip[i] = 1.0; continue; }
      double r = nzrate[i] / nzct;
      if (r == 1) r = (nzct - 1) / (double)nzct;
      o->pip[i] = r;
    }
  if (wppai) {
    if (o->wppa_count) memcpy(o->wppa_count, wppai, 8 * (size_t)nw);
    if (o->gwas) for (int w = 0; w < nw; ++w) { double r = wppai[w] / nzct; if (r == 1) r = (nzct - 1) / (double)nzct; o->gwas[w] = r; }
  }
  if (o->r_hat_final) memcpy(o->r_hat_final, r_hat, 8 * (size_t)m);
  free(xy); free(r_hat); free(tmp); free(yyi); free(g); free(xpx); free(vx); free(snptracker); free(nzrate); free(gsum);
  free(vargL); free(ifest); free(wppai); free(varediff);
#undef SS
#undef LD_UPDATE
#undef VAREI
#undef REDRAW_LOOP
  return 0;
}

int hbo_sbayesd(const hbo_sbayes_args* a, hbo_sbayes_out* o) { return sbayes_impl(a, o, 0); }
int hbo_sbayess(const hbo_sbayes_args* a, hbo_sbayes_out* o) { return sbayes_impl(a, o, 1); }
