/*
 * ref_entry.cpp -- TEST INFRASTRUCTURE ONLY (oracle/).  C entry points of oracle/_ref/libhibayes_ref.so: the reference's
 * own Bayes() (/root/reference/src/Bayes.cpp:60-1094), SBayesD() (SBayesD.cpp:5-609) and SBayesS() (SBayesS.cpp:21-679),
 * compiled unmodified from where they lie together with stats.cpp and solver.cpp against the stand-in headers of this
 * directory, called with plain buffers (the argument structs of hb_oracle.h) so that tests can put the reference's
 * outputs next to the oracle's.  Also here: what the third-party layer provides at link time -- the random draws
 * (replayed from the oracle's tape, see RcppArmadillo.h), reference-BLAS ddot_/daxpy_, Rcout, LAPACK names that the
 * paths under test never reach.
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <time.h>

#include "../hb_oracle.h"
#include "RcppArmadillo.h"
#include "bigmemory/BigMatrix.h"

using namespace Rcpp;
using namespace arma;

// ---- the reference's entry points (declarations only; definitions are the reference's own files) ------------------------
Rcpp::List Bayes(arma::vec& y, arma::mat& X, std::string model, arma::vec Pi, const Nullable<arma::vec> Kival, const Nullable<arma::mat> Ki,
                 const Nullable<arma::mat> C, const Nullable<CharacterMatrix> R, const Nullable<arma::vec> fold, const int niter, const int nburn,
                 const int thin, const Nullable<arma::vec> epsl_y_J, const Nullable<arma::sp_mat> epsl_Gi, const Nullable<arma::uvec> epsl_index,
                 const Nullable<double> dfvr, const Nullable<double> s2vr, const Nullable<double> vg, const Nullable<double> dfvg,
                 const Nullable<double> s2vg, const Nullable<double> ve, const Nullable<double> dfve, const Nullable<double> s2ve,
                 const Nullable<arma::uvec> windindx, const int outfreq, const int threads, const bool verbose);
Rcpp::List SBayesD(arma::mat sumstat, arma::mat ldm, std::string model, arma::vec Pi, const int niter, const int nburn, const int thin,
                   const Nullable<arma::vec> fold, const Nullable<arma::uvec> windindx, const Nullable<double> vg, const Nullable<double> dfvg,
                   const Nullable<double> s2vg, const Nullable<double> ve, const Nullable<double> dfve, const Nullable<double> s2ve,
                   const int outfreq, const int threads, const bool verbose);
Rcpp::List SBayesS(arma::mat sumstat, arma::sp_mat ldm, std::string model, arma::vec Pi, const int niter, const int nburn, const int thin,
                   const Nullable<arma::vec> fold, const Nullable<arma::uvec> windindx, const Nullable<double> vg, const Nullable<double> dfvg,
                   const Nullable<double> s2vg, const Nullable<double> ve, const Nullable<double> dfve, const Nullable<double> s2ve,
                   const int outfreq, const int threads, const bool verbose);

// tXXmat.cpp:80, :188, :608 and read_bed.cpp:235 (the exported dispatchers on the big.matrix type)
SEXP BigStat(SEXP pBigMat, const int threads);
SEXP tXXmat_Geno(SEXP pBigMat, const Nullable<double> chisq, const int threads, const bool verbose);
SEXP tXXmat_Chr(SEXP pBigMat, const NumericVector chr, const Nullable<double> chisq, const int threads, const bool verbose);
void read_bed(std::string bfile, const SEXP pBigMat, const long maxLine, const bool impt, const bool d, const int threads);

// ---- the tape -------------------------------------------------------------------------------------------------------
static const hbo_tape_entry* g_tape = nullptr;
static size_t g_tape_n = 0, g_tape_pos = 0;
static const char* const KIND[] = {"uniform", "normal", "gamma", "chi-square"};
// time stamps at chosen tape positions (bench.py's reference arm: the first draw of an iteration marks its start)
static uint64_t g_mark_pos[16];
static double g_mark_t[16];
static int g_nmarks = 0, g_next_mark = 0;
static double now_s() { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
extern "C" void hbref_set_time_marks(const uint64_t* pos, int n) {
  g_nmarks = n < 16 ? n : 16; g_next_mark = 0;
  for (int i = 0; i < g_nmarks; ++i) { g_mark_pos[i] = pos[i]; g_mark_t[i] = 0.0; }
}
extern "C" void hbref_get_time_marks(double* t, int n) { for (int i = 0; i < n && i < g_nmarks; ++i) t[i] = g_mark_t[i]; }
static double tape_pop(int kind, double param) {
  if (g_next_mark < g_nmarks && g_tape_pos == g_mark_pos[g_next_mark]) g_mark_t[g_next_mark++] = now_s();
  if (!g_tape) throw Rcpp::exception("libhibayes_ref: a random draw was requested but no tape is set");
  char buf[256];
  if (g_tape_pos >= g_tape_n) {
    snprintf(buf, sizeof buf, "tape exhausted: the reference asks for draw #%zu (%s) but the oracle consumed only %zu", g_tape_pos, KIND[kind], g_tape_n);
    throw Rcpp::exception(buf);
  }
  const hbo_tape_entry& e = g_tape[g_tape_pos];
  if (e.kind != kind) {
    snprintf(buf, sizeof buf, "tape mismatch at draw #%zu: the reference asks for %s, the oracle drew %s", g_tape_pos, KIND[kind], KIND[e.kind & 3]);
    throw Rcpp::exception(buf);
  }
  if (kind >= 2 && memcmp(&e.param, &param, sizeof(double)) != 0) {
    snprintf(buf, sizeof buf, "tape mismatch at draw #%zu (%s): the reference's shape %.17g, the oracle's %.17g", g_tape_pos, KIND[kind], param, e.param);
    throw Rcpp::exception(buf);
  }
  ++g_tape_pos;
  return e.value;
}
extern "C" double unif_rand(void) { return tape_pop(0, 0.0); }
extern "C" double norm_rand(void) { return tape_pop(1, 0.0); }
namespace R {
double rgamma(double shape, double scale) { return tape_pop(2, shape) * scale; }
double rchisq(double df) { return tape_pop(3, df); }
double rnorm(double mu, double sd) { return mu + sd * tape_pop(1, 0.0); }
double runif(double a, double b) { return a + (b - a) * tape_pop(0, 0.0); }
static double not_on_the_path(const char* f) { throw Rcpp::exception((std::string("libhibayes_ref: ") + f + " is not on the paths under test").c_str()); }
double rbeta(double, double) { return not_on_the_path("rbeta"); }
double rt(double) { return not_on_the_path("rt"); }
double rcauchy(double, double) { return not_on_the_path("rcauchy"); }
double rexp(double) { return not_on_the_path("rexp"); }
}  // namespace R
namespace arma { double mini_arma_randn() { return tape_pop(1, 0.0); } }

// ---- reference BLAS level 1 (netlib ddot/daxpy with unit stride: one accumulator, ascending) ---------------------------
// hbref_use_blas(path, prefix): forward both to a real BLAS (bench.py's reference arm: the bundled multi-threaded OpenBLAS,
// what an R installation linked against OpenBLAS runs); the default loops are what the bit-for-bit pin tests use.
typedef double (*ddot_fn)(const int*, const double*, const int*, const double*, const int*);
typedef void (*daxpy_fn)(const int*, const double*, const double*, const int*, double*, const int*);
static ddot_fn g_ddot = nullptr;
static daxpy_fn g_daxpy = nullptr;
extern "C" const char* hbref_last_error(void);
static char g_err[600];
extern "C" int hbref_use_blas(const char* path, const char* prefix) {
  g_ddot = nullptr; g_daxpy = nullptr;
  if (!path) return 0;
  void* h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
  if (!h) { snprintf(g_err, sizeof g_err, "hbref_use_blas: %s", dlerror()); return 1; }
  const std::string pre(prefix ? prefix : "");
  g_ddot = (ddot_fn)dlsym(h, (pre + "ddot_").c_str());
  g_daxpy = (daxpy_fn)dlsym(h, (pre + "daxpy_").c_str());
  if (!g_ddot || !g_daxpy) { g_ddot = nullptr; g_daxpy = nullptr; snprintf(g_err, sizeof g_err, "hbref_use_blas: %sddot_ / %sdaxpy_ not found", pre.c_str(), pre.c_str()); return 1; }
  return 0;
}
extern "C" double ddot_(const int* n, const double* x, const int* incx, const double* y, const int* incy) {
  if (g_ddot) return g_ddot(n, x, incx, y, incy);
  double s = 0.0;
  const int ix = *incx, iy = *incy;
  for (int i = 0; i < *n; ++i) s += x[(size_t)i * ix] * y[(size_t)i * iy];
  return s;
}
extern "C" void daxpy_(const int* n, const double* a, const double* x, const int* incx, double* y, const int* incy) {
  if (g_daxpy) { g_daxpy(n, a, x, incx, y, incy); return; }
  const int ix = *incx, iy = *incy;
  const double al = *a;
  if (al == 0.0) return;   // (netlib: quick return)
  for (int i = 0; i < *n; ++i) y[(size_t)i * iy] += al * x[(size_t)i * ix];
}
#define HB_NOLAPACK(name) throw Rcpp::exception("libhibayes_ref: LAPACK " name " is not on the paths under test")
extern "C" void dpotrf_(const char*, const int*, double*, const int*, int*) { HB_NOLAPACK("dpotrf"); }
extern "C" void dpotri_(const char*, const int*, double*, const int*, int*) { HB_NOLAPACK("dpotri"); }
extern "C" void dgetrf_(const int*, const int*, double*, const int*, int*, int*) { HB_NOLAPACK("dgetrf"); }
extern "C" void dgetri_(const int*, double*, const int*, const int*, double*, const int*, int*) { HB_NOLAPACK("dgetri"); }
extern "C" double dlange_(const char*, const int*, const int*, const double*, const int*, double*) { HB_NOLAPACK("dlange"); }
extern "C" void dgecon_(const char*, const int*, const double*, const int*, const double*, double*, double*, int*, int*) { HB_NOLAPACK("dgecon"); }
extern "C" void dsyevd_(const char*, const char*, const int*, double*, const int*, double*, double*, const int*, int*, const int*, int*) { HB_NOLAPACK("dsyevd"); }

// ---- printing ---------------------------------------------------------------------------------------------------------
namespace Rcpp {
Rostream::Rostream(bool err) : std::ostream(nullptr) { if (getenv("HB_REF_VERBOSE")) rdbuf(err ? std::cerr.rdbuf() : std::cout.rdbuf()); else rdbuf(&nb_); }
Rostream Rcout(false);
Rostream Rcerr(true);
}  // namespace Rcpp
extern "C" void Rprintf(const char* fmt, ...) { if (!getenv("HB_REF_VERBOSE")) return; va_list ap; va_start(ap, fmt); vprintf(fmt, ap); va_end(ap); }
extern "C" void REprintf(const char* fmt, ...) { if (!getenv("HB_REF_VERBOSE")) return; va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); }

// ---- plain-buffer wrappers ----------------------------------------------------------------------------------------------
extern "C" const char* hbref_last_error(void) { return g_err; }

static Nullable<double> opt(double v) { return v == v ? Nullable<double>(v) : Nullable<double>(); }
static void put(double* dst, const mat& v, size_t n) { if (dst) { if (v.n_elem != n) throw Rcpp::exception("libhibayes_ref: unexpected output length"); memcpy(dst, v.mem.data(), n * sizeof(double)); } }
static uvec to_uvec(const int32_t* p, size_t n) { uvec o(n); for (size_t i = 0; i < n; ++i) o.mem[i] = (uword)p[i]; return o; }
static vec to_vec(const double* p, size_t n) { vec o(n); memcpy(o.mem.data(), p, n * sizeof(double)); return o; }
static mat to_mat(const double* p, size_t r, size_t c) { mat o(r, c); memcpy(o.mem.data(), p, r * c * sizeof(double)); return o; }

struct TapeScope {
  TapeScope(const hbo_tape_entry* t, size_t n) { g_tape = t; g_tape_n = n; g_tape_pos = 0; }
  ~TapeScope() { g_tape = nullptr; }
};

extern "C" int hbref_bayes(const hbo_bayes_args* a, hbo_bayes_out* o, const hbo_tape_entry* tape, size_t ntape, size_t* consumed) {
  g_err[0] = 0;
  try {
    TapeScope ts(tape, ntape);
    const size_t n = a->n, m = a->m;
    vec y = to_vec(a->y, n);
    mat X(n, m);
    if (a->x_is_int8) { const int8_t* p = (const int8_t*)a->X; for (size_t i = 0; i < n * m; ++i) X.mem[i] = (double)p[i]; }
    else memcpy(X.mem.data(), a->X, n * m * sizeof(double));
    vec Pi = to_vec(a->Pi, a->n_fold);
    Nullable<vec> fold; if (a->fold) fold = Nullable<vec>(to_vec(a->fold, a->n_fold));
    Nullable<vec> Kival; Nullable<mat> Ki;
    if (a->nk) { Kival = Nullable<vec>(to_vec(a->Kival, a->nk)); Ki = Nullable<mat>(to_mat(a->Ki, n, a->nk)); }
    Nullable<mat> C; if (a->nc) C = Nullable<mat>(to_mat(a->C, n, a->nc));
    Nullable<CharacterMatrix> R;
    if (a->nr) {   // level codes -> labels whose lexical order is the order of the codes (makeZ() sorts the labels, Bayes.cpp:31-33)
      CharacterMatrix Rm((int)n, a->nr);
      char lab[32];
      for (int j = 0; j < a->nr; ++j) for (size_t i = 0; i < n; ++i) { snprintf(lab, sizeof lab, "L%09d", a->Rlev[(size_t)j * n + i]); Rm((int)i, j) = lab; }
      R = Nullable<CharacterMatrix>(Rm);
    }
    Nullable<vec> eyJ; Nullable<sp_mat> eGi; Nullable<uvec> eidx;
    if (a->ne) {
      eyJ = Nullable<vec>(to_vec(a->epsl_y_J, n));
      std::vector<long long> cp(a->qe + 1);
      for (int q = 0; q <= a->qe; ++q) cp[q] = a->Gi_colptr[q];
      eGi = Nullable<sp_mat>(sp_mat::from_csc(a->qe, a->qe, cp.data(), a->Gi_rowidx, a->Gi_val));
      eidx = Nullable<uvec>(to_uvec(a->epsl_index, a->ne));
    }
    Nullable<uvec> wind; if (a->windindx) wind = Nullable<uvec>(to_uvec(a->windindx, m));
    List res = Bayes(y, X, std::string(a->model), Pi, Kival, Ki, C, R, fold, a->niter, a->nburn, a->thin, eyJ, eGi, eidx, opt(a->dfvr), opt(a->s2vr),
                     opt(a->vg), opt(a->dfvg), opt(a->s2vg), opt(a->ve), opt(a->dfve), opt(a->s2ve), wind, 100, 1, getenv("HB_REF_VERBOSE") != nullptr);
    if (consumed) *consumed = g_tape_pos;
    o->Vg = res.get<double>("Vg"); o->Ve = res.get<double>("Ve"); o->h2 = res.get<double>("h2"); o->mu = res.get<double>("mu");
    put(o->alpha, res.get<vec>("alpha"), m);
    put(o->pi, res.get<vec>("pi"), a->n_fold);
    put(o->pip, res.get<vec>("pip"), m);
    put(o->g, res.get<vec>("g"), n);
    put(o->e, res.get<vec>("e"), n);
    if (a->nc) put(o->beta, res.get<vec>("beta"), a->nc);
    if (a->windindx && o->gwas) { const vec& w = res.get<vec>("gwas"); memcpy(o->gwas, w.mem.data(), w.n_elem * sizeof(double)); }
    const List& S = res.get<List>("MCMCsamples");
    const mat& vs = S.get<mat>("Vg");
    const size_t nrec = vs.n_elem;
    o->n_records_done = (int)nrec;
    put(o->vara_store, vs, nrec);
    put(o->vare_store, S.get<mat>("Ve"), nrec);
    put(o->hsq_store, S.get<mat>("h2"), nrec);
    put(o->mu_store, S.get<mat>("mu"), nrec);
    put(o->pi_store, S.get<mat>("pi"), a->n_fold * nrec);
    put(o->alpha_store, S.get<mat>("alpha"), m * nrec);
    if (a->nc) put(o->beta_store, S.get<mat>("beta"), a->nc * nrec);
    if (a->nr) {
      size_t nlev = 0; for (int j = 0; j < a->nr; ++j) nlev += a->nlev[j];
      put(o->vr, res.get<vec>("Vr"), a->nr);
      List r = res.get<List>("r");
      vec est = r[1];
      put(o->estR, est, nlev);
      put(o->vr_store, S.get<mat>("Vr"), a->nr * nrec);
      put(o->estR_store, S.get<mat>("r"), nlev * nrec);
    }
    if (a->ne) {
      o->Veps = res.get<double>("Veps"); o->J = res.get<double>("J");
      put(o->epsilon, res.get<vec>("epsilon"), a->qe);
      put(o->veps_store, S.get<mat>("Veps"), nrec);
      put(o->J_store, S.get<mat>("J"), nrec);
      put(o->epsilon_store, S.get<mat>("epsilon"), (size_t)a->qe * nrec);
    }
    return 0;
  } catch (const std::exception& e) {
    snprintf(g_err, sizeof g_err, "%s", e.what());
    if (consumed) *consumed = g_tape_pos;
    return 1;
  }
}

static int run_sbayes(int sparse, const hbo_sbayes_args* a, hbo_sbayes_out* o, const hbo_tape_entry* tape, size_t ntape, size_t* consumed) {
  g_err[0] = 0;
  try {
    TapeScope ts(tape, ntape);
    const size_t m = a->m;
    mat sumstat = to_mat(a->sumstat, m, 4);
    vec Pi = to_vec(a->Pi, a->n_fold);
    Nullable<vec> fold; if (a->fold) fold = Nullable<vec>(to_vec(a->fold, a->n_fold));
    Nullable<uvec> wind; if (a->windindx) wind = Nullable<uvec>(to_uvec(a->windindx, m));
    List res;
    if (sparse) {
      std::vector<long long> cp(m + 1);
      for (size_t q = 0; q <= m; ++q) cp[q] = a->ld_colptr[q];
      // (stored zeros are kept: an arma::sp_mat built from a dgCMatrix keeps what the caller stored only if non-zero)
      sp_mat ldm = sp_mat::from_csc(m, m, cp.data(), a->ld_rowidx, a->ld_val);
      res = SBayesS(sumstat, ldm, std::string(a->model), Pi, a->niter, a->nburn, a->thin, fold, wind, opt(a->vg), opt(a->dfvg), opt(a->s2vg),
                    opt(a->ve), opt(a->dfve), opt(a->s2ve), 100, 1, getenv("HB_REF_VERBOSE") != nullptr);
    } else {
      mat ldm = to_mat(a->ldm, m, m);
      res = SBayesD(sumstat, ldm, std::string(a->model), Pi, a->niter, a->nburn, a->thin, fold, wind, opt(a->vg), opt(a->dfvg), opt(a->s2vg),
                    opt(a->ve), opt(a->dfve), opt(a->s2ve), 100, 1, getenv("HB_REF_VERBOSE") != nullptr);
    }
    if (consumed) *consumed = g_tape_pos;
    o->Vg = res.get<double>("Vg"); o->Ve = res.get<double>("Ve"); o->h2 = res.get<double>("h2");
    put(o->alpha, res.get<vec>("alpha"), m);
    put(o->pi, res.get<vec>("pi"), a->n_fold);
    put(o->pip, res.get<vec>("pip"), m);
    if (a->windindx && o->gwas) { const vec& w = res.get<vec>("gwas"); memcpy(o->gwas, w.mem.data(), w.n_elem * sizeof(double)); }
    const List& S = res.get<List>("MCMCsamples");
    const mat& vs = S.get<mat>("Vg");
    const size_t nrec = vs.n_elem;
    o->n_records_done = (int)nrec;
    put(o->vara_store, vs, nrec);
    put(o->vare_store, S.get<mat>("Ve"), nrec);
    put(o->hsq_store, S.get<mat>("h2"), nrec);
    put(o->pi_store, S.get<mat>("pi"), a->n_fold * nrec);
    put(o->alpha_store, S.get<mat>("alpha"), m * nrec);
    return 0;
  } catch (const std::exception& e) {
    snprintf(g_err, sizeof g_err, "%s", e.what());
    if (consumed) *consumed = g_tape_pos;
    return 1;
  }
}
extern "C" int hbref_sbayesd(const hbo_sbayes_args* a, hbo_sbayes_out* o, const hbo_tape_entry* tape, size_t ntape, size_t* consumed) { return run_sbayes(0, a, o, tape, ntape, consumed); }
extern "C" int hbref_sbayess(const hbo_sbayes_args* a, hbo_sbayes_out* o, const hbo_tape_entry* tape, size_t ntape, size_t* consumed) { return run_sbayes(1, a, o, tape, ntape, consumed); }

// ---- LD builder and .bed decoder (no random numbers) --------------------------------------------------------------------------
// X: nid x m int8 column-major (a big.matrix of type char, what read_bed() fills and ldmat() hands over, R/ldm.r).
extern "C" int hbref_bigstat(const int8_t* X, int nid, int m, double* mean, double* sum, double* xx) {
  g_err[0] = 0;
  try {
    BigMatrix bm(const_cast<int8_t*>(X), nid, m, 1);
    List st(BigStat(XPtr<BigMatrix>(&bm), 1));
    const NumericVector a = st[0], b = st[1], c = st[2];
    for (int j = 0; j < m; ++j) { mean[j] = a[j]; sum[j] = b[j]; xx[j] = c[j]; }
    return 0;
  } catch (const std::exception& e) { snprintf(g_err, sizeof g_err, "%s", e.what()); return 1; }
}
// out: m x m column-major, zero where the reference's arma::sp_mat stores nothing.  chr NULL -> tXXmat_Geno, else tXXmat_Chr.
extern "C" int hbref_txxmat(const int8_t* X, int nid, int m, const int32_t* chr, int has_chisq, double chisq, double* out, long long* stored) {
  g_err[0] = 0;
  try {
    BigMatrix bm(const_cast<int8_t*>(X), nid, m, 1);
    XPtr<BigMatrix> xp(&bm);
    Nullable<double> cs; if (has_chisq) cs = Nullable<double>(chisq);
    SEXP res;
    if (chr) { NumericVector c(m); for (int j = 0; j < m; ++j) c[j] = chr[j]; res = tXXmat_Chr(xp, c, cs, 1, false); }
    else res = tXXmat_Geno(xp, cs, 1, false);
    memset(out, 0, sizeof(double) * (size_t)m * m);
    const bool is_sparse = chr != nullptr || (has_chisq && chisq > 0);   // (tXXmat.cpp:117-120, :520-523: which branch returns what)
    if (is_sparse) {
      const sp_mat& L = res->payload.get<sp_mat>();
      for (uword j = 0; j < L.n_cols; ++j) for (uword p = L.col_ptrs[j]; p < L.col_ptrs[j + 1]; ++p) out[(size_t)j * m + L.row_indices[p]] = L.values[p];
      if (stored) *stored = (long long)L.n_nonzero;
    } else {
      const mat& L = res->payload.get<mat>();
      memcpy(out, L.mem.data(), sizeof(double) * (size_t)m * m);
      if (stored) *stored = (long long)m * m;
    }
    return 0;
  } catch (const std::exception& e) { snprintf(g_err, sizeof g_err, "%s", e.what()); return 1; }
}
// read_bed<char>() on a file: out nid x m int8 column-major
extern "C" int hbref_read_bed(const char* bfile, int nid, int m, long maxLine, int impute, int dominance, int8_t* out) {
  g_err[0] = 0;
  try {
    BigMatrix bm(out, nid, m, 1);
    read_bed(std::string(bfile), XPtr<BigMatrix>(&bm), maxLine, impute != 0, dominance != 0, 1);
    return 0;
  } catch (const std::exception& e) { snprintf(g_err, sizeof g_err, "%s", e.what()); return 1; }
}
