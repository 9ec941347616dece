// host_bayes.cpp -- C++ host orchestration of the individual-level Gibbs sampler.
//
// Mirrors Rcpp::List Bayes(...) of the reference (/root/reference/src/Bayes.cpp:60-1094): same
// argument meaning, checks and error texts (:92-117, :293, :325, :356), same priors (:319-375),
// same per-iteration order (intercept :480, covariates :484, environmental random effects :496,
// single-step term :554, SNP sweep :586, variances :819-823, counters :826-845, records
// :848-882, early break :916) and the same outputs (:919-1040).  The SNP sweep and the
// reductions around it run on the GPU through hb_engine_*; there is no CPU fallback.
//
// The Rcpp file that would replace src/Bayes.cpp unpacks its arguments into hb_bayes_args and
// calls hb_bayes(); INTEGRATION.md shows it.
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <chrono>
#include <string>
#include <vector>

#include "../../include/hibayes_b200.h"
#include "hb_rng.h"

int hb_set_error(const char* fmt, ...);

namespace {

inline bool isna(double v) { return v != v; }

// Armadillo arrayops::accumulate / op_var::direct_var (two interleaved accumulators)
double acc_sum(const double* x, int n) {
  double a1 = 0.0, a2 = 0.0;
  int i, j;
  for (i = 0, j = 1; j < n; i += 2, j += 2) { a1 += x[i]; a2 += x[j]; }
  if (i < n) a1 += x[i];
  return a1 + a2;
}
double arma_var(const double* x, int n) {
  if (n < 2) return 0.0;
  const double mean = acc_sum(x, n) / (double)n;
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
double ddot(int n, const double* x, const double* y) {
  double s = 0.0;
  for (int i = 0; i < n; ++i) s += x[i] * y[i];
  return s;
}
void daxpy(int n, double a, const double* x, double* y) {
  for (int i = 0; i < n; ++i) y[i] += a * x[i];
}

struct EngineGuard {
  hb_engine* e = nullptr;
  ~EngineGuard() { hb_engine_destroy(e); }
};

struct FxGuard {
  hb_fx* f = nullptr;
  ~FxGuard() { hb_fx_destroy(f); }
};

double seconds_since(const std::chrono::steady_clock::time_point& t0) {
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

}  // namespace

#define HBCHK(call) do { if ((call) != 0) return 1; } while (0)

extern "C" int hb_bayes(const hb_bayes_args* a, hb_bayes_out* o) {
  if (!a || !o) return hb_set_error("hb_bayes: null argument");
  o->rounds_total = 0; o->tiles_total = 0;
  const int n = a->n, m = a->m;
  const int world = a->world > 1 ? a->world : 1;
  const double ntot = world > 1 ? (double)a->n_total : (double)n;   // individuals over all ranks
  if (world > 1) {
    if (!a->allreduce_sum_f64 || !a->allreduce_sum_i32_dev || !a->allgather_bytes || a->n_total < n)
      return hb_set_error("hb_bayes: world = %d needs n_total and the three collectives", world);
  }
  // sum over ranks, in place (no-op on one rank)
  auto allsum = [&](double* buf, size_t cnt) -> int {
    if (world <= 1) return 0;
    if (a->allreduce_sum_f64(a->comm_ctx, buf, cnt) != 0) return hb_set_error("hb_bayes: all-reduce failed");
    return 0;
  };
  const std::string model = a->model ? a->model : "";
  const hb_key_t KEY = hb_make_key(a->seed);

  // ---- Bayes.cpp:92-117
  for (int i = 0; i < n; ++i) if (isna(a->y[i])) return hb_set_error("NAs are not allowed in y.");
  const int model_index = (model == "BayesRR" ? 1 : (model == "BayesA" ? 2 : (model == "BayesB" || model == "BayesBpi" ? 3 :
                          (model == "BayesC" || model == "BayesCpi" || model == "BSLMM" ? 4 : (model == "BayesL" ? 5 : 6)))));
  bool fixpi = (model == "BayesB" || model == "BayesC");
  const int n_fold = a->n_fold;
  if (n_fold < 2) return hb_set_error("Pi should be a vector.");
  if (acc_sum(a->Pi, n_fold) != 1) return hb_set_error("sum of Pi should be 1.");
  if (a->Pi[0] == 1) return hb_set_error("all markers have no effect size.");
  for (int i = 0; i < n_fold; ++i)
    if (a->Pi[i] < 0 || a->Pi[i] > 1) return hb_set_error("elements of Pi should be at the range of [0, 1]");
  std::vector<double> Pi(a->Pi, a->Pi + n_fold);
  std::vector<double> fold_(std::max(n_fold, 2), 0.0);
  if (a->fold) std::copy(a->fold, a->fold + n_fold, fold_.begin());
  else {
    if (model == "BayesR") return hb_set_error("'fold' should be provided for BayesR model.");
    if (n_fold != 2) return hb_set_error("length of Pi and fold not equals.");
  }
  if (n_fold > HB_MAX_FOLD) return hb_set_error("this build supports at most %d mixture components", HB_MAX_FOLD);

  double vary = arma_var(a->y, n), ymean = acc_sum(a->y, n) / (double)n;
  if (world > 1) {   // the same two-pass variance over all ranks
    double s0 = acc_sum(a->y, n);
    HBCHK(allsum(&s0, 1));
    ymean = s0 / ntot;
    double acc[2] = {0.0, 0.0};
    for (int i = 0; i < n; ++i) { const double t = ymean - a->y[i]; acc[0] += t * t; acc[1] += t; }
    HBCHK(allsum(acc, 2));
    vary = (acc[0] - acc[1] * acc[1] / ntot) / (ntot - 1.0);
  }
  const double h2 = 0.5;
  const int niter = a->niter, nburn = a->nburn, thin = a->thin;
  const int n_records = (niter - nburn) / thin;  // :124

  // ---- covariates :126-147
  const int nc = a->nc;
  std::vector<double> cpc(nc), beta(nc, 0.0), betasum(nc, 0.0);
  for (int i = 0; i < nc; ++i) cpc[i] = ddot(n, a->C + (size_t)i * n, a->C + (size_t)i * n);
  // ---- environmental random effects :149-201
  const int nr = a->nr;
  const double dfr = isna(a->dfvr) ? -1 : a->dfvr;
  const double s2r = isna(a->s2vr) ? 0 : a->s2vr;
  std::vector<double> vrtmp(nr), vrv(nr, 0.0), vrsum(nr, 0.0);
  std::vector<int> R_off(nr + 1, 0);
  int n_levels = 0;
  for (int i = 0; i < nr; ++i) {
    vrtmp[i] = vary * (1 - h2) / (nr + 1);
    n_levels += a->nlev[i];
    R_off[i + 1] = n_levels;
  }
  std::vector<double> estR(n_levels, 0.0), estR_tmp(n_levels, 0.0), r_rhs(n_levels, 0.0), r_cnt(n_levels, 0.0),
      estRsum(n_levels, 0.0), ldiff(n_levels, 0.0);
  for (int i = 0; i < nr; ++i)
    for (int k = 0; k < n; ++k) r_cnt[R_off[i] + a->Rlev[(size_t)i * n + k]] += 1.0;
  // ---- single-step term :235-275
  const int ne = a->ne, qe = (a->Gi_colptr != nullptr || ne) ? a->qe : 0;
  double veps = 0, vepstmp = 0, JtJ = 0, epsl_J_beta = 0, vepssum = 0, Jsum = 0;
  std::vector<double> e_estR(qe, 0.0), e_tmp(qe, 0.0), e_rhs(qe, 0.0), e_cnt(qe, 0.0), e_sum(qe, 0.0);
  const bool have_eps = (a->Gi_colptr != nullptr && qe > 0);   // (a rank of a sharded run may hold none of the ne rows)
  if (ne && !a->Gi_colptr) return hb_set_error("variance-covariance matrix should be provided for epsilon term.");
  if (have_eps) {
    if (!a->epsl_y_J) return hb_set_error("epsl_y_J should be provided for epsilon term.");
    JtJ = ddot(n, a->epsl_y_J, a->epsl_y_J);
    for (int i = 0; i < ne; ++i) {
      if (a->epsl_index[i] < 1 || a->epsl_index[i] > qe) return hb_set_error("epsl_index out of range.");
      e_cnt[a->epsl_index[i] - 1] += 1.0;
    }
  }
  // BSLMM polygenic term :203-233
  const int nk = (a->Ki != nullptr) ? a->nk : 0;
  double vbtmp = 0;
  std::vector<double> k_c1(nk), k_c2(nk);
  if (nk) {
    if (!a->Kival) return hb_set_error("Ki and Kival should be provided together.");
    if (nk != n) return hb_set_error("variance-covariance matrix should be in square.");   // :221 (and :519 needs nk == n)
    if (world > 1) return hb_set_error("hb_bayes: the BSLMM polygenic term is not available with sharded individuals");
  }
  // with the individuals sharded over ranks these set-up sums run over all ranks' rows
  if (world > 1) {
    if (nc) HBCHK(allsum(cpc.data(), nc));
    if (nr) HBCHK(allsum(r_cnt.data(), n_levels));
    if (have_eps) { HBCHK(allsum(&JtJ, 1)); HBCHK(allsum(e_cnt.data(), qe)); }
  }

  int NnzSnp = 0;
  bool have_tracker = false;
  if (model == "BayesRR" || model == "BayesA" || model == "BayesL") {  // :288-292
    NnzSnp = m;
    Pi[0] = 0; Pi[1] = 1;
    fixpi = true;
  } else {
    if (model != "BayesR" && n_fold != 2)
      return hb_set_error("length of Pi should be 2, the first value is the proportion of non-effect markers.");
    have_tracker = true;
  }
  double dfvara_ = isna(a->dfvg) ? 4 : a->dfvg;
  if (dfvara_ <= 2) return hb_set_error("dfvg should not be less than 2.");
  if (niter < nburn) return hb_set_error("Number of total iteration ('niter') shold be larger than burn-in ('nburn').");

  // ---- device engine: load X, column statistics (:310-317), Gram band
  const auto t_setup = std::chrono::steady_clock::now();
  EngineGuard guard;
  hb_engine_config cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.device = a->device; cfg.n = n; cfg.m = m; cfg.lag_tiles = a->lag_tiles;
  // every SNP changes in every sweep of the dense models (RR/A/L): their tiles chain as a full triangular
  // recurrence, which favours short tiles; the mixture models change few SNPs per tile and favour long ones
  cfg.tile_snps = a->tile_snps ? a->tile_snps : ((model_index == 1 || model_index == 2 || model_index == 5) ? 64 : 256);
  cfg.n_slabs = a->n_slabs; cfg.seed = a->seed; cfg.rank = world > 1 ? a->rank : 0; cfg.world = world;
  HBCHK(hb_engine_create(&cfg, &guard.e));
  hb_engine* E = guard.e;
  if (a->x_type == 2) {
    const hb_bed_source* b = (const hb_bed_source*)a->X;
    HBCHK(hb_engine_load_bed(E, b->file, b->len, b->nid, b->rows, b->impt, b->dominance));
  } else if (a->x_type == 1) HBCHK(hb_engine_load_geno_i8(E, (const int8_t*)a->X, (size_t)n));
  else HBCHK(hb_engine_load_geno_f64(E, (const double*)a->X, (size_t)n));
  std::vector<double> xpx(m), sumx(m), vx(m);
  HBCHK(hb_engine_col_stats(E, xpx.data(), sumx.data()));
  HBCHK(allsum(xpx.data(), m));   // exact: integer-valued doubles
  HBCHK(allsum(sumx.data(), m));
  std::vector<uint8_t> active(m);
  int nvar0 = 0;
  for (int i = 0; i < m; ++i) {
    // var(x) from exact integer sums; zero exactly when the column is constant
    const bool constant = (ntot * xpx[i] == sumx[i] * sumx[i]);
    vx[i] = constant ? 0.0 : (xpx[i] - sumx[i] * sumx[i] / ntot) / (ntot - 1.0);
    active[i] = constant ? 0 : 1;
    nvar0 += constant ? 1 : 0;
  }
  const double sumvx = acc_sum(vx.data(), m);
  HBCHK(hb_engine_set_snp_info(E, xpx.data(), active.data()));
  HBCHK(hb_engine_build_gram(E));
  if (world > 1) {
    // the Gram band is a sum over individuals as well (exact int32), and the ranks exchange the IPC handles of
    // their dot accumulators so that the sweep kernels can add to each other over NVLink
    void* gptr = nullptr;
    uint64_t gcnt = 0;
    HBCHK(hb_engine_gram_device(E, &gptr, &gcnt));
    if (a->allreduce_sum_i32_dev(a->comm_ctx, gptr, (size_t)gcnt) != 0) return hb_set_error("hb_bayes: Gram all-reduce failed");
    std::vector<char> mine(64), all(64 * (size_t)world);
    HBCHK(hb_engine_ipc_handle(E, mine.data()));
    if (a->allgather_bytes(a->comm_ctx, mine.data(), all.data(), 64) != 0) return hb_set_error("hb_bayes: all-gather failed");
    HBCHK(hb_engine_set_peers(E, all.data()));
  }
  if (a->windindx) HBCHK(hb_engine_set_windows(E, a->windindx));
  int nw = 0;
  if (a->windindx) for (int i = 0; i < m; ++i) if (a->windindx[i] > nw) nw = a->windindx[i];
  o->seconds_setup = seconds_since(t_setup);

  // ---- priors :319-375
  double vara_ = isna(a->vg) ? ((dfvara_ - 2) / dfvara_) * vary * h2 : a->vg;
  vepstmp = vara_;
  vbtmp = vara_;   // :333
  double vare_ = isna(a->ve) ? vary * (1 - h2) / (nr + 1) : a->ve;
  const double dfvare_ = isna(a->dfve) ? -2 : a->dfve;
  const double s2vara_ = isna(a->s2vg) ? vara_ * (dfvara_ - 2) / dfvara_ : a->s2vg;
  double varg = vara_ / ((1 - Pi[0]) * sumvx);
  const double s2varg_ = s2vara_ / ((1 - Pi[0]) * sumvx);
  const double s2vare_ = isna(a->s2ve) ? 0 : a->s2ve;
  const double R2 = (dfvara_ - 2) / dfvara_;
  double lambda2 = 2 * (1 - R2) / (R2)*sumvx;
  double lambda = sqrt(lambda2);
  const double shape0 = 1.1;
  const double rate0 = (shape0 - 1) / lambda2;
  if (model == "BayesL") {
    std::vector<double> vargL(m, varg);
    HBCHK(hb_engine_set_vargL(E, vargL.data()));
  }
  std::vector<double> fold_snp_num(n_fold, 0.0), vara_fold(n_fold, 0.0);
  for (int j = 0; j < n_fold; ++j) vara_fold[j] = (vara_ / ((1 - Pi[0]) * sumvx)) * fold_[j];

  // ---- state :469-471
  double mu_, mu = ymean;
  std::vector<double> yadj(n);
  for (int i = 0; i < n; ++i) yadj[i] = a->y[i] - mu;
  HBCHK(hb_engine_set_residual(E, yadj.data()));
  // non-SNP effects (covariates :484-494, env. random effects :496-516, single-step J + epsilon :554-584) run on the
  // device on the engine's own residual and u (csrc/effects.cu): the vectors never cross PCIe inside the loop
  const bool side_effects = (nc > 0 || nr > 0 || have_eps || nk > 0);
  FxGuard fxg;
  if (side_effects) {
    hb_fx_desc fd;
    memset(&fd, 0, sizeof fd);
    fd.n = n; fd.nc = nc; fd.C = a->C; fd.nr = nr; fd.Rlev = a->Rlev; fd.nlev = a->nlev;
    if (have_eps) {
      fd.J = a->epsl_y_J; fd.ne = ne; fd.qe = qe; fd.epsl_index = a->epsl_index;
      fd.Gi_colptr = a->Gi_colptr; fd.Gi_rowidx = a->Gi_rowidx; fd.Gi_val = a->Gi_val;
    }
    fd.seed = a->seed;
    fd.nk = nk; fd.Ki = nk ? a->Ki : nullptr;
    HBCHK(hb_fx_create(E, &fd, &fxg.f));
    if (have_eps) HBCHK(hb_fx_eps_set_counts(fxg.f, e_cnt.data()));
  }
  hb_fx* FX = fxg.f;
  double sum_r = acc_sum(yadj.data(), n), sum_r2 = ddot(n, yadj.data(), yadj.data());
  if (world > 1) { double b2[2] = {sum_r, sum_r2}; HBCHK(allsum(b2, 2)); sum_r = b2[0]; sum_r2 = b2[1]; }

  double musum = 0, varasum = 0, varesum = 0, hsqsum = 0;
  std::vector<double> pisum(n_fold, 0.0), gtmp;
  int count = 0, nzct = 0, iter;
  double t_sweep = 0.0;

  for (iter = 0; iter < niter; ++iter) {
    const uint32_t it = (uint32_t)iter;
    // intercept :480-482
    mu_ = -(sum_r / ntot + sqrt(vare_ / ntot) * hb_draw_z(KEY, HB_DOM_ITER, it, HB_IT_MU, 0, 0));
    mu -= mu_;
    double mu_shift = mu_;
    double rnorm2 = sum_r2 + 2.0 * mu_ * sum_r + ntot * mu_ * mu_;
    if (side_effects) {
      HBCHK(hb_fx_axpy(FX, HB_FX_ONES, 0, mu_, 0.0));   // yadj += mu_ (:482)
      mu_shift = 0.0;
      // covariates :484-494
      for (int i = 0; i < nc; ++i) {
        const double oldgi = beta[i], v = cpc[i];
        double rhs;
        HBCHK(hb_fx_dot(FX, HB_FX_COV, i, &rhs));
        HBCHK(allsum(&rhs, 1));
        rhs += v * oldgi;
        const double gi = rhs / v + sqrt(vare_ / v) * hb_draw_z(KEY, HB_DOM_COV, it, (uint32_t)i, 0, 0);
        HBCHK(hb_fx_axpy(FX, HB_FX_COV, i, oldgi - gi, 0.0));
        beta[i] = gi;
      }
      // environmental random effects :496-516
      for (int i = 0; i < nr; ++i) {
        const int off = R_off[i], qr = a->nlev[i];
        HBCHK(hb_fx_level_sums(FX, i, r_rhs.data() + off));                    // Z' yadj
        HBCHK(allsum(r_rhs.data() + off, qr));
        for (int q = 0; q < qr; ++q) r_rhs[off + q] += r_cnt[off + q] * estR[off + q];   // + Z'Z estR
        for (int q = 0; q < qr; ++q) {
          const double l = r_cnt[off + q] + vare_ / vrtmp[i];
          estR_tmp[off + q] = r_rhs[off + q] / l + sqrt(vare_ / l) * hb_draw_z(KEY, HB_DOM_RAND, it, (uint32_t)(off + q), 0, 0);
        }
        for (int q = 0; q < qr; ++q) ldiff[off + q] = estR[off + q] - estR_tmp[off + q];
        HBCHK(hb_fx_level_apply(FX, i, ldiff.data() + off));
        vrtmp[i] = (ddot(qr, estR_tmp.data() + off, estR_tmp.data() + off) + s2r * dfr) /
                   hb_draw_chisq(KEY, HB_DOM_ITER, it, HB_IT_VR0 + (uint32_t)i, 0, qr + dfr);
        vrv[i] = arma_var(estR_tmp.data() + off, qr);
        for (int q = 0; q < qr; ++q) estR[off + q] = estR_tmp[off + q];
      }
      // BSLMM polygenic term :518-552 (block Gibbs sampler on the eigen-decomposition; the dense products run on the device)
      if (nk) {
        double emax = 0.0;
        for (int j = 0; j < nk; ++j) {
          const double ev = (a->Kival[j] * vare_) / (a->Kival[j] + vare_ / vbtmp);   // :531
          k_c2[j] = ev;
          emax = std::max(emax, fabs(ev));
        }
        for (int j = 0; j < nk; ++j) {
          if (!(k_c2[j] >= -1e-06 * emax))                                             // :533
            return hb_set_error("matrix is not positive definite, try to specify parameter 'lambda' with a small value, eg: 0.001 or bigger");
          k_c1[j] = k_c2[j] / vare_;
          k_c2[j] = sqrt(k_c2[j] < 0 ? 0.0 : k_c2[j]);                                 // :534-535
        }
        double quad;
        HBCHK(hb_fx_k_step(FX, iter, k_c1.data(), k_c2.data(), a->Kival, &quad));
        vbtmp = (quad + s2vara_ * dfvara_) / hb_draw_chisq(KEY, HB_DOM_ITER, it, HB_IT_VB, 0, dfvara_ + nk);   // :546-547
      }
      // single-step J + epsilon :554-584 (sparse Gauss-Seidel sampler, solver.cpp:131-140)
      if (have_eps) {
        const double oldgi = epsl_J_beta, v = JtJ;
        double rhs;
        HBCHK(hb_fx_dot(FX, HB_FX_J, 0, &rhs));
        HBCHK(allsum(&rhs, 1));
        rhs += v * oldgi;
        const double gi = rhs / v + sqrt(vare_ / v) * hb_draw_z(KEY, HB_DOM_ITER, it, HB_IT_J, 0, 0);
        HBCHK(hb_fx_axpy(FX, HB_FX_J, 0, oldgi - gi, -(oldgi - gi)));
        epsl_J_beta = gi;
        const double ratio = vare_ / vepstmp;
        if (world > 1) {   // Z'yadj.tail(ne) over all ranks' records (every rank then runs the same sampler)
          if (ne) HBCHK(hb_fx_eps_rhs(FX, e_rhs.data())); else std::fill(e_rhs.begin(), e_rhs.end(), 0.0);
          HBCHK(allsum(e_rhs.data(), qe));
          HBCHK(hb_fx_eps_set_rhs(FX, e_rhs.data()));
        } else {
          HBCHK(hb_fx_eps_rhs(FX, nullptr));
        }
        double quad;
        HBCHK(hb_fx_eps_sample(FX, iter, vare_, ratio, &quad));
        vepstmp = (quad + s2vara_ * dfvara_) / hb_draw_chisq(KEY, HB_DOM_ITER, it, HB_IT_VEPS, 0, dfvara_ + qe);
        veps = vepstmp;
      }
      double rn2;
      HBCHK(hb_fx_dot(FX, HB_FX_RESID, 0, &rn2));
      HBCHK(allsum(&rn2, 1));
      rnorm2 = rn2;
    }

    // SNP sweep :586-816 on the device
    hb_sweep_in in;
    memset(&in, 0, sizeof in);
    in.iter = iter; in.model_index = model_index; in.n_fold = n_fold;
    for (int j = 0; j < n_fold; ++j) { in.fold[j] = fold_[j]; in.logpi[j] = log(Pi[j]); }
    if (model_index == 6) for (int j = 0; j < n_fold; ++j) in.vara_fold[j] = vara_fold[j];
    else in.vara_fold[1] = varg;
    in.vare = vare_; in.dfvara = dfvara_; in.s2varg = s2varg_;
    in.lambda = lambda; in.lambda2 = lambda2;
    in.mu_shift = mu_shift; in.rnorm2_bound = rnorm2;
    hb_sweep_out so;
    HBCHK(hb_engine_sweep(E, &in, &so));
    { float a0, a1, a2; hb_engine_last_sweep_ms(E, &a0, &a1, &a2); t_sweep += 1e-3 * (a0 + a1 + a2); }
    {
      int tsz = 0;
      hb_engine_describe(E, nullptr, nullptr, &tsz, nullptr, nullptr, nullptr);
      o->rounds_total += so.rounds;
      if (o->rounds_trace) o->rounds_trace[iter] = so.rounds;
      if (o->sweep_ms_trace) { float a0, a1, a2; hb_engine_last_sweep_ms(E, &a0, &a1, &a2); o->sweep_ms_trace[iter] = a0 + a1 + a2; }
      o->tiles_total += (m + tsz - 1) / std::max(1, tsz);
    }

    switch (model_index) {
      case 1:
        varg = (so.varg_acc + s2varg_ * dfvara_) / hb_draw_chisq(KEY, HB_DOM_ITER, it, HB_IT_VARG, 0, dfvara_ + m - nvar0);  // :603
        break;
      case 2:
        break;
      case 3:
      case 4:
        fold_snp_num[1] = so.count[1];                          // :666-668, :710-712
        fold_snp_num[0] = m - nvar0 - fold_snp_num[1];
        NnzSnp = (int)fold_snp_num[1];
        if (model_index == 4)
          varg = (so.varg_acc + s2varg_ * dfvara_) / hb_draw_chisq(KEY, HB_DOM_ITER, it, HB_IT_VARG, 0, dfvara_ + NnzSnp);  // :713
        break;
      case 5: {
        const double shape = shape0 + m - nvar0;                 // :738-741
        const double rate = rate0 + so.sum_vargL / 2;
        lambda2 = hb_draw_gamma(KEY, HB_DOM_ITER, it, HB_IT_LAMBDA, 0, shape) * (1 / rate);
        lambda = sqrt(lambda2);
        break;
      }
      case 6:
        for (int j = 0; j < n_fold; ++j) fold_snp_num[j] = so.count[j];
        fold_snp_num[0] += nvar0;                                // snptracker == 0 also for skipped SNPs (:803-805)
        NnzSnp = m - (int)fold_snp_num[0];                       // :806
        varg = (so.varg_acc + s2varg_ * dfvara_) / hb_draw_chisq(KEY, HB_DOM_ITER, it, HB_IT_VARG, 0, dfvara_ + NnzSnp);  // :807
        for (int j = 0; j < n_fold; ++j) vara_fold[j] = varg * fold_[j];
        fold_snp_num[0] -= nvar0;                                // :813
        break;
    }
    if ((model_index == 3 || model_index == 4 || model_index == 6) && !fixpi) {  // rdirichlet_sample, stats.cpp:69-76
      for (int j = 0; j < n_fold; ++j)
        Pi[j] = hb_draw_gamma(KEY, HB_DOM_ITER, it, HB_IT_PI0 + (uint32_t)j, 0, fold_snp_num[j] + 1);
      const double tot = acc_sum(Pi.data(), n_fold);
      for (int j = 0; j < n_fold; ++j) Pi[j] /= tot;
    }
    sum_r = so.sum_r; sum_r2 = so.sum_r2;
    vara_ = so.var_u;                                                                                    // :819
    if (world > 1) {
      // per-iteration scalars are sums over the ranks' rows; var(u) with Armadillo's two accumulators about the
      // global mean
      double b3[3] = {so.sum_r, so.sum_r2, so.sum_u};
      HBCHK(allsum(b3, 3));
      sum_r = b3[0]; sum_r2 = b3[1];
      double acc[2];
      HBCHK(hb_engine_u_centered_sums(E, b3[2] / ntot, &acc[0], &acc[1]));
      HBCHK(allsum(acc, 2));
      vara_ = (acc[0] - acc[1] * acc[1] / ntot) / (ntot - 1.0);
    }
    vare_ = (sum_r2 + s2vare_ * dfvare_) / hb_draw_chisq(KEY, HB_DOM_ITER, it, HB_IT_VARE, 0, ntot + dfvare_);  // :823

    if (o->nnz_trace) o->nnz_trace[iter] = NnzSnp;
    if (o->vara_trace) o->vara_trace[iter] = vara_;
    if (o->vare_trace) o->vare_trace[iter] = vare_;
    if (o->varg_trace) o->varg_trace[iter] = varg;

    if (iter >= nburn) {  // :826-845
      if (have_tracker || a->windindx) HBCHK(hb_engine_accumulate_pip(E));
      nzct++;
    }
    if (iter >= nburn && (iter + 1 - nburn) % thin == 0) {  // :848-882
      musum += mu;
      if (o->mu_store) o->mu_store[count] = mu;
      if (!fixpi) {
        for (int j = 0; j < n_fold; ++j) pisum[j] += Pi[j];
        if (o->pi_store) for (int j = 0; j < n_fold; ++j) o->pi_store[(size_t)count * n_fold + j] = Pi[j];
      }
      varasum += vara_; varesum += vare_;
      if (o->vara_store) o->vara_store[count] = vara_;
      if (o->vare_store) o->vare_store[count] = vare_;
      HBCHK(hb_engine_accumulate_effects(E));
      if (o->alpha_store) HBCHK(hb_engine_get_effects(E, o->alpha_store + (size_t)count * m));
      for (int i = 0; i < nc; ++i) betasum[i] += beta[i];
      if (nc && o->beta_store) memcpy(o->beta_store + (size_t)count * nc, beta.data(), sizeof(double) * nc);
      double vt = vara_ + vare_;
      for (int i = 0; i < nr; ++i) { vt += vrv[i]; vrsum[i] += vrv[i]; }
      for (int q = 0; q < n_levels; ++q) estRsum[q] += estR[q];
      if (nk) HBCHK(hb_fx_k_accumulate(FX));   // :858
      if (have_eps) {
        vepssum += veps; Jsum += epsl_J_beta;
        HBCHK(hb_fx_eps_accumulate(FX));
        if (o->veps_store) o->veps_store[count] = veps;
        if (o->J_store) o->J_store[count] = epsl_J_beta;
        if (o->epsilon_store) HBCHK(hb_fx_eps_get(FX, o->epsilon_store + (size_t)count * qe, nullptr));
      }
      if (nr) {
        if (o->vr_store) for (int i = 0; i < nr; ++i) o->vr_store[(size_t)count * nr + i] = vrv[i];
        if (o->estR_store) memcpy(o->estR_store + (size_t)count * n_levels, estR.data(), sizeof(double) * n_levels);
      }
      hsqsum += vara_ / vt;
      if (o->hsq_store) o->hsq_store[count] = vara_ / vt;
      count++;
    }
    if (a->verbose && a->outfreq > 0 && (iter + 1) % a->outfreq == 0) {  // :884-914 (abridged)
      printf(" %d %d ", iter + 1, NnzSnp);
      for (int j = 0; j < n_fold; ++j) printf("%.4f ", Pi[j]);
      printf("%.4f %.4f %.4f\n", vara_, vare_, vara_ / (vara_ + vare_));
    }
    if (count == n_records) { ++iter; break; }  // :916
  }
  o->iters_done = iter;
  o->n_records_done = count;
  o->nzct = nzct;
  o->seconds_sweep = t_sweep;

  // ---- posterior summaries :919-1040
  const double rc = (double)count;
  o->Vg = varasum / rc; o->Ve = varesum / rc; o->h2 = hsqsum / rc;
  const double Mu = musum / rc;
  o->mu = Mu;
  std::vector<double> alpha(m, 0.0);
  HBCHK(hb_engine_get_effect_sums(E, alpha.data()));
  for (int i = 0; i < m; ++i) alpha[i] /= rc;
  if (nk) {   // :955-964: the polygenic values expressed as SNP effects and added to every stored sample
    std::vector<double> kv(n), ghat(m);
    HBCHK(hb_fx_k_ghat_vec(FX, a->Kival, sumvx, rc, kv.data()));
    HBCHK(hb_engine_xt_vec(E, kv.data(), ghat.data()));
    const double gm = acc_sum(ghat.data(), m) / (double)m;
    for (int i = 0; i < m; ++i) ghat[i] -= gm;
    for (int i = 0; i < m; ++i) alpha[i] += ghat[i];
    if (o->alpha_store)
      for (int c = 0; c < count; ++c)
        for (int i = 0; i < m; ++i) o->alpha_store[(size_t)c * m + i] += ghat[i];
  }
  if (o->alpha) memcpy(o->alpha, alpha.data(), sizeof(double) * m);
  if (o->e) {
    std::vector<double> xg(n);
    HBCHK(hb_engine_predict(E, alpha.data(), xg.data()));
    for (int i = 0; i < n; ++i) o->e[i] = a->y[i] - Mu * 1.0;
    for (int i = 0; i < nc; ++i) {
      const double b = betasum[i] / rc;
      for (int k = 0; k < n; ++k) o->e[k] -= a->C[(size_t)i * n + k] * b;
    }
    for (int i = 0; i < n; ++i) o->e[i] -= xg[i];
  }
  if (o->beta) for (int i = 0; i < nc; ++i) o->beta[i] = betasum[i] / rc;
  if (o->pi) {
    if (!fixpi) for (int j = 0; j < n_fold; ++j) o->pi[j] = pisum[j] / rc;
    else for (int j = 0; j < n_fold; ++j) o->pi[j] = Pi[j];
  }
  if (fixpi && o->pi_store)
    for (int c = 0; c < count; ++c) { o->pi_store[(size_t)c * n_fold] = Pi[0]; o->pi_store[(size_t)c * n_fold + 1] = Pi[1]; }
  if (have_eps) {
    HBCHK(hb_fx_eps_get(FX, nullptr, e_sum.data()));
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
    if (o->estR) memcpy(o->estR, estRsum.data(), sizeof(double) * n_levels);
    if (o->e)
      for (int i = 0; i < nr; ++i)
        for (int k = 0; k < n; ++k) o->e[k] -= estRsum[R_off[i] + a->Rlev[(size_t)i * n + k]];
  }
  if (o->g) HBCHK(hb_engine_get_u(E, o->g));  // the reference returns u as "g" (:1023)
  std::vector<double> nzrate(m, 0.0), wppa(nw, 0.0);
  HBCHK(hb_engine_get_pip_counts(E, nzrate.data(), nw ? wppa.data() : nullptr, nw));
  if (o->nzrate_count) memcpy(o->nzrate_count, nzrate.data(), sizeof(double) * m);
  if (o->tracker_final) {
    if (have_tracker) HBCHK(hb_engine_get_tracker(E, o->tracker_final));
    else memset(o->tracker_final, 0, sizeof(int32_t) * m);  // BayesRR/A/L keep no snptracker (:288-296)
  }
  if (o->pip) {  // :1026-1032
    if (!have_tracker) for (int i = 0; i < m; ++i) o->pip[i] = 1.0;
    else for (int i = 0; i < m; ++i) {
      double r = nzrate[i] / nzct;
      if (r == 1) r = (nzct - 1) / (double)nzct;
      o->pip[i] = r;
    }
  }
  if (nw) {
    if (o->wppa_count) memcpy(o->wppa_count, wppa.data(), sizeof(double) * nw);
    if (o->gwas) for (int w = 0; w < nw; ++w) {
      double r = wppa[w] / nzct;
      if (r == 1) r = (nzct - 1) / (double)nzct;
      o->gwas[w] = r;
    }
  }
  return 0;
}
