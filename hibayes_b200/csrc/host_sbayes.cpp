// host_sbayes.cpp -- C++ host orchestration of the dense-LD summary-statistics sampler.
//
// Mirrors Rcpp::List SBayesD(...) of the reference (/root/reference/src/SBayesD.cpp:5-609): same argument
// meaning, checks and error texts (:28-57, :71, :113, :126, :156), priors (:116-170), per-iteration order (SNP
// sweep :253-456, genetic variance :460-461, residual variance :466-468, counters :470-491, records :493-505,
// early break :530) and outputs (:532-578).  The sweep runs on the GPU through hb_ld_engine_*; no CPU fallback.
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/hibayes_b200.h"
#include "hb_rng.h"

int hb_set_error(const char* fmt, ...);

namespace {
inline bool isna(double v) { return v != v; }
double acc_sum(const double* x, int n) {   // arrayops::accumulate
  double a1 = 0.0, a2 = 0.0;
  int i, j;
  for (i = 0, j = 1; j < n; i += 2, j += 2) { a1 += x[i]; a2 += x[j]; }
  if (i < n) a1 += x[i];
  return a1 + a2;
}
struct LdGuard {
  hb_ld_engine* e = nullptr;
  ~LdGuard() { hb_ld_engine_destroy(e); }
};
}  // namespace

#define HBCHK(call) do { if ((call) != 0) return 1; } while (0)

static int sbayes_host(const hb_sbayes_args* a, hb_sbayes_out* o, bool sparse) {
  if (!a || !o) return hb_set_error("hb_sbayes: null argument");
  o->columns_total = 0; o->ld_entries_total = 0; o->ld_bytes_device = 0; o->rounds_total = 0; o->tiles_total = 0;
  if (sparse ? !(a->ld_colptr && a->ld_rowidx && a->ld_val) : !a->ldm) return hb_set_error("hb_sbayes: LD matrix missing");
  const int m = a->m;
  const std::string model = a->model ? a->model : "";
  const hb_key_t KEY = hb_make_key(a->seed);
  const int model_index = (model == "BayesRR" ? 1 : (model == "BayesA" ? 2 : (model == "BayesB" || model == "BayesBpi" ? 3 :
                          (model == "BayesC" || model == "BayesCpi" ? 4 : (model == "BayesL" ? 5 : 6)))));   // :28
  auto SS = [&](int k, int c) { return a->sumstat[(size_t)c * m + k]; };
  int n;   // :33-34
  { double acc = 0.0; int cnt = 0;
    for (int k = 0; k < m; ++k) if (std::isfinite(SS(k, 3))) { acc += SS(k, 3); cnt++; }
    if (!cnt) return hb_set_error("Lack of SE.");
    n = (int)(acc / cnt); }
  bool fixpi = (model == "BayesB" || model == "BayesC");
  const int n_fold = a->n_fold;
  if (n_fold < 2) return hb_set_error("Pi should be a vector.");
  if (acc_sum(a->Pi, n_fold) != 1) return hb_set_error("sum of Pi should be 1.");
  if (a->Pi[0] == 1) return hb_set_error("all markers have no effect size.");
  for (int i = 0; i < n_fold; ++i)
    if (a->Pi[i] < 0 || a->Pi[i] > 1) return hb_set_error("elements of Pi should be at the range of [0, 1]");
  if (n_fold > HB_MAX_FOLD) return hb_set_error("this build supports at most %d mixture components", HB_MAX_FOLD);
  std::vector<double> Pi(a->Pi, a->Pi + n_fold), fold_(n_fold, 0.0);
  if (a->fold) std::copy(a->fold, a->fold + n_fold, fold_.begin());
  else {
    if (model == "BayesR") return hb_set_error("'fold' should be provided for BayesR model.");
    if (n_fold != 2) return hb_set_error("length of Pi and fold not equals.");
  }
  const int niter = a->niter, nburn = a->nburn, thin = a->thin;
  const int n_records = (niter - nburn) / thin;
  int count = 0, nzct = 0, NnzSnp = 0;
  bool have_tracker = false;
  if (model == "BayesRR" || model == "BayesA" || model == "BayesL") {
    NnzSnp = m; Pi[0] = 0; Pi[1] = 1; fixpi = true;
  } else {
    if (model != "BayesR" && n_fold != 2)
      return hb_set_error("length of Pi should be 2, the first value is the proportion of non-effect markers.");
    have_tracker = true;
  }
  std::vector<double> xy(m, 0.0), r_hat(m, 0.0), yyi(m, 0.0), xpx(m), vx(m), g(m, 0.0), gsum(m, 0.0), nzrate(m, 0.0);
  std::vector<uint8_t> ifest(m, 1);
  std::vector<int32_t> tracker(m, 0);
  std::vector<double> varediff(sparse ? m : 0);
  if (sparse) {
    // SBayesS.cpp:109-113, 131-141; the device engine keeps the compressed columns as they are (hb_ld_engine_load_csc)
    for (int i = 0; i < m; ++i) {
      vx[i] = 0.0;
      for (int q = a->ld_colptr[i]; q < a->ld_colptr[i + 1]; ++q)
        if (a->ld_rowidx[q] == i) vx[i] = a->ld_val[q];
      varediff[i] = (m - (double)(a->ld_colptr[i + 1] - a->ld_colptr[i])) / m;
      xpx[i] = vx[i] * n;
    }
  } else {
    for (int i = 0; i < m; ++i) { vx[i] = a->ldm[(size_t)i * m + i]; xpx[i] = vx[i] * n; }   // :92-96
  }
  int count_y = 0, nvar0 = 0;
  for (int k = 0; k < m; ++k) {   // :100-112
    if (isna(SS(k, 1)) || isna(SS(k, 2)) || isna(SS(k, 3))) { ifest[k] = 0; nvar0++; }
    else {
      xy[k] = xpx[k] * SS(k, 1);
      r_hat[k] = xy[k];
      yyi[k] = xpx[k] * (SS(k, 1) * SS(k, 1) + (SS(k, 3) - 2) * SS(k, 2) * SS(k, 2));
      count_y++;
    }
  }
  if (count_y == 0) return hb_set_error("Lack of SE.");
  const double yy = acc_sum(yyi.data(), m) / count_y;
  const double vary = yy / (n - 1);
  const double h2 = 0.5;
  const double dfvara_ = isna(a->dfvg) ? 4 : a->dfvg;
  if (dfvara_ <= 2) return hb_set_error("dfvg should not be less than 2.");
  double vara_ = isna(a->vg) ? ((dfvara_ - 2) / dfvara_) * vary * h2 : a->vg;
  double vare_ = isna(a->ve) ? vary * (1 - h2) : a->ve;
  const double dfvare_ = isna(a->dfve) ? -2 : a->dfve;
  const double s2vara_ = isna(a->s2vg) ? vara_ * (dfvara_ - 2) / dfvara_ : a->s2vg;
  const double sumvx = acc_sum(vx.data(), m);
  double varg = vara_ / ((1 - Pi[0]) * sumvx);
  const double s2varg_ = s2vara_ / ((1 - Pi[0]) * sumvx);
  const double s2vare_ = isna(a->s2ve) ? 0 : a->s2ve;
  if (niter < nburn) return hb_set_error("Number of total iteration ('niter') shold be larger than burn-in ('nburn').");
  const double R2 = (dfvara_ - 2) / dfvara_;
  double lambda2 = 2 * (1 - R2) / (R2)*sumvx;
  double lambda = sqrt(lambda2);
  const double shape0 = 1.1, rate0 = (shape0 - 1) / lambda2;
  std::vector<double> fold_snp_num(n_fold, 0.0), vara_fold(n_fold, 0.0), pisum(n_fold, 0.0);
  for (int j = 0; j < n_fold; ++j) vara_fold[j] = (vara_ / ((1 - Pi[0]) * sumvx)) * fold_[j];
  int nw = 0;
  if (a->windindx) for (int i = 0; i < m; ++i) if (a->windindx[i] > nw) nw = a->windindx[i];
  std::vector<double> wppai(nw, 0.0);

  // ---- device engine: LD matrix and state
  LdGuard guard;
  HBCHK(hb_ld_engine_create(a->device, m, a->seed, &guard.e));
  hb_ld_engine* E = guard.e;
  if (sparse) {
    HBCHK(hb_ld_engine_load_csc(E, a->ld_colptr, a->ld_rowidx, a->ld_val));
    HBCHK(hb_ld_engine_set_sparse_info(E, varediff.data(), vx.data()));
  } else {
    HBCHK(hb_ld_engine_load_dense(E, a->ldm));
  }
  HBCHK(hb_ld_engine_set_state(E, xpx.data(), ifest.data(), xy.data(), r_hat.data()));
  if (model_index == 5) { std::vector<double> vl(m, varg); HBCHK(hb_ld_engine_set_vargL(E, vl.data())); }

  double varasum = 0, varesum = 0, hsqsum = 0, t_sweep = 0;
  int iter;
  for (iter = 0; iter < niter; ++iter) {
    const uint32_t it = (uint32_t)iter;
    hb_ld_sweep_in in;
    memset(&in, 0, sizeof in);
    in.iter = iter; in.model_index = model_index; in.n_fold = n_fold;
    for (int j = 0; j < n_fold; ++j) { in.fold[j] = fold_[j]; in.logpi[j] = log(Pi[j]); }
    if (model_index == 6) for (int j = 0; j < n_fold; ++j) in.vara_fold[j] = vara_fold[j];
    else in.vara_fold[1] = varg;
    in.vare = vare_; in.dfvara = dfvara_; in.s2varg = s2varg_; in.lambda = lambda; in.lambda2 = lambda2; in.nscale = n;
    in.sparse_mode = sparse ? 1 : 0; in.vara = vara_; in.vary = vary;
    hb_ld_sweep_out so;
    HBCHK(hb_ld_engine_sweep(E, &in, &so));
    t_sweep += 1e-3 * so.sweep_ms;
    o->columns_total += so.n_changed; o->ld_entries_total += (long long)so.ld_entries; o->rounds_total += so.rounds;
    o->tiles_total += (m + 255) / 256;
    switch (model_index) {
      case 1:
        varg = (so.varg_acc + s2varg_ * dfvara_) / hb_draw_chisq(KEY, HB_DOM_ITER, it, HB_IT_VARG, 0, dfvara_ + count_y);   // :269
        break;
      case 2:
        break;
      case 3:
      case 4:
        fold_snp_num[1] = so.count[1];                            // :318-320, :357-359
        fold_snp_num[0] = m - nvar0 - fold_snp_num[1];
        NnzSnp = (int)fold_snp_num[1];
        if (model_index == 4)
          varg = (so.varg_acc + s2varg_ * dfvara_) / hb_draw_chisq(KEY, HB_DOM_ITER, it, HB_IT_VARG, 0, dfvara_ + NnzSnp);   // :360
        break;
      case 5: {
        const double shape = shape0 + count_y, rate = rate0 + so.sum_vargL / 2;                    // :385-388
        lambda2 = hb_draw_gamma(KEY, HB_DOM_ITER, it, HB_IT_LAMBDA, 0, shape) * (1 / rate);
        lambda = sqrt(lambda2);
        break;
      }
      case 6:
        for (int j = 0; j < n_fold; ++j) fold_snp_num[j] = so.count[j];
        fold_snp_num[0] += nvar0;                                 // sum(snptracker == 0) counts the skipped SNPs too (:444)
        NnzSnp = m - (int)fold_snp_num[0];                        // :447
        varg = (so.varg_acc + s2varg_ * dfvara_) / hb_draw_chisq(KEY, HB_DOM_ITER, it, HB_IT_VARG, 0, dfvara_ + NnzSnp);   // :448
        for (int j = 0; j < n_fold; ++j) vara_fold[j] = varg * fold_[j];
        fold_snp_num[0] -= nvar0;                                 // :454
        break;
    }
    if ((model_index == 3 || model_index == 4 || model_index == 6) && !fixpi) {
      for (int j = 0; j < n_fold; ++j) Pi[j] = hb_draw_gamma(KEY, HB_DOM_ITER, it, HB_IT_PI0 + (uint32_t)j, 0, fold_snp_num[j] + 1);
      const double tot = acc_sum(Pi.data(), n_fold);
      for (int j = 0; j < n_fold; ++j) Pi[j] /= tot;
    }
    vara_ = (so.g_xy_minus_rhat + s2vara_ * dfvara_) / hb_draw_chisq(KEY, HB_DOM_ITER, it, HB_IT_VARA, 0, n + dfvara_);        // :460-461
    vare_ = (yy - so.g_xy_plus_rhat + s2vare_ * dfvare_) / hb_draw_chisq(KEY, HB_DOM_ITER, it, HB_IT_VARE, 0, n + dfvare_);    // :466-467
    if (vare_ < 0) vare_ = vara_ * 0.5;                                                                                           // :468
    if (o->nnz_trace) o->nnz_trace[iter] = NnzSnp;
    if (o->vara_trace) o->vara_trace[iter] = vara_;
    if (o->vare_trace) o->vare_trace[iter] = vare_;
    if (o->varg_trace) o->varg_trace[iter] = varg;
    const bool record = (iter >= nburn && (iter + 1 - nburn) % thin == 0);
    if (iter >= nburn) {   // :470-491
      if (have_tracker || nw) {
        HBCHK(hb_ld_engine_get(E, record ? g.data() : nullptr, tracker.data(), nullptr));
        if (have_tracker) for (int i = 0; i < m; ++i) if (tracker[i]) nzrate[i] += 1;
        for (int w = 0; w < nw; ++w) {
          bool any = false;
          for (int i = 0; i < m && !any; ++i) any = (a->windindx[i] == w + 1 && tracker[i]);
          if (any) wppai[w] += 1;
        }
      } else if (record) {
        HBCHK(hb_ld_engine_get(E, g.data(), nullptr, nullptr));
      }
      nzct++;
    }
    if (record) {   // :493-505
      if (!fixpi) {
        for (int j = 0; j < n_fold; ++j) pisum[j] += Pi[j];
        if (o->pi_store) for (int j = 0; j < n_fold; ++j) o->pi_store[(size_t)count * n_fold + j] = Pi[j];
      }
      varasum += vara_; varesum += vare_;
      if (o->vara_store) o->vara_store[count] = vara_;
      if (o->vare_store) o->vare_store[count] = vare_;
      for (int i = 0; i < m; ++i) gsum[i] += g[i];
      if (o->alpha_store) memcpy(o->alpha_store + (size_t)count * m, g.data(), 8 * (size_t)m);
      hsqsum += vara_ / (vara_ + vare_);
      if (o->hsq_store) o->hsq_store[count] = vara_ / (vara_ + vare_);
      count++;
    }
    if (a->verbose && a->outfreq > 0 && (iter + 1) % a->outfreq == 0) {
      printf(" %d %d ", iter + 1, NnzSnp);
      for (int j = 0; j < n_fold; ++j) printf("%.4f ", Pi[j]);
      printf("%.4f %.4f %.4f\n", vara_, vare_, vara_ / (vara_ + vare_));
    }
    if (count == n_records) { ++iter; break; }   // :530
  }
  o->iters_done = iter; o->n_records_done = count; o->nzct = nzct; o->n_used = n; o->seconds_sweep = t_sweep;
  { uint64_t lb = 0; hb_ld_engine_describe(E, nullptr, &lb, nullptr); o->ld_bytes_device = (long long)lb; }
  o->Vg = varasum / count; o->Ve = varesum / count; o->h2 = hsqsum / count;   // :535-545
  if (o->alpha) for (int i = 0; i < m; ++i) o->alpha[i] = gsum[i] / count;
  if (o->pi) for (int j = 0; j < n_fold; ++j) o->pi[j] = fixpi ? Pi[j] : pisum[j] / count;
  if (fixpi && o->pi_store) for (int c = 0; c < count; ++c) { o->pi_store[(size_t)c * n_fold] = Pi[0]; o->pi_store[(size_t)c * n_fold + 1] = Pi[1]; }
  if (o->nzrate_count) memcpy(o->nzrate_count, nzrate.data(), 8 * (size_t)m);
  if (o->tracker_final) {
    if (have_tracker) HBCHK(hb_ld_engine_get(E, nullptr, o->tracker_final, nullptr));
    else memset(o->tracker_final, 0, 4 * (size_t)m);
  }
  if (o->pip)   // :560-566
    for (int i = 0; i < m; ++i) {
      if (!have_tracker) { o->pip[i] = 1.0; continue; }
      double r = nzrate[i] / nzct;
      if (r == 1) r = (nzct - 1) / (double)nzct;
      o->pip[i] = r;
    }
  if (nw) {
    if (o->wppa_count) memcpy(o->wppa_count, wppai.data(), 8 * (size_t)nw);
    if (o->gwas) for (int w = 0; w < nw; ++w) { double r = wppai[w] / nzct; if (r == 1) r = (nzct - 1) / (double)nzct; o->gwas[w] = r; }
  }
  if (o->r_hat_final) HBCHK(hb_ld_engine_get(E, nullptr, nullptr, o->r_hat_final));
  return 0;
}

extern "C" int hb_sbayesd(const hb_sbayes_args* a, hb_sbayes_out* o) { return sbayes_host(a, o, false); }
extern "C" int hb_sbayess(const hb_sbayes_args* a, hb_sbayes_out* o) { return sbayes_host(a, o, true); }
