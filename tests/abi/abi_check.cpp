// abi_check.cpp -- a C++ caller of include/hibayes_b200.h, compiled with g++ and linked against libhibayes_b200.so:
// what the Rcpp shim of INTEGRATION.md would be, without R.  The struct layouts are the compiler's, not a ctypes mirror.
//   abi_check layout        prints sizeof / offsetof of every struct member the Python harness mirrors
//   abi_check run <device>  calls hb_bayes() and hb_sbayesd() on a small deterministic problem and prints the results
//                           as hex doubles (tests/test_abi_compiled.py feeds the same data through the ctypes wrappers)
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/hibayes_b200.h"

#define OFF(T, f) printf(#T "." #f " %zu\n", offsetof(T, f))

static uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

int main(int argc, char** argv) {
  if (argc >= 2 && !strcmp(argv[1], "layout")) {
    printf("sizeof.hb_engine_config %zu\nsizeof.hb_sweep_in %zu\nsizeof.hb_sweep_out %zu\nsizeof.hb_bayes_args %zu\n"
           "sizeof.hb_bayes_out %zu\nsizeof.hb_sbayes_args %zu\nsizeof.hb_sbayes_out %zu\nsizeof.hb_bed_source %zu\n"
           "sizeof.hb_ld_sweep_in %zu\nsizeof.hb_ld_sweep_out %zu\nsizeof.hb_fx_desc %zu\n",
           sizeof(hb_engine_config), sizeof(hb_sweep_in), sizeof(hb_sweep_out), sizeof(hb_bayes_args), sizeof(hb_bayes_out),
           sizeof(hb_sbayes_args), sizeof(hb_sbayes_out), sizeof(hb_bed_source), sizeof(hb_ld_sweep_in), sizeof(hb_ld_sweep_out),
           sizeof(hb_fx_desc));
    OFF(hb_engine_config, device); OFF(hb_engine_config, lag_tiles); OFF(hb_engine_config, seed); OFF(hb_engine_config, world);
    OFF(hb_sweep_in, fold); OFF(hb_sweep_in, vara_fold); OFF(hb_sweep_in, vare); OFF(hb_sweep_in, rnorm2_bound);
    OFF(hb_sweep_out, varg_acc); OFF(hb_sweep_out, var_u); OFF(hb_sweep_out, rounds);
    OFF(hb_bayes_args, y); OFF(hb_bayes_args, x_type); OFF(hb_bayes_args, model); OFF(hb_bayes_args, Pi); OFF(hb_bayes_args, C);
    OFF(hb_bayes_args, Rlev); OFF(hb_bayes_args, niter); OFF(hb_bayes_args, dfvr); OFF(hb_bayes_args, s2ve); OFF(hb_bayes_args, windindx);
    OFF(hb_bayes_args, seed); OFF(hb_bayes_args, ne); OFF(hb_bayes_args, epsl_y_J); OFF(hb_bayes_args, Gi_val); OFF(hb_bayes_args, device);
    OFF(hb_bayes_args, rank); OFF(hb_bayes_args, n_total); OFF(hb_bayes_args, comm_ctx); OFF(hb_bayes_args, allgather_bytes); OFF(hb_bayes_args, nk); OFF(hb_bayes_args, Kival); OFF(hb_bayes_args, Ki);
    OFF(hb_bayes_out, J); OFF(hb_bayes_out, beta); OFF(hb_bayes_out, epsilon); OFF(hb_bayes_out, mu_store); OFF(hb_bayes_out, beta_store);
    OFF(hb_bayes_out, tracker_final); OFF(hb_bayes_out, varg_trace); OFF(hb_bayes_out, n_records_done); OFF(hb_bayes_out, seconds_sweep);
    OFF(hb_bayes_out, rounds_total); OFF(hb_bayes_out, rounds_trace); OFF(hb_bayes_out, vr_store); OFF(hb_bayes_out, epsilon_store);
    OFF(hb_sbayes_args, sumstat); OFF(hb_sbayes_args, model); OFF(hb_sbayes_args, niter); OFF(hb_sbayes_args, vg); OFF(hb_sbayes_args, windindx);
    OFF(hb_sbayes_args, seed); OFF(hb_sbayes_args, device); OFF(hb_sbayes_args, ld_colptr); OFF(hb_sbayes_args, ld_val);
    OFF(hb_sbayes_out, alpha); OFF(hb_sbayes_out, alpha_store); OFF(hb_sbayes_out, r_hat_final); OFF(hb_sbayes_out, n_used);
    OFF(hb_sbayes_out, seconds_sweep); OFF(hb_sbayes_out, columns_total); OFF(hb_sbayes_out, tiles_total);
    OFF(hb_bed_source, len); OFF(hb_bed_source, rows); OFF(hb_bed_source, dominance);
    return 0;
  }
  if (argc >= 2 && !strcmp(argv[1], "run")) {
    const int device = argc >= 3 ? atoi(argv[2]) : 0;
    // ---- hb_bayes(): BayesCpi, n = 500 x m = 700 int8 genotypes, a covariate
    const int n = 500, m = 700;
    uint32_t s = 12345u;
    std::vector<int8_t> X((size_t)n * m);
    for (auto& x : X) x = (int8_t)(lcg(s) % 3);
    std::vector<double> y(n), cov(n);
    for (int i = 0; i < n; ++i) {
      cov[i] = (double)(lcg(s) % 1000) / 500.0 - 1.0;
      double v = 0.7 * cov[i] + (double)(lcg(s) % 2000) / 1000.0 - 1.0;
      for (int j = 0; j < 10; ++j) v += 0.25 * (double)X[(size_t)(j * 37) * n + i];
      y[i] = v;
    }
    const double Pi[2] = {0.9, 0.1};
    hb_bayes_args a;
    memset(&a, 0, sizeof a);
    a.n = n; a.m = m; a.y = y.data(); a.X = X.data(); a.x_type = 1; a.model = "BayesCpi"; a.n_fold = 2; a.Pi = Pi;
    a.nc = 1; a.C = cov.data(); a.niter = 10; a.nburn = 4; a.thin = 2;
    a.dfvr = a.s2vr = a.vg = a.dfvg = a.s2vg = a.ve = a.dfve = a.s2ve = HB_NA;
    a.seed = 2718; a.device = device;
    std::vector<double> alpha(m), pip(m), g(n), e(n), beta(1), pi(2);
    std::vector<int32_t> tracker(m);
    hb_bayes_out o;
    memset(&o, 0, sizeof o);
    o.alpha = alpha.data(); o.pip = pip.data(); o.g = g.data(); o.e = e.data(); o.beta = beta.data(); o.pi = pi.data();
    o.tracker_final = tracker.data();
    int rc = hb_bayes(&a, &o);
    printf("bayes.rc %d\n", rc);
    if (rc) { printf("bayes.error %s\n", hb_last_error()); return 0; }
    long long tsum = 0;
    double asum = 0, esum = 0;
    for (int j = 0; j < m; ++j) { tsum += tracker[j] * (j + 1); asum += alpha[j] * (j % 7 + 1); }
    for (int i = 0; i < n; ++i) esum += e[i] * (i % 5 + 1);
    printf("bayes.Vg %a\nbayes.Ve %a\nbayes.mu %a\nbayes.beta %a\nbayes.pi0 %a\nbayes.tracker_sum %lld\nbayes.alpha_sum %a\nbayes.e_sum %a\n"
           "bayes.records %d\n", o.Vg, o.Ve, o.mu, beta[0], pi[0], tsum, asum, esum, o.n_records_done);
    // ---- hb_sbayesd(): BayesR on a 300-SNP identity-plus-band LD matrix
    const int ms = 300;
    std::vector<double> ld((size_t)ms * ms, 0.0), ss((size_t)ms * 4);
    for (int j = 0; j < ms; ++j) {
      ld[(size_t)j * ms + j] = 0.4 + 0.001 * (j % 50);
      if (j + 1 < ms) { ld[(size_t)j * ms + j + 1] = 0.05; ld[(size_t)(j + 1) * ms + j] = 0.05; }
      ss[j] = 0.2; ss[ms + j] = ((double)(lcg(s) % 2000) / 1000.0 - 1.0) * 0.05; ss[2 * ms + j] = 0.02; ss[3 * ms + j] = 5000.0;
    }
    const double PiR[4] = {0.9, 0.05, 0.03, 0.02}, foldR[4] = {0, 1e-4, 1e-3, 1e-2};
    hb_sbayes_args sa;
    memset(&sa, 0, sizeof sa);
    sa.m = ms; sa.sumstat = ss.data(); sa.ldm = ld.data(); sa.model = "BayesR"; sa.n_fold = 4; sa.Pi = PiR; sa.fold = foldR;
    sa.niter = 12; sa.nburn = 4; sa.thin = 2; sa.vg = sa.dfvg = sa.s2vg = sa.ve = sa.dfve = sa.s2ve = HB_NA; sa.seed = 99; sa.device = device;
    std::vector<double> salpha(ms), spip(ms), spi(4);
    hb_sbayes_out so;
    memset(&so, 0, sizeof so);
    so.alpha = salpha.data(); so.pip = spip.data(); so.pi = spi.data();
    rc = hb_sbayesd(&sa, &so);
    printf("sbayesd.rc %d\n", rc);
    if (rc) { printf("sbayesd.error %s\n", hb_last_error()); return 0; }
    double sasum = 0, psum = 0;
    for (int j = 0; j < ms; ++j) { sasum += salpha[j] * (j % 7 + 1); psum += spip[j]; }
    printf("sbayesd.Vg %a\nsbayesd.Ve %a\nsbayesd.alpha_sum %a\nsbayesd.pip_sum %a\nsbayesd.columns %lld\n", so.Vg, so.Ve, sasum, psum, so.columns_total);
    return 0;
  }
  fprintf(stderr, "usage: abi_check layout | run [device]\n");
  return 2;
}
