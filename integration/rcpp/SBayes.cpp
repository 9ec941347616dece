// Drop-in bodies of hibayes' src/SBayesD.cpp and src/SBayesS.cpp: signatures (SBayesD.cpp:5-24, SBayesS.cpp:21-40) and the
// returned Rcpp::List (SBayesD.cpp:532-578) stay; the chain runs on the GPU through hb_sbayesd() / hb_sbayess().
#include "hb_dropin.h"

using namespace Rcpp;

static Rcpp::List run(bool sparse, arma::mat &sumstat, const arma::mat *ldm_d, const arma::sp_mat *ldm_s, std::string &model, arma::vec &Pi,
                      const int niter, const int nburn, const int thin, const Nullable<arma::vec> &fold, const Nullable<arma::uvec> &windindx,
                      const Nullable<double> &vg, const Nullable<double> &dfvg, const Nullable<double> &s2vg, const Nullable<double> &ve,
                      const Nullable<double> &dfve, const Nullable<double> &s2ve, const int outfreq, const bool verbose){
    const int m = sparse ? ldm_s->n_rows : ldm_d->n_rows, n_fold = Pi.n_elem;
    if((int)sumstat.n_rows != m)  throw Rcpp::exception("Number of SNPs not equals.");             // SBayesD.cpp:29-31
    hb_sbayes_args a;  std::memset(&a, 0, sizeof a);
    hb_sbayes_out  o;  std::memset(&o, 0, sizeof o);
    a.m = m;  a.sumstat = sumstat.memptr();      // m x 4: MAF, BETA, SE, NMISS (R/sbayes.r:209); NA_real_ is a NaN
    std::vector<int32_t> cp, ri;  std::vector<double> gv;
    if(sparse){ hb_dropin::csc(*ldm_s, cp, ri, gv);  a.ld_colptr = cp.data();  a.ld_rowidx = ri.data();  a.ld_val = gv.data(); }
    else a.ldm = ldm_d->memptr();
    a.model = model.c_str();  a.n_fold = n_fold;  a.Pi = Pi.memptr();
    arma::vec fold_v;
    if(fold.isNotNull()){ fold_v = as<arma::vec>(fold);  a.fold = fold_v.memptr();
        if(fold_v.n_elem != Pi.n_elem)  throw Rcpp::exception("length of Pi and fold not equals."); }
    a.niter = niter;  a.nburn = nburn;  a.thin = thin;
    a.vg = hb_dropin::opt(vg);  a.dfvg = hb_dropin::opt(dfvg);  a.s2vg = hb_dropin::opt(s2vg);
    a.ve = hb_dropin::opt(ve);  a.dfve = hb_dropin::opt(dfve);  a.s2ve = hb_dropin::opt(s2ve);
    std::vector<int32_t> wind;
    const int nw = hb_dropin::windows(windindx, wind);
    if(nw)  a.windindx = wind.data();
    a.outfreq = outfreq;  a.verbose = verbose;
    a.seed = hb_dropin::seed_from_r();
    const int n_records = (niter - nburn) / thin;
    arma::vec g = arma::zeros<arma::vec>(m), pi = arma::zeros<arma::vec>(n_fold), nzrate = arma::zeros<arma::vec>(m), wppai = arma::zeros<arma::vec>(nw);
    arma::mat vara_store = arma::zeros<arma::mat>(1, n_records), vare_store = arma::zeros<arma::mat>(1, n_records), hsq_store = arma::zeros<arma::mat>(1, n_records);
    arma::mat pi_store = arma::zeros<arma::mat>(n_fold, n_records), g_store = arma::zeros<arma::mat>(m, n_records);
    o.alpha = g.memptr();  o.pi = pi.memptr();  o.pip = nzrate.memptr();  o.gwas = nw ? wppai.memptr() : NULL;
    o.vara_store = vara_store.memptr();  o.vare_store = vare_store.memptr();  o.hsq_store = hsq_store.memptr();
    o.pi_store = pi_store.memptr();  o.alpha_store = g_store.memptr();
    if((sparse ? hb_sbayess(&a, &o) : hb_sbayesd(&a, &o)) != 0)  throw Rcpp::exception(hb_last_error());
    List results;
    List MCMCsample;
    results["Vg"] = o.Vg;
    results["Ve"] = o.Ve;
    results["h2"] = o.h2;
    MCMCsample["Vg"] = vara_store;
    MCMCsample["Ve"] = vare_store;
    MCMCsample["h2"] = hsq_store;
    results["alpha"] = g;
    MCMCsample["alpha"] = g_store;
    results["pi"] = pi;
    MCMCsample["pi"] = pi_store;
    results["pip"] = nzrate;
    if(nw)  results["gwas"] = wppai;
    results["MCMCsamples"] = MCMCsample;
    return results;
}

// [[Rcpp::export]]
Rcpp::List SBayesD(
    arma::mat sumstat,
    arma::mat ldm,
    std::string model,
    arma::vec Pi,
    const int niter = 50000,
    const int nburn = 20000,
    const int thin = 5,
    const Nullable<arma::vec> fold = R_NilValue,
    const Nullable<arma::uvec> windindx = R_NilValue,
    const Nullable<double> vg = R_NilValue,
    const Nullable<double> dfvg = R_NilValue,
    const Nullable<double> s2vg = R_NilValue,
    const Nullable<double> ve = R_NilValue,
    const Nullable<double> dfve = R_NilValue,
    const Nullable<double> s2ve = R_NilValue,
    const int outfreq = 100,
    const int threads = 0,
    const bool verbose = true
){
    return run(false, sumstat, &ldm, NULL, model, Pi, niter, nburn, thin, fold, windindx, vg, dfvg, s2vg, ve, dfve, s2ve, outfreq, verbose);
}

// [[Rcpp::export]]
Rcpp::List SBayesS(
    arma::mat sumstat,
    arma::sp_mat ldm,
    std::string model,
    arma::vec Pi,
    const int niter = 50000,
    const int nburn = 20000,
    const int thin = 5,
    const Nullable<arma::vec> fold = R_NilValue,
    const Nullable<arma::uvec> windindx = R_NilValue,
    const Nullable<double> vg = R_NilValue,
    const Nullable<double> dfvg = R_NilValue,
    const Nullable<double> s2vg = R_NilValue,
    const Nullable<double> ve = R_NilValue,
    const Nullable<double> dfve = R_NilValue,
    const Nullable<double> s2ve = R_NilValue,
    const int outfreq = 100,
    const int threads = 0,
    const bool verbose = true
){
    return run(true, sumstat, NULL, &ldm, model, Pi, niter, nburn, thin, fold, windindx, vg, dfvg, s2vg, ve, dfve, s2ve, outfreq, verbose);
}
