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
    arma::vec alpha_v = arma::zeros<arma::vec>(m), pi = arma::zeros<arma::vec>(n_fold), pip_v = arma::zeros<arma::vec>(m), gwas_v = arma::zeros<arma::vec>(nw);
    arma::mat vg_rec = arma::zeros<arma::mat>(1, n_records), ve_rec = arma::zeros<arma::mat>(1, n_records), h2_rec = arma::zeros<arma::mat>(1, n_records);
    arma::mat pi_rec = arma::zeros<arma::mat>(n_fold, n_records), alpha_rec = arma::zeros<arma::mat>(m, n_records);
    o.alpha = alpha_v.memptr();  o.pi = pi.memptr();  o.pip = pip_v.memptr();  o.gwas = nw ? gwas_v.memptr() : NULL;
    o.vara_store = vg_rec.memptr();  o.vare_store = ve_rec.memptr();  o.hsq_store = h2_rec.memptr();
    o.pi_store = pi_rec.memptr();  o.alpha_store = alpha_rec.memptr();
    if((sparse ? hb_sbayess(&a, &o) : hb_sbayesd(&a, &o)) != 0)  throw Rcpp::exception(hb_last_error());
    List out;
    List samples;
    out["Vg"] = o.Vg;
    out["Ve"] = o.Ve;
    out["h2"] = o.h2;
    samples["Vg"] = vg_rec;
    samples["Ve"] = ve_rec;
    samples["h2"] = h2_rec;
    out["alpha"] = alpha_v;
    samples["alpha"] = alpha_rec;
    out["pi"] = pi;
    samples["pi"] = pi_rec;
    out["pip"] = pip_v;
    if(nw)  out["gwas"] = gwas_v;
    out["MCMCsamples"] = samples;
    return out;
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
