// Drop-in body of hibayes' src/Bayes.cpp: the exported signature (Bayes.cpp:60-88), the named Rcpp::List it returns
// (:919-1040) and its error texts stay; the chain runs on the GPU through hb_bayes() (include/hibayes_b200.h).
#include "hb_dropin.h"

using namespace Rcpp;

// [[Rcpp::export]]
Rcpp::List Bayes(
    arma::vec &y,
    arma::mat &X,
    std::string model,
    arma::vec Pi,
    const Nullable<arma::vec> Kival = R_NilValue,
    const Nullable<arma::mat> Ki = R_NilValue,
    const Nullable<arma::mat> C = R_NilValue,
    const Nullable<CharacterMatrix> R = R_NilValue,
    const Nullable<arma::vec> fold = R_NilValue,
    const int niter = 50000,
    const int nburn = 20000,
    const int thin = 5,
    const Nullable<arma::vec> epsl_y_J = R_NilValue,
    const Nullable<arma::sp_mat> epsl_Gi = R_NilValue,
    const Nullable<arma::uvec> epsl_index = R_NilValue,
    const Nullable<double> dfvr = R_NilValue,
    const Nullable<double> s2vr = R_NilValue,
    const Nullable<double> vg = R_NilValue,
    const Nullable<double> dfvg = R_NilValue,
    const Nullable<double> s2vg = R_NilValue,
    const Nullable<double> ve = R_NilValue,
    const Nullable<double> dfve = R_NilValue,
    const Nullable<double> s2ve = R_NilValue,
    const Nullable<arma::uvec> windindx = R_NilValue,
    const int outfreq = 100,
    const int threads = 0,
    const bool verbose = true
){
    if(y.n_elem != X.n_rows)  throw Rcpp::exception("Number of individuals not equals.");          // Bayes.cpp:96
    hb_bayes_args a;  std::memset(&a, 0, sizeof a);
    hb_bayes_out  o;  std::memset(&o, 0, sizeof o);
    const int n = X.n_rows, m = X.n_cols, n_fold = Pi.n_elem;
    a.n = n;  a.m = m;  a.y = y.memptr();  a.X = X.memptr();  a.x_type = 0;      // an R numeric matrix: fp64, column-major
    a.model = model.c_str();  a.n_fold = n_fold;  a.Pi = Pi.memptr();
    arma::vec fold_v;
    if(fold.isNotNull()){ fold_v = as<arma::vec>(fold);  a.fold = fold_v.memptr();
        if(fold_v.n_elem != Pi.n_elem)  throw Rcpp::exception("length of Pi and fold not equals."); }   // :115
    arma::mat C_m;
    if(C.isNotNull()){ C_m = as<arma::mat>(C);
        if(C_m.n_rows != X.n_rows)  throw Rcpp::exception("Number of individuals does not match for covariates.");   // :133
        a.nc = C_m.n_cols;  a.C = C_m.memptr(); }
    // environmental random effects: labels -> 0-based codes in the order of the sorted labels (makeZ(), Bayes.cpp:24-58)
    std::vector<int32_t> lev, nlev;
    std::vector<std::string> r_levels;
    if(R.isNotNull()){
        CharacterMatrix R_ = as<CharacterMatrix>(R);
        if(R_.nrow() != n)  throw Rcpp::exception("Number of individuals does not match for environmental random effects.");   // :177
        a.nr = R_.ncol();
        lev.resize((size_t)n * a.nr);
        for(int j = 0; j < a.nr; j++){
            std::map<std::string, int> code;
            for(int i = 0; i < n; i++)  code[as<std::string>(R_(i, j))] = 0;
            if((int)code.size() == n)  throw Rcpp::exception("number of class of environmental random effects should be less than population size.");
            if(code.size() == 1)  throw Rcpp::exception("number of class of environmental random effects should be bigger than 1.");
            int k = 0;
            for(std::map<std::string, int>::iterator it = code.begin(); it != code.end(); ++it){ it->second = k++;  r_levels.push_back(it->first); }
            for(int i = 0; i < n; i++)  lev[(size_t)j * n + i] = code[as<std::string>(R_(i, j))];
            nlev.push_back(k);
        }
        a.Rlev = lev.data();  a.nlev = nlev.data();
    }
    int n_levels = 0;  for(size_t j = 0; j < nlev.size(); j++)  n_levels += nlev[j];
    // BSLMM: eigenvectors / eigenvalues of the relationship matrix (Bayes.cpp:218-233)
    arma::mat K_m;  arma::vec Kval_v;
    if(Ki.isNotNull()){ K_m = as<arma::mat>(Ki);  Kval_v = as<arma::vec>(Kival);
        if(K_m.n_cols != K_m.n_rows)  throw Rcpp::exception("variance-covariance matrix should be in square.");      // :221
        a.nk = K_m.n_cols;  a.Ki = K_m.memptr();  a.Kival = Kval_v.memptr(); }
    a.niter = niter;  a.nburn = nburn;  a.thin = thin;
    a.dfvr = hb_dropin::opt(dfvr);  a.s2vr = hb_dropin::opt(s2vr);  a.vg = hb_dropin::opt(vg);  a.dfvg = hb_dropin::opt(dfvg);
    a.s2vg = hb_dropin::opt(s2vg);  a.ve = hb_dropin::opt(ve);  a.dfve = hb_dropin::opt(dfve);  a.s2ve = hb_dropin::opt(s2ve);
    std::vector<int32_t> wind;
    const int nw = hb_dropin::windows(windindx, wind);
    if(nw)  a.windindx = wind.data();
    a.outfreq = outfreq;  a.verbose = verbose;
    // single-step term (Bayes.cpp:235-275): J covariate, sparse Gi, 1-based epsl_index
    arma::vec J_v;  std::vector<int32_t> eidx, cp, ri;  std::vector<double> gv;
    if(epsl_index.isNotNull()){
        arma::uvec ei = as<arma::uvec>(epsl_index);
        if(ei.n_elem){
            if(!epsl_Gi.isNotNull())    throw Rcpp::exception("variance-covariance matrix should be provided for epsilon term.");   // :260
            arma::sp_mat Gi = as<arma::sp_mat>(epsl_Gi);
            J_v = as<arma::vec>(epsl_y_J);
            if(Gi.n_cols != Gi.n_rows) throw Rcpp::exception("variance-covariance matrix should be in square.");                  // :263
            eidx.resize(ei.n_elem);
            for(arma::uword i = 0; i < ei.n_elem; i++)  eidx[i] = (int32_t)ei[i];    // stays 1-based, as R passes it (:255-256 subtracts 1)
            hb_dropin::csc(Gi, cp, ri, gv);
            a.ne = (int)eidx.size();  a.qe = (int)Gi.n_cols;  a.epsl_y_J = J_v.memptr();  a.epsl_index = eidx.data();
            a.Gi_colptr = cp.data();  a.Gi_rowidx = ri.data();  a.Gi_val = gv.data();
        }
    }
    a.seed = hb_dropin::seed_from_r();

    // outputs: vectors and matrices of the shapes the reference returns, filled in place by the driver
    const int n_records = (niter - nburn) / thin;
    arma::vec alpha = arma::zeros<arma::vec>(m), pi = arma::zeros<arma::vec>(n_fold), pip = arma::zeros<arma::vec>(m);
    arma::vec g = arma::zeros<arma::vec>(n), e = arma::zeros<arma::vec>(n), beta = arma::zeros<arma::vec>(a.nc);
    arma::vec gwas = arma::zeros<arma::vec>(nw), vr = arma::zeros<arma::vec>(a.nr), estR = arma::zeros<arma::vec>(n_levels);
    arma::vec epsilon = arma::zeros<arma::vec>(a.qe);
    arma::mat mu_rec = arma::zeros<arma::mat>(1, n_records), vg_rec = arma::zeros<arma::mat>(1, n_records);
    arma::mat ve_rec = arma::zeros<arma::mat>(1, n_records), h2_rec = arma::zeros<arma::mat>(1, n_records);
    arma::mat pi_rec = arma::zeros<arma::mat>(n_fold, n_records), alpha_rec = arma::zeros<arma::mat>(m, n_records);
    arma::mat beta_rec = arma::zeros<arma::mat>(a.nc, n_records), vr_rec = arma::zeros<arma::mat>(a.nr, n_records);
    arma::mat r_rec = arma::zeros<arma::mat>(n_levels, n_records), eps_rec = arma::zeros<arma::mat>(a.qe, n_records);
    arma::mat veps_rec = arma::zeros<arma::mat>(1, n_records), J_rec = arma::zeros<arma::mat>(1, n_records);
    o.alpha = alpha.memptr();  o.pi = pi.memptr();  o.pip = pip.memptr();  o.g = g.memptr();  o.e = e.memptr();
    o.beta = beta.memptr();  o.gwas = nw ? gwas.memptr() : NULL;  o.vr = vr.memptr();  o.estR = estR.memptr();  o.epsilon = epsilon.memptr();
    o.mu_store = mu_rec.memptr();  o.vara_store = vg_rec.memptr();  o.vare_store = ve_rec.memptr();  o.hsq_store = h2_rec.memptr();
    o.pi_store = pi_rec.memptr();  o.alpha_store = alpha_rec.memptr();  o.beta_store = beta_rec.memptr();
    o.vr_store = vr_rec.memptr();  o.estR_store = r_rec.memptr();
    o.veps_store = veps_rec.memptr();  o.J_store = J_rec.memptr();  o.epsilon_store = eps_rec.memptr();

    if(hb_bayes(&a, &o) != 0)  throw Rcpp::exception(hb_last_error());     // -> R's stop() through END_RCPP

    // the named list of Bayes.cpp:919-1040, in its order
    List out;
    List samples;
    if(a.nr){
        out["Vr"] = vr;
        samples["Vr"] = vr_rec;
    }
    out["Vg"] = o.Vg;
    out["Ve"] = o.Ve;
    out["h2"] = o.h2;
    samples["Vg"] = vg_rec;
    samples["Ve"] = ve_rec;
    samples["h2"] = h2_rec;
    out["mu"] = o.mu;
    samples["mu"] = mu_rec;
    if(a.nc){
        out["beta"] = beta;
        samples["beta"] = beta_rec;
    }
    out["alpha"] = alpha;
    samples["alpha"] = alpha_rec;
    out["pi"] = pi;
    samples["pi"] = pi_rec;
    if(a.ne){
        out["Veps"] = o.Veps;
        out["J"] = o.J;
        out["epsilon"] = epsilon;
        samples["Veps"] = veps_rec;
        samples["J"] = J_rec;
        samples["epsilon"] = eps_rec;
    }
    if(a.nr){
        List lr(2);
        lr[0] = wrap(r_levels.begin(), r_levels.end());
        lr[1] = wrap(estR);
        DataFrame r = lr;
        Rcpp::CharacterVector names(2);
        names[0] = "Levels";
        names[1] = "Estimation";
        r.attr("names") = names;
        out["r"] = r;
        samples["r"] = r_rec;
    }
    out["g"] = g;
    out["e"] = e;
    out["pip"] = pip;
    if(nw)  out["gwas"] = gwas;
    out["MCMCsamples"] = samples;
    return out;
}
