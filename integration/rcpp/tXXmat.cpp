// Drop-in bodies for the char big.matrix paths of hibayes' src/tXXmat.cpp: BigStat() (:43-98), tXXmat_Geno() (:100-206) and
// tXXmat_Chr() (:504-626) on the GPU through hb_ldmat_*() (tcgen05 kind::i8 Gram + the reference's epilogue arithmetic,
// include/hibayes_b200.h).  The exported signatures and what they return (list of three NumericVectors; arma::mat for the
// dense genome-wide matrix, arma::sp_mat everywhere else) stay.  The other big.matrix types and the _gwas variants keep the
// reference's CPU loops (a maintainer leaves those functions of the original file in place).
#include "hb_dropin.h"
#include <bigmemory/BigMatrix.h>
#include <bigmemory/MatrixAccessor.hpp>

using namespace Rcpp;

namespace {
struct Handle {   // hb_ldmat with the char matrix loaded; destroyed on every path out
  hb_ldmat* h;
  explicit Handle(XPtr<BigMatrix>& pMat) : h(NULL) {
    if(pMat->matrix_type() != 1)  throw Rcpp::exception("the GPU path takes a big.matrix of type char (what read_plink() builds)");
    const int n = pMat->nrow(), m = pMat->ncol();
    if(hb_ldmat_create(0, n, m, &h) != 0 || hb_ldmat_load_i8(h, reinterpret_cast<const int8_t*>(pMat->matrix()), (size_t)n) != 0){
        std::string msg = hb_last_error();  hb_ldmat_destroy(h);  throw Rcpp::exception(msg.c_str()); }
  }
  ~Handle() { hb_ldmat_destroy(h); }
};
SEXP sparse_result(Handle& H, const int32_t* chr, int has_chisq, double chisq, int m){
    long long nnz = 0;
    if(hb_ldmat_sparse(H.h, chr, has_chisq, chisq, &nnz) != 0)  throw Rcpp::exception(hb_last_error());
    std::vector<long long> cp(m + 1);  std::vector<int32_t> r32(nnz);
    arma::vec v = arma::zeros<arma::vec>(nnz);
    if(hb_ldmat_sparse_get(H.h, cp.data(), r32.data(), v.memptr()) != 0)  throw Rcpp::exception(hb_last_error());
    arma::uvec ri(nnz), cpu(m + 1);
    for(long long k = 0; k < nnz; k++)  ri[k] = r32[k];
    for(int j = 0; j <= m; j++)  cpu[j] = cp[j];
    return wrap(arma::sp_mat(ri, cpu, v, m, m));      // a dgCMatrix on the R side, as today
}
}  // namespace

// [[Rcpp::export]]
SEXP BigStat(SEXP pBigMat, const int threads = 0){
    XPtr<BigMatrix> xpMat(pBigMat);
    Handle H(xpMat);
    const int m = xpMat->ncol();
    NumericVector mean(m), sum(m), sd(m);
    if(hb_ldmat_stats(H.h, &mean[0], &sum[0], &sd[0]) != 0)  throw Rcpp::exception(hb_last_error());
    return List::create(Named("mean") = mean, Named("sum") = sum, Named("xx") = sd);
}

// [[Rcpp::export]]
SEXP tXXmat_Geno(SEXP pBigMat, const Nullable<double> chisq = R_NilValue, const int threads=0, const bool verbose=true){
    XPtr<BigMatrix> xpMat(pBigMat);
    Handle H(xpMat);
    const int m = xpMat->ncol();
    const bool sparse = chisq.isNotNull() && as<double>(chisq) > 0;                  // tXXmat.cpp:117-120
    if(sparse)  return sparse_result(H, NULL, 1, as<double>(chisq), m);
    arma::mat ldmat = arma::zeros<arma::mat>(m, m);
    if(hb_ldmat_dense(H.h, NULL, 0, 0.0, ldmat.memptr(), (size_t)m) != 0)  throw Rcpp::exception(hb_last_error());
    return wrap(ldmat);
}

// [[Rcpp::export]]
SEXP tXXmat_Chr(SEXP pBigMat, const NumericVector chr, const Nullable<double> chisq = R_NilValue, const int threads=0, const bool verbose=true){
    XPtr<BigMatrix> xpMat(pBigMat);
    Handle H(xpMat);
    const int m = xpMat->ncol();
    std::vector<int32_t> code(m);                                                    // only equality of the codes matters (:524-527)
    std::map<double, int32_t> lut;
    for(int j = 0; j < m; j++){
        std::map<double, int32_t>::iterator it = lut.find(chr[j]);
        if(it == lut.end())  it = lut.insert(std::make_pair((double)chr[j], (int32_t)lut.size())).first;
        code[j] = it->second;
    }
    const bool has = chisq.isNotNull();                                              // :520-523: any chisq, also 0
    return sparse_result(H, code.data(), has ? 1 : 0, has ? as<double>(chisq) : 0.0, m);   // both branches return an sp_mat
}
