// hb_dropin.h -- helpers shared by the drop-in bodies of Bayes(), SBayesD(), SBayesS() (integration/rcpp/*.cpp).
// These files replace src/Bayes.cpp, src/SBayesD.cpp, src/SBayesS.cpp of hibayes: same C++ signatures, same Rcpp::List,
// the work done by libhibayes_b200.so through its C ABI (include/hibayes_b200.h).  INTEGRATION.md says how they are built
// into the package; tests/test_dropin_rcpp.py builds them here and runs them on the GPU next to the reference's own files.
#ifndef HB_DROPIN_H
#define HB_DROPIN_H
#include <RcppArmadillo.h>
#include <R.h>
#include <Rmath.h>

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include <hibayes_b200.h>

namespace hb_dropin {
// Nullable<double> -> NaN = R_NilValue
inline double opt(const Rcpp::Nullable<double>& v) { return v.isNotNull() ? Rcpp::as<double>(v) : HB_NA; }
// The reference takes no seed argument: set.seed(seed) in ibrm() / sbrm() (R/bayes.r:151) seeds R's stream and the
// exported wrapper's RNGScope is active here.  Two uniforms from that stream make the 64-bit run key of the
// position-addressed generator (csrc/hb_rng.h), so the same set.seed() gives the same chain.
inline uint64_t seed_from_r() {
  const uint64_t hi = (uint64_t)(unif_rand() * 4294967296.0), lo = (uint64_t)(unif_rand() * 4294967296.0);
  return (hi << 32) | lo;
}
// Nullable<arma::uvec> windindx -> int32, returns the number of windows
inline int windows(const Rcpp::Nullable<arma::uvec>& windindx, std::vector<int32_t>& w) {
  if (!windindx.isNotNull()) return 0;
  arma::uvec v = Rcpp::as<arma::uvec>(windindx);
  w.resize(v.n_elem);
  int nw = 0;
  for (arma::uword i = 0; i < v.n_elem; ++i) { w[i] = (int32_t)v[i]; nw = std::max(nw, (int)v[i]); }
  return nw;
}
// arma::sp_mat -> int32 compressed columns, walked the way the reference walks it (SBayesS.cpp:132-141)
inline void csc(const arma::sp_mat& A, std::vector<int32_t>& colptr, std::vector<int32_t>& rowidx, std::vector<double>& val) {
  colptr.assign(A.n_cols + 1, 0);
  rowidx.clear(); val.clear();
  arma::sp_mat::const_iterator it, end;
  for (arma::uword j = 0; j < A.n_cols; ++j) {
    it = A.begin_col(j); end = A.end_col(j);
    for (; it != end; ++it) { rowidx.push_back((int32_t)it.row()); val.push_back(*it); }
    colptr[j + 1] = (int32_t)rowidx.size();
  }
}
}  // namespace hb_dropin
#endif
