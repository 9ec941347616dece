/*
 * RcppArmadillo.h (stand-in) -- TEST INFRASTRUCTURE ONLY (oracle/).  With mini_arma.h this is the third-party layer
 * (R, Rcpp, RcppArmadillo) that the reference's C++ sources include, reduced to what Bayes.cpp, SBayesD.cpp, SBayesS.cpp,
 * stats.cpp and solver.cpp touch, so that those files compile unmodified into oracle/_ref/libhibayes_ref.so.
 *
 * The one deliberate substitution is the random stream.  libR's generator is sequential and absent; the CPU oracle
 * (oracle/hb_oracle.c) draws by address.  Here unif_rand(), norm_rand(), R::rgamma(), R::rchisq() (and arma::randn)
 * REPLAY A TAPE: the list of variates the oracle consumed on the same inputs, in the order of the reference's own
 * calls.  Every pop checks the kind of draw and, for gamma / chi-square, that the shape the reference asks for is
 * bit-identical to the one the oracle used -- a mismatch (a different call order, one draw more or fewer, another
 * degree of freedom) aborts the run with the position on the tape.  So the compiled reference and the oracle see the
 * same random numbers, and every output that then agrees was computed by the reference's own statements.
 */
#ifndef HB_RCPPARMADILLO_SHIM_H
#define HB_RCPPARMADILLO_SHIM_H
#include <typeinfo>
#include <cstdarg>
#include <cstdio>
#include <iostream>
#include <map>
#include <memory>
#include <optional>
#include <sstream>
#include <string>
#include <vector>

#include "mini_arma.h"

// ---- R's C API, as far as it is named ------------------------------------------------------------------------------
namespace hb_shim {
// a value of any type (std::any would do, but its name collides with arma::any under the reference's using-directives)
class Box {
  std::shared_ptr<void> p_;
  const std::type_info* t_ = nullptr;
 public:
  Box() {}
  template <class T> explicit Box(const T& v) : p_(std::make_shared<T>(v)), t_(&typeid(T)) {}
  bool empty() const { return !p_; }
  template <class T> const T& get() const {
    if (!p_ || *t_ != typeid(T)) throw std::runtime_error(std::string("stand-in Rcpp: value holds ") + (t_ ? t_->name() : "nothing") + ", asked for " + typeid(T).name());
    return *static_cast<const T*>(p_.get());
  }
};
}  // namespace hb_shim
struct SEXPREC { hb_shim::Box payload; };
typedef SEXPREC* SEXP;
#define R_NilValue ((SEXP) nullptr)
#ifndef TRUE
#define TRUE 1
#endif
#ifndef FALSE
#define FALSE 0
#endif
extern "C" {
double unif_rand(void);
double norm_rand(void);
void Rprintf(const char* fmt, ...);
void REprintf(const char* fmt, ...);
}
namespace R {
double rgamma(double shape, double scale);
double rchisq(double df);
double rnorm(double mu, double sd);
double runif(double a, double b);
double rbeta(double a, double b);
double rt(double df);
double rcauchy(double loc, double scale);
double rexp(double scale);
}  // namespace R

// ---- BLAS / LAPACK entry points the reference calls directly (hibayes.h:21-29, solver.cpp) ---------------------------
extern "C" {
double ddot_(const int* n, const double* x, const int* incx, const double* y, const int* incy);
void daxpy_(const int* n, const double* a, const double* x, const int* incx, double* y, const int* incy);
void dpotrf_(const char* uplo, const int* n, double* a, const int* lda, int* info);
void dpotri_(const char* uplo, const int* n, double* a, const int* lda, int* info);
void dgetrf_(const int* m, const int* n, double* a, const int* lda, int* ipiv, int* info);
void dgetri_(const int* n, double* a, const int* lda, const int* ipiv, double* work, const int* lwork, int* info);
double dlange_(const char* norm, const int* m, const int* n, const double* a, const int* lda, double* work);
void dgecon_(const char* norm, const int* n, const double* a, const int* lda, const double* anorm, double* rcond, double* work, int* iwork, int* info);
void dsyevd_(const char* jobz, const char* uplo, const int* n, double* a, const int* lda, double* w, double* work, const int* lwork, int* iwork, const int* liwork, int* info);
}
#define ARMA_USE_LAPACK 1
#define arma_fortran(x) x##_

namespace Rcpp {
class exception : public std::exception {
  std::string msg_;
 public:
  explicit exception(const char* m) : msg_(m) {}
  exception(const char* m, bool) : msg_(m) {}
  const char* what() const noexcept override { return msg_.c_str(); }
};
inline void stop(const std::string& m) { throw exception(m.c_str()); }

struct NamePlaceholder {};
static const NamePlaceholder _ = NamePlaceholder();

// a value of any type travelling through List / wrap / SEXP
template <class T> inline SEXP box(const T& v) { SEXP s = new SEXPREC; s->payload = hb_shim::Box(v); return s; }   // (never freed: test processes are short)

template <class T> class Nullable {
  std::optional<T> v_;
 public:
  Nullable() {}
  Nullable(SEXP s) { if (s) v_ = s->payload.get<T>(); }
  Nullable(const T& v) : v_(v) {}
  bool isNotNull() const { return v_.has_value(); }
  bool isNull() const { return !v_.has_value(); }
  const T& get() const { if (!v_) throw exception("Nullable: value is NULL"); return *v_; }
};

// attribute sink: r.attr("names") = x compiles and is dropped
struct AttrProxy { template <class T> AttrProxy& operator=(const T&) { return *this; } };

template <class T> class Vector {
 protected:
  std::shared_ptr<std::vector<T>> d_;
 public:
  Vector() : d_(std::make_shared<std::vector<T>>()) {}
  Vector(int n) : d_(std::make_shared<std::vector<T>>(n)) {}
  Vector(size_t n) : d_(std::make_shared<std::vector<T>>(n)) {}
  Vector(SEXP s) : d_(std::make_shared<std::vector<T>>(s->payload.get<std::vector<T>>())) {}
  Vector(const std::vector<T>& v) : d_(std::make_shared<std::vector<T>>(v)) {}
  int length() const { return (int)d_->size(); }
  int size() const { return (int)d_->size(); }
  T& operator[](int i) { return (*d_)[i]; }
  const T& operator[](int i) const { return (*d_)[i]; }
  T& operator()(int i) { return (*d_)[i]; }
  void fill(const T& v) { std::fill(d_->begin(), d_->end(), v); }
  AttrProxy attr(const char*) { return AttrProxy(); }
  operator SEXP() const { return box(*d_); }
  const std::vector<T>& vec() const { return *d_; }
  typename std::vector<T>::iterator begin() { return d_->begin(); }
  typename std::vector<T>::iterator end() { return d_->end(); }
};
class NumericVector : public Vector<double> {
 public:
  using Vector<double>::Vector;
  NumericVector() {}
  static bool is_na(double v) { return v != v; }
};
class IntegerVector : public Vector<int> {
 public:
  using Vector<int>::Vector;
  IntegerVector() {}
};
class LogicalVector : public Vector<int> {
 public:
  using Vector<int>::Vector;
  LogicalVector() {}
};
inline LogicalVector rep(bool v, int n) { LogicalVector o(n); o.fill(v ? 1 : 0); return o; }
// external pointer: the SEXP carries a T*
template <class T> class XPtr {
  T* p_;
 public:
  XPtr(SEXP s) : p_(s->payload.get<T*>()) {}
  explicit XPtr(T* p) : p_(p) {}
  T* operator->() const { return p_; }
  T& operator*() const { return *p_; }
  operator SEXP() const { return box(p_); }
};
// R's NA_character_ is a sentinel object; here it is this string
static const char* const NA_STRING_SHIM = "\x01NA\x01";
class CharacterVector : public Vector<std::string> {
 public:
  using Vector<std::string>::Vector;
  CharacterVector() {}
  static bool is_na(const std::string& v) { return v == NA_STRING_SHIM; }
};
class CharacterMatrix {
  int nr_ = 0, nc_ = 0;
  std::shared_ptr<std::vector<std::string>> d_;
 public:
  CharacterMatrix() : d_(std::make_shared<std::vector<std::string>>()) {}
  CharacterMatrix(int nr, int nc) : nr_(nr), nc_(nc), d_(std::make_shared<std::vector<std::string>>((size_t)nr * nc)) {}
  int nrow() const { return nr_; }
  int ncol() const { return nc_; }
  std::string& operator()(int i, int j) { return (*d_)[(size_t)j * nr_ + i]; }
  const std::string& operator()(int i, int j) const { return (*d_)[(size_t)j * nr_ + i]; }
  CharacterVector operator()(NamePlaceholder, int j) const {
    std::vector<std::string> c(d_->begin() + (size_t)j * nr_, d_->begin() + (size_t)(j + 1) * nr_);
    return CharacterVector(c);
  }
  static bool is_na(const std::string& v) { return v == NA_STRING_SHIM; }
};

template <class T> struct NamedValue { std::string name; T value; };
struct Named {
  std::string name;
  explicit Named(const char* n) : name(n) {}
  template <class T> NamedValue<T> operator=(const T& v) const { return NamedValue<T>{name, v}; }
};

class List {
  struct Slot { std::string name; hb_shim::Box value; };
  std::shared_ptr<std::vector<Slot>> d_;
 public:
  struct Proxy {
    hb_shim::Box* a;
    template <class T> Proxy& operator=(const T& v) { *a = hb_shim::Box(v); return *this; }
    Proxy& operator=(SEXP s) { *a = s ? s->payload : hb_shim::Box(); return *this; }
    template <class T> operator T() const { return a->get<T>(); }
  };
  List() : d_(std::make_shared<std::vector<Slot>>()) {}
  List(SEXP s) : d_(s->payload.get<List>().d_) {}
  operator SEXP() const { return box(*this); }
  explicit List(int n) : d_(std::make_shared<std::vector<Slot>>(n)) {}
  Proxy operator[](int i) { return Proxy{&(*d_)[i].value}; }
  Proxy operator[](const char* name) { return (*this)[std::string(name)]; }
  Proxy operator[](const std::string& name) {
    for (auto& s : *d_) if (s.name == name) return Proxy{&s.value};
    d_->push_back(Slot{name, hb_shim::Box()});
    return Proxy{&d_->back().value};
  }
  bool has(const std::string& name) const { for (auto& s : *d_) if (s.name == name) return true; return false; }
  template <class T> const T& get(const std::string& name) const {
    for (auto& s : *d_) if (s.name == name) return s.value.template get<T>();
    throw exception(("List: no element named " + name).c_str());
  }
  template <class... A> static List create(const NamedValue<A>&... nv) { List l; (l.d_->push_back(Slot{nv.name, hb_shim::Box(nv.value)}), ...); return l; }
  AttrProxy attr(const char*) { return AttrProxy(); }
  int size() const { return (int)d_->size(); }
};
typedef List DataFrame;

// as<>
template <class T, class U> struct Conv { static T go(const U& u) { return T(u); } };
template <class T> struct Conv<T, NumericVector> {
  static T go(const NumericVector& v) { T o; o.init(v.size(), 1); for (int i = 0; i < v.size(); ++i) o.mem[i] = v[i]; return o; }
};
template <> struct Conv<NumericVector, NumericVector> { static NumericVector go(const NumericVector& v) { return v; } };
template <class T, class U> inline T as(const Nullable<U>& n) { return Conv<T, U>::go(n.get()); }
template <class T> inline T as(const T& v) { return v; }
template <class T> inline T as(const NumericVector& v) { return Conv<T, NumericVector>::go(v); }
template <class T> inline T as(const CharacterVector& v) { return T(v.vec()); }   // std::vector<std::string>

// wrap
template <class It> inline SEXP wrap(It b, It e) { return box(std::vector<typename std::iterator_traits<It>::value_type>(b, e)); }
template <class T> inline SEXP wrap(const T& v) { return box(v); }

// output streams: the reference prints progress; the stand-in drops it unless HB_REF_VERBOSE is set
class Rostream : public std::ostream {
  struct NullBuf : std::streambuf { int overflow(int c) override { return c; } } nb_;
 public:
  explicit Rostream(bool err);
};
extern Rostream Rcout;
extern Rostream Rcerr;
}  // namespace Rcpp

#endif
