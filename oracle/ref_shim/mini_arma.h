/*
 * mini_arma.h -- TEST INFRASTRUCTURE ONLY (oracle/).  A small, eager stand-in for the part of the Armadillo API that
 * the reference's Bayes.cpp / SBayesD.cpp / SBayesS.cpp / stats.cpp / solver.cpp use, so that those files can be
 * compiled UNMODIFIED, from where they lie under /root/reference/src, into oracle/_ref/libhibayes_ref.so
 * (oracle/Makefile, target _ref).  Armadillo itself (and R, Rcpp) is absent from this image.
 *
 * Nothing here restates the reference: it is the third-party layer under it.  Where Armadillo's arithmetic order is
 * visible in fp64 results it follows Armadillo's published algorithms:
 *   arrayops::accumulate / op_mean::direct_mean  two interleaved accumulators (sum, mean)
 *   op_var::direct_var                           two-pass with the (acc3^2 / n) correction, norm_type 0
 *   op_dot::direct_dot                           n <= 32: two accumulators; larger: BLAS ddot (unit stride, one accumulator)
 *   sparse x dense, dot(sparse column, dense)    one accumulator over the stored entries in column order
 *   dense matrix x vector                        reference-BLAS dgemv order (N: column AXPYs; T: one dot per column)
 * No expression templates: every operator returns a concrete object.
 */
#ifndef HB_MINI_ARMA_H
#define HB_MINI_ARMA_H
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <ostream>
#include <limits>
#include <stdexcept>
#include <string>
#include <vector>

namespace arma {
using std::endl;
typedef unsigned long long uword;
typedef long long sword;
typedef int blas_int;
struct datum { static constexpr double eps = 2.220446049250313e-16; static constexpr double pi = 3.14159265358979323846; };

template <class T> struct Mat;
template <class T> struct Col;
template <class T> struct SpMat;

// ---- reductions with Armadillo's operation order --------------------------------------------------------------
template <class T> inline T acc2(const T* x, uword n) {
  T a1 = T(0), a2 = T(0);
  uword i, j;
  for (i = 0, j = 1; j < n; i += 2, j += 2) { a1 += x[i]; a2 += x[j]; }
  if (i < n) a1 += x[i];
  return a1 + a2;
}
inline double direct_mean(const double* x, uword n) { return acc2(x, n) / double(n); }   // (finite inputs)
inline double direct_var(const double* x, uword n, int norm_type = 0) {
  if (n < 2) return 0.0;
  const double mu = direct_mean(x, n);
  double acc_2 = 0.0, acc_3 = 0.0;
  uword i, j;
  for (i = 0, j = 1; j < n; i += 2, j += 2) {
    const double ti = mu - x[i], tj = mu - x[j];
    acc_2 += ti * ti + tj * tj;
    acc_3 += ti + tj;
  }
  if (i < n) { const double ti = mu - x[i]; acc_2 += ti * ti; acc_3 += ti; }
  const double nn = double(n), nd = norm_type == 0 ? nn - 1.0 : nn;
  return (acc_2 - acc_3 * acc_3 / nn) / nd;
}
inline double direct_dot(const double* a, const double* b, uword n) {
  if (n <= 32) {
    double v1 = 0.0, v2 = 0.0;
    uword i, j;
    for (i = 0, j = 1; j < n; i += 2, j += 2) { v1 += a[i] * b[i]; v2 += a[j] * b[j]; }
    if (i < n) v1 += a[i] * b[i];
    return v1 + v2;
  }
  double s = 0.0;
  for (uword i = 0; i < n; ++i) s += a[i] * b[i];
  return s;
}

// ---- views -----------------------------------------------------------------------------------------------------
template <class T> struct subview_col {   // contiguous piece of a column (col(), subvec(), tail(), unsafe_col())
  T* p; uword n_elem; uword n_rows; static const uword n_cols = 1;
  subview_col(T* p_, uword n) : p(p_), n_elem(n), n_rows(n) {}
  subview_col& operator=(const Mat<T>& v);
  subview_col& operator=(const subview_col& v) { if (v.n_elem != n_elem) throw std::logic_error("subview_col: size"); std::memmove(p, v.p, n_elem * sizeof(T)); return *this; }
  subview_col& operator+=(const Mat<T>& v);
  subview_col& operator-=(const Mat<T>& v);
  void fill(T v) { for (uword i = 0; i < n_elem; ++i) p[i] = v; }
  void zeros() { fill(T(0)); }
  T& operator[](uword i) { return p[i]; }
  T operator[](uword i) const { return p[i]; }
  T& operator()(uword i) { return p[i]; }
  const T* memptr() const { return p; }
};
template <class T> struct subview_row {
  Mat<T>* m; uword r;
  void fill(T v);
};
template <class T> struct subview_elem {
  Mat<T>* m; std::vector<uword> idx;
  void fill(T v);
  void zeros() { fill(T(0)); }
  // Armadillo checks the indices when the view is evaluated (ARMA_NO_DEBUG is not set by the reference's Makevars)
  void check() const { for (uword i : idx) if (i >= m->n_elem) throw std::logic_error("Mat::elem(): index out of bounds"); }
};
template <class T> struct diagview {
  Mat<T>* m;
  diagview& operator+=(T v);
  diagview& operator=(const Mat<T>& v);
  operator Col<T>() const;
};
template <class T> struct each_col_view {
  Mat<T>* m;
  void operator+=(const Mat<T>& v);
  void operator-=(const Mat<T>& v);
};

// ---- dense -----------------------------------------------------------------------------------------------------
template <class T> struct Mat {
  typedef T elem_type;
  uword n_rows = 0, n_cols = 0, n_elem = 0;
  std::vector<T> mem;
  bool is_col = false;
  Mat() {}
  Mat(uword r, uword c) { init(r, c); }
  Mat(const subview_col<T>& s) { is_col = true; init(s.n_elem, 1); std::copy(s.p, s.p + s.n_elem, mem.begin()); }
  Mat(const subview_elem<T>& s) { s.check(); is_col = true; init(s.idx.size(), 1); for (uword i = 0; i < n_elem; ++i) mem[i] = s.m->mem[s.idx[i]]; }
  void init(uword r, uword c) { n_rows = r; n_cols = c; n_elem = r * c; mem.assign(n_elem, T(0)); }
  // resize keeps the elements (column-major, as Armadillo's op_resize)
  void resize(uword r, uword c) {
    Mat<T> o; o.is_col = is_col; o.init(r, c);
    for (uword j = 0; j < std::min(c, n_cols); ++j) for (uword i = 0; i < std::min(r, n_rows); ++i) o.mem[j * r + i] = mem[j * n_rows + i];
    n_rows = r; n_cols = c; n_elem = r * c; mem.swap(o.mem);
  }
  void resize(uword n) { if (is_col || n_cols <= 1) resize(n, 1); else resize(1, n); }
  void set_size(uword r, uword c) { init(r, c); }
  void set_size(uword n) { init(n, 1); }
  Mat& zeros() { std::fill(mem.begin(), mem.end(), T(0)); return *this; }
  Mat& zeros(uword n) { init(n, 1); return *this; }
  Mat& zeros(uword r, uword c) { init(r, c); return *this; }
  Mat& ones() { std::fill(mem.begin(), mem.end(), T(1)); return *this; }
  Mat& ones(uword n) { init(n, 1); return ones(); }
  Mat& fill(T v) { std::fill(mem.begin(), mem.end(), v); return *this; }
  T* memptr() { return mem.data(); }
  const T* memptr() const { return mem.data(); }
  T* colptr(uword j) { return mem.data() + j * n_rows; }
  const T* colptr(uword j) const { return mem.data() + j * n_rows; }
  T& operator[](uword i) { return mem[i]; }
  const T& operator[](uword i) const { return mem[i]; }
  T& operator()(uword i) { return mem[i]; }
  const T& operator()(uword i) const { return mem[i]; }
  T& operator()(uword i, uword j) { return mem[j * n_rows + i]; }
  const T& operator()(uword i, uword j) const { return mem[j * n_rows + i]; }
  T& at(uword i, uword j) { return mem[j * n_rows + i]; }
  bool is_empty() const { return n_elem == 0; }
  bool has_nan() const { for (const T& v : mem) if (v != v) return true; return false; }
  subview_col<T> col(uword j) { return subview_col<T>(colptr(j), n_rows); }
  subview_col<T> col(uword j) const { return subview_col<T>(const_cast<T*>(colptr(j)), n_rows); }
  subview_col<T> unsafe_col(uword j) { return col(j); }
  subview_row<T> row(uword r) { return subview_row<T>{this, r}; }
  subview_col<T> subvec(uword a, uword b) { return subview_col<T>(mem.data() + a, b - a + 1); }
  subview_col<T> subvec(uword a, uword b) const { return subview_col<T>(const_cast<T*>(mem.data()) + a, b - a + 1); }
  subview_col<T> tail(uword k) { return subview_col<T>(mem.data() + (n_elem - k), k); }
  subview_col<T> head(uword k) { return subview_col<T>(mem.data(), k); }
  subview_elem<T> elem(const Mat<uword>& ix) { subview_elem<T> s; s.m = this; s.idx.assign(ix.mem.begin(), ix.mem.end()); return s; }
  subview_elem<T> elem(const Mat<uword>& ix) const { subview_elem<T> s; s.m = const_cast<Mat<T>*>(this); s.idx.assign(ix.mem.begin(), ix.mem.end()); return s; }
  subview_elem<T> operator()(const Mat<uword>& ix) const { return elem(ix); }
  diagview<T> diag() { return diagview<T>{this}; }
  each_col_view<T> each_col() { return each_col_view<T>{this}; }
  Mat<T> t() const {
    Mat<T> o(n_cols, n_rows);
    for (uword j = 0; j < n_cols; ++j) for (uword i = 0; i < n_rows; ++i) o.mem[i * n_cols + j] = mem[j * n_rows + i];
    return o;
  }
  Mat& operator+=(const Mat& b) { chk(b); for (uword i = 0; i < n_elem; ++i) mem[i] += b.mem[i]; return *this; }
  Mat& operator-=(const Mat& b) { chk(b); for (uword i = 0; i < n_elem; ++i) mem[i] -= b.mem[i]; return *this; }
  Mat& operator/=(const Mat& b) { chk(b); for (uword i = 0; i < n_elem; ++i) mem[i] /= b.mem[i]; return *this; }
  Mat& operator%=(const Mat& b) { chk(b); for (uword i = 0; i < n_elem; ++i) mem[i] *= b.mem[i]; return *this; }
  Mat& operator+=(T v) { for (T& x : mem) x += v; return *this; }
  Mat& operator-=(T v) { for (T& x : mem) x -= v; return *this; }
  Mat& operator*=(T v) { for (T& x : mem) x *= v; return *this; }
  Mat& operator/=(T v) { for (T& x : mem) x /= v; return *this; }
  void chk(const Mat& b) const { if (b.n_elem != n_elem) throw std::logic_error("mini_arma: element-wise operation on different sizes"); }
};
template <class T> struct Col : Mat<T> {
  Col() { this->is_col = true; this->n_cols = 1; }
  explicit Col(uword n) { this->is_col = true; this->init(n, 1); }
  Col(const Mat<T>& m) : Mat<T>(m) { this->is_col = true; if (m.n_cols != 1 && m.n_elem) { if (m.n_rows != 1) throw std::logic_error("mini_arma: matrix is not a vector"); this->n_rows = m.n_elem; this->n_cols = 1; } if (!m.n_elem) { this->n_rows = 0; this->n_cols = 1; } }
  Col(const subview_col<T>& s) : Mat<T>(s) {}
  Col(const subview_elem<T>& s) : Mat<T>(s) {}
  Col(const std::vector<T>& v) { this->is_col = true; this->init(v.size(), 1); this->mem = v; }
};
typedef Mat<double> mat;
typedef Col<double> vec;
typedef Col<double> colvec;
typedef Mat<double> rowvec;
typedef Col<uword> uvec;
typedef Col<sword> ivec;
typedef Mat<uword> umat;

template <class T> subview_col<T>& subview_col<T>::operator=(const Mat<T>& v) { if (v.n_elem != n_elem) throw std::logic_error("subview_col: size"); std::copy(v.mem.begin(), v.mem.end(), p); return *this; }
template <class T> subview_col<T>& subview_col<T>::operator+=(const Mat<T>& v) { if (v.n_elem != n_elem) throw std::logic_error("subview_col: size"); for (uword i = 0; i < n_elem; ++i) p[i] += v.mem[i]; return *this; }
template <class T> subview_col<T>& subview_col<T>::operator-=(const Mat<T>& v) { if (v.n_elem != n_elem) throw std::logic_error("subview_col: size"); for (uword i = 0; i < n_elem; ++i) p[i] -= v.mem[i]; return *this; }
template <class T> void subview_row<T>::fill(T v) { for (uword j = 0; j < m->n_cols; ++j) (*m)(r, j) = v; }
template <class T> void subview_elem<T>::fill(T v) { check(); for (uword i : idx) m->mem[i] = v; }
template <class T> diagview<T>& diagview<T>::operator+=(T v) { for (uword i = 0; i < std::min(m->n_rows, m->n_cols); ++i) (*m)(i, i) += v; return *this; }
template <class T> diagview<T>& diagview<T>::operator=(const Mat<T>& v) { for (uword i = 0; i < std::min(m->n_rows, m->n_cols); ++i) (*m)(i, i) = v.mem[i]; return *this; }
template <class T> diagview<T>::operator Col<T>() const { Col<T> o(std::min(m->n_rows, m->n_cols)); for (uword i = 0; i < o.n_elem; ++i) o.mem[i] = (*m)(i, i); return o; }
template <class T> void each_col_view<T>::operator+=(const Mat<T>& v) { for (uword j = 0; j < m->n_cols; ++j) for (uword i = 0; i < m->n_rows; ++i) (*m)(i, j) += v.mem[i]; }
template <class T> void each_col_view<T>::operator-=(const Mat<T>& v) { for (uword j = 0; j < m->n_cols; ++j) for (uword i = 0; i < m->n_rows; ++i) (*m)(i, j) -= v.mem[i]; }

// generators
template <class V = vec> inline V zeros(uword n) { V o; o.init(n, 1); return o; }
template <class V = mat> inline V zeros(uword r, uword c) { V o; o.init(r, c); return o; }
template <class V = vec> inline V ones(uword n) { V o; o.init(n, 1); o.ones(); return o; }
double mini_arma_randn();   // (RcppArmadillo draws through R's generator; provided by the R shim: the tape)
template <class V = vec> inline V randn(uword n) { V o; o.init(n, 1); for (uword i = 0; i < n; ++i) o.mem[i] = mini_arma_randn(); return o; }

// element-wise binary operators (dense); results keep the shape of the left operand
#define HB_EW(op, sym)                                                                                     \
  template <class T> inline Mat<T> operator sym(const Mat<T>& a, const Mat<T>& b) { a.chk(b); Mat<T> o(a); for (uword i = 0; i < a.n_elem; ++i) o.mem[i] = a.mem[i] op b.mem[i]; return o; }
HB_EW(+, +) HB_EW(-, -) HB_EW(*, %) HB_EW(/, /)
#undef HB_EW
template <class T> inline Mat<T> operator+(const Mat<T>& a, T s) { Mat<T> o(a); for (T& x : o.mem) x += s; return o; }
template <class T> inline Mat<T> operator+(T s, const Mat<T>& a) { Mat<T> o(a); for (T& x : o.mem) x = s + x; return o; }
template <class T> inline Mat<T> operator-(const Mat<T>& a, T s) { Mat<T> o(a); for (T& x : o.mem) x -= s; return o; }
template <class T> inline Mat<T> operator-(T s, const Mat<T>& a) { Mat<T> o(a); for (T& x : o.mem) x = s - x; return o; }
template <class T> inline Mat<T> operator*(const Mat<T>& a, T s) { Mat<T> o(a); for (T& x : o.mem) x *= s; return o; }
template <class T> inline Mat<T> operator*(T s, const Mat<T>& a) { Mat<T> o(a); for (T& x : o.mem) x *= s; return o; }
template <class T> inline Mat<T> operator/(const Mat<T>& a, T s) { Mat<T> o(a); for (T& x : o.mem) x /= s; return o; }
template <class T> inline Mat<T> operator/(T s, const Mat<T>& a) { Mat<T> o(a); for (T& x : o.mem) x = s / x; return o; }
// int scalars with double objects (1 / Kval, fold_snp_num + 1, Kval * 2 ...)
inline mat operator+(const mat& a, int s) { return a + double(s); }
inline mat operator-(const mat& a, int s) { return a - double(s); }
inline mat operator*(const mat& a, int s) { return a * double(s); }
inline mat operator/(const mat& a, int s) { return a / double(s); }
inline mat operator+(int s, const mat& a) { return double(s) + a; }
inline mat operator-(int s, const mat& a) { return double(s) - a; }
inline mat operator*(int s, const mat& a) { return double(s) * a; }
inline mat operator/(int s, const mat& a) { return double(s) / a; }
// subviews on either side of an operator are materialised first
template <class T> inline Mat<T> operator-(const subview_col<T>& a, const Mat<T>& b) { return Mat<T>(a) - b; }
template <class T> inline Mat<T> operator-(const Mat<T>& a, const subview_col<T>& b) { return a - Mat<T>(b); }
template <class T> inline Mat<T> operator+(const subview_col<T>& a, const Mat<T>& b) { return Mat<T>(a) + b; }
template <class T> inline Mat<T> operator+(const Mat<T>& a, const subview_col<T>& b) { return a + Mat<T>(b); }
template <class T> inline Mat<T> operator%(const Mat<T>& a, const subview_elem<T>& b) { return a % Mat<T>(b); }
template <class T> inline Mat<T> operator-(const Mat<T>& a) { Mat<T> o(a); for (T& x : o.mem) x = -x; return o; }

// matrix product, reference-BLAS order
inline mat operator*(const mat& A, const mat& B) {
  if (A.n_cols != B.n_rows) throw std::logic_error("mini_arma: matrix product of incompatible sizes");
  mat C(A.n_rows, B.n_cols);
  if (A.n_rows == 1) {   // row vector x matrix: one dot per column of B
    for (uword j = 0; j < B.n_cols; ++j) { double s = 0.0; for (uword k = 0; k < A.n_cols; ++k) s += A.mem[k] * B(k, j); C.mem[j] = s; }
    return C;
  }
  for (uword j = 0; j < B.n_cols; ++j)
    for (uword k = 0; k < A.n_cols; ++k) {
      const double b = B(k, j);
      const double* a = A.colptr(k);
      double* c = C.colptr(j);
      for (uword i = 0; i < A.n_rows; ++i) c[i] += b * a[i];
    }
  return C;
}
inline mat operator*(const mat& A, const subview_col<double>& b) { return A * mat(b); }
// A.t() * v without forming the transpose would be the same sums; the reference writes X.t() * v, which Armadillo
// hands to dgemv('T'): one dot per column of X -- identical to transposing first and taking row dots in order.

// comparisons -> uvec
#define HB_CMP(sym)                                                                                                                   \
  template <class T, class S> inline uvec operator sym(const Mat<T>& a, S s) { uvec o(a.n_elem); for (uword i = 0; i < a.n_elem; ++i) o.mem[i] = a.mem[i] sym T(s) ? 1 : 0; return o; }
HB_CMP(==) HB_CMP(!=) HB_CMP(<) HB_CMP(<=) HB_CMP(>) HB_CMP(>=)
#undef HB_CMP

// functions
template <class T> inline T sum(const Mat<T>& a) { return acc2(a.mem.data(), a.n_elem); }
template <class T> inline T sum(const subview_col<T>& a) { return acc2(a.p, a.n_elem); }
template <class T> inline T accu(const Mat<T>& a) { return acc2(a.mem.data(), a.n_elem); }
inline double mean(const mat& a) { return direct_mean(a.mem.data(), a.n_elem); }
inline double mean(const subview_elem<double>& s) { return mean(mat(s)); }
inline double var(const mat& a) { return direct_var(a.mem.data(), a.n_elem); }
inline double var(const subview_col<double>& a) { return direct_var(a.p, a.n_elem); }
inline double stddev(const mat& a) { return std::sqrt(direct_var(a.mem.data(), a.n_elem)); }
inline mat mean(const mat& a, int dim) {   // dim 1: mean of every row (op_mean: running sums over the columns)
  if (dim != 1) throw std::logic_error("mini_arma: mean(M, dim) is only provided for dim = 1");
  mat o(a.n_rows, 1);
  for (uword j = 0; j < a.n_cols; ++j) for (uword i = 0; i < a.n_rows; ++i) o.mem[i] += a(i, j);
  for (uword i = 0; i < a.n_rows; ++i) o.mem[i] /= double(a.n_cols);
  return o;
}
inline mat stddev(const mat& a, int norm_type, int dim) {
  if (dim != 1) throw std::logic_error("mini_arma: stddev(M, n, dim) is only provided for dim = 1");
  mat o(a.n_rows, 1);
  std::vector<double> tmp(a.n_cols);
  for (uword i = 0; i < a.n_rows; ++i) { for (uword j = 0; j < a.n_cols; ++j) tmp[j] = a(i, j); o.mem[i] = std::sqrt(direct_var(tmp.data(), a.n_cols, norm_type)); }
  return o;
}
inline double dot(const mat& a, const mat& b) { a.chk(b); return direct_dot(a.mem.data(), b.mem.data(), a.n_elem); }
inline double dot(const subview_col<double>& a, const subview_col<double>& b) { return direct_dot(a.p, b.p, a.n_elem); }
inline double dot(const subview_col<double>& a, const mat& b) { return direct_dot(a.p, b.mem.data(), a.n_elem); }
inline double dot(const mat& a, const subview_col<double>& b) { return direct_dot(a.mem.data(), b.p, b.n_elem); }
inline double norm(const mat& a, int p) { if (p != 2) throw std::logic_error("mini_arma: norm p"); double s = 0.0; for (double v : a.mem) s += v * v; return std::sqrt(s); }
#define HB_FN(name, expr) inline mat name(const mat& a) { mat o(a); for (double& x : o.mem) x = (expr); return o; }
HB_FN(square, x * x) HB_FN(log, std::log(x)) HB_FN(exp, std::exp(x)) HB_FN(sqrt, std::sqrt(x)) HB_FN(abs, std::fabs(x))
#undef HB_FN
inline mat square(const subview_col<double>& a) { return square(mat(a)); }
template <class T> inline T max(const Mat<T>& a) { if (!a.n_elem) throw std::logic_error("max(): empty"); return *std::max_element(a.mem.begin(), a.mem.end()); }
template <class T> inline T min(const Mat<T>& a) { if (!a.n_elem) throw std::logic_error("min(): empty"); return *std::min_element(a.mem.begin(), a.mem.end()); }
inline uvec find(const Mat<uword>& c) { uvec o; for (uword i = 0; i < c.n_elem; ++i) if (c.mem[i]) o.mem.push_back(i); o.n_rows = o.n_elem = o.mem.size(); o.n_cols = 1; return o; }
inline uvec find_finite(const mat& a) { uvec o; for (uword i = 0; i < a.n_elem; ++i) if (std::isfinite(a.mem[i])) o.mem.push_back(i); o.n_rows = o.n_elem = o.mem.size(); o.n_cols = 1; return o; }
template <class T> inline Col<T> unique(const Mat<T>& a) { std::vector<T> v(a.mem); std::sort(v.begin(), v.end()); v.erase(std::unique(v.begin(), v.end()), v.end()); return Col<T>(v); }
template <class T> inline bool any(const Mat<T>& a) { for (const T& v : a.mem) if (v != T(0)) return true; return false; }
template <class T> inline bool any(const subview_elem<T>& s) { s.check(); for (uword i : s.idx) if (s.m->mem[i] != T(0)) return true; return false; }
template <class T> inline bool all(const Mat<T>& a) { for (const T& v : a.mem) if (v == T(0)) return false; return true; }
inline double as_scalar(const mat& a) { if (a.n_elem != 1) throw std::logic_error("as_scalar(): not 1 x 1"); return a.mem[0]; }
inline vec diagvec(const mat& a) { vec o(std::min(a.n_rows, a.n_cols)); for (uword i = 0; i < o.n_elem; ++i) o.mem[i] = a(i, i); return o; }
template <class V> struct conv_to {
  template <class T> static V from(const Mat<T>& a) { V o; o.init(a.n_rows, a.n_cols); for (uword i = 0; i < a.n_elem; ++i) o.mem[i] = static_cast<typename V::elem_type>(a.mem[i]); if (a.n_cols != 1 && a.n_rows == 1) { o.n_rows = a.n_elem; o.n_cols = 1; } return o; }
};

// ---- sparse (compressed columns, rows ascending inside a column) ------------------------------------------------
template <class T> struct SpMat {
  uword n_rows = 0, n_cols = 0, n_nonzero = 0;
  std::vector<uword> col_ptrs{0};
  std::vector<uword> row_indices;
  std::vector<T> values;
  SpMat() {}
  SpMat(uword r, uword c) { resize0(r, c); }
  // compressed columns as given (Armadillo's SpMat(rowind, colptr, values, n_rows, n_cols)); zeros are not stored
  SpMat(const Mat<uword>& rowind, const Mat<uword>& colptr, const Mat<T>& vals, uword r, uword c) {
    resize0(r, c);
    for (uword j = 0; j < c; ++j) {
      for (uword p = colptr.mem[j]; p < colptr.mem[j + 1]; ++p) if (vals.mem[p] != T(0)) { row_indices.push_back(rowind.mem[p]); values.push_back(vals.mem[p]); }
      col_ptrs[j + 1] = row_indices.size();
    }
    n_nonzero = values.size();
  }
  void resize0(uword r, uword c) { n_rows = r; n_cols = c; n_nonzero = 0; col_ptrs.assign(c + 1, 0); row_indices.clear(); values.clear(); }
  void resize(uword r, uword c) {
    if (n_nonzero) throw std::logic_error("mini_arma: SpMat::resize of a non-empty matrix is not provided");
    resize0(r, c);
  }
  void set(uword i, uword j, T v) {
    if (i >= n_rows || j >= n_cols) throw std::logic_error("mini_arma: SpMat index out of range");
    const uword a = col_ptrs[j], b = col_ptrs[j + 1];
    const uword pos = std::lower_bound(row_indices.begin() + a, row_indices.begin() + b, i) - row_indices.begin();
    if (pos < b && row_indices[pos] == i) {
      if (v != T(0)) { values[pos] = v; return; }
      row_indices.erase(row_indices.begin() + pos); values.erase(values.begin() + pos);
      for (uword c = j + 1; c <= n_cols; ++c) --col_ptrs[c];
      --n_nonzero;
      return;
    }
    if (v == T(0)) return;
    row_indices.insert(row_indices.begin() + pos, i); values.insert(values.begin() + pos, v);
    for (uword c = j + 1; c <= n_cols; ++c) ++col_ptrs[c];
    ++n_nonzero;
  }
  T get(uword i, uword j) const {
    const uword a = col_ptrs[j], b = col_ptrs[j + 1];
    const uword pos = std::lower_bound(row_indices.begin() + a, row_indices.begin() + b, i) - row_indices.begin();
    return (pos < b && row_indices[pos] == i) ? values[pos] : T(0);
  }
  struct elem_proxy {
    SpMat* m; uword i, j;
    operator T() const { return m->get(i, j); }
    elem_proxy& operator=(T v) { m->set(i, j, v); return *this; }
    elem_proxy& operator=(const elem_proxy& o) { m->set(i, j, o.m->get(o.i, o.j)); return *this; }   // A(j, i) = A(i, j) = v
    elem_proxy& operator+=(T v) { m->set(i, j, m->get(i, j) + v); return *this; }
  };
  elem_proxy operator()(uword i, uword j) { return elem_proxy{this, i, j}; }
  T operator()(uword i, uword j) const { return get(i, j); }
  struct const_iterator {
    const SpMat* m = nullptr; uword pos = 0;
    T operator*() const { return m->values[pos]; }
    uword row() const { return m->row_indices[pos]; }
    const_iterator& operator++() { ++pos; return *this; }
    const_iterator operator++(int) { const_iterator t = *this; ++pos; return t; }
    bool operator!=(const const_iterator& o) const { return pos != o.pos; }
    bool operator==(const const_iterator& o) const { return pos == o.pos; }
  };
  typedef const_iterator const_col_iterator;
  typedef const_iterator iterator;
  typedef const_iterator col_iterator;
  const_iterator begin_col(uword j) const { return const_iterator{this, col_ptrs[j]}; }
  const_iterator end_col(uword j) const { return const_iterator{this, col_ptrs[j + 1]}; }
  struct col_view { const SpMat* m; uword j; };
  col_view col(uword j) const { return col_view{this, j}; }
  SpMat t() const {
    SpMat o(n_cols, n_rows);
    std::vector<uword> cnt(n_rows + 1, 0);
    for (uword r : row_indices) ++cnt[r + 1];
    for (uword r = 0; r < n_rows; ++r) cnt[r + 1] += cnt[r];
    o.col_ptrs = cnt; o.row_indices.resize(n_nonzero); o.values.resize(n_nonzero); o.n_nonzero = n_nonzero;
    std::vector<uword> fillp(cnt.begin(), cnt.end() - 1);
    for (uword j = 0; j < n_cols; ++j) for (uword p = col_ptrs[j]; p < col_ptrs[j + 1]; ++p) { const uword q = fillp[row_indices[p]]++; o.row_indices[q] = j; o.values[q] = values[p]; }
    return o;
  }
  static SpMat from_csc(uword r, uword c, const long long* colptr, const int* rowidx, const T* val) {
    SpMat o(r, c);
    for (uword j = 0; j < c; ++j) {
      std::vector<std::pair<uword, T>> e;
      for (long long p = colptr[j]; p < colptr[j + 1]; ++p) if (val[p] != T(0)) e.push_back({(uword)rowidx[p], val[p]});
      std::sort(e.begin(), e.end(), [](const std::pair<uword, T>& a, const std::pair<uword, T>& b) { return a.first < b.first; });
      for (auto& x : e) { o.row_indices.push_back(x.first); o.values.push_back(x.second); }
      o.col_ptrs[j + 1] = o.row_indices.size();
    }
    o.n_nonzero = o.values.size();
    return o;
  }
  SpMat& operator+=(const SpMat& b) { *this = *this + b; return *this; }
};
typedef SpMat<double> sp_mat;

template <class T> inline SpMat<T> operator+(const SpMat<T>& a, const SpMat<T>& b) {
  if (a.n_rows != b.n_rows || a.n_cols != b.n_cols) throw std::logic_error("mini_arma: sparse addition of different sizes");
  SpMat<T> o(a.n_rows, a.n_cols);
  for (uword j = 0; j < a.n_cols; ++j) {
    uword p = a.col_ptrs[j], q = b.col_ptrs[j];
    const uword pe = a.col_ptrs[j + 1], qe = b.col_ptrs[j + 1];
    while (p < pe || q < qe) {
      uword r; T v;
      if (q >= qe || (p < pe && a.row_indices[p] < b.row_indices[q])) { r = a.row_indices[p]; v = a.values[p]; ++p; }
      else if (p >= pe || b.row_indices[q] < a.row_indices[p]) { r = b.row_indices[q]; v = b.values[q]; ++q; }
      else { r = a.row_indices[p]; v = a.values[p] + b.values[q]; ++p; ++q; }
      if (v != T(0)) { o.row_indices.push_back(r); o.values.push_back(v); }
    }
    o.col_ptrs[j + 1] = o.row_indices.size();
  }
  o.n_nonzero = o.values.size();
  return o;
}
template <class T> inline SpMat<T> operator*(const SpMat<T>& a, T s) {
  SpMat<T> o(a.n_rows, a.n_cols);
  for (uword j = 0; j < a.n_cols; ++j) {
    for (uword p = a.col_ptrs[j]; p < a.col_ptrs[j + 1]; ++p) { const T v = a.values[p] * s; if (v != T(0)) { o.row_indices.push_back(a.row_indices[p]); o.values.push_back(v); } }
    o.col_ptrs[j + 1] = o.row_indices.size();
  }
  o.n_nonzero = o.values.size();
  return o;
}
template <class T> inline SpMat<T> operator*(T s, const SpMat<T>& a) { return a * s; }
// sparse x sparse: column j of the product = sum over the stored entries of b's column j, in order
template <class T> inline SpMat<T> operator*(const SpMat<T>& a, const SpMat<T>& b) {
  if (a.n_cols != b.n_rows) throw std::logic_error("mini_arma: sparse product of incompatible sizes");
  SpMat<T> o(a.n_rows, b.n_cols);
  std::vector<T> acc(a.n_rows, T(0));
  std::vector<char> hit(a.n_rows, 0);
  for (uword j = 0; j < b.n_cols; ++j) {
    std::vector<uword> rows;
    for (uword q = b.col_ptrs[j]; q < b.col_ptrs[j + 1]; ++q) {
      const uword k = b.row_indices[q];
      for (uword p = a.col_ptrs[k]; p < a.col_ptrs[k + 1]; ++p) { const uword r = a.row_indices[p]; if (!hit[r]) { hit[r] = 1; rows.push_back(r); } acc[r] += a.values[p] * b.values[q]; }
    }
    std::sort(rows.begin(), rows.end());
    for (uword r : rows) { if (acc[r] != T(0)) { o.row_indices.push_back(r); o.values.push_back(acc[r]); } acc[r] = T(0); hit[r] = 0; }
    o.col_ptrs[j + 1] = o.row_indices.size();
  }
  o.n_nonzero = o.values.size();
  return o;
}
// sparse x dense vector / matrix: out(row, c) += value * x(col, c), stored entries in column order
inline mat operator*(const sp_mat& a, const mat& x) {
  if (a.n_cols != x.n_rows) throw std::logic_error("mini_arma: sparse x dense of incompatible sizes");
  mat o(a.n_rows, x.n_cols);
  o.is_col = x.n_cols == 1;
  for (uword c = 0; c < x.n_cols; ++c)
    for (uword j = 0; j < a.n_cols; ++j) { const double xj = x(j, c); for (uword p = a.col_ptrs[j]; p < a.col_ptrs[j + 1]; ++p) o(a.row_indices[p], c) += a.values[p] * xj; }
  return o;
}
inline mat operator*(const sp_mat& a, const subview_col<double>& x) { return a * mat(x); }
// dense (row vector or matrix) x sparse: out(r, j) = sum over the stored entries of column j
inline mat operator*(const mat& x, const sp_mat& a) {
  if (x.n_cols != a.n_rows) throw std::logic_error("mini_arma: dense x sparse of incompatible sizes");
  mat o(x.n_rows, a.n_cols);
  for (uword j = 0; j < a.n_cols; ++j)
    for (uword r = 0; r < x.n_rows; ++r) { double s = 0.0; for (uword p = a.col_ptrs[j]; p < a.col_ptrs[j + 1]; ++p) s += x(r, a.row_indices[p]) * a.values[p]; o(r, j) = s; }
  return o;
}
inline double dot(const sp_mat::col_view& c, const mat& x) {
  double s = 0.0;
  for (uword p = c.m->col_ptrs[c.j]; p < c.m->col_ptrs[c.j + 1]; ++p) s += c.m->values[p] * x.mem[c.m->row_indices[p]];
  return s;
}
inline vec diagvec(const sp_mat& a) { vec o(std::min(a.n_rows, a.n_cols)); for (uword i = 0; i < o.n_elem; ++i) o.mem[i] = a.get(i, i); return o; }
}  // namespace arma
#endif
