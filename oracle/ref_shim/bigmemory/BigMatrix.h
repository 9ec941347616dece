/* stand-in for bigmemory's BigMatrix (TEST INFRASTRUCTURE ONLY; see ../RcppArmadillo.h): a column-major matrix of
 * char / short / int / float / double in caller-owned memory, as much of the class as tXXmat.cpp and read_bed.cpp use. */
#ifndef HB_SHIM_BIGMATRIX_H
#define HB_SHIM_BIGMATRIX_H
#include <climits>
#include <cstddef>
#include "../RcppArmadillo.h"
typedef long index_type;
class BigMatrix {
  void* data_; index_type nrow_, ncol_; int type_;
 public:
  BigMatrix(void* data, index_type nrow, index_type ncol, int type) : data_(data), nrow_(nrow), ncol_(ncol), type_(type) {}
  index_type nrow() const { return nrow_; }
  index_type ncol() const { return ncol_; }
  int matrix_type() const { return type_; }   // 1 char, 2 short, 4 int, 6 float, 8 double
  void* matrix() { return data_; }
};
#ifndef NA_CHAR
#define NA_CHAR CHAR_MIN
#define NA_SHORT SHRT_MIN
#endif
#ifndef NA_INTEGER
#define NA_INTEGER INT_MIN
#define NA_REAL (__builtin_nan("1954"))
#endif
#endif
