/* stand-in for bigmemory's MatrixAccessor (TEST INFRASTRUCTURE ONLY): accessor[column][row] on a BigMatrix */
#ifndef HB_SHIM_MATRIXACCESSOR_H
#define HB_SHIM_MATRIXACCESSOR_H
#include "BigMatrix.h"
template <typename T> class MatrixAccessor {
  T* p_; index_type nrow_;
 public:
  explicit MatrixAccessor(BigMatrix& bm) : p_(static_cast<T*>(bm.matrix())), nrow_(bm.nrow()) {}
  T* operator[](index_type col) { return p_ + col * nrow_; }
};
#endif
