/* stand-in for RcppProgress (TEST INFRASTRUCTURE ONLY): never aborts, displays nothing */
#ifndef HB_SHIM_PROGRESS_H
#define HB_SHIM_PROGRESS_H
#include "progress_bar.hpp"
class Progress {
 public:
  Progress(unsigned long, bool, ProgressBar&) {}
  Progress(unsigned long, bool) {}
  bool increment(unsigned long = 1) { return true; }
  static bool check_abort() { return false; }
};
#endif
